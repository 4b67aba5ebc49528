#!/usr/bin/env python
"""Summarise ncu outputs into the tracked text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep [regex]    > profiles/rNN_<kernel>.md
    python tools/ncu_summary.py traffic  a.ncu-rep b.ncu-rep ...            > profiles/rNN_traffic.json
    python tools/ncu_summary.py ranges   gpurun_out/ranges.csv name1,name2.. N  > profiles/rNN_traffic.json
      (ncu --replay-mode range of `bench.py --profile-ranges`: one range per kernel family, N launches each)

`launches` reads the CSV written by `ncu --metrics gpu__time_duration.sum --csv --log-file ...`;
`full` reads a `--set full` report through `ncu -i ... --page raw/source --csv` (no GPU needed).
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("dycon::<unnamed>::", "").replace("(anonymous namespace)::", "")
    return name[-70:]


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if "spin_kernel" in r[ki]:        # torch.cuda._sleep: bench.py's pacing kernel, not part of the step
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        agg.setdefault(short(r[ki]), []).append(v)
    total = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {sum(v):.1f} | {100 * sum(v) / total:.1f}% |")
    print(f"\nsum of profiled launches: {total:.1f} us (cold-cache, serialised by ncu: compare shares, not absolutes)")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, regex=None):
    rows = ncu_csv(rep, "raw")
    h = rows[0]
    ki = h.index("Kernel Name")
    idx = [(m, h.index(m)) for m in RAW if m in h]
    units = rows[1]
    print("## raw metrics (one line per captured launch)\n")
    for r in rows[2:]:
        if regex and not re.search(regex, r[ki]):
            continue
        print(f"### `{short(r[ki])}`")
        for m, i in idx:
            print(f"- {m} = {r[i]} {units[i]}")
        print()
    src = ncu_csv(rep, "source")
    secs, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    print("## warp-stall sampling (source page)\n")
    for sec in secs:
        if sec["name"] in seen or (regex and not re.search(regex, sec["name"])) or not sec["rows"]:
            continue
        seen.add(sec["name"])
        h, data = sec["rows"][0], sec["rows"][1:]
        si = h.index("# Samples") if "# Samples" in h else h.index("Warp Stall Sampling (All Samples)")
        so = h.index("Source")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
        agg = {h[i]: sum(int(r[i]) for r in data if len(r) > i and r[i].isdigit()) for i in stall}
        print(f"### `{short(sec['name'])}` -- {tot} samples")
        print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.0f}%" for k, v in
                                            sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        print("\n| samples | SASS | top stalls |\n|---:|---|---|")
        for r in sorted((r for r in data if len(r) > si and r[si].isdigit()), key=lambda r: -int(r[si]))[:12]:
            st = sorted(((h[i][6:], int(r[i])) for i in stall if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
            print(f"| {r[si]} | `{r[so].strip()[:70]}` | {', '.join(f'{k} {v}' for k, v in st)} |")
        print()


def traffic(reps):
    """DRAM bytes read / written and duration of the first captured launch of every kernel (JSON)."""
    import json
    out = collections.OrderedDict()
    for rep in reps:
        rows = ncu_csv(rep, "raw")
        h, units = rows[0], rows[1]
        ki = h.index("Kernel Name")
        col = {m: h.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0,
                 "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
        val = lambda r, m: float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
        for r in rows[2:]:
            name = re.sub(r"^.*>::", "", short(r[ki]))
            if name not in out:
                out[name] = {"dram_read_bytes": val(r, "dram__bytes_read.sum"),
                             "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                             "duration_us": round(val(r, "gpu__time_duration.sum"), 3)}
    print(json.dumps({"source": "ncu --set full, B200, BraTS19 shape B=4 N=1728 D=256 fp16, one launch per kernel, caches "
                                "flushed by ncu between replays (see the step_kernels / ema summaries next to this file)",
                      "kernels": out}, indent=1))


def ranges(path, names, per_range):
    """CSV of `ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
    -> {"families": {name: {dram_read_bytes, dram_write_bytes, dram_bytes_per_launch, us_per_launch}}}."""
    import json
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5 and r[0].isdigit()]
    out = {}
    for r in rows:
        rid, metric, unit, val = int(r[0]), r[-3], r[-2], float(r[-1].replace(",", ""))
        if rid >= len(names):
            continue
        d = out.setdefault(names[rid], {})
        if metric == "dram__bytes_read.sum":
            d["dram_read_bytes"] = val / per_range
        elif metric == "dram__bytes_write.sum":
            d["dram_write_bytes"] = val / per_range
        elif metric == "gpu__time_duration.sum":
            d["us_per_launch_under_ncu"] = val / 1e3 / per_range
    for d in out.values():
        d["dram_bytes_per_launch"] = d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
    print(json.dumps({"how": "ncu --replay-mode range over bench.py --profile-ranges: every range = %d back-to-back launches of one "
                             "kernel family on 4 rotating input sets (452 MB > 126 MB L2); bytes are per launch. Outputs that are "
                             "rewritten in place every launch (the 28 MB UnCL gradient, the 7 MB FeCL gradient) stay in the 126 MB "
                             "L2 and show up as fewer DRAM writes than the algorithmic bytes." % per_range,
                      "families": out}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "ranges":
        ranges(sys.argv[2], sys.argv[3].split(","), int(sys.argv[4]))
    elif sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
