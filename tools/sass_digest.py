#!/usr/bin/env python
"""profiles/sass_digest.txt: what the shipped library is made of (cuobjdump -sass | grep -c per mnemonic and kernel).

    python tools/sass_digest.py > profiles/sass_digest.txt

UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDGSTS = cp.async (B200_PROFILING.md, "What proves a Blackwell-native kernel").  HMMA (legacy
mma.sync) must be absent."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dycon_paper_replication_b200", "_dycon_b200.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "BRA.U.ANY", "LDGSTS", "MUFU", "HMMA", "RED.E", "ATOM"]


def short_name(pretty):
    """kernel name with its template arguments, without the parameter list and the namespaces"""
    s = pretty.replace("dycon::(anonymous namespace)::", "").replace("void ", "")
    depth, out = 0, []
    for ch in s:                       # cut at the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out).strip()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn.endswith(".") and mn in line):
                per[cur][mn] += 1
        if "/*" in line and ";" in line:
            per[cur]["instructions"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# SASS digest of {os.path.relpath(SO, ROOT)} ({os.path.getsize(SO)} bytes; sm_100a)")
    print("# totals: " + ", ".join(f"{mn} {total[mn]}" for mn in MNEMONICS) + f", instructions {total['instructions']}")
    print("# kernels using tcgen05 / TMA (counts per kernel):")
    for name, pretty in zip(per, demangle):
        c = per[name]
        if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
            print(f"{short_name(pretty)}: " + ", ".join(f"{mn} {c[mn]}" for mn in MNEMONICS if c[mn]) + f", instructions {c['instructions']}")
    print("# other kernels:")
    for name, pretty in zip(per, demangle):
        c = per[name]
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]):
            print(f"{short_name(pretty)}: " + ", ".join(f"{mn} {c[mn]}" for mn in MNEMONICS if c[mn]) + f", instructions {c['instructions']}")


if __name__ == "__main__":
    main()
