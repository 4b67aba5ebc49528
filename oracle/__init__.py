"""Oracle for the DyCON loss hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement of the reference algorithm
(``code/utils/dycon_losses.py`` and ``update_ema_variables`` in
``code/train_DyCON_BraTS19.py`` of rogeliorjr/DyCON_Paper_Replication).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline.  The product package
(``dycon_paper_replication_b200``) never imports this package.

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 8c), so the oracle is pinned by *executing the unmodified
reference module* in the build container: ``oracle/make_golden.py`` imports
``/root/reference/code/utils/dycon_losses.py`` by path, runs it on seeded
inputs in fp32 and fp64 and commits inputs + outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` then checks both restatements in this package
against those fixtures (no access to ``/root/reference`` needed at test time).

Two independent restatements:

* ``oracle.torch_port`` -- operation-for-operation torch restatement that lets
  autograd produce the gradients (same cost profile as the reference; this is
  what ``bench.py`` times as the CPU baseline, kind "port").
* ``oracle.closed_form`` -- numpy float64 closed forms with hand-derived
  gradients; this is the specification the CUDA kernels
  implement.
"""
