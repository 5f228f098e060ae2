"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the cfm_b200 hot path.

Contents
  oracle/ot/            stand-in for the un-vendored POT package (solver bodies)
  oracle/coupling.py    restatement of OTPlanSampler's glue + matcher formulas
  oracle/vector_field.py  reference MLP, torch_wrapper, torchdyn-style dopri5

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and
--impl reference).  Nothing in cfm_b200/ may import from here; the product path
fails loudly when the CUDA library is missing instead of routing through CPU code.

Parity status: exact-OT pinned (scipy LSA + reference tests + golden fixtures);
Sinkhorn plan values, MLP-through-dopri5 trajectories: PARITY UNPINNED at the
POT / torchdyn boundary (no golden vectors exist in the reference; both packages
are absent from /root/reference and from this image).
"""
