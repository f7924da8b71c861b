"""CPU oracle for the SkyEye detector forward path (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch-fp32 / plain-C restatement of the reference's
algorithm for the hot path (backbone -> neck -> CLA -> transformer heads ->
decode -> NMS).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package (``skyeye``) never imports anything from here and fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against the reference itself, executed in the build
container (``tests/golden/make_golden.py`` imports ``/root/reference`` with the
repairs R1-R4 of SURVEY.md §0.2, loads the oracle's state dict into the
reference modules and stores the reference outputs under ``tests/golden``).
"""
