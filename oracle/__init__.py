"""CPU oracle for the two DetectorFreeSfM hot paths -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain functional PyTorch (CPU, fp32) and plain C, the
algorithms of the reference's pairwise coarse matcher (third_party/LoFTR) and of its
multi-view refinement matcher (src/MultiviewMatcher + third_party/RoIAlign.pytorch).
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and only as the checker.
The product path (``detectorfreesfm_b200``) never imports this package and fails
loudly when the CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"):
  * RoIAlign: pinned against the reference's own known-answer vector
    (third_party/RoIAlign.pytorch/README.md:42-96) and against the reference's own
    C++ (``oracle/_ref``, compiled from the sources where they lie).
  * LoFTR / MultiviewMatcher: the reference holds no golden tensors for these; the
    restatement is pinned against the reference modules themselves, imported in the
    build container from /root/reference (tests/test_oracle_vs_reference.py), and the
    outputs of that import are committed as fixtures under tests/golden/ together
    with the generating script (tests/golden/make_golden.py).
"""
