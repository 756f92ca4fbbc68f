"""CPU oracle for the Qaintensor.jl hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / pure-Python restatement of the reference's
algorithms for the one hot path this repository accelerates (tensor-network
contraction along the default / network2graph order, SVD truncation, and the
sliced-contraction / MPS extensions built from those primitives).  Every
function cites the reference file:line it follows (paths are relative to the
reference checkout, e.g. ``src/contract.jl:242-264``).

Rules ("Oracle"):

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import this package;
* the product package ``qaintensor.jl_b200`` never imports it and has no CPU
  fallback -- it fails loudly when ``libqaintensor_cuda.so`` is missing;
* the reference (pure Julia, with its arithmetic in the un-vendored packages
  TensorOperations 3.1.0, LightGraphs 1.3.5, LinearAlgebra/LAPACK) cannot be
  built or imported in this environment (no ``julia`` binary), so there is no
  ``oracle/_ref``.  The oracle is pinned against every literal known-answer in
  the reference's own tests (``test/test_treewidth.jl``, ``test/test_helper.jl``,
  ``test/test_mpo.jl:81``), against the reference tests' cross-path relations
  with fixed seeds, and against independent dense linear algebra (DFT matrix,
  ``kron`` state vectors, ``numpy.linalg.svd``).  Numeric amplitude /
  singular-value vectors do not exist in the reference (its tests are unseeded
  and compare paths with ``≈``), and the reference cannot be executed here, so
  for those quantities the status is **parity unpinned** against the reference
  itself: they are pinned only by relation (dense ``U*psi``, DFT, LAPACK) and
  by the oracle-generated fixtures of ``tests/golden/``.  The exact contraction
  *order* for a given network is pinned by the literal subroutine known-answers
  plus the hand-traced vector of SURVEY.md Appendix B.

Index conventions: every public oracle function speaks the reference's
conventions -- 1-based tensor / leg / wire numbers, column-major reshapes
(``order="F"``), axis ``i`` of a Julia array is numpy axis ``i-1``.
"""
