"""CPU oracle for the SIA2D hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``odinn.jl_b200/``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

PARITY UNPINNED BY STORED NUMBERS: the reference (ODINN.jl v1.1.0) ships no
golden vectors for ``SIA2D!`` or its VJPs, Julia is not available in the build
environment, and the forward kernel lives in the un-vendored Huginn.jl
(compat 0.13.3).  The oracle is pinned instead by (1) the reference's own
operator-transpose identities (test/SIA2D_adjoint_utils.jl), (2) the reference's
VJP-vs-finite-difference protocol and thresholds (test/SIA2D_adjoint.jl), with
the in-tree hand-written adjoint transcribed verbatim and checked against finite
differences of the restated forward, and (3) the Halfar similarity solution.
"""
