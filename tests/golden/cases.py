"""Solver parity cases shared by the golden generator and the tests.  Each case: name ->
(matrix spec, numEvals, keyword arguments of harness.solve)."""
from primme_b200 import api as G, matrices as M

MATRICES = {
    "aniso3d": lambda: M.laplacian_nd((13, 17, 19)),
    "lap3d_20": lambda: M.laplacian_nd((20, 20, 20)),     # config C2 at N=20 (degenerate spectrum)
    "lap1d_500": lambda: M.laplacian_1d(500),
    "lap2d": lambda: M.laplacian_nd((30, 41)),
    "lap1d_100": lambda: M.laplacian_1d(100),              # matrix of examples/ex_eigs_dseq.c
    "powerlaw_4k": lambda: M.power_law_symmetric(4000, mean_degree=10.0, seed=11),
}

# exact_counts: outer iterations / restarts / matvecs must equal the reference's (non-degenerate
# spectra, where the control flow is not at the mercy of rounding in the summation order)
CASES = {
    "c2_small": ("lap3d_20", 10, dict(method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=4, maxBasisSize=40, eps=1e-10, aNorm=12.0), False),
    "aniso_b4_smallest": ("aniso3d", 8, dict(method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=4, maxBasisSize=40, eps=1e-10), True),
    "aniso_b4_largest": ("aniso3d", 8, dict(target=G.primme_largest, method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=4, maxBasisSize=40, eps=1e-10), True),
    "aniso_b1_cgs": ("aniso3d", 5, dict(method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=1, eps=1e-10), True),
    "aniso_b8_m64_largest_k20": ("aniso3d", 20, dict(target=G.primme_largest, method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=8, maxBasisSize=64, eps=1e-8), True),
    "lap1d_b2_gdk": ("lap1d_500", 4, dict(method=G.PRIMME_GD_plusK, maxBlockSize=2, eps=1e-9), True),
    "lap2d_b3_locking": ("lap2d", 6, dict(method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=3, locking=1, eps=1e-10), True),
    "lap2d_b1_locking": ("lap2d", 6, dict(method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=1, locking=1, eps=1e-10), True),
    "lap2d_b1_jacobi": ("lap2d", 4, dict(method=G.PRIMME_GD_Olsen_plusK, jacobi=True, eps=1e-10), True),
    "lap2d_b4_jacobi_jdolsen": ("lap2d", 4, dict(method=G.PRIMME_JD_Olsen_plusK, maxBlockSize=4, jacobi=True, eps=1e-10), True),
    "lap2d_closest_abs_b2": ("lap2d", 4, dict(target=G.primme_closest_abs, targetShifts=[3.1], method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=2, eps=1e-9), False),
    "lap2d_arnoldi": ("lap2d", 2, dict(method=G.PRIMME_Arnoldi, eps=1e-7), True),
    "lap2d_lobpcg": ("lap2d", 3, dict(method=G.PRIMME_LOBPCG_OrthoBasis, eps=1e-8), True),
    "ex_eigs_dseq_gdk": ("lap1d_100", 10, dict(method=G.PRIMME_GD_Olsen_plusK, eps=1e-9, jacobi=True), True),
    "powerlaw_b8_largest": ("powerlaw_4k", 12, dict(target=G.primme_largest, method=G.PRIMME_GD_Olsen_plusK, maxBlockSize=8, maxBasisSize=64, eps=1e-8), True),
}
