"""GPU parity of the whole hot path through the C-ABI: cublas_dprimme / dprimme of the product
against the committed reference fixture and the invariants of the reference's check_solution."""
import numpy as np
import pytest

import harness as H
import solver_checks as SC
from golden.cases import CASES

pytestmark = pytest.mark.gpu


# Cases whose outer-iteration / restart / matvec counts on the GPU are NOT identical to the reference's although
# the fixture marks them exact for the CPU kernels (different summation order in the panels): they must still be
# within 5 %.  Every other `exact_counts` case is asserted IDENTICAL on the GPU.
GPU_CLOSE = set()


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_matches_reference_fixture(name):
    r = SC.run_case("product", name)
    got, want = SC.check_against_golden(name, r, counts="close" if name in GPU_CLOSE else "exact")
    assert r["launches"] > 0  # the CUDA path did the work
    print(name, "counts", got, "reference", want, "IDENTICAL" if got == want else "deviation")


def test_c2_full_size_properties():
    """config C2 at full size (n = 10^6): analytic eigenvalues of the 3-D Laplacian,
    orthonormality and residuals (no CPU oracle at this size)."""
    from primme_b200 import api, matrices as M
    N = 100
    csr = M.laplacian_nd((N, N, N))
    r = H.solve("product", csr, 10, method=api.PRIMME_GD_Olsen_plusK, maxBlockSize=4, maxBasisSize=40,
                eps=1e-10, aNorm=12.0)
    assert r["ret"] == 0
    lam1 = 4 * np.sin(np.pi * np.arange(1, 4) / (2 * (N + 1))) ** 2
    exact = np.sort([lam1[i] + lam1[j] + lam1[k] for i in range(3) for j in range(3) for k in range(3)])[:10]
    assert np.allclose(r["evals"], exact, rtol=1e-10, atol=0)
    SC.check_invariants(csr, r, 1e-10, 12.0)
    s = r["stats"]
    # reference on this container: 1045 outer / 195 restarts / 3919 matvecs
    assert abs(s["numMatvecs"] - 3919) < 0.1 * 3919, s


def test_host_contract_dprimme():
    """dprimme with HOST evecs and a HOST matvec callback (reference dprimme contract)"""
    import ctypes as C
    from primme_b200 import api, matrices as M
    lib = H.lib_product()
    ok = H.lib_oracle_kernels()  # only its host CSR callback is used (test infrastructure)
    ip, ix, da = M.laplacian_nd((13, 17, 19))
    n = len(ip) - 1
    A = H.CsrHost(n, ip.ctypes.data, ix.ctypes.data, da.ctypes.data, 1, None, 0.0, 0)
    p = api.new_params(lib, n, numEvals=8, maxBlockSize=4, maxBasisSize=40, eps=1e-10)
    p.matrix = C.addressof(A)
    p.matrixMatvec = C.cast(ok.csr_host_matvec, C.c_void_p).value
    assert lib.primme_set_method(api.PRIMME_GD_Olsen_plusK, C.byref(p)) == 0
    evals, rn, evecs = np.zeros(8), np.zeros(8), np.zeros((8, n))
    rc = lib.dprimme(evals.ctypes.data, evecs.ctypes.data, rn.ctypes.data, C.byref(p))
    assert rc == 0
    g = SC.GOLDEN["aniso_b4_smallest"]
    assert np.allclose(evals, g["evals"], rtol=1e-10)
    assert p.stats.numMatvecs == g["numMatvecs"] or abs(p.stats.numMatvecs - g["numMatvecs"]) < 0.05 * g["numMatvecs"]


@pytest.mark.parametrize("kw", [dict(target="closest_abs", targetShifts=[1.3]),
                                dict(target="closest_geq", targetShifts=[0.5, 1.0, 1.3], locking=1),
                                dict(target="closest_abs", targetShifts=[1.3], maxBlockSize=3),
                                dict(target="closest_abs", targetShifts=[1.3], projection=2),
                                dict(target="closest_geq", targetShifts=[1.3], locking=1, maxBlockSize=2, projection=2)])
def test_refined_extraction_product_matches_reference(kw):
    """primme_proj_refined (and, with projection=2, primme_proj_harmonic) on the GPU (Q next to V and W: residual utility, block-ortho sweep with R
    factor, VWXR sweep for Q*hU) against the unmodified reference: same eigenpairs, residuals below the
    tolerance, counts within the rounding drift of interior targets"""
    from primme_b200 import api, matrices as M
    kw = dict(kw)
    kw["target"] = getattr(api, "primme_" + kw["target"])
    csr = M.laplacian_nd((7, 11, 13))
    k = 3
    proj = kw.pop("projection", api.primme_proj_refined)
    ref = H.solve("reference", csr, k, projection=proj, eps=1e-8, **kw)
    got = H.solve("product", csr, k, projection=proj, eps=1e-8, **kw)
    assert ref["ret"] == 0 and got["ret"] == 0 and got["initSize"] == k and got["launches"] > 0
    # both must return eigenvalues of A on the right side of the shift (with clustered interior values a
    # Davidson run may skip one: the sets are compared to the spectrum, pairwise only on the first two)
    n = len(csr[0]) - 1
    exact = np.linalg.eigvalsh(M.csr_matvec(*csr, np.eye(n)))
    for r in (ref, got):
        assert all(np.abs(exact - e).min() < 1e-6 for e in r["evals"])
        if kw["target"] == api.primme_closest_geq:
            assert np.all(r["evals"] >= min(kw["targetShifts"]) - 1e-6)
    assert np.abs(np.sort(got["evals"])[:2] - np.sort(ref["evals"])[:2]).max() <= 1e-6
    X = got["evecs"]
    res = np.linalg.norm(M.csr_matvec(*csr, X) - X * got["evals"], axis=0)
    assert res.max() < 1e-8 * 12 * 1.1
    assert np.abs(X.T @ X - np.eye(k)).max() < 1e-7
    assert abs(got["stats"]["numOuterIterations"] - ref["stats"]["numOuterIterations"]) <= 0.35 * ref["stats"]["numOuterIterations"]


@pytest.mark.parametrize("method,k", [("PRIMME_LOBPCG_OrthoBasis", 12), ("PRIMME_STEEPEST_DESCENT", 11)])
def test_block_sizes_above_the_panel_width(method, k):
    """the reference's LOBPCG / steepest-descent presets set maxBlockSize = numEvals (primme_interface.c:455,465):
    blocks wider than the kernels' 8-column panels run in chunks of 8"""
    from primme_b200 import api, matrices as M
    csr = M.laplacian_nd((40, 31))
    kw = dict(method=getattr(api, method), eps=1e-9, aNorm=8.0)
    got = H.solve("product", csr, k, **kw)
    ref = H.solve("reference", csr, k, **kw)
    assert got["ret"] == 0 and ref["ret"] == 0
    assert got["params"].maxBlockSize == k > 8
    assert np.allclose(got["evals"], ref["evals"], rtol=1e-10)
    SC.check_invariants(csr, got, 1e-9, 8.0)
    assert got["launches"] > 0
