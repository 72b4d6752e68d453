"""SVD front end on the GPU (SURVEY 8f rank 2, config C4 family): dprimme_svds with the reference's
host contract against the unmodified reference, and cublas_dprimme_svds with the built-in CSR
operator (A and A' SpMM on device blocks) against a dense SVD."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import svds_harness as S
import test_svds_cpu as T
from primme_b200 import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["tall_largest", "wide_largest", "tall_largest_block", "tall_locking"])
def test_svds_product_host_contract_matches_reference(case):
    m, n, per_row, seed, k, target, kw = T.CASES[case]
    csr = S.random_rect(m, n, per_row, seed)
    ref = S.solve("reference", csr, (m, n), k, target=target, method_stage1=api.PRIMME_GD_Olsen_plusK, **kw)
    got = S.solve("product", csr, (m, n), k, target=target, method_stage1=api.PRIMME_GD_Olsen_plusK, **kw)
    T.check(case, got)
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-10)
    # same decisions up to rounding: counts within 3 % of the reference's
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= max(3, 0.03 * ref["stats"][key]), (got["stats"], ref["stats"])


@pytest.mark.parametrize("case", ["hybrid_loose", "hybrid_tight", "hybrid_wide", "hybrid_block", "augmented"])
def test_svds_two_stage_product_host_contract_matches_reference(case):
    m, n, per_row, seed, k, preset, kw = T.TWO_STAGE[case]
    csr = S.random_rect(m, n, per_row, seed)
    args = dict(method=preset, method_stage1=api.PRIMME_GD_Olsen_plusK, method_stage2=api.PRIMME_GD_Olsen_plusK, **kw)
    ref = S.solve("reference", csr, (m, n), k, **args)
    got = S.solve("product", csr, (m, n), k, **args)
    T.check_two_stage(case, got)
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-11)
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= max(3, 0.05 * ref["stats"][key]), (got["stats"], ref["stats"])


@pytest.mark.parametrize("shape,k,preset", [((40000, 9000), 6, S.primme_svds_normalequations),
                                            ((7000, 30000), 5, S.primme_svds_normalequations),
                                            ((12000, 3000), 4, S.primme_svds_hybrid),
                                            ((3000, 9000), 3, S.primme_svds_augmented)])
def test_cublas_dprimme_svds_builtin_operator(shape, k, preset):
    m, n = shape
    csr = S.random_rect(m, n, 6, 11)
    lib = H.lib_product()
    S.declare(lib)
    rp = np.ascontiguousarray(csr[0], dtype=np.int64)
    ci = np.ascontiguousarray(csr[1], dtype=np.int32)
    va = np.ascontiguousarray(csr[2], dtype=np.float64)
    ctx = C.c_void_p()
    assert lib.pb200_ctx_create(C.byref(ctx), -1) == 0
    A = C.c_void_p()
    assert lib.pb200_csr_create(ctx, m, n, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
    lib.pb200_csr_build_transpose.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.pb200_csr_build_transpose(ctx, A) == 0
    p = lib.primme_svds_params_create()
    for name, v in (("m", m), ("n", n), ("numSvals", k), ("target", S.primme_svds_largest), ("printLevel", 0),
                    ("maxBlockSize", 2), ("matrix", A.value),
                    ("matrixMatvec", C.cast(lib.primme_b200_svds_csr_matvec, C.c_void_p).value)):
        S.set_member(lib, p, name, v)
    S.set_member(lib, p, "eps", 1e-9 if preset == S.primme_svds_normalequations else 1e-12)
    assert lib.primme_svds_set_method(preset, api.PRIMME_GD_Olsen_plusK, api.PRIMME_GD_Olsen_plusK, p) == 0
    inner = S.get_member(lib, p, "primme")  # address of the first-stage primme_params
    inner_p = C.cast(C.c_void_p(inner), C.POINTER(api.PrimmeParams))
    lib.primme_b200_attach_ctx(inner_p, ctx)
    dsvecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * (m + n) * k, C.byref(dsvecs)) == 0
    svals, rn = np.zeros(k), np.zeros(k)
    l0 = lib.pb200_ctx_launches(ctx)
    rc = lib.cublas_dprimme_svds(svals.ctypes.data, dsvecs, rn.ctypes.data, p)
    assert rc == 0 and S.get_member(lib, p, "initSize") == k
    assert lib.pb200_ctx_launches(ctx) > l0
    host = np.zeros((m + n) * k)
    assert lib.pb200_copy_d2h(ctx, dsvecs, (m + n) * k, host.ctypes.data, (m + n) * k, (m + n) * k, 1, 8) == 0
    U = host[: m * k].reshape(k, m).T
    V = host[m * k:].reshape(k, n).T
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    As = sp.csr_matrix((va, ci, rp), shape=(m, n))
    want = np.sort(spl.svds(As, k=k, which="LM", tol=1e-12, return_singular_vectors=False))[::-1]
    assert np.allclose(np.sort(svals)[::-1], want, rtol=1e-8)
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-8 and np.abs(U.T @ U - np.eye(k)).max() < 1e-6
    assert np.linalg.norm(As @ V - U * svals, axis=0).max() < 1e-7 * want[0]
    if preset != S.primme_svds_normalequations:
        # the second stage reaches what the normal equations cannot: both residuals at eps |A|
        assert np.linalg.norm(As @ V - U * svals, axis=0).max() < 1e-11 * want[0]
        assert np.linalg.norm(As.T @ U - V * svals, axis=0).max() < 1e-11 * want[0]
    lib.primme_b200_attach_ctx(inner_p, None)
    lib.pb200_free(ctx, dsvecs)
    lib.primme_svds_params_destroy(p)
    lib.pb200_csr_destroy(ctx, A)
    lib.pb200_ctx_destroy(ctx)
