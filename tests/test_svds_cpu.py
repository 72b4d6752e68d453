"""SVD front end, normal equations: the host logic (primme_b200/src/svds.c
linked against the CPU restatement of the kernels) against the UNMODIFIED reference's dprimme_svds
on the same matrices, callbacks and parameters -- singular values, residual norms, vectors and the
iteration / matvec counts."""
import numpy as np
import pytest

import harness as H
import svds_harness as S
from primme_b200 import api

CASES = {
    # name: (m, n, per_row, seed, numSvals, target, extra)
    "tall_largest": (600, 150, 6, 1, 6, S.primme_svds_largest, dict(eps=1e-10)),
    "wide_largest": (150, 500, 8, 2, 5, S.primme_svds_largest, dict(eps=1e-10)),
    "tall_largest_block": (500, 200, 5, 3, 8, S.primme_svds_largest, dict(eps=1e-9, maxBlockSize=4, maxBasisSize=32)),
    "tall_smallest": (300, 60, 5, 4, 3, S.primme_svds_smallest, dict(eps=1e-8)),
    "tall_locking": (400, 120, 6, 5, 5, S.primme_svds_largest, dict(eps=1e-10, locking=1)),
}


def check(case, r):
    m, n, per_row, seed, k, target, kw = CASES[case]
    csr = S.random_rect(m, n, per_row, seed)
    A = S.dense(csr, (m, n))
    sv = np.linalg.svd(A, compute_uv=False)
    want = sv[:k] if target == S.primme_svds_largest else sv[::-1][:k]
    assert r["ret"] == 0 and r["initSize"] == k
    got = np.sort(r["svals"])[::-1] if target == S.primme_svds_largest else np.sort(r["svals"])
    assert np.allclose(got, want, rtol=0, atol=10 * kw["eps"] * sv[0])
    U, V = r["U"], r["V"]
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-8
    # triplet residuals ||A v - sigma u|| and ||A' u - sigma v|| within the normal-equations accuracy
    R1 = A @ V - U * r["svals"]
    R2 = A.T @ U - V * r["svals"]
    assert np.linalg.norm(R1, axis=0).max() < 1e-8 * sv[0]
    assert np.all(np.linalg.norm(R2, axis=0) <= np.maximum(r["rnorms"] * 10, 1e-9 * sv[0]))


@pytest.mark.parametrize("case", sorted(CASES))
def test_svds_hostcheck_matches_reference(case):
    m, n, per_row, seed, k, target, kw = CASES[case]
    csr = S.random_rect(m, n, per_row, seed)
    ref = S.solve("reference", csr, (m, n), k, target=target, method_stage1=api.PRIMME_GD_Olsen_plusK, **kw)
    got = S.solve("hostcheck", csr, (m, n), k, target=target, method_stage1=api.PRIMME_GD_Olsen_plusK, **kw)
    check(case, ref)
    check(case, got)
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-10)
    assert got["stats"] == ref["stats"], (got["stats"], ref["stats"])
    assert abs(got["aNorm"] - ref["aNorm"]) <= 1e-12 * ref["aNorm"]


def test_svds_out_of_scope_is_refused():
    m, n = 50, 20
    csr = S.random_rect(m, n, 4, 9)
    r = S.solve("hostcheck", csr, (m, n), 2, method=S.primme_svds_hybrid, eps=1e-8)
    assert r["ret"] == api.PRIMME_FUNCTION_UNAVAILABLE and r["initSize"] == 0


@pytest.mark.parametrize("m,n,target,preset,stage1", [
    (300, 100, S.primme_svds_largest, S.primme_svds_normalequations, api.PRIMME_DEFAULT_METHOD),
    (100, 300, S.primme_svds_smallest, S.primme_svds_normalequations, api.PRIMME_GD_Olsen_plusK),
    (500, 500, S.primme_svds_largest, S.primme_svds_default, api.PRIMME_DEFAULT_METHOD),
    (200, 50, S.primme_svds_closest_abs, S.primme_svds_augmented, api.PRIMME_JDQMR),
])
def test_svds_parameter_api_matches_reference(m, n, target, preset, stage1):
    """primme_svds_initialize / set_member / set_method / get_member: every scalar member of
    primme_svds_params and the derived first- and second-stage primme_params equal the reference's
    (primme_svds_interface.c:107-420)"""
    import ctypes as C
    vals = {}
    for which in ("reference", "hostcheck"):
        lib = {"reference": H.lib_reference, "hostcheck": H.lib_hostcheck}[which]()
        S.declare(lib)
        p = lib.primme_svds_params_create()
        for name, v in (("m", m), ("n", n), ("numSvals", 3), ("target", target), ("maxBasisSize", 24), ("maxBlockSize", 2),
                        ("aNorm", 7.5), ("eps", 1e-7), ("locking", 1), ("printLevel", 0), ("mLocal", m), ("nLocal", n)):
            S.set_member(lib, p, name, v)
        shifts = (C.c_double * 1)(0.25)
        if target == S.primme_svds_closest_abs:
            S.set_member(lib, p, "numTargetShifts", 1)
            S.set_member(lib, p, "targetShifts", C.addressof(shifts))
        assert lib.primme_svds_set_method(preset, stage1, api.PRIMME_DEFAULT_METHOD, p) == 0
        got = {name: S.get_member(lib, p, name) for name, (ident, kind) in S.LABEL.items()
               if kind in ("I", "D") and not name.startswith("stats_")}
        for stage in ("primme", "primmeStage2"):
            inner = C.cast(C.c_void_p(S.get_member(lib, p, stage)), C.POINTER(api.PrimmeParams)).contents
            for f in ("n", "nLocal", "numEvals", "target", "numTargetShifts", "locking", "initSize", "numOrthoConst",
                      "maxBasisSize", "minRestartSize", "maxBlockSize", "maxMatvecs", "aNorm", "eps", "printLevel",
                      "dynamicMethodSwitch", "initBasisMode"):
                got[stage + "." + f] = getattr(inner, f)
            got[stage + ".maxPrevRetain"] = inner.restartingParams.maxPrevRetain
            got[stage + ".maxInnerIterations"] = inner.correctionParams.maxInnerIterations
            got[stage + ".precondition"] = inner.correctionParams.precondition
            got[stage + ".convTest"] = inner.correctionParams.convTest
            got[stage + ".projection"] = inner.projectionParams.projection
            pr = inner.correctionParams.projectors
            got[stage + ".projectors"] = (pr.LeftQ, pr.LeftX, pr.RightQ, pr.RightX, pr.SkewQ, pr.SkewX)
        vals[which] = got
        lib.primme_svds_params_destroy(p)
    diff = {k: (vals["reference"][k], vals["hostcheck"][k]) for k in vals["reference"] if vals["reference"][k] != vals["hostcheck"][k]}
    assert not diff, diff
