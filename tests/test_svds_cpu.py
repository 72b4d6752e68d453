"""SVD front end: normal equations, augmented operator and the two-stage hybrid: the host logic (primme_b200/src/svds.c
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


# hybrid (normal equations, then the augmented operator started from the first stage's triplets) and the
# augmented operator alone; fixed methods in both stages so that the counts are comparable
TWO_STAGE = {
    # name: (m, n, per_row, seed, numSvals, preset, extra)
    "hybrid_loose": (600, 150, 6, 1, 5, S.primme_svds_hybrid, dict(eps=1e-6)),          # second stage has nothing left to do
    "hybrid_tight": (600, 150, 6, 1, 5, S.primme_svds_hybrid, dict(eps=1e-12)),         # second stage refines every triplet
    "hybrid_wide": (150, 500, 8, 2, 4, S.primme_svds_hybrid, dict(eps=1e-12)),          # AA' first
    "hybrid_block": (500, 200, 5, 3, 6, S.primme_svds_hybrid, dict(eps=1e-11, maxBlockSize=2, maxBasisSize=24)),
    "hybrid_locking": (400, 120, 6, 5, 5, S.primme_svds_hybrid, dict(eps=1e-12, locking=1)),
    "augmented": (600, 150, 6, 1, 5, S.primme_svds_augmented, dict(eps=1e-8)),
    "augmented_wide": (120, 300, 6, 7, 3, S.primme_svds_augmented, dict(eps=1e-9)),
}


def check_two_stage(case, r):
    m, n, per_row, seed, k, preset, kw = TWO_STAGE[case]
    A = S.dense(S.random_rect(m, n, per_row, seed), (m, n))
    sv = np.linalg.svd(A, compute_uv=False)
    assert r["ret"] == 0 and r["initSize"] == k
    assert np.allclose(np.sort(r["svals"])[::-1], sv[:k], rtol=0, atol=10 * kw["eps"] * sv[0])
    U, V = r["U"], r["V"]
    # every column of U and of V comes back normalised (primme_svds_c.c:958-985)
    assert np.allclose(np.linalg.norm(U, axis=0), 1, atol=1e-12) and np.allclose(np.linalg.norm(V, axis=0), 1, atol=1e-12)
    # the reported residual norm is the triplet's: sqrt(|A v - s u|^2 + |A' u - s v|^2) < eps |A|
    R = np.sqrt(np.linalg.norm(A @ V - U * r["svals"], axis=0) ** 2 + np.linalg.norm(A.T @ U - V * r["svals"], axis=0) ** 2)
    assert np.all(R < max(kw["eps"], 1e-15) * r["aNorm"] * 1.05 + 1e-13 * sv[0]), (R, kw["eps"] * r["aNorm"])


@pytest.mark.parametrize("case", sorted(TWO_STAGE))
def test_svds_two_stage_hostcheck_matches_reference(case):
    m, n, per_row, seed, k, preset, kw = TWO_STAGE[case]
    csr = S.random_rect(m, n, per_row, seed)
    args = dict(method=preset, method_stage1=api.PRIMME_GD_Olsen_plusK, method_stage2=api.PRIMME_GD_Olsen_plusK, **kw)
    ref = S.solve("reference", csr, (m, n), k, **args)
    got = S.solve("hostcheck", csr, (m, n), k, **args)
    check_two_stage(case, ref)
    check_two_stage(case, got)
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-12)
    assert got["stats"] == ref["stats"], (got["stats"], ref["stats"])
    assert abs(got["aNorm"] - ref["aNorm"]) <= 1e-12 * ref["aNorm"]


@pytest.mark.parametrize("case", ["hybrid_tight", "hybrid_wide", "augmented_wide"])
def test_svds_two_stage_device_contract_code_path(case):
    """cublas_dprimme_svds of the host-check build: same host logic, but every vector operation of the
    front end goes through the C-ABI (here its CPU restatement) instead of host loops"""
    m, n, per_row, seed, k, preset, kw = TWO_STAGE[case]
    csr = S.random_rect(m, n, per_row, seed)
    args = dict(method=preset, method_stage1=api.PRIMME_GD_Olsen_plusK, method_stage2=api.PRIMME_GD_Olsen_plusK, **kw)
    ref = S.solve("reference", csr, (m, n), k, **args)
    got = S.solve("hostcheck", csr, (m, n), k, device_entry=True, **args)
    check_two_stage(case, got)
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-12)
    assert got["stats"] == ref["stats"], (got["stats"], ref["stats"])


# smallest singular values through the hybrid method: the second stage is the augmented operator with
# refined extraction, closest_geq to lower bounds of the first stage's values, hard locking
SMALLEST = {
    "tall": (300, 60, 5, 4, 3, dict(eps=1e-11)),
    "wide": (120, 300, 5, 4, 2, dict(eps=1e-11)),
    "tall_jdqmr": (400, 90, 6, 8, 2, dict(eps=1e-12)),
}


@pytest.mark.parametrize("case", sorted(SMALLEST))
def test_svds_smallest_hybrid_hostcheck_matches_reference(case):
    m, n, per_row, seed, k, kw = SMALLEST[case]
    csr = S.random_rect(m, n, per_row, seed)
    st2 = api.PRIMME_JDQMR if "jdqmr" in case else api.PRIMME_GD_Olsen_plusK
    args = dict(target=S.primme_svds_smallest, method=S.primme_svds_hybrid, method_stage1=api.PRIMME_GD_Olsen_plusK,
                method_stage2=st2, maxMatvecs=400000, **kw)
    ref = S.solve("reference", csr, (m, n), k, **args)
    got = S.solve("hostcheck", csr, (m, n), k, **args)
    A = S.dense(csr, (m, n))
    sv = np.linalg.svd(A, compute_uv=False)[::-1][:k]
    for r in (ref, got):
        assert r["ret"] == 0 and r["initSize"] == k
        assert np.allclose(np.sort(r["svals"]), sv, rtol=0, atol=1e-9 * r["aNorm"])
        U, V = r["U"], r["V"]
        R = np.sqrt(np.linalg.norm(A @ V - U * r["svals"], axis=0) ** 2 + np.linalg.norm(A.T @ U - V * r["svals"], axis=0) ** 2)
        assert np.all(R < kw["eps"] * r["aNorm"] * 1.05 + 1e-13 * r["aNorm"])
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-10)
    for key in ("numOuterIterations", "numMatvecs"):
        assert abs(got["stats"][key] - ref["stats"][key]) <= max(4, 0.05 * ref["stats"][key]), (got["stats"], ref["stats"])


@pytest.mark.parametrize("preset", [S.primme_svds_normalequations, S.primme_svds_hybrid, S.primme_svds_augmented])
@pytest.mark.parametrize("shape", [(300, 80), (90, 260)])
def test_svds_constraints_and_initial_guesses(preset, shape):
    """svecs on input = [Uc U0 Vc V0] (primme_svds_c.c:640-668): the two leading triplets as orthogonality
    constraints, perturbed next ones as initial guesses; the solver must return the triplets that follow"""
    m, n = shape
    csr = S.random_rect(m, n, 6, 17)
    A = S.dense(csr, (m, n))
    U, sv, Vt = np.linalg.svd(A, full_matrices=False)
    rng = np.random.default_rng(3)
    cons = (U[:, :2].copy(), Vt[:2].T.copy())
    guess = (U[:, 2:4] + 1e-3 * rng.standard_normal((m, 2)), Vt[2:4].T + 1e-3 * rng.standard_normal((n, 2)))
    args = dict(method=preset, method_stage1=api.PRIMME_GD_Olsen_plusK, method_stage2=api.PRIMME_GD_Olsen_plusK,
                eps=1e-9 if preset == S.primme_svds_normalequations else 1e-11, constraints=cons, guesses=guess,
                maxMatvecs=40000)
    got = S.solve("hostcheck", csr, (m, n), 3, **args)
    # AA' (m < n) with constraints in the two-stage method: the reference hands its second stage right vectors
    # written at a wrong offset (see below) and does not converge (-203); compared to the dense SVD only
    ref = got if (m < n and preset == S.primme_svds_hybrid) else S.solve("reference", csr, (m, n), 3, **args)
    for r in (ref, got):
        assert r["ret"] == 0 and r["initSize"] == 3
        assert np.allclose(np.sort(r["svals"])[::-1], sv[2:5], rtol=1e-8)
    # orthogonal to the constraints, and triplets of A.  (Checked on ours only: with AA' and constraints the
    # reference writes V = A'U/sigma at an offset computed with the wrong leading dimension,
    # primme_svds_c.c:929-941 uses primme->nLocal = mLocal for the n-long right constraints.)
    assert np.abs(got["V"].T @ cons[1]).max() < 1e-7 and np.abs(got["U"].T @ cons[0]).max() < 1e-7
    assert np.linalg.norm(A @ got["V"] - got["U"] * got["svals"], axis=0).max() < 1e-7 * sv[0]
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-10)
    assert got["stats"] == ref["stats"], (got["stats"], ref["stats"])
    # device-contract code path of the same (vector shuffles through the C-ABI)
    dev = S.solve("hostcheck", csr, (m, n), 3, device_entry=True, **args)
    assert dev["ret"] == 0 and np.allclose(dev["svals"], ref["svals"], rtol=1e-10) and dev["stats"] == got["stats"]


def test_svds_default_method_runs_like_reference():
    """primme_svds_default = hybrid with PRIMME_DEFAULT_METHOD in both stages (run-time method choice:
    values and residuals are the criterion, not counts)"""
    m, n, k = 500, 120, 4
    csr = S.random_rect(m, n, 6, 21)
    ref = S.solve("reference", csr, (m, n), k, method=S.primme_svds_default, eps=1e-12)
    got = S.solve("hostcheck", csr, (m, n), k, method=S.primme_svds_default, eps=1e-12)
    assert ref["ret"] == 0 and got["ret"] == 0 and got["initSize"] == k
    assert np.allclose(got["svals"], ref["svals"], rtol=1e-12)
    # reported norms carry the sqrt(2) of the augmented normalisation (primme_svds_c.c:1017-1021)
    for r in (ref, got):
        assert np.all(r["rnorms"] < 2 * 1e-12 * r["aNorm"])


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
