"""CPU tests (no GPU): the product's host control code (primme_b200/src/*.c) linked against the
oracle kernels (oracle/_build/libprimme_hostcheck.so) must reproduce the UNMODIFIED reference:
eigenvalues to 1e-10 relative, and identical outer-iteration / restart / matvec counts on
non-degenerate spectra.  The fixture comes from tests/golden/make_golden.py; when the reference
build is available (build container) it is also re-run live."""
import numpy as np
import pytest

import harness as H
import solver_checks as SC
from golden.cases import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_hostcheck_matches_reference_fixture(name):
    r = SC.run_case("hostcheck", name)
    SC.check_against_golden(name, r, counts="exact")


@pytest.mark.skipif(not H.have_reference(), reason="reference build (oracle/_ref) not available")
@pytest.mark.parametrize("name", ["aniso_b4_smallest", "lap2d_b3_locking", "lap2d_b1_jacobi"])
def test_fixture_is_current(name):
    """the committed fixture equals what the reference produces here (pins the oracle)"""
    r = SC.run_case("reference", name)
    g = SC.GOLDEN[name]
    assert np.allclose(r["evals"], g["evals"], rtol=1e-13, atol=0)
    s = r["stats"]
    assert (s["numOuterIterations"], s["numRestarts"], s["numMatvecs"]) == (
        g["numOuterIterations"], g["numRestarts"], g["numMatvecs"])


def test_dynamic_method():
    """PRIMME_DYNAMIC: GD+k <-> JDQMR chosen at run time from wall-clock timings by the reference's
    cost model (main_iter.c:427-437,601-624,1181-1187,1943-2440; dav_dynamic.c), so only
    eigenvalues / residuals are a parity criterion, plus the recommendation left in
    dynamicMethodSwitch (-1 GD+k, -2 JDQMR_ETol, -3 dynamic)"""
    from primme_b200 import api, matrices as M
    csr = M.laplacian_1d(100)
    r = H.solve("hostcheck", csr, 10, method=api.PRIMME_DYNAMIC, eps=1e-9, jacobi=True)
    SC.check_invariants(csr, r, 1e-9, 0.0)
    exact = 2 - 2 * np.cos(np.pi * np.arange(1, 11) / 101)
    assert np.allclose(r["evals"], exact, rtol=1e-10)
    assert r["params"].dynamicMethodSwitch in (-1, -2, -3)
    # a problem large enough for restarts: both methods get measured, the eigenvalues equal the
    # reference's (non-degenerate spectrum)
    csr = M.laplacian_nd((12, 11, 10))
    ref = H.solve("reference", csr, 6, method=api.PRIMME_DYNAMIC, eps=1e-9, jacobi=True)
    got = H.solve("hostcheck", csr, 6, method=api.PRIMME_DYNAMIC, eps=1e-9, jacobi=True)
    assert got["ret"] == 0 and np.allclose(got["evals"], ref["evals"], rtol=1e-9)
    SC.check_invariants(csr, got, 1e-9, 0.0)
    assert got["params"].dynamicMethodSwitch in (-1, -2, -3)


def test_initial_guesses_and_constraints():
    """warm start (initSize) and orthogonality constraints (numOrthoConst), init.c:173-208"""
    from primme_b200 import api, matrices as M
    csr = M.laplacian_nd((12, 15))
    n = 180
    base = H.solve("hostcheck", csr, 4, method=api.PRIMME_GD_Olsen_plusK, eps=1e-10)
    ref = H.solve("reference", csr, 4, method=api.PRIMME_GD_Olsen_plusK, eps=1e-10) if H.have_reference() else None
    # warm start with slightly perturbed solutions converges in far fewer iterations
    rng = np.random.default_rng(0)
    guess = base["evecs"] + 1e-6 * rng.standard_normal((n, 4))
    warm = H.solve("hostcheck", csr, 4, method=api.PRIMME_GD_Olsen_plusK, eps=1e-10, initSize=4, init_vecs=guess)
    SC.check_invariants(csr, warm, 1e-10, 0.0)
    assert warm["stats"]["numMatvecs"] < base["stats"]["numMatvecs"]
    if ref is not None:
        rw = H.solve("reference", csr, 4, method=api.PRIMME_GD_Olsen_plusK, eps=1e-10, initSize=4, init_vecs=guess)
        assert rw["stats"]["numMatvecs"] == warm["stats"]["numMatvecs"]
    # constraints: deflate the two lowest, the solver must return pairs 3.. of the spectrum
    cons = base["evecs"][:, :2]
    defl = H.solve("hostcheck", csr, 2, method=api.PRIMME_GD_Olsen_plusK, eps=1e-10, numOrthoConst=2, init_vecs=cons)
    assert defl["ret"] == 0
    assert np.allclose(defl["evals"], base["evals"][2:4], rtol=1e-9)


def test_error_codes():
    """input validation returns the reference's codes (primme_c.c:438-538)"""
    import ctypes as C
    from primme_b200 import api
    lib = H.lib_hostcheck()
    p = api.new_params(lib, 10, numEvals=20, method=api.PRIMME_GD)
    ev = np.zeros(20)
    assert lib.dprimme(ev.ctypes.data, ev.ctypes.data, ev.ctypes.data, C.byref(p)) == -7  # no matvec
    # what IS outside the scope is refused before any work: e.g. a mass matrix (generalised problem)
    p = api.new_params(lib, 100, numEvals=2)
    p.matrixMatvec = 1
    p.massMatrixMatvec = 1
    assert lib.dprimme(ev.ctypes.data, ev.ctypes.data, ev.ctypes.data, C.byref(p)) == -44


def test_custom_conv_test_sees_unit_ritz_vectors():
    """A user convTestFun is handed the Ritz VECTOR as evec (reference auxiliary_eigs_normal.c:408-443),
    never the residual block of the fused candidates sweep; with a block size > 1 and no preconditioner
    (the shape that fuses with the built-in test) the counts equal the reference's with the same callback."""
    import ctypes as C
    from primme_b200 import api, matrices as M
    csr = M.laplacian_nd((14, 11))
    n = len(csr[0]) - 1
    CONV = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                       C.c_void_p, C.POINTER(C.c_int))
    out = {}
    for which in ("reference", "hostcheck"):
        norms = []

        def conv(eval_, evec, rnorm, isconv, primme, ierr, norms=norms):
            if evec:
                v = np.ctypeslib.as_array(C.cast(evec, C.POINTER(C.c_double)), shape=(n,))
                norms.append(float(np.linalg.norm(v)))
            isconv[0] = 1 if rnorm[0] < 1e-9 * 8.0 else 0
            ierr[0] = 0

        cb = CONV(conv)

        def tweak(p, cb=cb):
            p.convTestFun = C.cast(cb, C.c_void_p).value

        r = H.solve(which, csr, 4, method=api.PRIMME_GD_Olsen_plusK, maxBlockSize=2, maxBasisSize=20, aNorm=8.0,
                    eps=1e-9, tweak=tweak)
        assert r["ret"] == 0
        assert len(norms) > 10 and np.allclose(norms, 1.0, atol=1e-8), (which, min(norms), max(norms))
        out[which] = (r["stats"]["numOuterIterations"], r["stats"]["numMatvecs"], r["evals"])
    assert out["reference"][:2] == out["hostcheck"][:2]
    assert np.allclose(out["reference"][2], out["hostcheck"][2], rtol=1e-10)
