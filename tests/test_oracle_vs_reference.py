"""Pins the oracle: the plain-C restatement of the hot-path kernels (oracle/kernels_ref.c) against
the reference's OWN functions, called directly in the unmodified library built from
/root/reference (oracle/_ref/libprimme_ref.so):
   Num_update_VWXR_dprimme   (src/eigs/auxiliary_eigs_normal.c:155)  <-> pb200_dvwxr
   update_projection_dprimme (src/eigs/update_projection.c:81)       <-> pb200_dortho_sweep (Gram)
   Bortho_block_dprimme      (src/eigs/ortho.c:429)                  <-> span/orthonormality of the
                                                                         host block ortho + sweep
Needs the reference build, i.e. runs in the build container only."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from primme_b200 import api

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="reference build (oracle/_ref) not available")


class PrimmeContext(C.Structure):
    """reference src/include/common.h:610-641 (non-profile build)"""
    _fields_ = [("primme", C.c_void_p), ("primme_svds", C.c_void_p), ("printLevel", C.c_int),
                ("outputFile", C.c_void_p), ("report", C.c_void_p), ("mm", C.c_void_p),
                ("numProcs", C.c_int), ("procID", C.c_int), ("mpicomm", C.c_void_p),
                ("bcast", C.c_void_p), ("globalSum", C.c_void_p), ("queue", C.c_void_p)]


@pytest.fixture(scope="module")
def ref():
    lib = H.lib_reference()
    lib.primme_get_context.restype = PrimmeContext
    lib.primme_get_context.argtypes = [C.POINTER(api.PrimmeParams)]
    lib.primme_free_context.argtypes = [PrimmeContext]
    return lib


def ref_ctx(lib, n):
    p = api.new_params(lib, n, numEvals=1, method=api.PRIMME_GD_Olsen_plusK)
    p.nLocal = n
    p.numProcs = 1
    ctx = lib.primme_get_context(C.byref(p))
    return p, ctx


def test_vwxr_matches_reference_function(ref):
    ok = H.lib_oracle_kernels()
    rng = np.random.default_rng(3)
    n, m, nh, ld = 1300, 24, 12, 1311
    V = rng.standard_normal((m + 6, ld))
    W = rng.standard_normal((m + 6, ld))
    h = rng.standard_normal((nh, m))
    theta = rng.standard_normal(nh)
    rs, nconv, bs = 8, 2, 4   # restart size, converged, next block

    # ---- reference: in-place restart call exactly as restart_soft_locking issues it ----
    Vr, Wr = V.copy(), W.copy()
    Rn_r = np.zeros(bs)
    G_r = np.zeros((rs, rs + 1)); H_r = np.zeros((rs, rs + 2))
    p, ctx = ref_ctx(ref, n)
    dp = lambda a, off=0: C.c_void_p(a.ctypes.data + 8 * off)
    I, L = C.c_int, C.c_int64
    f = ref.Num_update_VWXR_dprimme
    f.restype = C.c_int
    NUL = C.c_void_p(None)
    rc = f(dp(Vr), dp(Wr), NUL, L(n), I(m), L(ld), dp(h), I(nh), I(m), dp(theta),
           dp(Vr), I(0), I(rs), L(ld),                       # X0 = V*h(:,0:rs)  (in place)
           dp(Vr, ld * rs), I(nconv), I(nconv + bs), L(ld),  # X1 = next block's Ritz vectors
           NUL, I(0), I(0), L(0),                            # X2
           dp(Wr), I(0), I(rs), L(ld),                       # Wo
           dp(Wr, ld * rs), I(nconv), I(nconv + bs), L(ld), dp(Rn_r),   # R, Rnorms
           NUL, I(0), I(0), L(0), NUL, I(0), I(0), L(0), NUL, I(0), I(0), L(0),   # BX0..2
           NUL, I(0), I(0),                                  # rnorms
           dp(G_r), I(rs), I(rs + 1), dp(H_r), I(rs), I(rs + 2),
           NUL, I(0), I(0), ctx)
    assert rc == 0
    ref.primme_free_context(ctx)

    # ---- oracle ----
    Vo, Wo = V.copy(), W.copy()
    Rn_o = np.zeros(bs)
    G_o = np.zeros((rs, rs + 1)); H_o = np.zeros((rs, rs + 2))
    o = api.VwxrOut()
    o.X[0] = api.VwxrCols(Vo.ctypes.data, ld, 0, rs)
    o.X[1] = api.VwxrCols(Vo.ctypes.data + 8 * ld * rs, ld, nconv, nconv + bs)
    o.Wo = api.VwxrCols(Wo.ctypes.data, ld, 0, rs)
    o.R = api.VwxrCols(Wo.ctypes.data + 8 * ld * rs, ld, nconv, nconv + bs)
    o.Rnorms_host = Rn_o.ctypes.data
    o.nG, o.G_host, o.ldG = rs, G_o.ctypes.data, rs + 1
    o.nH, o.H_host, o.ldH = rs, H_o.ctypes.data, rs + 2
    octx = C.c_void_p()
    ok.pb200_ctx_create(C.byref(octx), -1)
    assert ok.pb200_dvwxr(octx, n, Vo.ctypes.data, Wo.ctypes.data, m, ld, h.ctypes.data, m, nh, theta.ctypes.data, C.byref(o)) == 0
    ok.pb200_ctx_destroy(octx)

    assert np.allclose(Vo[:, :n], Vr[:, :n], rtol=1e-12, atol=1e-12)
    assert np.allclose(Wo[:, :n], Wr[:, :n], rtol=1e-12, atol=1e-12)
    assert np.allclose(Rn_o, Rn_r, rtol=1e-12)
    # the reference fills the upper triangle of G and H (symmetric flag); compare that part
    iu = np.triu_indices(rs)
    assert np.allclose(G_o[:, :rs].T[iu], G_r[:, :rs].T[iu], rtol=1e-11, atol=1e-9)
    assert np.allclose(H_o[:, :rs].T[iu], H_r[:, :rs].T[iu], rtol=1e-11, atol=1e-9)


def test_projection_panel_matches_reference_function(ref):
    ok = H.lib_oracle_kernels()
    rng = np.random.default_rng(4)
    n, m, b, ld, ldz = 900, 18, 4, 905, 30
    V = rng.standard_normal((m + b, ld))
    W = rng.standard_normal((m + b, ld))
    Zr = np.zeros((ldz, ldz))
    p, ctx = ref_ctx(ref, n)
    f = ref.update_projection_dprimme
    f.restype = C.c_int
    rc = f(C.c_void_p(V.ctypes.data), C.c_int64(ld), C.c_void_p(W.ctypes.data), C.c_int64(ld),
           C.c_void_p(Zr.ctypes.data), C.c_int64(ldz), C.c_int64(n), C.c_int(m), C.c_int(b), C.c_int(1), ctx)
    assert rc == 0
    ref.primme_free_context(ctx)
    Zo = np.zeros((b, ldz))
    octx = C.c_void_p()
    ok.pb200_ctx_create(C.byref(octx), -1)
    assert ok.pb200_dortho_sweep(octx, n, None, 0, 0, V.ctypes.data, m + b, ld, W.ctypes.data + 8 * ld * m, b, ld,
                                 None, 0, None, 0, 0, Zo.ctypes.data, ldz) == 0
    ok.pb200_ctx_destroy(octx)
    assert np.allclose(Zo[:, : m + b], Zr[m: m + b, : m + b], rtol=1e-12, atol=1e-10)
