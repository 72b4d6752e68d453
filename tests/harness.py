"""Test harness: loaders for the product, the CPU host-check build and the unmodified reference,
and one `solve` function that drives any of them through the same ctypes `primme_params`.

Roles (see oracle/README in DESIGN.md):
  product    primme_b200/libprimme_b200.so              needs a GPU
  hostcheck  oracle/_build/libprimme_hostcheck.so       product host C code + oracle CPU kernels
  reference  oracle/_ref/libprimme_ref.so               UNMODIFIED PRIMME 3.2 built from /root/reference
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primme_b200 import api  # noqa: E402

HOSTCHECK = os.path.join(ROOT, "oracle", "_build", "libprimme_hostcheck.so")
ORACLE_KERNELS = os.path.join(ROOT, "oracle", "_build", "liboracle_kernels.so")
REFERENCE = os.path.join(ROOT, "oracle", "_ref", "libprimme_ref.so")
REFERENCE_SKEWQ = os.path.join(ROOT, "oracle", "_ref", "libprimme_ref_skewq.so")


def _ensure_built(path):
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "oracle"], cwd=ROOT, check=False)
    return os.path.exists(path)


def have_reference():
    return _ensure_built(REFERENCE) or os.path.exists(REFERENCE)


def have_gpu():
    try:
        lib = api.load_library()
    except OSError:
        return False
    return lib.pb200_device_count() > 0


def lib_product():
    return api.load_library()


def lib_hostcheck():
    _ensure_built(HOSTCHECK)
    return api.load_library(HOSTCHECK)


def lib_oracle_kernels():
    _ensure_built(ORACLE_KERNELS)
    return api.load_library(ORACLE_KERNELS)


def lib_reference():
    _ensure_built(REFERENCE)
    return api.load_library(REFERENCE)


def lib_reference_skewq():
    """the reference with the two one-line fixes of oracle/Makefile (refskewq): the only build of it that can run
    the skew-Q projector with a preconditioner; used by tests/test_jdqmr_cpu.py for that configuration alone"""
    _ensure_built(REFERENCE_SKEWQ)
    return api.load_library(REFERENCE_SKEWQ)


class CsrHost(C.Structure):
    """mirror of oracle/csr_host.c:csr_host"""
    _fields_ = [("n", C.c_int64), ("rowptr", C.c_void_p), ("colind", C.c_void_p), ("vals", C.c_void_p),
                ("nthreads", C.c_int), ("diag", C.c_void_p), ("minabs", C.c_double), ("use_shifts", C.c_int)]


def solve(which, csr, numEvals, target=api.primme_smallest, method=api.PRIMME_GD_Olsen_plusK, jacobi=False,
          nthreads=1, init_vecs=None, projectors=None, tweak=None, **kw):
    """Run dprimme through `which` in {"reference", "hostcheck", "product"} on the CSR triple.
    Returns dict(evals, rnorms, evecs (n x k, column order), ret, stats, initSize)."""
    indptr, indices, data = csr
    n = len(indptr) - 1
    rp = np.ascontiguousarray(indptr, dtype=np.int64)
    ci = np.ascontiguousarray(indices, dtype=np.int32)
    va = np.ascontiguousarray(data, dtype=np.float64)
    diag = None
    if jacobi:
        diag = np.zeros(n)
        rows = np.repeat(np.arange(n), np.diff(rp))
        m = rows == ci
        diag[rows[m]] = va[m]

    lib = {"reference": lib_reference, "reference_skewq": lib_reference_skewq, "hostcheck": lib_hostcheck,
           "product": lib_product}[which]()
    p = api.new_params(lib, n, numEvals=numEvals, target=target, **kw)
    keep = []
    if which.startswith("reference"):
        ok = lib_oracle_kernels()
        A = CsrHost(n, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, nthreads,
                    diag.ctypes.data if diag is not None else None, 1e-12 if jacobi else 0.0, 1)
        keep.append(A)
        p.matrix = C.addressof(A)
        p.matrixMatvec = C.cast(ok.csr_host_matvec, C.c_void_p).value
        if jacobi:
            p.preconditioner = C.addressof(A)
            p.applyPreconditioner = C.cast(ok.csr_host_jacobi, C.c_void_p).value
    else:
        if jacobi:
            # the callbacks must be known before primme_set_method decides `precondition`
            p.applyPreconditioner = C.cast(lib.primme_b200_jacobi_apply, C.c_void_p).value
    if method is not None:
        assert lib.primme_set_method(method, C.byref(p)) == 0
    if projectors is not None:   # (LeftQ, LeftX, RightQ, RightX, SkewQ, SkewX) on top of the preset
        pr = p.correctionParams.projectors
        pr.LeftQ, pr.LeftX, pr.RightQ, pr.RightX, pr.SkewQ, pr.SkewX = projectors

    if tweak is not None:        # last-minute edits of the struct (custom callbacks, ...)
        tweak(p)
    ncols = p.numOrthoConst + max(numEvals, p.initSize)
    evals = np.zeros(numEvals)
    rnorms = np.zeros(numEvals)
    evecs = np.zeros((ncols, n))
    if init_vecs is not None:
        evecs[: init_vecs.shape[1], :] = init_vecs.T
    p.ldevecs = n

    if which.startswith("reference"):
        rc = lib.dprimme(evals.ctypes.data, evecs.ctypes.data, rnorms.ctypes.data, C.byref(p))
    else:
        ctx = C.c_void_p()
        assert lib.pb200_ctx_create(C.byref(ctx), -1) == 0
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        jac = None
        ddiag = C.c_void_p()
        if jacobi:
            assert lib.pb200_malloc(ctx, 8 * n, C.byref(ddiag)) == 0
            assert lib.pb200_copy_h2d(ctx, diag.ctypes.data, n, ddiag, n, n, 1, 8) == 0
            jac = api.Jacobi(ddiag.value, 1e-12, 1)
            keep.append(jac)
            p.preconditioner = C.addressof(jac)
        rc = lib.primme_b200_dprimme_csr(evals.ctypes.data, evecs.ctypes.data, rnorms.ctypes.data, C.byref(p),
                                         rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0)
        launches = lib.pb200_ctx_launches(ctx)
        if jacobi:
            lib.pb200_free(ctx, ddiag)
        lib.primme_b200_attach_ctx(C.byref(p), None)
        lib.pb200_ctx_destroy(ctx)
    out = dict(evals=evals, rnorms=rnorms, evecs=evecs[p.numOrthoConst:].T.copy(), ret=rc,
               stats=api.stats_dict(p), initSize=p.initSize, params=p)
    if not which.startswith("reference"):
        out["launches"] = launches
    return out


def zsolve(which, csr, numEvals, target=api.primme_smallest, method=api.PRIMME_GD_Olsen_plusK, jacobi=False,
           nthreads=1, tweak=None, projectors=None, **kw):
    """zprimme through `which` in {"reference", "hostcheck", "product"}: csr = (indptr, indices, complex values) of a
    Hermitian matrix.  Returns dict(evals, rnorms, evecs (n x k complex), ret, stats)."""
    indptr, indices, data = csr
    n = len(indptr) - 1
    rp = np.ascontiguousarray(indptr, dtype=np.int64)
    ci = np.ascontiguousarray(indices, dtype=np.int32)
    va = np.ascontiguousarray(data, dtype=np.complex128)
    diag = None
    if jacobi:
        diag = np.zeros(n)
        rows = np.repeat(np.arange(n), np.diff(rp))
        m = rows == ci
        diag[rows[m]] = va[m].real
    lib = {"reference": lib_reference, "reference_skewq": lib_reference_skewq, "hostcheck": lib_hostcheck,
           "product": lib_product}[which]()
    p = api.new_params(lib, n, numEvals=numEvals, target=target, **kw)
    keep = []
    if which.startswith("reference"):
        ok = lib_oracle_kernels()
        A = CsrHost(n, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, nthreads,
                    diag.ctypes.data if diag is not None else None, 1e-12 if jacobi else 0.0, 1)
        keep.append(A)
        p.matrix = C.addressof(A)
        p.matrixMatvec = C.cast(ok.csr_host_zmatvec, C.c_void_p).value
        if jacobi:
            p.preconditioner = C.addressof(A)
            p.applyPreconditioner = C.cast(ok.csr_host_zjacobi, C.c_void_p).value
    elif jacobi:
        p.applyPreconditioner = C.cast(lib.primme_b200_zjacobi_apply, C.c_void_p).value
    if method is not None:
        assert lib.primme_set_method(method, C.byref(p)) == 0
    if projectors is not None:   # (LeftQ, LeftX, RightQ, RightX, SkewQ, SkewX) on top of the preset
        pr = p.correctionParams.projectors
        pr.LeftQ, pr.LeftX, pr.RightQ, pr.RightX, pr.SkewQ, pr.SkewX = projectors
    if tweak is not None:
        tweak(p)
    ncols = p.numOrthoConst + max(numEvals, p.initSize)
    evals = np.zeros(numEvals)
    rnorms = np.zeros(numEvals)
    evecs = np.zeros((ncols, n), dtype=np.complex128)
    p.ldevecs = n
    vp = C.c_void_p
    if which.startswith("reference"):
        lib.zprimme.restype, lib.zprimme.argtypes = C.c_int, [vp, vp, vp, C.POINTER(api.PrimmeParams)]
        rc = lib.zprimme(evals.ctypes.data, evecs.ctypes.data, rnorms.ctypes.data, C.byref(p))
    else:
        lib.primme_b200_zprimme_csr.restype = C.c_int
        lib.primme_b200_zprimme_csr.argtypes = [vp, vp, vp, C.POINTER(api.PrimmeParams), vp, vp, vp, C.c_int]
        ctx = C.c_void_p()
        assert lib.pb200_ctx_create(C.byref(ctx), -1) == 0
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        ddiag = C.c_void_p()
        if jacobi:
            assert lib.pb200_malloc(ctx, 8 * n, C.byref(ddiag)) == 0
            assert lib.pb200_copy_h2d(ctx, diag.ctypes.data, n, ddiag, n, n, 1, 8) == 0
            jac = api.Jacobi(ddiag.value, 1e-12, 1)
            keep.append(jac)
            p.preconditioner = C.addressof(jac)
        rc = lib.primme_b200_zprimme_csr(evals.ctypes.data, evecs.ctypes.data, rnorms.ctypes.data, C.byref(p),
                                         rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0)
        if jacobi:
            lib.pb200_free(ctx, ddiag)
        lib.primme_b200_attach_ctx(C.byref(p), None)
        lib.pb200_ctx_destroy(ctx)
    return dict(evals=evals, rnorms=rnorms, evecs=evecs[p.numOrthoConst:].T.copy(), ret=rc,
                stats=api.stats_dict(p), initSize=p.initSize, params=p)
