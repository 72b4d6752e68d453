"""ctypes binding of the C-ABI (include/primme*.h, include/primme_b200.h).

This is the thin Python host layer of the package: it mirrors ``primme_params`` byte for byte
(reference include/primme_eigs.h:166-253), loads the product library
``primme_b200/libprimme_b200.so`` and offers a SciPy-like ``eigsh`` in the spirit of the
reference's Python binding (reference Python/primme.pyx: ``eigsh``).  The same ``PrimmeParams``
structure is used by the tests to drive the UNMODIFIED reference library (oracle/_ref) -- the
struct ABI is identical by construction, which is itself checked by tests/test_abi.py.

The product library has no CPU path: on a machine without a CUDA device the solvers return
PRIMME_FUNCTION_UNAVAILABLE (-44) and ``eigsh`` raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PRIMME_INT = C.c_int64

# enums (reference include/primme_eigs.h:47-107,256-273)
primme_proj_default, primme_proj_RR, primme_proj_harmonic, primme_proj_refined = range(4)
primme_smallest, primme_largest, primme_closest_geq, primme_closest_leq, primme_closest_abs, primme_largest_abs = range(6)
(PRIMME_DEFAULT_METHOD, PRIMME_DYNAMIC, PRIMME_DEFAULT_MIN_TIME, PRIMME_DEFAULT_MIN_MATVECS, PRIMME_Arnoldi,
 PRIMME_GD, PRIMME_GD_plusK, PRIMME_GD_Olsen_plusK, PRIMME_JD_Olsen_plusK, PRIMME_RQI, PRIMME_JDQR, PRIMME_JDQMR,
 PRIMME_JDQMR_ETol, PRIMME_STEEPEST_DESCENT, PRIMME_LOBPCG_OrthoBasis, PRIMME_LOBPCG_OrthoBasis_Window) = range(16)
METHODS = {name: val for name, val in globals().items() if name.startswith("PRIMME_") and isinstance(val, int)}
PRIMME_FUNCTION_UNAVAILABLE = -44
PRIMME_MAIN_ITER_FAILURE = -3


class PrimmeStats(C.Structure):
    _fields_ = [(n, PRIMME_INT) for n in (
        "numOuterIterations", "numRestarts", "numMatvecs", "numPreconds", "numGlobalSum", "numBroadcast",
        "volumeGlobalSum", "volumeBroadcast")] + [(n, C.c_double) for n in (
        "flopsDense", "numOrthoInnerProds", "elapsedTime", "timeMatvec", "timePrecond", "timeOrtho",
        "timeGlobalSum", "timeBroadcast", "timeDense", "estimateMinEVal", "estimateMaxEVal",
        "estimateLargestSVal", "estimateBNorm", "estimateInvBNorm", "maxConvTol", "estimateResidualError")] + [
        ("lockingIssue", PRIMME_INT)]


class JDProjectors(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("LeftQ", "LeftX", "RightQ", "RightX", "SkewQ", "SkewX")]


class ProjectionParams(C.Structure):
    _fields_ = [("projection", C.c_int)]


class CorrectionParams(C.Structure):
    _fields_ = [("precondition", C.c_int), ("robustShifts", C.c_int), ("maxInnerIterations", C.c_int),
                ("projectors", JDProjectors), ("convTest", C.c_int), ("relTolBase", C.c_double)]


class RestartingParams(C.Structure):
    _fields_ = [("maxPrevRetain", C.c_int)]


class PrimmeParams(C.Structure):
    pass


BLOCK_OP = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(PRIMME_INT), C.c_void_p, C.POINTER(PRIMME_INT),
                       C.POINTER(C.c_int), C.POINTER(PrimmeParams), C.POINTER(C.c_int))
GLOBAL_SUM = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(PrimmeParams),
                         C.POINTER(C.c_int))
BCAST = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int), C.POINTER(PrimmeParams), C.POINTER(C.c_int))

PrimmeParams._fields_ = [
    ("n", PRIMME_INT),
    ("matrixMatvec", C.c_void_p), ("matrixMatvec_type", C.c_int),
    ("applyPreconditioner", C.c_void_p), ("applyPreconditioner_type", C.c_int),
    ("massMatrixMatvec", C.c_void_p), ("massMatrixMatvec_type", C.c_int),
    ("numProcs", C.c_int), ("procID", C.c_int), ("nLocal", PRIMME_INT), ("commInfo", C.c_void_p),
    ("globalSumReal", C.c_void_p), ("globalSumReal_type", C.c_int),
    ("broadcastReal", C.c_void_p), ("broadcastReal_type", C.c_int),
    ("numEvals", C.c_int), ("target", C.c_int), ("numTargetShifts", C.c_int),
    ("targetShifts", C.POINTER(C.c_double)),
    ("dynamicMethodSwitch", C.c_int), ("locking", C.c_int), ("initSize", C.c_int), ("numOrthoConst", C.c_int),
    ("maxBasisSize", C.c_int), ("minRestartSize", C.c_int), ("maxBlockSize", C.c_int),
    ("maxMatvecs", PRIMME_INT), ("maxOuterIterations", PRIMME_INT), ("iseed", PRIMME_INT * 4),
    ("aNorm", C.c_double), ("BNorm", C.c_double), ("invBNorm", C.c_double), ("eps", C.c_double),
    ("orth", C.c_int), ("internalPrecision", C.c_int),
    ("printLevel", C.c_int), ("outputFile", C.c_void_p),
    ("matrix", C.c_void_p), ("preconditioner", C.c_void_p), ("massMatrix", C.c_void_p),
    ("ShiftsForPreconditioner", C.POINTER(C.c_double)), ("initBasisMode", C.c_int),
    ("ldevecs", PRIMME_INT), ("ldOPs", PRIMME_INT),
    ("projectionParams", ProjectionParams), ("restartingParams", RestartingParams),
    ("correctionParams", CorrectionParams), ("stats", PrimmeStats),
    ("convTestFun", C.c_void_p), ("convTestFun_type", C.c_int), ("convtest", C.c_void_p),
    ("monitorFun", C.c_void_p), ("monitorFun_type", C.c_int), ("monitor", C.c_void_p),
    ("queue", C.c_void_p), ("profile", C.c_char_p),
]


class VwxrCols(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("cb", C.c_int), ("ce", C.c_int)]


class VwxrOut(C.Structure):
    _fields_ = [("X", VwxrCols * 3), ("Wo", VwxrCols), ("R", VwxrCols), ("Rnorms_host", C.c_void_p),
                ("rb", C.c_int), ("re", C.c_int), ("rnorms_host", C.c_void_p),
                ("nG", C.c_int), ("G_host", C.c_void_p), ("ldG", C.c_int),
                ("nH", C.c_int), ("H_host", C.c_void_p), ("ldH", C.c_int),
                ("P_host", C.c_void_p), ("ldP", C.c_int), ("R2", C.c_void_p), ("ldR2", C.c_int64)]


class Jacobi(C.Structure):
    _fields_ = [("diag_dev", C.c_void_p), ("minabs", C.c_double), ("use_shifts", C.c_int)]


PRODUCT_LIB = os.path.join(HERE, "libprimme_b200.so")
_cache = {}


def _declare(lib):
    """argtypes/restype for the entry points both the product and the oracle export"""
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        "pb200_device_count": (i32, []),
        "pb200_ctx_create": (i32, [P(vp), i32]),
        "pb200_ctx_destroy": (i32, [vp]),
        "pb200_ctx_sync": (i32, [vp]),
        "pb200_ctx_launches": (i64, [vp]),
        "pb200_ctx_nranks": (i32, [vp]),
        "pb200_ctx_set_profiling": (i32, [vp, i32]),
        "pb200_ctx_get_profile": (i32, [vp, i32, P(i64), P(dbl), P(dbl)]),
        "pb200_malloc": (i32, [vp, C.c_size_t, P(vp)]),
        "pb200_free": (i32, [vp, vp]),
        "pb200_copy_h2d": (i32, [vp, vp, i64, vp, i64, i64, i32, i32]),
        "pb200_copy_d2h": (i32, [vp, vp, i64, vp, i64, i64, i32, i32]),
        "pb200_copy_d2d": (i32, [vp, vp, i64, vp, i64, i64, i32, i32]),
        "pb200_csr_create": (i32, [vp, i64, i64, i64, vp, vp, vp, i32, i32, P(vp)]),
        "pb200_csr_destroy": (i32, [vp, vp]),
        "pb200_csr_build_transpose": (i32, [vp, vp]),
        "pb200_dspmm": (i32, [vp, vp, vp, i64, vp, i64, i32]),
        "pb200_dspmm_t": (i32, [vp, vp, vp, i64, vp, i64, i32]),
        "pb200_dortho_sweep": (i32, [vp, i64, vp, i32, i64, vp, i32, i64, vp, i32, i64, vp, i32, vp, i32, i32, vp, i32]),
        "pb200_dvwxr": (i32, [vp, i64, vp, vp, i32, i64, vp, i32, i32, vp, P(VwxrOut)]),
        "pb200_dvwxr_can_fuse_gram": (i32, [vp, i64, vp, vp, i32, i64, i32, P(VwxrOut)]),
        "pb200_dpermute_columns": (i32, [vp, i64, vp, i64, vp, i32]),
        "pb200_dcopy_columns": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, i32]),
        "pb200_daxpy_columns": (i32, [vp, i64, vp, vp, i64, vp, i64, i32]),
        "pb200_dscale_columns": (i32, [vp, i64, vp, vp, i64, i32]),
        "pb200_dcolumn_dots": (i32, [vp, i64, vp, i64, vp, i64, i32, vp]),
        "pb200_dresidual_inplace": (i32, [vp, i64, vp, vp, i64, vp, i64, i32, vp]),
        "pb200_djacobi": (i32, [vp, i64, vp, vp, dbl, vp, i64, vp, i64, i32]),
    }
    for name, (res, args) in sig.items():
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
    for name in ("dprimme", "cublas_dprimme"):
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.restype, f.argtypes = i32, [vp, vp, vp, P(PrimmeParams)]
    if hasattr(lib, "primme_initialize"):
        lib.primme_initialize.restype, lib.primme_initialize.argtypes = None, [P(PrimmeParams)]
        lib.primme_set_method.restype, lib.primme_set_method.argtypes = i32, [i32, P(PrimmeParams)]
    if hasattr(lib, "primme_b200_dprimme_csr"):
        lib.primme_b200_dprimme_csr.restype = i32
        lib.primme_b200_dprimme_csr.argtypes = [vp, vp, vp, P(PrimmeParams), vp, vp, vp, i32]
        lib.primme_b200_attach_ctx.restype, lib.primme_b200_attach_ctx.argtypes = i32, [P(PrimmeParams), vp]
    return lib


def load_library(path=None):
    """Load (once) a shared library exporting the primme / pb200 C-ABI.  Default: the product."""
    path = path or PRODUCT_LIB
    if path not in _cache:
        if not os.path.exists(path):
            raise OSError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()' "
                "or `make lib`).  There is no pure-Python or CPU implementation of this package.")
        _cache[path] = _declare(C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2)))
    return _cache[path]


def new_params(lib, n, numEvals=1, target=primme_smallest, method=None, **kw):
    """primme_initialize + member assignment + primme_set_method, like the reference examples
    (examples/ex_eigs_dseq.c:60-95)."""
    p = PrimmeParams()
    lib.primme_initialize(C.byref(p))
    p.n = n
    p.numEvals = numEvals
    p.target = target
    p.printLevel = 0
    for key, val in kw.items():
        if key == "maxPrevRetain":
            p.restartingParams.maxPrevRetain = val
        elif key == "projection":
            p.projectionParams.projection = val
        elif key == "iseed":
            for i in range(4):
                p.iseed[i] = val[i]
        elif key == "targetShifts":
            arr = (C.c_double * len(val))(*val)
            p._keep_shifts = arr
            p.targetShifts = C.cast(arr, C.POINTER(C.c_double))
            p.numTargetShifts = len(val)
        else:
            setattr(p, key, val)
    if method is not None:
        rc = lib.primme_set_method(method, C.byref(p))
        if rc != 0:
            raise ValueError("primme_set_method failed")
    return p


def stats_dict(p):
    s = p.stats
    return {k: getattr(s, k) for k, _ in PrimmeStats._fields_}


def _target_from_which(which, sigma):
    """`which` / `sigma` of the reference's Python eigsh (Python/primme.pyx:508-544)"""
    table = {"LM": (primme_largest_abs, True), "LA": (primme_largest, False), "SA": (primme_smallest, False),
             "SM": (primme_closest_abs, True), "CLT": (primme_closest_leq, True), "CGT": (primme_closest_geq, True)}
    if isinstance(which, str):
        if which not in table:
            raise ValueError("which must be one of 'LM', 'SM', 'LA', 'SA', 'CLT', 'CGT' or a number")
        target, shifted = table[which]
        sigma = (0.0 if sigma is None else float(sigma)) if shifted else None
    else:
        if sigma is not None:
            raise ValueError("Giving a numeric value in `which`, and also giving `sigma`. Set only one of those.")
        target, sigma = primme_closest_abs, float(which)
    return target, sigma


def eigsh_csr(indptr, indices, data, k=6, sigma=None, which="SA", v0=None, ncv=None, maxiter=None, tol=0.0,
              return_eigenvectors=True, lock=None, method=PRIMME_GD_Olsen_plusK, maxBlockSize=1, maxBasisSize=0,
              minRestartSize=0, maxPrevRetain=0, aNorm=0.0, projection=None, lib=None, return_stats=False,
              raise_for_unconverged=True, **kw):
    """Eigenpairs of the symmetric CSR matrix (indptr, indices, data) on the GPU.

    SciPy-flavoured front end with the option names of the reference's Python ``eigsh``
    (Python/primme.pyx:284-600): ``which`` in 'LA', 'SA', 'LM', 'SM', 'CLT', 'CGT' or a number, ``sigma``,
    ``v0`` (n x i initial guesses), ``ncv`` (maxBasisSize), ``maxiter`` (maxOuterIterations), ``lock``,
    ``projection`` ('RR', 'refined', 'harmonic').  Host CSR in, host eigenpairs out; the matrix upload, the
    Davidson iteration and the download all go through ``primme_b200_dprimme_csr``."""
    lib = lib or load_library()
    n = len(indptr) - 1
    target, sigma = _target_from_which(which, sigma)
    extra = dict(kw)
    if sigma is not None:
        extra["targetShifts"] = [sigma]
    if ncv:
        maxBasisSize = int(ncv)
    if maxiter:
        extra["maxOuterIterations"] = int(maxiter)
    if lock is not None:
        extra["locking"] = 1 if lock else 0
    if minRestartSize:
        extra["minRestartSize"] = int(minRestartSize)
    if maxPrevRetain:
        extra["maxPrevRetain"] = int(maxPrevRetain)
    if projection is not None:
        extra["projection"] = {"RR": primme_proj_RR, "refined": primme_proj_refined, "harmonic": primme_proj_harmonic}[projection]
    init = None
    if v0 is not None:
        init = np.atleast_2d(np.asarray(v0, dtype=np.float64).T).T.reshape(n, -1)
        extra["initSize"] = min(init.shape[1], k)
    p = new_params(lib, n, numEvals=k, target=target, maxBlockSize=maxBlockSize, maxBasisSize=maxBasisSize,
                   eps=tol, aNorm=aNorm, method=method, **extra)
    rp = np.ascontiguousarray(indptr, dtype=np.int64)
    ci = np.ascontiguousarray(indices, dtype=np.int32)
    va = np.ascontiguousarray(data, dtype=np.float64)
    evals = np.zeros(k)
    rnorms = np.zeros(k)
    evecs = np.zeros((max(k, p.initSize), n))  # column-major n x k
    if init is not None:
        evecs[:p.initSize] = init[:, :p.initSize].T
    p.ldevecs = n
    rc = lib.primme_b200_dprimme_csr(evals.ctypes.data, evecs.ctypes.data, rnorms.ctypes.data, C.byref(p),
                                     rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0)
    if rc == PRIMME_FUNCTION_UNAVAILABLE:
        raise RuntimeError("primme_b200: no CUDA device (or feature outside the build's scope); "
                           "this package has no CPU fallback")
    if rc != 0 and (raise_for_unconverged or rc != PRIMME_MAIN_ITER_FAILURE):
        raise RuntimeError(f"dprimme returned {rc}")
    nconv = max(p.initSize, 0) if rc != 0 else k
    out = (evals[:nconv], evecs[:nconv].T.copy()) if return_eigenvectors else (evals[:nconv],)
    if return_stats:
        out = out + (dict(stats_dict(p), rnorms=rnorms[:nconv], initSize=p.initSize),)
    return out if len(out) > 1 else out[0]


def svds_csr(indptr, indices, data, shape, k=6, ncv=None, tol=0.0, which="LM", maxiter=None,
             return_singular_vectors=True, method="hybrid", methodStage1=PRIMME_DEFAULT_METHOD,
             methodStage2=PRIMME_DEFAULT_METHOD, maxBlockSize=0, lib=None, return_stats=False):
    """Singular triplets of the m x n CSR matrix on the GPU, with the option names of the reference's Python
    ``svds`` (Python/primme.pyx:1074-1400): ``which`` 'LM' / 'SM' / a number (closest singular values),
    ``method`` 'hybrid' (default) / 'normalequations' / 'augmented'.  Returns (U, s, Vt) like SciPy.

    The matrix and its transposed copy are uploaded once; ``cublas_dprimme_svds`` runs with the built-in
    device operator (``primme_b200_svds_csr_matvec``); the triplets come back to the host."""
    import re
    lib = lib or load_library()
    m, n = shape
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(os.path.dirname(here), "include", "primme_svds.h")).read()
    label = {mm.group(1): (int(mm.group(2)), mm.group(4)) for mm in
             re.finditer(r"X\((\w+),\s*(\d+),\s*([\w\.]+),\s*(\w+)\)", txt[txt.index("PRIMME_SVDS_PARAM_TABLE"):])}
    vp, i32 = C.c_void_p, C.c_int
    lib.primme_svds_params_create.restype = vp
    lib.primme_svds_set_member.argtypes = [vp, i32, vp]
    lib.primme_svds_get_member.argtypes = [vp, i32, vp]
    lib.primme_svds_set_method.argtypes = [i32, i32, i32, vp]
    lib.primme_svds_params_destroy.argtypes = [vp]
    lib.cublas_dprimme_svds.argtypes = [vp, vp, vp, vp]
    lib.pb200_csr_build_transpose.argtypes = [vp, vp]

    def put(p, name, value):
        ident, kind = label[name]
        v = C.c_int64(int(value)) if kind == "I" else C.c_double(float(value)) if kind == "D" else None
        rc = lib.primme_svds_set_member(p, ident, C.byref(v) if v is not None else C.c_void_p(value))
        if rc:
            raise ValueError(name)

    def get(p, name):
        ident, kind = label[name]
        v = C.c_int64() if kind == "I" else C.c_double() if kind == "D" else C.c_void_p()
        lib.primme_svds_get_member(p, ident, C.byref(v))
        return v.value

    rp = np.ascontiguousarray(indptr, dtype=np.int64)
    ci = np.ascontiguousarray(indices, dtype=np.int32)
    va = np.ascontiguousarray(data, dtype=np.float64)
    ctx, A, dsvecs = C.c_void_p(), C.c_void_p(), C.c_void_p()
    if lib.pb200_ctx_create(C.byref(ctx), -1) != 0:
        raise RuntimeError("primme_b200: no CUDA device; this package has no CPU fallback")
    p = None
    try:
        if lib.pb200_csr_create(ctx, m, n, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) != 0 \
                or lib.pb200_csr_build_transpose(ctx, A) != 0:
            raise RuntimeError("primme_b200: could not upload the matrix")
        p = lib.primme_svds_params_create()
        shifts = None
        if which == "LM":
            target = 0
        elif which == "SM":
            target = 1
        else:
            target = 2
            shifts = (C.c_double * 1)(float(which))
        for name, v in (("m", m), ("n", n), ("numSvals", k), ("target", target), ("printLevel", 0), ("eps", tol),
                        ("matrix", A.value), ("matrixMatvec", C.cast(lib.primme_b200_svds_csr_matvec, C.c_void_p).value)):
            put(p, name, v)
        if shifts is not None:
            put(p, "numTargetShifts", 1)
            put(p, "targetShifts", C.addressof(shifts))
        if ncv:
            put(p, "maxBasisSize", ncv)
        if maxiter:
            put(p, "maxMatvecs", maxiter)
        if maxBlockSize:
            put(p, "maxBlockSize", maxBlockSize)
        preset = {"hybrid": 1, "normalequations": 2, "augmented": 3}[method]
        if lib.primme_svds_set_method(preset, methodStage1, methodStage2, p) != 0:
            raise ValueError("primme_svds_set_method")
        inner = C.cast(C.c_void_p(get(p, "primme")), C.POINTER(PrimmeParams))
        lib.primme_b200_attach_ctx(inner, ctx)
        if lib.pb200_malloc(ctx, 8 * (m + n) * k, C.byref(dsvecs)) != 0:
            raise MemoryError
        svals, rn = np.zeros(k), np.zeros(k)
        rc = lib.cublas_dprimme_svds(svals.ctypes.data, dsvecs, rn.ctypes.data, p)
        lib.primme_b200_attach_ctx(inner, None)
        if rc != 0:
            raise RuntimeError(f"dprimme_svds returned {rc}")
        kk = get(p, "initSize")
        host = np.zeros((m + n) * k)
        lib.pb200_copy_d2h(ctx, dsvecs, (m + n) * k, host.ctypes.data, (m + n) * k, (m + n) * kk, 1, 8)
        U = host[: m * kk].reshape(kk, m).T.copy()
        Vt = host[m * kk: (m + n) * kk].reshape(kk, n).copy()
        stats = {s_: get(p, "stats_" + s_) for s_ in ("numOuterIterations", "numRestarts", "numMatvecs")}
        stats.update(rnorms=rn[:kk], aNorm=get(p, "aNorm"))
    finally:
        if p:
            lib.primme_svds_params_destroy(p)
        if dsvecs:
            lib.pb200_free(ctx, dsvecs)
        if A:
            lib.pb200_csr_destroy(ctx, A)
        lib.pb200_ctx_destroy(ctx)
    out = (U, svals[:kk], Vt) if return_singular_vectors else (svals[:kk],)
    if return_stats:
        out = out + (stats,)
    return out if len(out) > 1 else out[0]
