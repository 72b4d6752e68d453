"""Drives dprimme_svds of the reference, of the host-logic check library and of the product through
the parameter API the three share (primme_svds_params_create / set_member / get_member /
set_method), so no ctypes mirror of primme_svds_params is needed."""
import ctypes as C
import os
import re

import numpy as np

import harness as H
from primme_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_txt = open(os.path.join(ROOT, "include", "primme_svds.h")).read()
LABEL = {m.group(1): (int(m.group(2)), m.group(4)) for m in
         re.finditer(r"X\((\w+),\s*(\d+),\s*([\w\.]+),\s*(\w+)\)", _txt[_txt.index("PRIMME_SVDS_PARAM_TABLE"):])}

primme_svds_largest, primme_svds_smallest, primme_svds_closest_abs = 0, 1, 2
primme_svds_default, primme_svds_hybrid, primme_svds_normalequations, primme_svds_augmented = 0, 1, 2, 3


class CsrRect(C.Structure):
    """mirror of oracle/csr_host.c:csr_host_rect"""
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("rowptr", C.c_void_p), ("colind", C.c_void_p), ("vals", C.c_void_p)]


def declare(lib):
    vp, i32 = C.c_void_p, C.c_int
    lib.primme_svds_params_create.restype, lib.primme_svds_params_create.argtypes = vp, []
    lib.primme_svds_params_destroy.restype, lib.primme_svds_params_destroy.argtypes = i32, [vp]
    lib.primme_svds_set_member.restype, lib.primme_svds_set_member.argtypes = i32, [vp, i32, vp]
    lib.primme_svds_get_member.restype, lib.primme_svds_get_member.argtypes = i32, [vp, i32, vp]
    lib.primme_svds_set_method.restype, lib.primme_svds_set_method.argtypes = i32, [i32, i32, i32, vp]
    for f in ("dprimme_svds", "cublas_dprimme_svds"):
        if hasattr(lib, f):
            getattr(lib, f).restype, getattr(lib, f).argtypes = i32, [vp, vp, vp, vp]


def set_member(lib, p, name, value):
    ident, kind = LABEL[name]
    if kind == "I":
        v = C.c_int64(int(value))
        rc = lib.primme_svds_set_member(p, ident, C.byref(v))
    elif kind == "D":
        v = C.c_double(float(value))
        rc = lib.primme_svds_set_member(p, ident, C.byref(v))
    else:  # pointers and functions travel as the value itself
        rc = lib.primme_svds_set_member(p, ident, C.c_void_p(value))
    assert rc == 0, (name, rc)


def get_member(lib, p, name):
    ident, kind = LABEL[name]
    if kind == "I":
        v = C.c_int64()
    elif kind == "D":
        v = C.c_double()
    else:
        v = C.c_void_p()
    assert lib.primme_svds_get_member(p, ident, C.byref(v)) == 0, name
    return v.value


def solve(which, csr, shape, numSvals, target=primme_svds_largest, method=primme_svds_normalequations,
          method_stage1=api.PRIMME_DEFAULT_METHOD, method_stage2=api.PRIMME_DEFAULT_METHOD, device_entry=False, constraints=None, guesses=None, **kw):
    """dprimme_svds through `which` in {"reference", "hostcheck", "product"}; host contract (host
    svecs, host matvec callback from oracle/csr_host.c).  Returns dict(svals, rnorms, U, V, ret, stats)."""
    m, n = shape
    indptr, indices, data = csr
    rp = np.ascontiguousarray(indptr, dtype=np.int64)
    ci = np.ascontiguousarray(indices, dtype=np.int32)
    va = np.ascontiguousarray(data, dtype=np.float64)
    lib = {"reference": H.lib_reference, "hostcheck": H.lib_hostcheck, "product": H.lib_product}[which]()
    declare(lib)
    ok = H.lib_oracle_kernels()
    A = CsrRect(m, n, rp.ctypes.data, ci.ctypes.data, va.ctypes.data)
    p = lib.primme_svds_params_create()
    set_member(lib, p, "m", m)
    set_member(lib, p, "n", n)
    set_member(lib, p, "numSvals", numSvals)
    set_member(lib, p, "target", target)
    set_member(lib, p, "matrix", C.addressof(A))
    set_member(lib, p, "matrixMatvec", C.cast(ok.csr_host_svds_matvec, C.c_void_p).value)
    set_member(lib, p, "printLevel", 0)
    for k, v in kw.items():
        set_member(lib, p, k, v)
    # svecs on input = [Uc U0 Vc V0]: orthogonality constraints and initial guesses, left then right
    nc = constraints[0].shape[1] if constraints is not None else 0
    ng = guesses[0].shape[1] if guesses is not None else 0
    if nc:
        set_member(lib, p, "numOrthoConst", nc)
    if ng:
        set_member(lib, p, "initSize", ng)
    assert lib.primme_svds_set_method(method, method_stage1, method_stage2, p) == 0
    svals, rn = np.zeros(numSvals), np.zeros(numSvals)
    ncols = nc + max(numSvals, ng)
    svecs = np.zeros((m + n) * ncols)
    if nc or ng:
        left = [a for a in (constraints[0] if nc else None, guesses[0] if ng else None) if a is not None]
        right = [a for a in (constraints[1] if nc else None, guesses[1] if ng else None) if a is not None]
        L, Rr = np.hstack(left), np.hstack(right)
        svecs[: m * (nc + ng)] = L.T.ravel()
        svecs[m * (nc + ng): (m + n) * (nc + ng)] = Rr.T.ravel()
    # device_entry: cublas_dprimme_svds of the host-check library, where "device" memory is host memory --
    # runs the device-contract code path (kernels of the C-ABI for every vector operation) on the CPU
    entry = lib.cublas_dprimme_svds if device_entry else lib.dprimme_svds
    rc = entry(svals.ctypes.data, svecs.ctypes.data, rn.ctypes.data, p)
    k = get_member(lib, p, "initSize")
    kt = k + nc   # the constraints come back in front of the triplets found
    out = dict(ret=rc, svals=svals, rnorms=rn, initSize=k,
               U=svecs[: m * kt].reshape(kt, m).T[:, nc:].copy(), V=svecs[m * kt: (m + n) * kt].reshape(kt, n).T[:, nc:].copy(),
               stats={s: get_member(lib, p, "stats_" + s) for s in ("numOuterIterations", "numRestarts", "numMatvecs")},
               aNorm=get_member(lib, p, "aNorm"))
    lib.primme_svds_params_destroy(p)
    return out


def random_rect(m, n, per_row, seed):
    """sparse m x n matrix with `per_row` entries per row in U(-1,1) plus a graded diagonal band so
    that the extreme singular values are separated"""
    rng = np.random.default_rng(seed)
    cols = np.empty((m, per_row), dtype=np.int64)
    for i in range(m):
        cols[i] = np.sort(rng.choice(n, size=per_row, replace=False))
    vals = rng.uniform(-1.0, 1.0, size=(m, per_row))
    d = min(m, n)
    for i in range(d):
        j = np.searchsorted(cols[i], i % n)
        if j < per_row and cols[i, j] == i % n:
            vals[i, j] += 3.0 + 10.0 * i / d
        else:
            cols[i, 0], vals[i, 0] = i % n, 3.0 + 10.0 * i / d
            o = np.argsort(cols[i])
            cols[i], vals[i] = cols[i][o], vals[i][o]
    # duplicates after the forced diagonal: keep them (CSR allows repeated columns; they add up)
    indptr = np.arange(0, m * per_row + 1, per_row, dtype=np.int64)
    return indptr, cols.ravel().astype(np.int32), vals.ravel()


def dense(csr, shape):
    m, n = shape
    A = np.zeros((m, n))
    ip, ix, da = csr
    for i in range(m):
        for k in range(ip[i], ip[i + 1]):
            A[i, ix[k]] += da[k]
    return A
