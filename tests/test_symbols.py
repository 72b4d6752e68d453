"""The product library loads on a machine without a GPU, exports every symbol the public headers
declare, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

from primme_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for hdr in ("primme_b200.h",):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b((?:pb200|primme_b200)_\w+)\s*\(", txt))
    names -= {"primme_b200_jacobi"}
    eigs = open(os.path.join(ROOT, "include", "primme_eigs.h")).read()
    for base in re.findall(r"^PRIMME_DECLARE_SOLVERS_\((\w+),", eigs, flags=re.M):
        names |= {base, "magma_" + base, "cublas_" + base}
    names |= set(re.findall(r"^(?:int|void|primme_params \*)\s*\*?(primme_\w+)\(", eigs, flags=re.M))
    # SVD front end: every solver flavour and the parameter API
    svds = open(os.path.join(ROOT, "include", "primme_svds.h")).read()
    for base in re.findall(r"^PRIMME_DECLARE_SVDS_ALL_\((\w+),", svds, flags=re.M):
        names |= {base, "magma_" + base, "cublas_" + base}
    names |= set(re.findall(r"^(?:int|void|primme_svds_params \*)\s*\*?(primme_svds_\w+)\(", svds, flags=re.M))
    # reference-internal entry points of the reference's own test driver
    internal = open(os.path.join(ROOT, "include", "primme_ref_internal.h")).read()
    internal = re.sub(r"/\*.*?\*/", "", internal, flags=re.S)
    names |= set(re.findall(r"^(?:int|void|double|primme_context_mirror)\s+(\w+)\(", internal, flags=re.M))
    return names


def test_all_declared_symbols_are_exported():
    lib = api.load_library()
    missing = [n for n in sorted(declared_functions()) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    lib = api.load_library()
    if lib.pb200_device_count() > 0:
        return  # on the GPU box this property cannot be observed
    ctx = C.c_void_p()
    assert lib.pb200_ctx_create(C.byref(ctx), -1) == -44
    p = api.new_params(lib, 10, numEvals=2, method=api.PRIMME_GD_Olsen_plusK)
    p.matrixMatvec = 1  # never called
    ev = np.zeros(40)
    assert lib.dprimme(ev.ctypes.data, ev.ctypes.data, ev.ctypes.data, C.byref(p)) == api.PRIMME_FUNCTION_UNAVAILABLE
    assert p.initSize == 0


def test_params_api_matches_reference_defaults():
    """primme_initialize / primme_set_method defaults of the benchmark configurations
    (C2 -> 40/16/4/4/locking 0; C5 -> 64/24/8/8; C3 -> 35/21/1/1/locking 1)"""
    lib = api.load_library()
    p = api.new_params(lib, 10**6, numEvals=10, maxBasisSize=40, maxBlockSize=4, method=api.PRIMME_GD_Olsen_plusK)
    assert (p.maxBasisSize, p.minRestartSize, p.maxBlockSize, p.restartingParams.maxPrevRetain, p.locking) == (40, 16, 4, 4, 0)
    p = api.new_params(lib, 10**7, numEvals=20, target=api.primme_largest, maxBasisSize=64, maxBlockSize=8,
                       method=api.PRIMME_GD_Olsen_plusK)
    assert (p.maxBasisSize, p.minRestartSize, p.maxBlockSize, p.restartingParams.maxPrevRetain, p.locking) == (64, 24, 8, 8, 0)
    p = api.new_params(lib, 5 * 10**5, numEvals=8, target=api.primme_closest_abs, targetShifts=[0.5],
                       method=api.PRIMME_JDQMR_ETol)
    assert (p.maxBasisSize, p.minRestartSize, p.maxBlockSize, p.restartingParams.maxPrevRetain, p.locking) == (35, 21, 1, 1, 1)
    assert p.correctionParams.maxInnerIterations == -1 and p.correctionParams.projectors.LeftX == 1
