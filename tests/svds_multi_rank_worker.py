"""Worker of tests/test_multi_rank.py::test_two_rank_row_partitioned_svds: one rank of a row-partitioned
dprimme_svds (reference include/primme_svds.h: mLocal / nLocal, globalSumReal; config C4 of BASELINE.json is
this layout on 2 GPUs) on the CPU host-check build, torch.distributed/gloo standing in for NCCL.  The rows of
A and of the left vectors are split in contiguous blocks, the right vectors likewise; the user matvec owns
the exchange: y = A x gathers x, y = A' x sums the local products and keeps its slice."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
import svds_harness as S  # noqa: E402
from primme_b200 import api  # noqa: E402

SVDS_MATVEC = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                          C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int))
SVDS_GSUM = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int))


def split(total, world):
    return [total * (r + 1) // world - total * r // world for r in range(world)]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m, n, k = int(os.environ.get("PB_M", "500")), int(os.environ.get("PB_N", "140")), 4
    preset = getattr(S, os.environ.get("PB_PRESET", "primme_svds_normalequations"))
    A = S.dense(S.random_rect(m, n, 6, 31), (m, n))
    mc, nc = split(m, world), split(n, world)
    mlo, nlo = sum(mc[:rank]), sum(nc[:rank])
    mloc, nloc = mc[rank], nc[rank]
    Aloc = A[mlo:mlo + mloc]
    lib = H.lib_hostcheck()
    S.declare(lib)

    def view(ptr, ld, rows, b):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(b * ld,)).reshape(b, ld)[:, :rows]

    def matvec(x, ldx, y, ldy, bs, trans, p, ierr):
        b = bs[0]
        if not trans[0]:
            xl = np.ascontiguousarray(view(x, ldx[0], nloc, b))
            parts = [torch.empty((b, c), dtype=torch.float64) for c in nc]
            dist.all_gather(parts, torch.from_numpy(xl))
            view(y, ldy[0], mloc, b)[:] = (Aloc @ torch.cat(parts, dim=1).numpy().T).T
        else:
            z = torch.from_numpy(np.ascontiguousarray((Aloc.T @ view(x, ldx[0], mloc, b).T).T))
            dist.all_reduce(z)
            view(y, ldy[0], nloc, b)[:] = z.numpy()[:, nlo:nlo + nloc]
        ierr[0] = 0

    def gsum(send, recv, count, p, ierr):
        c = count[0]
        t = torch.from_numpy(np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_double)), shape=(c,)).copy())
        dist.all_reduce(t)
        np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_double)), shape=(c,))[:] = t.numpy()
        ierr[0] = 0

    mv, gs = SVDS_MATVEC(matvec), SVDS_GSUM(gsum)
    p = lib.primme_svds_params_create()
    for name, v in (("m", m), ("n", n), ("mLocal", mloc), ("nLocal", nloc), ("numProcs", world), ("procID", rank),
                    ("numSvals", k), ("target", S.primme_svds_largest), ("printLevel", 0), ("eps", 1e-11),
                    ("matrixMatvec", C.cast(mv, C.c_void_p).value), ("globalSumReal", C.cast(gs, C.c_void_p).value)):
        S.set_member(lib, p, name, v)
    assert lib.primme_svds_set_method(preset, api.PRIMME_GD_Olsen_plusK, api.PRIMME_GD_Olsen_plusK, p) == 0
    svals, rn = np.zeros(k), np.zeros(k)
    svecs = np.zeros((mloc + nloc) * k)
    entry = lib.cublas_dprimme_svds if os.environ.get("PB_DEVICE_ENTRY") == "1" else lib.dprimme_svds
    rc = entry(svals.ctypes.data, svecs.ctypes.data, rn.ctypes.data, p)
    kk = S.get_member(lib, p, "initSize")
    U = torch.from_numpy(svecs[: mloc * kk].reshape(kk, mloc).copy())
    V = torch.from_numpy(svecs[mloc * kk: (mloc + nloc) * kk].reshape(kk, nloc).copy())
    Us = [torch.empty((kk, c), dtype=torch.float64) for c in mc]
    Vs = [torch.empty((kk, c), dtype=torch.float64) for c in nc]
    dist.all_gather(Us, U)
    dist.all_gather(Vs, V)
    if rank == 0:
        U, V = torch.cat(Us, dim=1).numpy().T, torch.cat(Vs, dim=1).numpy().T
        sv = np.linalg.svd(A, compute_uv=False)
        print("RESULT " + json.dumps(dict(
            rc=rc, initSize=kk, svals=svals.tolist(), exact=sv[:k].tolist(), rnorms=rn.tolist(),
            res=np.sqrt(np.linalg.norm(A @ V - U * svals, axis=0) ** 2 + np.linalg.norm(A.T @ U - V * svals, axis=0) ** 2).tolist(),
            orthU=float(np.abs(U.T @ U - np.eye(kk)).max()), orthV=float(np.abs(V.T @ V - np.eye(kk)).max()),
            aNorm=S.get_member(lib, p, "aNorm"), globalsums=S.get_member(lib, p, "stats_numGlobalSum"),
            matvecs=S.get_member(lib, p, "stats_numMatvecs"))))
    lib.primme_svds_params_destroy(p)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
