"""Worker of tests/test_multi_rank.py: one rank of a row-sharded solve (PRIMME's SPMD model,
reference include/primme_eigs.h:187-198, examples/ex_eigs_mpi.c) on the CPU host-check build, with
torch.distributed/gloo standing in for NCCL: globalSumReal = all_reduce(SUM) on the host panel,
the SpMV halo = all_gather of the block."""
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
from primme_b200 import api, matrices as M  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    shape = (8, 11, 13)
    ip, ix, da = M.laplacian_nd(shape)
    n = len(ip) - 1
    lo, hi = n * rank // world, n * (rank + 1) // world
    nloc = hi - lo
    counts = [n * (r + 1) // world - n * r // world for r in range(world)]
    lip, lix, lda = ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], da[ip[lo]:ip[hi]]
    lib = H.lib_hostcheck()

    def matvec(x, ldx, y, ldy, bs, p, ierr):
        b = bs[0]
        xl = np.ctypeslib.as_array(C.cast(x, C.POINTER(C.c_double)), shape=(b * ldx[0],)).reshape(b, ldx[0])[:, :nloc]
        parts = [torch.empty((b, c), dtype=torch.float64) for c in counts]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(xl)))   # halo exchange
        xfull = torch.cat(parts, dim=1).numpy().T
        yl = M.csr_matvec(lip, lix, lda, xfull)
        yo = np.ctypeslib.as_array(C.cast(y, C.POINTER(C.c_double)), shape=(b * ldy[0],)).reshape(b, ldy[0])
        yo[:, :nloc] = yl.T
        ierr[0] = 0

    def gsum(send, recv, count, p, ierr):
        c = count[0]
        s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_double)), shape=(c,))
        t = torch.from_numpy(s.copy())
        dist.all_reduce(t)
        r = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_double)), shape=(c,))
        r[:] = t.numpy()
        ierr[0] = 0

    def precond(x, ldx, y, ldy, bs, p, ierr):
        # Jacobi, (diag - shift)^{-1} on the local rows, shifts from primme.ShiftsForPreconditioner (NULL: none)
        b = bs[0]
        xl = np.ctypeslib.as_array(C.cast(x, C.POINTER(C.c_double)), shape=(b * ldx[0],)).reshape(b, ldx[0])
        yo = np.ctypeslib.as_array(C.cast(y, C.POINTER(C.c_double)), shape=(b * ldy[0],)).reshape(b, ldy[0])
        sh = p.contents.ShiftsForPreconditioner
        for j in range(b):
            d = ldiag - (sh[j] if sh else 0.0)
            d = np.where(np.abs(d) < 1e-12, 1e-12, d)
            yo[j, :nloc] = xl[j, :nloc] / d
        ierr[0] = 0

    rows = np.repeat(np.arange(n), np.diff(ip))
    diag = np.zeros(n)
    diag[rows[rows == ix]] = da[rows == ix]
    ldiag = diag[lo:hi]
    mv, gs, pc = api.BLOCK_OP(matvec), api.GLOBAL_SUM(gsum), api.BLOCK_OP(precond)
    k = 6
    extra = {}
    proj = os.environ.get("PB_PROJ", "")
    if proj:
        # interior pairs with the QR factorisation of (A - tau I) V carried next to V and W
        k = 3
        extra = dict(target=api.primme_closest_abs, targetShifts=[0.35],
                     projection=api.primme_proj_refined if proj == "refined" else api.primme_proj_harmonic)
    p = api.new_params(lib, n, numEvals=k, maxBlockSize=int(os.environ.get("PB_BS", "3")), eps=1e-8 if proj else 1e-10,
                       numProcs=world, procID=rank, nLocal=nloc, **extra)
    p.matrixMatvec = C.cast(mv, C.c_void_p).value
    p.globalSumReal = C.cast(gs, C.c_void_p).value
    if os.environ.get("PB_JACOBI"):   # before primme_set_method: the preset decides `precondition` from it
        p.applyPreconditioner = C.cast(pc, C.c_void_p).value
    method = getattr(api, os.environ.get("PB_METHOD", "PRIMME_GD_Olsen_plusK"))
    assert lib.primme_set_method(method, C.byref(p)) == 0
    p.ldevecs = nloc
    evals, rn, evecs = np.zeros(k), np.zeros(k), np.zeros((k, nloc))
    rc = lib.cublas_dprimme(evals.ctypes.data, evecs.ctypes.data, rn.ctypes.data, C.byref(p))
    # gather the eigenvectors to check orthonormality / residuals globally on rank 0
    parts = [torch.empty((k, c), dtype=torch.float64) for c in counts]
    dist.all_gather(parts, torch.from_numpy(evecs))
    if rank == 0:
        X = torch.cat(parts, dim=1).numpy().T
        AX = M.csr_matvec(ip, ix, da, X)
        res = np.linalg.norm(AX - X * evals, axis=0)
        print("RESULT " + json.dumps(dict(rc=rc, evals=evals.tolist(), rnorms=rn.tolist(), res=res.tolist(),
                                          orth=float(np.abs(X.T @ X - np.eye(k)).max()),
                                          matvecs=p.stats.numMatvecs, globalsums=p.stats.numGlobalSum)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
