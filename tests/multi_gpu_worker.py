"""Worker of tests/test_multi_gpu.py and of bench.py's sharded mode: one rank = one GPU of a
row-sharded solve.  NCCL communicator created through the library's own C-ABI
(pb200_comm_unique_id / pb200_ctx_comm_init); the unique id travels over torch.distributed."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def setup_rank(lib, api, csr, rank, world, local):
    """returns (ctx, A_local, D, counts, (lo, hi)) with the NCCL communicator attached"""
    import torch
    import torch.distributed as dist
    ip, ix, da = csr
    n = len(ip) - 1
    counts = np.array([n * (r + 1) // world - n * r // world for r in range(world)], dtype=np.int64)
    lo, hi = n * rank // world, n * (rank + 1) // world
    ctx = C.c_void_p()
    assert lib.pb200_ctx_create(C.byref(ctx), local) == 0
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        assert lib.pb200_comm_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        uid = uid.cuda()
    dist.broadcast(uid, 0)
    raw = bytes(uid.cpu().numpy().tobytes())
    assert lib.pb200_ctx_comm_init(ctx, world, rank, raw) == 0
    # peer-memory panel exchange: all-gather the IPC handles of the exchange buffers
    if dist.get_backend() == "nccl" or torch.cuda.is_available():
        hbuf = (C.c_char * 64)()
        assert lib.pb200_ctx_peer_export(ctx, hbuf) == 0
        mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        rc = lib.pb200_ctx_peer_attach(ctx, world, rank, blob)
        assert rc in (0, 1), rc
    lip = np.ascontiguousarray(ip[lo:hi + 1] - ip[lo], dtype=np.int64)
    lix = np.ascontiguousarray(ix[ip[lo]:ip[hi]], dtype=np.int32)
    lda = np.ascontiguousarray(da[ip[lo]:ip[hi]], dtype=np.float64)
    A = C.c_void_p()
    assert lib.pb200_csr_create(ctx, hi - lo, n, len(lix), lip.ctypes.data, lix.ctypes.data, lda.ctypes.data, 0, 0, C.byref(A)) == 0
    D = C.c_void_p()
    assert lib.pb200_dist_csr_create(ctx, A, counts.ctypes.data, world, C.byref(D)) == 0
    return ctx, A, D, counts, (lo, hi)


def declare(lib):
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.pb200_comm_unique_id.restype, lib.pb200_comm_unique_id.argtypes = i32, [vp]
    lib.pb200_ctx_comm_init.restype, lib.pb200_ctx_comm_init.argtypes = i32, [vp, i32, i32, C.c_char_p]
    lib.pb200_ctx_comm_free.restype, lib.pb200_ctx_comm_free.argtypes = i32, [vp]
    lib.pb200_dist_csr_create.restype, lib.pb200_dist_csr_create.argtypes = i32, [vp, vp, vp, i32, C.POINTER(vp)]
    lib.pb200_dist_csr_destroy.restype, lib.pb200_dist_csr_destroy.argtypes = i32, [vp, vp]
    lib.pb200_ctx_peer_export.restype, lib.pb200_ctx_peer_export.argtypes = i32, [vp, vp]
    lib.pb200_ctx_peer_attach.restype, lib.pb200_ctx_peer_attach.argtypes = i32, [vp, i32, i32, C.c_char_p]
    lib.pb200_ctx_peer_active.restype, lib.pb200_ctx_peer_active.argtypes = i32, [vp]


def sharded_solve(lib, api, ctx, D, n, nloc, rank, world, devecs, evals, rn, **workload):
    p = api.new_params(lib, n, numProcs=world, procID=rank, nLocal=nloc, **workload)
    assert lib.primme_set_method(api.PRIMME_GD_Olsen_plusK, C.byref(p)) == 0
    p.ldevecs = max(nloc, 1)
    p.matrix = D
    p.matrixMatvec = C.cast(lib.primme_b200_dist_csr_matvec, C.c_void_p).value
    lib.primme_b200_attach_ctx(C.byref(p), ctx)
    rc = lib.cublas_dprimme(evals.ctypes.data, devecs, rn.ctypes.data, C.byref(p))
    lib.primme_b200_attach_ctx(C.byref(p), None)
    return rc, p


def main():
    import torch
    import torch.distributed as dist
    import harness as H
    from primme_b200 import api, matrices as M
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    lib = H.lib_product()
    declare(lib)
    shape = (32, 29, 37)
    csr = M.laplacian_nd(shape)
    n = len(csr[0]) - 1
    ctx, A, D, counts, (lo, hi) = setup_rank(lib, api, csr, rank, world, local)
    nloc = hi - lo
    k = 6
    devecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * max(nloc, 1) * k, C.byref(devecs)) == 0
    evals, rn = np.zeros(k), np.zeros(k)
    rc, p = sharded_solve(lib, api, ctx, D, n, nloc, rank, world, devecs, evals, rn, numEvals=k,
                          maxBlockSize=int(os.environ.get("PB_BS", "4")), maxBasisSize=40, eps=1e-10)
    X = np.zeros((k, nloc))
    assert lib.pb200_copy_d2h(ctx, devecs, nloc, X.ctypes.data, nloc, nloc, k, 8) == 0
    parts = [torch.empty((k, int(c)), dtype=torch.float64) for c in counts]
    if len(set(counts.tolist())) == 1:
        dist.all_gather(parts, torch.from_numpy(X))
    else:
        raise SystemExit("test uses equal shards")
    if rank == 0:
        Xf = torch.cat(parts, dim=1).numpy().T
        AX = M.csr_matvec(*csr, Xf)
        res = np.linalg.norm(AX - Xf * evals, axis=0)
        print("RESULT " + json.dumps(dict(rc=rc, evals=evals.tolist(), res=res.tolist(),
                                          orth=float(np.abs(Xf.T @ Xf - np.eye(k)).max()),
                                          matvecs=p.stats.numMatvecs, launches=lib.pb200_ctx_launches(ctx),
                                          peer_exchange=lib.pb200_ctx_peer_active(ctx))))
    lib.pb200_free(ctx, devecs)
    lib.pb200_dist_csr_destroy(ctx, D)
    lib.pb200_csr_destroy(ctx, A)
    lib.pb200_ctx_destroy(ctx)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
