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
    lib.pb200_ddist_spmm.restype, lib.pb200_ddist_spmm.argtypes = i32, [vp, vp, vp, i64, vp, i64, i32]
    lib.pb200_zdist_spmm.restype, lib.pb200_zdist_spmm.argtypes = i32, [vp, vp, vp, i64, vp, i64, i32]
    lib.pb200_dist_csr_info.restype = i32
    lib.pb200_dist_csr_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]


def dist_spmm_check(lib, ctx, csr, rank, world, counts, lo, hi, complex_vals=False, reps=3, bs=(1, 4, 8, 11)):
    """Y_local = (A X)[lo:hi] through the compacted peer-memory halo; several blocks back to back so that
    both gather buffers, the flags and the acknowledgements are exercised.  Returns max abs error / scale."""
    from primme_b200 import matrices as M
    ip, ix, da = csr
    n = len(ip) - 1
    nloc = hi - lo
    rng = np.random.default_rng(17)
    vals = da + (1j * rng.standard_normal(len(da)) if complex_vals else 0)
    es = 16 if complex_vals else 8
    lip = np.ascontiguousarray(ip[lo:hi + 1] - ip[lo], dtype=np.int64)
    lix = np.ascontiguousarray(ix[ip[lo]:ip[hi]], dtype=np.int32)
    lva = np.ascontiguousarray(vals[ip[lo]:ip[hi]], dtype=np.complex128 if complex_vals else np.float64)
    A, D = C.c_void_p(), C.c_void_p()
    assert lib.pb200_csr_create(ctx, nloc, n, len(lix), lip.ctypes.data, lix.ctypes.data, lva.ctypes.data, 0,
                                1 if complex_vals else 0, C.byref(A)) == 0
    assert lib.pb200_dist_csr_create(ctx, A, counts.ctypes.data, world, C.byref(D)) == 0
    nl, nh, sent, peer = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    lib.pb200_dist_csr_info(D, C.byref(nl), C.byref(nh), C.byref(sent), C.byref(peer))
    rows = np.repeat(np.arange(n), np.diff(ip))
    worst = 0.0
    for b in bs:
        dX, dY = C.c_void_p(), C.c_void_p()
        assert lib.pb200_malloc(ctx, es * max(nloc, 1) * b, C.byref(dX)) == 0
        assert lib.pb200_malloc(ctx, es * max(nloc, 1) * b, C.byref(dY)) == 0
        for rep in range(reps):
            X = rng.standard_normal((b, n)) + (1j * rng.standard_normal((b, n)) if complex_vals else 0)
            Xl = np.ascontiguousarray(X[:, lo:hi])
            assert lib.pb200_copy_h2d(ctx, Xl.ctypes.data, nloc, dX, nloc, nloc, b, es) == 0
            fn = lib.pb200_zdist_spmm if complex_vals else lib.pb200_ddist_spmm
            assert fn(ctx, D, dX, nloc, dY, nloc, b) == 0
            Yl = np.zeros((b, nloc), dtype=X.dtype)
            assert lib.pb200_copy_d2h(ctx, dY, nloc, Yl.ctypes.data, nloc, nloc, b, es) == 0
            for j in range(b):
                prod = vals * X[j, ix]
                ref = np.bincount(rows, weights=prod.real, minlength=n)
                if complex_vals:
                    ref = ref + 1j * np.bincount(rows, weights=prod.imag, minlength=n)
                worst = max(worst, float(np.abs(Yl[j] - ref[lo:hi]).max() / (np.abs(ref).max() + 1)))
        lib.pb200_free(ctx, dX), lib.pb200_free(ctx, dY)
    lib.pb200_dist_csr_destroy(ctx, D)
    lib.pb200_csr_destroy(ctx, A)
    return dict(err=worst, nhalo=nh.value, sent=sent.value, peer_halo=peer.value)


def svds_dist_check(lib, api, ctx, rank, world, m=30000, n=9000, per_row=8, k=5, eps=1e-9):
    """Row-partitioned cublas_dprimme_svds with the built-in operator (config C4's layout): rank r owns rows
    [m-range r] of A and rows [n-range r] of A^T; normal equations, GD_Olsen_plusK.  Returns on rank 0 the
    singular values, those of a sparse SVD of the whole matrix, the triplet residuals and orthogonality."""
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import svds_harness as S
    from primme_b200 import matrices as M
    S.declare(lib)
    ip, ix, da = M.random_rectangular(m, n, per_row=per_row, seed=2024)
    As = sp.csr_matrix((da, ix, ip), shape=(m, n))
    At = As.T.tocsr()
    At.sort_indices()
    mc = np.array([m * (r + 1) // world - m * r // world for r in range(world)], dtype=np.int64)
    nc = np.array([n * (r + 1) // world - n * r // world for r in range(world)], dtype=np.int64)
    mlo, nlo = int(mc[:rank].sum()), int(nc[:rank].sum())
    mloc, nloc = int(mc[rank]), int(nc[rank])

    def shard(mat, lo, cnt, ncols_global, counts):
        sub = mat[lo:lo + cnt]
        rp = np.ascontiguousarray(sub.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(sub.indices, dtype=np.int32)
        va = np.ascontiguousarray(sub.data, dtype=np.float64)
        A, D = C.c_void_p(), C.c_void_p()
        assert lib.pb200_csr_create(ctx, cnt, ncols_global, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
        assert lib.pb200_dist_csr_create(ctx, A, counts.ctypes.data, world, C.byref(D)) == 0
        return A, D

    A1, D1 = shard(As, mlo, mloc, n, nc)     # y = A x: x partitioned like the right vectors
    A2, D2 = shard(At, nlo, nloc, m, mc)     # y = A^T x: x partitioned like the left vectors

    class Op(C.Structure):
        _fields_ = [("A", C.c_void_p), ("At", C.c_void_p)]

    op = Op(D1.value, D2.value)
    p = lib.primme_svds_params_create()
    for name, v in (("m", m), ("n", n), ("mLocal", mloc), ("nLocal", nloc), ("numProcs", world), ("procID", rank),
                    ("numSvals", k), ("target", S.primme_svds_largest), ("printLevel", 0), ("eps", eps), ("maxBlockSize", 2),
                    ("matrix", C.addressof(op)),
                    ("matrixMatvec", C.cast(lib.primme_b200_svds_dist_csr_matvec, C.c_void_p).value)):
        S.set_member(lib, p, name, v)
    assert lib.primme_svds_set_method(S.primme_svds_normalequations, api.PRIMME_GD_Olsen_plusK, api.PRIMME_GD_Olsen_plusK, p) == 0
    inner = S.get_member(lib, p, "primme")
    inner_p = C.cast(C.c_void_p(inner), C.POINTER(api.PrimmeParams))
    lib.primme_b200_attach_ctx(inner_p, ctx)
    dsvecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * (mloc + nloc) * k, C.byref(dsvecs)) == 0
    svals, rn = np.zeros(k), np.zeros(k)
    rc = lib.cublas_dprimme_svds(svals.ctypes.data, dsvecs, rn.ctypes.data, p)
    kk = S.get_member(lib, p, "initSize")
    host = np.zeros((mloc + nloc) * k)
    assert lib.pb200_copy_d2h(ctx, dsvecs, (mloc + nloc) * k, host.ctypes.data, (mloc + nloc) * k, (mloc + nloc) * k, 1, 8) == 0
    U = torch.from_numpy(host[: mloc * k].reshape(k, mloc).copy())
    V = torch.from_numpy(host[mloc * k:].reshape(k, nloc).copy())
    Us = [torch.empty((k, int(c)), dtype=torch.float64) for c in mc]
    Vs = [torch.empty((k, int(c)), dtype=torch.float64) for c in nc]
    dist.all_gather(Us, U)
    dist.all_gather(Vs, V)
    out = None
    if rank == 0:
        Uf, Vf = torch.cat(Us, dim=1).numpy().T, torch.cat(Vs, dim=1).numpy().T
        want = np.sort(spl.svds(As, k=k, which="LM", tol=1e-12, return_singular_vectors=False))[::-1]
        out = dict(rc=rc, initSize=kk, svals=np.sort(svals)[::-1].tolist(), exact=want.tolist(),
                   res=float(np.linalg.norm(As @ Vf - Uf * svals, axis=0).max() / want[0]),
                   orthU=float(np.abs(Uf.T @ Uf - np.eye(k)).max()), orthV=float(np.abs(Vf.T @ Vf - np.eye(k)).max()),
                   matvecs=S.get_member(lib, p, "stats_numMatvecs"))
    lib.primme_b200_attach_ctx(inner_p, None)
    lib.pb200_free(ctx, dsvecs)
    lib.primme_svds_params_destroy(p)
    for A_, D_ in ((A1, D1), (A2, D2)):
        lib.pb200_dist_csr_destroy(ctx, D_)
        lib.pb200_csr_destroy(ctx, A_)
    return out


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
    # halo exchange on a matrix without locality (power law), real and complex
    pl = M.power_law_rows(30011, mean_degree=12.0, seed=3)
    npl = len(pl[0]) - 1
    cpl = np.array([npl * (r + 1) // world - npl * r // world for r in range(world)], dtype=np.int64)
    spmm_real = dist_spmm_check(lib, ctx, pl, rank, world, cpl, npl * rank // world, npl * (rank + 1) // world)
    spmm_cplx = dist_spmm_check(lib, ctx, pl, rank, world, cpl, npl * rank // world, npl * (rank + 1) // world,
                                complex_vals=True, bs=(1, 3, 8))
    errs = torch.tensor([spmm_real["err"], spmm_cplx["err"]], dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    svds = svds_dist_check(lib, api, ctx, rank, world) if os.environ.get("PB_BS", "4") == "4" else None
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
                                          peer_exchange=lib.pb200_ctx_peer_active(ctx),
                                          spmm_err=errs.tolist(), spmm_halo=spmm_real, svds=svds)))
    lib.pb200_free(ctx, devecs)
    lib.pb200_dist_csr_destroy(ctx, D)
    lib.pb200_csr_destroy(ctx, A)
    lib.pb200_ctx_destroy(ctx)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
