#!/usr/bin/env python
"""bench.py -- headline measurement of the Davidson inner-loop hot path on B200.

Workload (BASELINE.json configs[1], "C2"): dprimme, 3-D 7-point Laplacian CSR n = 10^6
(nnz = 6 940 000), 10 smallest eigenpairs, GD_Olsen_plusK, maxBasisSize 40, maxBlockSize 4,
eps 1e-10, aNorm 12.  A "step" is one complete solve (time-to-converge); the metric is
matvecs/s = stats.numMatvecs / elapsed, summed over ranks.

  value     : matrix and eigenvector storage already resident in HBM (cublas_dprimme contract)
  e2e       : the same solve through the host-facing C-ABI call primme_b200_dprimme_csr: the
              host CSR upload and the eigenvector download are inside the timed region
  roofline  : the kernel kind with the largest share of device time, algorithmic bytes / CUDA-
              event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref, built from
              /root/reference by oracle/Makefile) on the host cores, on a bounded sample
              (maxMatvecs) of the same workload

Multi-GPU (--gpus N under torchrun, one rank per GPU): the SAME problem is row-sharded over the
ranks (PRIMME's SPMD model): every rank owns n/N rows of A, V, W and the eigenvectors; panels are
all-reduced on the device with NCCL and the SpMV halo is one grouped NCCL exchange per block
(primme_b200_dist_csr_matvec).  Total work is fixed => "scaling": "strong".  --multi replicas runs
N independent solves instead (weak).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=100, help="grid points per side (n = side^3)")
    ap.add_argument("--ref-matvecs", type=int, default=120,
                    help="bounded sample of the reference arm: matvecs per step")
    ap.add_argument("--ref-threads", type=int, default=0,
                    help="host threads of the reference arm (0: min(cores, 16); OpenBLAS on 100+ "
                         "threads is slower on these tall-skinny panels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "none"],
                    help="clock/throttle sampling during the timed region (none: measurement experiments only)")
    ap.add_argument("--sampler-period", type=float, default=0.25)
    ap.add_argument("--c5-n", type=int, default=10**7,
                    help="rows of the power-law matrix of config C5 reported in the `c5` block (0: skip)")
    ap.add_argument("--c5-steps", type=int, default=2)
    ap.add_argument("--c4-m", type=int, default=10**6,
                    help="rows of the rectangular matrix of config C4 (columns = 2/5 of it), `c4` block at N <= 2 (0: skip)")
    ap.add_argument("--c4-max-matvecs", type=int, default=4000)
    ap.add_argument("--c3-n", type=int, default=500000,
                    help="rows of the complex Hermitian matrix of config C3 reported in the `c3` block at N=1 (0: skip)")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="N>1: row-sharded solve with NCCL (strong scaling) or independent replicas")
    return ap.parse_args()


WORKLOAD = dict(numEvals=10, maxBasisSize=40, maxBlockSize=4, eps=1e-10, aNorm=12.0)


def workload_name(side):
    return (f"dprimme 3D 7-pt Laplacian CSR n={side}^3, 10 smallest, GD_Olsen_plusK, "
            f"maxBasisSize=40, blockSize=4, eps=1e-10")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index, mode="nvml", period=0.25):
        self.rows = []
        self.stop = False
        self.index = index
        self.mode = mode
        self.period = period
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        # in-process NVML when available (no process spawn, no driver re-initialisation while the
        # launch-latency-bound solver is being timed); nvidia-smi otherwise
        if self.mode == "none":
            return
        try:
            if self.mode != "nvml":
                raise RuntimeError("nvidia-smi requested")
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = get_reasons(h)
                act = lambda bit: "Active" if (r & bit) else "Not Active"
                self.rows.append([str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)])
                time.sleep(self.period)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(max(self.period, 0.5))

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kind, algorithmic_bytes_per_launch):
    """DRAM bytes per launch of the dominant kernel family: the ratio (dram__bytes_read.sum +
    dram__bytes_write.sum) / algorithmic bytes of the committed `ncu --set full` capture
    (profiles/ncu_traffic_r02.json, made by scripts/gpu_r2j.sh) applied to this run's
    algorithmic bytes per launch.  None when there is no capture for the family."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as f:
            ratio = json.load(f)["family_ratio"][kind]
        return ratio * algorithmic_bytes_per_launch
    except Exception:
        return None


def ref_threads(args):
    return args.ref_threads if args.ref_threads > 0 else min(os.cpu_count() or 1, 16)


def run_reference(args, csr, rank):
    """the reference's own CPU path on a bounded sample of the workload"""
    import harness as H
    from primme_b200 import api
    ncores = ref_threads(args)
    times, mvs = [], []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        kw = {k: v for k, v in WORKLOAD.items() if k != "numEvals"}
        if args.ref_matvecs > 0:     # bounded sample; <= 0: the whole solve (time-to-converge of the reference)
            kw["maxMatvecs"] = args.ref_matvecs
        r = H.solve("reference", csr, WORKLOAD["numEvals"], method=api.PRIMME_GD_Olsen_plusK, nthreads=ncores, **kw)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(r["stats"]["elapsedTime"] or dt)
            mvs.append(r["stats"]["numMatvecs"])
    return sum(mvs) / sum(times), sum(times) / len(times), ncores, mvs[0]


C5 = dict(numEvals=20, maxBasisSize=64, maxBlockSize=8, eps=1e-8)


def run_c5(args, lib, api, M, ctx, rank, world, dist):
    """Config C5 (BASELINE.json configs[4]): dprimme, symmetric power-law CSR n = 1e7, nnz ~ 1.5e8, 20 largest,
    GD_Olsen_plusK, maxBlockSize 8, maxBasisSize 64, eps 1e-8; the SAME matrix row-sharded over the ranks
    (every rank generates only its rows), compacted halo pushed over NVLink peer memory, panels all-reduced
    inside the kernels.  Returns the `c5` block of the JSON line (rank 0) -- the strong-scaling curve of the
    configuration the 8-GPU target is quoted on."""
    import torch
    n = args.c5_n
    lo, hi = n * rank // world, n * (rank + 1) // world
    nloc = hi - lo
    t0 = time.perf_counter()
    ip, ix, da = M.power_law_rows(n, lo, hi)
    t_gen = time.perf_counter() - t0
    nnz_loc = len(ix)
    A, D = C.c_void_p(), C.c_void_p()
    assert lib.pb200_csr_create(ctx, nloc, n, nnz_loc, ip.ctypes.data, ix.ctypes.data, da.ctypes.data, 0, 0, C.byref(A)) == 0
    del ip, ix, da
    counts = np.array([n * (r + 1) // world - n * r // world for r in range(world)], dtype=np.int64)
    halo = {}
    if world > 1:
        import multi_gpu_worker as MG
        MG.declare(lib)
        assert lib.pb200_dist_csr_create(ctx, A, counts.ctypes.data, world, C.byref(D)) == 0
        nl, nh, sent, peer = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        lib.pb200_dist_csr_info(D, C.byref(nl), C.byref(nh), C.byref(sent), C.byref(peer))
        halo = dict(rows_received_per_block=nh.value, rows_pushed_per_block=sent.value, peer_memory=peer.value)
    k = C5["numEvals"]
    devecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * max(nloc, 1) * k, C.byref(devecs)) == 0
    evals, rn = np.zeros(k), np.zeros(k)

    def solve():
        p = api.new_params(lib, n, target=api.primme_largest, **C5)
        assert lib.primme_set_method(api.PRIMME_GD_Olsen_plusK, C.byref(p)) == 0
        if world > 1:
            p.numProcs, p.procID, p.nLocal, p.ldevecs = world, rank, nloc, max(nloc, 1)
            p.matrix = D
            p.matrixMatvec = C.cast(lib.primme_b200_dist_csr_matvec, C.c_void_p).value
        else:
            p.ldevecs = n
            p.matrix = A
            p.matrixMatvec = C.cast(lib.primme_b200_csr_matvec, C.c_void_p).value
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        rc = lib.cublas_dprimme(evals.ctypes.data, devecs, rn.ctypes.data, C.byref(p))
        lib.primme_b200_attach_ctx(C.byref(p), None)
        assert rc == 0, rc
        return p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solve()  # warm-up: allocations, gather-layout timing, first-launch attributes
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mv = 0
    for _ in range(args.c5_steps):
        p = solve()
        mv += p.stats.numMatvecs
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stats = api.stats_dict(p)
    lib.pb200_ctx_set_profiling(ctx, 1)
    tw = time.perf_counter()
    solve()
    wall = time.perf_counter() - tw
    kinds = ["spmm", "ortho_sweep", "vwxr", "utils", "panel_reduce"]
    kern = {}
    dev_ms = 0.0
    for i, name in enumerate(kinds):
        cnt, pms, pby = C.c_int64(), C.c_double(), C.c_double()
        lib.pb200_ctx_get_profile(ctx, i, C.byref(cnt), C.byref(pms), C.byref(pby))
        dev_ms += pms.value
        kern[name] = {"GBps": round((pby.value / 1e9) / (pms.value / 1e3), 1) if pms.value > 0 else 0.0,
                      "ms": round(pms.value, 3), "launches": cnt.value}
    lib.pb200_ctx_set_profiling(ctx, 0)
    nnz = nnz_loc
    if world > 1:
        t = torch.tensor([float(nnz_loc)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        nnz = int(t.item())
    out = {"workload": f"dprimme power-law CSR n={n}, 20 largest, GD_Olsen_plusK, blockSize=8, maxBasisSize=64, eps=1e-8",
           "n": n, "nnz": nnz, "n_gpus": world, "scaling": "strong",
           "ms_per_solve": ms / args.c5_steps, "matvecs_per_s": mv / (ms * 1e-3),
           "matvecs_per_solve": stats["numMatvecs"], "outer_iterations": stats["numOuterIterations"],
           "restarts": stats["numRestarts"], "largest_eval": float(evals[0]), "max_resnorm": float(rn.max()),
           "kernels_rank0": kern, "device_time_share_of_solve_rank0": dev_ms / (wall * 1e3),
           "halo_rank0": dict(halo, bytes_pushed_per_block=halo.get("rows_pushed_per_block", 0) * 64),
           "generate_s_rank0": round(t_gen, 1)}
    lib.pb200_free(ctx, devecs)
    if world > 1:
        lib.pb200_dist_csr_destroy(ctx, D)
    lib.pb200_csr_destroy(ctx, A)
    return out


def run_c4(args, lib, api, M, ctx, rank, world, dist):
    """Config C4 (BASELINE.json configs[3]): dprimme_svds, rectangular CSR 10^6 x 4*10^5, nnz 2*10^7, 10 largest
    singular values, primme_svds_normalequations (stage 1 pinned to GD_Olsen_plusK), eps 1e-8,
    rows of A and of A^T partitioned over the ranks (named on 2 GPUs), built-in device operator."""
    import torch
    import scipy.sparse as sp
    import svds_harness as S
    S.declare(lib)
    m, n, k = args.c4_m, args.c4_m * 2 // 5, 10
    t0 = time.perf_counter()
    ip, ix, da = M.random_rectangular(m, n, per_row=20, seed=2024)
    As = sp.csr_matrix((da, ix, ip), shape=(m, n))
    At = As.T.tocsr()
    At.sort_indices()
    t_gen = time.perf_counter() - t0
    mc = np.array([m * (r + 1) // world - m * r // world for r in range(world)], dtype=np.int64)
    nc = np.array([n * (r + 1) // world - n * r // world for r in range(world)], dtype=np.int64)
    mlo, nlo, mloc, nloc = int(mc[:rank].sum()), int(nc[:rank].sum()), int(mc[rank]), int(nc[rank])
    keep = []

    def shard(mat, lo, cnt, ncols_global, counts):
        sub = mat[lo:lo + cnt]
        rp = np.ascontiguousarray(sub.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(sub.indices, dtype=np.int32)
        va = np.ascontiguousarray(sub.data, dtype=np.float64)
        A, D = C.c_void_p(), C.c_void_p()
        assert lib.pb200_csr_create(ctx, cnt, ncols_global, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
        assert lib.pb200_dist_csr_create(ctx, A, counts.ctypes.data, world, C.byref(D)) == 0
        keep.append((A, D))
        return D

    class Op(C.Structure):
        _fields_ = [("A", C.c_void_p), ("At", C.c_void_p)]

    import multi_gpu_worker as MG
    MG.declare(lib)
    op = Op(shard(As, mlo, mloc, n, nc).value, shard(At, nlo, nloc, m, mc).value)
    nnz = int(As.nnz)
    del As, At
    dsvecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * (mloc + nloc) * k, C.byref(dsvecs)) == 0
    svals, rn = np.zeros(k), np.zeros(k)

    def solve():
        p = lib.primme_svds_params_create()
        for name, v in (("m", m), ("n", n), ("mLocal", mloc), ("nLocal", nloc), ("numProcs", world), ("procID", rank),
                        ("numSvals", k), ("target", S.primme_svds_largest), ("printLevel", 0), ("eps", 1e-8),
                        ("maxBlockSize", 4), ("maxMatvecs", args.c4_max_matvecs), ("matrix", C.addressof(op)),
                        ("matrixMatvec", C.cast(lib.primme_b200_svds_dist_csr_matvec, C.c_void_p).value)):
            S.set_member(lib, p, name, v)
        assert lib.primme_svds_set_method(S.primme_svds_normalequations, api.PRIMME_GD_Olsen_plusK, api.PRIMME_GD_Olsen_plusK, p) == 0
        inner_p = C.cast(C.c_void_p(S.get_member(lib, p, "primme")), C.POINTER(api.PrimmeParams))
        lib.primme_b200_attach_ctx(inner_p, ctx)
        rc = lib.cublas_dprimme_svds(svals.ctypes.data, dsvecs, rn.ctypes.data, p)
        out = dict(rc=rc, converged=S.get_member(lib, p, "initSize"), matvecs=S.get_member(lib, p, "stats_numMatvecs"),
                   outer=S.get_member(lib, p, "stats_numOuterIterations"))
        lib.primme_b200_attach_ctx(inner_p, None)
        lib.primme_svds_params_destroy(p)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solve()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    r = solve()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"workload": f"cublas_dprimme_svds rectangular CSR {m}x{n}, 10 largest, normal equations, GD_Olsen_plusK block 4, eps=1e-8",
           "m": m, "n": n, "nnz": nnz, "n_gpus": world, "ms_per_solve": ms, "operator_applications": r["matvecs"],
           "applications_per_s": r["matvecs"] / (ms * 1e-3), "outer_iterations": r["outer"], "rc": r["rc"],
           "converged": r["converged"], "max_matvecs": args.c4_max_matvecs,
           "svals": [float(x) for x in np.sort(svals)[::-1]], "max_resnorm": float(rn.max()), "generate_s_rank0": round(t_gen, 1)}
    lib.pb200_free(ctx, dsvecs)
    for A_, D_ in keep:
        lib.pb200_dist_csr_destroy(ctx, D_)
        lib.pb200_csr_destroy(ctx, A_)
    return out


def run_c3(args, lib, api, M, ctx):
    """Config C3 (BASELINE.json configs[2]): zprimme, complex Hermitian CSR n = 5e5 (matrices.C3_MATRIX), 8 pairs
    closest to sigma = 0.5, JDQMR_ETol, Jacobi preconditioner, 1 GPU.  Matrix and eigenvectors resident in HBM."""
    import torch
    n = args.c3_n
    ip, ix, da = M.hermitian_c3(n, **M.C3_MATRIX)
    rows = np.repeat(np.arange(n), np.diff(ip))
    diag = np.zeros(n)
    diag[rows[rows == ix]] = da[rows == ix].real
    A = C.c_void_p()
    assert lib.pb200_csr_create(ctx, n, n, len(ix), ip.ctypes.data, ix.ctypes.data, da.ctypes.data, 0, 1, C.byref(A)) == 0
    ddiag, devecs = C.c_void_p(), C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * n, C.byref(ddiag)) == 0
    assert lib.pb200_copy_h2d(ctx, diag.ctypes.data, n, ddiag, n, n, 1, 8) == 0
    k = 8
    assert lib.pb200_malloc(ctx, 16 * n * k, C.byref(devecs)) == 0
    jac = api.Jacobi(ddiag.value, 1e-12, 1)
    evals, rn = np.zeros(k), np.zeros(k)
    lib.cublas_zprimme.restype = C.c_int
    lib.cublas_zprimme.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(api.PrimmeParams)]

    def solve():
        p = api.new_params(lib, n, numEvals=k, target=api.primme_closest_abs, targetShifts=[0.5], eps=1e-10)
        p.applyPreconditioner = C.cast(lib.primme_b200_zjacobi_apply, C.c_void_p).value
        assert lib.primme_set_method(api.PRIMME_JDQMR_ETol, C.byref(p)) == 0
        p.preconditioner = C.addressof(jac)
        p.matrix = A
        p.matrixMatvec = C.cast(lib.primme_b200_csr_matvec, C.c_void_p).value
        p.ldevecs = n
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        rc = lib.cublas_zprimme(evals.ctypes.data, devecs, rn.ctypes.data, C.byref(p))
        lib.primme_b200_attach_ctx(C.byref(p), None)
        assert rc == 0, rc
        return p

    solve()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.pb200_ctx_launches(ctx)
    ev0.record()
    mv = 0
    for _ in range(args.c5_steps):
        p = solve()
        mv += p.stats.numMatvecs
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = lib.pb200_ctx_launches(ctx) - l0
    st = api.stats_dict(p)
    lib.pb200_ctx_set_profiling(ctx, 1)
    solve()
    kern = {}
    for i, name in enumerate(["spmm", "ortho_sweep", "vwxr", "utils", "panel_reduce"]):
        cnt, pms, pby = C.c_int64(), C.c_double(), C.c_double()
        lib.pb200_ctx_get_profile(ctx, i, C.byref(cnt), C.byref(pms), C.byref(pby))
        kern[name] = {"GBps": round((pby.value / 1e9) / (pms.value / 1e3), 1) if pms.value > 0 else 0.0,
                      "ms": round(pms.value, 3), "launches": cnt.value}
    lib.pb200_ctx_set_profiling(ctx, 0)
    out = {"workload": f"zprimme complex Hermitian CSR n={n}, 8 closest to 0.5, JDQMR_ETol, Jacobi, eps=1e-10",
           "n": n, "nnz": int(len(ix)), "dtype": "c128", "ms_per_solve": ms / args.c5_steps,
           "matvecs_per_s": mv / (ms * 1e-3), "matvecs_per_solve": st["numMatvecs"],
           "outer_iterations": st["numOuterIterations"], "restarts": st["numRestarts"],
           "evals": [float(x) for x in np.sort(evals)], "max_resnorm": float(rn.max()),
           "gpu_launches_per_solve": launches / args.c5_steps, "kernels": kern}
    lib.pb200_free(ctx, devecs), lib.pb200_free(ctx, ddiag)
    lib.pb200_csr_destroy(ctx, A)
    return out


def main():
    args = parse()
    # OpenBLAS reads its thread count when the library is first loaded (the product links the same
    # OpenBLAS for its tiny host-side algebra and pins it to one thread during a solve)
    os.environ["OPENBLAS_NUM_THREADS"] = str(ref_threads(args))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # stdout carries exactly one JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from primme_b200 import api, matrices as M

    csr = M.laplacian_nd((args.side,) * 3)
    n = len(csr[0]) - 1
    nnz = len(csr[1])

    if args.impl == "reference":
        if rank != 0:
            return 0
        val, sec, ncores, mv = run_reference(args, csr, rank)
        line = {"impl": "reference", "metric": "matvecs_per_s", "value": val, "unit": "matvecs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.side), "n": n, "nnz": nnz},
                "cpu_baseline": {"value": val, "unit": "matvecs/s", "cores": ncores, "kind": "reference",
                                 "sample": (f"first {mv} matvecs of the solve (maxMatvecs bound), " if args.ref_matvecs > 0
                                            else f"the whole solve ({mv} matvecs, time to converge {sec:.1f} s), ") +
                                           f"OpenBLAS {ncores} threads + {ncores}-thread CSR callback"},
                "e2e": {"value": val, "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    lib = api.load_library()
    if lib.pb200_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")

    rp = np.ascontiguousarray(csr[0], dtype=np.int64)
    ci = np.ascontiguousarray(csr[1], dtype=np.int32)
    va = np.ascontiguousarray(csr[2], dtype=np.float64)
    k = WORKLOAD["numEvals"]

    sharded = world > 1 and args.multi == "sharded"
    peer_active = False
    D = None
    nloc = n
    if sharded:
        import multi_gpu_worker as MG
        MG.declare(lib)
        ctx, A, D, counts, (lo, hi) = MG.setup_rank(lib, api, csr, rank, world, local)
        nloc = hi - lo
        peer_active = bool(lib.pb200_ctx_peer_active(ctx))
    else:
        ctx = C.c_void_p()
        assert lib.pb200_ctx_create(C.byref(ctx), local) == 0
        A = C.c_void_p()
        assert lib.pb200_csr_create(ctx, n, n, nnz, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0
    devecs = C.c_void_p()
    assert lib.pb200_malloc(ctx, 8 * max(nloc, 1) * k, C.byref(devecs)) == 0
    evals, rn = np.zeros(k), np.zeros(k)

    def make_params():
        p = api.new_params(lib, n, target=api.primme_smallest, **WORKLOAD)
        assert lib.primme_set_method(api.PRIMME_GD_Olsen_plusK, C.byref(p)) == 0
        p.ldevecs = n
        return p

    def resident_solve():
        p = make_params()
        if sharded:
            p.numProcs, p.procID, p.nLocal, p.ldevecs = world, rank, nloc, max(nloc, 1)
            p.matrix = D
            p.matrixMatvec = C.cast(lib.primme_b200_dist_csr_matvec, C.c_void_p).value
        else:
            p.matrix = A
            p.matrixMatvec = C.cast(lib.primme_b200_csr_matvec, C.c_void_p).value
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        rc = lib.cublas_dprimme(evals.ctypes.data, devecs, rn.ctypes.data, C.byref(p))
        lib.primme_b200_attach_ctx(C.byref(p), None)
        assert rc == 0, rc
        return p

    # pinned host buffers for the end-to-end leg
    hevecs = torch.empty((k, n), dtype=torch.float64).pin_memory()
    h_rp = torch.from_numpy(rp).pin_memory()
    h_ci = torch.from_numpy(ci).pin_memory()
    h_va = torch.from_numpy(va).pin_memory()

    def e2e_solve():
        p = make_params()
        lib.primme_b200_attach_ctx(C.byref(p), ctx)
        rc = lib.primme_b200_dprimme_csr(evals.ctypes.data, hevecs.data_ptr(), rn.ctypes.data, C.byref(p),
                                         h_rp.data_ptr(), h_ci.data_ptr(), h_va.data_ptr(), 0)
        lib.primme_b200_attach_ctx(C.byref(p), None)
        assert rc == 0, rc
        return p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.pb200_ctx_launches(ctx)
        ev0.record()
        mv = 0
        for _ in range(steps):
            p = fn()
            mv += p.stats.numMatvecs
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            if not sharded:  # replicas: every rank did its own matvecs
                m = torch.tensor([mv], device="cuda", dtype=torch.float64)
                dist.all_reduce(m, op=dist.ReduceOp.SUM)
                mv = float(m.item())
        return ms, mv, lib.pb200_ctx_launches(ctx) - l0, p

    for _ in range(args.warmup):
        resident_solve()
    with ClockSampler(local, args.sampler, args.sampler_period) as cs:
        ms, mv, launches, p = timed(resident_solve, args.steps)
    clocks = cs.summary()
    value = mv / (ms * 1e-3)
    stats = api.stats_dict(p)

    if sharded:
        # end to end at N > 1: upload this rank's CSR shard, solve, download its eigenvector rows
        hl = torch.empty((k, max(nloc, 1)), dtype=torch.float64).pin_memory()
        lo_nz, hi_nz = int(rp[lo]), int(rp[hi])
        s_rp = torch.from_numpy(np.ascontiguousarray(rp[lo:hi + 1] - rp[lo])).pin_memory()
        s_ci = torch.from_numpy(np.ascontiguousarray(ci[lo_nz:hi_nz])).pin_memory()
        s_va = torch.from_numpy(np.ascontiguousarray(va[lo_nz:hi_nz])).pin_memory()
        cnts = np.ascontiguousarray(counts, dtype=np.int64)

        def e2e_sharded():
            A2, D2 = C.c_void_p(), C.c_void_p()
            assert lib.pb200_csr_create(ctx, nloc, n, hi_nz - lo_nz, s_rp.data_ptr(), s_ci.data_ptr(), s_va.data_ptr(),
                                        0, 0, C.byref(A2)) == 0
            assert lib.pb200_dist_csr_create(ctx, A2, cnts.ctypes.data, world, C.byref(D2)) == 0
            p = make_params()
            p.numProcs, p.procID, p.nLocal, p.ldevecs = world, rank, nloc, max(nloc, 1)
            p.matrix = D2
            p.matrixMatvec = C.cast(lib.primme_b200_dist_csr_matvec, C.c_void_p).value
            lib.primme_b200_attach_ctx(C.byref(p), ctx)
            rc = lib.cublas_dprimme(evals.ctypes.data, devecs, rn.ctypes.data, C.byref(p))
            lib.primme_b200_attach_ctx(C.byref(p), None)
            assert rc == 0, rc
            assert lib.pb200_copy_d2h(ctx, devecs, nloc, hl.data_ptr(), nloc, nloc, k, 8) == 0
            lib.pb200_dist_csr_destroy(ctx, D2)
            lib.pb200_csr_destroy(ctx, A2)
            return p

        e2e_fn = e2e_sharded
        h2d = (s_rp.numel() * 8 + s_ci.numel() * 4 + s_va.numel() * 8) * world
        d2h = 8 * n * k + 16 * k
    else:
        e2e_fn = e2e_solve
        h2d = rp.nbytes + ci.nbytes + va.nbytes
        d2h = 8 * n * k + 16 * k
    e2e_fn()  # warm the pinned-path once
    ms_e, mv_e, _, _ = timed(e2e_fn, max(1, min(args.steps, 3)))
    e2e_value = mv_e / (ms_e * 1e-3)

    # per-kernel roofline from one profiled solve (CUDA events on the kernels' stream)
    lib.pb200_ctx_set_profiling.restype = C.c_int
    lib.pb200_ctx_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.pb200_ctx_get_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pb200_ctx_set_profiling(ctx, 1)
    t0 = time.perf_counter()
    resident_solve()
    prof_wall = time.perf_counter() - t0
    kinds = ["spmm", "ortho_sweep", "vwxr", "utils", "panel_reduce"]
    prof = {}
    for i, name in enumerate(kinds):
        cnt, pms, pby = C.c_int64(), C.c_double(), C.c_double()
        lib.pb200_ctx_get_profile(ctx, i, C.byref(cnt), C.byref(pms), C.byref(pby))
        prof[name] = dict(launches=cnt.value, ms=pms.value, bytes=pby.value,
                          gbs=(pby.value / 1e9) / (pms.value / 1e3) if pms.value > 0 else 0.0)
    lib.pb200_ctx_set_profiling(ctx, 0)
    dom = max(("spmm", "ortho_sweep", "vwxr"), key=lambda kname: prof[kname]["ms"])
    peak, peak_src = measured_peak()
    roofline = {"bound": "hbm", "kernel": dom, "achieved": prof[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": prof[dom]["gbs"] / peak, "peak_source": peak_src,
                "traffic": ncu_traffic(dom, prof[dom]["bytes"] / max(1, prof[dom]["launches"])),
                "traffic_source": "profiles/ncu_traffic_r02.json (ncu --set full DRAM bytes / algorithmic bytes at the C2 shapes, cold L2)",
                "launches": prof[dom]["launches"], "avg_launch_us": 1e3 * prof[dom]["ms"] / max(1, prof[dom]["launches"]),
                "algorithmic_bytes_per_launch": prof[dom]["bytes"] / max(1, prof[dom]["launches"]),
                "all_kernels": {kname: {"GBps": round(v["gbs"], 1), "ms": round(v["ms"], 3), "launches": v["launches"]}
                                for kname, v in prof.items()},
                "device_time_share_of_solve": sum(v["ms"] for v in prof.values()) / (prof_wall * 1e3)}

    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        try:
            a2 = argparse.Namespace(**vars(args))
            a2.warmup, a2.steps = 0, 1
            val, sec, ncores, mvr = run_reference(a2, csr, 0)
            cpu_baseline = {"value": val, "unit": "matvecs/s", "cores": ncores, "kind": "reference",
                            "sample": f"first {mvr} matvecs of the same solve ({sec:.1f} s), "
                                      f"OpenBLAS {ncores} threads + {ncores}-thread CSR callback"}
        except Exception as e:  # the reference build is test infrastructure; never fatal here
            cpu_baseline = {"value": None, "unit": "matvecs/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"unavailable: {e}"}

    c5 = None
    if args.c5_n > 0 and (sharded or world == 1):
        try:
            c5 = run_c5(args, lib, api, M, ctx, rank, world, dist if world > 1 else None)
        except Exception as e:  # never lose the headline line
            c5 = {"error": repr(e)}

    c3 = None
    if args.c3_n > 0 and world == 1:
        try:
            c3 = run_c3(args, lib, api, M, ctx)
        except Exception as e:
            c3 = {"error": repr(e)}

    c4 = None
    if args.c4_m > 0 and world <= 2 and (sharded or world == 1):
        try:
            c4 = run_c4(args, lib, api, M, ctx, rank, world, dist if world > 1 else None)
        except Exception as e:
            c4 = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "matvecs_per_s", "value": value, "unit": "matvecs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if (sharded or world == 1) else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.side), "n": n, "nnz": nnz,
                           "parallelism": (f"row-sharded x{world}, " + ("in-kernel panel all-reduce over NVLink peer memory" if peer_active
                                                                         else "NCCL panel all-reduce") + " + compacted SpMV halo pushed over NVLink peer memory" if sharded
                                           else "replicas only" if args.gpus > 1 else "single GPU"),
                           "l2": "working set per sweep (V,W 2x320 MB at n=1e6) exceeds the 126 MB L2; no flush",
                           "outer_iterations": stats["numOuterIterations"], "restarts": stats["numRestarts"],
                           "matvecs_per_solve": stats["numMatvecs"], "time_to_converge_s": ms / args.steps / 1e3,
                           "host_timers_s": {"matvec": stats["timeMatvec"], "ortho": stats["timeOrtho"],
                                             "dense_vwxr": stats["timeDense"], "elapsed": stats["elapsedTime"]}},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "matvecs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e / max(1, min(args.steps, 3))},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "c5": c5, "c3": c3, "c4": c4}
        print(json.dumps(line))

    lib.pb200_free(ctx, devecs)
    if D is not None:
        lib.pb200_dist_csr_destroy(ctx, D)
    lib.pb200_csr_destroy(ctx, A)
    lib.pb200_ctx_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
