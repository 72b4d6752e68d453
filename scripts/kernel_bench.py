#!/usr/bin/env python
"""Per-kernel microbenchmark at the shapes of config C2 (n = 10^6, m = 28/40, b = 4) and C5 per GPU
(n = 1.25e6, m = 56/64, b = 8): CUDA-event time (the context's profiler) and achieved GB/s on
algorithmic bytes.  Also the target command for `ncu --set full -k regex:...` captures.
Usage: python scripts/kernel_bench.py [--reps 20] [--config c2|c5]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primme_b200 import api, matrices as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--mbar", type=int, default=0, help="basis size of the steady-state shapes (default 28 / 48)")
    ap.add_argument("--only", default="", help="comma list of name prefixes to run (default all)")
    args = ap.parse_args()
    lib = api.load_library(os.environ.get("PB200_LIB"))  # PB200_LIB: experimental build variants
    ctx = C.c_void_p()
    assert lib.pb200_ctx_create(C.byref(ctx), 0) == 0
    if args.config == "c2":
        n, mmax, b, mbar, rs = args.side ** 3, 40, 4, 28, 20
        csr = M.laplacian_nd((args.side,) * 3)
    else:
        n, mmax, b, mbar, rs = 1250000, 64, 8, 48, 32
        csr = M.power_law_rows(n, mean_degree=15.0, seed=7)
    if args.mbar > 0:
        mbar = args.mbar
    ld = (n + 15) // 16 * 16
    rng = np.random.default_rng(0)

    def dev(cols):
        p = C.c_void_p()
        assert lib.pb200_malloc(ctx, 8 * ld * cols, C.byref(p)) == 0
        blk = rng.standard_normal((min(cols, 8), ld)) / np.sqrt(n)
        for c0 in range(0, cols, blk.shape[0]):
            nc = min(blk.shape[0], cols - c0)
            lib.pb200_copy_h2d(ctx, blk.ctypes.data, ld, C.c_void_p(p.value + 8 * ld * c0), ld, ld, nc, 8)
        return p

    V, W = dev(mmax), dev(mmax)
    off = lambda p, col: C.c_void_p(p.value + 8 * ld * col)
    rp, ci, va = (np.ascontiguousarray(csr[0], np.int64), np.ascontiguousarray(csr[1], np.int32),
                  np.ascontiguousarray(csr[2], np.float64))
    A = C.c_void_p()
    assert lib.pb200_csr_create(ctx, n, n, len(ci), rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 0, 0, C.byref(A)) == 0

    results = {}

    only = [x for x in args.only.split(",") if x]

    def measure(name, kind, fn):
        if only and not any(name.startswith(o) for o in only):
            return
        fn(); fn()
        lib.pb200_ctx_set_profiling(ctx, 1)
        for _ in range(args.reps):
            fn()
        cnt, ms, by = C.c_int64(), C.c_double(), C.c_double()
        lib.pb200_ctx_get_profile(ctx, kind, C.byref(cnt), C.byref(ms), C.byref(by))
        lib.pb200_ctx_set_profiling(ctx, 0)
        results[name] = dict(launches=cnt.value, us=1e3 * ms.value / max(1, cnt.value),
                             MB=by.value / max(1, cnt.value) / 1e6,
                             GBps=(by.value / 1e9) / (ms.value / 1e3) if ms.value > 0 else 0)
        print(f"{name:34s} {results[name]['us']:9.1f} us  {results[name]['MB']:8.1f} MB  {results[name]['GBps']:8.1f} GB/s", flush=True)

    P = np.zeros((b, mmax + b + 1))
    Cm = rng.standard_normal((b, mmax + 1)) * 1e-3
    Y = np.eye(b) + 1e-3 * rng.standard_normal((b, b))
    m = mbar
    tag = "v3" if os.environ.get("PB200_SPMM_V3", "1") != "0" else "v2"
    measure(f"spmm {tag} b={b}", 0, lambda: lib.pb200_dspmm(ctx, A, off(V, m), ld, off(W, m), ld, b))
    if only == ["spmm"]:
        print(json.dumps(results))
        return
    measure(f"ortho gram m={m} b={b}", 1, lambda: lib.pb200_dortho_sweep(
        ctx, n, None, 0, ld, V, m, ld, off(V, m), b, ld, None, 0, None, 0, 1, P.ctypes.data, mmax + b + 1))
    measure(f"ortho update+gram m={m} b={b}", 1, lambda: lib.pb200_dortho_sweep(
        ctx, n, None, 0, ld, V, m, ld, off(V, m), b, ld, Cm.ctypes.data, mmax + 1, Y.ctypes.data, b, 1,
        P.ctypes.data, mmax + b + 1))
    measure(f"projection m+b={m + b} b={b}", 1, lambda: lib.pb200_dortho_sweep(
        ctx, n, None, 0, ld, V, m + b, ld, off(W, m), b, ld, None, 0, None, 0, 0, P.ctypes.data, mmax + b + 1))
    h = rng.standard_normal((mmax, mmax)) / np.sqrt(mmax)
    theta = rng.standard_normal(mmax)
    Rn = np.zeros(mmax)

    def cand():
        o = api.VwxrOut()
        o.X[0] = api.VwxrCols(off(V, m).value, ld, 0, b)
        o.R = api.VwxrCols(off(W, m).value, ld, 0, b)
        o.Rnorms_host = Rn.ctypes.data
        return lib.pb200_dvwxr(ctx, n, V, W, m, ld, h.ctypes.data, mmax, b, theta.ctypes.data, C.byref(o))

    measure(f"vwxr candidates m={m} b={b}", 2, cand)
    Pp = np.zeros((b, mmax + 8))

    def candp():
        o = api.VwxrOut()
        o.X[0] = api.VwxrCols(off(V, m).value, ld, 0, b)
        o.R = api.VwxrCols(off(W, m).value, ld, 0, b)
        o.Rnorms_host = Rn.ctypes.data
        o.P_host, o.ldP = Pp.ctypes.data, mmax + 8
        return lib.pb200_dvwxr(ctx, n, V, W, m, ld, h.ctypes.data, mmax, b, theta.ctypes.data, C.byref(o))

    measure(f"vwxr candidates+gram m={m} b={b}", 2, candp)
    G = np.zeros((mmax, mmax))
    Hm = np.zeros((mmax, mmax))

    def restart():
        o = api.VwxrOut()
        o.X[0] = api.VwxrCols(V.value, ld, 0, rs)
        o.Wo = api.VwxrCols(W.value, ld, 0, rs)
        o.X[1] = api.VwxrCols(off(V, rs).value, ld, 0, b)
        o.R = api.VwxrCols(off(W, rs).value, ld, 0, b)
        o.Rnorms_host = Rn.ctypes.data
        o.nG, o.G_host, o.ldG = rs, G.ctypes.data, mmax
        o.nH, o.H_host, o.ldH = rs, Hm.ctypes.data, mmax
        return lib.pb200_dvwxr(ctx, n, V, W, mmax, ld, h.ctypes.data, mmax, rs + b, theta.ctypes.data, C.byref(o))

    measure(f"vwxr restart m={mmax} rs={rs} b={b}", 2, restart)
    if os.environ.get("PB200_BENCH_OOP"):
        V2, W2 = dev(rs + b), dev(rs + b)

        def restart_oop():
            o = api.VwxrOut()
            o.X[0] = api.VwxrCols(V2.value, ld, 0, rs)
            o.Wo = api.VwxrCols(W2.value, ld, 0, rs)
            o.X[1] = api.VwxrCols(off(V2, rs).value, ld, 0, b)
            o.R = api.VwxrCols(off(W2, rs).value, ld, 0, b)
            o.Rnorms_host = Rn.ctypes.data
            o.nG, o.G_host, o.ldG = rs, G.ctypes.data, mmax
            o.nH, o.H_host, o.ldH = rs, Hm.ctypes.data, mmax
            return lib.pb200_dvwxr(ctx, n, V, W, mmax, ld, h.ctypes.data, mmax, rs + b, theta.ctypes.data, C.byref(o))

        measure(f"vwxr restart out of place", 2, restart_oop)
    # host-visible latency of one call at a tiny size (launch + panel reduction + delivery): the
    # fixed cost of every synchronisation point of the outer iteration
    import time
    nt = 1024

    def wall(name, fn, reps=300):
        for _ in range(20):
            fn()
        lib.pb200_ctx_sync(ctx)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        lib.pb200_ctx_sync(ctx)
        us = (time.perf_counter() - t0) / reps * 1e6
        results[name] = dict(wall_us=us)
        print(f"{name:34s} {us:9.1f} us wall per call (n={nt})", flush=True)

    wall("latency gram", lambda: lib.pb200_dortho_sweep(
        ctx, nt, None, 0, ld, V, m, ld, off(V, m), b, ld, None, 0, None, 0, 1, P.ctypes.data, mmax + b + 1))
    wall("latency update+gram", lambda: lib.pb200_dortho_sweep(
        ctx, nt, None, 0, ld, V, m, ld, off(V, m), b, ld, Cm.ctypes.data, mmax + 1, Y.ctypes.data, b, 1,
        P.ctypes.data, mmax + b + 1))

    def cand_small():
        o = api.VwxrOut()
        o.X[0] = api.VwxrCols(off(V, m).value, ld, 0, b)
        o.R = api.VwxrCols(off(W, m).value, ld, 0, b)
        o.Rnorms_host = Rn.ctypes.data
        return lib.pb200_dvwxr(ctx, nt, V, W, m, ld, h.ctypes.data, mmax, b, theta.ctypes.data, C.byref(o))

    wall("latency vwxr candidates", cand_small)
    print(json.dumps(results))


if __name__ == "__main__":
    main()
