"""N > 1 host logic on CPU: two gloo ranks run the row-sharded solve (host-check build); the
result must equal the single-rank spectrum, with residuals and orthonormality checked globally."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("bs,method", [(1, "PRIMME_GD_Olsen_plusK"), (3, "PRIMME_GD_Olsen_plusK"),
                                       (1, "PRIMME_JDQMR_ETol"), (2, "PRIMME_JDQMR"),
                                       # PRIMME_JDQR with a (Jacobi) preconditioner: the skew-Q projector, K^{-1}Q and
                                       # M = Q'K^{-1}Q reduced over the ranks
                                       (1, "PRIMME_JDQR+jacobi"), (2, "PRIMME_JDQR+jacobi")])
def test_two_rank_row_sharded_solve(bs, method):
    env = dict(os.environ, PB_BS=str(bs), PB_METHOD=method.split("+")[0], OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    if method.endswith("+jacobi"):
        env["PB_JACOBI"] = "1"
    port = 29511 + bs + (10 if "JDQMR" in method else 0) + (20 if "JDQR" in method else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "multi_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["rc"] == 0
    shape = (8, 11, 13)
    lam = [2 - 2 * np.cos(np.pi * np.arange(1, s + 1) / (s + 1)) for s in shape]
    exact = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())[:6]
    assert np.allclose(out["evals"], exact, rtol=1e-10)
    assert out["orth"] < 1e-7
    anorm = 12.0
    assert max(out["res"]) < 1e-10 * anorm * 1.1
    assert out["globalsums"] > 0


@pytest.mark.parametrize("preset,device_entry", [("primme_svds_normalequations", "0"), ("primme_svds_hybrid", "0"),
                                                 ("primme_svds_hybrid", "1"), ("primme_svds_augmented", "1")])
def test_two_rank_row_partitioned_svds(preset, device_entry):
    """dprimme_svds / cublas_dprimme_svds with numProcs = 2 (the layout of BASELINE config C4): rows of A, of the
    left and of the right vectors split over the ranks; singular values against a dense SVD, triplet
    residuals and orthonormality of the gathered vectors"""
    env = dict(os.environ, PB_PRESET=preset, PB_DEVICE_ENTRY=device_entry, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    port = 29560 + (1 if "hybrid" in preset else 0) + 2 * int(device_entry) + (4 if "augmented" in preset else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "svds_multi_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["rc"] == 0 and out["initSize"] == 4
    assert np.allclose(out["svals"], out["exact"], rtol=1e-10)
    assert out["orthV"] < 1e-8 and out["orthU"] < 1e-6
    tol = 1e-11 * out["aNorm"]
    if preset == "primme_svds_normalequations":
        assert max(out["res"]) < 1e-7 * out["aNorm"]      # accuracy of the normal equations
    else:
        assert max(out["res"]) < 1.05 * tol and max(out["rnorms"]) < 2 * tol
    assert out["globalsums"] > 0


@pytest.mark.parametrize("proj,bs", [("refined", 2), ("harmonic", 1)])
def test_two_rank_refined_and_harmonic_extraction(proj, bs):
    """row-sharded solve with Q next to V and W: the panels of the Q orthogonalisation, of Q'V and of the Q
    restart are reduced over the ranks like every other panel"""
    env = dict(os.environ, PB_BS=str(bs), PB_METHOD="PRIMME_GD_Olsen_plusK", PB_PROJ=proj, OMP_NUM_THREADS="1",
               OPENBLAS_NUM_THREADS="1")
    port = 29580 + bs + (3 if proj == "harmonic" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "multi_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["rc"] == 0
    shape = (8, 11, 13)
    lam = [2 - 2 * np.cos(np.pi * np.arange(1, s + 1) / (s + 1)) for s in shape]
    spec = (lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel()
    exact = spec[np.argsort(np.abs(spec - 0.35))][:3]
    assert np.allclose(np.sort(out["evals"]), np.sort(exact), rtol=1e-7)
    assert out["orth"] < 1e-7 and max(out["res"]) < 1e-8 * 12.0 * 1.1
