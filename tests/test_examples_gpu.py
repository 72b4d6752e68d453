"""The reference's own example examples/ex_eigs_dseq.c (config C1: 1-D Laplacian n=100, 10 smallest,
PRIMME_DYNAMIC, diagonal preconditioner, host dprimme contract), compiled UNCHANGED against
include/ and linked against the product library by `make examples` in the build container, runs on
the GPU and prints the right eigenvalues."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "examples", "ex_eigs_dseq")


@pytest.mark.skipif(not os.path.exists(EXE), reason="example binary not built (needs /root/reference at build time)")
def test_ex_eigs_dseq_runs_unchanged():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    evals = [float(m.group(1)) for m in re.finditer(r"Eval\[\d+\]:\s*([-0-9.eE+]+)", r.stdout)]
    assert len(evals) >= 10, r.stdout[-2000:]
    exact = 2 - 2 * np.cos(np.pi * np.arange(1, 11) / 101)
    assert np.allclose(sorted(evals[:10]), exact, rtol=1e-8)
