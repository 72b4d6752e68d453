"""The reference's own regression drivers (tests/driver.c, tests/driversvds.c + tests/COMMON, compiled
UNCHANGED by oracle/Makefile) linked against the HOST-CHECK build -- the product's host control code
over the CPU restatement of the kernels (test infrastructure, oracle/kernels_ref.c) -- run the
reference's hand-written double-precision configurations here, without a GPU, and pass the drivers'
own check_solution / check_solution_svds against the STORED golden solutions tests/sol_*: the host
logic is pinned to the reference's golden vectors on every CPU run, the kernels on the GPU run
(tests/test_driver_gpu.py)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRV = os.path.join(ROOT, "oracle", "_ref", "driver")
DATA = os.path.join(HERE, "golden", "driver")


def run(binary, cfg):
    path = os.path.join(DRV, binary)
    if not os.path.exists(path):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    return subprocess.run([path, cfg], cwd=DATA, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("cfg", ["test_001", "test_002", "test_003", "test_004", "test_005", "test_006", "test_007"])
def test_eigs_driver_hostcheck_passes_golden(cfg):
    """all seven hand-written double-precision eigenvalue configurations of the reference (007: harmonic
    extraction, closest_abs)"""
    r = run("primme_double_hostcheck", cfg)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("cfg", ["test_101", "test_102", "test_103", "test_104", "test_105", "test_106"])
def test_complex_eigs_driver_hostcheck_passes_golden(cfg):
    """the reference's six complex Hermitian configurations (mhd1280b.mtx; zprimme): tests/driver.c compiled
    UNCHANGED with -DUSE_DOUBLECOMPLEX, check_solution against the stored sol_10N_doublecomplex"""
    r = run("primme_doublecomplex_hostcheck", cfg)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("cfg", ["test_201", "test_202", "test_203", "test_204", "test_205", "test_206", "test_207"])
def test_svds_driver_hostcheck_passes_golden(cfg):
    """201, 202, 207: largest triplets (hybrid / augmented); 203-206: smallest triplets, where the second
    stage runs JDQMR with refined extraction and closest_geq shifts on the augmented operator"""
    r = run("primmesvds_double_hostcheck", cfg)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("cfg", ["test_201", "test_207"])
def test_svds_driver_prints_the_configuration_like_the_reference(cfg):
    """primme_svds_display_params / primme_display_params: the text the drivers print (and the
    reference's config reader parses back) is identical to the unmodified reference's"""
    ours = run("primmesvds_double_hostcheck", cfg).stdout.split("Error in")[0].split("Sval[")[0]
    ref = run("primmesvds_double_ref", cfg).stdout.split("Error in")[0].split("Sval[")[0]
    assert "primme_svds.methodStage2" in ours
    assert ours == ref
