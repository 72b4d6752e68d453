"""The reference's own regression driver (tests/driver.c + tests/COMMON, compiled UNCHANGED from the
reference tree by oracle/Makefile and linked against the product library) runs the reference's
double-precision configurations test_001..test_007 on the GPU: it reads LUNDA.mtx, calls dprimme
through the public API with the reference's host CSR matvec / Jacobi preconditioner callbacks, then
check_solution (tests/COMMON/ioandtest.c:86-150) verifies eigenvalues, residual norms,
orthogonality and the angle to the STORED reference solutions tests/sol_00N_double, and
checkInterface exercises primme_get_member/set_member on every field.  Exit code 0 = all checks
passed.  test_006 selects PRIMME_DEFAULT_MIN_TIME (JDQMR_ETol with the Jacobi preconditioner of the
driver), test_007 harmonic extraction."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver", "primme_double_b200")
DRIVER_SVDS = os.path.join(ROOT, "oracle", "_ref", "driver", "primmesvds_double_b200")
DATA = os.path.join(HERE, "golden", "driver")


def run(cfg, driver=DRIVER):
    return subprocess.run([driver, cfg], cwd=DATA, capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["test_201", "test_202", "test_205", "test_206", "test_207"])
def test_reference_svds_driver_passes_its_own_checks(cfg):
    """tests/driversvds.c, unchanged: rect.mtx / lund_b.mtx; 5 largest triplets with the default two-stage
    hybrid (201: eps 1e-6, 202: eps 1e-12) and with the augmented operator alone (207); the smallest
    triplet with the Jacobi-type preconditioner of the driver (205, 206: second stage = JDQMR with refined
    extraction on the augmented operator); check_solution_svds against the stored sol_20Nsvds_double.
    (203 and 204 -- 5 smallest, tens of thousands of iterations -- run on the CPU host-check build,
    tests/test_driver_cpu.py.)"""
    if not os.path.exists(DRIVER_SVDS):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    r = run(cfg, DRIVER_SVDS)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["test_001", "test_002", "test_003", "test_004", "test_005", "test_006", "test_007"])
def test_reference_driver_passes_its_own_checks(cfg):
    if not os.path.exists(DRIVER):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    r = run(cfg)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


DRIVER_Z = os.path.join(ROOT, "oracle", "_ref", "driver", "primme_doublecomplex_b200")


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["test_101", "test_102", "test_103", "test_104", "test_105", "test_106"])
def test_reference_complex_driver_passes_its_own_checks(cfg):
    """tests/driver.c compiled UNCHANGED with -DUSE_DOUBLECOMPLEX: the reference's six complex Hermitian
    configurations (mhd1280b.mtx) through zprimme on the GPU -- GD_Olsen_plusK with locking and a large basis
    (101), a basis of 3 (102), 50 pairs with soft locking (103), DEFAULT_MIN_TIME interior (104), Jacobi
    preconditioner interior (105) and largest (106); check_solution against the stored sol_10N_doublecomplex"""
    if not os.path.exists(DRIVER_Z):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    r = run(cfg, DRIVER_Z)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_reference_driver_links_against_the_product():
    """CPU-side: the unchanged driver links (every reference-internal symbol it needs is exported)
    and, without a GPU, the product refuses to run instead of computing on the host"""
    if not os.path.exists(DRIVER):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    import ctypes as C
    from primme_b200 import api
    lib = api.load_library()
    for sym in ("primme_get_context", "primme_free_context", "Mem_pop_frame", "Mem_pop_clean_frame",
                "Num_dot_dprimme", "Num_gemv_dprimme", "Num_larnv_dprimme", "ortho_single_iteration_dprimme"):
        assert hasattr(lib, sym), sym
    if lib.pb200_device_count() <= 0:
        r = run("test_001")
        assert r.returncode != 0
