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


@pytest.mark.parametrize("cfg", ["test_001", "test_002", "test_003", "test_004", "test_005", "test_006"])
def test_eigs_driver_hostcheck_passes_golden(cfg):
    r = run("primme_double_hostcheck", cfg)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("cfg", ["test_201", "test_202", "test_207"])
def test_svds_driver_hostcheck_passes_golden(cfg):
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


@pytest.mark.parametrize("cfg", ["test_203", "test_204", "test_205", "test_206"])
def test_svds_driver_out_of_scope_config_is_refused_before_any_work(cfg):
    """smallest singular values with the hybrid method: the second stage would need refined
    extraction; refused up front (no first stage is run for a result that cannot be finished)"""
    r = run("primmesvds_double_hostcheck", cfg)
    assert r.returncode != 0
    assert "outside the scope of this build" in r.stdout + r.stderr
    assert "1st Matvecs     : 0" in r.stdout


def test_svds_monitor_reports_like_the_reference(tmp_path):
    """primme_svds.monitorFun (default reporter at printLevel 3) through both stages of the hybrid method with
    fixed methods: the OUT lines -- iteration, converged count, block index, matvecs, singular value, stage --
    are the reference's, line by line, until rounding-level residuals (1e-14) first reorder an event"""
    import re
    cfg = open(os.path.join(DATA, "test_202")).read().replace("printLevel = 1", "printLevel = 3")
    cfg += "primme.method = PRIMME_GD_Olsen_plusK\nprimmeStage2.method = PRIMME_GD_Olsen_plusK\n"
    name = "test_monitor_tmp"
    path = os.path.join(DATA, name)
    open(path, "w").write(cfg)
    try:
        ours = run("primmesvds_double_hostcheck", name)
        ref = run("primmesvds_double_ref", name)
    finally:
        os.remove(path)
    assert ours.returncode == 0 and ref.returncode == 0

    def out_lines(text):
        rows = []
        for ln in text.splitlines():
            m = re.match(r"OUT (\d+) conv (\d+) blk (\d+) MV (\d+) Sec \S+ SV\s+(\S+) \|r\| (\S+) stage (\d)", ln)
            if m:
                rows.append((int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), round(float(m.group(5)), 6), int(m.group(7))))
        return rows

    a, b = out_lines(ours.stdout), out_lines(ref.stdout)
    assert len(b) > 300 and abs(len(a) - len(b)) <= 0.02 * len(b)
    assert a[:100] == b[:100]
    # (on rect.mtx the first stage already reaches eps: the second has nothing left to report)
    assert ("Lock striplet" in ours.stdout) == ("Lock striplet" in ref.stdout)
    assert "#Converged" in ours.stdout
