"""The reference's generated interface tests (`make tests_primme_interface all_tests_double`,
tests/Makefile:99-189): 3192 configurations -- 14 preset methods x Laplacians of size 0...100 x 0...100 wanted
pairs x smallest / largest / closest_abs / closest_geq x Rayleigh-Ritz / refined extraction -- each run by the
reference's own driver (compiled unchanged, linked against the host-check build: product host code over the
CPU restatement of the kernels) and verified by its check_solution against the STORED solutions
tests/sol_testi-*_double.  ALL 3192 must pass, like the unmodified reference: the presets whose block size equals
numEvals (LOBPCG_OrthoBasis, STEEPEST_DESCENT with up to 100 pairs) run their blocks in chunks of the kernels'
8-column panels."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

import gen_interface_configs as G

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver", "primme_double_hostcheck")


def out_of_scope(name, method):
    return False


def test_generated_interface_configurations(tmp_path):
    if not os.path.exists(DRIVER):
        pytest.skip("driver binary not built (needs the reference tree at build time)")
    work = str(tmp_path / "iface")
    names = G.write_all(work)
    os.symlink(os.path.join(HERE, "golden", "driver", "tests"), os.path.join(work, "tests"))
    assert len(names) == 3192

    def run(item):
        name, method = item
        r = subprocess.run([DRIVER, name + ".F"], cwd=work, capture_output=True, text=True, timeout=120)
        return name, method, r.returncode, r.stdout[-400:]

    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(run, names))
    wrong = []
    refused = 0
    for name, method, rc, tail in results:
        if out_of_scope(name, method):
            refused += 1
            if rc == 0 or "-44" not in tail:
                wrong.append((name, rc, "expected a refusal with -44"))
        elif rc != 0:
            wrong.append((name, rc, tail[-200:]))
    assert not wrong, wrong[:10]
    assert refused == 0 and len(results) == 3192
    shutil.rmtree(work, ignore_errors=True)
