"""The public structs, labels and enum values are byte-compatible with the reference's headers
(reference include/primme_eigs.h:166-253,286-378; include/primme_svds.h).  The golden values were
recorded from /root/reference/include by tests/golden/make_abi_golden.py."""
import ctypes as C
import json
import os

import pytest

import abi_probe
from primme_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "abi_golden.json")))
OB = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"  # LAPACK of the build


def test_headers_match_reference_layout():
    mine = abi_probe.probe(os.path.join(ROOT, "include"))
    assert mine == GOLD


def test_ctypes_mirror_matches_headers():
    assert C.sizeof(api.PrimmeParams) == GOLD["sizeof_primme_params"]
    assert C.sizeof(api.PrimmeStats) == GOLD["sizeof_primme_stats"]
    for name in ("n", "nLocal", "numEvals", "iseed", "eps", "ldOPs", "queue", "profile", "matrix"):
        assert getattr(api.PrimmeParams, name).offset == GOLD["off_eigs_" + name], name
    assert api.PrimmeParams.stats.offset == GOLD["off_eigs_stats_numOuterIterations"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference headers not mounted")
def test_golden_is_current():
    assert abi_probe.probe("/root/reference/include") == GOLD


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference examples not mounted")
def test_reference_example_compiles_against_our_headers(tmp_path):
    """examples/ex_eigs_dseq.c (the plumbing config C1) compiles unchanged against include/ and links
    against the product library; running it needs a GPU (see tests/test_examples_gpu.py)."""
    import subprocess
    exe = tmp_path / "ex_eigs_dseq"
    lib = os.path.join(ROOT, "primme_b200")
    r = subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), "/root/reference/examples/ex_eigs_dseq.c",
                        "-o", str(exe), "-L", lib, "-lprimme_b200", f"-Wl,-rpath,{lib}",
                        f"-Wl,-rpath-link,{OB}", "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
