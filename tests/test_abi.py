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


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/include"), reason="reference tree not mounted")
def test_internal_context_mirror_matches_reference(tmp_path):
    """primme_context crosses the boundary BY VALUE in the reference-internal entry points the
    reference's test driver links (include/primme_ref_internal.h, primme_b200/src/ref_internal.c):
    size and field offsets equal those of src/include/common.h:610-641"""
    import subprocess
    src = tmp_path / "c.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "common.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(primme_context), offsetof(primme_context, primme),'
                   'offsetof(primme_context, printLevel), offsetof(primme_context, report), offsetof(primme_context, mm),'
                   'offsetof(primme_context, numProcs), offsetof(primme_context, bcast), offsetof(primme_context, globalSum),'
                   'offsetof(primme_context, queue));\nprintf("%zu %zu %zu\\n", sizeof(primme_frame), offsetof(primme_frame, keep_frame),'
                   'offsetof(primme_frame, prev));return 0;}\n')
    exe = tmp_path / "c"
    subprocess.run(["gcc", "-DNDEBUG", "-I/root/reference/src/include", "-I/root/reference/include", str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    ref = [int(x) for x in out]
    lib = api.load_library()
    mine = [lib.primme_b200_ref_context_size()] + [lib.primme_b200_ref_context_offset(i) for i in range(8)]
    assert mine == ref[:9]
    assert ref[9:] == [24, 8, 16]  # primme_frame {prev_alloc, keep_frame, prev}
