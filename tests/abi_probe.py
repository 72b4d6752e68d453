"""Compile a probe against a set of PRIMME headers and return {name: value} of sizes, offsets and
enum values (used by tests/test_abi.py and tests/golden/make_abi_golden.py)."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def field_names():
    """member paths from our X-macro tables (the reference has the same members)"""
    paths = {"primme_params": [], "primme_svds_params": []}
    for hdr, key in (("primme_eigs.h", "primme_params"), ("primme_svds.h", "primme_svds_params")):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        for m in re.finditer(r"X\((\w+),\s*(\d+),\s*([\w\.]+),\s*(\w+)\)", txt):
            paths[key].append((m.group(1), int(m.group(2)), m.group(3)))
    return paths


def probe(incdir):
    paths = field_names()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "primme.h"', 'int main(void){']
    lines.append('printf("sizeof_primme_params %zu\\n", sizeof(primme_params));')
    lines.append('printf("sizeof_primme_svds_params %zu\\n", sizeof(primme_svds_params));')
    lines.append('printf("sizeof_primme_stats %zu\\n", sizeof(primme_stats));')
    for name, ident, path in paths["primme_params"]:
        lines.append(f'printf("off_eigs_{name} %zu\\n", offsetof(primme_params, {path}));')
        lines.append(f'printf("label_eigs_{name} %d\\n", (int)PRIMME_{name});')
    for name, ident, path in paths["primme_svds_params"]:
        lines.append(f'printf("off_svds_{name} %zu\\n", offsetof(primme_svds_params, {path}));')
        lines.append(f'printf("label_svds_{name} %d\\n", (int)PRIMME_SVDS_{name});')
    for e in ("primme_largest_abs", "primme_proj_refined", "primme_init_user", "primme_adaptive",
              "primme_event_profile", "primme_orth_explicit_I", "primme_op_int", "PRIMME_LOBPCG_OrthoBasis_Window",
              "primme_string", "primme_svds_closest_abs", "primme_svds_augmented", "primme_svds_op_augmented",
              "PRIMME_FUNCTION_UNAVAILABLE", "PRIMME_LAPACK_FAILURE", "PRIMME_MAIN_ITER_FAILURE"):
        lines.append(f'printf("enum_{e} %d\\n", (int){e});')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write("\n".join(lines))
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-I", incdir, src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return {ln.split()[0]: int(ln.split()[1]) for ln in out.strip().splitlines()}
