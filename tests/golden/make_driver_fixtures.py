#!/usr/bin/env python
"""Copies the DATA the reference's own test driver needs for its double-precision regression
configurations into tests/golden/driver/ (the reference tree is not on the GPU box): the matrix
tests/LUNDA.mtx, the configurations tests/tests/test_00N with the `_double` suffix appended to the
solution-file name exactly as tests/Makefile:108-119 does (sed 's/sol_[^ ]*/&_double/'), and the
stored solutions tests/tests/sol_00N_double that check_solution (tests/COMMON/ioandtest.c:86-150)
compares against.  No source code is copied.  Run in the build container:
    python tests/golden/make_driver_fixtures.py [/root/reference]"""
import os
import re
import shutil
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "driver")
os.makedirs(os.path.join(out, "tests"), exist_ok=True)
shutil.copyfile(os.path.join(ref, "tests", "LUNDA.mtx"), os.path.join(out, "LUNDA.mtx"))
for i in range(1, 8):
    name = f"test_{i:03d}"
    txt = open(os.path.join(ref, "tests", "tests", name)).read()
    open(os.path.join(out, name), "w").write(re.sub(r"(sol_[^ \n]*)", r"\1_double", txt))
    shutil.copyfile(os.path.join(ref, "tests", "tests", f"sol_{i:03d}_double"),
                    os.path.join(out, "tests", f"sol_{i:03d}_double"))
# complex Hermitian configurations test_101...106 (mhd1280b.mtx, sol_10N_doublecomplex)
shutil.copyfile(os.path.join(ref, "tests", "mhd1280b.mtx"), os.path.join(out, "mhd1280b.mtx"))
for i in range(101, 107):
    name = f"test_{i:03d}"
    txt = open(os.path.join(ref, "tests", "tests", name)).read()
    open(os.path.join(out, name), "w").write(re.sub(r"(sol_[^ \n]*)", r"\1_doublecomplex", txt))
    shutil.copyfile(os.path.join(ref, "tests", "tests", f"sol_{i:03d}_doublecomplex"),
                    os.path.join(out, "tests", f"sol_{i:03d}_doublecomplex"))
print("wrote", sorted(os.listdir(out)))
