#!/usr/bin/env python
"""one-off helper of the zprimme port: retype the listed identifiers from double to SCALAR in a C file
(declarations, parameter lists and the malloc/calloc statements that initialise them)"""
import re
import sys

path, names = sys.argv[1], sys.argv[2:]
s = open(path).read()
alt = "|".join(re.escape(n) for n in names)
# declarations / parameters:  [const] double *NAME   and  , *NAME in the same declaration is left to the compiler
s = re.sub(r"\bdouble (\*+)(%s)\b" % alt, r"SCALAR \1\2", s)
# allocation statements assigning to NAME
def fix_alloc(m):
    st = m.group(0)
    return st.replace("(double *)", "(SCALAR *)").replace("sizeof(double)", "sizeof(SCALAR)")
s = re.sub(r"\b(?:%s) = \(double \*\)(?:malloc|calloc)\([^;]*;" % alt, fix_alloc, s)
open(path, "w").write(s)
