"""Generate tests/golden/solver_golden.json by running the UNMODIFIED reference
(oracle/_ref/libprimme_ref.so, built from /root/reference by oracle/Makefile) on the cases of
tests/golden/cases.py.  Run in the build container:  python tests/golden/make_golden.py
The JSON is committed; the GPU box has no /root/reference and only reads the fixture."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness as H  # noqa: E402
from golden.cases import CASES, MATRICES  # noqa: E402

out = {}
for name, (mat, k, kw, exact) in CASES.items():
    csr = MATRICES[mat]()
    r = H.solve("reference", csr, k, **kw)
    s = r["stats"]
    out[name] = dict(matrix=mat, numEvals=k, ret=r["ret"], initSize=r["initSize"],
                     evals=[float(v) for v in r["evals"]], rnorms=[float(v) for v in r["rnorms"]],
                     numOuterIterations=s["numOuterIterations"], numRestarts=s["numRestarts"],
                     numMatvecs=s["numMatvecs"], numPreconds=s["numPreconds"], exact_counts=exact,
                     eps=kw.get("eps", 0.0), aNorm=kw.get("aNorm", 0.0),
                     estimateLargestSVal=s["estimateLargestSVal"])
    print(name, out[name]["numOuterIterations"], out[name]["numRestarts"], out[name]["numMatvecs"])
with open(os.path.join(HERE, "solver_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
