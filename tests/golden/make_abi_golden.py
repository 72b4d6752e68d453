"""Record sizeof/offsetof of the public structs from the REFERENCE headers
(/root/reference/include) into tests/golden/abi_golden.json (committed)."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from abi_probe import probe  # noqa: E402

out = probe("/root/reference/include")
with open(os.path.join(HERE, "abi_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(out), "entries")
