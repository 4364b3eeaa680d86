#!/usr/bin/env python
"""tests/golden/make_golden_tables.py -- constructor tables of the UNMODIFIED reference class
(/root/reference/src/orb_extractor/ORBextractor.cc:419-474 through its accessors, inc/ORBextractor.h:63-83),
printed by `oracle/_ref/ref_extract tables` (oracle/ref_main.cpp) and stored bit-exactly in ctor_tables.json.

    make -C oracle all && python tests/golden/make_golden_tables.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import CTOR_TABLE_CASES, parse_tables  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_extract")

out = {}
for c in CTOR_TABLE_CASES:
    r = subprocess.run([REF, "tables"] + [repr(x) if isinstance(x, float) else str(x) for x in c], capture_output=True, text=True, check=True)
    out["/".join(str(x) for x in c)] = parse_tables(r.stdout)
with open(os.path.join(HERE, "ctor_tables.json"), "w") as f:
    json.dump(out, f, sort_keys=True, separators=(",", ":"))
print("wrote %d cases" % len(out))
