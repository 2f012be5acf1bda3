#!/usr/bin/env python3
"""Build experimental variants of the library next to the product build (kernel tuning runs; see tools/quick_perf.py).
usage: python tools/build_variants.py name=DEFINE[,DEFINE...] ...   e.g.  p6=BRO_PARSE_MIN_BLOCKS=6"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brotli_rs_b200 import build as b

for spec in sys.argv[1:]:
    name, defs = spec.split("=", 1)
    print(b.build(defines=defs.split(","), out="lib_%s.so" % name))
