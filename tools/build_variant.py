#!/usr/bin/env python
"""Build a named experiment variant of libp3p.so next to the product library (CPU box: nvcc cross-compiles).
usage: python tools/build_variant.py NAME [nvcc flags ...]   ->  pixelspointspolygons_b200/variants/libp3p_NAME.so
Select it on the GPU box with P3P_LIB=<path> (pixelspointspolygons_b200/build.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelspointspolygons_b200 import build

name, flags = sys.argv[1], sys.argv[2:]
d = os.path.join(ROOT, "pixelspointspolygons_b200", "variants")
os.makedirs(d, exist_ok=True)
print(build.build_library(force=True, out=os.path.join(d, f"libp3p_{name}.so"), extra_flags=flags))
