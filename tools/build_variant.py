#!/usr/bin/env python
"""Build an experiment copy of the library with extra -D flags (A/B timing; never the product library):

    python tools/build_variant.py <tag> -DMSDA_BWD_UNCOND=1 ...   ->  mdqe_cvpr2023_b200/libmsda_b200_<tag>.so
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200 import build as b  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(b.PKG_DIR, f"libmsda_b200_{tag}.so")
print(b.build(force=True, extra_flags=flags, lib_path=out, obj_dir=os.path.join(b.PKG_DIR, "_obj_" + tag)))
