#!/usr/bin/env python
"""Run a script (bench.py, a micro-benchmark, pytest) against the EXPERIMENTS build of the library:

    TAG_TC_HALO=1 python tools/run_exp.py bench.py --steps 5 --no-cpu-baseline
    TAG_TC_DEBUG=4 python tools/run_exp.py tools/tc_microbench.py

libtag_b200_exp.so (build.py --experiments, -DTAG_EXPERIMENTS) is the only build that reads the TAG_TC_DEBUG / TAG_TC_HALO /
TAG_TC_PAIR / TAG_K1_DEBUG / TAG_FRAME_TABLE switches; results can be WRONG with them set, so the product library has none."""
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("video-gen-evals_b200.build")
lib = os.path.join(os.path.dirname(b.__file__), "libtag_b200_exp.so")
if not os.path.exists(lib):
    lib = b.build(experiments=True)
L = importlib.import_module("video-gen-evals_b200._lib")
L.LIB_PATH = lib
if len(sys.argv) < 2:
    sys.exit("usage: run_exp.py <script.py | -m module> [args...]")
if sys.argv[1] == "-m":
    sys.argv = sys.argv[2:]
    runpy.run_module(sys.argv[0], run_name="__main__", alter_sys=True)
else:
    sys.argv = sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
