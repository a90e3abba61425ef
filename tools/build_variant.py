#!/usr/bin/env python3
"""Build the C-ABI library with extra -D defines into variants/<name>/libtcgnn_b200.so (A/B experiments on the GPU box:
`LD_LIBRARY_PATH=variants/<name> python tools/quick.py ...` -- the binding finds the library through RUNPATH, which
LD_LIBRARY_PATH precedes).  The products are ignored by git (*.so, *.o) and travel with gpurun.

    python tools/build_variant.py trace -DTCGNN_DEBUG_SWITCHES
"""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import build as b  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "variants", name)
os.makedirs(os.path.join(out, "obj"), exist_ok=True)
jobs, objects = [], []
for src in b.CU_SOURCES + b.CPP_SOURCES:
    s = os.path.join(b.CSRC, src)
    o = os.path.join(out, "obj", os.path.splitext(src)[0] + ".o")
    objects.append(o)
    if src.endswith(".cu"):
        jobs.append([b.NVCC] + b.NVCC_FLAGS + extra + ["-c", s, "-o", o])
    else:
        jobs.append(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-I", b.INCLUDE, "-c", s, "-o", o])
with cf.ThreadPoolExecutor(max_workers=8) as pool:
    for r in pool.map(lambda c: subprocess.run(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), jobs):
        if r.returncode != 0:
            sys.exit(r.stdout)
lib = os.path.join(out, b.LIB_NAME)
subprocess.run([b.NVCC, "-shared", "-o", lib] + objects + ["-cudart", "shared", "-Xlinker",
               f"-rpath={os.path.join(b.CUDA_HOME, 'lib64')}", "-lpthread"], check=True)
print("built", lib)
