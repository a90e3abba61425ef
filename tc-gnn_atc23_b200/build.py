#!/usr/bin/env python3
"""In-tree build of the native code (no setup.py, no JIT cache):

  csrc/*.cu, csrc/sgt_cpu.cpp  --nvcc/g++-->  build/*.o  -->  libtcgnn_b200.so   (C ABI, include/tcgnn_b200.h)
  csrc/binding.cpp             --g++------->  TCGNN.cpython-*.so                  (torch extension module `TCGNN`)

Everything is compiled for sm_100a only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`).
nvcc cross-compiles without a GPU.  Outputs stay inside this directory so they travel to the
GPU box with the repository snapshot.  Re-runs only rebuild what is out of date.

    python tc-gnn_atc23_b200/build.py [--force] [--no-binding] [--verbose] [--debug-switches]

`--debug-switches` compiles the profiling switches in (-DTCGNN_DEBUG_SWITCHES: TCGNN_ABLATE / TCGNN_TRACE /
TCGNN_PRESET, used by tools/trace.py); the production library does not contain them -- some of them produce wrong
results on purpose.  Switching between the two kinds of build needs --force.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
INCLUDE = os.path.join(ROOT, "include")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

LIB_NAME = "libtcgnn_b200.so"
CU_SOURCES = ["capi.cu", "plan.cu", "round_pack.cu", "spmm_tc.cu", "sddmm_tc.cu", "sgt_gpu.cu", "graph_ops.cu",
              "host_entry.cu", "umma_probe.cu"]
CPP_SOURCES = ["sgt_cpu.cpp"]
HEADERS = ["common.cuh", "plan.h", "scan.cuh", os.path.join(INCLUDE, "tcgnn_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE,
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _run(cmd: list[str], verbose: bool) -> str:
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if verbose:
        print(res.stdout)
    return res.stdout


def ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX")


def lib_path() -> str:
    return os.path.join(HERE, LIB_NAME)


def module_path() -> str:
    return os.path.join(HERE, "TCGNN" + ext_suffix())


def build_library(force: bool = False, verbose: bool = False, debug_switches: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-DTCGNN_DEBUG_SWITCHES"] if debug_switches else []
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    objects = []
    for src in CU_SOURCES + CPP_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objects.append(o)
        if not force and _newer(o, [s] + headers):
            continue
        if src.endswith(".cu"):
            cmd = [NVCC] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
        else:
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-I", INCLUDE, "-c", s, "-o", o]
        jobs.append((src, cmd))
    logs = {}
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            futs = {pool.submit(_run, cmd, False): src for src, cmd in jobs}
            for f in cf.as_completed(futs):
                logs[futs[f]] = f.result()
        with open(os.path.join(OBJ, "ptxas.log"), "a") as fh:
            for src, out in logs.items():
                fh.write(f"==== {src}\n{out}\n")
                if verbose:
                    print(f"==== {src}\n{out}")
    lib = lib_path()
    if force or jobs or not _newer(lib, objects):
        _run([NVCC, "-shared", "-o", lib] + objects +
             ["-cudart", "shared", "-Xlinker", f"-rpath={os.path.join(CUDA_HOME, 'lib64')}", "-lpthread"], verbose)
    return lib


def build_binding(force: bool = False, verbose: bool = False) -> str:
    import torch  # noqa: F401  (only needed for its headers / libraries)
    tdir = os.path.dirname(torch.__file__)
    src = os.path.join(CSRC, "binding.cpp")
    out = module_path()
    if not force and _newer(out, [src, os.path.join(INCLUDE, "tcgnn_b200.h"), lib_path()]):
        return out
    pyinc = sysconfig.get_paths()["include"]
    cmd = [
        "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
        "-DTORCH_EXTENSION_NAME=TCGNN", "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1",
        "-I", INCLUDE, "-I", os.path.join(tdir, "include"),
        "-I", os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
        "-I", pyinc, "-I", os.path.join(CUDA_HOME, "include"),
        src, "-o", out,
        "-L", HERE, "-l:" + LIB_NAME, "-L", os.path.join(tdir, "lib"),
        "-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-ltorch_cuda",
        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{os.path.join(tdir, 'lib')}",
    ]
    _run(cmd, verbose)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--no-binding", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--debug-switches", action="store_true")
    args = ap.parse_args()
    lib = build_library(args.force or args.debug_switches, args.verbose, args.debug_switches)
    print("built", lib)
    if not args.no_binding:
        print("built", build_binding(args.force, args.verbose))


if __name__ == "__main__":
    main()
