"""Build the product: libbam_gpu.so (CUDA, sm_100a) and the bam_gpu host driver.

nvcc cross-compiles without a GPU.  Everything is built IN-TREE so that the binaries travel to
the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", "g++"]
LIB_SOURCES = ["bam_capi.cu", "bam_logpost.cu", "bam_logpost_chain.cu", "bam_logpost_tile.cu", "bam_sampler.cu", "bam_sampler_warp_a.cu", "bam_sampler_warp_b.cu",
               "bam_sampler_warp_c.cu", "bam_latent.cu", "bam_predict.cu", "bam_comm.cu", "bam_host.cpp"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "bam_gpu.h"))
    hostdir = os.path.join(CSRC, "host")
    if os.path.isdir(hostdir):
        out += [os.path.join(hostdir, f) for f in os.listdir(hostdir) if f.endswith(".h")]
    return out


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout + r.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in LIB_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _newer([s] + hdrs, o):
            extra = ["-Xptxas", "-v"] if verbose and src.endswith(".cu") else []
            jobs.append([NVCC] + NVFLAGS + extra + ["-c", s, "-o", o])
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), 10)) as ex:
            logs = list(ex.map(_run, jobs))
    so = os.path.join(HERE, "libbam_gpu.so")
    if jobs or not os.path.exists(so):
        _run([NVCC] + ARCH + ["-shared", "-o", so] + objs + ["-lcudart", "-ldl"])
    if verbose:
        print("\n".join(logs))
    return so


def build_driver(force: bool = False) -> str | None:
    hostdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hostdir):
        return None
    srcs = sorted(os.path.join(hostdir, f) for f in os.listdir(hostdir) if f.endswith(".cpp"))
    if not srcs or not os.path.exists(os.path.join(hostdir, "bam_gpu_main.cpp")):
        return None
    exe = os.path.join(HERE, "bam_gpu")
    so = os.path.join(HERE, "libbam_gpu.so")
    if force or _newer(srcs + _headers() + [so], exe):
        _run(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), "-I", CSRC, "-o", exe] + srcs +
             ["-L", HERE, "-lbam_gpu", "-pthread", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def build_all(force: bool = False, verbose: bool = False):
    so = build_library(force, verbose)
    exe = build_driver(force)
    return so, exe


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
