"""The C ABI from a plain C99 client (tests/c/abi_check.c, gcc -std=c99 -pedantic -Wall -Werror): the header compiles as C,
the struct layout is the one the Fortran bind(C) types of bindings/bam_gpu_binding.f90 imply, every export the Fortran
module declares exists in the library, and the whole path runs through the ABI (GPU) / fails loudly without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_check.c")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from bam_b200 import build

    build.build_library()
    out = str(tmp_path_factory.mktemp("abi") / "abi_check")
    lib = os.path.join(ROOT, "bam_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                        "-L", lib, "-lbam_gpu", "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def fortran_types():
    """{type name: [(field, kind)]} of the bind(C) derived types, in declaration order"""
    text = open(os.path.join(ROOT, "bindings", "bam_gpu_binding.f90")).read()
    out = {}
    for m in re.finditer(r"type, bind\(C\) :: (\w+)\n(.*?)end type", text, flags=re.S):
        fields = []
        for ln in m.group(2).splitlines():
            ln = ln.split("!")[0].strip()
            mm = re.match(r"(type\(c_ptr\)|integer\(c_int\)|real\(c_double\))\s*::\s*(.*)", ln)
            if mm:
                fields += [(f.strip(), mm.group(1)) for f in mm.group(2).split(",")]
        out[m.group(1)] = fields
    return out


def implied_layout(fields):
    """offsets a companion C compiler gives the interoperable struct (natural alignment, LP64)"""
    size = {"type(c_ptr)": 8, "integer(c_int)": 4, "real(c_double)": 8}
    off, o, amax = {}, 0, 1
    for name, kind in fields:
        s = size[kind]
        o = (o + s - 1) // s * s
        off[name] = o
        o += s
        amax = max(amax, s)
    return off, (o + amax - 1) // amax * amax


def test_layout_is_the_one_the_fortran_types_imply(exe):
    r = subprocess.run([exe, "layout"], capture_output=True, text=True)
    assert r.returncode == 0
    got = dict(ln.split() for ln in r.stdout.splitlines())
    ft = fortran_types()
    assert set(ft) == {"bamgpu_problem", "bamgpu_sizes"}
    for tname, fields in ft.items():
        off, size = implied_layout(fields)
        c_fields = [k.split(".")[1] for k in got if k.startswith(tname + ".") and not k.endswith(".sizeof")]
        assert [f for f, _ in fields] == c_fields, tname  # same members, same order
        for f, _ in fields:
            assert int(got[f"{tname}.{f}"]) == off[f], (tname, f)
        assert int(got[f"{tname}.sizeof"]) == size
    text = open(os.path.join(ROOT, "bindings", "bam_gpu_binding.f90")).read()
    for const in ("BAMGPU_MESS_LEN", "BAMGPU_PRIOR_NPAR", "BAMGPU_NENV"):
        assert re.search(rf"{const} = {got[const]}\b", text), const


def test_fortran_module_declares_every_export():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bam_gpu.h")).read(), flags=re.S)
    exports = set(re.findall(r"\b(bamgpu_[A-Za-z0-9_]+)\s*\(", hdr))
    bound = set(re.findall(r"bind\(C, name='(bamgpu_\w+)'\)", open(os.path.join(ROOT, "bindings", "bam_gpu_binding.f90")).read()))
    assert bound <= exports, bound - exports
    assert exports == bound, sorted(exports - bound)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_c_client_fails_loudly_without_a_gpu(exe):
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_client_runs_the_whole_path(exe):
    from bam_b200.capi import Problem
    from oracle.oracle_api import Oracle

    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().splitlines()[-1].startswith("OK")
    lp = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("logpost ")]
    assert len(lp) == 4
    H = [-0.3, -0.1, 0.1, 0.4, 0.8, 1.2, 1.6, 1.9, 2.3, 2.8, 3.4, 4.1]
    Q = [4.1, 12.9, 25.0, 47.8, 85.0, 130.2, 181.0, 260.4, 420.9, 700.3, 1105.0, 1712.0]
    prob = Problem("BaRatin", np.array(H), np.array(Q), Yu=0.05 * np.array(Q), partype=[0] * 6,
                   priors_theta=[("Gaussian", [-0.5, 1]), ("LogNormal", [3.9, 0.5]), ("Gaussian", [1.5, 0.05]),
                                 ("Gaussian", [1.5, 0.5]), ("LogNormal", [4.9, 0.5]), ("Gaussian", [1.67, 0.05])],
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000]), ("Uniform", [0, 1000])],
                   xtra_i=np.array([1, 0, 0, 1], dtype=np.int32))
    x0 = np.array([-0.5, 53.0, 1.5, 1.5, 144.0, 1.67, 1.0, 0.1])
    x = np.array([x0 * (1.0 + 0.01 * k * np.where(np.arange(8) % 2, 1, -1)) for k in range(4)])
    x[3, 3] = -1.0
    fx, feas, isnull = Oracle(prob).logpost(x)
    for k, row in enumerate(lp):
        assert int(row[3]) == feas[k] and int(row[4]) == isnull[k]
        if feas[k]:
            assert abs(float(row[2]) - fx[k]) <= 1e-12 * abs(fx[k])
    assert not feas[3]
