"""CPU: the C-ABI library loads and exports every symbol include/bam_gpu.h declares; without a GPU
the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bam_b200 import build

    build.build_library()
    from bam_b200.capi import load_library

    return load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bam_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bamgpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_name_helpers(lib):
    assert lib.bamgpu_dist_id(b"Gaussian") == 1 and lib.bamgpu_dist_id(b"FlatPrior+") == 7
    assert lib.bamgpu_dist_id(b"Nope") == 0
    assert lib.bamgpu_sigmafunc_id(b"Linear") == 2 and lib.bamgpu_sigmafunc_npar(4) == 3
    assert lib.bamgpu_model_id(b"MDL_BaRatin") == 2 and lib.bamgpu_model_id(b"BaRatin") == 2


def test_prior_corr_to_M_and_rhat(lib):
    from bam_b200.capi import _ptr_d

    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    C = A @ A.T
    dd = np.sqrt(np.diag(C))
    C = C / dd[:, None] / dd[None, :]
    np.fill_diagonal(C, 1.0)
    Cf = np.asfortranarray(C)
    M = np.zeros((6, 6), order="F")
    mess = ctypes.create_string_buffer(250)
    assert lib.bamgpu_prior_corr_to_M(6, _ptr_d(Cf), _ptr_d(M), mess) == 0
    assert np.allclose(M, np.linalg.inv(C) - np.eye(6), rtol=1e-10, atol=1e-10)
    bad = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 1.0]]))
    M2 = np.zeros((2, 2), order="F")
    assert lib.bamgpu_prior_corr_to_M(2, _ptr_d(bad), _ptr_d(M2), mess) == 1 and b"positive definite" in mess.value
    # Gelman-Rubin on synthetic chains
    K, P, n = 8, 3, 500
    chains = rng.standard_normal((K, n, P)) + np.array([0.0, 0.0, 1.0]) * np.arange(K)[:, None, None]
    count = np.full(K, float(n))
    mean = chains.mean(axis=1)
    m2 = ((chains - mean[:, None, :]) ** 2).sum(axis=1)
    out = np.zeros(P)
    assert lib.bamgpu_rhat(K, P, _ptr_d(count), _ptr_d(np.ascontiguousarray(mean)), _ptr_d(np.ascontiguousarray(m2)),
                           _ptr_d(out), mess) == 0
    assert out[0] < 1.02 and out[1] < 1.02 and out[2] > 2.0


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly(lib):
    from bam_b200.capi import BamError, BamGpu
    from tests.helpers import problem_from_fixture

    p, _ = problem_from_fixture("baratin")
    with pytest.raises(BamError) as e:
        BamGpu(p)
    assert "no CUDA device" in str(e.value)


def test_host_layout_errors(lib):
    """create() reports the reference's error convention (err>0 + message) for bad problems"""
    from bam_b200.capi import BamError, BamGpu, Problem

    X, Y = np.zeros((4, 1)), np.zeros((4, 1))
    bad = Problem("NoSuchModel", X, Y, partype=[0], priors_theta=[("Gaussian", [0, 1])], sigma_func=["Constant"],
                  priors_gamma=[("Uniform", [0, 1])])
    with pytest.raises(BamError) as e:
        BamGpu(bad)
    assert "Unavailable [model%ID]" in str(e.value) or "no CUDA device" in str(e.value)
