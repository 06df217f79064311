"""Shared helpers for the test-suite: golden fixtures and synthetic problems."""
import json
import os

import numpy as np

from bam_b200.capi import Problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def prior_corr_to_M(C):
    C = np.asarray(C, dtype=np.float64)
    return np.linalg.inv(C) - np.eye(C.shape[0])


def problem_from_fixture(name, with_copula=True):
    d = load_fixture(name)
    dd = dict(d)
    if d.get("priorC") is not None and with_copula:
        k = int(round(len(d["priorC"]) ** 0.5))
        C = np.asarray(d["priorC"]).reshape(k, k).T  # column-major -> matrix
        dd["priorM"] = prior_corr_to_M(C).flatten(order="F").tolist()
    else:
        dd["priorM"] = None
    return Problem.from_dict(dd), d


def perturbed_vectors(start, K, rel=0.02, seed=0, abs_floor=0.01):
    """K vectors around `start` (row 0 is start itself)."""
    rng = np.random.default_rng(seed)
    start = np.asarray(start, dtype=np.float64)
    x = np.tile(start, (K, 1))
    scale = np.maximum(np.abs(start) * rel, abs_floor * rel)
    x[1:] += rng.standard_normal((K - 1, start.size)) * scale
    return x


def relerr(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
