"""Shared test helpers: golden fixture access and error metrics."""
from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

import cases

GOLDEN = Path(__file__).resolve().parent / "golden"


@lru_cache(maxsize=None)
def _npz(name):
    return np.load(GOLDEN / name, allow_pickle=False)


def golden(case_name: str, dt: str, what: str):
    """-> (kind, array, float64 sum); kind is 'full' or 'sampled' (every SAMPLE_STRIDE-th)."""
    z = _npz("core.npz")
    key = f"{case_name}/{dt}/{what}"
    s = float(z[key + "/sum"])
    if key + "/full" in z.files:
        return "full", z[key + "/full"], s
    return "sampled", z[key + "/sampled"], s


def golden_insum(case_name: str) -> float:
    return float(_npz("core.npz")[f"{case_name}/insum"])


def module_golden():
    return _npz("module.npz")


def view_like_golden(kind: str, arr):
    a = np.asarray(arr)
    return a if kind == "full" else a.reshape(-1)[::cases.SAMPLE_STRIDE]


def max_abs(a, b) -> float:
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))) \
        if np.asarray(a).size else 0.0


def allclose_frac(a, b, rtol, atol) -> float:
    """Fraction of elements violating |a-b| <= atol + rtol*|b|."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.mean(np.abs(a - b) > atol + rtol * np.abs(b)))


def check_inputs_unchanged(c):
    arrs = cases.make_inputs(c)
    got = cases.input_checksum(arrs)
    want = golden_insum(c.name)
    assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), \
        f"input generator drifted for {c.name}: {got} vs {want}"
    return arrs
