import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def disorder():
    out = {}
    for L in (4, 20):
        out[L] = (pd.read_csv(os.path.join(GOLDEN, f"hs_L{L}.csv")).values,
                  pd.read_csv(os.path.join(GOLDEN, f"phis_L{L}.csv")).values)
    return out


@pytest.fixture(scope="session")
def known():
    with open(os.path.join(GOLDEN, "known_answers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def gate_counts():
    with open(os.path.join(GOLDEN, "gate_counts.json")) as fh:
        return json.load(fh)


def golden_csv(name):
    return pd.read_csv(os.path.join(GOLDEN, name))


def family_z(n_points, alpha=0.0027):
    """Per-point |z| threshold that keeps the FAMILY-WISE false-alarm rate of `n_points` independent comparisons
    at the single-comparison 3-sigma level (two-sided alpha = 0.0027, Bonferroni): BASELINE.json's north_star asks
    for estimates 'within 3-sigma'; checking N points each at 3.0 sigma would reject a correct simulator with
    probability 1 - 0.9973^N (11 % at N = 43), so the per-point bound is widened to keep the family at 0.27 %.
    n_points = 1 gives exactly 3.0."""
    from scipy.stats import norm
    return float(norm.isf(alpha / (2.0 * n_points)))
