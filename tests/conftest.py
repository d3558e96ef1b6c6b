import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_eval():
    return dict(np.load(os.path.join(GOLDEN, "kat_evaluator.npz")))


@pytest.fixture(scope="session")
def ml100k():
    """The reference's ratio-0.8 ml-100k split (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "ml100k_split.npz"))
    return {
        "num_users": int(z["num_users"]), "num_items": int(z["num_items"]),
        "train_indptr": z["train_indptr"].astype(np.int64),
        "train_indices": z["train_indices"].astype(np.int32),
        "test_indptr": z["test_indptr"].astype(np.int64),
        "test_indices": z["test_indices"].astype(np.int32),
    }


@pytest.fixture(scope="session")
def gowalla():
    """dataset/gowalla.{train,test} ('given' split) as loaded by the reference's data.Dataset
    (tests/golden/make_golden.py gowalla) + what the reference computed on it."""
    z = np.load(os.path.join(GOLDEN, "gowalla_split.npz"))
    with open(os.path.join(GOLDEN, "kat_gowalla.json")) as f:
        kat = json.load(f)
    return {
        "num_users": int(z["num_users"]), "num_items": int(z["num_items"]),
        "train_indptr": z["train_indptr"].astype(np.int64),
        "train_indices": z["train_indices"].astype(np.int32),
        "test_indptr": z["test_indptr"].astype(np.int64),
        "test_indices": z["test_indices"].astype(np.int32),
        "kat": kat, "adj": dict(np.load(os.path.join(GOLDEN, "kat_gowalla_adj.npz"))),
    }


def gowalla_tables(g):
    """The tables of make_golden.py::gowalla (RandomState(11); noise + mean of 3 held-out items)."""
    rng = np.random.RandomState(11)
    V = (rng.randn(g["num_items"], 64) * .1).astype(np.float32)
    U = (rng.randn(g["num_users"], 64) * .1).astype(np.float32)
    tp, ti = g["test_indptr"], g["test_indices"]
    for u in range(g["num_users"]):
        if tp[u + 1] > tp[u]:
            U[u] += V[ti[tp[u]:tp[u + 1]][:3]].mean(0) * np.float32(1.5)
    return U, V


@pytest.fixture(scope="session")
def golden_ml100k_eval():
    with open(os.path.join(GOLDEN, "kat_ml100k_eval.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_sampler():
    with open(os.path.join(GOLDEN, "kat_sampler.json")) as f:
        return json.load(f)


def parse_result_string(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def random_csr(rs, num_rows, num_cols, degrees):
    """Random CSR with sorted, duplicate-free rows of about the requested degrees (fast)."""
    rows = [np.unique(rs.randint(0, num_cols, int(k))) for k in degrees]
    indptr = np.zeros(num_rows + 1, np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)
