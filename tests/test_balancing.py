"""BalancingSampler: the oracle's restatement against the picks the reference itself
produced (tests/golden/make_golden_balancing.py), and the drop-in sampler's host logic with the OracleEngine.  CPU."""
import os
import pickle

import numpy as np
import pytest
import torch

from helpers import LabeledIndexDataset, OracleEngine, make_strategy
from oracle import al_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_balancing.npz")))


def masks(g, tag):
    n = g[f"{tag}_emb"].shape[0]
    lb = np.zeros(n, dtype=bool)
    lb[g[f"{tag}_labeled"]] = True
    avail = ~lb
    avail[g[f"{tag}_eval"]] = False
    return avail, lb


def balancing_strategy(g, tag, engine=None, **kw):
    emb, ys, c = torch.from_numpy(g[f"{tag}_emb"]), g[f"{tag}_ys"], int(g[f"{tag}_classes"])
    return make_strategy("BalancingSampler", torch.zeros(len(ys), c), emb, g[f"{tag}_eval"], g[f"{tag}_labeled"], 64,
                         engine=engine, dataset=LabeledIndexDataset(ys, c), **kw)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference(bgold, tag):
    g = bgold
    avail, lb = masks(g, tag)
    diag = {}
    np.random.seed(11)
    picks, cost = O.balancing_query(torch.from_numpy(g[f"{tag}_emb"]), torch.from_numpy(g[f"{tag}_ys"]), avail, lb,
                                    float(g[f"{tag}_budget"]), int(g[f"{tag}_classes"]), diag)
    assert picks == g[f"{tag}_picks"].tolist() and cost == int(g[f"{tag}_budget"])
    assert diag["balancing_steps"] == int(g[f"{tag}_balancing_steps"])
    if tag == "b":                                     # both branches of balancing_sampler.py:81-124 are exercised
        assert 0 < diag["balancing_steps"] < cost


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sampler_plumbing_matches_reference(bgold, tag):
    g = bgold
    s = balancing_strategy(g, tag, engine=OracleEngine())
    np.random.seed(11)
    idx, cost = s.query(float(g[f"{tag}_budget"]))
    assert [int(i) for i in idx] == g[f"{tag}_picks"].tolist() and cost == int(g[f"{tag}_budget"])
    after = np.random.random_sample()
    avail, lb = masks(g, tag)                           # the RNG stream was consumed exactly like the reference's
    np.random.seed(11)
    O.balancing_query(torch.from_numpy(g[f"{tag}_emb"]), torch.from_numpy(g[f"{tag}_ys"]), avail, lb,
                      float(g[f"{tag}_budget"]), int(g[f"{tag}_classes"]))
    assert after == np.random.random_sample()
    s.update(idx, cost)                                 # strategy.py:470: nothing labeled twice, nothing from eval
    assert not set(int(i) for i in idx) & set(g[f"{tag}_eval"].tolist())
    pickle.dumps(s)                                     # the embedding cache and the engine stay out of pickles
    # second round under --freeze_feature reuses the cached embeddings (balancing_sampler.py:34-37,55-57)
    cached = s._bal_cache[0]
    idx2, cost2 = s.query(5.0)
    assert s._bal_cache[0] is cached and cost2 == 5 and not set(map(int, idx2)) & set(map(int, idx))


def test_budget_is_clamped_to_the_available_rows(bgold):
    g = bgold
    s = balancing_strategy(g, "b", engine=OracleEngine())
    avail, _ = masks(g, "b")
    np.random.seed(3)
    idx, cost = s.query(1e9)
    assert cost == int(avail.sum()) == len(set(int(i) for i in idx))
