"""The oracle against the LIVE reference on fresh random inputs (beyond the committed golden fixtures).

Runs only where the reference checkout exists (the build container: /root/reference, imported in place, read-only,
comet_ml stubbed); skipped everywhere else -- the GPU box has no reference and nothing under `-m gpu` depends on it.
Every comparison is exact: the oracle runs the same torch-CPU / NumPy arithmetic in the same order."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import al_oracle as O

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, GOLD)
    import make_golden as G
    import make_golden_balancing as GB
    import make_golden_mase as GM
    get_strategy, Exp = G._import_reference()
    return dict(get_strategy=get_strategy, Exp=Exp, G=G, GM=GM, GB=GB)


def pool(rng, n, n_eval, n_lab):
    ev = rng.choice(n, size=n_eval, replace=False)
    rest = np.setdiff1d(np.arange(n), ev)
    lab = rng.choice(rest, size=n_lab, replace=False)
    lb = np.zeros(n, dtype=bool)
    lb[lab] = True
    return ev, lab, lb


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_margin_query(ref, seed):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    n, c = 300 + 17 * seed, (10, 100, 37)[seed]
    ev, lab, lb = pool(rng, n, 20, 40)
    logits = torch.randn(n, c) * 3
    s = ref["G"].make_strategy(ref["get_strategy"], ref["Exp"], "MarginSampler", logits, torch.zeros(n, 4), ev, lab, 64)
    idx, cost = s.query(50.0)
    avail = O.available_query_idxs(lb, ev, shuffle=False)
    mine, mcost = O.uncertainty_query(logits[avail], avail, 50.0, O.MODE_MARGIN)
    assert mine == idx and mcost == cost


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pairwise_and_both_coreset_modes(ref, seed):
    rng = np.random.default_rng(10 + seed)
    torch.manual_seed(10 + seed)
    cs = ref["get_strategy"]("CoresetSampler")
    obj = cs.__new__(cs)
    m, d = 200 + 31 * seed, (16, 48, 7)[seed]
    feat = torch.relu(torch.randn(m, d))
    ind = np.zeros(m, dtype=bool)
    ind[rng.choice(m, size=25, replace=False)] = True
    d2 = obj.get_pairwise_l2_dist(feat)
    assert torch.equal(O.pairwise_l2_dist(feat), d2)
    assert O.coreset(d2, ind, 30, randomize=False) == list(obj.coreset(d2, ind, 30, randomize=False))
    assert O.coreset_streaming(feat, ind, 30) == list(obj.coreset(d2, ind, 30, randomize=False))
    np.random.seed(seed)
    want = list(obj.coreset(d2, ind, 30, randomize=True))
    np.random.seed(seed)
    assert O.coreset(d2, ind, 30, randomize=True) == want
    np.random.seed(seed)                                      # the O(N)-memory form with NumPy-exact sampling arithmetic
    assert O.coreset_streaming(feat, ind, 30, randomize=True) == want
    none = np.zeros(m, dtype=bool)                            # cold start: minimax centre / uniform draw
    assert O.coreset(d2, none, 4, randomize=False) == list(obj.coreset(d2, none, 4, randomize=False))


@pytest.mark.parametrize("seed", [0, 1])
def test_gradient_embeddings(ref, seed):
    torch.manual_seed(20 + seed)
    n, c, d, bs = (45, 70)[seed], (10, 24)[seed], (12, 40)[seed], (16, 32)[seed]
    lg, h = torch.randn(n, c) * 3, torch.relu(torch.randn(n, d))
    s = ref["G"].make_strategy(ref["get_strategy"], ref["Exp"], "BADGESampler", lg, h, [], [], bs)
    for pooled in (False, True):
        want = s.get_gradient_embeddings(list(range(n)), use_adaptive_pool=pooled)
        assert torch.equal(O.gradient_embeddings(lg, h, bs, use_adaptive_pool=pooled), want)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mase_and_base(ref, seed):
    rng = np.random.default_rng(30 + seed)
    torch.manual_seed(30 + seed)
    n, m, c, bs, budget = 260, (8, 24, 16)[seed], (5, 12, 40)[seed], 64, 36
    ev, lab, lb = pool(rng, n, 15, 30)
    # + rand: no all-zero rows -> no exactly equal margins (the reference's torch.sort is not stable on ties)
    emb, w, b = torch.relu(torch.randn(n, m)) + 0.05 * torch.rand(n, m), torch.randn(c, m) * 0.5, torch.randn(c) * 0.1
    avail = O.available_query_idxs(lb, ev, shuffle=False)
    picks = {}
    for name in ("MASESampler", "BASESampler"):
        s = ref["GM"].make_strategy(ref["get_strategy"], ref["Exp"], name, emb, w, b, ev, lab, bs)
        with ref["GM"]._cuda_is_identity():
            picks[name] = s.query(float(budget))[0]
            mm, pc, pred, _ = s.compute_margins(avail)
    omm, opc, opred = O.mase_margins(emb[avail], w, b, bs)
    assert torch.equal(omm, mm) and torch.equal(opc, pc) and torch.equal(opred, pred)
    assert O.mase_query(omm, avail, float(budget))[0] == picks["MASESampler"]
    assert avail[O.base_select(omm, opc, opred, budget, c)].tolist() == picks["BASESampler"]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_balancing(ref, seed):
    rng = np.random.default_rng(40 + seed)
    torch.manual_seed(40 + seed)
    n, m, c, budget = 240, 10, (4, 6, 3)[seed], (30, 24, 40)[seed]
    probs = np.array([0.5, 0.25, 0.15, 0.05, 0.03, 0.02][:c])
    ys = rng.choice(c, size=n, p=probs / probs.sum())
    emb = torch.relu(torch.randn(c, m)[torch.from_numpy(ys)] * 2 + torch.randn(n, m))
    ev, lab, lb = pool(rng, n, 12, (40, 25, 60)[seed])
    idx, cost = ref["GB"].run_reference(ref["get_strategy"], ref["Exp"], emb, ys, c, ev, lab, 64, budget, seed)
    avail = ~lb
    avail[ev] = False
    np.random.seed(seed)
    mine, mcost = O.balancing_query(emb, torch.from_numpy(ys), avail, lb, float(budget), c)
    assert mine == idx and mcost == cost
