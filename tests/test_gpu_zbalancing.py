"""BalancingSampler on the GPU (libalq.so through ctypes): the masked ratio arg-min against torch, and the drop-in
sampler end to end against the picks the reference itself produced."""
import numpy as np
import pytest
import torch

from test_balancing import balancing_strategy, bgold  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from active_learning_b200.engine import Engine
    return Engine()


@pytest.mark.parametrize("n", [1, 50, 1024, 100003])
def test_ratio_argmin_matches_torch(eng, n):
    torch.manual_seed(n)
    num = torch.rand(n) + 0.1
    den = torch.rand(n) + 0.1
    if n > 50:                                            # quantised: exact ties, the lowest available index wins
        num, den = torch.round(num * 4) / 4 + 0.25, torch.round(den * 4) / 4 + 0.25
    avail = (torch.rand(n) < 0.6).to(torch.uint8)
    avail[n // 2] = 1
    for nm in (num, None):
        r = (torch.ones(n) if nm is None else nm) / den
        want = int(torch.where(avail.bool(), r, torch.tensor(float("inf"))).min(dim=0).indices)
        got = eng.ratio_argmin(None if nm is None else nm.cuda(), den.cuda(), avail.cuda())
        assert got == want and avail[got] == 1
    assert eng.ratio_argmin(num.cuda(), den.cuda(), torch.zeros(n, dtype=torch.uint8).cuda()) == -1


@pytest.mark.parametrize("tag", ["a", "b"])
def test_balancing_sampler_end_to_end(bgold, tag):  # noqa: F811
    g = bgold
    s = balancing_strategy(g, tag)
    np.random.seed(11)
    idx, cost = s.query(float(g[f"{tag}_budget"]))
    assert [int(i) for i in idx] == g[f"{tag}_picks"].tolist() and cost == int(g[f"{tag}_budget"])
    s.update(idx, cost)
    assert s._bal_cache[0].is_cuda                        # embeddings stayed on the device
