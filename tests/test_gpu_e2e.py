"""The drop-in samplers end to end on the GPU (libalq.so through ctypes) against the index lists
the reference itself produced (tests/golden/reference_golden.npz)."""
import pickle

import numpy as np
import pytest
import torch

from helpers import make_strategy

pytestmark = pytest.mark.gpu


def _pool(gold):
    return int(gold["e2e_n"]), gold["e2e_eval_idxs"], gold["e2e_labeled"]


def test_margin_confidence_entropy_samplers(gold):
    n, ev, lab = _pool(gold)
    for tag in ("f32_c10", "f32_c1000"):
        s = make_strategy("MarginSampler", torch.from_numpy(gold[f"margin_{tag}_logits"]),
                          torch.zeros(n, 4), ev, lab, 128)
        np.random.seed(7)
        idx, cost = s.query(60.0)
        assert idx == gold[f"margin_{tag}_picks"].tolist() and cost == 60
    s = make_strategy("ConfidenceSampler", torch.from_numpy(gold["confidence_f32_logits"]),
                      torch.zeros(n, 4), ev, lab, 128)
    np.random.seed(7)
    assert s.query(60.0)[0] == gold["confidence_f32_picks"].tolist()
    logits = torch.from_numpy(gold["margin_f32_c1000_logits"])
    s = make_strategy("EntropySampler", logits, torch.zeros(n, 4), ev, lab, 128)
    idx, _ = s.query(60.0)
    pool = s.available_query_idxs(shuffle=False)
    ent = -(torch.softmax(logits, 1) * torch.log_softmax(logits, 1)).sum(1)[pool]
    assert idx == pool[np.argsort(-ent.numpy(), kind="stable")[:60]].tolist()


@pytest.mark.parametrize("name,sub,parts", [("CoresetSampler", False, 1), ("CoresetSampler", True, 1),
                                            ("PartitionedCoresetSampler", True, 3),
                                            ("BADGESampler", True, 1),
                                            ("PartitionedBADGESampler", True, 3)])
@pytest.mark.parametrize("etag", ["int", "f32"])
def test_coreset_family_matches_reference(gold, name, sub, parts, etag):
    n, ev, lab = _pool(gold)
    kw = dict(partitions=parts)
    if sub:
        kw.update(subset_labeled=60, subset_unlabeled=300)
    s = make_strategy(name, torch.from_numpy(gold["e2e_logits"]),
                      torch.from_numpy(gold[f"e2e_emb_{etag}"]), ev, lab, 64, **kw)
    np.random.seed(21)
    idx, cost = s.query(50.0)
    assert cost == 50
    assert [int(i) for i in idx] == gold[f"e2e_{name}_{'sub' if sub else 'all'}_{etag}"].tolist()
    s.update(idx, cost)
    pickle.loads(pickle.dumps(s))                       # still picklable after a device query


def test_two_rounds_with_cache(gold):
    """freeze_feature + no subsets: round 2 reuses the cached embedding slab (coreset_sampler.py:112-121)."""
    n, ev, lab = _pool(gold)
    s = make_strategy("CoresetSampler", torch.from_numpy(gold["e2e_logits"]),
                      torch.from_numpy(gold["e2e_emb_int"]), ev, lab, 64)
    np.random.seed(21)
    idx1, c1 = s.query(50.0)
    assert idx1 == gold["e2e_CoresetSampler_all_int"].tolist()
    s.update(idx1, c1)
    idx2, c2 = s.query(50.0)
    assert c2 == 50 and not set(idx2) & set(idx1) and not set(idx2) & set(lab.tolist())
    s.update(idx2, c2)
