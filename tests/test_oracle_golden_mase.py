"""MASE / BASE restatement in oracle/al_oracle.py against vectors produced by the reference itself
(tests/golden/make_golden_mase.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import al_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAGS = ("a", "b", "t", "d")     # C=10 / C=1000 / exact ties / duplicated class rows


@pytest.fixture(scope="module")
def mgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_mase.npz")))


def pool_of(g, tag):
    n = g[f"{tag}_emb"].shape[0]
    lb = np.zeros(n, dtype=bool)
    lb[g[f"{tag}_labeled"]] = True
    return O.available_query_idxs(lb, g[f"{tag}_eval"], shuffle=False)


def head_of(g, tag):
    return (torch.from_numpy(g[f"{tag}_emb"]), torch.from_numpy(g[f"{tag}_weight"]),
            torch.from_numpy(g[f"{tag}_bias"]))


def base_chunks(budget, c):
    """(class, start, count) of every non-empty per-class chunk of a BASE pick list (base_sampler.py:24-26)."""
    out, at = [], 0
    for cls in range(c):
        cnt = budget // c + int(cls < budget % c)
        if cnt:
            out.append((cls, at, cnt))
            at += cnt
    return out


def check_base_replay(picks_pos, ref_pos, mm, pc, pred, budget, c):
    """Tie-tolerant BASE comparison: replay the reference's own pick list class by class.  Given the rows the
    reference had taken before class c, a stable selection must produce the same ascending key sequence as the
    reference's chunk (the reference's torch.sort is not stable, so tied rows may differ)."""
    assert len(set(picks_pos)) == len(picks_pos) == len(ref_pos)
    for cls, at, cnt in base_chunks(budget, c):
        key = torch.where(pred == cls, mm, pc[:, cls]).clone()
        if at:
            key[torch.as_tensor(ref_pos[:at])] = float("inf")
        mine = O.select_smallest(key, cnt)
        assert key[mine].tolist() == key[torch.as_tensor(ref_pos[at:at + cnt])].tolist(), cls


@pytest.mark.parametrize("tag", TAGS)
def test_margins_match_reference(mgold, tag):
    g = mgold
    pool = pool_of(g, tag)
    emb, w, b = head_of(g, tag)
    mm, pc, pred = O.mase_margins(emb[pool], w, b, int(g[f"{tag}_bs"]))
    assert torch.equal(pred, torch.from_numpy(g[f"{tag}_pred"]))
    assert torch.equal(pc, torch.from_numpy(g[f"{tag}_per_class"]))        # same library, same order: bit-exact
    assert torch.equal(mm, torch.from_numpy(g[f"{tag}_min_margins"]))
    if tag == "d":                                                         # duplicated rows: NaN -> inf branch taken
        assert int(torch.isinf(pc).sum(dim=1).max()) == 2


@pytest.mark.parametrize("tag", TAGS)
def test_mase_and_base_picks_match_reference(mgold, tag):
    g = mgold
    pool = pool_of(g, tag)
    emb, w, b = head_of(g, tag)
    budget, c = int(g[f"{tag}_budget"]), w.shape[0]
    mm, pc, pred = O.mase_margins(emb[pool], w, b, int(g[f"{tag}_bs"]))
    idx, cost = O.mase_query(mm, pool, float(budget))
    assert cost == budget
    pos_of = {int(p): i for i, p in enumerate(pool)}
    ref = g[f"{tag}_mase_picks"].tolist()
    if tag in ("a", "b", "d"):
        assert idx == ref
    else:                                                                   # exact ties: compare the score sequence
        assert [float(mm[pos_of[i]]) for i in idx] == [float(mm[pos_of[i]]) for i in ref]
    base = O.base_select(mm, pc, pred, budget, c)
    ref_pos = [pos_of[int(i)] for i in g[f"{tag}_base_picks"]]
    if tag in ("a", "b"):
        assert base.tolist() == ref_pos
    check_base_replay(base.tolist(), ref_pos, mm, pc, pred, budget, c)


@pytest.mark.parametrize("tag", TAGS)
def test_closed_form_equals_broadcast_form(mgold, tag):
    """|z_pred - z_c| / |w_pred - w_c| (what the CUDA kernel evaluates) against the reference's broadcast
    arithmetic: same predictions, same inf pattern, radii within fp32 rounding of the two evaluation orders."""
    g = mgold
    pool = pool_of(g, tag)
    emb, w, b = head_of(g, tag)
    logits = torch.nn.functional.linear(emb[pool], w, b)
    mm2, pc2, pred2 = O.mase_margins_closed_form(logits, w)
    pc = torch.from_numpy(g[f"{tag}_per_class"])
    assert torch.equal(pred2, torch.from_numpy(g[f"{tag}_pred"]))
    assert torch.equal(torch.isinf(pc2), torch.isinf(pc))
    fin = torch.isfinite(pc)
    torch.testing.assert_close(pc2[fin], pc[fin], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(mm2, torch.from_numpy(g[f"{tag}_min_margins"]), rtol=1e-4, atol=2e-6)
    if tag in ("a", "b"):                                                   # well-separated fixtures: same picks
        budget = int(g[f"{tag}_budget"])
        assert O.mase_query(mm2, pool, budget)[0] == g[f"{tag}_mase_picks"].tolist()
        pos_of = {int(p): i for i, p in enumerate(pool)}
        assert O.base_select(mm2, pc2, pred2, budget, w.shape[0]).tolist() == \
            [pos_of[int(i)] for i in g[f"{tag}_base_picks"]]
