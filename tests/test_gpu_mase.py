"""K6 (MASE / BASE) on the GPU through the C ABI: against the vectors the reference itself produced
(tests/golden/reference_golden_mase.npz) and against the oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from helpers import HeadNet, make_strategy
from oracle import al_oracle as O
from test_oracle_golden_mase import check_base_replay, head_of, pool_of

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-4, 2e-6       # north star: per-sample scores within 1e-4 fp32


@pytest.fixture(scope="module")
def eng():
    from active_learning_b200.engine import Engine
    return Engine()


@pytest.fixture(scope="module")
def mgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_mase.npz")))


@pytest.mark.parametrize("c,m", [(2, 1), (10, 32), (37, 50), (100, 2048), (1000, 64), (1000, 512)])
def test_class_gap_inv(eng, c, m):
    torch.manual_seed(c + m)
    w = torch.randn(c, m)
    if c >= 10:
        w[7] = w[3]                                     # coinciding class rows -> +inf, not NaN
    g, gmin = (t.cpu() for t in eng.class_gap_inv(w.cuda()))
    assert g.shape == (c, (c + 3) & ~3)
    assert torch.equal(gmin[:c], g[:, :c].min(dim=1).values) and torch.equal(g[:, :c], g[:, :c].T)
    gf = torch.where(torch.isinf(g[:, :c]), torch.zeros(()), g[:, :c])
    want_ratio = (gf.max(dim=1).values / gmin[:c]).clamp_min(1.0).max() if c > 2 else gmin[c]
    torch.testing.assert_close(gmin[c], want_ratio, rtol=1e-6, atol=0)
    den = ((w.double()[:, None, :] - w.double()[None, :, :]) ** 2).sum(dim=2)
    ref = (1.0 / den.sqrt()).float()
    ref.fill_diagonal_(float("inf"))
    assert torch.equal(torch.isinf(g[:, :c]), torch.isinf(ref))
    fin = torch.isfinite(ref)
    torch.testing.assert_close(g[:, :c][fin], ref[fin], rtol=1e-5, atol=0)
    assert torch.isinf(g[:, c:]).all()
    if c >= 10:
        assert torch.isinf(g[3, 7]) and torch.isinf(g[7, 3])


@pytest.mark.parametrize("tag", ["a", "b", "t", "d"])
def test_margins_match_reference_golden(eng, mgold, tag):
    g = mgold
    pool = pool_of(g, tag)
    emb, w, b = head_of(g, tag)
    logits = torch.nn.functional.linear(emb[pool], w, b)
    ginv = eng.class_gap_inv(w.cuda())
    mm, pred, radius = eng.mase_margins(logits.cuda(), ginv, want_per_class=True)
    ref_pc = torch.from_numpy(g[f"{tag}_per_class"])
    assert torch.equal(pred.cpu().long(), torch.from_numpy(g[f"{tag}_pred"]))
    assert torch.equal(torch.isinf(radius.cpu()), torch.isinf(ref_pc))
    fin = torch.isfinite(ref_pc)
    torch.testing.assert_close(radius.cpu()[fin], ref_pc[fin], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(mm.cpu(), torch.from_numpy(g[f"{tag}_min_margins"]), rtol=RTOL, atol=ATOL)
    mm2, pred2, none = eng.mase_margins(logits.cuda(), ginv, want_per_class=False)
    assert none is None and torch.equal(mm2, mm) and torch.equal(pred2, pred)


@pytest.mark.parametrize("n,c", [(1, 10), (257, 10), (300, 37), (200, 12), (333, 128), (300, 500), (1000, 1000),
                                 (513, 1024), (64, 2048), (40, 4100),
                                 (5000, 1000), (4099, 12), (6000, 128), (4500, 500), (4200, 2048)])   # pipelined kernel (n >= 4096)
def test_margins_match_closed_form_oracle(eng, n, c):
    torch.manual_seed(n * 3 + c)
    w = torch.randn(c, 48) * 0.3
    logits = torch.randn(n, c) * 2.0
    ginv = eng.class_gap_inv(w.cuda())
    mm, pred, radius = eng.mase_margins(logits.cuda(), ginv, want_per_class=True)
    rmm, rpc, rpred = O.mase_margins_closed_form(logits, w)
    assert torch.equal(pred.cpu().long(), rpred)
    assert torch.equal(torch.isinf(radius.cpu()), torch.isinf(rpc))
    fin = torch.isfinite(rpc)
    torch.testing.assert_close(radius.cpu()[fin], rpc[fin], rtol=5e-6, atol=1e-30)
    torch.testing.assert_close(mm.cpu(), rmm, rtol=5e-6, atol=1e-30)
    mm2, pred2, _ = eng.mase_margins(logits.cuda(), ginv, want_per_class=False)    # table reads pruned by gmin:
    assert torch.equal(mm2, mm) and torch.equal(pred2, pred)                        # the exact same minimum


def test_pruned_minimum_is_exact_on_adversarial_rows(eng):
    """Min-only path (classes skipped when gap * gmin >= r0) against the all-classes path, bit for bit: near-equal
    logits, equal maxima, -inf / +inf entries, coinciding class rows (inf table entries), a wide spread of distances."""
    torch.manual_seed(9)
    c = 1000
    w = torch.randn(c, 16) * torch.logspace(-3, 1, c)[:, None]       # |w_a - w_c| spans 4 decades
    w[11], w[500] = w[10], w[499]
    z = torch.randn(600, c)
    z[:100] = torch.round(z[:100])                                   # many equal maxima
    z[100:200] = 1.0 + 1e-6 * torch.randn(100, c)                    # everything close to the boundary
    z[200:220, ::3] = float("-inf")
    z[220:230] = float("-inf")                                       # degenerate rows
    z[230:240, 5] = float("inf")
    z[240:260] = 0.0
    z[260:300, 10] = 50.0                                            # predicted class has a coinciding twin (11)
    gap = eng.class_gap_inv(w.cuda())
    full, pred_a, radius = eng.mase_margins(z.cuda(), gap, want_per_class=True)
    pruned, pred_b, _ = eng.mase_margins(z.cuda(), gap, want_per_class=False)
    assert torch.equal(pred_a, pred_b)
    assert torch.equal(full.cpu().view(torch.int32), pruned.cpu().view(torch.int32))
    assert torch.equal(full, radius.min(dim=1).values)
    zbig = z.repeat(8, 1).cuda()                                     # n >= 4096: the pipelined kernel
    a, pa, rad = eng.mase_margins(zbig, gap, want_per_class=True)
    b, pb, _ = eng.mase_margins(zbig, gap, want_per_class=False)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)) and torch.equal(pa, pb)
    assert torch.equal(a[:600].view(torch.int32), full.view(torch.int32)) and torch.equal(rad[:600].view(torch.int32), radius.view(torch.int32))
    for cc in (10, 12, 128, 500):                                    # the other register tilings
        zz = z[:, :cc].contiguous().cuda()
        gg = eng.class_gap_inv(w[:cc].cuda())
        a, _, _ = eng.mase_margins(zz, gg, want_per_class=True)
        b, _, _ = eng.mase_margins(zz, gg, want_per_class=False)
        assert torch.equal(a.cpu().view(torch.int32), b.cpu().view(torch.int32)), cc


def test_margins_on_strided_slab_and_argmax_ties(eng):
    rng = np.random.default_rng(5)
    w = torch.randn(1000, 32)
    ginv = eng.class_gap_inv(w.cuda())
    slab = torch.from_numpy(rng.integers(-3, 4, size=(400, 1024)).astype(np.float32)).cuda()   # heavy arg-max ties
    view = slab[:, :1000]
    mm, pred, radius = eng.mase_margins(view, ginv, want_per_class=True)
    rmm, rpc, rpred = O.mase_margins_closed_form(view.cpu(), w)
    assert torch.equal(pred.cpu().long(), rpred)            # lowest index among equal maxima, like torch's CPU max
    torch.testing.assert_close(mm.cpu(), rmm, rtol=5e-6, atol=1e-30)
    assert (mm.cpu() == 0).sum() > 100                      # tied maxima sit exactly on a boundary
    assert torch.equal(torch.isinf(radius.cpu()), torch.isinf(rpc))


@pytest.mark.parametrize("n,c,budget", [(5000, 10, 333), (3000, 1000, 1500), (700, 37, 700), (90000, 100, 1000),
                                        (6000, 50, 500), (20000, 1000, 10000), (4000, 1003, 3999), (300, 20, 300)])
def test_base_select_matches_oracle(eng, n, c, budget):
    """Sequential class loop and parallel candidate lists + in-order resolve, both against the oracle's loop."""
    torch.manual_seed(n + c)
    w = torch.randn(c, 24) * 0.3
    logits = torch.randn(n, c) * 2.0
    if n in (5000, 6000):                                   # quantised logits: ties inside every per-class sort
        logits = torch.round(logits * 2) / 2
    ginv = eng.class_gap_inv(w.cuda())
    mm, pred, radius = eng.mase_margins(logits.cuda(), ginv, want_per_class=True)
    ref = O.base_select(mm.cpu(), radius.cpu(), pred.cpu().long(), budget, c)
    try:
        for impl in (1, 2, 0):
            eng.set_option("base_impl", impl)
            got = eng.base_select(mm, radius, pred, budget).cpu().numpy()
            assert got.tolist() == ref.tolist(), impl           # bit-exact: same keys, same stable tie-break
    finally:
        eng.set_option("base_impl", 0)


@pytest.mark.parametrize("tag", ["a", "b", "t", "d"])
def test_mase_base_samplers_end_to_end(mgold, tag):
    g = mgold
    pool = pool_of(g, tag)
    pos_of = {int(p): i for i, p in enumerate(pool)}
    budget = int(g[f"{tag}_budget"])
    ref_mm = torch.from_numpy(g[f"{tag}_min_margins"])
    ref_pc, ref_pred = torch.from_numpy(g[f"{tag}_per_class"]), torch.from_numpy(g[f"{tag}_pred"])
    picks = {}
    for name in ("MASESampler", "BASESampler"):
        net = HeadNet(*head_of(g, tag))
        s = make_strategy(name, None, None, g[f"{tag}_eval"], g[f"{tag}_labeled"], int(g[f"{tag}_bs"]), net=net)
        idx, cost = s.query(float(budget))
        assert cost == budget and len(set(idx)) == budget
        s.update(idx, cost)
        picks[name] = idx
    ref_mase, ref_base = g[f"{tag}_mase_picks"].tolist(), g[f"{tag}_base_picks"].tolist()
    if tag in ("a", "b", "d"):                              # keys separated by >= 1e-4: the reference's own lists
        assert picks["MASESampler"] == ref_mase
        assert picks["BASESampler"] == ref_base
    else:                                                   # exact ties (the reference's sort is not stable there):
        torch.testing.assert_close(ref_mm[[pos_of[i] for i in picks["MASESampler"]]],     # same margin sequence,
                                   ref_mm[[pos_of[i] for i in ref_mase]], rtol=RTOL, atol=ATOL)
        from active_learning_b200.engine import Engine     # and BASE == the stable class loop on the same margins
        eng = Engine()
        emb, w, b = head_of(g, tag)
        logits = torch.nn.functional.linear(emb[pool], w, b)        # small integers: exact on any device
        mm, pred, radius = eng.mase_margins(logits.cuda(), eng.class_gap_inv(w.cuda()), want_per_class=True)
        want = O.base_select(mm.cpu(), radius.cpu(), pred.cpu().long(), budget, w.shape[0])
        assert [pos_of[i] for i in picks["BASESampler"]] == want.tolist()
        check_base_replay(want.tolist(), [pos_of[i] for i in ref_base], ref_mm, ref_pc, ref_pred, budget, w.shape[0])


def test_compute_margins_api_and_label_collection(mgold):
    g = mgold
    net = HeadNet(*head_of(g, "a"))
    s = make_strategy("BASESampler", None, None, g["a_eval"], g["a_labeled"], int(g["a_bs"]), net=net)
    pool = s.available_query_idxs(boolean=False, shuffle=False)
    mm, pc, pred, true = s.compute_margins(pool)
    assert not mm.is_cuda and pred.dtype == torch.int64 and len(true) == len(pool)
    torch.testing.assert_close(pc[torch.isfinite(pc)], torch.from_numpy(g["a_per_class"])[torch.isfinite(pc)],
                               rtol=RTOL, atol=ATOL)


def test_base_double_selection_raises_like_the_reference(eng):
    """All class rows coincide -> every key is +inf -> the second class re-selects taken rows: the reference's
    `assert len(labeled_idxs) == len(set(labeled_idxs))` (base_sampler.py:40)."""
    from active_learning_b200._lib import AlqError
    w = torch.ones(2, 8)
    logits = torch.zeros(64, 2)
    ginv = eng.class_gap_inv(w.cuda())
    mm, pred, radius = eng.mase_margins(logits.cuda(), ginv, want_per_class=True)
    assert torch.isinf(mm).all()
    with pytest.raises(AlqError, match="selected twice"):
        eng.base_select(mm, radius, pred, 40)
    net = HeadNet(torch.zeros(64, 8), w, torch.zeros(2))
    s = make_strategy("BASESampler", None, None, [], [], 16, net=net, mase_self_check=False)
    with pytest.raises(AssertionError):
        s.query(40.0)


def test_bad_arguments_are_errors(eng):
    from active_learning_b200._lib import AlqError
    w = torch.randn(10, 8).cuda()
    ginv = eng.class_gap_inv(w)
    with pytest.raises(AlqError):
        eng.mase_margins(torch.randn(5, 12).cuda(), ginv)            # class count mismatch
    with pytest.raises(AlqError):
        eng.mase_margins(torch.randn(5, 10), ginv)                   # CPU tensor: no fallback
    mm, pred, radius = eng.mase_margins(torch.randn(50, 10).cuda(), ginv, want_per_class=True)
    with pytest.raises(AlqError):
        eng.base_select(mm, radius, pred, 51)                        # budget > n
    with pytest.raises(AlqError):
        eng.base_select(mm, radius, pred.long(), 5)                  # pred must be int32
    assert eng.base_select(mm, radius, pred, 0).numel() == 0
