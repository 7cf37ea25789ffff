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


class _EncoderNet(torch.nn.Module):
    """The reference's model layout (resnet_simclr.py:13-41): `encoder` (here: table lookup -> Linear -> BatchNorm ->
    ReLU, so it has parameters AND buffers) followed by the `linear` head; under --freeze_feature only `linear` moves."""

    def __init__(self, table, d, c):
        super().__init__()

        class Enc(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.register_buffer("table", table.clone())
                self.fc = torch.nn.Linear(table.shape[1], d)
                self.bn = torch.nn.BatchNorm1d(d)
                self.calls = 0

            def forward(self, x):
                self.calls += x.shape[0]
                return torch.relu(self.bn(self.fc(self.table[x.long()])))

        self.encoder = Enc()
        self.linear = torch.nn.Linear(d, c)

    def forward(self, x, return_features=False, specify_input_layer=None):
        if specify_input_layer:
            return self.linear(x)
        h = self.encoder(x)
        out = self.linear(h)
        return (out, h) if return_features else out


@pytest.mark.parametrize("name", ["MarginSampler", "CoresetSampler", "BADGESampler"])
def test_embedding_cache_over_two_rounds_equals_uncached(gold, name):
    """SURVEY.md section 8f rank 1 on the GPU: a net with the reference's encoder / linear layout, --freeze_feature.
    The cached sampler (encoder runs once per pool row over all rounds, logits = linear(cached embedding)) must pick
    what the uncached one picks in every round -- also after the head was 'trained' between the rounds, and after an
    encoder BUFFER changed (BatchNorm statistics: the cache must notice and refill)."""
    n, ev, lab = _pool(gold)
    torch.manual_seed(3)
    table = torch.randn(n, 24)
    nets = [_EncoderNet(table, 32, 10) for _ in range(2)]
    nets[1].load_state_dict(nets[0].state_dict())
    for net in nets:
        net.eval()
    kw = dict(freeze_feature=True)
    if name != "MarginSampler":
        kw.update(subset_labeled=60, subset_unlabeled=300)
    from helpers import IndexDataset
    strategies = []
    for net, use_cache in zip(nets, (True, False)):
        s = make_strategy(name, torch.zeros(n, 10), torch.zeros(n, 32), ev, lab, 64, dataset=IndexDataset(n, 10),
                          cache_embeddings=use_cache, **kw)
        s.net = s.feature_net = net          # the lookup net make_strategy installed is replaced by the encoder net
        strategies.append(s)
    cached, plain = strategies
    for rnd in range(3):
        np.random.seed(40 + rnd)
        a, ca = cached.query(30.0)
        np.random.seed(40 + rnd)
        b, cb = plain.query(30.0)
        assert [int(i) for i in a] == [int(i) for i in b] and ca == cb == 30, rnd
        cached.update(a, ca)
        plain.update(b, cb)
        with torch.no_grad():                # "training" moves the head only (frozen encoder)
            for s in strategies:
                s.net.linear.weight.add_(0.01 * (rnd + 1) * torch.sign(s.net.linear.weight))
        if rnd == 1:                         # an encoder buffer changes: the cache has to be rebuilt, not reused
            with torch.no_grad():
                for s in strategies:
                    s.net.encoder.bn.running_mean.add_(0.25)
    assert cached._emb_cache is not None
    # the encoder saw every queried row once per cache generation (2 generations), the uncached one once per round
    assert nets[0].encoder.calls < nets[1].encoder.calls
