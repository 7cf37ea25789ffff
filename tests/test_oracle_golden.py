"""The oracle (oracle/al_oracle.py) against vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import al_oracle as O


def _pool(gold):
    n = int(gold["e2e_n"])
    lb = np.zeros(n, dtype=bool)
    lb[gold["e2e_labeled"]] = True
    return n, lb, gold["e2e_eval_idxs"]


def _same_selection(got, ref, pool, scores, exact_order):
    """Without exact score ties the pick lists are identical.  With ties the reference's order is
    whatever torch.sort(stable=False) does (not stable on CPU: observed in the dyadic golden), so
    the fixed tie-break of this build (stable: lowest pool position first) can only be compared
    through the score sequence and the strictly-below-the-boundary set."""
    if exact_order:
        assert got == ref
        return
    pos_of = {int(g): i for i, g in enumerate(pool)}
    sc = scores.numpy()
    s_got = [sc[pos_of[g]] for g in got]
    s_ref = [sc[pos_of[g]] for g in ref]
    assert s_got == s_ref                      # same ascending score sequence
    boundary = s_got[-1]
    assert {g for g, s in zip(got, s_got) if s < boundary} == \
        {g for g, s in zip(ref, s_ref) if s < boundary}
    tied = [g for g, s in zip(got, s_got) if s == boundary]
    all_tied = [int(pool[i]) for i in np.flatnonzero(sc == boundary)]
    assert tied == all_tied[:len(tied)]        # stable: earliest pool positions win the tie


def test_margin_matches_reference(gold):
    n, lb, ev = _pool(gold)
    for tag in ("f32_c10", "f32_c1000", "dyadic_c10"):
        np.random.seed(7)
        pool = O.available_query_idxs(lb, ev, shuffle=False)
        logits = torch.from_numpy(gold[f"margin_{tag}_logits"])[pool]
        idx, cost = O.uncertainty_query(logits, pool, 60.0, O.MODE_MARGIN)
        assert cost == 60
        _same_selection(idx, gold[f"margin_{tag}_picks"].tolist(), pool,
                        O.softmax_scores(logits, O.MODE_MARGIN), exact_order=not tag.startswith("dyadic"))


def test_confidence_matches_reference_minus_line41(gold):
    assert int(gold["confidence_raises"]) == 1  # the shipped bug is real
    n, lb, ev = _pool(gold)
    for tag in ("f32", "dyadic"):
        np.random.seed(7)
        pool = O.available_query_idxs(lb, ev, shuffle=True)  # confidence_sampler.py:19
        logits = torch.from_numpy(gold[f"confidence_{tag}_logits"])[pool]
        idx, _ = O.uncertainty_query(logits, pool, 60.0, O.MODE_LEAST_CONFIDENCE)
        _same_selection(idx, gold[f"confidence_{tag}_picks"].tolist(), pool,
                        O.softmax_scores(logits, O.MODE_LEAST_CONFIDENCE), exact_order=tag == "f32")


def test_pairwise_and_coreset_match_reference(gold):
    ind = gold["cs_indicator"]
    for tag in ("int", "f32"):
        feat = torch.from_numpy(gold[f"cs_{tag}_feat"])
        d2 = O.pairwise_l2_dist(feat)
        assert np.array_equal(d2.numpy(), gold[f"cs_{tag}_d2"])
        assert O.coreset(d2, ind, 40) == gold[f"cs_{tag}_greedy"].tolist()
        np.random.seed(11)
        assert O.coreset(d2, ind, 40, randomize=True) == gold[f"cs_{tag}_d2sample"].tolist()
        none = np.zeros_like(ind)
        assert O.coreset(d2, none, 6) == gold[f"cs_{tag}_greedy_cold"].tolist()
        np.random.seed(12)
        assert O.coreset(d2, none, 6, randomize=True) == gold[f"cs_{tag}_d2sample_cold"].tolist()


def test_nan_retry_branch_matches_reference(gold):
    feat = torch.from_numpy(gold["cs_dup_feat"])
    np.random.seed(13)
    with np.errstate(invalid="ignore"):
        got = O.coreset(O.pairwise_l2_dist(feat), gold["cs_dup_indicator"], 5, randomize=True)
    assert got == gold["cs_dup_d2sample"].tolist()


def test_streaming_equals_dense_on_exact_fixture(gold):
    """Running-min form == the reference's O(N*L) re-gather form (SURVEY.md finding 4)."""
    feat = torch.from_numpy(gold["cs_int_feat"])
    ind = gold["cs_indicator"]
    assert O.coreset_streaming(feat, ind, 40) == gold["cs_int_greedy"].tolist()
    np.random.seed(11)
    assert O.coreset_streaming(feat, ind, 40, randomize=True) == gold["cs_int_d2sample"].tolist()
    np.random.seed(11)
    us = np.random.random_sample(40)
    assert O.coreset_streaming(feat, ind, 40, randomize=True, uniforms=us) == \
        gold["cs_int_d2sample"].tolist()


def test_gradient_embeddings_match_reference(gold):
    lg, hg = torch.from_numpy(gold["ge_logits"]), torch.from_numpy(gold["ge_emb"])
    assert np.array_equal(O.gradient_embeddings(lg, hg, 16).numpy(), gold["ge_full"])
    assert np.array_equal(O.gradient_embeddings(lg, hg, 16, True).numpy(), gold["ge_pooled"])
    lg2, hg2 = torch.from_numpy(gold["ge2_logits"]), torch.from_numpy(gold["ge2_emb"])
    assert np.array_equal(O.gradient_embeddings(lg2, hg2, 4, True).numpy(), gold["ge2_pooled"])


def test_badge_factors_reproduce_gradient_embedding(gold):
    """Closed form a_i (x) h_i (incl. the 1/bs of the short last batch) and its pooled form."""
    lg, hg = torch.from_numpy(gold["ge_logits"]), torch.from_numpy(gold["ge_emb"])
    a = O.badge_factors(lg, 16)
    full = (a[:, :, None] * hg[:, None, :]).reshape(len(a), -1)
    np.testing.assert_allclose(full.numpy(), gold["ge_full"], rtol=0, atol=2e-7)
    pa, ph = O.pooled_factors(a, hg)
    pooled = (pa[:, :, None] * ph[:, None, :]).reshape(len(a), -1)
    np.testing.assert_allclose(pooled.numpy(), gold["ge_pooled"], rtol=0, atol=2e-7)
    lg2, hg2 = torch.from_numpy(gold["ge2_logits"]), torch.from_numpy(gold["ge2_emb"])
    pa, ph = O.pooled_factors(O.badge_factors(lg2, 4), hg2)
    assert pa.shape[1] == 16 and ph.shape[1] == 32
    pooled = (pa[:, :, None] * ph[:, None, :]).reshape(len(pa), -1)
    np.testing.assert_allclose(pooled.numpy(), gold["ge2_pooled"], rtol=0, atol=2e-7)


def test_numpy_pairwise_sum_restatement_is_bit_exact():
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 9, 127, 128, 129, 255, 1000, 13000, 80000, 130001):
        a = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
        assert O.np_pairwise_sum_f32(a) == np.sum(a), n
        leaves = O.pairwise_leaves(n)
        assert sum(m for _, m in leaves) == n
        assert all(m <= 128 for _, m in leaves)


def test_choice_is_cdf_search(gold):
    p, us = gold["choice_p"], gold["choice_uniforms"]
    got = [O.choice_from_uniform(p, u) for u in us]
    assert got == gold["choice_draws"].tolist()


def test_end_to_end_queries_match_reference(gold):
    """Full query() restated from oracle pieces: bookkeeping + RNG order + scoring."""
    n, lb, ev = _pool(gold)
    logits = torch.from_numpy(gold["e2e_logits"])
    for etag in ("int", "f32"):
        emb = torch.from_numpy(gold[f"e2e_emb_{etag}"])
        for name, sub, parts, badge in (("CoresetSampler", False, 1, False),
                                        ("CoresetSampler", True, 1, False),
                                        ("PartitionedCoresetSampler", True, 3, False),
                                        ("BADGESampler", True, 1, True),
                                        ("PartitionedBADGESampler", True, 3, True)):
            sl, su = (60, 300) if sub else (None, None)
            np.random.seed(21)
            union, lab, unl = O.idxs_for_coreset(lb, ev, sl, su)

            def embed(idxs, pooled=False):
                idxs = np.asarray(idxs)
                if badge:
                    return O.gradient_embeddings(logits[idxs], emb[idxs], 64, pooled)
                return emb[idxs]

            if parts == 1:
                d2 = O.pairwise_l2_dist(embed(union))
                ind = lb[union]
                picks = O.coreset(d2, ind, 50, randomize=badge)
                got = np.array(union)[picks].tolist()
            else:
                lab_parts = O.partition_idxs(lab, parts)
                unl_parts = O.partition_idxs(unl, parts)
                got = []
                for i in range(parts):
                    rows = np.concatenate((lab_parts[i], unl_parts[i]))
                    d2 = O.pairwise_l2_dist(embed(rows, pooled=True))
                    ind = np.zeros(len(rows), dtype=bool)
                    ind[:len(lab_parts[i])] = True
                    b_i = int(50 / parts) + int(i < 50 % parts)
                    got += list(rows[O.coreset(d2, ind, b_i, randomize=badge)])
                got = sorted(int(g) for g in got)
            key = f"e2e_{name}_{'sub' if sub else 'all'}_{etag}"
            assert got == gold[key].tolist(), key


def test_factored_streaming_equals_materialised_rows():
    """O.coreset_streaming(factors=...) -- the oracle form used at BASELINE config 4's dimensions, where the
    2048*1000-d gradient embedding cannot be formed -- against the dense path on the materialised a (x) h rows
    (badge_sampler.py:40), on exact-arithmetic fixtures: identical picks, both modes."""
    rng = np.random.default_rng(9)
    n, c, d, l0, b = 700, 8, 16, 100, 50
    a = torch.from_numpy(rng.integers(-1, 2, size=(n, c)).astype(np.float32))
    h = torch.from_numpy(rng.integers(-1, 2, size=(n, d)).astype(np.float32))
    g = (a[:, :, None] * h[:, None, :]).reshape(n, -1)
    ind = np.zeros(n, dtype=bool)
    ind[rng.choice(n, l0, replace=False)] = True
    us = rng.random(b)
    dist = O.pairwise_l2_dist(g)
    for randomize in (False, True):
        cert = []
        got = O.coreset_streaming(h, ind, b, randomize=randomize, uniforms=us, factors=a, certificate=cert)
        assert got == O.coreset_streaming(g, ind, b, randomize=randomize, uniforms=us)
        assert len(cert) == b and min(cert) >= 0.0
        if not randomize:
            assert got == O.coreset(dist, ind, b)
