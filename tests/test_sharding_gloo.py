"""N > 1 host logic on CPU: world_size-2 gloo process groups, OracleEngine injected.
Sharded results must equal the single-process results (and therefore the reference's)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, path, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import OracleEngine, make_strategy
    from active_learning_b200.sharding import ShardGroup
    gold = dict(np.load(path))
    n, ev, lab = int(gold["e2e_n"]), gold["e2e_eval_idxs"], gold["e2e_labeled"]
    group = ShardGroup()
    res = {}
    # row-sharded uncertainty sampler, incl. a tie-heavy (dyadic) pool
    for tag in ("f32_c10", "dyadic_c10"):
        s = make_strategy("MarginSampler", torch.from_numpy(gold[f"margin_{tag}_logits"]),
                          torch.zeros(n, 4), ev, lab, 128, engine=OracleEngine())
        s._shard_group = group
        np.random.seed(7)
        res[f"margin_{tag}"] = s.query(60.0)[0]
    # partitions dealt to ranks
    for name in ("PartitionedCoresetSampler", "PartitionedBADGESampler"):
        s = make_strategy(name, torch.from_numpy(gold["e2e_logits"]), torch.from_numpy(gold["e2e_emb_int"]),
                          ev, lab, 64, engine=OracleEngine(), partitions=3, subset_labeled=60,
                          subset_unlabeled=300)
        s._shard_group = group
        np.random.seed(21)
        res[name] = [int(i) for i in s.query(50.0)[0]]
    # global CoreSet / BADGE: rows forwarded per rank, replicated by one all-gather, loop sharded by candidate row
    # (leaf-aligned for the D^2 draw); incl. the cold start (nothing labeled)
    for name, kw in (("CoresetSampler", {}), ("CoresetSampler", dict(subset_labeled=60, subset_unlabeled=300)),
                     ("BADGESampler", dict(subset_labeled=60, subset_unlabeled=300))):
        s = make_strategy(name, torch.from_numpy(gold["e2e_logits"]), torch.from_numpy(gold["e2e_emb_int"]),
                          ev, lab, 64, engine=OracleEngine(), **kw)
        s._shard_group = group
        np.random.seed(21)
        res[f"global_{name}_{'sub' if kw else 'all'}"] = [int(i) for i in s.query(50.0)[0]]
    for name in ("CoresetSampler", "BADGESampler"):
        for sharded in (True, False):
            s = make_strategy(name, torch.from_numpy(gold["e2e_logits"]), torch.from_numpy(gold["e2e_emb_int"]),
                              ev, [], 64, engine=OracleEngine())
            if sharded:
                s._shard_group = group
            np.random.seed(5)
            res[f"cold_{name}_{'multi' if sharded else 'single'}"] = [int(i) for i in s.query(12.0)[0]]
    # MASE rows sharded + merged; BASE margins sharded, class loop replicated on the gathered margins
    from helpers import HeadNet
    mg = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_mase.npz")))
    for name in ("MASESampler", "BASESampler"):
        net = HeadNet(torch.from_numpy(mg["a_emb"]), torch.from_numpy(mg["a_weight"]), torch.from_numpy(mg["a_bias"]))
        s = make_strategy(name, None, None, mg["a_eval"], mg["a_labeled"], int(mg["a_bs"]), engine=OracleEngine(), net=net)
        s._shard_group = group
        res[name] = s.query(float(mg["a_budget"]))[0]
    assert group.row_range(11, 0) == (0, 6) and group.row_range(11, 1) == (6, 11)
    # ragged row gather (labeled rows of the global CoreSet query): rank order, padding dropped, empty rank ok
    counts = [3, 0] if world == 2 else [3] * world
    mine = torch.full((counts[rank], 5), float(rank + 1))
    allr = group.all_gather_rows(mine, counts)
    assert allr.shape == (sum(counts), 5) and bool((allr[:3] == 1.0).all())
    vec = group.all_gather_rows(torch.arange(2 + rank, dtype=torch.int64), [2 + q for q in range(world)])
    assert vec.tolist() == [i for q in range(world) for i in range(2 + q)]
    assert sorted(group.my_partitions(5, 0) + group.my_partitions(5, 1)) == list(range(5))
    if rank == 0:
        np.save(out_path, np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_equals_single_process(gold):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = os.path.join(tempfile.mkdtemp(), "res.npy")
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    mp.spawn(_worker, args=(2, port, path, out), nprocs=2, join=True)
    res = np.load(out, allow_pickle=True)[0]
    assert res["margin_f32_c10"] == gold["margin_f32_c10_picks"].tolist()
    assert res["PartitionedCoresetSampler"] == gold["e2e_PartitionedCoresetSampler_sub_int"].tolist()
    assert res["PartitionedBADGESampler"] == gold["e2e_PartitionedBADGESampler_sub_int"].tolist()
    assert res["global_CoresetSampler_all"] == gold["e2e_CoresetSampler_all_int"].tolist()
    assert res["global_CoresetSampler_sub"] == gold["e2e_CoresetSampler_sub_int"].tolist()
    assert res["global_BADGESampler_sub"] == gold["e2e_BADGESampler_sub_int"].tolist()
    for name in ("CoresetSampler", "BADGESampler"):
        assert res[f"cold_{name}_multi"] == res[f"cold_{name}_single"] and len(res[f"cold_{name}_multi"]) == 12
    mg = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_mase.npz")))
    assert res["MASESampler"] == mg["a_mase_picks"].tolist()
    assert res["BASESampler"] == mg["a_base_picks"].tolist()
    # tie-heavy pool: the sharded result equals the single-process stable selection
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import OracleEngine, make_strategy
    n, ev, lab = int(gold["e2e_n"]), gold["e2e_eval_idxs"], gold["e2e_labeled"]
    s = make_strategy("MarginSampler", torch.from_numpy(gold["margin_dyadic_c10_logits"]),
                      torch.zeros(n, 4), ev, lab, 128, engine=OracleEngine())
    assert res["margin_dyadic_c10"] == s.query(60.0)[0]
