"""Multi-GPU selection loop vs the single-GPU loop on the same data (run under torchrun, >= 2 GPUs).
Picks must be identical for every world size: per-row arithmetic does not depend on the sharding and
the D^2 draw runs on a replicated array."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_learning_b200.engine import Engine

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
eng = Engine(local).comm_init()
dev = eng.device
ok = True

def case(n, l, d, c, b, ints, seed, variant):
    global ok
    g = torch.Generator(device="cpu").manual_seed(seed)
    if ints:
        X = torch.randint(-1, 2, (n, d), generator=g).float(); Y = torch.randint(-1, 2, (l, d), generator=g).float()
        XA = torch.randint(-1, 2, (n, max(c, 4)), generator=g).float(); YA = torch.randint(-1, 2, (l, max(c, 4)), generator=g).float()
    else:
        X = torch.relu(torch.randn(n, d, generator=g)); Y = torch.relu(torch.randn(l, d, generator=g))
        XA = torch.randn(n, max(c, 4), generator=g) * .05; YA = torch.randn(l, max(c, 4), generator=g) * .05
    X, Y, XA, YA = X.to(dev), Y.to(dev), XA.to(dev), YA.to(dev)
    us = np.random.default_rng(seed).random(b)
    cand_pos = np.sort(np.random.default_rng(seed + 1).choice(n + l, n, replace=False)).astype(np.int32)   # positions in the full array
    # uneven shards
    cuts = [0] + sorted(np.random.default_rng(seed + 2).choice(np.arange(1, n), world - 1, replace=False).tolist()) + [n]
    lo, hi = cuts[rank], cuts[rank + 1]
    for sample in (False, True):
        fac = c > 0
        def run(xs, xas, pos, shard):
            xn = eng.row_norm2(xs); yn = eng.row_norm2(Y)
            xan = eng.row_norm2(xas) if fac else None; yan = eng.row_norm2(YA) if fac else None
            mind = eng.min_dist(xs, xn, Y, yn, xas if fac else None, xan, YA if fac else None, yan)
            return eng.greedy_select(xs, xn, mind, [0, xs.shape[0]], [b], a=xas if fac else None, an=xan,
                                     uniforms=us if sample else None,
                                     vpos=torch.as_tensor(pos, device=dev) if sample else None,
                                     full_n=[n + l] if sample else None, variant=variant,
                                     shard_off=shard, vpos_all=torch.as_tensor(cand_pos, device=dev) if (sample and shard is not None) else None)
        single = run(X, XA, cand_pos, None)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        multi = run(X[lo:hi].contiguous(), XA[lo:hi].contiguous(), cand_pos[lo:hi], cuts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        same = np.array_equal(single, multi)
        allsame = torch.tensor([int(same)], device=dev); dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"n={n} d={d} c={c} b={b} ints={ints} sample={sample} variant={variant}: "
                  f"{'OK' if allsame.item() else 'MISMATCH'} ({dt*1e3:.1f} ms multi)", flush=True)
            if not allsame.item():
                k = next((i for i, (a_, b_) in enumerate(zip(single, multi)) if a_ != b_), -1)
                print("  first diff at", k, single[max(0,k-2):k+3], multi[max(0,k-2):k+3])
        ok = ok and bool(allsame.item())

for variant in (1, 2):
    case(6001, 700, 512, 0, 64, True, 1, variant)
    case(6001, 700, 512, 40, 64, True, 2, variant)
    case(3000, 300, 2048, 0, 40, False, 3, variant)
    case(3000, 300, 2048, 1000, 40, False, 4, variant)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    case(80000, 50000, 2048, 0, 2000, False, 5, 2)
    case(80000, 50000, 2048, 1000, 2000, False, 6, 2)
dist.barrier()
if rank == 0:
    print("MGPU GREEDY", "PASS" if ok else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
