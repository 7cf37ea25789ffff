"""Debug aid: the sharded fused tail on an all-ties pool, both ranking routes, against the separate kernels."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_learning_b200.engine import Engine            # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
eng = Engine(local).comm_init()
dev = eng.device
for (n_per, c, b, seed, dyadic) in ((9000, 64, 3000, 12, True), (9000, 8, 3000, 13, True), (20000, 1000, 10000, 11, False), (5000, 40, 4999, 14, False)):
    g = torch.Generator(device=dev).manual_seed(seed)
    n_tot = n_per * world
    logits = (torch.randint(-2, 3, (n_tot, c), generator=g, device=dev).float() if dyadic
              else torch.randn(n_tot, c, generator=g, device=dev) * 3)
    lo = rank * n_per
    shard = logits[lo:lo + n_per].contiguous()
    for mode in (0, 1):
        truth = eng.select_smallest(eng.score_softmax(logits, mode), b)
        for buckets in (1, 0):
            eng.set_option("tail_buckets", buckets)
            _, one = eng.uncertainty_tail(logits, mode, b)
            torch.cuda.synchronize()
            dist.barrier()
            _, got = eng.uncertainty_tail_sharded(shard, mode, b, lo, n_per, n_per)
            torch.cuda.synchronize()
            msg = ""
            try:
                eng.comm_check()
            except Exception as exc:
                msg = str(exc)[:80]
            d1 = (truth != one).nonzero().flatten()
            d2 = (truth != got).nonzero().flatten()
            if rank == 0:
                print(f"n/rank {n_per} c {c} b {b} mode {mode} buckets {buckets}: one-GPU fused wrong at {d1.numel()} "
                      f"sharded wrong at {d2.numel()} {msg}", flush=True)
                if d2.numel():
                    k = int(d2[0])
                    print("   first diff", k, truth[max(0, k - 2):k + 4].tolist(), got[max(0, k - 2):k + 4].tolist(),
                          "last diff", int(d2[-1]), flush=True)
            dist.barrier()
eng.set_option("tail_buckets", 1)
dist.destroy_process_group()
