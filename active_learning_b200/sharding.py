"""Multi-GPU plumbing of the query path: one process per GPU, torch.distributed for the (few,
small) exchanges.  Rows of the scoring kernels and partitions of the partitioned samplers are
independent, so the data path needs no collective; what is exchanged is

  * uncertainty samplers : each rank's local top-B (score, global position) -> one all-gather of
                           G*B 12-byte records -> the same stable merge on every rank;
  * partitioned samplers : partition i runs on rank i % G; the picked row lists are gathered once.

Works with backend "nccl" (GPU) and "gloo" (the CPU tests of this host logic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class ShardGroup:
    def __init__(self, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.pg = process_group
        self.rank = dist.get_rank(process_group)
        self.world_size = dist.get_world_size(process_group)

    # ---- row sharding (K1 / K2) -----------------------------------------------------------------
    def row_range(self, n, rank=None):
        """Contiguous block of rows owned by `rank`: sizes differ by at most one."""
        r = self.rank if rank is None else rank
        base, rem = divmod(int(n), self.world_size)
        lo = r * base + min(r, rem)
        return lo, lo + base + (1 if r < rem else 0)

    def merge_smallest(self, scores_local, pos_local, row_lo, budget, engine):
        """Global `budget` smallest (score, global position) pairs from every rank's local winners.

        scores_local: this rank's score vector; pos_local: its local top-B positions in ascending
        (score, position) order (K1b output); row_lo: global position of local row 0.
        The gathered array is ordered by rank and, inside a rank, by (score, position); ranks own
        ascending position ranges, so among equal scores array order == global-position order and
        K1b's "lowest array position first" tie-break IS the global tie-break.  The merge is one
        more K1b launch over G*B candidates; only the B winners go to the host."""
        dev = scores_local.device
        b = int(budget)
        k = int(pos_local.numel())
        s = torch.full((b,), float("inf"), dtype=torch.float32, device=dev)
        g = torch.full((b,), -1, dtype=torch.int64, device=dev)
        if k:
            idx = pos_local.long()
            s[:k] = scores_local[idx]
            g[:k] = idx + int(row_lo)
        s_all = [torch.empty_like(s) for _ in range(self.world_size)]
        g_all = [torch.empty_like(g) for _ in range(self.world_size)]
        dist.all_gather(s_all, s, group=self.pg)
        dist.all_gather(g_all, g, group=self.pg)
        s_cat, g_cat = torch.cat(s_all), torch.cat(g_all)
        sel = engine.select_smallest(s_cat, b)
        out = g_cat[sel.long()].cpu().numpy()
        assert (out >= 0).all()
        return out

    # ---- partition dealing ------------------------------------------------------------------------
    def my_partitions(self, n_parts, rank=None):
        r = self.rank if rank is None else rank
        return [i for i in range(n_parts) if i % self.world_size == r]

    def all_gather_dict(self, mine: dict) -> dict:
        """Union of per-rank {partition: picks} dicts (picks are small int arrays)."""
        payload = {int(k): np.asarray(v).tolist() for k, v in mine.items()}
        out = [None] * self.world_size
        dist.all_gather_object(out, payload, group=self.pg)
        merged = {}
        for d in out:
            for k, v in d.items():
                merged[int(k)] = np.asarray(v, dtype=np.int64)
        return merged
