"""BalancingSampler on the device: "Active Learning for Imbalanced Datasets" (WACV 2020).

Query skeleton of /root/reference/src/query_strategies/balancing_sampler.py:26-134, one pick per step.  The class
histogram test (:66-82), the RNG draw of the random branch (:124) and the mask updates (:127-128) are host bookkeeping
and stay literally the reference's (tiny torch-CPU tensors, so every promotion and division rounds the same way).
What moves to the GPU is the part that scales with the pool: the embeddings never leave the device, class centres are
kept as running per-class sums, the two distance fields of a balancing step are K3 calls (`alq_min_dist`, min against
the rarest centre and max over the majority centres) and the masked ratio arg-min is `alq_ratio_argmin`; one int32
returns to the host per step.
"""
from __future__ import annotations

import numpy as np
import torch

from .strategy import EngineMixin


class BalancingQuery(EngineMixin):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.freeze_feature = kwargs["freeze_feature"]          # required, like balancing_sampler.py:24
        self._bal_cache = None

    def _pool_embeddings(self):
        """Embeddings [n_pool, D padded to x4] on the device and the pool's labels (:39-58); kept across rounds under
        --freeze_feature like the reference's saved_embeddings / saved_ys."""
        if self.freeze_feature and getattr(self, "_bal_cache", None) is not None:
            return self._bal_cache
        labels = []
        self.feature_net.eval()
        _, emb = self._forward_pool(list(range(self.n_pool)), self.feature_net, want_features=True, labels_out=labels)
        ys = torch.cat([torch.as_tensor(b).reshape(-1) for b in labels], dim=0).to(torch.int64)
        if self.freeze_feature:
            self._bal_cache = (emb, ys)
        return emb, ys

    def query(self, budget):
        self.feature_net = self.net
        idxs_for_query = self.available_query_idxs(boolean=True)
        idxs_labeled = self.already_labeled_idxs(boolean=True)
        labeled_idxs_cur_rd = []
        emb, ys = self._pool_embeddings()
        eng = self.get_engine()
        dev = emb.device
        C = self.num_classes
        budget = int(min(idxs_for_query.sum(), budget))
        if budget <= 0:
            return labeled_idxs_cur_rd, 0
        # device state: per-class embedding sums, row norms, availability mask
        ys_dev = ys.to(dev)
        lab_rows = torch.from_numpy(np.flatnonzero(idxs_labeled)).to(dev)
        sums = torch.zeros((C, emb.shape[1]), dtype=torch.float32, device=dev)
        if lab_rows.numel():
            sums.index_add_(0, ys_dev[lab_rows], emb[lab_rows])
        xn = eng.row_norm2(emb)
        avail = torch.from_numpy(idxs_for_query.astype(np.uint8)).to(dev)
        ys_labeled_count = torch.bincount(ys[torch.from_numpy(idxs_labeled)], minlength=C)[:C]      # (C,) int64, :66-67
        query_count = 0
        for _ in range(budget):
            mean_labeled_count = ys_labeled_count.float().mean()
            maj_classes = ys_labeled_count > mean_labeled_count
            maj_classes_avgcount = ys_labeled_count[maj_classes].sum() / maj_classes.sum()
            minor_classes = ys_labeled_count <= mean_labeled_count
            minor_classes_avgcount = ys_labeled_count[minor_classes].sum() / minor_classes.sum()
            if budget - query_count <= minor_classes.sum() * (maj_classes_avgcount - minor_classes_avgcount):   # :81-82
                rarest_class_count, rarest_class = ys_labeled_count.min(dim=0)
                rows = torch.cat([rarest_class.reshape(1), torch.nonzero(maj_classes).reshape(-1)]).to(dev)
                denom = ys_labeled_count.to(torch.float32).to(dev)[rows] + 1e-5                      # :88
                centres = sums[rows] / denom[:, None]
                cn = eng.row_norm2(centres)
                d_major = eng.min_dist(emb, xn, centres[1:], cn[1:], reduce_max=True)               # :109-114
                d_rare = None
                if rarest_class_count != 0:                                                          # :104-107
                    d_rare = eng.min_dist(emb, xn, centres[:1], cn[:1])                              # :98-101
                query_idx = np.int64(eng.ratio_argmin(d_rare, d_major, avail))                       # :115-121
            else:
                query_idx = np.random.choice(np.where(idxs_for_query.squeeze() == True)[0])          # noqa: E712  :124
            idxs_for_query[query_idx] = False
            idxs_labeled[query_idx] = True
            labeled_idxs_cur_rd.append(query_idx)
            query_count += 1
            cls = int(ys[int(query_idx)])
            ys_labeled_count[cls] += 1
            sums[cls] += emb[int(query_idx)]
            avail[int(query_idx)] = 0
        return labeled_idxs_cur_rd, query_count
