"""Small, fixed sequence of the hot kernels at BASELINE shapes, for ncu (never a bench number).

    ncu ... python tools/prof_target.py [margin] [coreset] [badge] [k3]
PROF_STEPS (default 24) selection steps per persistent launch -- ncu replays the whole launch per metric pass."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_learning_b200.engine import Engine  # noqa: E402

which = set(sys.argv[1:]) or {"margin", "coreset", "badge"}
eng = Engine(0)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
N, C, D, L, B = 80000, 1000, 2048, 50000, 10000
if "margin" in which:
    logits = torch.randn(N, C, device=dev, generator=g) * 3
    for _ in range(3):
        s, pos = eng.uncertainty_tail(logits, 0, B)          # K1 + K1b fused (one cooperative launch)
    s = eng.score_softmax(logits, 0)                         # the two separate kernels, for comparison
    pos = eng.select_smallest(s, B)
    torch.cuda.synchronize()
    del logits
if which & {"coreset", "badge", "k3"}:
    X = torch.relu(torch.randn(N, D, device=dev, generator=g))
    Y = torch.relu(torch.randn(L, D, device=dev, generator=g))
    xn, yn = eng.row_norm2(X), eng.row_norm2(Y)
    steps = int(os.environ.get("PROF_STEPS", "24"))
    if which & {"coreset", "k3"}:
        mind = eng.min_dist(X, xn, Y, yn)
        if "coreset" in which:
            for _ in range(2):
                eng.greedy_select(X, xn, mind.clone(), [0, N], [steps])
    if "badge" in which:
        lx = torch.randn(N, C, device=dev, generator=g) * 3
        ly = torch.randn(L, C, device=dev, generator=g) * 3
        XA, xan = eng.badge_factors(lx, 128)
        YA, yan = eng.badge_factors(ly, 128)
        del lx, ly
        mind = eng.min_dist(X, xn, Y, yn, XA, xan, YA, yan)
        us = np.random.default_rng(0).random(steps)
        vpos = torch.arange(L, L + N, dtype=torch.int32, device=dev)
        for _ in range(2):
            eng.greedy_select(X, xn, mind.clone(), [0, N], [steps], a=XA, an=xan, uniforms=us, vpos=vpos, full_n=[N + L])
    torch.cuda.synchronize()
print("prof target done", eng.launches)
