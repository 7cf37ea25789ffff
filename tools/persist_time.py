"""Selection-loop timing at the north-star shapes on one GPU: per-step launches (variant 2) vs the persistent
cooperative kernel (variant 3).  JSON lines on stdout.  env: PT_STEPS (default 1000), ALQ_TILE_KB, ALQ_MAX_STAGES."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_learning_b200.engine import Engine  # noqa: E402

eng = Engine(0)
dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
N, C, D, L = int(os.environ.get("PT_ROWS", "80000")), 1000, 2048, 50000
B = int(os.environ.get("PT_STEPS", "1000"))
variants = [int(v) for v in os.environ.get("PT_VARIANTS", "2,3").split(",")]
kinds = os.environ.get("PT_KINDS", "dense,factored").split(",")
X = torch.relu(torch.randn(N, D, device=dev, generator=g))
xn = eng.row_norm2(X)
PEAK = 6569.6
for fac in (False, True):
    if ("factored" if fac else "dense") not in kinds:
        continue
    XA = xan = None
    if fac:
        XA, xan = eng.badge_factors(torch.randn(N, C, device=dev, generator=g) * 3, 128)
    mind = torch.rand(N, device=dev, generator=g) * 1000
    us = np.random.default_rng(0).random(B)
    vpos = torch.arange(L, L + N, dtype=torch.int32, device=dev)
    ref = None
    for variant in variants:
        for rep in range(2):
            m = mind.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            picks, ms = eng.greedy_select(X, xn, m, [0, N], [B], a=XA, an=xan, uniforms=us if fac else None,
                                          vpos=vpos if fac else None, full_n=[N + L] if fac else None, time_steps=True,
                                          variant=variant)
            e1.record()
            torch.cuda.synchronize()
        loop_us = e0.elapsed_time(e1) * 1e3 / (B - 1)
        row_bytes = 4 * D + 12 + (4 * C if fac else 0)
        tm = eng.last_greedy_timing
        if ref is None:
            ref = picks
        print(json.dumps({"kind": "factored" if fac else "dense", "variant": variant, "steps": B,
                          "loop_us_per_step": round(loop_us, 2), "loop_frac": round(N * row_bytes / loop_us / 1e3 / PEAK, 4),
                          "stream_us": round(tm["stream_ms"] * 1e3, 2), "select_us": round(tm["select_ms"] * 1e3, 2),
                          "stream_frac": round(N * row_bytes / max(tm["stream_ms"], 1e-9) / 1e6 / PEAK, 4),
                          "select_phases_us": [round(v * 1e3, 2) for v in tm["select_phases_ms"]],
                          "picks_equal_first_variant": bool(np.array_equal(ref, picks)),
                          "tile_kb": os.environ.get("ALQ_TILE_KB"), "max_stages": os.environ.get("ALQ_MAX_STAGES")}), flush=True)
