"""Step-kernel tile/stage sweep (CUDA-event mean over a 400-step loop)."""
import os, sys, subprocess, json
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from active_learning_b200.engine import Engine
    eng = Engine(0); dev = eng.device
    g = torch.Generator(device=dev).manual_seed(0)
    N, C, D, B = 80000, 1000, 2048, 400
    X = torch.relu(torch.randn(N, D, device=dev, generator=g)); xn = eng.row_norm2(X)
    out = {}
    for fac in (False, True):
        XA = xan = None
        if fac:
            XA, xan = eng.badge_factors(torch.randn(N, C, device=dev, generator=g) * 3, 128)
        mind = torch.rand(N, device=dev, generator=g) * 1000
        us = np.random.default_rng(0).random(B)
        vpos = torch.arange(50000, 50000 + N, dtype=torch.int32, device=dev)
        for _ in range(2):
            _, ms = eng.greedy_select(X, xn, mind.clone(), [0, N], [B], a=XA, an=xan, uniforms=us if fac else None,
                                      vpos=vpos if fac else None, full_n=[N + 50000] if fac else None, time_steps=True)
        out["factored" if fac else "dense"] = round(ms * 1e3, 2)
    print(json.dumps(out))
else:
    for kb, st in ((16, 16), (8, 16), (12, 16), (24, 16), (32, 16), (48, 16), (16, 8), (16, 12), (64, 16)):
        env = dict(os.environ, ALQ_TILE_KB=str(kb), ALQ_MAX_STAGES=str(st))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(f"tile {kb:3d} KB  max stages {st:2d}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]}", flush=True)
