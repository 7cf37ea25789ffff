import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from active_learning_b200.engine import Engine
    eng = Engine(0); dev = eng.device
    g = torch.Generator(device=dev).manual_seed(0)
    lg = torch.randn(80000, 1000, device=dev, generator=g) * 3
    out = torch.empty(80000, device=dev)
    for _ in range(5): eng.score_softmax(lg, 0, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): eng.score_softmax(lg, 0, out=out)
    e1.record(); torch.cuda.synchronize()
    k1 = e0.elapsed_time(e1) / 200 * 1e3
    a = None
    for _ in range(3): a = eng.badge_factors(lg, 128)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): a = eng.badge_factors(lg, 128)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"k1_us": round(k1, 2), "k2_us": round(e0.elapsed_time(e1) / 50 * 1e3, 2)}))
else:
    for kb, sp in ((16, 1), (16, 2), (16, 4), (8, 1), (8, 2), (12, 2), (12, 3), (24, 2), (24, 3), (32, 4)):
        env = dict(os.environ, ALQ_ROW_TILE_KB=str(kb), ALQ_ROW_SPLIT=str(sp))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(f"row tile {kb:3d} KB split {sp}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]}", flush=True)
