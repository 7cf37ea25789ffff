#!/bin/bash
# the driver's scaling sequence on one 8-GPU box: N = 1, 2, 4, 8
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 exit=$?"
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "N=$N exit=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open(f"gpurun_out/scale_n{n}.json"))
        w=d.get("workloads",{})
        print(n, round(d["value"]/1e6,1),"M/s", round(d["ms_per_step"]*1e3,1),"us", "k1",round(d["roofline"]["frac"],3),
              {k:(round(v.get("ms_per_step",0),1), round(v.get("breakdown_ms",{}).get("loop_ms",0)/10,1)) for k,v in w.items()})
    except Exception as e:
        print(n,"ERR",e)
PY
