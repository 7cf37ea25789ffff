#!/bin/bash
mkdir -p gpurun_out
ALQ_PERSIST_DEBUG=2 PT_VARIANTS=3 PT_KINDS=factored PT_STEPS=400 timeout 300 python tools/persist_time.py > gpurun_out/r2i.jsonl 2> gpurun_out/r2i_skew.err
cat gpurun_out/r2i.jsonl; grep -c skew gpurun_out/r2i_skew.err
