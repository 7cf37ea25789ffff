#!/bin/bash
# L2-residency sweep of the persistent selection loop (evict_last hints on the head of every CTA's segment)
for mb in 0 48 72 96 112; do
  echo "== 10000 rows, resident $mb MB"; ALQ_L2_RESIDENT_MB=$mb PT_ROWS=10000 PT_VARIANTS=3 PT_STEPS=1500 timeout 200 python tools/persist_time.py 2>/dev/null | cut -c1-210
done
for mb in 0 64 96; do
  echo "== 80000 rows, resident $mb MB"; ALQ_L2_RESIDENT_MB=$mb PT_VARIANTS=3 PT_STEPS=600 timeout 200 python tools/persist_time.py 2>/dev/null | cut -c1-210
done
timeout 300 python tools/pool_forward_time.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pool fwd', {k:(round(v['value']),round(v['ms_per_query'])) for k,v in d.items() if isinstance(v,dict) and 'value' in v})"
