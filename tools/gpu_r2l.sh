#!/bin/bash
timeout 900 python tools/pool_forward_time.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prefetch, loader args as the reference', {k:(round(v['value']),round(v['ms_per_query'])) for k,v in d.items() if isinstance(v,dict) and 'value' in v})"
ALQ_PREFETCH=0 timeout 900 python tools/pool_forward_time.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no prefetch', {k:(round(v['value']),round(v['ms_per_query'])) for k,v in d.items() if isinstance(v,dict) and 'value' in v})"
