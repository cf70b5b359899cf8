#!/bin/bash
# last sanity run of the final tree
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2 )
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('step %.2f e2e %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['state_checksum']['T'])"
