#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sp in 0 1; do OM_SPLIT=$sp python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('split $sp', d['config']['launch'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'state', d['config']['state_bytes_per_instance'], 'errs', d['instance_errors'])"; done
SKIP_TESTS=1 bash tools/quick_bench.sh
