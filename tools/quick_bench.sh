#!/bin/bash
# quick GPU check: parity tests (optional: SKIP_TESTS=1) + C3 bench line
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('warps',d['config']['launch']['warps_per_cta'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'state', d['config']['state_bytes_per_instance'], 'errs', d['instance_errors'])"
