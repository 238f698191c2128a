#!/bin/bash
# quick GPU check: parity tests (optional: SKIP_TESTS=1) + C3 bench line (+ c4/c5 with ALL=1)
if [ -z "$SKIP_TESTS" ]; then timeout -k 10 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c3 warps',d['config']['launch']['warps_per_cta'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'e2e %.4e'%d['e2e']['value'], 'state', d['config']['state_bytes_per_instance'], 'errs', d['instance_errors'])"
if [ -n "$ALL" ]; then
timeout -k 10 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c5', d['config']['launch'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'errs', d['instance_errors'])"
timeout -k 10 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c2', d['config']['launch'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'errs', d['instance_errors'])"
timeout -k 10 300 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c4', d['config']['launch'], d['config']['state_bytes_per_instance'], 'value %.4e'%d['value'],'decisions/s %.4e'%d['decisions_per_s'],'ms/step %.1f'%d['ms_per_step'],'kernel_ms %.4f'%d['roofline']['kernel_ms'],'frac %.3f'%d['roofline']['frac'])"
fi
