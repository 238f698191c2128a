#!/bin/bash
# occupancy sweep of the episode kernel: WARPS="18 20 22" INSTANCES=16384 bash tools/occ_sweep.sh
for w in ${WARPS:-4 8 12 17 20 22}; do OM_WARPS_PER_CTA=$w timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${INSTANCES:+--instances $INSTANCES} 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('warps',d['config']['launch']['warps_per_cta'],'instances',d['config']['instances_per_gpu'],'value %.4e'%d['value'],'kernel_ms %.2f'%d['roofline']['kernel_ms'])"; done
