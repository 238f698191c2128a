#!/bin/bash
# ncu --set full capture of one om_step_kernel launch in the middle of a C4 episode
TAG=${1:-rX}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4_plain.json 2> gpurun_out/${TAG}_c4_plain.err || exit 1
cat gpurun_out/${TAG}_c4_plain.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c4 value %.4e'%d['value'],'decisions/s %.4e'%d['decisions_per_s'],'ms/step %.1f'%d['ms_per_step'],'kernel_ms %.4f'%d['roofline']['kernel_ms'],'frac %.3f'%d['roofline']['frac'], d['env_steps_per_episode'])"
ncu --set full --clock-control none --import-source on -k regex:om_step_kernel -s 200 -c 1 -o gpurun_out/${TAG}_step_prof -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4_ncu.log 2>&1; echo "ncu rc=$?"
