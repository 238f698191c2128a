#!/bin/bash
# one ncu --set full capture of the episode kernel on the C3 shard (run after a plain bench has exited 0)
TAG=${1:-rX}
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
