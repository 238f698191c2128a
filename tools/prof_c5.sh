#!/bin/bash
TAG=${1:-rX}
ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 -o gpurun_out/${TAG}_c5_prof -f \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c5_ncu.log 2>&1; echo "ncu rc=$?"
