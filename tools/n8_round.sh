#!/bin/bash
# 8-GPU lines of the headline workload and of configs[4] (run with gpurun --gpus 8): bash tools/n8_round.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
for w in c3 c5; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --workload $w --steps 5 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_${w}_n8.json 2> gpurun_out/${TAG}_bench_${w}_n8.err
  echo "$w n8 rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_${w}_n8.json | head -c 400; echo
done
