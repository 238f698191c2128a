#!/bin/bash
TAG=${1:-rX}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --workload c3 --steps 5 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_c3_n8.json 2> gpurun_out/${TAG}_bench_c3_n8.err
echo "c3 n8 rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_c3_n8.json').read().strip().split(chr(10))[-1]); print('value %.4e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d['n_gpus'])"
