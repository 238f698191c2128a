#!/bin/bash
# Full GPU check: parity tests, all bench workloads, launch list and ncu --set full captures of the episode kernel (C3 shard),
# the step kernel (C4) and the split episode kernel (C5).
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout -k 10 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout -k 10 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; echo "bench c5 rc=$?"
timeout -k 10 600 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
timeout -k 10 600 python bench.py --workload ablation --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ablation.json 2> gpurun_out/${TAG}_bench_ablation.err; echo "bench ablation rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-counters --no-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-counters --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:om_step_kernel -s 200 -c 1 -o gpurun_out/${TAG}_step_prof -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-counters > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "ncu step rc=$?"
ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 -o gpurun_out/${TAG}_c5_prof -f \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-counters > gpurun_out/${TAG}_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
cat gpurun_out/${TAG}_bench_c3.json
