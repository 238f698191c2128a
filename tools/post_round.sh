#!/bin/bash
# after `gpurun -- bash tools/gpu_round.sh TAG` (same build still in the tree): summaries for profiles/
TAG=$1; NOTE=${2:-}
SO=openmines_b200/libopenmines_b200.so
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep "${TAG}: ${NOTE} om_run_kernel<philox>, C3 shard (16384 instances). ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-counters" > profiles/${TAG}_ncu_summary.md
python tools/sass_attrib.py gpurun_out/${TAG}_prof.ncu-rep $SO _Z13om_run_kernelILi1ELb0ELb0ELb0EEv9RunParams >> profiles/${TAG}_ncu_summary.md
if [ -f gpurun_out/${TAG}_step_prof.ncu-rep ]; then
python tools/ncu_summary.py gpurun_out/${TAG}_step_prof.ncu-rep "${TAG}: ${NOTE} om_step_kernel<philox>, one launch in the middle of a C4 episode (16384 envs). ncu --set full --clock-control none --import-source on -k regex:om_step_kernel -s 200 -c 1 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-counters" > profiles/${TAG}_step_ncu_summary.md
python tools/sass_attrib.py gpurun_out/${TAG}_step_prof.ncu-rep $SO _Z14om_step_kernelILi1ELb0EEv9RunParams >> profiles/${TAG}_step_ncu_summary.md
fi
if [ -f gpurun_out/${TAG}_c5_prof.ncu-rep ]; then
python tools/ncu_summary.py gpurun_out/${TAG}_c5_prof.ncu-rep "${TAG}: ${NOTE} om_run_kernel<philox, split>, C5 shard (1024 instances of the 500-truck mine). ncu --set full --clock-control none --import-source on -k regex:om_run_kernel -s 1 -c 1 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-counters" > profiles/${TAG}_c5_ncu_summary.md
python tools/sass_attrib.py gpurun_out/${TAG}_c5_prof.ncu-rep $SO _Z13om_run_kernelILi1ELb0ELb1ELb0EEv9RunParams >> profiles/${TAG}_c5_ncu_summary.md
fi
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
for w in c3 c2 c4 c5 ablation reference; do [ -s gpurun_out/${TAG}_bench_$w.json ] && cp gpurun_out/${TAG}_bench_$w.json profiles/${TAG}_bench_$w.json; done
cp gpurun_out/${TAG}_pytest.log profiles/${TAG}_pytest.log
grep -E "duration|inst_executed.sum |issue_active|no_instruction|dram__bytes" profiles/${TAG}_ncu_summary.md
