#!/bin/bash
# compute-sanitizer pass over every kernel family (GPU box): bash tools/sanitize.sh TAG   -> gpurun_out/TAG_sanitize_*.log
# Only this repo's kernels are instrumented (--kernel-name regex:om_): torch's own fill / copy kernels are not our subject.
TAG=${1:-rX}
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_sanitize_plain.log; exit 1; }
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout -k 10 ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --kernel-name regex:om_ --print-limit 30 --error-exitcode 7 \
      python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_${tool}.log | tail -1)"
done
