#!/bin/bash
# Memory-safety pass over every kernel family (GPU box): bash tools/sanitize.sh TAG   -> gpurun_out/TAG_sanitize_*.log
# 1. plain run of tools/sanitize_target.py (every leg checks its results against the oracle / goldens);
# 2. the same under the bounds-checking build (OM_DEBUG_LIB=1: -DOM_BOUNDS_CHECK, every state access checks its index);
# 3. compute-sanitizer memcheck / racecheck / initcheck / synccheck where the pool allows it (only this repo's kernels are
#    instrumented: --kernel-name regex:om_).
TAG=${1:-rX}
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_sanitize_plain.log; exit 1; }
echo "plain: $(tail -1 gpurun_out/${TAG}_sanitize_plain.log)"
OM_DEBUG_LIB=1 python -c "from openmines_b200 import _lib; _lib.build()" && OM_DEBUG_LIB=1 python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_bounds.log 2>&1
echo "bounds rc=$? : $(tail -1 gpurun_out/${TAG}_sanitize_bounds.log)"
# 2b. warp-convergence probes alone (production code + __activemask() checks): at scale
OM_DEBUG_LIB=probes OM_EXTRA_NVCC_FLAGS="-DOM_CONV_PROBES" OM_LIB_PATH=$PWD/openmines_b200/libprobes.so python -c "from openmines_b200 import _lib; _lib.build()" && \
  OM_LIB_PATH=$PWD/openmines_b200/libprobes.so python tools/conv_probe_run.py > gpurun_out/${TAG}_sanitize_probes.log 2>&1
echo "probes rc=$? : $(grep -c 'convergence violations 0 ' gpurun_out/${TAG}_sanitize_probes.log) of 3 workloads with 0 convergence violations"
[ -n "$SKIP_COMPUTE_SANITIZER" ] && exit 0
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout -k 10 ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --kernel-name regex:om_ --print-limit 30 --error-exitcode 7 \
      python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|closed' gpurun_out/${TAG}_sanitize_${tool}.log | tail -1)"
done
