#!/bin/bash
# Bench one workload with a variant build of the engine: tools/variant_bench.sh LIB.so WORKLOAD [bench args]
LIBV=$1; WL=$2; shift 2
OM_LIB_PATH=$LIBV timeout -k 10 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-counters --no-extras "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d.get('roofline',{})
print('$LIBV $WL', d['config'].get('launch'), 'value %.4e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.4f'%r.get('kernel_ms',0), 'errs', d.get('instance_errors'))"
