for w in 4 8 12 17; do OM_WARPS_PER_CTA=$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('warps',d['config']['launch']['warps_per_cta'],'value %.3e'%d['value'],'kernel_ms',d['roofline']['kernel_ms'])"; done
