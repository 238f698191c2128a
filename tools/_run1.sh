timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "draw_tape or marginals or rl_step or det_all" 2>&1 | tail -2
for l in openmines_b200/libopenmines_b200.so variants/lib_prev.so; do tools/variant_bench.sh $l c3; tools/variant_bench.sh $l c5; tools/variant_bench.sh $l c4 --steps 1; tools/variant_bench.sh $l c2; done
