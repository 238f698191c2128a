for l in openmines_b200/libopenmines_b200.so variants/lib_y_swboth.so variants/lib_y_swdisp.so variants/lib_y_swphase.so openmines_b200/libopenmines_b200.so; do tools/variant_bench.sh $l c3; done
for l in openmines_b200/libopenmines_b200.so variants/lib_y_swboth.so; do tools/variant_bench.sh $l c2; tools/variant_bench.sh $l c5; tools/variant_bench.sh $l c4 --steps 1; done
