for l in openmines_b200/libopenmines_b200.so variants/lib_r_n_all.so variants/lib_r_both_all.so; do tools/variant_bench.sh $l c3; done
for l in openmines_b200/libopenmines_b200.so variants/lib_r_d_all.so variants/lib_r_both_all.so; do tools/variant_bench.sh $l c5; done
tools/variant_bench.sh openmines_b200/libopenmines_b200.so c4 --steps 1
