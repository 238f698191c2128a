timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c3 c5 c2; do tools/variant_bench.sh openmines_b200/libopenmines_b200.so $w; done
tools/variant_bench.sh openmines_b200/libopenmines_b200.so c4 --steps 1
