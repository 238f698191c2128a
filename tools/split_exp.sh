for w in 17 20; do OM_SPLIT=1 OM_WARPS_PER_CTA=$w SKIP_TESTS=1 bash tools/quick_bench.sh; done
