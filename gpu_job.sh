BENCH_ONLY="cfg1" BENCH_NPTS=1000000 python tools/bench_configs.py 2>&1 | tail -5 | cut -c1-400
