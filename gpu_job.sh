export BENCH_ONLY="cfg2"
run() { timeout 300 python tools/bench_configs.py 2>/dev/null | grep '"dmma"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$1', d['mode'], round(d['frac_of_fp_peak'], 4))"; }
run good
cp fastchebinterp.jl_b200/libfastcheb_cuda.so /tmp/good.so
cp tmp_exp/libfake4.so fastchebinterp.jl_b200/libfastcheb_cuda.so; run fake4_of_16_steps
cp tmp_exp/libfake1.so fastchebinterp.jl_b200/libfastcheb_cuda.so; run fake1_of_16_steps
