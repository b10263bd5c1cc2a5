set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu 2>&1 | tail -15
BENCH_ONLY="2-D" timeout 300 python tools/bench_configs.py > gpurun_out/configs_2d.jsonl 2> gpurun_out/configs_2d.err
cut -c1-260 gpurun_out/configs_2d.jsonl
tail -3 gpurun_out/configs_2d.err
