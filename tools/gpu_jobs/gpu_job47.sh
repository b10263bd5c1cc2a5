# randomized stress with fresh seeds after the gradient-kernel change (product basis a: every n; b-e: fits, shared points, N-d)
mkdir -p gpurun_out
for s in ${STRESS_SEEDS:-11 12 13}; do echo "seed $s"; STRESS_SEED=$s timeout 600 python tools/stress_r02.py 2>&1 | tail -7; done | tee gpurun_out/r02_stress_job47.txt
for o in ${FUZZ_OFFSETS:-101 202}; do echo "fuzz offset $o"; FASTCHEB_FUZZ_OFFSET=$o timeout 600 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2; done | tee -a gpurun_out/r02_stress_job47.txt
