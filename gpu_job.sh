for e in 0 1 2 3; do echo "EXP $e"; ./bench_micro/dmma_var_exp$e 4194304 | cut -c1-220; done
