echo base; python tools/probe_pb1d.py 2>&1 | grep cfg4_f64_val
for t in 256 128 96 64; do echo "P4 threads $t"; FASTCHEB_PB_P4=1 FASTCHEB_PB_THREADS=$t python tools/probe_pb1d.py 2>&1 | grep cfg4_f64_val; done
