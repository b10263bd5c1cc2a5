for o in 1 2 4; do echo "fit oversub $o"; FASTCHEB_FIT_OVERSUB=$o python tools/probe_fit1d.py 2>&1 | tail -4; done
