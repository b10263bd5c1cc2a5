for o in 1 2 4 8; do echo "oversub $o"; FASTCHEB_PB_OVERSUB=$o python tools/probe_pb1d.py 2>&1 | grep -E "^cfg"; done
