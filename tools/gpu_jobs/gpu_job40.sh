FIT_HOWMANY=16384 FASTCHEB_FITND_PROF=1 FIT_ONLY=65x65-float64 python tools/probe_fit_nd.py 2>&1 | tail -9 | cut -c1-160
