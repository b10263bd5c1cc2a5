for b in 2 3 4 6; do echo "65x65 BLOCK=$b"; for s in 65x65-float64 65x65-float32; do FASTCHEB_FITND_BLOCK=$b FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | grep "\"$s" | cut -c1-50,95-175; done; done
for b in 6 8 12 16 24; do echo "33x33 BLOCK=$b"; for s in 33x33-float64 33x33-float32; do FASTCHEB_FITND_BLOCK=$b FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | grep "\"$s" | cut -c1-50,95-175; done; done
for b in 2 3 4; do echo "17^3 BLOCK=$b"; FASTCHEB_FITND_BLOCK=$b FIT_ONLY=17x17x17-float64 python tools/probe_fit_nd.py 2>&1 | grep "\"17x17x17-" | cut -c1-50,95-175; done
