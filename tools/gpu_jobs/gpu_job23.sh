for s in 65x65-float64 65x65-float32; do
echo base; FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175
echo ablate-passes; FASTCHEB_FITND_ABLATE=1 FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175
for w in 2 8; do echo warps $w; FASTCHEB_FITND_WARPS=$w FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175; done
done
