for dyn in "" 0 1; do echo "DYN=$dyn"; FASTCHEB_FIT_DYN=$dyn FIT_ONLY=65-float python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-170; done
