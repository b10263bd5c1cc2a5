for o in 1 2 4; do echo "oversub $o"; FASTCHEB_CL_OVERSUB=$o FASTCHEB_LO1D=0 FASTCHEB_LOND=0 python tools/probe_loworder.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(f\"  {d['leg']:28s} {d['points_per_s']:.3e} hbm {d['frac_hbm']:.3f} fp {d['frac_fp']:.3f}\")
"; done
