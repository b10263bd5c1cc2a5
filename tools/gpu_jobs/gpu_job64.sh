for o in 1 2 4 8; do echo "oversub $o"; FASTCHEB_LO_OVERSUB=$o python tools/probe_loworder.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(f\"  {d['leg']:28s} {d['points_per_s']:.3e} hbm {d['frac_hbm']:.3f} fp {d['frac_fp']:.3f}\")
"; FASTCHEB_MANY_OVERSUB=$o python tools/probe_many.py 2>&1 | tail -6; done
