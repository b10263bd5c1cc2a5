timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/probe_interp.py 2>&1 | head -6
python tools/latency_probe.py 2>&1 | tail -9
