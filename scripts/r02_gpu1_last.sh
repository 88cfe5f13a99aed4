#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_faces.py tests/test_gpu_parity.py -m gpu -q -k "order or golden or qft" 2>&1 | tail -2
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['passes_per_step'], d['norm2_after'])"
