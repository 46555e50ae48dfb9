#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()}, d['combine']['value'])"
