#!/bin/bash
# quick bench line of BASELINE config 3: ms/step + phases
python bench.py --config c3 $BQ3_ARGS --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.3f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['phases_ms_per_step'].items()}, {k: round(v,3) for k,v in d['roofline'].get('poisson_ms_per_step',{}).items()}, d['roofline']['kernel'])"
