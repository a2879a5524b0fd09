#!/bin/bash
# usage: tools/bench_short.sh [bench args]  -> one-line summary
python bench.py --steps 6 --warmup 3 --no-cpu "$@" 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f e2e %.2f ms/step %.1f nits %s lits %d ksp_ms %.0f setup_ms %.0f asm_ms %.1f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], sum(d['nits']), sum(d['lits']), d['phase_ms']['ksp'], d['phase_ms']['pc_setup'], d['phase_ms']['assemble'], d['gpu_launches']))"
