#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "4 stack" "2 refine" "4 refine"; do
  set -- $cfg
  ( timeout 500 python bench.py --gpus 1 --mult $1 --scale $2 --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 ) > gpurun_out/c_bench_n1x$1_$2.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c_bench_*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value %.1f ms/step %.1f nits %s lits %s failed %s dt %s phase %s'%(d['value'],d['ms_per_step'],d['nits'],d['lits'],d['failed'],['%.2e'%x for x in d['dt_days']],{k:round(v) for k,v in d['phase_ms'].items()}))
        elif line.strip(): print(f, line[:300])
PY
