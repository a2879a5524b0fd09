#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 2>&1 | tail -2 ) > gpurun_out/b_bench_n$N.log 2>&1
if [ -n "$2" ]; then ( timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 6 --warmup 3 --scale refine 2>&1 | tail -2 ) > gpurun_out/b_bench_n${N}_refine.log 2>&1; fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b_bench_n*.log')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value %.1f ms/step %.1f nits %s lits %s failed %s dt %s phase %s'%(d['value'],d['ms_per_step'],d['nits'],d['lits'],d['failed'],['%.2e'%x for x in d['dt_days']],{k:round(v) for k,v in d['phase_ms'].items()}))
PY
