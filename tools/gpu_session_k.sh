#!/bin/bash
# bench only at N ranks (default exchanges)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/k_bench_n$N.err | tail -1 ) > gpurun_out/k_bench_n$N.json
python - <<PY
import json
f='gpurun_out/k_bench_n$N.json'
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.2f e2e %.2f ms/step %.1f nits %s lits %s failed %s spmv %.4f ms | %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['nits'],d.get('lits'),d.get('failed'),d['roofline']['ms_per_launch'],d['config'].get('exchanges','')[:40]))
except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-1500:])
PY
