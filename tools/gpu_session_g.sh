#!/bin/bash
# multi-rank check: consistency test + bench at N ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then (timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/g_pytest_multi.log; fi
( timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 2>&1 | tail -1 ) > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err
( timeout 200 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/g_bench_ref_n$N.json 2>&1
cat gpurun_out/g_pytest_multi.log 2>/dev/null | tail -2
python - <<PY
import json
for f in ['gpurun_out/g_bench_n$N.json','gpurun_out/g_bench_ref_n$N.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.1f nits %s lits %s failed %s'%(d['value'],d['ms_per_step'],d.get('nits'),d.get('lits'),d.get('failed')))
    except Exception as e: print(f,'ERR',e, open(f).read()[-600:])
PY
