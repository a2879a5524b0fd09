#!/bin/bash
# peer-memory mailboxes at N ranks: consistency check (both gather thresholds), bench with the mailboxes and NCCL-only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*" | tail -6 ) > gpurun_out/h_check_n$N.log
( TPB_MG_GATHER=300 timeout 300 $TR --master-port 29512 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*" | tail -6 ) > gpurun_out/h_check_g300_n$N.log
( timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/h_bench_n$N.err | tail -1 ) > gpurun_out/h_bench_n$N.json
if [ "${2:-ab}" = "ab" ]; then
( TPB_P2P=0 timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/h_bench_nccl_n$N.err | tail -1 ) > gpurun_out/h_bench_nccl_n$N.json
fi
cat gpurun_out/h_check_n$N.log gpurun_out/h_check_g300_n$N.log | cut -c1-220
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/h_bench*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f e2e %.2f ms/step %.1f nits %s lits %s failed %s | %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d.get('nits'),d.get('lits'),d.get('failed'),d['config'].get('exchanges')))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
