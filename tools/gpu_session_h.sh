#!/bin/bash
# peer-memory exchanges at N ranks: consistency check (both gather thresholds), bench per TPB_P2P mask
# (15 = everything fused, 7 = mailboxes with separate halo kernels, 0 = NCCL only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
MASKS=${2:-"15 7 0"}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep "rank " | cut -c1-230 ) > gpurun_out/h_check_n$N.log
( TPB_MG_GATHER=300 timeout 300 $TR --master-port 29512 tests/mgpu_check.py 2>&1 | grep "rank " | cut -c1-230 ) > gpurun_out/h_check_g300_n$N.log
port=29520
for m in $MASKS; do
  port=$((port+1))
  ( TPB_P2P=$m timeout 400 $TR --master-port $port bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/h_bench_p$m\_n$N.err | tail -1 ) > gpurun_out/h_bench_p$m\_n$N.json
done
cat gpurun_out/h_check_n$N.log gpurun_out/h_check_g300_n$N.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/h_bench_p*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f e2e %.2f ms/step %.1f lits %s failed %s spmv %.4f ms asm %.4f ms ksp %.0f | %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d.get('lits'),d.get('failed'),d['roofline']['ms_per_launch'],d['roofline_assembly']['ms_per_launch'],d['phase_ms']['ksp'],d['config'].get('exchanges','')[:40]))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
