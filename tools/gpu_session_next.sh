#!/bin/bash
# First GPU session of the next round (run with `gpurun --gpus 2 -- bash tools/gpu_session_next.sh 2`, then 4 / 8):
# the diagonal-dominance stop of the multigrid hierarchies on slabs (TPB_MG_DD_DIST=1, written in r1 but never run on
# more than one GPU).  Consistency check both ways, then bench A/B.  If the check is green and the bench gains what
# N=1 gained (x1.24), make it the default (csrc/tpb_pc.cu dd_on_slabs) and drop the environment switch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( TPB_MG_DD_DIST=1 timeout 300 $TR --master-port 29611 tests/mgpu_check.py 2>&1 | grep "rank " | cut -c1-230 ) > gpurun_out/n_check_dd_n$N.log
( TPB_MG_DD_DIST=1 TPB_MG_GATHER=300 timeout 300 $TR --master-port 29612 tests/mgpu_check.py 2>&1 | grep "rank " | cut -c1-230 ) > gpurun_out/n_check_dd_g300_n$N.log
( TPB_MG_DD_DIST=1 timeout 400 $TR --master-port 29613 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/n_bench_dd_n$N.err | tail -1 ) > gpurun_out/n_bench_dd_n$N.json
( timeout 400 $TR --master-port 29614 bench.py --gpus $N --steps 6 --warmup 3 2>gpurun_out/n_bench_n$N.err | tail -1 ) > gpurun_out/n_bench_n$N.json
cat gpurun_out/n_check_dd_n$N.log gpurun_out/n_check_dd_g300_n$N.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/n_bench*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f e2e %.2f ms/step %.1f nits %s lits %s failed %s ksp %.0f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['nits'],d.get('lits'),d.get('failed'),d['phase_ms']['ksp']))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
