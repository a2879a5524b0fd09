cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for v in "TPB_X=1" "TPB_MG_GATHER=8"; do
  env $v $TR --master-port 29513 bench.py --gpus 2 --no-cpu 2>gpurun_out/r2l_$v.err | tail -1 > gpurun_out/r2l_$v.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_$v.json').read().strip().splitlines()[-1])
print('$v value %.1f e2e %.1f ms/step %.1f nits %d lits %d phase %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], sum(d['nits']), sum(d['lits']), d['phase_ms']))
PY
done
