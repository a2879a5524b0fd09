#!/bin/bash
# Multi-GPU session: gpurun --gpus N -- bash tools/gpu_session_multi.sh <tag> <N> [sanitize]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-m}; N=${2:-2}; SAN=${3:-}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  ( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest_multi.log
  tail -4 gpurun_out/${TAG}_pytest_multi.log
else
  ( timeout 600 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep "rank " | cut -c1-260 ) > gpurun_out/${TAG}_check_n$N.log
  ( timeout 900 $TR --master-port 29512 tests/mgpu_bench_check.py 3 2>&1 | grep "rank " | cut -c1-400 ) > gpurun_out/${TAG}_benchcheck_n$N.log
  cat gpurun_out/${TAG}_check_n$N.log gpurun_out/${TAG}_benchcheck_n$N.log
fi
( timeout 900 $TR --master-port 29513 bench.py --gpus $N 2>gpurun_out/${TAG}_bench_n$N.err | tail -1 ) > gpurun_out/${TAG}_bench_n$N.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N value %.2f e2e %.2f ms/step %.1f nits %d lits %d failed %s phase %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],sum(d['nits']),sum(d['lits']),d.get('failed'),d['phase_ms']))
except Exception as e:
    print('bench ERR', e); print(open('gpurun_out/${TAG}_bench_n$N.err').read()[-1500:])
PY
for mode in $EXTRA_SCALES; do
  ( timeout 900 $TR --master-port 29515 bench.py --gpus $N --no-cpu --scale $mode 2>gpurun_out/${TAG}_bench_${mode}_n$N.err | tail -1 ) > gpurun_out/${TAG}_bench_${mode}_n$N.json
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_${mode}_n$N.json').read().strip().splitlines()[-1])
    print('N=$N $mode value %.2f e2e %.2f ms/step %.1f nits %d lits %d failed %s phase %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],sum(d['nits']),sum(d['lits']),d.get('failed'),d['phase_ms']))
except Exception as e:
    print('bench $mode ERR', e); print(open('gpurun_out/${TAG}_bench_${mode}_n$N.err').read()[-1500:])
PY
done
if [ -n "$SAN" ]; then
  for tool in memcheck racecheck; do
    ( timeout 1500 $TR --master-port 29514 --no-python compute-sanitizer --tool $tool --print-limit 20 python tests/mgpu_check.py 2>&1 | grep -E "rank |ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" | cut -c1-260 | head -60 ) > gpurun_out/${TAG}_sanitizer_${tool}_n$N.log
    tail -6 gpurun_out/${TAG}_sanitizer_${tool}_n$N.log
  done
fi
