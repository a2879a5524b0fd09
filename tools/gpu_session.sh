#!/bin/bash
# One GPU session: tests, bench, launch list.  usage: gpurun -- bash tools/gpu_session.sh <tag> [pytest-args]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-s}
shift
( timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu 2>gpurun_out/${TAG}_bench.err | tail -1 ) > gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_pytest.log
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value %.2f e2e %.2f ms/step %.1f nits %s lits %s failed %s phase %s launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['nits'],d.get('lits'),d.get('failed'),d['phase_ms'],d['gpu_launches']))
    for k in ('roofline','roofline_assembly','roofline_dominant_by_share'):
        r=d[k]; print(k, '%.3f ms  %.0f GB/s frac %.3f'%(r['ms_per_launch'],r['achieved'],r['frac']))
except Exception as e:
    print('bench ERR', e); print(open('gpurun_out/${TAG}_bench.err').read()[-1500:])
PY
