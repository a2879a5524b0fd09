#!/bin/bash
# One GPU session: tests, bench (both arms), launch list.  usage: gpurun -- bash tools/gpu_session.sh <tag> [pytest-args]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-s}
shift
( timeout 2400 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25 ) > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
( timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 ) > gpurun_out/${TAG}_bench.json
( timeout 1200 python bench.py --impl reference 2>gpurun_out/${TAG}_ref.err | tail -1 ) > gpurun_out/${TAG}_ref.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('value %.2f e2e %.2f ms/step %.1f nits %d lits %d failed %s phase %s launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],sum(d['nits']),sum(d['lits']),d.get('failed'),d['phase_ms'],d['gpu_launches']))
    for k in ('roofline','roofline_spmv','roofline_assembly'):
        r=d[k]; print(k, '%.3f ms  %.0f GB/s frac %.3f'%(r['ms_per_launch'],r['achieved'],r['frac']))
    print('cpu_baseline', d.get('cpu_baseline'))
except Exception as e:
    print('bench ERR', e); print(open('gpurun_out/${TAG}_bench.err').read()[-1500:])
try:
    d=json.loads(open('gpurun_out/${TAG}_ref.json').read().strip().splitlines()[-1])
    print('REF value %.3f ms/step %.0f nits %d lits %d same_config %s cores %d roof %s'%(d['value'],d['ms_per_step'],sum(d['nits']),sum(d['lits']),d['config']['same_config'],d['cpu_baseline']['cores'],d['cpu_baseline']['roofline']))
except Exception as e:
    print('ref ERR', e); print(open('gpurun_out/${TAG}_ref.err').read()[-1500:])
PY
