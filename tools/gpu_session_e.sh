#!/bin/bash
# A/B of programmatic dependent launch in the multigrid kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/e_pytest.log
TPB_PDL=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/e_bench_nopdl.json 2> gpurun_out/e_bench_nopdl.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/e_bench_pdl.json 2> gpurun_out/e_bench_pdl.err
TPB_PDL=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/e_bench_nopdl2.json 2> gpurun_out/e_bench_nopdl2.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/e_bench_pdl2.json 2> gpurun_out/e_bench_pdl2.err
tail -3 gpurun_out/e_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f e2e %.2f ms/step %.1f lits %s ksp %.0f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['lits'],d['phase_ms']['ksp']))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
