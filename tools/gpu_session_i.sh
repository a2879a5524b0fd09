#!/bin/bash
# A/B of the plane-marching shared-memory assembly kernel against the thread-per-cell kernel (one GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/i_pytest.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/i_bench_tile.json 2> gpurun_out/i_bench_tile.err
TPB_ASM_TILE=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/i_bench_cell.json 2> gpurun_out/i_bench_cell.err
for kz in 9 11 29 43; do
TPB_ASM_KZ=$kz timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/i_bench_kz$kz.json 2> gpurun_out/i_bench_kz$kz.err
done
tail -4 gpurun_out/i_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f e2e %.2f asm %.4f ms frac %.3f spmv frac %.3f phase %s'%(d['value'],d['e2e']['value'],d['roofline_assembly']['ms_per_launch'],d['roofline_assembly']['frac'],d['roofline']['frac'],d['phase_ms']))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-500:])
PY
