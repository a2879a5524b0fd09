cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f e2e %.2f ms/step %.1f lits %d phase %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],sum(d['lits']),d['phase_ms']))"
bash tools/gpu_evidence.sh r2j 22
du -sh gpurun_out
