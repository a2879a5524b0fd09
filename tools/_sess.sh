cd "$GRAFT_REPO_ROOT"
python tools/dbg_asm.py 2>&1 | tail -14
for v in 1 0; do echo "== TPB_GRAPH_KEEP=$v"; TPB_GRAPH_KEEP=$v python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['phase_ms'], d['e2e']['ms_per_step'], d['ms_per_step'])"; done
