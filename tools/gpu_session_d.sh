#!/bin/bash
# single-GPU: tests, bench, then the ncu launch list of the same bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/d_pytest.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/d_bench1.json 2> gpurun_out/d_bench1.err
if [ -n "$1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 4000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_d.log 2>&1
fi
tail -3 gpurun_out/d_pytest.log
