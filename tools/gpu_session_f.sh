#!/bin/bash
# final evidence for the round: tests, bench, launch list and one --set full capture (single GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/f_pytest.log
timeout 400 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 4000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"assemble_kernel|props_kernel|spmv_kernel|ilu_half_kernel|restrict_kernel|tail_kernel|rbgs_kernel|mdot_kernel" --launch-skip 0 --launch-count 60 -o gpurun_out/prof_r1d python tools/prof_kernels.py > gpurun_out/ncu_f2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
tail -3 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; tail -2 gpurun_out/ncu_f2.log
