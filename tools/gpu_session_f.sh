#!/bin/bash
# evidence for the round on ONE GPU: tests, smoke, bench, launch list, --set full captures of the hot kernels.
# gpurun copies back at most 64 MiB of gpurun_out/: the full captures are limited to a handful of launches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
if [ "${1:-full}" = "full" ]; then
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 3000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"assemble_kernel|props_kernel|spmv_kernel" --launch-skip 2 --launch-count 3 -o gpurun_out/prof_r1d_asm_spmv python tools/prof_kernels.py spmv > gpurun_out/ncu_f2.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name regex:"ilu_half_kernel|restrict_kernel|tail_kernel|rbgs_first_kernel|mdot_kernel|maxpy_kernel" --launch-skip 0 --launch-count 18 -o gpurun_out/prof_r1d_pc python tools/prof_kernels.py pc > gpurun_out/ncu_f3.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name regex:"rbgs_kernel" --launch-skip 0 --launch-count 4 -o gpurun_out/prof_r1d_rbgs python tools/prof_kernels.py pc > gpurun_out/ncu_f4.log 2>&1
for r in gpurun_out/prof_r1d_asm_spmv gpurun_out/prof_r1d_pc gpurun_out/prof_r1d_rbgs; do
  ncu -i $r.ncu-rep --page raw --csv > $r.raw.csv 2>/dev/null
done
fi
du -sh gpurun_out
tail -3 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; tail -2 gpurun_out/ncu_f3.log; cut -c1-400 gpurun_out/f_bench.json
