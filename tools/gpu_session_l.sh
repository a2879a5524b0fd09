#!/bin/bash
# last evidence pass of the round on ONE GPU (small outputs): tests, bench, --set full captures per kernel family, smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/l_pytest.log
timeout 300 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name regex:"assemble_kernel|props_kernel|spmv_kernel" --launch-skip 2 --launch-count 3 -o gpurun_out/prof_r1e_asm_spmv python tools/prof_kernels.py spmv > gpurun_out/ncu_l2.log 2>&1
timeout 200 ncu --set full --clock-control none --kernel-name regex:"ilu_half_kernel|restrict_kernel|tail_kernel|rbgs_first_kernel|mdot_kernel|maxpy_kernel" --launch-skip 0 --launch-count 18 -o gpurun_out/prof_r1e_pc python tools/prof_kernels.py pc > gpurun_out/ncu_l3.log 2>&1
timeout 200 ncu --set full --clock-control none --kernel-name regex:"rbgs_kernel" --launch-skip 0 --launch-count 4 -o gpurun_out/prof_r1e_rbgs python tools/prof_kernels.py pc > gpurun_out/ncu_l4.log 2>&1
for r in gpurun_out/prof_r1e_asm_spmv gpurun_out/prof_r1e_pc gpurun_out/prof_r1e_rbgs; do
  ncu -i $r.ncu-rep --page raw --csv > $r.raw.csv 2>/dev/null
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1
du -sh gpurun_out
tail -2 gpurun_out/l_pytest.log; tail -2 gpurun_out/l_smoke.log; tail -2 gpurun_out/ncu_l3.log; tail -1 gpurun_out/ncu_l4.log; cut -c1-300 gpurun_out/l_bench.json
