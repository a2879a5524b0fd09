#!/bin/bash
# assembly kernel matrix (one GPU): thread-per-cell vs plane-marching tile shapes, with/without prefetch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/j_matrix.log
for cfg in 0 1 2 3 4 11 12 13 14; do
  echo "== TPB_ASM_TILE=$cfg" >> gpurun_out/j_matrix.log
  TPB_ASM_TILE=$cfg timeout 120 python tools/microbench_one.py 2>&1 | grep assemble >> gpurun_out/j_matrix.log
done
for kz in 6 9 12; do
  echo "== TPB_ASM_TILE=2 KZ=$kz" >> gpurun_out/j_matrix.log
  TPB_ASM_TILE=2 TPB_ASM_KZ=$kz timeout 120 python tools/microbench_one.py 2>&1 | grep assemble >> gpurun_out/j_matrix.log
done
(TPB_ASM_TILE=12 timeout 300 python -m pytest tests/test_gpu_assembly.py -m gpu -x -q 2>&1 | tail -2) >> gpurun_out/j_matrix.log
cat gpurun_out/j_matrix.log
