#!/bin/bash
# multi-rank multigrid check: regressions on one GPU, 2-rank consistency check at two gather thresholds, N=2 bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/a_pytest.log 2>&1
( timeout 300 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^\*\*\|OMP_NUM" | tail -12 ) > gpurun_out/a_mgpu_default.log 2>&1
( TPB_MG_GATHER=300 timeout 300 $TR --master-port 29512 tests/mgpu_check.py 2>&1 | grep -v "^\*\*\|OMP_NUM" | tail -12 ) > gpurun_out/a_mgpu_g300.log 2>&1
( timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | tail -3 ) > gpurun_out/a_bench_n2.log 2>&1
( timeout 400 python bench.py --gpus 1 --mult 2 --steps 6 --warmup 3 --no-cpu 2>&1 | tail -3 ) > gpurun_out/a_bench_n1x2.log 2>&1
( TPB_GRAPH_NCCL=1 timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | tail -3 ) > gpurun_out/a_bench_n2_graphnccl.log 2>&1
tail -n 4 gpurun_out/a_*.log
