#!/bin/bash
# launch list of one late time step: gpurun -- bash tools/gpu_prof.sh <tag> [W] [opts...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-p}; W=${2:-12}; shift; shift
python tools/prof_step.py $W "$@" > gpurun_out/${TAG}_plain.log 2>&1 || { tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_step.py $W "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_summary.md
cat gpurun_out/${TAG}_summary.md
