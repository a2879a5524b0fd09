#!/bin/bash
# Evidence session on one GPU: sanitizer on the smoke path, launch list of a late time step, --set full captures of the
# top kernels.  usage: gpurun -- bash tools/gpu_evidence.sh <tag> [W]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-e}; W=${2:-22}
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  grep -E "smoke:|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_${tool}.log | cut -c1-200
done
python tools/prof_step.py $W > gpurun_out/${TAG}_plain.log 2>&1; tail -1 gpurun_out/${TAG}_plain.log
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_step.py $W > gpurun_out/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_summary.md; head -14 gpurun_out/${TAG}_summary.md
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"line_smooth_kernel|spmv_kernel|assemble_kernel|props_kernel|ilu_half|restrict_line|mdot_kernel" --launch-count 14 -o gpurun_out/${TAG}_full python tools/prof_kernels.py all > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_full.ncu-rep; rm -f gpurun_out/${TAG}_full.ncu-rep   # the raw csv travels back, the report is too large
