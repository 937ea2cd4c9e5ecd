#!/bin/bash
# compute-sanitizer over the kernels.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun -- 'tools/sanitize.sh memcheck'      gpurun -- 'tools/sanitize.sh racecheck'
# Logs go to gpurun_out/sanitize_<tool>.log; copy the summary lines into profiles/.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
python tests/perf/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 3 --print-limit 20 python tests/perf/sanitize_target.py > gpurun_out/sanitize_${tool}.log 2>&1
rc=$?
echo "compute-sanitizer --tool $tool: exit $rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target|Error|hazard" gpurun_out/sanitize_${tool}.log | head -20
exit 0
