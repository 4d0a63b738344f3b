#!/bin/bash
# compute-sanitizer over the unit-size workload.  Run on the GPU box:
#   bash scripts/run_sanitizer.sh [tools...]      (default: memcheck racecheck synccheck initcheck)
# Logs -> gpurun_out/sanitizer_<tool>.log ; one-line verdicts -> gpurun_out/sanitizer_summary.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck synccheck initcheck}"
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  small=0
  [ "$tool" = racecheck ] && small=1
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  start=$(date +%s)
  GPB_SANITIZER_SMALL=$small timeout 1500 compute-sanitizer --tool "$tool" $extra --print-limit 20 --error-exitcode 7 \
      python scripts/sanitizer_workload.py > "gpurun_out/sanitizer_${tool}.log" 2>&1
  rc=$?
  end=$(date +%s)
  echo "$tool rc=$rc seconds=$((end - start)) :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|workload OK' "gpurun_out/sanitizer_${tool}.log" | tr '\n' ' ')" \
      | tee -a gpurun_out/sanitizer_summary.txt
done
