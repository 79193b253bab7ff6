#!/bin/bash
# compute-sanitizer over the toy-shape GPU parity tests of every kernel family that hand-rolls mbarrier / TMEM / cluster
# protocols (GEMM 1-CTA + CTA pair + dgrad/wgrad, attention single / pair / causal / bias / backward, conv, LN-modulate
# and the training row kernels).  Run on a GPU box:   bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]
# Writes gpurun_out/sanitize_<tool>.log and a one-line-per-tool summary gpurun_out/sanitize_summary.json; copy the
# summary to profiles/ to commit it.  The selection keeps each tool under a few minutes (sanitizers slow kernels 10-100x).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-"memcheck racecheck synccheck"}
NODES=$(python scripts/sanitize_select.py)
echo "{" > gpurun_out/sanitize_summary.json
first=1
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  timeout 2400 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
      python -m pytest $NODES -m gpu -q -p no:cacheprovider > $log 2>&1
  rc1=$?
  rc2=0
  errs=$(grep -c "========= .*\(Invalid\|Race\|hazard\|Barrier error\|Uninitialized\)" $log || true)
  summ=$(grep "ERROR SUMMARY" $log | tr '\n' ';')
  passed=$(grep -E "passed|failed" $log | tr '\n' ';')
  [ $first -eq 0 ] && echo "," >> gpurun_out/sanitize_summary.json
  first=0
  printf '"%s": {"rc": [%d, %d], "error_lines": %s, "summary": "%s", "pytest": "%s"}' "$tool" $rc1 $rc2 "${errs:-0}" "$summ" "$passed" >> gpurun_out/sanitize_summary.json
  echo "[$tool] rc=$rc1,$rc2 $summ $passed"
done
echo "}" >> gpurun_out/sanitize_summary.json
