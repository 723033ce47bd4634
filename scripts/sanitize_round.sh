#!/bin/bash
# compute-sanitizer evidence for the final build (run under gpurun): memcheck, racecheck and synccheck over
# scripts/sanitize.py (float layers, byte-coded layers, scan form), plus racecheck with the multi-warp heavy-tile
# kernel.  Summaries land in gpurun_out/sanitizer_<tool>[_mw].log; copy them to profiles/.
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitizer_${tool}_${TAG}.full 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_${tool}_${TAG}.full
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize .*OK|exit code|Error|error|hazard" gpurun_out/sanitizer_${tool}_${TAG}.full | head -40 > gpurun_out/sanitizer_${tool}_${TAG}.log
done
B200NAV_MW_HEAVY=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitizer_racecheck_mw_${TAG}.full 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_racecheck_mw_${TAG}.full
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize .*OK|exit code|Error|error|hazard" gpurun_out/sanitizer_racecheck_mw_${TAG}.full | head -40 > gpurun_out/sanitizer_racecheck_mw_${TAG}.log
tail -n 4 gpurun_out/sanitizer_*_${TAG}.log
