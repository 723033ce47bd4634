#!/bin/bash
# compute-sanitizer evidence for the final build (run under gpurun) over scripts/sanitize.py (float layers, byte-coded
# layers, scan form; fleets of 2-3 robots):
#   memcheck, synccheck          default kernels (small fleets: the multi-warp pipeline tile kernel)
#   racecheck                    one warp per tile (B200NAV_MW_HEAVY=0): must report 0 hazards
#   racecheck_mw                 the multi-warp pipeline kernel: its warps synchronise through release / acquire progress
#                                words in shared memory, which racecheck does not model - it reports those words and
#                                the cells they guard (expected; every report must be pipe_load / pipe_store / lds_u8 /
#                                sts_u8 inside himm_apply_list_pipe)
# Summaries land in gpurun_out/sanitizer_<tool>_<tag>.log; copy them to profiles/.
TAG=${1:-r2}
mkdir -p gpurun_out
run() { # name, tool, env
  env $3 timeout 900 compute-sanitizer --tool $2 --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitizer_$1_${TAG}.full 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$1_${TAG}.full
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize .*OK|exit code|Error|error" gpurun_out/sanitizer_$1_${TAG}.full | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -30 > gpurun_out/sanitizer_$1_${TAG}.log
}
run memcheck memcheck B200NAV_MW_HEAVY=1
run synccheck synccheck B200NAV_MW_HEAVY=1
run racecheck racecheck B200NAV_MW_HEAVY=0
run racecheck_mw racecheck B200NAV_MW_HEAVY=1
tail -n 8 gpurun_out/sanitizer_*_${TAG}.log
# the GPU test suite itself under memcheck (everything but the subprocess-spawning tests) and the HIMM + VFH+ tests
# under racecheck with one warp per tile
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests -m gpu -q -k "not bench_contract and not fleet" > gpurun_out/sanitizer_memcheck_suite_${TAG}.full 2>&1
B200NAV_MW_HEAVY=0 timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_himm_gpu.py tests/test_vfh_gpu.py -m gpu -q -k "not c2_sized and not c3_sized and not c4_full and not more_tiles" > gpurun_out/sanitizer_racecheck_suite_${TAG}.full 2>&1
for n in memcheck_suite racecheck_suite; do
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_${n}_${TAG}.full > gpurun_out/sanitizer_${n}_${TAG}.log
  cat gpurun_out/sanitizer_${n}_${TAG}.log
done
