#!/bin/bash
# Runs on the GPU box (under gpurun): the bench line, the ncu launch list of the same command and one full capture of
# the dominant kernel.  Outputs land in gpurun_out/ and are summarised into profiles/ by scripts/summarise_profiles.py.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 40 --warmup 8 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"himm_|vfh_update|grid_fill|grid_to_occ" -c 80 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:himm_tile -s 10 -c 1 -o gpurun_out/prof_tile_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vfh_update -s 10 -c 1 -o gpurun_out/prof_vfh_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:himm_prep -s 10 -c 1 -o gpurun_out/prof_prep_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
ls -la gpurun_out | tail -8
