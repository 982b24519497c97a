#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): collects the evidence files that get copied into profiles/.
# usage: tools/collect_profiles.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
kill $SMI
tail -c 400 $OUT/bench_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"proposal_kernel|field_kernel|xf_kernel|hoist|finish_kernel|minmax|pdf_kernel" -c 70 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"proposal_kernel|field_kernel|xf_kernel|hoist_tc|pdf_kernel" -s 5 -c 5 -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
NJF_LIB=$PWD/neural-jacobian-field_b200/lib/libnjf_b200_prof.so timeout 300 python tools/phase_profile.py > $OUT/phase_$TAG.json 2>&1
tail -c 300 $OUT/phase_$TAG.json
