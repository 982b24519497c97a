#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): collects the evidence files that get copied into profiles/.
# usage: tools/collect_profiles.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "${SKIP_BENCH:-}" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  tail -c 400 $OUT/bench_$TAG.json
fi
# launch list of two timed frames (cold-cache, serialised: compare SHARES with the CUDA-event breakdown, not absolutes)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"proposal_kernel|field_kernel|xf_kernel|hoist|finish_kernel|minmax|pdf_kernel" -c 200 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
# one frame under --set full: warm-up = 3 frames x 28 launches (NJF_BENCH_LAUNCHES_PER_FRAME printed by bench.py)
timeout 900 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --import-source on -k regex:"proposal_kernel|field_kernel|xf_kernel|hoist_tc|pdf_kernel|finish_kernel" -s ${SKIP:-78} -c ${COUNT:-26} -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
# the report itself is too large to travel back (gpurun_out/ is capped at 64 MiB): keep the raw metric page as CSV
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/prof_$TAG.ncu-rep
NJF_LIB=$PWD/neural-jacobian-field_b200/lib/libnjf_b200_prof.so timeout 300 python tools/phase_profile.py > $OUT/phase_$TAG.json 2> $OUT/phase_$TAG.err
tail -c 300 $OUT/phase_$TAG.json
