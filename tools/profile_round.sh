#!/bin/bash
# Round evidence (run on the GPU box through gpurun): bench line, ncu launch list of the timed steps, and
# `--set full` captures of the dominant kernels.  Outputs land in gpurun_out/ (keep them under 64 MiB in total).
set -u
tag=${1:-r02_final}
what=${2:-all}
mkdir -p gpurun_out
if [ "$what" = all ] || [ "$what" = bench ]; then
  python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  tail -c 3000 gpurun_out/${tag}_bench_n1.json
fi
# launch list of exactly two resident steps (CUDA-graph kernel nodes are profiled one by one).  The conv autotuner times
# its candidates with CUDA events, which a profiler distorts: the ncu runs use the kernels' own default tilings
export STP3_CONV_AUTOTUNE=0
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --profiler-range > gpurun_out/${tag}_ncu_bench.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
# full-section captures: the conv kernels of the temporal model (18) + first decoder convs, and the lift-splat kernels
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_igemm|aspp_fused|block_fused" -c 16 \
    -o gpurun_out/${tag}_conv_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline --profiler-range \
    > gpurun_out/${tag}_ncu_conv.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"lift_splat|bev_finalize" -c 2 \
    -o gpurun_out/${tag}_liftsplat_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline --profiler-range \
    > gpurun_out/${tag}_ncu_ls.log 2>&1
ls -la gpurun_out/${tag}_*
du -sh gpurun_out
