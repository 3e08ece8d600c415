#!/bin/bash
# `ncu --set full` capture of the dominant kernel (gemm_tc_kernel: tcgen05 GEMM + implicit-GEMM conv) on the round's kernels,
# the source of `roofline.traffic` in bench.py.  Run on the GPU box (gpurun), then summarise locally:
#
#   gpurun -- tools/capture_gemm_full.sh                 # -> gpurun_out/r2_gemm_full_raw.csv + r2_gemm_launch_list.json
#   python tools/summarize_gemm_full.py                  # -> profiles/r2_gemm_full.summary.{json,txt}
#
# Captured: the first 66 gemm_tc launches of ONE eager UNet evaluation at the bench shape (UNet batch 16, 16 frames, 32x32
# latents) = conv/linear launches of the level-0 and level-1 down blocks (K = 320 / 640 linears, N = 320 / 640 convs, GEGLU,
# stride-2 convs).  ncu replays each launch ~40 times: its times are cold-cache / serialised — only the DRAM byte counts and
# the pipe utilisations are taken from it, the launch durations behind `roofline.achieved` are measured live by bench.py.
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_step.py --fast-init --list-gemm gpurun_out/r2_gemm_launch_list.json > gpurun_out/r2_gemm_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -c ${NCU_COUNT:-66} \
    -f -o gpurun_out/r2_gemm_full python tools/profile_step.py --fast-init >> gpurun_out/r2_gemm_full.log 2>&1
tail -3 gpurun_out/r2_gemm_full.log
# gpurun brings back at most 64 MiB: export the metric pages here and drop the report (~200 MB with sources)
ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_full_raw.csv
if [ -z "$KEEP_NCU_REP" ]; then rm -f gpurun_out/r2_gemm_full.ncu-rep; fi
