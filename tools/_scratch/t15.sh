mkdir -p gpurun_out
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_gn.txt 2>&1; head -1 gpurun_out/r2_breakdown_gn.txt; grep groupnorm gpurun_out/r2_breakdown_gn.txt
