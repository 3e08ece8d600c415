mkdir -p gpurun_out
python tools/step_breakdown.py --fast-init --detail "M=16384 N=1280 K=1280" > gpurun_out/r2_breakdown_q.txt 2>&1; grep -A60 "launches matching" gpurun_out/r2_breakdown_q.txt | head -60
SEER_RESIDUAL_STREAM=fp32 python tools/step_breakdown.py --fast-init --detail "M=16384 N=1280 K=1280" > gpurun_out/r2_breakdown_q32.txt 2>&1; grep -A60 "launches matching" gpurun_out/r2_breakdown_q32.txt | head -50
