mkdir -p gpurun_out
python tools/step_breakdown.py --fast-init --detail "M=" > gpurun_out/r2_breakdown_r.txt 2>&1; grep "spec=176" gpurun_out/r2_breakdown_r.txt | sort -k3,5 | uniq -c -f 2 | head -5; grep "spec=176" gpurun_out/r2_breakdown_r.txt | awk '{print $3,$4,$5,$6,$7, $12,$13,$14,$15,$16}' | sort | uniq -c
SEER_GEMM_HEAVY16=0 python tools/step_breakdown.py --fast-init --detail "M=" > gpurun_out/r2_breakdown_r0.txt 2>&1; grep "spec=176" gpurun_out/r2_breakdown_r0.txt | awk '{print $3,$4,$5,$6,$7, $12,$13,$14,$15,$16}' | sort | uniq -c
head -3 gpurun_out/r2_breakdown_r.txt; head -3 gpurun_out/r2_breakdown_r0.txt
