mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attention_tc_persist -c 3 -f -o gpurun_out/r2_attn40_full python tools/attn_bench.py --iters 1 > gpurun_out/r2_attn40_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_attn40_full.ncu-rep --page raw --csv > gpurun_out/r2_attn40_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_attn40_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_attn40_sass.csv 2>/dev/null
ls -la gpurun_out/; 
