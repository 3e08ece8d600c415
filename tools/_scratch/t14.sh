mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attention_tc_persist --launch-skip 5 -c 1 -f -o gpurun_out/r2_attn40_cross python tools/attn_bench.py --iters 1 > gpurun_out/r2_attn40_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_attn40_cross.ncu-rep --page raw --csv > gpurun_out/r2_attn40_cross_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_attn40_cross.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_attn40_cross_sass.csv 2>/dev/null
ncu -i gpurun_out/r2_attn40_cross.ncu-rep --page source --csv > gpurun_out/r2_attn40_cross_src.csv 2>/dev/null
rm -f gpurun_out/r2_attn40_cross.ncu-rep
ls -la gpurun_out/
