mkdir -p gpurun_out
for v in new old new old; do
  if [ $v = old ]; then export SEER_GEMM_LIGHT16_OFF=1; else unset SEER_GEMM_LIGHT16_OFF; fi
  echo "== $v"
  python tools/step_breakdown.py --fast-init --detail "K=1280\|K=2560\|K=5120" > gpurun_out/r2_breakdown_l16_$v.txt 2>&1
  head -1 gpurun_out/r2_breakdown_l16_$v.txt
  grep "gemm M=262144 N=320 K=1280 \|gemm M=65536 N=640 K=2560 \|gemm M=16384 N=1280 K=5120 \|gemm M=4096 N=1280 K=5120 " gpurun_out/r2_breakdown_l16_$v.txt | head -4
done
