mkdir -p gpurun_out
for ne in 8 4; do
  echo "== SEER_GEMM_NEPI=$ne"
  SEER_GEMM_NEPI=$ne python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_nepi$ne.txt 2>&1
  head -1 gpurun_out/r2_breakdown_nepi$ne.txt
  grep "gemm M=262144 N=2560 K=320\|gemm M=262144 N=960\|gemm M=262144 N=320 K=1280\|gemm M=65536 N=1920\|gemm M=65536 N=5120\|gemm M=65536 N=640 K=2560\|gemm M=16384 N=3840" gpurun_out/r2_breakdown_nepi$ne.txt
done
