mkdir -p gpurun_out
for rep in 1 2; do
for lib in libseer_b200.so libseer_b200_nofast.so; do
  echo "== $lib (rep $rep)"
  SEER_B200_LIB=$PWD/seervideoldm_b200/$lib python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_u_${lib}_$rep.txt 2>&1
  head -1 gpurun_out/r2_breakdown_u_${lib}_$rep.txt
  grep "gemm M=262144 N=2560 K=320\|gemm M=262144 N=320 K=320 \|gemm M=65536 N=640 K=640 \|gemm M=262144 N=960\|gemm M=16384 N=1280 K=1280\|gemm M=262144 N=320 K=1280\|conv3x3 M=262144 N=320 K=2880\|gemm M=65536 N=1920" gpurun_out/r2_breakdown_u_${lib}_$rep.txt
done
done
