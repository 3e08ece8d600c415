mkdir -p gpurun_out
for n in 0 1 2 3; do
  lib=seervideoldm_b200/libseer_b200_poly$n.so; [ $n = 2 ] && lib=seervideoldm_b200/libseer_b200.so
  echo "== SEER_ATTN_POLY=$n ($lib)"; SEER_B200_LIB=$PWD/$lib python tools/attn_bench.py 2>&1 | grep "d=40"
done > gpurun_out/r2_attn_poly_ab.txt 2>&1
cat gpurun_out/r2_attn_poly_ab.txt
python tools/cfg_p2p_loopback.py > gpurun_out/r2_cfg_p2p_loopback.txt 2>&1; echo "loopback rc=$?"; cat gpurun_out/r2_cfg_p2p_loopback.txt | tail -5
python -m pytest tests/test_kernels_gpu.py tests/test_cfg_p2p_gpu.py tests/test_bench_shapes_gpu.py -x -q -k "attention or p2p or config5" > gpurun_out/r2_t25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t25_pytest.log
python tools/vae_bench.py > gpurun_out/r2_vae_bench.txt 2>&1; echo "vae rc=$?"; grep "^VAE" gpurun_out/r2_vae_bench.txt
