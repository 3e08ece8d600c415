mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or conv" > gpurun_out/r2_t33_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t33_pytest.log
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_t.txt 2>&1; head -24 gpurun_out/r2_breakdown_t.txt
