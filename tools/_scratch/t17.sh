mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q -k "rope or test_step_vs_oracle or golden" > gpurun_out/r2_t37_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t37_pytest.log
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_rope.txt 2>&1; head -1 gpurun_out/r2_breakdown_rope.txt; grep "rope\|groupnorm M=262144 C=320" gpurun_out/r2_breakdown_rope.txt | head
