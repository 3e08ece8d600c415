mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q -k "small_linear or test_step_vs_oracle or golden or fp32" > gpurun_out/r2_t34_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t34_pytest.log
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_v.txt 2>&1; head -1 gpurun_out/r2_breakdown_v.txt; grep "small_linear\|conv_in\|timestep" gpurun_out/r2_breakdown_v.txt
python tools/step_breakdown.py --fast-init --clips 1 --frames 12 > gpurun_out/r2_breakdown_v_b1.txt 2>&1; head -1 gpurun_out/r2_breakdown_v_b1.txt; grep "small_linear\|conv_in\|timestep" gpurun_out/r2_breakdown_v_b1.txt
