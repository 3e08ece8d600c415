mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bridge.csv python tools/profile_step.py --fast-init > gpurun_out/r2_launches.log 2>&1; echo "launch list rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sthv2_b1.csv python tools/profile_step.py --fast-init --clips 1 --frames 12 >> gpurun_out/r2_launches.log 2>&1; echo "launch list b1 rc=$?"
bash tools/capture_gemm_full.sh; echo "capture rc=$?"
bash tools/sanitize.sh memcheck racecheck synccheck > gpurun_out/r2_sanitize_final.out 2>&1; echo "sanitize rc=$?"; tail -5 gpurun_out/r2_sanitize_final.out
SEER_EMULATE_BF16_STREAM=1 python -m pytest tests/test_bench_shapes_gpu.py -x -q -s -k "bridge_step or ddim_loop_bf16" > gpurun_out/r2_bf16_stream_emulation.log 2>&1; echo "emu rc=$?"; grep -i "rel\|passed\|failed" gpurun_out/r2_bf16_stream_emulation.log | tail
python -m pytest tests/test_bench_shapes_gpu.py -x -q -s -k "bridge_step or ddim_loop_bf16" > gpurun_out/r2_bf16_stream_baseline.log 2>&1; grep -i "rel\|passed\|failed" gpurun_out/r2_bf16_stream_baseline.log | tail
du -sh gpurun_out
