mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -x -q -s -k "residual_stream or bf16_concat or conv_in_bf16 or test_step_vs_oracle or test_batch_independence or cfg_shared or bridge_step or ddim_loop or test_groupnorm or test_conv_in_out or golden" > gpurun_out/r2_t30_pytest.log 2>&1; echo "pytest rc=$?"; grep -i "rel-L2\|passed\|failed\|error" gpurun_out/r2_t30_pytest.log | tail -25
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_p.json').read().strip().splitlines()[-1]); print('bf16 stream', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
SEER_RESIDUAL_STREAM=fp32 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p_fp32stream.json 2> gpurun_out/r2_bench_p2.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_p_fp32stream.json').read().strip().splitlines()[-1]); print('fp32 stream', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
python tools/step_breakdown.py > gpurun_out/r2_breakdown_p.txt 2>&1; head -30 gpurun_out/r2_breakdown_p.txt
