mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py -x -q -s -k "conv_in or test_step_vs_oracle or golden or cfg_shared or batch_independence or bridge_step or ddim_loop" > gpurun_out/r2_t35_pytest.log 2>&1; echo "pytest rc=$?"; grep -i "rel-L2\|passed\|failed\|error" gpurun_out/r2_t35_pytest.log | tail -14
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_w.txt 2>&1; head -1 gpurun_out/r2_breakdown_w.txt; grep "small_linear\|conv_in\|N=320 K=64" gpurun_out/r2_breakdown_w.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_r.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
