mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py tests/test_vae_gpu.py -x -q -s -k "groupnorm or statistics or test_step_vs_oracle or golden or bridge_step or ddim_loop or vae or fp32" > gpurun_out/r2_t36_pytest.log 2>&1; echo "pytest rc=$?"; grep -i "rel-L2\|passed\|failed\|error" gpurun_out/r2_t36_pytest.log | tail -22
python tools/step_breakdown.py --fast-init > gpurun_out/r2_breakdown_gn2.txt 2>&1; head -1 gpurun_out/r2_breakdown_gn2.txt; grep groupnorm gpurun_out/r2_breakdown_gn2.txt | head -6
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_s.json 2> gpurun_out/r2_bench_s.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_s.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
