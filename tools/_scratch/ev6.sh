mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
python bench.py --verify > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['verify']['bit_identical'], d['clocks'])"
python bench.py --config sthv2 --no-cpu-baseline > gpurun_out/r2_bench_final_sthv2.json 2> gpurun_out/r2_bench_final_sthv2.err; echo "sthv2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final_sthv2.json').read().strip().splitlines()[-1]); print('sthv2', d['value'], d['ms_per_step'])"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bridge_final.csv python tools/profile_step.py --fast-init > gpurun_out/r2_launches.log 2>&1; echo "launch list rc=$?"
python tools/step_breakdown.py > gpurun_out/r2_breakdown_final.txt 2>&1; head -3 gpurun_out/r2_breakdown_final.txt
