mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_final_smoke.log
python bench.py --verify > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_o.json
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bridge.csv python tools/profile_step.py --fast-init > gpurun_out/r2_launches.log 2>&1; echo "launch list rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sthv2_b1.csv python tools/profile_step.py --fast-init --clips 1 --frames 12 >> gpurun_out/r2_launches.log 2>&1; echo "launch list b1 rc=$?"
bash tools/capture_gemm_full.sh; echo "capture rc=$?"
ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2_gemm_full*
bash tools/sanitize.sh memcheck racecheck synccheck > gpurun_out/r2_sanitize_final.out 2>&1; echo "sanitize rc=$?"; tail -5 gpurun_out/r2_sanitize_final.out
python bench.py --config sweep64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_sweep64.json 2> gpurun_out/r2_bench_sweep64.err; echo "sweep64 rc=$?"; cat gpurun_out/r2_bench_sweep64.json
