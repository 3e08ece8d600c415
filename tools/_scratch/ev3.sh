mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
python bench.py --verify > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_final.json
python bench.py --impl reference > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_final_ref.json
python bench.py --config sthv2 --no-cpu-baseline > gpurun_out/r2_bench_final_sthv2.json 2> gpurun_out/r2_bench_final_sthv2.err; echo "sthv2 rc=$?"; cut -c1-200 gpurun_out/r2_bench_final_sthv2.json
python bench.py --config stress --no-cpu-baseline > gpurun_out/r2_bench_final_stress.json 2> gpurun_out/r2_bench_final_stress.err; echo "stress rc=$?"; cut -c1-200 gpurun_out/r2_bench_final_stress.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bridge_final.csv python tools/profile_step.py --fast-init > gpurun_out/r2_launches.log 2>&1; echo "launch list rc=$?"
bash tools/capture_gemm_full.sh; echo "capture rc=$?"
python tools/step_breakdown.py > gpurun_out/r2_breakdown_final.txt 2>&1; head -3 gpurun_out/r2_breakdown_final.txt
python tools/vae_bench.py > gpurun_out/r2_vae_bench.txt 2>&1; grep "^VAE" gpurun_out/r2_vae_bench.txt
du -sh gpurun_out
