mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --verify > gpurun_out/r2_bench_final_2gpu.json 2> gpurun_out/r2_bench_final_2gpu.err; echo "bench2 rc=$?"; tail -1 gpurun_out/r2_bench_final_2gpu.json | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d.get('verify'))"
python -m pytest tests/test_cfg_p2p_gpu.py -x -q > gpurun_out/r2_final_p2p_pytest.log 2>&1; echo "p2p pytest rc=$?"; tail -2 gpurun_out/r2_final_p2p_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/cfg_branch_split_check.py > gpurun_out/r2_cfg_branch_split_2gpu.txt 2>&1; echo "split rc=$?"; grep -v "^\s*$\|OMP\|\*\*\*" gpurun_out/r2_cfg_branch_split_2gpu.txt | tail -14
python bench.py --config stress --no-cpu-baseline > gpurun_out/r2_bench_final_stress.json 2> gpurun_out/r2_bench_final_stress.err; echo "stress rc=$?"; cut -c1-220 gpurun_out/r2_bench_final_stress.json
