mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 --verify > gpurun_out/r2_bench_final_8gpu.json 2> gpurun_out/r2_bench_final_8gpu.err; echo "bench8 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final_8gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d.get('verify'), d['clocks'])"
tail -3 gpurun_out/r2_bench_final_8gpu.err
