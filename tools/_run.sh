mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config cfg4train --train-encoder --steps 6 --warmup 3 > gpurun_out/r3x_cfg4train_full_8gpu.json 2> gpurun_out/r3x_cfg4train_full_8gpu.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r3x_bench_8gpu.json 2> gpurun_out/r3x_bench_8gpu.err; echo rc=$?
python -c "
import json
for f in ('r3x_cfg4train_full_8gpu','r3x_bench_8gpu'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce'))
"
