mkdir -p gpurun_out
timeout 600 python bench.py --config cfg4train --steps 10 --warmup 3 > gpurun_out/r3j_cfg4train.json 2> gpurun_out/r3j_cfg4train.err; echo rc=$?; tail -3 gpurun_out/r3j_cfg4train.err; python -c "
import json; d=json.loads(open('gpurun_out/r3j_cfg4train.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'], d['gpu_launches'], d['applied_steps'], d['skipped_steps'])"
