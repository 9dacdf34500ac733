mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err; echo bench rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/r3v_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline_encoder']['phases_ms_per_step'])"
