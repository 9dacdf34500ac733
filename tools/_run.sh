mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/r3r_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r3r_pytest.log
timeout 600 python bench.py > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; echo bench rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/r3r_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity']['teacher_forced']); print(d['roofline_encoder']['phases_ms_per_step'])"
