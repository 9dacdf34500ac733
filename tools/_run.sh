mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r3h_pytest.log
timeout 300 python tools/gemm_sweep.py 32 > gpurun_out/r3h_sweep.log 2>&1; echo sweep rc=$?; cat gpurun_out/r3h_sweep.log
timeout 600 python bench.py > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; echo bench rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/r3h_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline_encoder']['phases_ms_per_step'])"
