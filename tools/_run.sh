mkdir -p gpurun_out
( timeout 1500 python tools/ncu_traffic.py --tag r02 > gpurun_out/r3w_traffic.log 2>&1; echo traffic rc=$? )
cp gpurun_out/traffic.json profiles/traffic.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3w_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r3w_pytest.log
timeout 600 python bench.py > gpurun_out/r3w_bench.json 2> gpurun_out/r3w_bench.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3w_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 600 python bench.py --config cfg4 > gpurun_out/r3w_cfg4.json 2> gpurun_out/r3w_cfg4.err; echo cfg4 rc=$?
timeout 600 python bench.py --config cfg5 > gpurun_out/r3w_cfg5.json 2> gpurun_out/r3w_cfg5.err; echo cfg5 rc=$?
python -c "
import json
for f in ('r3w_bench','r3w_cfg4','r3w_cfg5'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('roofline',{}).get('traffic'))
"
