mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_semantic_gpu.py -x -q 2>&1 | tail -4
for st in 0 1; do
SNAPB200_ROOT_STAGED=$st timeout 300 ncu -k regex:gemm_tc_kernel --metrics gpu__time_duration.sum --clock-control none -c 4 --csv --log-file gpurun_out/r3u_root_$st.csv python tools/gemm_one.py root > /dev/null 2>&1; echo staged=$st; grep "gemm_tc_kernel" gpurun_out/r3u_root_$st.csv | awk -F'","' '{gsub(/"/,"",$NF); print $NF}' | tail -2
done
timeout 600 python bench.py > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err; echo bench rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/r3u_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline_encoder']['phases_ms_per_step'])"
