timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -s -k shifted 2>&1 | grep -A3 "^shift "
