timeout 400 python -m pytest tests/test_gpu_dense.py tests/test_gpu_random_shapes.py tests/test_gpu_variants.py -x -q 2>&1 | tail -4 | tee gpurun_out/r02_t15_tests.log
STEPS=100 bash benchmarks/ab.sh - 2>&1 | cut -c1-120 | tee gpurun_out/r02_t15_ab.log
