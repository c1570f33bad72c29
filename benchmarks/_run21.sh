timeout 400 python -m pytest tests/test_gpu_dense.py tests/test_gpu_variants.py tests/test_gpu_random_shapes.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_t21_tests.log
timeout 200 python benchmarks/cabi_latency.py 2>&1 | tail -1 | tee gpurun_out/r02_t21_cabi.txt
RAGARC_TC_PDL=0 timeout 200 python benchmarks/cabi_latency.py 2>&1 | tail -1 | tee -a gpurun_out/r02_t21_cabi.txt
STEPS=100 bash benchmarks/ab.sh - "RAGARC_TC_PDL=0" 2>&1 | cut -c1-110 | tee gpurun_out/r02_t21_ab.log
