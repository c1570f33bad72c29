timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py tests/test_gpu_native_index.py tests/test_gpu_plugins.py tests/test_gpu_c_abi.py tests/test_gpu_sparse_fusion_pool.py -q 2>&1 | tail -8 | tee gpurun_out/r02_t17_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python benchmarks/profile_plugin.py 2>&1 | head -3 | tee gpurun_out/r02_t17_plugin.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 100 --quick --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e']))" | tee gpurun_out/r02_t17_e2e.log
