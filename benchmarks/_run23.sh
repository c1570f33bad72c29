CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_plugins.py -x -q -k "overlapped or pipeline" 2>&1 | tail -3 | tee gpurun_out/r02_t23_tests.log
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/r02_t23_tests.log
for extra in "--overlap" ""; do CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 100 --quick --no-cpu-baseline $extra 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=1 $extra q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f ov %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], d['e2e']['ms_per_step'], d['config']['overlapped_select']))"; done | tee gpurun_out/r02_t23_ab.log
for extra in "--overlap" ""; do timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --quick $extra 2>gpurun_out/r02_t23_err.log | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=2 $extra q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], d['e2e']['ms_per_step'], d['config']['parallelism']))"; done | tee -a gpurun_out/r02_t23_ab.log
tail -5 gpurun_out/r02_t23_err.log | cut -c1-300
