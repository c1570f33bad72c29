timeout 400 python -m pytest tests/test_gpu_dense.py tests/test_gpu_multi.py -x -q 2>&1 | tail -4
for e in "RAGARC_OWNER_SIGNAL=0" "RAGARC_OWNER_SIGNAL=1"; do env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --quick 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$e', 'q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f sync %.4f par %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['sync_ms_per_step'], d['config']['parallelism']))"; done
CUDA_VISIBLE_DEVICES=0 ncu --set full --import-source on --clock-control none -k regex:merge_lists_kernel -s 3 -c 1 -o gpurun_out/r02_merge_lists -f python benchmarks/tc_stats.py > gpurun_out/r02_ncu_merge.log 2>&1
CUDA_VISIBLE_DEVICES=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_quick.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline --no-graph > /dev/null 2>&1
