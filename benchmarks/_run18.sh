run() { timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 --steps 100 --warmup 5 "${@:2}" 2>gpurun_out/r02_scale_err_$1.log | grep '^{' ; }
run 8 > gpurun_out/r02_scale_n8.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_scale_n8.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("N=8 full: q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f verified %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["merge_kernel_ms"], d["e2e"]["ms_per_step"], d["config"]["verified"]))
ex=d["extra"]; print("c4:", {k: ex["c4"].get(k) for k in ("value","ms_per_step","step_roofline_frac","verified")} if "error" not in ex["c4"] else ex["c4"])
print("c5:", [(b["batch"], round(b["ms_per_step"],3), round(b["host_in_host_out_ms"],3), round(b["roofline"]["frac"],3)) for b in ex["c5"]["batches"]] if "error" not in ex["c5"] else ex["c5"])
PY
RAGARC_OWNER_SIGNAL=1 run 8 --quick | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=8 signal: q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], d['e2e']['ms_per_step']))"
RAGARC_EXCHANGE=nccl run 8 --quick | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=8 nccl allgather: q/s %.0f step %.4f par %s' % (d['value'], d['ms_per_step'], d['config']['parallelism']))"
CUDA_VISIBLE_DEVICES=0,1,2,3 run 4 > gpurun_out/r02_scale_n4.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_scale_n4.json').read().strip().splitlines()[-1]); r=d['roofline']; print('N=4 full: q/s %.0f step %.4f main %.4f merge %.4f e2e %.4f c4 %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], d['e2e']['ms_per_step'], d['extra']['c4'].get('value')))"
