for S in 37 74 81 111 148 222 296; do
  echo "S=$S"; RAGARC_DENSE_S=$S timeout 200 python benchmarks/run_configs.py --only c5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'], round(d['kernel_ms'],3), round(d['hbm_gbs']), round(d['merge_ms'],3))
"
done
