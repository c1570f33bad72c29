#!/bin/bash
# A/B of environment-selected kernel variants inside ONE box lease (box-to-box power variance is
# larger than most effects).  usage: ab.sh "VAR=val ..." "VAR=val ..." ...   ("-" = defaults)
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then e=""; else e="$v"; fi
  env $e timeout 600 python bench.py --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline --quick 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-28s q/s %8.0f step %.4f main %.4f merge %.4f seed %.4f e2e %.4f clk %s sched %s' % ('$v', d['value'], d['ms_per_step'], r['kernel_ms'], r['merge_kernel_ms'], r['setup_ms'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['config'].get('schedule')))"
done; done
