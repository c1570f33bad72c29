timeout 300 python benchmarks/cabi_latency.py 2>&1 | tail -1 | tee gpurun_out/r02_cabi_latency.txt
bash benchmarks/sanitize.sh 2>&1 | tee gpurun_out/r02_sanitize_summary.txt
