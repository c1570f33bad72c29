timeout 600 python benchmarks/bm25_large.py 2>&1 | tail -2 | tee gpurun_out/r02_bm25_large.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:bm25 -c 6 --csv --log-file gpurun_out/r02_bm25_large_ncu.csv python benchmarks/bm25_large.py 1000000 > /dev/null 2>&1
tail -8 gpurun_out/r02_bm25_large_ncu.csv | cut -c1-400
