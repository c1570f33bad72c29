# launch list of the default bench command (eager, so that every kernel is its own launch), per-launch times
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline --no-graph > /dev/null 2>&1
# full capture of one step's kernels: cluster launch, left-over launch, merge
ncu --set full --import-source on --clock-control none -k regex:"dense_tc_kernel|merge_lists_kernel|normalize_cast" -s 9 -c 4 -o gpurun_out/r02_final_step -f python benchmarks/tc_stats.py > gpurun_out/r02_final_ncu.log 2>&1
ls -la gpurun_out/r02_final_step.ncu-rep
