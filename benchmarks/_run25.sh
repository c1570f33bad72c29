STEPS=100 bash benchmarks/ab.sh - "RAGARC_TC_RHO=1.0" "RAGARC_TC_RHO=1.04" "RAGARC_TC_PUB_WAIT=15000" 2>&1 | cut -c1-118 | tee gpurun_out/r02_t25_ab.log
