timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02_t16_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_t16_smoke.log
