#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (memcheck everywhere, racecheck on the
# kernels that do shared-memory read-modify-write).  Slow: run on its own gpurun call.
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
timeout 900 $CS --tool memcheck python -m pytest -x -q -m gpu \
  tests/test_gpu_dense.py -k "tcgen05_path or ties or k_larger or ascending or rows_subset or c1_fp32 or owner_push or rung or spill" \
  > gpurun_out/sanitize_memcheck_dense.log 2>&1; echo "memcheck dense rc=$?"; tail -4 gpurun_out/sanitize_memcheck_dense.log
timeout 900 $CS --tool memcheck python -m pytest -x -q -m gpu \
  tests/test_gpu_sparse_fusion_pool.py tests/test_gpu_native_index.py tests/test_gpu_tails.py tests/test_gpu_comm.py \
  tests/test_gpu_parity_full.py tests/test_gpu_plugins.py \
  -k "not c2_shape and not doc_range and not large_random and not c3_every and not c4_shard and not c2_dense and not huggingface and not two_gpus and not per_gpu" \
  > gpurun_out/sanitize_memcheck_misc.log 2>&1; echo "memcheck misc rc=$?"; tail -4 gpurun_out/sanitize_memcheck_misc.log
timeout 900 $CS --tool racecheck python -m pytest -x -q -m gpu \
  tests/test_gpu_sparse_fusion_pool.py -k "bm25_small or fewer_matches or rrf or pool_all_masked" \
  > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
