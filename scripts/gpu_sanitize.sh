#!/bin/bash
# compute-sanitizer over the spec-vector and unit tests of every path (memcheck), racecheck over the kernels that use shared memory / shuffles
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 9 python -m pytest tests/test_gpu_verify.py tests/test_gpu_commit.py tests/test_gpu_cells.py tests/test_gpu_proofs_recovery.py tests/test_setup_json.py tests/test_gpu_primitives.py -m gpu -x -q > gpurun_out/san_memcheck_full.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_full.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 --error-exitcode 9 python -m pytest "tests/test_gpu_verify.py::test_synthetic_cell_batches" "tests/test_gpu_verify.py::test_synthetic_blob_batch_accept_and_reject" tests/test_gpu_cells.py -m gpu -x -q > gpurun_out/san_racecheck_full.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck_full.log | tail -3
