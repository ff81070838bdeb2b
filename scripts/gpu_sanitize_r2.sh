#!/bin/bash
# round 2, final build: compute-sanitizer memcheck over the spec-vector and unit tests of every path, racecheck over the kernels that use
# shared memory / shuffles (incl. the new item reduce and the column path), initcheck over the verifiers
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 9 python -m pytest tests/test_gpu_verify.py tests/test_gpu_commit.py tests/test_gpu_cells.py tests/test_gpu_proofs_recovery.py tests/test_setup_json.py tests/test_gpu_primitives.py tests/test_gpu_multidev.py -m gpu -x -q > gpurun_out/r02_san_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_san_memcheck.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 --error-exitcode 9 python -m pytest "tests/test_gpu_verify.py::test_synthetic_cell_batches" "tests/test_gpu_verify.py::test_synthetic_blob_batch_accept_and_reject" "tests/test_gpu_verify.py::test_large_verdict_uses_column_grouping" "tests/test_gpu_verify.py::test_optimistic_combined_pass_agrees_with_per_verdict_checks" tests/test_gpu_cells.py -m gpu -x -q > gpurun_out/r02_san_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_san_racecheck.log | tail -3
timeout 1200 compute-sanitizer --tool initcheck --print-limit 10 --error-exitcode 9 python -m pytest tests/test_gpu_verify.py -m gpu -x -q > gpurun_out/r02_san_initcheck.log 2>&1
echo "initcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_san_initcheck.log | tail -3
