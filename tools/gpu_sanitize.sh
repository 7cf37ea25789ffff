#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory pipelines) over a reduced test set
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SEL='tests/test_gpu_greedy.py::test_greedy_matches_reference_golden tests/test_gpu_greedy.py::test_d2_sampling_matches_reference_golden tests/test_gpu_greedy.py::test_nan_retry_branch_matches_reference_golden tests/test_gpu_greedy.py::test_min_dist_exact_on_integer_rows tests/test_gpu_scores.py::test_margin_and_confidence_match_reference_golden tests/test_gpu_scores.py::test_select_all_equal_scores tests/test_gpu_scores.py::test_badge_factors_match_oracle_and_reference'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest $SEL -x -q -m gpu > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_greedy.py::test_greedy_matches_reference_golden tests/test_gpu_scores.py::test_select_all_equal_scores -x -q -m gpu > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
