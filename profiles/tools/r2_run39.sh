#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_memcheck.txt \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy.py -q -x -m gpu \
    -k "warp_per_env_kernel_equals and iris_softmax-5 or resident_minibatch or reset_pipeline or ring_only_step_is or ring_front_end and natural or dense_policy and 4321 and 0-" > gpurun_out/r2_memcheck_pytest.txt 2>&1
echo "exit $?"; tail -4 gpurun_out/r2_memcheck_pytest.txt; grep -c "Invalid\|Error" gpurun_out/r2_memcheck.txt; tail -5 gpurun_out/r2_memcheck.txt
