#!/bin/bash
# compute-sanitizer memcheck over the block / frame parity tests that exercise K1 and K2 on small inputs.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/${1:-memcheck}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $O.sanitizer.log \
  python -m pytest tests/test_gpu_block.py -m gpu -x -q \
  -k "sizes_sweep or acceleration or output_too_small or error_kinds or random_garbage or dense_dependencies or errors_and_capacity or dictionary_decode or large_blocks" \
  > $O.pytest.log 2>&1
echo "exit $?" >> $O.pytest.log
tail -3 $O.pytest.log
grep -c "Invalid\|out of bounds\|misaligned" $O.sanitizer.log
tail -5 $O.sanitizer.log
