#!/bin/bash
# One GPU call: every training-step test in its own process (a faulting kernel poisons only its own CUDA context), compute-sanitizer on the first
# failure, then short `bench.py --workload train` runs of the variants.  Logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/train_tests.log
export CUDA_LAUNCH_BLOCKING=1
fail=0
for t in test_train_step_trains_the_point_encoder test_lmm_train_mode_backward_and_optimizer_step "test_train_step_gradients_match_autograd[point]"; do
  echo "=== $t" >> gpurun_out/train_tests.log
  timeout 240 python -m pytest "tests/test_gpu_train.py::$t" -q -x 2>&1 | tail -40 >> gpurun_out/train_tests.log
  rc=${PIPESTATUS[0]}
  echo "rc=$rc" >> gpurun_out/train_tests.log
  if [ $rc -ne 0 ]; then
    fail=$((fail+1))
    if [ $fail -eq 1 ]; then
      timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_train.py::$t" -q -x 2>&1 | grep -v "^$" | head -120 > gpurun_out/train_sanitizer.log
    fi
  fi
done
grep -E "^===|^rc=|passed|failed|Error|error|assert" gpurun_out/train_tests.log | tail -60
unset CUDA_LAUNCH_BLOCKING
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "teacher_forced or attention_seam or encode_cond or padded" 2>&1 | tail -3
