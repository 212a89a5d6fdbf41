#!/bin/bash
# One GPU call: every training-step test in its own process (a faulting kernel poisons only its own CUDA context), compute-sanitizer on the first
# failure, then short `bench.py --workload train` runs of the variants.  Logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/train_tests.log
export CUDA_LAUNCH_BLOCKING=1
fail=0
for t in "test_attention_backward_matches_autograd[mma]" "test_train_step_gradients_match_autograd[point]" "test_train_step_gradients_match_autograd[point_latent]" \
         test_train_step_with_dropout_and_padding "test_train_step_mid_size[default]" "test_train_step_mid_size[stats_pass]" "test_train_step_mid_size[wmma]" \
         "test_train_step_mid_size[recompute]" "test_train_step_mid_size[recompute+stats_pass]" "test_train_step_mid_size[recompute+wmma]" \
         test_lmm_train_mode_backward_and_optimizer_step test_flat_trainer_steps_reduce_the_loss; do
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
for v in "" "--debug train_recompute=1"; do
  timeout 300 python bench.py --workload train --steps 2 --warmup 1 $v > "gpurun_out/bench_train${v// /_}.json" 2> gpurun_out/bench_train.err
  echo "bench [$v] rc=$?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['loss_history'])" "gpurun_out/bench_train${v// /_}.json"; tail -3 gpurun_out/bench_train.err
done
