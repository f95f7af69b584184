#!/bin/bash
# Runs on the GPU box under gpurun: parity tests (each group in its own process so a CUDA fault in
# one kernel cannot poison the others), smoke, a short bench and (PROFILE=1) the ncu launch list.
# Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 900"
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n ${TAILN:-4} gpurun_out/$name.log; }
if [ -z "$SKIP_TESTS" ]; then
run kernels   $PT tests/test_kernels_gpu.py
run step      $PT tests/test_step_gpu.py
run smoke     python __graft_entry__.py smoke
fi
TAILN=2 run bench     python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --batch ${BENCH_BATCH:-64}
if [ -n "$TRAFFIC" ]; then   # DRAM bytes of every launch of one step at the bench batch (application replay: one app run per pass)
  run ncu_traffic ncu --profile-from-start off --replay-mode application --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/traffic.csv python scripts/profile_step.py --batch ${BENCH_BATCH:-64}
fi
if [ -n "$PROFILE" ]; then
  run ncu_list ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py --batch ${PROFILE_BATCH:-4}
fi
