#!/bin/bash
# Runs on the GPU box under gpurun: parity tests (each group in its own process so a CUDA fault in
# one kernel cannot poison the others), smoke, and a short bench.  Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 900"
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 4 gpurun_out/$name.log; }
run k_other   $PT tests/test_kernels_gpu.py -k "not gemm"
run k_gemm32  $PT tests/test_kernels_gpu.py -k "gemm and float32"
run k_gemm16  $PT tests/test_kernels_gpu.py -k "gemm and not float32"
run step32    $PT tests/test_step_gpu.py -k "float32 or replay"
run step16    $PT tests/test_step_gpu.py -k "not float32 and not replay"
run smoke     python __graft_entry__.py smoke
run bench     python bench.py --steps 3 --warmup 3 --batch ${BENCH_BATCH:-4}
