#!/bin/bash
# Runs on the GPU box under gpurun: one `ncu --set full` capture per kernel regex (first matching launch after
# SKIP earlier ones) of one profiled train step.  Reports -> gpurun_out/ncu_<tag>.ncu-rep (read back on the
# CPU box with scripts/ncu_summary.py).  usage: ncu_capture.sh tag@regex[@skip]  (regex over the demangled name incl. template args) ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=@ read -r tag rx skip <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      --kernel-name-base demangled -k "regex:$rx" -s "${skip:-0}" -c 1 -f -o "gpurun_out/ncu_$tag" \
      python scripts/profile_step.py --model "${PROFILE_MODEL:-swin_t}" --batch "${PROFILE_BATCH:-4}" > "gpurun_out/ncu_$tag.log" 2>&1
  echo "ncu $tag exit $?"
done
