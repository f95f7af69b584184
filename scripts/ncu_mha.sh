#!/bin/bash
# ncu --set full captures of the attention kernels on the CLIP ViT-B/16 shapes (scripts/mha_bench.py); usage: ncu_mha.sh tag@regex@skip ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for spec in "$@"; do
  IFS=@ read -r tag rx skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s "${skip:-2}" -c 1 -f -o "gpurun_out/ncu_$tag" python scripts/mha_bench.py --reps 1 > "gpurun_out/ncu_$tag.log" 2>&1
  echo "ncu $tag exit $?"
done
