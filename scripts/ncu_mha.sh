cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for spec in "mha_fwd@mha_fwd_kernel@2" "mha_dq@mha_bwd_dq_kernel@2" "mha_dkv@mha_bwd_dkv_kernel@2"; do
  IFS=@ read -r tag rx skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s "$skip" -c 1 -f -o "gpurun_out/ncu_r02a_$tag" python scripts/mha_bench.py --reps 1 > "gpurun_out/ncu_r02a_$tag.log" 2>&1
  echo "ncu $tag exit $?"
done
