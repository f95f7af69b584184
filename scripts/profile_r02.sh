cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
# 1. final bench lines (with baselines) for both workloads
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_swin_t_b64.json 2> gpurun_out/r02c_bench_swin_t_b64.err
python bench.py --model clip_b16 --batch 64 --steps 10 --warmup 3 > gpurun_out/r02c_bench_clip_b16_b64.json 2> gpurun_out/r02c_bench_clip_b16_b64.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02c_bench_reference_arm_swin_t.json 2>/dev/null
python bench.py --impl reference --model clip_b16 --steps 4 --warmup 1 > gpurun_out/r02c_bench_reference_arm_clip_b16.json 2>/dev/null
# 2. ncu launch lists (durations + DRAM bytes of every launch of one step)
for m in "clip_b16 16" "swin_t 16"; do set -- $m
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02c_traffic_$1_b$2.csv python scripts/profile_step.py --model $1 --batch $2 > gpurun_out/r02c_traffic_$1.log 2>&1
  echo "launch list $1 exit $?"
done
# 3. --set full captures of the new kernels
bash scripts/ncu_mha.sh r02c_mha_tc_fwd@mha_tc_fwd_kernel@2 r02c_mha_tc_bwd@mha_tc_bwd_kernel@2 r02c_tt_mha_bwd@tt_bwd_kernel@2 r02c_tt_mha_fwd@tt_fwd_kernel@2
PROFILE_BATCH=16 bash scripts/ncu_capture.sh "r02c_gemm_clip_qkv@gemm_tc_kernel@6" "r02c_ln_bwd_clip@ln_bwd_rr_kernel@4"
