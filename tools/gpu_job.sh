set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python bench.py --workload encoder > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err
mkdir -p /tmp/ncu
NCU="ncu --set full --clock-control none --graph-profiling node --profile-from-start off"
$NCU --import-source on -k regex:dec_cross_attn_kernel -s 100 -c 1 -f -o /tmp/ncu/prof_dec_cross python tools/full_profile.py large-v3 120 > gpurun_out/ncu_a.log 2>&1
$NCU -k regex:"gemm_bf16_kernel|dec_ln_kernel|dec_self_attn_kernel" -s 2000 -c 14 -f -o /tmp/ncu/prof_dec_chain python tools/full_profile.py large-v3 120 > gpurun_out/ncu_c.log 2>&1
$NCU -k regex:"gemm_bf16_kernel|encoder_attention_kernel|log_mel_kernel|layernorm_kernel" -s 8 -c 8 -f -o /tmp/ncu/prof_encoder python tools/full_profile.py large-v3 120 > gpurun_out/ncu_d.log 2>&1
WDR_PROFILE_DTW_PASS=1 ncu --set full --clock-control none --profile-from-start off -k regex:"dtwp_cross_attn_kernel|dtwp_self_attn_kernel" -s 4 -c 2 -f -o /tmp/ncu/prof_dtwpass python tools/dtw_pass_profile.py large-v3 120 > gpurun_out/ncu_e.log 2>&1
for f in prof_dec_cross prof_dec_chain prof_encoder prof_dtwpass; do python tools/ncu_summary.py /tmp/ncu/$f.ncu-rep gpurun_out/${f}_summary.csv; done
ncu -i /tmp/ncu/prof_dec_cross.ncu-rep --page details --csv > gpurun_out/prof_dec_cross_details.csv 2>/dev/null
ls -la /tmp/ncu; du -sh gpurun_out
cat gpurun_out/bench_full.json | cut -c1-1200
