set -x
# round 2: one `ncu --set full` capture per changed / headline kernel (each command has already run clean without ncu); condensed by
# tests/ncu_summary.py into gpurun_out/r2_ncu_summaries.txt (copied to profiles/)
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_bf16x3_in_src_30720 python tests/prof_gemm.py x3_in_src > gpurun_out/ncu1.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_bf16x3_w1_reduce_epi4 python tests/prof_gemm.py x3_w1_reduce > gpurun_out/ncu2.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_tf32_proj_215040 python tests/prof_gemm.py tf32_proj_full > gpurun_out/ncu3.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_f16_slf_in_215040 python tests/prof_gemm.py f16_slf_in_full > gpurun_out/ncu4.log 2>&1
$NCU -k regex:gauss_reduce -s 3 -c 1 -o gpurun_out/r2_gauss_f32 python tests/prof_gauss.py 0 > gpurun_out/ncu5.log 2>&1
$NCU -k regex:gauss_reduce -s 60 -c 1 -o gpurun_out/r2_gauss_bf16 python tests/prof_gauss.py 0 > gpurun_out/ncu6.log 2>&1
$NCU -k regex:gather_rows -s 4 -c 1 -o gpurun_out/r2_gather_rows python tests/prof_datapipe.py 2304 > gpurun_out/ncu7.log 2>&1
$NCU -k regex:moe_attn_rg -s 2 -c 1 -o gpurun_out/r2_moe_attn_rg python bench.py --steps 1 --warmup 1 --no-graph --hot-only > gpurun_out/ncu8.log 2>&1
$NCU -k regex:attn_long -s 1 -c 1 -o gpurun_out/r2_attn_long python -m pytest tests/test_gpu_tome_block.py -q -x -k "stack and b577 and bf16" > gpurun_out/ncu9.log 2>&1
python tests/ncu_summary.py gpurun_out/r2_gemm_bf16x3_in_src_30720.ncu-rep gpurun_out/r2_gemm_bf16x3_w1_reduce_epi4.ncu-rep gpurun_out/r2_gemm_tf32_proj_215040.ncu-rep gpurun_out/r2_gemm_f16_slf_in_215040.ncu-rep gpurun_out/r2_gauss_f32.ncu-rep gpurun_out/r2_gauss_bf16.ncu-rep gpurun_out/r2_gather_rows.ncu-rep gpurun_out/r2_moe_attn_rg.ncu-rep gpurun_out/r2_attn_long.ncu-rep gpurun_out/r2_gemm_attn_pipe.ncu-rep > gpurun_out/r2_ncu_summaries.txt
rm -f gpurun_out/r2_gemm_bf16x3*.ncu-rep gpurun_out/r2_gemm_tf32*.ncu-rep gpurun_out/r2_gemm_f16*.ncu-rep gpurun_out/r2_gather*.ncu-rep gpurun_out/r2_moe*.ncu-rep gpurun_out/r2_attn_long.ncu-rep
ls -la gpurun_out/ | head -40
