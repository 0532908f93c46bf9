set -x
# one `ncu --set full` capture per dominant kernel (each command has already run clean without ncu); condensed by tests/ncu_summary.py
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r1_gemm_bf16x3_cg2 python tests/prof_gemm.py x3_in_src > gpurun_out/ncu1.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r1_gemm_f16 python tests/prof_gemm.py f16_slf_in > gpurun_out/ncu2.log 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r1_gemm_tf32_proj_cg2 python tests/prof_gemm.py tf32_proj > gpurun_out/ncu3.log 2>&1
$NCU -k regex:tome_merge_wavg -s 23 -c 1 -o gpurun_out/r1_tome_merge python tests/prof_tome.py 592 > gpurun_out/ncu4.log 2>&1
$NCU -k regex:tome_match_tiled -s 23 -c 1 -o gpurun_out/r1_tome_match_tiled python tests/prof_tome.py 592 > gpurun_out/ncu4b.log 2>&1
$NCU -k regex:attn_small_f16 -s 2 -c 1 -o gpurun_out/r1_attn_small_f16 python bench.py --steps 1 --warmup 1 --no-graph --hot-only > gpurun_out/ncu5.log 2>&1
$NCU -k regex:attn_mma16 -s 4 -c 1 -o gpurun_out/r1_attn_mma16_planes python bench.py --steps 1 --warmup 1 --no-graph --hot-only > gpurun_out/ncu6.log 2>&1
$NCU -k regex:gemm_attn_pipe_kernel -s 2 -c 2 -o gpurun_out/r1_gemm_attn_pipe python bench.py --steps 1 --warmup 1 --no-graph --hot-only > gpurun_out/ncu6b.log 2>&1
$NCU -k regex:gauss_reduce -s 3 -c 1 -o gpurun_out/r1_gauss_f32 python tests/prof_gauss.py 0 > gpurun_out/ncu7.log 2>&1
python tests/ncu_summary.py gpurun_out/r1_*.ncu-rep > gpurun_out/r1_ncu_summaries_v3.txt
rm -f gpurun_out/r1_*.ncu-rep
ls -la gpurun_out/
