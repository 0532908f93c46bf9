#!/bin/bash
# compute-sanitizer passes over the hot path: memcheck and racecheck of smoke() (both precision modes: every
# kernel family of the forward), of the small block / ToMe / datapipe tests.  Summaries land in gpurun_out/ (copied to profiles/).
out=gpurun_out/r2_sanitizer.txt
: > $out
run() {  # tool, label, command...
  tool=$1; label=$2; shift 2
  echo "== compute-sanitizer --tool $tool : $label" >> $out
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 "$@" > gpurun_out/san_tmp.log 2>&1
  echo "exit code $?" >> $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke\[|Error|error:" gpurun_out/san_tmp.log | sort | uniq -c | head -20 >> $out
}
SEL='test_projection or test_temp_moe and 6-60-7-7 or test_qst_grounding_and_head and 8 or test_gather_rows or test_gemm_reduce_epilogue and 5-60-7-1'
run memcheck "smoke(): one forward of B=4 in fp32 and bf16 mode" python __graft_entry__.py smoke
run racecheck "smoke(): one forward of B=4 in fp32 and bf16 mode" python __graft_entry__.py smoke
run memcheck "block / datapipe / reduce-epilogue tests (small cases)" python -m pytest tests/test_gpu_blocks.py tests/test_gpu_datapipe.py tests/test_gpu_gemm.py -q -x -k "$SEL"
run memcheck "ToMe single steps + 23-step chain (golden, 2 frames)" python -m pytest tests/test_gpu_tome.py -q -x -k "single_steps or chain_577"
cat $out
