#!/bin/bash
# compute-sanitizer passes over the kernel tests (the race / synchronisation tooling the reference lacks).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      (default: memcheck racecheck synccheck)
# Runs a SMALL selection of tests/test_kernels_gpu.py (one case per kernel family: tcgen05 GEMM + epilogue kinds, implicit
# conv incl. stride-2 and upsample phases, the three tcgen05 attention kernels, norms, elementwise) — the sanitizer slows
# kernels down 10-100x.  Logs go to gpurun_out/sanitize_<tool>.log; exit status is nonzero if any tool reports an error.
# The mbarrier pipelines of the tcgen05 kernels are additionally protected by bounded waits (mbar_wait traps after ~12 s),
# so a deadlock fails the run instead of hanging the GPU.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
SEL='test_gemm_plain[256-320-320] or test_gemm_epilogues or test_gemm_geglu or test_gemm_bf16_token_stream_kinds[4096-320-320] or test_gemm_v2_layernorm_fold[False] or test_conv3x3[5-8-128-320] or test_conv3x3_stride2_implicit[3-16-64-160] or test_upsample_conv_phases[4-8-64-160] or test_conv3x3_bf16_out_with_statistics or test_attention_spatial[40-1-100] or test_attention_spatial[40-5-256] or test_attention_spatial[80-2-64] or test_attention_spatial[160-3-16] or test_attention_cross_77[40-3-1024] or test_attention_scta[80-1-4-16] or test_attention_scta[160-2-5-4] or test_attention_scta[40-2-4-32] or test_groupnorm[2-48-640-320] or test_layernorm[100-320] or test_rope_matches_oracle or test_conv_in_out or test_cfg_ddim_update_bit_exact'
rc=0
for tool in $TOOLS; do
  log=gpurun_out/sanitize_$tool.log
  echo "== compute-sanitizer --tool $tool" | tee "$log"
  timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 77 --launch-timeout 600 \
      python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" >> "$log" 2>&1
  st=$?
  tail -4 "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$log" | tail -3
  [ $st -ne 0 ] && rc=$st
done
exit $rc
