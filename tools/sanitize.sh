#!/bin/bash
# compute-sanitizer passes over the library's kernels (GPU box only; SURVEY.md section 5 asks for
# them).  Not part of the timed or tested path -- run by hand:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# memcheck / racecheck slow kernels down by 10-100x: the bounded waits of the tcgen05 kernels trap
# after ~3e9 cycles, so the small configurations of tests/gpu_checks.py are used, one process per
# check (a trap poisons the CUDA context).
set -u
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
CHECKS="warp_hrflow_s4 warp_lrflow_bd4 pool_upsample module_ops downsample_bd conv_tc_halo_64 conv_tc_tap_convT epilogues_tc conv_chain_1tile conv_chain_ragged_repeat step_golden_g1"
for tool in memcheck racecheck; do
  for c in $CHECKS; do
    timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 \
      python tests/gpu_diag.py --one $c > "$OUT/${tool}_$c.log" 2>&1
    echo "$tool $c rc=$? $(grep -c 'ERROR SUMMARY' "$OUT/${tool}_$c.log") $(grep 'ERROR SUMMARY' "$OUT/${tool}_$c.log" | tail -1)"
  done
done
