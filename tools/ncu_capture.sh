#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then the ncu launch list of a short bench and `--set full`
# captures of the hot kernels at the train512 level-0 shapes.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
for spec in "dw_fwd128:dwconv3x3_strip" "dw_fwd_aff:dwconv3x3_strip" "gemm64:gemm_tc_nt" "dw_bwd_mask:dwconv3x3_bwd_strip" "dw_bwd_aff:dwconv3x3_bwd_strip" "pw_bwd_fused64:pw_bwd_fused" "pw_bwd_fused128:pw_bwd_fused" "gemm_fold_dgrad:gemm_tc_nt" "gemm_fold_wgrad:gemm_tc_wgrad" "bn_bwd_apply:bn_bwd_apply" "fused64:sepconv_fused" "bn_act:bn_act"; do
  name=${spec%%:*}; pat=${spec##*:}
  python tools/kernel_micro.py $name 2 > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o gpurun_out/prof_$name python tools/kernel_micro.py $name 2 > gpurun_out/ncu_$name.log 2>&1
done
# summarise on the box (gpurun merges at most 64 MiB back): tables into gpurun_out/profiles_out/, then keep only three reports
UNET_PROFILES_OUT=gpurun_out/profiles_out python tools/ncu_summarize.py r01
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/prof_*.ncu-rep; do case "$f" in *pw_bwd_fused64*|*dw_bwd_aff*|*dw_fwd_aff*) ;; *) rm -f "$f";; esac; done
