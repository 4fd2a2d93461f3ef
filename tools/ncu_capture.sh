#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then the ncu launch list of a short bench and `--set full`
# captures of the hot kernels at the train512 level-0 shapes.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-infer"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
TAG=${1:-r02}
# r02 adds the tensor-bound GEMMs of SURVEY 8d (tensor-pipe evidence) and the Conv2DTranspose / dropout variants
for spec in "pw_bneck2:gemm_tc_nt" "pw_dec4b1:gemm_tc_nt" "pw_enc4b2:gemm_tc_nt" "convt_dec4:gemm_tc_nt" "convt_dec3:gemm_tc_nt" "convt_dec2:gemm_tc_nt" \
            "convt_dec3_nodrop:gemm_tc_nt" "gemm64:gemm_tc_nt" "dw_bwd_mask:dwconv3x3_bwd_strip" "dw_bwd_aff:dwconv3x3_bwd_strip" "dw_bwd_drop256:dwconv3x3_bwd_strip" \
            "dw_fwd_aff:dwconv3x3_strip" "pw_bwd_fused64:pw_bwd_fused" "fused64:sepconv_fused" "fused128_64:sepconv_fused"; do
  name=${spec%%:*}; pat=${spec##*:}
  python tools/kernel_micro.py $name 2 > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o gpurun_out/prof_$name python tools/kernel_micro.py $name 2 > gpurun_out/ncu_$name.log 2>&1
done
# summarise on the box (gpurun merges at most 64 MiB back): tables into gpurun_out/profiles_out/, then keep only three reports
UNET_PROFILES_OUT=gpurun_out/profiles_out python tools/ncu_summarize.py $TAG
ls -la gpurun_out/*.ncu-rep
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-infer --breakdown > gpurun_out/live_bench.json 2> gpurun_out/live_breakdown.txt
python tools/share_table.py gpurun_out/live_breakdown.txt gpurun_out/launches.csv > gpurun_out/profiles_out/shares_$TAG.md
for f in gpurun_out/prof_*.ncu-rep; do case "$f" in *dw_bwd_mask*|*convt_dec3.ncu-rep|*pw_bneck2*) ;; *) rm -f "$f";; esac; done
