#!/bin/bash
# Same-box A/B of the round's switches (box-to-box spread of the step time is ~3 %, so only runs of ONE gpurun call compare):
# the default build, everything that has a switch turned off, and each switch off alone.  One JSON line per run.
#   gpurun --timeout 900 -- 'bash tools/knob_ablation.sh gpurun_out/r02_knob_ablation.jsonl'
OUT=${1:-gpurun_out/knob_ablation.jsonl}
: > $OUT
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-profile"
run() { # label, env...
  local label=$1; shift
  env "$@" $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(dict(run='$label', ms_per_step=d['ms_per_step'], samples_per_s=d['value'])))" >> $OUT
}
run default X=1
run all_off DEEPCAM_B200_TC_HALO=0 DEEPCAM_B200_FORK_ASPP=0 DEEPCAM_B200_EARLY_WEIGHTS=0 DEEPCAM_B200_DWW_TILE=0 DEEPCAM_B200_DW_PARITY_SPLIT=0 DEEPCAM_B200_DW_S2_TILE=0 DEEPCAM_B200_DW_TMA=0 DEEPCAM_B200_WGRAD_HALO=0
run halo_off DEEPCAM_B200_TC_HALO=0
run fork_aspp_off DEEPCAM_B200_FORK_ASPP=0
run early_weights_off DEEPCAM_B200_EARLY_WEIGHTS=0
run dww_tile_off DEEPCAM_B200_DWW_TILE=0
run dw_parity_split_off DEEPCAM_B200_DW_PARITY_SPLIT=0
run dw_s2_tile_off DEEPCAM_B200_DW_S2_TILE=0
run dw_tma_off DEEPCAM_B200_DW_TMA=0
run wgrad_halo_off DEEPCAM_B200_WGRAD_HALO=0
run deterministic_on DEEPCAM_B200_DETERMINISTIC=1
run fuse_bn_dw_on DEEPCAM_B200_FUSE_BN_DW=1
run default_again X=1
cat $OUT
