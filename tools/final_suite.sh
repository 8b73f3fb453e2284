#!/bin/bash
# One-box record of a build (1 GPU): GPU tests, bench line + kernel-class profile, ncu launch list of graph-replayed steps, ncu --set
# full of the kernels added in the second half of round 2, compute-sanitizer over their tests, same-box switch ablation.
#   gpurun --timeout 1500 -- 'bash tools/final_suite.sh r02c'
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/${TAG}_gpu_tests.log 2>&1; tail -3 $O/${TAG}_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 2>/dev/null > $O/${TAG}_bench_1gpu.json; cut -c1-200 $O/${TAG}_bench_1gpu.json
cp $O/bench_kernel_classes.json $O/${TAG}_kernel_classes.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/${TAG}_launches_step_graph.csv \
  python bench.py --steps 3 --warmup 3 --no-profile --no-cpu-baseline --no-gpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_tc_halo" -c 3 -o $O/${TAG}_wgrad_halo \
  python tools/kbench.py --only small_wgrad --reps 1 --no-graph > $O/${TAG}_ncu_wgh.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"dw_bwd_weight_tile|dw_s1d1_tile_kernel|dw_fwd_s2" -c 6 -o $O/${TAG}_dw_tma \
  python tools/kbench.py --only "dw_fwd 2x384,dw_bwd_weight 2x384,dw_fwd 2x48,dw_bwd_weight 2x48" --reps 1 --no-graph > $O/${TAG}_ncu_dw.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -q -x \
  -k "test_depthwise or deterministic or entry3x3s2 or conv2_32_64 or three_logit or pack" > $O/${TAG}_sanitizer_memcheck.log 2>&1; tail -4 $O/${TAG}_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -q -x \
  -k "test_conv_tcgen05 and (entry3x3s2 or conv2_32_64)" > $O/${TAG}_sanitizer_racecheck.log 2>&1; tail -4 $O/${TAG}_sanitizer_racecheck.log
bash tools/knob_ablation.sh $O/${TAG}_knob_ablation.jsonl
