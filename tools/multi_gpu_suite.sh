#!/bin/bash
# Multi-GPU measurement set of one round (run under `gpurun --gpus 8`): DDP gradient parity over NCCL, the headline bench at
# N = 2 / 4 / 8 (with the exposed-communication measurement), the eval path of BASELINE.json configs[3] at N = 8 and the
# local-batch x LARS sweep points of configs[4].  Every JSON line lands in gpurun_out/<tag>_*.json(l); copy what should be
# judged into profiles/.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/multi_gpu_suite.sh r02'
TAG=${1:-rNN}
SHORT=${2:-}          # "short": DDP parity + the N = 1 / 2 / 4 / 8 bench lines + eval at N = 8, no sweep
OUT=gpurun_out
mkdir -p $OUT
run() { # nproc, extra args...
  local n=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" 2>>$OUT/${TAG}_multi.err
}
python -m pytest tests/test_parallel_nccl_gpu.py -m gpu -q -p no:cacheprovider > $OUT/${TAG}_ddp_test.log 2>&1; tail -3 $OUT/${TAG}_ddp_test.log
python bench.py --steps 20 --warmup 5 --no-profile --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_bench_1gpu_samebox.json 2>>$OUT/${TAG}_multi.err
for n in 2 4 8; do run $n --steps 20 --warmup 5 --no-profile > $OUT/${TAG}_bench_${n}gpu.json; done
run 8 --mode eval --eval-samples 64 > $OUT/${TAG}_eval_8gpu.json
python bench.py --mode eval --eval-samples 64 > $OUT/${TAG}_eval_1gpu.json 2>>$OUT/${TAG}_multi.err
if [ "$SHORT" = "short" ]; then for f in $OUT/${TAG}_bench_*gpu*.json $OUT/${TAG}_eval_*gpu.json; do echo "== $f"; cut -c1-200 $f; done; exit 0; fi
: > $OUT/${TAG}_sweep_lars.jsonl
for b in 1 4 8; do run 8 --steps 10 --warmup 3 --no-profile --optimizer lars --local-batch $b >> $OUT/${TAG}_sweep_lars.jsonl; done
run 8 --steps 10 --warmup 3 --no-profile --optimizer lars --local-batch 2 >> $OUT/${TAG}_sweep_lars.jsonl
run 2 --steps 10 --warmup 3 --no-profile --optimizer lars --local-batch 4 >> $OUT/${TAG}_sweep_lars.jsonl
run 4 --steps 10 --warmup 3 --no-profile --optimizer lars --local-batch 4 >> $OUT/${TAG}_sweep_lars.jsonl
for f in $OUT/${TAG}_bench_*gpu.json $OUT/${TAG}_eval_*gpu.json $OUT/${TAG}_sweep_lars.jsonl; do echo "== $f"; cut -c1-260 $f; done
