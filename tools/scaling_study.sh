#!/bin/bash
# Scaling study of BASELINE's north star on N GPUs of one box:  gpurun --gpus N -- 'bash tools/scaling_study.sh N'
# (strong scaling: every workload's frame is fixed; lines -> gpurun_out/r2_scaling_n<N>.jsonl, one JSON line per (workload, N))
N=${1:-1}
SET=${2:-all}          # all | core (bun69k + box caustics 4 M / 64 M: what an 8-GPU call can afford)
OUT=gpurun_out/r2_scaling_n$N.jsonl
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/r2_scaling_build_n$N.log 2>&1
if [ "$SET" = "all" ]; then
[ -f scenes/gen/soup_1048576.cli ] || python tools/make_synth.py soup 1048576 > /dev/null
[ -f scenes/gen/grid_4.cli ] || python tools/make_synth.py grid 4 > /dev/null
fi
run() {
  if [ "$N" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu "$@" 2>> gpurun_out/r2_scaling.err | grep '^{' >> $OUT
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu "$@" 2>> gpurun_out/r2_scaling.err | grep '^{' >> $OUT; fi
}
run --workload bun69k
run --workload box_caustics
run --workload box_caustics --photons 64000000
if [ "$SET" = "all" ]; then
run --workload box_caustics --photons 16000000
run --workload synth --synth soup:1048576
run --workload synth --synth grid:4
run --workload sierp
fi
tail -n 7 $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], d['config']['workload'][:60], d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'))"
