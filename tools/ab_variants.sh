#!/bin/bash
# A/B of tuning variants on the GPU box: bash tools/ab_variants.sh "<bench args>" lib1.so lib2.so ...   (results: gpurun_out/ab.jsonl)
args="$1"; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib" >> gpurun_out/ab.jsonl
  DRT_LIB="$lib" python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 $args 2>>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'stages_ms_rank0_last_step', 'frame_crc32', 'gpu_launches')}))" >> gpurun_out/ab.jsonl
done
