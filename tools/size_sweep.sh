#!/bin/bash
# Synthetic size sweep of SURVEY 8(d) at N = 1 (soups 2^18 / 2^22, replicated-bunny grids k = 8 / 16; 2^16, 2^20 and k = 4 are in the scaling study):
#   gpurun -- 'bash tools/size_sweep.sh'   -> gpurun_out/r2_sizes.jsonl (bench lines incl. build_info: .cli interpretation, device ordering, host shape)
# Soups stop at 2^22 here: the drop-in boundary is the reference's TEXT scene format, 5 lines per triangle -- 2^24 triangles are 84 M lines / 3 GB of
# .cli and most of a minute of interpretation per rank before the first ray; the device-side ordering itself was measured to 2^24 objects (r2b_refbvh_time.jsonl).
OUT=gpurun_out/r2_sizes.jsonl
mkdir -p gpurun_out; : > $OUT
python __graft_entry__.py > gpurun_out/r2_sizes_build.log 2>&1
for s in soup:262144 soup:4194304 grid:8 grid:16; do
  timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu --no-e2e --workload synth --synth $s 2>> gpurun_out/r2_sizes.err | grep '^{' >> $OUT
done
python - <<'P'
import json
for l in open("gpurun_out/r2_sizes.jsonl"):
    d = json.loads(l); print(d["config"]["workload"][:28], d["value"], d["ms_per_step"], d["stages_ms_rank0_last_step"], d["build_info"])
P
