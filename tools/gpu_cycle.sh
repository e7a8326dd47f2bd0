#!/bin/bash
# One GPU measurement cycle, run on the box through gpurun:  gpurun --timeout 2400 -- 'bash tools/gpu_cycle.sh <tag> [tests|notests]'
# Order matters: every ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
#   1. pytest -m gpu                                     -> gpurun_out/<tag>_pytest.log
#   2. bench.py (both arms, default flags)               -> gpurun_out/<tag>_bench.json, <tag>_ref.json
#   3. launch list of bench.py (gpu__time_duration.sum)  -> gpurun_out/<tag>_launches.csv
#   4. ncu --set full of the lean k_trace / k_light / k_shade at 1080p x 4 spp (one batch per level) -> gpurun_out/<tag>_full.ncu-rep
set -u
TAG=${1:-r2}
MODE=${2:-tests}
OUT=gpurun_out
mkdir -p $OUT
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1 || { echo "build failed"; tail -20 $OUT/${TAG}_build.log; exit 1; }
if [ "$MODE" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 $OUT/${TAG}_bench.json
PROF="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 600 $PROF > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $PROF > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
PROF2="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --res 1920x1080 --spp 4"
timeout 600 $PROF2 > $OUT/${TAG}_plain2.log 2>&1 && \
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'k_trace|k_light|k_shade' -c 6 -f -o $OUT/${TAG}_full $PROF2 > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
