# N=1 sweep over the shipped configs and synthetic scenes; one JSON line per run into gpurun_out/sweep_r1.jsonl
out=gpurun_out/sweep_r1.jsonl; : > $out
for w in t01 sierp planets box_gi; do python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> $out; done
for n in 1000000 4000000 16000000 64000000; do python bench.py --workload box_caustics --photons $n --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> $out; done
for sc in soup:65536 soup:262144 soup:1048576 grid:2 grid:4 grid:8; do for a in 1 2; do python bench.py --workload synth --synth $sc --accel $a --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> $out; done; done
python bench.py --accel 2 --steps 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> $out
python bench.py --accel 0 --steps 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> $out
python - <<'PY'
import json
for l in open("gpurun_out/sweep_r1.jsonl"):
    try: d=json.loads(l)
    except Exception: print("BAD", l[:100]); continue
    print("%-70s accel=%-26s %9.1f Mrays/s %9.1f ms  lbvh_ms=%.2f stages=%s" % (d["config"]["workload"][:70], d["config"]["accel"], d["value"], d["ms_per_step"], d["accel_info"]["lbvh_build_ms"], d["stages_ms_rank0_last_step"]))
PY
