O=gpurun_out/r02; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -1 $O/gpu_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<PY
import json
for f in ("bench_cfg3.json","bench_cfg5.json"):
    d=json.loads(open("$O/"+f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["clocks"], (d.get("cpu_baseline") or {}).get("value"))
PY
