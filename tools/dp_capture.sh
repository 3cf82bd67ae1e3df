#!/bin/bash
# Data-parallel evidence on N GPUs of one box: gpurun --gpus N -- 'bash tools/dp_capture.sh N [indep]'
#   indep: also run N independent one-GPU replicas at the same time (the no-exchange reference with every GPU under load)
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q > $O/dp_tests_n2.log 2>&1; echo "dp tests rc=$?"; tail -2 $O/dp_tests_n2.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-roofline > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > $O/bench_n1_same_box_as_n$N.json 2>/dev/null
if [ "${2:-}" = indep ]; then
  pids=""
  for i in $(seq 0 $((N - 1))); do
    CUDA_VISIBLE_DEVICES=$i timeout 300 python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > $O/indep_n${N}_gpu$i.json 2>/dev/null &
    pids="$pids $!"
  done
  wait $pids
fi
python - <<PY
import glob, json
for f in ["bench_n$N.json", "bench_n1_same_box_as_n$N.json"] + sorted(g.split("/")[-1] for g in glob.glob("$O/indep_n${N}_gpu*.json")):
    try:
        d = json.loads(open("$O/" + f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["value"]), d.get("dp_replicas_in_sync"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
PY
