# 2-GPU data-parallel timing sweep (run under `gpurun --gpus 2`): what a replica pays for the gradient exchange
O=gpurun_out/r02; mkdir -p $O
show() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'], 2), d.get('dp_replicas_in_sync'), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'failed', e)
PY
}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/n2_$name.json 2> $O/n2_$name.err
  show $O/n2_$name.json
}
# two independent replicas, one per GPU, at the same time: the no-exchange reference with both GPUs under load
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > $O/n2_indep_gpu0.json 2>/dev/null &
P0=$!
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > $O/n2_indep_gpu1.json 2>/dev/null &
P1=$!
wait $P0 $P1
show $O/n2_indep_gpu0.json $O/n2_indep_gpu1.json
run default
run ctas4 NCCL_MAX_CTAS=4
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run dyn SCT_GEMM_DYNAMIC=1
run dyn_ctas8 SCT_GEMM_DYNAMIC=1 NCCL_MAX_CTAS=8
