# 2-GPU data-parallel timing sweep (run under `gpurun --gpus 2`): bucket size / wire dtype of the gradient exchange
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-roofline > gpurun_out/r2_n2_$name.json 2> gpurun_out/r2_n2_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_n2_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],2), d['dp_replicas_in_sync'], d['clocks']['sm_mhz'])
"
}
run v3_dyn_b128 SCT_GEMM_DYNAMIC=1
run v3_static_b128 SCT_GEMM_DYNAMIC=0
run v3_dyn_b32 SCT_GEMM_DYNAMIC=1 SCT_DP_BUCKET_MB=32
python bench.py --steps 8 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/r2_n1_ref3.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2_n1_ref3.json').read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['clocks'])
"
