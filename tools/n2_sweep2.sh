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
run default2
run simple NCCL_PROTO=Simple
run simple_c32 NCCL_PROTO=Simple NCCL_MIN_CTAS=32
run b64 SCT_DP_BUCKET_MB=64
run simple_b64 NCCL_PROTO=Simple SCT_DP_BUCKET_MB=64
run fp32wire_simple NCCL_PROTO=Simple SCT_DP_GRAD_DTYPE=fp32
