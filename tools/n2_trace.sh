O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-roofline --trace $O/trace_n2.json > $O/trace_n2.log 2>&1
gzip -f $O/trace_n2.json.rank0 $O/trace_n2.json.rank1; ls -la $O | grep trace
