#!/bin/bash
# Round-2 evidence capture (run on a B200 via: gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh').
# Everything lands in gpurun_out/r02/; tools/summarise_profiles.py turns it into the files committed under profiles/.
set -u
O=gpurun_out/r02
mkdir -p $O
T="timeout 400"
$T python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"
for c in cfg3 cfg1 cfg2 cfg4 cfg5; do
  extra="--no-cpu-baseline"; [ $c = cfg3 ] && extra=""
  $T python bench.py --config $c --steps 10 --warmup 3 $extra > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
done
$T python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
$T python bench.py --config cfg1 --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm_cfg1.json 2> $O/bench_reference_arm_cfg1.err; echo "ref cfg1 rc=$?"
$T python tools/attn_bench.py > $O/attn_bench.txt 2>/dev/null; echo "attn_bench rc=$?"
$T python tools/gemm_bench.py > $O/gemm_bench.txt 2>/dev/null
$T python tools/ffn_bench.py > $O/ffn_bench.txt 2>/dev/null
$T python tools/rowwise_bench.py > $O/rowwise_bench.txt 2>/dev/null
$T python bench.py --steps 2 --warmup 3 --trace $O/trace.json > /dev/null 2>&1 && gzip -f $O/trace.json
# ncu (the plain runs above exited 0 with the same arguments): launch list of ONE step with DRAM bytes, then full sets
$T python bench.py --ncu-step --no-cpu-baseline > $O/ncu_step_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $O/launches.csv python bench.py --ncu-step --no-cpu-baseline > $O/ncu_step.log 2>&1
echo "ncu launches rc=$?"
$T python tools/ffn_bench.py > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 6 -c 4 -o $O/gemm_prof \
    python tools/ffn_bench.py > $O/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
$T python tools/attn_bench.py --once > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd2_kernel|attn_bwd_dkdv|attn_bwd_dq_ds" -s 3 -c 3 \
    -o $O/attn_prof python tools/attn_bench.py --once > $O/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la $O
