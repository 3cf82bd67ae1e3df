"""Benchmark of the SCT-GAN adversarial train step (BASELINE.json metric: GAN train-step tokens/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full adversarial optimisation step (generator + discriminator losses, backward, gradient
all-reduce when N > 1, three clips, AdamW) over one synthetic batch of the cfg3 per-GPU shard:
B = 32 contracts/GPU, contract seq S = path seq P = target seq T = 1024, default model.py hyper-parameters
(262.6 M parameters, dropout 0.3), bf16 tensor-core arithmetic with fp32 master weights / residual stream.
tokens = B*S contract tokens per step per GPU (weak scaling: 32 contracts per GPU at every N).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the public
API (SmartContractTrainer.train_step) from pinned HOST buffers with the H2D copies and a D2H read of the
loss inside the timed region.  `roofline` times the dominant kernel family (the tcgen05 GEMM) with CUDA
events on the launching stream in one extra instrumented step; `cpu_baseline` times the CPU oracle (a port
of the reference's PyTorch path: the Python reference itself cannot travel to the GPU box) on a bounded
sample.  `--impl reference` times that same CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gan_train_step_tokens_per_sec"
UNIT = "tokens/s"
CFG = dict(B=32, S=1024, P=1024, lines_per=12)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1404.8), d.get("hbm_gbs", 6449.4), "measured"
    return 1400.0, 6650.0, "fallback"


def synthetic_batch(B, S, P, vocab, seed, lines_per, device="cpu", pin=False):
    """ids ~ U{3..V-1}, prefix masks with len ~ U{L/2..L}, token_to_line = arange(S)//12,
    targets drawn independently (augmented-target branch), labels sparse."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (B, S), generator=g)
    ast = torch.randint(3, vocab, (B, P), generator=g)
    tgt = torch.randint(3, vocab, (B, S), generator=g)
    ls = torch.randint(S // 2, S + 1, (B,), generator=g)
    lp = torch.randint(max(1, P // 2), P + 1, (B,), generator=g)
    batch = dict(
        input_ids=ids, attention_mask=(torch.arange(S)[None, :] < ls[:, None]).long(),
        ast_input_ids=ast, ast_attention_mask=(torch.arange(P)[None, :] < lp[:, None]).long(),
        target_ids=tgt, token_to_line=(torch.arange(S) // lines_per)[None, :].expand(B, S).contiguous(),
        contract_vulnerabilities=(torch.rand(B, 8, generator=g) < 0.2).float(),
        vulnerable_lines=(torch.rand(B, 1024, 8, generator=g) < 0.01).float())
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return {k: v.to(device) for k, v in batch.items()} if device != "cpu" else batch


def redraw_1d_params(model, seed=0):
    """The reference zero-initialises all LayerNorm gammas and biases (model.py:290-294); that would make
    every logit 0.  gamma ~ N(1, 0.1), beta/bias ~ N(0, 0.02) (SURVEY §7 hard part 1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                noise = torch.randn(p.shape, generator=g)
                p.copy_(1.0 + 0.1 * noise if (n.endswith("weight")) else 0.02 * noise)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_step_time(sample_B, sample_S, sample_P, steps, warmup):
    """The reference's CPU path (oracle port of SCT-GAN/model.py forward + train.py loss + backward + clips +
    AdamW) on a bounded sample; returns (seconds per step, tokens per step, threads)."""
    from oracle import sct_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from sct_gan_b200 import SmartContractTransformer  # only for the key -> shape table of the default model

    cfg = dict(O.DEFAULT_CFG)
    shapes = {k: tuple(v.shape) for k, v in SmartContractTransformer(**cfg).state_dict().items()}
    sd = O.synth_state_dict(shapes, 0)
    names = [k for k, v in sd.items() if v.is_floating_point() and k not in ("pos_encoder.pe", "path_embedding.weight")]
    for k in names:
        sd[k].requires_grad_(True)
    sd["path_embedding.weight"] = sd["ast_embedding.weight"]
    batch = O.make_batch(sample_B, sample_S, sample_P, cfg["vocab_size"], seed=1234)
    state, times = {}, []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.forward_train(sd, cfg, batch, torch.float32)
        loss = O.step_losses(out, batch)["total_loss"]
        for k in names:
            sd[k].grad = None
        loss.backward()
        with torch.no_grad():
            params = {k: sd[k] for k in names}
            grads = {k: sd[k].grad for k in names}
            O.clip_and_adamw(params, grads, state)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), sample_B * sample_S, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sB, sS, sP = 1, CFG["S"], CFG["P"]
    sec, toks, threads = cpu_reference_step_time(sB, sS, sP, max(1, min(args.steps, 40)), max(0, min(args.warmup, 5)))
    v = toks / sec
    sample = f"{sB} contract(s) of the workload (S=P=T={sS}) per step, fp32, {threads} host threads, dropout off"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_name():
    tag = "cfg3 per-GPU shard" if (CFG["B"], CFG["S"], CFG["P"]) == (32, 1024, 1024) else "custom shape"
    return (f"{tag}: full adversarial train step (generator+discriminator losses, backward, clips, AdamW), "
            f"B={CFG['B']}/GPU, S=T={CFG['S']}, P={CFG['P']}, default SmartContractTransformer (262.6M params, dropout 0.3, "
            f"use_gan)")


# ------------------------------------------------------------------------------------------------
_RESULT_OUT = None


def emit(obj):
    """The result line goes to the process's ORIGINAL stdout, alone (see claim_stdout)."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def claim_stdout():
    """stdout must carry ONE JSON line, but native libraries write there too (NCCL prints its version banner on file
    descriptor 1 whatever NCCL_DEBUG_FILE says): keep a private handle on the real stdout for the result and point
    fd 1 at stderr for everybody else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=CFG["B"])
    ap.add_argument("--seq", type=int, default=CFG["S"])
    ap.add_argument("--path", type=int, default=None, help="execution-path sequence length (default: --seq)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vuln-heads", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of replaying the captured step")
    ap.add_argument("--ncu-step", action="store_true",
                    help="warm up, then run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--torch-profile", default=None, help="write a torch.profiler table of one step to this file and exit")
    args = ap.parse_args()
    CFG["B"], CFG["S"], CFG["P"] = args.batch, args.seq, (args.path or args.seq)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    from sct_gan_b200 import SmartContractTrainer, SmartContractTransformer, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries ONE JSON line: whatever NCCL logs (its version banner at NCCL_DEBUG=WARN, INFO traces) goes to
        # stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if "SCT_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["SCT_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.load().sct_device_check() == 0, _lib.last_error()
    W = max(3, args.warmup)
    B, S, P = CFG["B"], CFG["S"], CFG["P"]

    torch.manual_seed(0)
    model = SmartContractTransformer(use_gan=True, max_length=max(1024, S))
    redraw_1d_params(model, 0)
    model = model.to(dev)
    trainer = SmartContractTrainer(model, use_augmentation=True, use_gan=True,
                                   compute_vuln_heads=not args.no_vuln_heads, use_cuda_graph=not args.no_graph)
    n_lines = (S - 1) // CFG["lines_per"] + 1  # = token_to_line.max() + 1, known on the host by construction
    batch = synthetic_batch(B, S, P, model.vocab_size, 1234 + rank, CFG["lines_per"], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        trainer.train_step(batch, n_lines=n_lines)
    barrier()
    if args.ncu_step:
        torch.cuda.cudart().cudaProfilerStart()
        trainer.train_step(batch, n_lines=n_lines)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        emit({"ncu_step": "done", "launches_per_step_through_cabi": _lib.Stats.launches // (W + 1)})
        return
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile

        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            trainer.train_step(batch, n_lines=n_lines)
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        with open(args.torch_profile, "w") as f:
            f.write(f"wall_ms_under_profiler {wall * 1e3:.2f}\n")
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
        return
    # ---- value: K steps, inputs resident in HBM, CUDA events, max over ranks
    _lib.Stats.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        t_cpu = time.perf_counter()
        for _ in range(args.steps):
            trainer.train_step(batch, n_lines=n_lines)
        cpu_ms_per_step = (time.perf_counter() - t_cpu) * 1e3 / args.steps  # host time to enqueue (incl. its syncs)
        e1.record()
        barrier()
    launches = _lib.Stats.launches
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = world * B * S / (ms_per_step * 1e-3)

    # ---- e2e: public API from pinned host buffers, H2D + D2H inside the timed region
    host = synthetic_batch(B, S, P, model.vocab_size, 4321 + rank, CFG["lines_per"], pin=True)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    for _ in range(1):
        trainer.train_step(host, n_lines=n_lines)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        res = trainer.train_step(host, n_lines=n_lines)  # pinned host tensors -> H2D inside the step
        _ = res["total_loss"].item()
    e3.record()
    barrier()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * S / (ms2.item() / args.steps * 1e-3)

    # ---- roofline: one extra instrumented step, CUDA events around every launch of the dominant kernels
    tf_peak, hbm_peak, peak_src = peaks()
    _lib.Stats.timed = ("sct_gemm_bf16_nt", "sct_gemm_bf16_nn", "sct_gemm_bf16_tn", "sct_attn_fwd_strided", "sct_attn_bwd",
                        "sct_add_dropout_ln_fwd")
    _lib.Stats.events = []
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    was_graph, trainer.use_cuda_graph = trainer.use_cuda_graph, False  # events need real (eager) launches
    _lib.Stats.events = None
    trainer.train_step(batch, n_lines=n_lines)  # untimed: lets the eager allocator pool grow next to the graph's
    torch.cuda.synchronize()
    _lib.Stats.events = []
    e4.record()
    trainer.train_step(batch, n_lines=n_lines)
    e5.record()
    trainer.use_cuda_graph = was_graph
    torch.cuda.synchronize()
    ev, _lib.Stats.events = _lib.Stats.events, None
    step_ms_instr = e4.elapsed_time(e5)
    fam = {}
    for name, work, a, b in ev:
        f = fam.setdefault(name, [0.0, 0.0, 0])
        f[0] += work
        f[1] += a.elapsed_time(b)
        f[2] += 1
    gemm = [fam.get(k, [0, 0, 0]) for k in ("sct_gemm_bf16_nt", "sct_gemm_bf16_nn", "sct_gemm_bf16_tn")]
    g_work, g_ms, g_n = (sum(x[i] for x in gemm) for i in range(3))
    achieved = g_work / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    kernels = {}
    for k, (w, t, n) in fam.items():
        unit = "GB/s" if k == "sct_add_dropout_ln_fwd" else "TFLOP/s"
        rate = (w / (t * 1e-3) / (1e9 if unit == "GB/s" else 1e12)) if t > 0 else 0.0
        kernels[k] = {"launches": n, "ms": round(t, 3), "share_of_step": round(t / ms_per_step, 4),
                      "achieved": round(rate, 1), "unit": unit,
                      "frac": round(rate / (hbm_peak if unit == "GB/s" else tf_peak), 4)}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01m_gemm_traffic.json")
    if os.path.exists(tpath):  # DRAM bytes per launch of this kernel family from the committed ncu pass of the same step
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (sct_gemm_bf16_nt/nn/tn: tcgen05 + TMEM + TMA)",
                "achieved": round(achieved, 1), "peak": tf_peak, "unit": "TFLOP/s",
                "frac": round(achieved / tf_peak, 4), "traffic": traffic,
                "traffic_note": "ncu dram__bytes_read+write per launch, averaged over the 258 GEMM launches of one step "
                                "(profiles/r01m_gemm_traffic.json); algorithmic A+B+D bytes average ~160 MB per launch",
                "flops_per_launch": round(g_work / max(g_n, 1), 0),
                "peak_source": f"{peak_src} (sustained bf16)",
                "launches_per_step": g_n, "ms_per_step_in_kernel": round(g_ms, 3),
                "share_of_step": round(g_ms / ms_per_step, 4),
                "note": "timed with CUDA events around every launch in one extra eager (non-graph) step after the timed region",
                "by_call": kernels}

    in_sync = None
    if world > 1:  # data-parallel sanity: after the same number of steps every replica must hold identical weights
        probe = torch.stack([model.output_layer.weight.detach().double().sum(),
                             model.encoder.layers[0].linear1.weight.detach().double().sum(),
                             model.embedding.weight.detach().double().sum()])
        gathered = [torch.empty_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        in_sync = all(torch.equal(g, gathered[0]) for g in gathered)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, toks, threads = cpu_reference_step_time(1, S, P, 4, 1)
        cpu = {"value": toks / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"1 contract of the workload (S=P=T={S}) per step, full step incl. backward/clips/AdamW, fp32, "
                         f"1 warm-up + 4 timed steps, {sec:.1f} s/step"}
    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(), "global_batch": world * B, "seq_len": S, "path_len": P,
                       "parallelism": f"dp{world}", "vuln_heads": not args.no_vuln_heads,
                       "cuda_graph": not args.no_graph,
                       "l2": "no explicit flush: each step streams several GB of activations/weights (>> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "host_enqueue_ms_per_step": round(cpu_ms_per_step, 2), "dp_replicas_in_sync": in_sync,
        })
    if world > 1:
        # Tearing the NCCL communicator down while captured graphs still reference its collectives hangs
        # (seen on 2 GPUs): drain the device, make sure every rank got here, then leave without the destructor.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
