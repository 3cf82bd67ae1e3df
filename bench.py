"""Benchmark of the SCT-GAN adversarial train step (BASELINE.json metric: GAN train-step tokens/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg1|cfg2|cfg3|cfg4|cfg5]

Default workload = cfg3 per-GPU shard (BASELINE.json configs[2], the configuration the metric is quoted on): one
"step" = one full adversarial optimisation step (generator + discriminator losses, vulnerability heads, syntax penalty,
line metrics, backward, gradient all-reduce when N > 1, three clips, AdamW) over one synthetic batch of B = 32
contracts/GPU, contract seq S = path seq P = target seq T = 1024, default model.py hyper-parameters (262.6 M
parameters, dropout 0.3), bf16 tensor-core arithmetic with fp32 master weights / residual stream.  tokens = B*S
contract tokens per step per GPU (weak scaling: 32 contracts per GPU at every N).
Other named workloads (BASELINE.json configs): cfg1 (B=8, S=512, P=128, full step), cfg2 (generator-only teacher-forced
step: cross-entropy loss only, no vulnerability heads / discriminator loss, B=64, S=P=512), cfg4 (long context, S=T=4096,
P=1024, B=8), cfg5 (generation: B=128, S=512, 512 new tokens with the KV cache; metric = new tokens/s).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the public API
(SmartContractTrainer.train_step / model.forward) from pinned HOST buffers with the H2D copies and a D2H read of the
result inside the timed region.  `roofline` times EVERY C-ABI call of one extra eager step with CUDA events on the
launching stream (`by_call`: every kernel family against its own roofline); the headline family is the tcgen05 GEMM.
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference (oracle/_ref, byte-compiled from
/root/reference/SCT-GAN by oracle/build_ref.py) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import glob
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
CONFIGS = {
    "cfg1": dict(B=8, S=512, P=128, mode="full", note="BASELINE.json configs[0]: the reference's CPU-runnable case"),
    "cfg2": dict(B=64, S=512, P=512, mode="generator", note="configs[1]: generator-only teacher-forced step"),
    "cfg3": dict(B=32, S=1024, P=1024, mode="full", note="configs[2]: per-GPU shard of B=256 over 8 GPUs"),
    "cfg4": dict(B=8, S=4096, P=1024, mode="full", note="configs[3]: long contract, max_length=4096"),
    "cfg5": dict(B=128, S=512, P=512, mode="decode", new_tokens=512, note="configs[4]: generation with KV cache"),
}
CFG = dict(CONFIGS["cfg3"], name="cfg3", lines_per=12)


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


class StubTokenizer:
    """Stands in for the RoBERTa tokenizer SoliditySyntaxLoss is built from (train.py:284-311; no network here): a
    fixed token -> id table over the synthetic vocabulary, so the syntax-penalty scan runs its real rules."""
    unk_token_id = 3

    def __init__(self):
        from sct_gan_b200.syntax import KEYWORD_FOLLOWERS

        toks = sorted(set(KEYWORD_FOLLOWERS) | {f for v in KEYWORD_FOLLOWERS.values() for f in v}
                      | {";", "(", ")", "{", "}"})
        self.table = {t: 100 + 37 * i for i, t in enumerate(toks)}

    def convert_tokens_to_ids(self, tok):
        return self.table.get(tok, self.unk_token_id)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.skip = [], None, index, 0

    def start(self):
        """Launch nvidia-smi (it needs a second or more to come up on an 8-GPU box: call before the warm-up steps)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def ready(self):
        return bool(self.rows) or self.proc is None or self.proc.poll() is not None

    def __enter__(self):
        """Start of the timed region: samples from before it are not used."""
        if self.proc is None and not self.rows:
            self.start()
        self.skip = len(self.rows)
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
        rows = self.rows[self.skip:] or self.rows  # samples taken inside the region (all of them if it was too short)
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) on a bounded sample of the workload
def cpu_sample_shape():
    """cfg1 is the reference's own CPU-runnable case and runs whole; the larger workloads are sampled as
    contracts of the same (S, P, T) — the reference's cost per contract does not depend on how many share a batch,
    except for the spatial-penalty loop (B*1024 Python iterations at S = 1024), which a small sample UNDER-counts."""
    if CFG["name"] == "cfg1":
        return CFG["B"], CFG["S"], CFG["P"]
    return 1, CFG["S"], CFG["P"]


def cpu_reference_run(budget_s, max_steps, warmup):
    """Times the CPU arm for about `budget_s` seconds.  Returns a dict for `cpu_baseline`."""
    from oracle import ref_loader

    sB, sS, sP = cpu_sample_shape()
    threads = os.cpu_count() or 1
    if CFG["mode"] == "decode":
        return cpu_port_decode(budget_s, threads)
    if ref_loader.available():
        from oracle import ref_step

        rs = ref_step.ReferenceStep(max_length=max(1024, sS), threads=threads)
        batch = ref_step.synthetic_batch(sB, sS, sP, 50265, 1234)
        kind, step = "reference", (lambda: rs.step(batch))
        what = "UNMODIFIED reference (oracle/_ref: SCT-GAN/model.py + train.py loss classes, restated batch-loop body)"
    else:  # /root/reference was not present when the repo was built: the oracle port stands in
        kind, step = "port", port_step_fn(sB, sS, sP)
        what = "oracle port of the reference path (oracle/_ref not built)"
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + max_steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start + dt > budget_s and times:
            break
    sec = sum(times) / len(times)
    mode = "generator + discriminator + vulnerability-head losses" if CFG["mode"] == "full" else \
        "full reference step (the reference has no generator-only mode)"
    return {"value": sB * sS / sec, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{sB} contract(s) of the workload (S=T={sS}, P={sP}) per step; {what}; {mode}, backward, 3 clips, "
                      f"AdamW; fp32, dropout 0.3, {threads} host threads; {warmup} warm-up + {len(times)} timed step(s), "
                      f"{sec:.2f} s/step"}, sec


def port_step_fn(sB, sS, sP):
    from oracle import sct_oracle as O
    from sct_gan_b200 import SmartContractTransformer  # only for the key -> shape table of the default model

    cfg = dict(O.DEFAULT_CFG, max_length=max(1024, sS))
    shapes = {k: tuple(v.shape) for k, v in SmartContractTransformer(**cfg).state_dict().items()}
    sd = O.synth_state_dict(shapes, 0)
    names = [k for k, v in sd.items() if v.is_floating_point() and k not in ("pos_encoder.pe", "path_embedding.weight")]
    for k in names:
        sd[k].requires_grad_(True)
    sd["path_embedding.weight"] = sd["ast_embedding.weight"]
    batch = O.make_batch(sB, sS, sP, cfg["vocab_size"], seed=1234)
    state = {}

    def step():
        out = O.forward_train(sd, cfg, batch, torch.float32)
        loss = O.step_losses(out, batch)["total_loss"]
        for k in names:
            sd[k].grad = None
        loss.backward()
        with torch.no_grad():
            O.clip_and_adamw({k: sd[k] for k in names}, {k: sd[k].grad for k in names}, state)

    return step


def cpu_port_decode(budget_s, threads):
    """cfg5 on CPU: the reference's sampling loop (model.py:862-930: the whole prefix is re-decoded for every new
    token) as restated by the oracle (greedy), 2 contracts x 8 new tokens.  The reference's own loop cannot be
    bounded without editing it (it runs to max_length or its stop rules), hence kind = "port"."""
    from oracle import sct_oracle as O
    from sct_gan_b200 import SmartContractTransformer

    torch.set_num_threads(threads)
    cfg = dict(O.DEFAULT_CFG)
    shapes = {k: tuple(v.shape) for k, v in SmartContractTransformer(**cfg).state_dict().items()}
    sd = O.synth_state_dict(shapes, 0)
    sB, n_new = 2, 8
    batch = O.make_batch(sB, CFG["S"], CFG["P"], cfg["vocab_size"], seed=1234)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.generate_greedy(sd, cfg, batch, n_new)
    sec = time.perf_counter() - t0
    return {"value": sB * n_new / sec, "unit": "new tokens/s", "cores": threads, "kind": "port",
            "sample": f"{sB} contracts (S=P={CFG['S']}) x {n_new} new tokens, oracle port of the reference's re-decode-the-"
                      f"prefix sampling loop, fp32, {threads} host threads, {sec:.1f} s"}, sec / n_new


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu, sec = cpu_reference_run(budget_s=150.0, max_steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    v = cpu["value"]
    emit({
        "impl": "reference", "metric": metric_name(), "value": v, "unit": cpu["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_record(int(os.environ.get("WORLD_SIZE", "1")), args),
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": cpu["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def config_record(world, args):
    """The `config` object of the result line: the same for both arms (the reference arm times a bounded sample of this
    workload on the host, described in its `cpu_baseline.sample`)."""
    B, S, P = CFG["B"], CFG["S"], CFG["P"]
    if CFG["mode"] == "decode":
        return {"workload": workload_name(), "global_batch": world * B, "seq_len": S, "new_tokens": CFG["new_tokens"],
                "parallelism": f"dp{world} (independent replicas)", "cuda_graph": True,
                "l2": "K/V caches + weights per step (> 2 GB) exceed the 126 MB L2"}
    full = CFG["mode"] == "full"
    heads = full and not args.no_vuln_heads
    return {"workload": workload_name(), "global_batch": world * B, "seq_len": S, "path_len": P,
            "parallelism": f"dp{world}", "vuln_heads": heads, "syntax_penalty": full, "line_metrics": heads,
            "cuda_graph": not args.no_graph,
            "l2": "no explicit flush: each step streams several GB of activations/weights (>> 126 MB L2)"}


def metric_name():
    return "generation_new_tokens_per_sec" if CFG["mode"] == "decode" else METRIC


def workload_name():
    name, B, S, P = CFG["name"], CFG["B"], CFG["S"], CFG["P"]
    tag = f"{name} ({CFG.get('note', '')})" if (B, S, P) == tuple(CONFIGS.get(name, {}).get(k) for k in "BSP") else \
        f"{name} with custom shape"
    model = "default SmartContractTransformer (262.6M params, dropout 0.3, use_gan)"
    if CFG["mode"] == "decode":
        return (f"{tag}: autoregressive generation (inference.py / model.py:862-930 path), B={B}, S=P={S}, "
                f"{CFG['new_tokens']} new tokens, greedy, KV cache, {model}")
    if CFG["mode"] == "generator":
        return (f"{tag}: generator-only teacher-forced step (token cross-entropy only: no vulnerability heads, no "
                f"discriminator loss), forward + backward + clips + AdamW, B={B}/GPU, S=T={S}, P={P}, {model}")
    return (f"{tag}: full adversarial train step (generator + discriminator + vulnerability-head losses, syntax penalty, "
            f"line metrics, backward, clips, AdamW), B={B}/GPU, S=T={S}, P={P}, {model}")


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


# kernel family -> (roofline class, unit); everything a C-ABI call can be is listed (bench `by_call`)
TENSOR_CALLS = ("sct_gemm_bf16_nt", "sct_gemm_bf16_nn", "sct_gemm_bf16_tn", "sct_attn_fwd_strided", "sct_attn_bwd",
                "sct_gemm_bf16_nt_gelu", "sct_gemm_bf16_nn_mul")


def traffic_record():
    """DRAM bytes per launch of the headline kernel family from the newest committed ncu pass (the file names the
    commit and command it was captured with); None when no capture is committed."""
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "*gemm_traffic.json")))
    if not cands:
        return None, None
    with open(cands[-1]) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), {"file": os.path.relpath(cands[-1], ROOT), "captured_at_commit": d.get("commit"),
                                            "launches": d.get("launches"), "command": d.get("command")}


def instrumented_roofline(run_step, ms_per_step):
    """One extra EAGER pass with CUDA events around every C-ABI call (events need real launches, not a graph replay)."""
    from sct_gan_b200 import _lib

    tf_peak, hbm_peak, peak_src = peaks()
    _lib.Stats.timed = None  # every call
    _lib.Stats.events = None
    run_step()  # untimed: lets the eager allocator pool grow next to the graph's
    torch.cuda.synchronize()
    _lib.Stats.events = []
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    run_step()
    e5.record()
    torch.cuda.synchronize()
    ev, _lib.Stats.events = _lib.Stats.events, None
    fam = {}
    for name, work, a, b in ev:
        f = fam.setdefault(name, [0.0, 0.0, 0])
        f[0] += work
        f[1] += a.elapsed_time(b)
        f[2] += 1
    gemm = [fam.get(k, [0, 0, 0]) for k in TENSOR_CALLS if k.startswith("sct_gemm")]
    g_work, g_ms, g_n = (sum(x[i] for x in gemm) for i in range(3))
    achieved = g_work / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    kernels = {}
    for k, (w, t, n) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        tensor = k in TENSOR_CALLS
        unit = "TFLOP/s" if tensor else "GB/s"
        rate = (w / (t * 1e-3) / (1e12 if tensor else 1e9)) if t > 0 else 0.0
        kernels[k] = {"launches": n, "ms": round(t, 3), "share_of_step": round(t / ms_per_step, 4),
                      "achieved": round(rate, 1), "unit": unit, "bound": "tensor" if tensor else "hbm",
                      "frac": round(rate / (tf_peak if tensor else hbm_peak), 4) if w > 0 else None}
    traffic, tsrc = traffic_record()
    return {"bound": "tensor", "kernel": "gemm_kernel (sct_gemm_bf16_nt/nn/tn and fused-epilogue variants: tcgen05 + TMEM + TMA)",
            "achieved": round(achieved, 1), "peak": tf_peak, "unit": "TFLOP/s",
            "frac": round(achieved / tf_peak, 4), "traffic": traffic, "traffic_source": tsrc,
            "flops_per_launch": round(g_work / max(g_n, 1), 0),
            "peak_source": f"{peak_src} (sustained bf16 {tf_peak} TFLOP/s; HBM copy {hbm_peak} GB/s for the hbm-bound calls)",
            "launches_per_step": g_n, "ms_per_step_in_kernel": round(g_ms, 3),
            "share_of_step": round(g_ms / ms_per_step, 4),
            "instrumented_step_ms": round(e4.elapsed_time(e5), 3),
            "note": "CUDA events around every C-ABI call of one extra eager (non-graph) step after the timed region; "
                    "hbm-bound calls: algorithmic bytes / time against the measured copy bandwidth",
            "by_call": kernels}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--path", type=int, default=None, help="execution-path sequence length")
    ap.add_argument("--new-tokens", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vuln-heads", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of replaying the captured step")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true",
                    help="warm up, then run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--torch-profile", default=None, help="write a torch.profiler table of one step to this file and exit")
    ap.add_argument("--trace", default=None, help="write a chrome trace (CUPTI kernel timeline) of two steps and exit")
    args = ap.parse_args()
    CFG.update(CONFIGS[args.config], name=args.config)
    if args.batch:
        CFG["B"] = args.batch
    if args.seq:
        CFG["S"] = args.seq
        if CFG["name"] != "cfg4" and not args.path:
            CFG["P"] = args.seq
    if args.path:
        CFG["P"] = args.path
    if args.new_tokens:
        CFG["new_tokens"] = args.new_tokens
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    from sct_gan_b200 import SmartContractTrainer, SmartContractTransformer, _lib
    from sct_gan_b200.syntax import SoliditySyntaxRules

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(0)
    model = SmartContractTransformer(use_gan=True, max_length=max(1024, S))
    redraw_1d_params(model, 0)
    model = model.to(dev)
    if CFG["mode"] == "decode":
        run_decode(args, model, dev, world, rank, local, barrier)
        return finish(world)
    full = CFG["mode"] == "full"
    heads = full and not args.no_vuln_heads
    rules = SoliditySyntaxRules(StubTokenizer(), model.vocab_size, dev) if full else None
    trainer = SmartContractTrainer(model, use_augmentation=True, use_gan=full, compute_vuln_heads=heads,
                                   use_cuda_graph=not args.no_graph, syntax_rules=rules, line_metrics=heads)
    n_lines = (S - 1) // CFG["lines_per"] + 1  # = token_to_line.max() + 1, known on the host by construction
    batch = synthetic_batch(B, S, P, model.vocab_size, 1234 + rank, CFG["lines_per"], device=dev)

    clk = ClockSampler(local)
    if not (args.ncu_step or args.trace or args.torch_profile):
        clk.start()
    for _ in range(W):
        trainer.train_step(batch, n_lines=n_lines)
    t_wait = time.perf_counter()
    while not clk.ready() and time.perf_counter() - t_wait < 10.0:  # keep the GPU under load until nvidia-smi reports
        trainer.train_step(batch, n_lines=n_lines)
        torch.cuda.synchronize()
    barrier()
    if args.ncu_step:
        torch.cuda.cudart().cudaProfilerStart()
        trainer.train_step(batch, n_lines=n_lines)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        emit({"ncu_step": "done", "launches_per_step_through_cabi": _lib.Stats.launches // (W + 1)})
        return finish(world, trainer)
    if args.trace:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                trainer.train_step(batch, n_lines=n_lines)
            torch.cuda.synchronize()
        prof.export_chrome_trace(args.trace if world == 1 else f"{args.trace}.rank{rank}")
        return finish(world, trainer)
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile

        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            trainer.train_step(batch, n_lines=n_lines)
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        with open(args.torch_profile, "w") as f:
            f.write(f"wall_ms_under_profiler {wall * 1e3:.2f}\n")
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=80, max_name_column_width=90))
        return finish(world, trainer)
    # ---- value: K steps, inputs resident in HBM, CUDA events, max over ranks
    _lib.Stats.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clk:
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

    roofline = None
    if not args.no_roofline:
        was_graph, trainer.use_cuda_graph = trainer.use_cuda_graph, False
        roofline = instrumented_roofline(lambda: trainer.train_step(batch, n_lines=n_lines), ms_per_step)
        trainer.use_cuda_graph = was_graph

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
        cpu, _ = cpu_reference_run(budget_s=25.0, max_steps=3, warmup=1 if CFG["name"] != "cfg1" else 0)
    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_record(world, args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "host_enqueue_ms_per_step": round(cpu_ms_per_step, 2), "dp_replicas_in_sync": in_sync,
        })
    finish(world, trainer)


def run_decode(args, model, dev, world, rank, local, barrier):
    """cfg5: B sequences, `new_tokens` greedy tokens each through the KV-cached decode step (one CUDA graph per step
    signature).  A "step" = one whole generation call; value = new tokens / s."""
    import torch.distributed as dist

    from sct_gan_b200 import _lib

    B, S, P, n_new = CFG["B"], CFG["S"], CFG["P"], CFG["new_tokens"]
    model.eval()
    batch = synthetic_batch(B, S, P, model.vocab_size, 1234 + rank, CFG["lines_per"], device=dev)
    host = synthetic_batch(B, S, P, model.vocab_size, 4321 + rank, CFG["lines_per"], pin=True)
    keys = ("input_ids", "attention_mask", "ast_input_ids", "ast_attention_mask")

    def gen(b):
        with torch.no_grad():
            return model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], ast_input_ids=b["ast_input_ids"],
                         ast_attention_mask=b["ast_attention_mask"], target_ids=None, greedy=True,
                         max_new_tokens=n_new, compute_vuln_heads=False)["generated_sequence"]

    W = max(3, args.warmup)
    for _ in range(W):
        gen(batch)
    barrier()
    if args.trace:
        from torch.profiler import ProfilerActivity, profile

        CFG["new_tokens"], n_new = 24, 24  # a short call is enough to see the step's kernels
        gen(batch)
        gen(batch)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            gen(batch)
            torch.cuda.synchronize()
        prof.export_chrome_trace(args.trace)
        return
    _lib.Stats.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            gen(batch)
        e1.record()
        barrier()
    launches = _lib.Stats.launches
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = world * B * n_new / (ms_per_step * 1e-3)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    d2h = 0
    for _ in range(args.steps):
        seq = gen({k: host[k].to(dev, non_blocking=True) for k in keys}).cpu()
        d2h = seq.numel() * seq.element_size()
    e3.record()
    barrier()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * n_new / (ms2.item() / args.steps * 1e-3)
    # HBM floor of one decode step: decoder + vocab weights once (bf16) + self-attention K/V read up to the
    # current position + cross-attention K/V of the S memory rows, per layer
    d, L, V = model.d_model, len(model.decoder.layers), model.vocab_size
    w_bytes = 2.0 * (L * (3 * d * d + d * d + d * d + 2 * d * d + d * d + 2 * d * 2048) + V * d)
    self_kv = sum(B * (t + 1) * 2 * d * 2.0 * L for t in range(n_new)) / n_new
    cross_kv = B * S * 2 * d * 2.0 * L * 0.75  # ragged masks: keys past the last valid one are skipped (~3/4 kept)
    floor_bytes = w_bytes + self_kv + cross_kv
    _, hbm_peak, peak_src = peaks()
    step_ms = ms_per_step / n_new
    achieved = floor_bytes / (step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "decode step (KV-cached decoder layer stack + vocab projection)",
                "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                "traffic": None, "bytes_per_decode_step": round(floor_bytes), "ms_per_decode_step": round(step_ms, 4),
                "peak_source": f"{peak_src} (HBM copy)",
                "note": "algorithmic bytes of one decode step (weights once, K/V caches once) / measured time per step"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_port_decode(25.0, os.cpu_count() or 1)
    if rank == 0:
        emit({
            "metric": metric_name(), "value": value, "unit": "new tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_record(world, args),
            "e2e": {"value": e2e_value, "unit": "new tokens/s",
                    "h2d_bytes_per_step": sum(host[k].numel() * host[k].element_size() for k in keys),
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
        })


def finish(world, trainer=None):
    """Orderly teardown: release the captured graphs (they hold the NCCL kernels), then destroy the process group.
    A watchdog ends the process if the communicator teardown stalls anyway (the JSON line is already out)."""
    if trainer is not None:
        trainer.close()
    if world > 1:
        import threading

        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()

        def bail():
            sys.stderr.write("bench.py: process-group teardown stalled; leaving without it\n")
            sys.stderr.flush()
            os._exit(0)

        t = threading.Timer(30.0, bail)
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    main()
