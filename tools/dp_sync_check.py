"""Data-parallel sanity check on N GPUs: after every optimisation step all replicas must hold bit-identical weights.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/dp_sync_check.py [--graph] [--torch-tail] [--steps 4]
Prints, per step, the parameters (if any) whose checksum differs between ranks."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic batch + 1-D parameter re-draw helpers)
from sct_gan_b200 import SmartContractTrainer, SmartContractTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--torch-tail", action="store_true")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seq", type=int, default=256)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = SmartContractTransformer(use_gan=True)
    bench.redraw_1d_params(model, 0)
    model = model.to(dev)
    trainer = SmartContractTrainer(model, use_augmentation=True, use_gan=True, use_cuda_graph=args.graph,
                                   fused_optimizer=not args.torch_tail)
    S = args.seq
    n_lines = (S - 1) // 12 + 1
    batch = bench.synthetic_batch(args.batch, S, S, model.vocab_size, 1234 + rank, 12, device=dev)
    names = [n for n, _ in model.named_parameters()]

    def checksums(grad):
        vals = []
        for _, p in model.named_parameters():
            t = p.grad if grad else p
            vals.append(t.detach().double().sum() if t is not None else torch.zeros((), dtype=torch.float64, device=dev))
        return torch.stack(vals)

    def compare(tag, grad):
        mine = checksums(grad)
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        bad = [names[i] for i in range(len(names)) if any(not torch.equal(v[i], allv[0][i]) for v in allv)]
        if rank == 0:
            print(f"{tag}: {len(bad)} of {len(names)} differ" + (f"  e.g. {bad[:6]}" if bad else ""), flush=True)

    compare("init weights", False)
    for step in range(args.steps):
        out = trainer.train_step(batch, n_lines=n_lines)
        torch.cuda.synchronize()
        compare(f"step {step} grads (after all-reduce + clips)", True)
        compare(f"step {step} weights", False)
        if rank == 0:
            print(f"   loss {float(out['total_loss']):.5f} stepped {bool(out['stepped'])}", flush=True)
    trainer.close()  # captured graphs hold the NCCL kernels: release them before the communicator goes
    dist.barrier()
    sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
