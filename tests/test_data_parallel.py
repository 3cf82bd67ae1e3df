"""Host-side data-parallel logic on CPU (gloo, world_size 2): bucketed gradient mean all-reduce reproduces the
single-process gradient of the concatenated batch, including parameters that receive no gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(16, 32)
        self.dead = nn.Embedding(10, 16)  # like disc_grammar_embedding: never used
        self.b = nn.Linear(32, 4)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _worker(rank, world, port, bucket_bytes, overlapped, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sct_gan_b200.trainer import OverlappedGradReducer, allreduce_mean_grads

    torch.manual_seed(1)
    x = torch.randn(8, 16)
    y = torch.randn(8, 4)
    m = Toy()
    shard = slice(rank * 4, rank * 4 + 4)
    if overlapped:  # buckets launched from post-accumulate-grad hooks during backward, joined afterwards
        red = OverlappedGradReducer(list(m.parameters()), world, None, bucket_bytes)
        ((m(x[shard]) - y[shard]) ** 2).mean().backward()
        red.finish()
    else:
        ((m(x[shard]) - y[shard]) ** 2).mean().backward()
        allreduce_mean_grads(list(m.parameters()), world, None, bucket_bytes)
    # confidence exchange used for the 0.3 / 0.8 GAN branches (trainer.compute_losses)
    c = torch.tensor([0.2 + 0.4 * rank])
    dist.all_reduce(c)
    c /= world
    if rank == 0:
        ret["grads"] = {n: (p.grad.clone() if p.grad is not None else None) for n, p in m.named_parameters()}
        ret["conf"] = c.item()
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_process():
    world = 2
    for bucket_bytes, overlapped in ((64, False), (1 << 20, False), (64, True), (1 << 20, True)):
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), bucket_bytes, overlapped, ret), nprocs=world, join=True)
        torch.manual_seed(1)
        x = torch.randn(8, 16)
        y = torch.randn(8, 4)
        m = Toy()
        ((m(x) - y) ** 2).mean().backward()
        for n, p in m.named_parameters():
            g = ret["grads"][n]
            if p.grad is None:
                assert g is None, n
            else:
                assert torch.allclose(g, p.grad, atol=1e-6), n
        assert abs(ret["conf"] - 0.4) < 1e-6
