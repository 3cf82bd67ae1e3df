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


# ------------------------------------------------------------------------------------------------
# The loss block's data-dependent branches under data parallelism (trainer.compute_losses): the floors / > 1 / > 5
# rescale of the line loss (train.py:1185-1194), the "batch held a vulnerable line" focal-loss switch
# (train.py:1174-1184) and the spatial penalty's cross-sample line sums (train.py:174-245, live at S == 1024) must
# take the decision / value of the single-process step on the concatenated batch.
def _loss_inputs():
    g = torch.Generator().manual_seed(5)
    B, S, C = 4, 1024, 8
    ll = torch.randn(B, S, C, generator=g) * 0.5
    ll[:2] -= 6.0   # shard 0: confident negatives (tiny local loss, no vulnerable line)
    ll[2:] += 3.0   # shard 1: wrong everywhere (large local loss)
    cl = torch.randn(B, C, generator=g)
    vl = torch.zeros(B, S, C)
    vl[2:, 5:40:7, 1] = 1.0
    vl[3, 100, 3] = 1.0
    batch = {"contract_vulnerabilities": (torch.rand(B, C, generator=g) < 0.3).float(), "vulnerable_lines": vl,
             "token_to_line": (torch.arange(S) // 12)[None, :].expand(B, S).contiguous()}
    return ll, cl, batch


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.zeros(3))


def _loss_grads(ll, cl, batch, pg_world):
    from sct_gan_b200.trainer import SmartContractTrainer

    tr = SmartContractTrainer(_Tiny(), use_gan=False, use_augmentation=False, fused_optimizer=False)
    tr.focal.copy_(torch.tensor([5.0, 2.0, 0.2]))  # large alpha: the global line loss lands above the 5.0 threshold
    ll = ll.clone().requires_grad_(True)
    cl = cl.clone().requires_grad_(True)
    gen = (ll.mean() * 0.0 + 1.5)
    out = {"gen_ce_loss": gen, "contract_vulnerability_logits": cl, "line_vulnerability_logits": ll,
           "discriminator_logits": None}
    res = tr.compute_losses(out, batch, 0.0, n_lines=86)
    res["total_loss"].backward()
    return ll.grad, cl.grad, bool(tr._pending_has_line), float(res["line_vuln_loss"])


def _loss_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ll, cl, batch = _loss_inputs()
    sh = slice(rank * 2, rank * 2 + 2)
    g_ll, g_cl, has, lv = _loss_grads(ll[sh], cl[sh], {k: v[sh] for k, v in batch.items()}, world)
    ret[rank] = (g_ll, g_cl, has, lv)
    dist.barrier()
    dist.destroy_process_group()


def test_loss_block_branches_match_single_process_under_data_parallel():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loss_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    ll, cl, batch = _loss_inputs()
    g_ll, g_cl, has, lv = _loss_grads(ll, cl, batch, 1)
    assert has and lv > 0.5  # the >5 branch fired on the global value (6.9.. * 0.1)
    for r in range(world):
        sh = slice(r * 2, r * 2 + 2)
        d_ll, d_cl, d_has, _ = ret[r]
        assert d_has == has  # rank 0's shard holds no vulnerable line, the batch does
        # averaged over ranks, the data-parallel gradient of a shard = single-process gradient of that shard
        assert torch.allclose(d_ll / world, g_ll[sh], atol=1e-7, rtol=1e-4), r
        assert torch.allclose(d_cl / world, g_cl[sh], atol=1e-7, rtol=1e-4), r
