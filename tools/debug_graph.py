import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sct_oracle as O
from sct_gan_b200 import SmartContractTrainer, SmartContractTransformer, ops

cfg = {**O.DEFAULT_CFG, **dict(num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=256, max_length=128, vocab_size=512, dropout=0.3)}
batch = O.make_batch(2, 64, 32, 512, seed=3, device="cuda")
n_lines = int(batch["token_to_line"].max()) + 1

def run(use_graph, lr, heads, steps=5):
    torch.manual_seed(5)
    m = SmartContractTransformer(**cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(O.synth_state_dict(shapes, 3))
    m = m.cuda()
    tr = SmartContractTrainer(m, learning_rate=lr, use_augmentation=True, use_gan=True, use_cuda_graph=use_graph, compute_vuln_heads=heads)
    out = []
    for _ in range(steps):
        res = tr.train_step(batch, n_lines=n_lines)
        ep = m._drop_epoch[0].item()
        out.append((round(res["gen_loss"].item(), 5), round(res["total_loss"].item(), 5), ep))
    return out

for heads in (False, True):
    for lr in (0.0, 1e-4):
        print("heads", heads, "lr", lr)
        print("  eager", run(False, lr, heads))
        print("  graph", run(True, lr, heads))
