"""sct_gan_b200 — B200-native implementation of the SCT-GAN adversarial train-step hot path.

    from sct_gan_b200 import SmartContractTransformer, SmartContractTrainer

`SmartContractTransformer` is a drop-in for the reference's class of the same name (SCT-GAN/model.py:23);
its hot path runs in hand-written sm_100a CUDA behind the C ABI of include/sct_b200.h (libsct_b200.so,
built in-tree by `python -m sct_gan_b200.build`).  There is no CPU or PyTorch-eager fallback for that path.
"""
from .model import PositionalEncoding, SmartContractTransformer  # noqa: F401
from .syntax import SoliditySyntaxRules  # noqa: F401
from .trainer import SmartContractTrainer  # noqa: F401

__all__ = ["SmartContractTransformer", "PositionalEncoding", "SmartContractTrainer", "SoliditySyntaxRules"]
