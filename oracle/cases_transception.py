"""Golden-case table of the networks/Transception.py variant, shared by oracle/make_golden_transception.py (reference
side), tests/test_oracle_transception.py (oracle side) and tests/test_gpu_transception.py (CUDA side) — TEST
INFRASTRUCTURE ONLY.  Each row: name, module path inside ``Transception`` (identical in the reference and in the drop-in
mirror), a builder for the module's positional arguments, and the oracle restatement called as ``fn(sd, path, *args)``."""
import torch

from oracle import fixtures as FX
from oracle import transception_oracle as TO

MODEL_SEED = 1234


def seeded_model(num_classes=9, perturb=True, concat='original'):
    """The drop-in parameter mirror with the reference's same-seed init (+ perturbed affine / bias tensors), eval mode."""
    from transception_b200 import Transception
    torch.manual_seed(MODEL_SEED)
    net = Transception(num_classes=num_classes, concat=concat)
    if perturb:
        FX.randomise(net)
    return net.eval()


def _pe(k):
    return lambda sd, p, x: TO.patch_embed_fuse(sd, p, x, k)[0]


CASES = [
    ("pe_fuse_3x3_s2", "backbone.patch_embed2_1", lambda: (FX.rand(2, 64, 56, 56, seed=41),), _pe(3)),
    ("pe_fuse_1x1_s3", "backbone.patch_embed3_2", lambda: (FX.rand(2, 128, 28, 28, seed=42),), _pe(1)),
    ("pe_fuse_3x3_s4", "backbone.patch_embed4_1", lambda: (FX.rand(3, 320, 14, 14, seed=43),), _pe(3)),
    ("fuse_attn_s2", "backbone.block2.0.attn", lambda: (FX.rand(2, 1460, 128, seed=44),), TO.fuse_efficient_attention),
    ("fuse_attn_s3", "backbone.block3.1.attn", lambda: (FX.rand(2, 340, 320, seed=45),), TO.fuse_efficient_attention),
    ("fuse_attn_s4", "backbone.block4.1.attn", lambda: (FX.rand(3, 74, 512, seed=46),), TO.fuse_efficient_attention),
    ("fuse_attn_s2_hot", "backbone.block2.1.attn", lambda: (FX.rand(2, 1460, 128, seed=47, scale=5.0),), TO.fuse_efficient_attention),
    ("fuse_block_s2", "backbone.block2.1", lambda: (FX.rand(2, 1460, 128, seed=48), 676, 784, 26, 26, 28, 28), TO.fuse_block),
    ("fuse_block_s3", "backbone.block3.0", lambda: (FX.rand(2, 340, 320, seed=49), 144, 196, 12, 12, 14, 14), TO.fuse_block),
    ("fuse_block_s4", "backbone.block4.0", lambda: (FX.rand(3, 74, 512, seed=50), 25, 49, 5, 5, 7, 7), TO.fuse_block),
]
CASE_NAMES = [c[0] for c in CASES]
BY_NAME = {c[0]: c for c in CASES}
