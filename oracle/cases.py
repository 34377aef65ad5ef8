"""Golden-case table shared by oracle/make_golden.py (reference side), tests/test_oracle.py (oracle side) and
tests/test_gpu_parity.py (CUDA side) — TEST INFRASTRUCTURE ONLY.

Each row: name, module path inside MSTransception (identical in the reference and in the drop-in mirror), a
builder for the module's positional arguments, and the oracle restatement called as ``fn(sd, path, *args)``.
One row (at least) per function of SURVEY.md §8a.
"""
from oracle import fixtures as FX
from oracle import mstr_oracle as O


def _pe(sd, p, x):
    return O.patch_embed(sd, p, x)[0]


def _hw(fn):
    return lambda sd, p, x, size: fn(sd, p, x, size[0], size[1])


def _enc(L):
    return lambda sd, p, x, size: O.mhca_encoder(sd, p, x, size[0], size[1], L)


def _stage(L):
    return lambda sd, p, xs: O.mhca_stage(sd, p, xs, L)


def _blayer(ch):
    return lambda sd, p, maps: O.bridge_layer(sd, p, O.bridge_tokens(maps), ch)


CASES = [
    ("patch_embed_c3", "backbone.patch_embed1", lambda: (FX.image(2, 3, seed=11),), _pe),
    ("efficient_block", "backbone.block1.0", lambda: (FX.rand(2, 3136, 64, seed=12), 56, 56), O.efficient_block),
    ("efficient_attention", "backbone.block1.1.attn", lambda: (FX.rand(2, 64, 56, 56, seed=13),), O.efficient_attention),
    ("mixffn_s1", "backbone.block1.0.mlp", lambda: (FX.rand(1, 3136, 64, seed=14), 56, 56), O.mixffn_skip),
    ("mixffn_s3", "backbone.mhca_stage3.mhca_blks.1.MHCA_layers.2.mlp", lambda: (FX.rand(3, 196, 128, seed=14), 14, 14), O.mixffn_skip),
    ("mixffn_s4", "backbone.mhca_stage4.mhca_blks.0.MHCA_layers.0.mlp", lambda: (FX.rand(3, 49, 320, seed=14), 7, 7), O.mixffn_skip),
    ("mixffn_b4", "bridge.bridge_layer2.mixffn4", lambda: (FX.rand(3, 49, 512, seed=14), 7, 7), O.mixffn_skip),
    ("ripm_s2", "backbone.patch_embed_stage2", lambda: (FX.rand(2, 64, 56, 56, seed=15),), O.patch_embed_stage),
    ("ripm_s3", "backbone.patch_embed_stage3", lambda: (FX.rand(2, 128, 28, 28, seed=15),), O.patch_embed_stage),
    ("ripm_s4", "backbone.patch_embed_stage4", lambda: (FX.rand(2, 320, 14, 14, seed=15),), O.patch_embed_stage),
    ("resblock_s2", "backbone.mhca_stage2.InvRes", lambda: (FX.rand(2, 64, 28, 28, seed=16),), O.resblock),
    ("resblock_s4", "backbone.mhca_stage4.InvRes", lambda: (FX.rand(2, 320, 7, 7, seed=16),), O.resblock),
    ("mb_attn_s2", "backbone.mhca_stage2.mhca_blks.1.MHCA_layers.0.factoratt_crpe", lambda: (FX.rand(2, 784, 64, seed=17), (28, 28)), _hw(O.factor_att)),
    ("mb_attn_s3", "backbone.mhca_stage3.mhca_blks.1.MHCA_layers.0.factoratt_crpe", lambda: (FX.rand(2, 196, 128, seed=17), (14, 14)), _hw(O.factor_att)),
    ("mb_attn_s4", "backbone.mhca_stage4.mhca_blks.1.MHCA_layers.0.factoratt_crpe", lambda: (FX.rand(2, 49, 320, seed=17), (7, 7)), _hw(O.factor_att)),
    ("mhca_block_s3", "backbone.mhca_stage3.mhca_blks.0.MHCA_layers.5", lambda: (FX.rand(2, 196, 128, seed=18), (14, 14)), _hw(O.mhca_block)),
    ("mhca_encoder_s2", "backbone.mhca_stage2.mhca_blks.2", lambda: (FX.rand(2, 784, 64, seed=18), (28, 28)), _enc(3)),
    ("mhca_encoder_s4", "backbone.mhca_stage4.mhca_blks.2", lambda: (FX.rand(2, 49, 320, seed=18), (7, 7)), _enc(3)),
    ("coordatt_s2", "backbone.mhca_stage2.aggregate", lambda: (FX.rand(2, 256, 28, 28, seed=19),), O.coord_att),
    ("coordatt_s3", "backbone.mhca_stage3.aggregate", lambda: (FX.rand(2, 512, 14, 14, seed=19),), O.coord_att),
    ("coordatt_s4", "backbone.mhca_stage4.aggregate", lambda: (FX.rand(2, 1280, 7, 7, seed=19),), O.coord_att),
    ("mhca_stage2", "backbone.mhca_stage2", lambda: ([FX.rand(2, 64, 28, 28, seed=20 + i) for i in range(3)],), _stage(3)),
    ("mhca_stage4", "backbone.mhca_stage4", lambda: ([FX.rand(2, 320, 7, 7, seed=20 + i) for i in range(3)],), _stage(3)),
    ("scale_reduce", "bridge.bridge_layer2.attn.scale_reduce", lambda: (FX.rand(2, 6076, 64, seed=31),), O.scale_reduce),
    ("bridge_self_attn", "bridge.bridge_layer3.attn", lambda: (FX.rand(2, 6076, 64, seed=32),), O.bridge_self_atten),
    ("bridge_self_attn_hot", "bridge.bridge_layer3.attn", lambda: (FX.rand(2, 6076, 64, seed=32, scale=6.0),), O.bridge_self_atten),
    ("bridge_channel_attn", "bridge.bridge_layer1.attn", lambda: (FX.rand(2, 6076, 64, seed=33),), O.bridge_channel_atten),
    ("bridge_layer1", "bridge.bridge_layer1", lambda: (FX.bridge_maps(34),), _blayer(True)),
    ("bridge_layer2", "bridge.bridge_layer2", lambda: (FX.bridge_maps(34),), _blayer(False)),
    ("bridge_block", "bridge", lambda: (FX.bridge_maps(35, bs=1),), O.bridge_block),
]
CASE_NAMES = [c[0] for c in CASES]
BY_NAME = {c[0]: c for c in CASES}


def flatten_out(out):
    import torch
    if isinstance(out, torch.Tensor):
        return [out]
    return [o for o in out if isinstance(o, torch.Tensor)]
