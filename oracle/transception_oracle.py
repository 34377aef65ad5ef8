"""CPU oracle for the earlier model variant ``networks/Transception.py::Transception`` — TEST INFRASTRUCTURE, NOT
PRODUCT CODE.

Functional, ``state_dict``-driven restatement (plain fp32 PyTorch ops on the CPU) of the reference forward with the
default structure ``Transception(num_classes, head_count=1, dil_conv=1, token_mlp_mode="mix_skip")`` and both fusion
modes (``concat='original'`` = 1x1 conv, anything else = ``SK_Block``).  Only ``tests/`` and ``bench.py``'s CPU legs
may import it.

Pinning: ``oracle/make_golden_transception.py`` runs the REAL reference (``/root/reference/networks/Transception.py``
imported through ``oracle/ref_shim.py``) in the authoring container on seeded weights / inputs and stores its outputs in
``tests/golden/transception_golden.pt``; ``tests/test_oracle_transception.py`` re-checks this file against them.

Line references are to ``/root/reference/networks/Transception.py`` unless prefixed ``EffSegformer.py``.
"""
import torch
import torch.nn.functional as F

from oracle import mstr_oracle as O

DIMS = (64, 128, 320, 512)


# ---- OverlapPatchEmbeddings_fuse: EffSegformer.py:117-131 -------------------------------------------------------
def patch_embed_fuse(sd, p, x, k, stride=2, padding=0, dilation=2):
    """Dilated strided conv + LayerNorm -> tokens. (k, padding) = (3, 0) / (1, 0) for the two branches when dil_conv=1
    (Transception.py:383-387); returns (tokens, H, W)."""
    assert sd[p + '.proj.weight'].shape[-1] == k
    px = F.conv2d(x, sd[p + '.proj.weight'], sd[p + '.proj.bias'], stride, padding, dilation)
    H, W = px.shape[2:]
    return O.layernorm(sd, p + '.norm', O.map_to_tokens(px)), H, W


# ---- FuseEfficientAttention.forward: :49-87 (head_count = 1) -------------------------------------------------------
def fuse_efficient_attention(sd, p, x):
    """Linear k/q/v on [B, N, C] tokens, then ``.reshape(b, C, n)`` — a raw reinterpretation of the [N, C] buffer, not
    a transpose (:54-57) — softmax over n (keys) / over channels (queries), C x C context, reprojection of the
    transposed result."""
    B, N, C = x.shape
    k = O.linear(sd, p + '.keys', x).reshape(B, C, N)
    q = O.linear(sd, p + '.queries', x).reshape(B, C, N)
    v = O.linear(sd, p + '.values', x).reshape(B, C, N)
    context = F.softmax(k, dim=2) @ v.transpose(1, 2)
    att = (context.transpose(1, 2) @ F.softmax(q, dim=1)).reshape(B, C, N)
    return O.linear(sd, p + '.reprojection', att.permute(0, 2, 1))


# ---- EfficientTransformerBlockFuse.forward: :213-250 (two-branch case) -------------------------------------------
def fuse_block(sd, p, x, n1, n2, H1, W1, H2, W2):
    assert x.shape[1] == n1 + n2
    tx = x + fuse_efficient_attention(sd, p + '.attn', O.layernorm(sd, p + '.norm1', x))
    z1, z2 = tx[:, :n1], tx[:, n1:]
    m1 = z1 + O.mixffn_skip(sd, p + '.mlp1', O.layernorm(sd, p + '.norm2', z1), H1, W1)
    m2 = z2 + O.mixffn_skip(sd, p + '.mlp2', O.layernorm(sd, p + '.norm2', z2), H2, W2)
    return torch.cat((m1, m2), 1)


# ---- SK_Block.forward: :328-358 ----------------------------------------------------------------------------------
def sk_block(sd, p, maps):
    bs, c = maps[0].shape[:2]
    feats = torch.stack(maps, 0)
    S = sum(maps).mean(-1).mean(-1)
    Z = O.linear(sd, p + '.fc', S)
    w = torch.stack([O.linear(sd, '%s.fcs.%d' % (p, i), Z).view(bs, c, 1, 1) for i in range(len(maps))], 0)
    V = (torch.softmax(w, dim=0) * feats).sum(0)
    y = F.relu(O.conv(sd, p + '.conv_bn_ac.0', V))
    return O.batchnorm_eval(sd, p + '.conv_bn_ac.2', y)


# ---- one inception stage of MiT_3inception.forward: :455-483 (stage 2; 3 and 4 are the same code) ------------------
def fuse_stage(sd, p, x, stage, n_layers=2, concat='original'):
    """x: [B, C_in, H, H] map of the previous stage -> [B, C_out, H/2, H/2]."""
    side = x.shape[2] // 2
    x1, H1, W1 = patch_embed_fuse(sd, '%s.patch_embed%d_1' % (p, stage), x, 3)
    x2, H2, W2 = patch_embed_fuse(sd, '%s.patch_embed%d_2' % (p, stage), x, 1)
    n1, n2 = x1.shape[1], x2.shape[1]
    t = torch.cat((x1, x2), 1)
    for i in range(n_layers):
        t = fuse_block(sd, '%s.block%d.%d' % (p, stage, i), t, n1, n2, H1, W1, H2, W2)
    t = O.layernorm(sd, '%s.norm%d' % (p, stage), t)
    b = t.shape[0]
    m1 = t[:, :n1].reshape(b, H1, W1, -1).permute(0, 3, 1, 2)
    m2 = t[:, n1:].reshape(b, H2, W2, -1).permute(0, 3, 1, 2)
    m1 = F.interpolate(m1, [side, side])          # nearest (the default mode), :471
    if concat == 'original':
        return O.conv(sd, '%s.conv1_1_s%d' % (p, stage), torch.cat((m1, m2), 1))
    return sk_block(sd, '%s.sk_concat%d' % (p, stage), [m1, m2])


# ---- MiT_3inception.forward: :440-551 --------------------------------------------------------------------------
def mit_3inception(sd, p, x, concat='original', layers=(2, 2, 2, 2)):
    B = x.shape[0]
    t, H, W = O.patch_embed(sd, p + '.patch_embed1', x)
    for i in range(layers[0]):
        t = O.efficient_block(sd, '%s.block1.%d' % (p, i), t, H, W)
    t = O.layernorm(sd, p + '.norm1', t)
    x = t.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
    outs = [x]
    for stage in (2, 3, 4):
        x = fuse_stage(sd, p, x, stage, layers[stage - 1], concat)
        outs.append(x)
    return outs


# ---- Transception.forward: :1038-1057 --------------------------------------------------------------------------
def forward(sd, x, concat='original', return_all=False):
    if x.shape[1] == 1:
        x = x.repeat(1, 3, 1, 1)
    enc = mit_3inception(sd, 'backbone', x, concat)
    b, c = enc[3].shape[:2]
    t3 = O.decoder_layer(sd, 'decoder_3', enc[3].permute(0, 2, 3, 1).reshape(b, -1, c))
    t2 = O.decoder_layer(sd, 'decoder_2', t3, enc[2].permute(0, 2, 3, 1))
    t1 = O.decoder_layer(sd, 'decoder_1', t2, enc[1].permute(0, 2, 3, 1))
    logits = O.decoder_layer(sd, 'decoder_0', t1, enc[0].permute(0, 2, 3, 1), is_last=True)
    if return_all:
        return {'enc': enc, 'logits': logits}
    return logits
