"""CPU oracle for the TransCeption hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, ``state_dict``-driven restatement (plain fp32/fp64 PyTorch ops on the CPU) of
the reference forward ``networks/MSTr.py::MSTransception`` with default structure.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it; the product path (``transception_b200``) never does.

Pinning: the reference has no golden vectors or tests of its own.  This
restatement is pinned against *outputs of the reference itself run in the authoring
container* (``oracle/make_golden.py`` imports ``/root/reference`` through
``oracle/ref_shim.py`` and writes ``tests/golden/*.pt``): whole-model logits and every
encoder/bridge map on a seeded input, plus per-module fixtures with seeded weights.
``tests/test_oracle.py`` re-checks the oracle against those fixtures on every run.

Every function takes ``sd`` (a reference-format state_dict, or any mapping of tensors), a key
``prefix`` and activations, and cites the reference lines it restates
(``/root/reference/networks/MSTr.py`` unless stated otherwise).
"""
import math
import torch
import torch.nn.functional as F

TOK = (3136, 4704, 5684, 6076)  # bridge token offsets at 224x224 (MSTr.py:2228-2231)


def bridge_geometry(ntok):
    """Stage-1 side S and the cumulative token offsets of the 4-scale pyramid whose 64-wide token count is ``ntok``.

    The reference hard-codes the 224x224 values (S = 56, offsets 3136/4704/5684/6076: MSTr.py:2228-2231, :2394-2397,
    :2432-2435) and therefore fails on any other input size.  This is the generalisation SURVEY.md section 8c asks for
    (56/28/14/7 -> S, S/2, S/4, S/8); at ntok = 6076 it returns exactly the reference constants, which
    tests/test_oracle.py pins against the live reference's outputs."""
    s2 = ntok * 16 // 31
    S = int(math.isqrt(s2))
    if S * S * 31 != ntok * 16 or S % 8:
        raise ValueError("token count %d is not a 4-scale pyramid" % ntok)
    sides = (S, S // 2, S // 4, S // 8)
    offs, acc = [], 0
    for hw, mult in zip(sides, (1, 2, 5, 8)):
        acc += hw * hw * mult
        offs.append(acc)
    return S, sides, tuple(offs)


def _p(sd, key):
    return sd[key]


def linear(sd, p, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def layernorm(sd, p, x, eps=1e-5):
    w = sd[p + '.weight']
    return F.layer_norm(x, (w.shape[0],), w, sd[p + '.bias'], eps)


# Training-row tests set this: nn.BatchNorm2d in train mode normalises with the statistics of the batch
# (F.batch_norm(training=True), what the reference's 21 BatchNorm2d layers do under model.train(), trainer.py:113);
# the running statistics are not touched here.  Pinned against the live reference by tests/golden/train_golden.pt.
BN_TRAIN = False


def batchnorm_eval(sd, p, x, eps=1e-5):
    """nn.BatchNorm2d: eval mode (running statistics) unless BN_TRAIN."""
    if BN_TRAIN:
        return F.batch_norm(x, None, None, sd[p + '.weight'], sd[p + '.bias'], True, 0.1, eps)
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'],
                        False, 0.1, eps)


def conv(sd, p, x, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + '.weight'], sd.get(p + '.bias'), stride, padding, 1, groups)


def tokens_to_map(x, H, W):
    B, N, C = x.shape
    return x.transpose(1, 2).reshape(B, C, H, W)


def map_to_tokens(x):
    return x.flatten(2).transpose(1, 2)


# ---- Mix-FFN: MSTr.py:21-31 (DWConv), :58-61 (MixFFN_skip.forward) ------------------
def mixffn_skip(sd, p, x, H, W):
    h = linear(sd, p + '.fc1', x)
    C4 = h.shape[-1]
    dw = map_to_tokens(conv(sd, p + '.dwconv.dwconv', tokens_to_map(h, H, W), 1, 1, C4))
    ax = F.gelu(layernorm(sd, p + '.norm1', dw + h))  # fc1(x) evaluated once; algebraically identical
    return linear(sd, p + '.fc2', ax)


# ---- efficient attention: MSTr.py:106-143 --------------------------------------------
def efficient_attention(sd, p, x_map):
    n, C, h, w = x_map.shape
    keys = conv(sd, p + '.keys', x_map).reshape(n, C, h * w)
    queries = conv(sd, p + '.queries', x_map).reshape(n, C, h * w)
    values = conv(sd, p + '.values', x_map).reshape(n, C, h * w)
    key = F.softmax(keys, dim=2)
    query = F.softmax(queries, dim=1)
    context = key @ values.transpose(1, 2)
    att = (context.transpose(1, 2) @ query).reshape(n, C, h, w)
    return conv(sd, p + '.reprojection', att)


# ---- EfficientTransformerBlock.forward: MSTr.py:164-173 ------------------------------
def efficient_block(sd, p, x, H, W):
    n1 = tokens_to_map(layernorm(sd, p + '.norm1', x), H, W)
    tx = x + map_to_tokens(efficient_attention(sd, p + '.attn', n1))
    return tx + mixffn_skip(sd, p + '.mlp', layernorm(sd, p + '.norm2', tx), H, W)


# ---- stem: MSTr.py:299-304 -----------------------------------------------------------
def patch_embed(sd, p, x, stride=4, padding=3):
    px = conv(sd, p + '.proj', x, stride, padding)
    H, W = px.shape[2:]
    return layernorm(sd, p + '.norm', map_to_tokens(px)), H, W


# ---- RIPM: MSTr.py:355-362 (DWConv2d_BN), :725-732 (Patch_Embed_stage) ----------------
def dwsep_bn_hs(sd, p, x, stride):
    C = x.shape[1]
    y = conv(sd, p + '.dwconv', x, stride, 1, C)
    y = conv(sd, p + '.pwconv', y)
    return F.hardswish(batchnorm_eval(sd, p + '.bn', y))


def patch_embed_stage(sd, p, x, n_path=3):
    outs = []
    for i in range(n_path):
        x = dwsep_bn_hs(sd, '%s.patch_embeds.%d.patch_conv' % (p, i), x, 2 if i == 0 else 1)
        outs.append(x)
    return outs


# ---- ResBlock: MSTr.py:1042-1050, Conv2d_BN :399-404 ----------------------------------
def resblock(sd, p, x):
    C = x.shape[1]
    f = F.hardswish(batchnorm_eval(sd, p + '.conv1.bn', conv(sd, p + '.conv1.conv', x)))
    f = F.hardswish(batchnorm_eval(sd, p + '.norm', conv(sd, p + '.dwconv', f, 1, 1, C)))
    f = batchnorm_eval(sd, p + '.conv2.bn', conv(sd, p + '.conv2.conv', f))
    return x + f


# ---- MB transformer: MSTr.py:744-752 (CPE), :801-823 (CRPE), :852-886 (FactorAtt), :935-946 (block) ----
def conv_pos_enc(sd, p, x, H, W):
    feat = tokens_to_map(x, H, W)
    return map_to_tokens(conv(sd, p + '.proj', feat, 1, 1, feat.shape[1]) + feat)


def conv_rel_pos_enc(sd, p, q, v, H, W, splits=(2, 3, 3), windows=(3, 5, 7)):
    B, h, N, Ch = q.shape
    v_img = v.transpose(2, 3).reshape(B, h * Ch, H, W)  # 'B h (H W) Ch -> B (h Ch) H W'
    outs, c0 = [], 0
    for i, (hs, win) in enumerate(zip(splits, windows)):
        c1 = c0 + hs * Ch
        outs.append(conv(sd, '%s.conv_list.%d' % (p, i), v_img[:, c0:c1], 1, win // 2, hs * Ch))
        c0 = c1
    cv = torch.cat(outs, 1).reshape(B, h, Ch, H * W).transpose(2, 3)
    return q * cv


def factor_att(sd, p, x, H, W, heads=8):
    B, N, C = x.shape
    Ch = C // heads
    qkv = linear(sd, p + '.qkv', x).reshape(B, N, 3, heads, Ch).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    ctx = torch.einsum('bhnk,bhnv->bhkv', k.softmax(dim=2), v)
    fa = torch.einsum('bhnk,bhkv->bhnv', q, ctx)
    out = (Ch ** -0.5) * fa + conv_rel_pos_enc(sd, p + '.crpe', q, v, H, W)
    return linear(sd, p + '.proj', out.transpose(1, 2).reshape(B, N, C))


def mhca_block(sd, p, x, H, W, heads=8):
    x = conv_pos_enc(sd, p + '.cpe', x, H, W)
    x = x + factor_att(sd, p + '.factoratt_crpe', layernorm(sd, p + '.norm1', x, 1e-6), H, W, heads)
    return x + mixffn_skip(sd, p + '.mlp', layernorm(sd, p + '.norm2', x, 1e-6), H, W)


def mhca_encoder(sd, p, x, H, W, n_layers, heads=8):
    B = x.shape[0]
    for i in range(n_layers):
        x = mhca_block(sd, '%s.MHCA_layers.%d' % (p, i), x, H, W, heads)
    return x.reshape(B, H, W, -1).permute(0, 3, 1, 2)


# ---- IFF: MSTr.py:1270-1286 (silu_swish), :1322-1348 (CoordAtt.forward) ---------------
def silu_swish(t):
    return t * torch.minimum(F.silu(t + 3) / 6, torch.ones_like(t))


def coord_att(sd, p, x):
    n, c, h, w = x.shape
    x_h = x.mean(dim=3, keepdim=True)                      # [n,c,h,1]
    x_w = x.mean(dim=2, keepdim=True).permute(0, 1, 3, 2)  # [n,c,w,1]
    y = conv(sd, p + '.conv1', torch.cat([x_h, x_w], dim=2))
    y = silu_swish(batchnorm_eval(sd, p + '.bn1', y))
    y_h, y_w = torch.split(y, [h, w], dim=2)
    a_h = conv(sd, p + '.conv_h', y_h).sigmoid()
    a_w = conv(sd, p + '.conv_w', y_w.permute(0, 1, 3, 2)).sigmoid()
    return conv(sd, p + '.conv_in_out', x * a_w * a_h)


def mhca_stage(sd, p, inputs, n_layers, heads=8):
    outs = [resblock(sd, p + '.InvRes', inputs[0])]
    for i, x in enumerate(inputs):
        H, W = x.shape[2:]
        outs.append(mhca_encoder(sd, '%s.mhca_blks.%d' % (p, i), map_to_tokens(x), H, W, n_layers, heads))
    return coord_att(sd, p + '.aggregate', torch.cat(outs, dim=1))


# ---- encoder: MSTr.py:1709-1744 --------------------------------------------------------
def msvit(sd, p, x):
    B = x.shape[0]
    t, H, W = patch_embed(sd, p + '.patch_embed1', x)
    for i in range(2):
        t = efficient_block(sd, '%s.block1.%d' % (p, i), t, H, W)
    t = layernorm(sd, p + '.norm1', t)
    cur = t.reshape(B, H, W, -1).permute(0, 3, 1, 2)
    outs = [cur]
    for s, L in ((2, 3), (3, 8), (4, 3)):
        paths = patch_embed_stage(sd, '%s.patch_embed_stage%d' % (p, s), cur)
        cur = mhca_stage(sd, '%s.mhca_stage%d' % (p, s), paths, L)
        outs.append(cur)
    return outs


# ---- bridge: MSTr.py:2225-2249, :2267-2292, :2309-2353, :2373-2409, :2422-2442 --------
def scale_reduce(sd, p, x):
    B, N, C = x.shape
    _, (h0, h1, h2, _h3), tok = bridge_geometry(N)
    t0 = x[:, :tok[0]].reshape(B, h0, h0, C).permute(0, 3, 1, 2)
    t1 = x[:, tok[0]:tok[1]].reshape(B, h1, h1, C * 2).permute(0, 3, 1, 2)
    t2 = x[:, tok[1]:tok[2]].reshape(B, h2, h2, C * 5).permute(0, 3, 1, 2)
    t3 = x[:, tok[2]:tok[3]]
    s0 = conv(sd, p + '.sr0', t0, 8).reshape(B, C, -1).permute(0, 2, 1)
    s1 = conv(sd, p + '.sr1', t1, 4).reshape(B, C, -1).permute(0, 2, 1)
    s2 = conv(sd, p + '.sr2', t2, 2).reshape(B, C, -1).permute(0, 2, 1)
    return layernorm(sd, p + '.norm', torch.cat([s0, s1, s2, t3], -2))


def bridge_self_atten(sd, p, x):
    B, N, C = x.shape
    q = linear(sd, p + '.q', x)
    r = scale_reduce(sd, p + '.scale_reduce', x)
    kv = linear(sd, p + '.kv', r).reshape(B, -1, 2, C)
    k, v = kv[:, :, 0], kv[:, :, 1]
    attn = ((q @ k.transpose(-2, -1)) * (C ** -0.5)).softmax(dim=-1)
    return linear(sd, p + '.proj', attn @ v)


def bridge_channel_atten(sd, p, x):
    B, N, C = x.shape
    k = linear(sd, p + '.k', x).reshape(B, C, N)  # raw reinterpretation, not a transpose (MSTr.py:2312-2314)
    q = linear(sd, p + '.q', x).reshape(B, C, N)
    v = linear(sd, p + '.v', x).reshape(B, C, N)
    context = F.softmax(k, dim=2) @ v.transpose(1, 2)
    att = context.transpose(1, 2) @ F.softmax(q, dim=1)
    return linear(sd, p + '.proj', att.permute(0, 2, 1))


def bridge_tokens(maps):
    B = maps[0].shape[0]
    return torch.cat([m.permute(0, 2, 3, 1).reshape(B, -1, 64) for m in maps], -2)


def bridge_layer(sd, p, x, ch_att):
    B, _, C = x.shape
    attn = bridge_channel_atten if ch_att else bridge_self_atten
    tx1 = x + attn(sd, p + '.attn', layernorm(sd, p + '.norm1', x))
    tx = layernorm(sd, p + '.norm2', tx1)
    parts, off = [], 0
    sides = bridge_geometry(x.shape[1])[1]
    for i, (hw, mult) in enumerate(zip(sides, (1, 2, 5, 8))):
        n = hw * hw * mult
        t = tx[:, off:off + n].reshape(B, -1, C * mult)
        parts.append(mixffn_skip(sd, '%s.mixffn%d' % (p, i + 1), t, hw, hw).reshape(B, -1, C))
        off += n
    return tx1 + torch.cat(parts, -2)


def bridge_block(sd, p, maps, ch_att_list=(True, False, False, False)):
    x = bridge_tokens(maps)
    for i in range(4):
        x = bridge_layer(sd, '%s.bridge_layer%d' % (p, i + 1), x, ch_att_list[i])
    B, _, C = x.shape
    outs, off = [], 0
    for hw, mult in zip(bridge_geometry(x.shape[1])[1], (1, 2, 5, 8)):
        n = hw * hw * mult
        outs.append(x[:, off:off + n].reshape(B, hw, hw, C * mult).permute(0, 3, 1, 2))
        off += n
    return outs


# ---- decoder: MSTr.py:184-201, :212-227, :273-290 --------------------------------------
def patch_expand(sd, p, x, H, W, scale):
    B = x.shape[0]
    x = F.linear(x, sd[p + '.expand.weight'])
    C = x.shape[-1]
    c = C // (scale * scale)
    x = x.view(B, H, W, scale, scale, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * scale * W * scale, c)
    return layernorm(sd, p + '.norm', x)


def decoder_layer(sd, p, x1, x2=None, is_last=False):
    if x2 is None:
        side = int(math.isqrt(x1.shape[1]))
        return patch_expand(sd, p + '.layer_up', x1, side, side, 2)
    b, h, w, c = x2.shape
    t = linear(sd, p + '.concat_linear', torch.cat([x1, x2.reshape(b, -1, c)], dim=-1))
    t = efficient_block(sd, p + '.layer_former_1', t, h, w)
    t = efficient_block(sd, p + '.layer_former_2', t, h, w)
    if is_last:
        up = patch_expand(sd, p + '.layer_up', t, h, w, 4)
        return conv(sd, p + '.last_layer', up.view(b, 4 * h, 4 * w, -1).permute(0, 3, 1, 2))
    return patch_expand(sd, p + '.layer_up', t, h, w, 2)


# ---- top level: MSTr.py:2826-2852 ------------------------------------------------------
def forward(sd, x, ch_att_list=(True, False, False, False), return_all=False):
    """sd: reference-format state_dict; x: [B,1|3,H,H] with H = 224 in the reference (any multiple of 32 here, see
    bridge_geometry). Returns logits (and, optionally, every map)."""
    if x.shape[1] == 1:
        x = x.repeat(1, 3, 1, 1)
    enc = msvit(sd, 'backbone', x)
    br = bridge_block(sd, 'bridge', enc, ch_att_list)
    b, c = br[3].shape[:2]
    t3 = decoder_layer(sd, 'decoder_3', br[3].permute(0, 2, 3, 1).reshape(b, -1, c))
    t2 = decoder_layer(sd, 'decoder_2', t3, br[2].permute(0, 2, 3, 1))
    t1 = decoder_layer(sd, 'decoder_1', t2, br[1].permute(0, 2, 3, 1))
    logits = decoder_layer(sd, 'decoder_0', t1, br[0].permute(0, 2, 3, 1), is_last=True)
    if return_all:
        return {'enc': enc, 'bridge': br, 'logits': logits}
    return logits


def strip(sd, prefix):
    """View of ``sd`` with ``prefix.`` removed from the keys that carry it (for per-module use)."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix + '.')}
