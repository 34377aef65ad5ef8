"""Drop-in ``MSTransception`` for the reference's ``networks.MSTr``.

Boundary: the ``nn.Module`` surface is the reference's plugin interface —
``from networks.MSTr import MSTransception`` (reference ``train_MSTransception.py:12``,
``test.py:14``).  Every class below keeps the reference's constructor signature, the
attribute names (=> identical ``state_dict`` keys, 2 200 of them, dead and aliased ones
included) and the reference's construction order (=> identical same-seed initialisation,
reference ``MSTr.py:2759-2823``).  The modules only *hold* parameters; all arithmetic is
done by hand-written sm_100a kernels reached through the C-ABI library
(``transception_b200.ops`` -> ``libtransception_sm100.so``).  There is no PyTorch/CPU
fallback: calling ``forward`` without the CUDA library raises.

Layout: activations are tokens-major / NHWC fp32 end to end.  Wherever the reference's
sub-module boundary is an NCHW map, this file returns an NCHW-*shaped* view of NHWC
storage (torch ``channels_last`` strides): same shapes and values as the reference, no
transposition kernels.

Only the default structure of the entry scripts is built (``Stage_3or4=3``,
``concat='coord'``, ``have_bridge='original'``; any ``br_ch_att_list`` permutation); other
structural values raise ``NotImplementedError``.
"""
import torch
import torch.nn as nn

from . import autograd as tcx_autograd
from . import ops

_LN_EPS = 1e-5


_SIDE = {}


def _side_stream(device):
    """One auxiliary stream per device for module-level parallel branches (streams, not a tracing compiler)."""
    key = torch.device(device).index
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device=device)
    return st


TRAIN_BRANCH_STREAMS = True    # training row: independent branches of a stage run on their own streams (set False for A/B)
_BRANCH = {}


def _parallel_branches(device, fns, inputs=()):
    """Run independent branch closures on per-device auxiliary streams, forked from and joined back to the current stream
    (events only: capturable as parallel branches of a CUDA graph).  Autograd runs each node's backward on the stream its forward
    ran on, so the backward of the branches overlaps the same way.  ``inputs``: tensors allocated on the current stream that the
    branches read (and keep for backward) — recorded on the branch streams so the caching allocator does not hand their memory
    out again while a branch kernel may still be reading it; branch outputs are recorded on the current stream likewise."""
    if not TRAIN_BRANCH_STREAMS or len(fns) < 2:
        return [f() for f in fns]
    key = torch.device(device).index
    pool = _BRANCH.setdefault(key, [])
    while len(pool) < len(fns):
        pool.append(torch.cuda.Stream(device=device))
    main = torch.cuda.current_stream(device)
    outs = []
    for st, f in zip(pool, fns):
        st.wait_stream(main)
        for t in inputs:
            t.record_stream(st)
        with torch.cuda.stream(st):
            outs.append(f())
    for st, o in zip(pool, outs):
        main.wait_stream(st)
        o.record_stream(main)
    return outs


# Parameters shared by the blocks of an MHCAEncoder (its ConvPosEnc / ConvRelPosEnc): in the training row every block takes its
# own alias from a fan-out node (autograd.FanOutFn), so the per-block gradients are added once, by one kernel.  Per thread:
# nn.DataParallel runs its replicas in threads.
_FANOUT = __import__("threading").local()


def _shared(p):
    """The alias reserved for this use of a shared parameter, or the parameter itself outside an MHCAEncoder training forward."""
    d = getattr(_FANOUT, "d", None)
    if d:
        pool = d.get(id(p))
        if pool:
            return pool.pop()
    return p


def _nhwc(x):
    """NCHW-shaped tensor (any strides) -> contiguous [B,H,W,C] (free for channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()


def _as_nchw(x_nhwc):
    """[B,H,W,C] storage -> NCHW-shaped view (channels_last strides)."""
    return x_nhwc.permute(0, 3, 1, 2)


def _xavier_convs(mods):
    for m in mods:
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)


def _recording(x, w):
    """True when autograd is recording this call: the training-row nodes (autograd.py) run instead of the fused forward."""
    return torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)


def _bn(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


def _check_eval_bn(mod, x=None):
    """Eval-mode BatchNorm path: fine without autograd; gradients through running statistics are not built."""
    if x is not None and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters())):
        raise NotImplementedError(
            "transception_b200: backward through eval-mode BatchNorm (running statistics) is not built; "
            "call .train() for a training step or wrap inference in torch.no_grad()")


def _lin_nhwc(x, conv):
    """1x1 convolution on an NHWC map as a Linear autograd node."""
    w = conv.weight
    return tcx_autograd.linear(x, w.reshape(w.shape[0], w.shape[1]), conv.bias)


# --------------------------------------------------------------------------------------
# Mix-FFN (reference MSTr.py:21-31 DWConv, :48-61 / :889-902 MixFFN_skip)
# --------------------------------------------------------------------------------------
class DWConv(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, groups=dim)

    def forward(self, x, H, W):
        B, N, C = x.shape
        if _recording(x, self.dwconv.weight):
            return tcx_autograd.dwconv_tokens(x, H, W, self.dwconv.weight, self.dwconv.bias, False)
        return ops.dwconv_tokens(x.contiguous(), H, W, self.dwconv.weight, self.dwconv.bias, add_input=False)


class MixFFN_skip(nn.Module):
    """fc2(GELU(LN(dw3x3(fc1 x) + fc1 x))); ``norm2``/``norm3`` are dead parameters kept for
    state_dict compatibility (reference MSTr.py:56-57)."""

    def __init__(self, c1, c2):
        super().__init__()
        self.fc1 = nn.Linear(c1, c2)
        self.dwconv = DWConv(c2)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(c2, c1)
        self.norm1 = nn.LayerNorm(c2)
        self.norm2 = nn.LayerNorm(c2)
        self.norm3 = nn.LayerNorm(c2)

    def args(self):
        return (self.fc1.weight, self.fc1.bias, self.dwconv.dwconv.weight, self.dwconv.dwconv.bias,
                self.norm1.weight, self.norm1.bias, self.norm1.eps, self.fc2.weight, self.fc2.bias)

    def forward(self, x, H, W, residual=None):
        """``residual`` (extension, optional): added to the result inside the fc2 kernel — the caller's skip connection."""
        if _recording(x, self.fc1.weight):
            # training row: forward + backward on the library's kernels (transception_b200/autograd.py)
            return tcx_autograd.mixffn_skip(x, H, W, *self.args(), residual=residual)
        return ops.mixffn_skip(x, H, W, *self.args(), residual=residual)


# --------------------------------------------------------------------------------------
# Stage-1 / decoder efficient attention (reference MSTr.py:80-173)
# --------------------------------------------------------------------------------------
class EfficientAttention(nn.Module):
    def __init__(self, in_channels, key_channels, value_channels, head_count=1):
        super().__init__()
        if head_count != 1 or key_channels != in_channels or value_channels != in_channels:
            raise NotImplementedError("EfficientAttention is built for head_count=1, key=value=in channels "
                                      "(the only configuration reachable from MSTransception)")
        self.in_channels, self.key_channels = in_channels, key_channels
        self.head_count, self.value_channels = head_count, value_channels
        self.keys = nn.Conv2d(in_channels, key_channels, 1)
        self.queries = nn.Conv2d(in_channels, key_channels, 1)
        self.values = nn.Conv2d(in_channels, value_channels, 1)
        self.reprojection = nn.Conv2d(value_channels, in_channels, 1)

    def args(self):
        return (self.keys.weight, self.keys.bias, self.queries.weight, self.queries.bias,
                self.values.weight, self.values.bias, self.reprojection.weight, self.reprojection.bias)

    def forward(self, input_):
        x = _nhwc(input_)
        B, H, W, C = x.shape
        if _recording(input_, self.keys.weight):
            y = tcx_autograd.eff_attn(x.reshape(B, H * W, C), *self.args())
        else:
            y = ops.eff_attn(x.view(B, H * W, C), *self.args(), residual=None, reinterpret=False)
        return _as_nchw(y.view(B, H, W, C))


class EfficientTransformerBlock(nn.Module):
    """``head_count`` is accepted and ignored exactly like the reference (MSTr.py:154-155)."""

    def __init__(self, in_dim, key_dim, value_dim, head_count=1, token_mlp='mix'):
        super().__init__()
        if token_mlp != 'mix_skip':
            raise NotImplementedError("only token_mlp='mix_skip' is built")
        self.norm1 = nn.LayerNorm(in_dim)
        self.attn = EfficientAttention(in_channels=in_dim, key_channels=key_dim,
                                       value_channels=value_dim, head_count=1)
        self.norm2 = nn.LayerNorm(in_dim)
        self.mlp = MixFFN_skip(in_dim, int(in_dim * 4))

    def forward(self, x, H, W):
        if _recording(x, self.norm1.weight):
            # training row: every op is an autograd node backed by the library's forward + backward kernels; the two
            # residual additions are the only ATen arithmetic (MSTr.py:164-173)
            n1, n2 = self.norm1, self.norm2
            xn, xr = tcx_autograd.layernorm_res(x, n1.weight, n1.bias, n1.eps)
            tx = tcx_autograd.eff_attn(xn, *self.attn.args(), residual=xr)          # x + attn(LN(x)): the add is the kernel's epilogue
            xn, xr = tcx_autograd.layernorm_res(tx, n2.weight, n2.bias, n2.eps)
            return self.mlp(xn, H, W, residual=xr)
        return ops.eff_block(x, H, W, self.norm1.weight, self.norm1.bias, self.norm1.eps, self.attn.args(),
                             self.norm2.weight, self.norm2.bias, self.mlp.args())


# --------------------------------------------------------------------------------------
# Decoder (reference MSTr.py:176-290) — SURVEY §8f rank 1
# --------------------------------------------------------------------------------------
def _patch_expand_train(x, H, W, w, scale, norm):
    """PatchExpand / FinalPatchExpand_X4 (MSTr.py:184-201, :212-227) as autograd nodes: expand GEMM and LayerNorm on the
    library's kernels, the pixel shuffle between them is a permuted copy."""
    B = x.shape[0]
    y = tcx_autograd.linear(x, w)
    c = y.shape[-1] // (scale * scale)
    y = y.view(B, H, W, scale, scale, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * scale * W * scale, c)
    return tcx_autograd.layernorm(y, norm.weight, norm.bias, norm.eps)


class PatchExpand(nn.Module):
    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False) if dim_scale == 2 else nn.Identity()
        self.norm = norm_layer(dim // dim_scale)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        if _recording(x, self.expand.weight):
            return _patch_expand_train(x, H, W, self.expand.weight, 2, self.norm)
        return ops.patch_expand(x.contiguous(), H, W, self.expand.weight, 2,
                                self.norm.weight, self.norm.bias, self.norm.eps)


class FinalPatchExpand_X4(nn.Module):
    def __init__(self, input_resolution, dim, dim_scale=4, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.dim_scale = dim_scale
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        if _recording(x, self.expand.weight):
            return _patch_expand_train(x, H, W, self.expand.weight, self.dim_scale, self.norm)
        return ops.patch_expand(x.contiguous(), H, W, self.expand.weight, self.dim_scale,
                                self.norm.weight, self.norm.bias, self.norm.eps)


class MyDecoderLayer(nn.Module):
    def __init__(self, input_size, in_out_chan, head_count, token_mlp_mode, n_class=9,
                 norm_layer=nn.LayerNorm, is_last=False):
        super().__init__()
        dims, out_dim, key_dim, value_dim = in_out_chan
        if not is_last:
            self.concat_linear = nn.Linear(dims * 2, out_dim)
            self.layer_up = PatchExpand(input_resolution=input_size, dim=out_dim, dim_scale=2, norm_layer=norm_layer)
            self.last_layer = None
        else:
            self.concat_linear = nn.Linear(dims * 4, out_dim)
            self.layer_up = FinalPatchExpand_X4(input_resolution=input_size, dim=out_dim, dim_scale=4,
                                                norm_layer=norm_layer)
            self.last_layer = nn.Conv2d(out_dim, n_class, 1)
        self.layer_former_1 = EfficientTransformerBlock(out_dim, key_dim, value_dim, head_count, token_mlp_mode)
        self.layer_former_2 = EfficientTransformerBlock(out_dim, key_dim, value_dim, head_count, token_mlp_mode)
        # reference MSTr.py:255-269 (pre-order walk: xavier on every Linear/Conv2d weight, zero biases)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x1, x2=None):
        if x2 is None:
            return self.layer_up(x1)
        b, h, w, c = x2.shape
        if _recording(x1, self.concat_linear.weight):
            return self._forward_train(x1, x2.reshape(b, h * w, c), h, w)
        cat_linear_x = ops.concat_linear(x1, x2.reshape(b, h * w, c), self.concat_linear.weight, self.concat_linear.bias)
        t1 = self.layer_former_1(cat_linear_x, h, w)
        t2 = self.layer_former_2(t1, h, w)
        if self.last_layer is not None:
            up = self.layer_up
            # fused: expand GEMM -> pixel shuffle x4 -> LN(64) -> 1x1 conv to classes, NCHW logits
            return ops.final_expand_head(t2, h, w, up.expand.weight, up.norm.weight, up.norm.bias, up.norm.eps,
                                         self.last_layer.weight, self.last_layer.bias)
        return self.layer_up(t2)


    def _forward_train(self, x1, x2, h, w):
        """Training row (MSTr.py:273-290): Linear / LayerNorm / block nodes of autograd.py; the concatenation, the pixel
        shuffles and the NCHW view of the logits are copies."""
        t = tcx_autograd.linear(torch.cat([x1, x2], dim=-1), self.concat_linear.weight, self.concat_linear.bias)
        t = self.layer_former_2(self.layer_former_1(t, h, w), h, w)
        if self.last_layer is None:
            return self.layer_up(t)
        # expand Linear node, then pixel shuffle + LayerNorm + 1x1 conv to n_class planes as ONE node writing NCHW logits
        up, cw, cb = self.layer_up, self.last_layer.weight, self.last_layer.bias
        if t.shape[-1] != 64 or cw.shape[0] > 16:
            raise NotImplementedError("the training class head is built for dim 64 and at most 16 classes")
        e = tcx_autograd.linear(t, up.expand.weight)
        return tcx_autograd.final_head(e, h, w, up.norm.weight, up.norm.bias, up.norm.eps, cw, cb)


# --------------------------------------------------------------------------------------
# Stage-1 stem (reference MSTr.py:292-304)
# --------------------------------------------------------------------------------------
class OverlapPatchEmbeddings(nn.Module):
    def __init__(self, img_size=224, patch_size=7, stride=4, padding=1, in_ch=3, dim=768):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_ch, dim, patch_size, stride, padding)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        k, s, p = self.proj.kernel_size[0], self.proj.stride[0], self.proj.padding[0]
        H = (x.shape[2] + 2 * p - k) // s + 1
        W = (x.shape[3] + 2 * p - k) // s + 1
        if _recording(x, self.proj.weight):
            # training row (MSTr.py:299-304): conv node (fused stem kernel; the patch matrix is built only for the weight
            # gradient) + LayerNorm node
            w = self.proj.weight
            if tuple(w.shape) == (64, 3, 7, 7) and s == 4 and p == 3 and x.shape[1] in (1, 3) and not x.requires_grad:
                y = tcx_autograd.patch_embed_conv(x, w, self.proj.bias)
                return tcx_autograd.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps), H, W
            if x.shape[1] == 1 and w.shape[1] == 3:
                x = x.repeat(1, 3, 1, 1)                      # MSTr.py:2828-2829
            patches = torch.nn.functional.unfold(x, k, padding=p, stride=s).transpose(1, 2)
            kk = patches.shape[-1]
            pad = (-kk) % 4
            patches = torch.nn.functional.pad(patches, (0, pad))
            wmat = torch.nn.functional.pad(w.reshape(w.shape[0], kk), (0, pad))
            y = tcx_autograd.linear(patches, wmat, self.proj.bias)
            return tcx_autograd.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps), H, W
        y = ops.patch_embed_ln(x.contiguous(), self.proj.weight, self.proj.bias, s, p,
                               self.norm.weight, self.norm.bias, self.norm.eps)
        return y, H, W


# --------------------------------------------------------------------------------------
# RIPM: ResInception Patch Merging (reference MSTr.py:309-404, :670-732, :996-1050)
# --------------------------------------------------------------------------------------
class DWConv2d_BN(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=1, stride=1, norm_layer=nn.BatchNorm2d,
                 act_layer=nn.Hardswish, bn_weight_init=1, norm_cfg="BN"):
        super().__init__()
        if in_ch != out_ch or kernel_size != 3 or act_layer is not nn.Hardswish:
            raise NotImplementedError("DWConv2d_BN is built for the RIPM shape (3x3, in==out, Hardswish)")
        self.dwconv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, (kernel_size - 1) // 2, groups=out_ch, bias=False)
        self.pwconv = nn.Conv2d(out_ch, out_ch, 1, 1, 0, bias=False)
        self.bn = nn.BatchNorm2d(out_ch)
        self.act = act_layer()
        _xavier_convs(self.modules())
        self.bn.weight.data.fill_(bn_weight_init)
        self.bn.bias.data.zero_()

    def nhwc(self, x_nhwc, out=None):
        if self.training:
            # training row (MSTr.py:355-362): depthwise conv, 1x1 conv and BatchNorm(batch statistics)+Hardswish nodes
            y = tcx_autograd.dwconv3x3_nhwc(x_nhwc, self.dwconv.weight, self.dwconv.stride[0])
            y = tcx_autograd.bn_act(_lin_nhwc(y, self.pwconv), self.bn, ops.ACT_HARDSWISH)
            return y if out is None else out.copy_(y)
        _check_eval_bn(self, x_nhwc)
        return ops.ripm_dwsep_bn_hs(x_nhwc, self.dwconv.stride[0], self.dwconv.weight, self.pwconv.weight,
                                    *_bn(self.bn), out=out)

    def forward(self, x):
        return _as_nchw(self.nhwc(_nhwc(x)))


class Conv2d_BN(nn.Module):
    """Parameter holder for the 1x1 conv + BN inside ``ResBlock``."""

    def __init__(self, in_ch, out_ch, kernel_size=1, stride=1, pad=0, dilation=1, groups=1,
                 bn_weight_init=1, act_layer=None, norm_cfg="BN"):
        super().__init__()
        if kernel_size != 1 or stride != 1 or groups != 1:
            raise NotImplementedError("Conv2d_BN is built for 1x1/stride 1 only")
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, pad, dilation, groups, bias=False)
        self.bn = nn.BatchNorm2d(out_ch)
        nn.init.constant_(self.bn.weight, bn_weight_init)
        nn.init.constant_(self.bn.bias, 0)
        _xavier_convs(self.modules())
        self.hardswish = act_layer is nn.Hardswish
        if act_layer is not None and not self.hardswish:
            raise NotImplementedError("Conv2d_BN: only Hardswish / no activation are built")
        self.act_layer = act_layer() if act_layer is not None else nn.Identity()

    def nhwc_train(self, x_nhwc):
        return tcx_autograd.bn_act(_lin_nhwc(x_nhwc, self.conv), self.bn, ops.ACT_HARDSWISH if self.hardswish else ops.ACT_NONE)

    def forward(self, x):
        xh = _nhwc(x)
        if self.training:
            return _as_nchw(self.nhwc_train(xh))
        _check_eval_bn(self, x)
        B, H, W, C = xh.shape
        y = ops.linear_bn_act(xh.view(-1, C), self.conv.weight, *_bn(self.bn), hardswish=self.hardswish)
        return _as_nchw(y.view(B, H, W, -1))


class DWCPatchEmbed(nn.Module):
    def __init__(self, in_chans=3, embed_dim=768, patch_size=16, stride=1, pad=0,
                 act_layer=nn.Hardswish, norm_cfg='BN'):
        super().__init__()
        self.stride = stride
        self.patch_conv = DWConv2d_BN(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                                      act_layer=nn.Hardswish, norm_cfg=norm_cfg)

    def forward(self, x):
        return self.patch_conv(x)


class Patch_Embed_stage(nn.Module):
    """Three chained dw3x3 -> pw1x1 -> BN -> Hardswish blocks; returns the three intermediate maps.
    The three outputs are written into one stacked [3,B,H,W,C] buffer so the Multi-Branch stage can run its
    three branches as one grouped launch."""

    def __init__(self, embed_dim, num_path=3, isPool=False, norm_cfg=dict(type="BN")):
        super().__init__()
        self.patch_embeds = nn.ModuleList([
            DWCPatchEmbed(in_chans=embed_dim, embed_dim=embed_dim, patch_size=3,
                          stride=2 if isPool and idx == 0 else 1, pad=1, norm_cfg='BN')
            for idx in range(num_path)])

    def nhwc(self, x_nhwc):
        B, H, W, C = x_nhwc.shape
        if self.training:
            outs, cur = [], x_nhwc
            for pe in self.patch_embeds:
                cur = pe.patch_conv.nhwc(cur)
                outs.append(cur)
            return torch.stack(outs, 0)
        s0 = self.patch_embeds[0].stride
        Ho, Wo = (H + 2 - 3) // s0 + 1, (W + 2 - 3) // s0 + 1
        stacked = torch.empty((len(self.patch_embeds), B, Ho, Wo, C), device=x_nhwc.device, dtype=x_nhwc.dtype)
        cur = x_nhwc
        for i, pe in enumerate(self.patch_embeds):
            cur = pe.patch_conv.nhwc(cur, out=stacked[i])
        return stacked

    def forward(self, x):
        stacked = self.nhwc(_nhwc(x))
        return [_as_nchw(stacked[i]) for i in range(stacked.shape[0])]


class ResBlock(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.Hardswish, norm_cfg="BN"):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.conv1 = Conv2d_BN(in_features, hidden_features, act_layer=act_layer, norm_cfg=norm_cfg)
        self.dwconv = nn.Conv2d(hidden_features, hidden_features, 3, 1, 1, bias=False, groups=hidden_features)
        self.norm = nn.BatchNorm2d(hidden_features)
        self.act = act_layer()
        self.conv2 = Conv2d_BN(hidden_features, out_features, norm_cfg=norm_cfg)
        # reference MSTr.py:1025-1039 via self.apply (children first): conv1.conv, dwconv, conv2.conv
        _xavier_convs([self.conv1.conv, self.dwconv, self.conv2.conv])

    def nhwc(self, x_nhwc):
        if self.training:
            # training row (MSTr.py:1042-1050)
            f = self.conv1.nhwc_train(x_nhwc)
            f = tcx_autograd.bn_act(tcx_autograd.dwconv3x3_nhwc(f, self.dwconv.weight, 1), self.norm, ops.ACT_HARDSWISH)
            return x_nhwc + self.conv2.nhwc_train(f)
        _check_eval_bn(self, x_nhwc)
        return ops.resblock(x_nhwc, self.conv1.conv.weight, _bn(self.conv1.bn), self.dwconv.weight, _bn(self.norm),
                            self.conv2.conv.weight, _bn(self.conv2.bn))

    def forward(self, x):
        return _as_nchw(self.nhwc(_nhwc(x)))


# --------------------------------------------------------------------------------------
# Multi-Branch transformer (reference MSTr.py:734-993)
# --------------------------------------------------------------------------------------
class ConvPosEnc(nn.Module):
    def __init__(self, dim, k=3):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, k, 1, k // 2, groups=dim)

    def forward(self, x, size):
        H, W = size
        if _recording(x, self.proj.weight):
            return tcx_autograd.dwconv_tokens(x, H, W, _shared(self.proj.weight), _shared(self.proj.bias), True)
        return ops.dwconv_tokens(x.contiguous(), H, W, self.proj.weight, self.proj.bias, add_input=True)


class ConvRelPosEnc(nn.Module):
    def __init__(self, Ch, h, window):
        super().__init__()
        if isinstance(window, int):
            window = {window: h}
        elif not isinstance(window, dict):
            raise ValueError()
        self.window = window
        self.conv_list = nn.ModuleList()
        self.head_splits = []
        for cur_window, cur_head_split in window.items():
            ch = cur_head_split * Ch
            self.conv_list.append(nn.Conv2d(ch, ch, kernel_size=(cur_window, cur_window),
                                            padding=(cur_window // 2, cur_window // 2), groups=ch))
            self.head_splits.append(cur_head_split)
        self.channel_splits = [x * Ch for x in self.head_splits]

    def forward(self, q, v, size):
        """q, v: [B,h,N,Ch] -> q * dwconv(v)  (reference MSTr.py:801-823)."""
        H, W = size
        return ops.crpe(q.contiguous(), v.contiguous(), H, W, [c.weight for c in self.conv_list],
                        [c.bias for c in self.conv_list], self.head_splits)


class FactorAtt_ConvRelPosEnc(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 shared_crpe=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.crpe = shared_crpe

    def args(self):
        c = self.crpe
        return (self.num_heads, self.scale, self.qkv.weight, self.qkv.bias,
                [m.weight for m in c.conv_list], [m.bias for m in c.conv_list], list(c.head_splits),
                self.proj.weight, self.proj.bias)

    def forward(self, x, size, residual=None):
        H, W = size
        if _recording(x, self.qkv.weight):
            heads, _, qkvw, qkvb, cw, cb, splits, pw, pb = self.args()
            cw, cb = [_shared(w) for w in cw], [_shared(b) for b in cb]
            ops._check_crpe(splits, cw, heads)
            return tcx_autograd.factor_att(x, H, W, heads, qkvw, qkvb, cw, cb, pw, pb, residual=residual)
        return ops.mb_factor_attn(x.contiguous(), H, W, *self.args(), residual=residual)


class MHCABlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=3, drop_path=0.0, qkv_bias=True, qk_scale=None,
                 norm_layer='LN', shared_cpe=None, shared_crpe=None):
        super().__init__()
        self.cpe = shared_cpe
        self.crpe = shared_crpe
        self.factoratt_crpe = FactorAtt_ConvRelPosEnc(dim, num_heads=num_heads, qkv_bias=qkv_bias,
                                                      qk_scale=qk_scale, shared_crpe=shared_crpe)
        self.mlp = MixFFN_skip(dim, dim * mlp_ratio)
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)

    def forward(self, x, size):
        if _recording(x, self.norm1.weight):
            # training row (MSTr.py:935-946): position-encoding conv, attention and Mix-FFN are autograd nodes on the
            # library's kernels; the two residual additions are ATen adds
            n1, n2 = self.norm1, self.norm2
            x = self.cpe(x, size)
            xn, xr = tcx_autograd.layernorm_res(x, n1.weight, n1.bias, n1.eps)
            x = self.factoratt_crpe(xn, size, residual=xr)
            xn, xr = tcx_autograd.layernorm_res(x, n2.weight, n2.bias, n2.eps)
            return self.mlp(xn, size[0], size[1], residual=xr)
        return ops.mhca_blocks(x.contiguous().unsqueeze(0), size[0], size[1], [[self]])[0]


class MHCAEncoder(nn.Module):
    def __init__(self, dim, num_layers=1, num_heads=8, mlp_ratio=3, drop_path_list=[], qk_scale=None,
                 crpe_window={3: 2, 5: 3, 7: 3}):
        super().__init__()
        self.num_layers = num_layers
        self.cpe = ConvPosEnc(dim, k=3)
        self.crpe = ConvRelPosEnc(Ch=dim // num_heads, h=num_heads, window=crpe_window)
        self.MHCA_layers = nn.ModuleList([
            MHCABlock(dim, num_heads=num_heads, mlp_ratio=mlp_ratio, drop_path=drop_path_list[idx],
                      qk_scale=qk_scale, shared_cpe=self.cpe, shared_crpe=self.crpe)
            for idx in range(self.num_layers)])

    def run_blocks_train(self, x, size):
        """The blocks one after the other on tokens x [B, N, C] (training row); the parameters they share (ConvPosEnc,
        ConvRelPosEnc) are handed out as per-block aliases of a fan-out node, see _shared."""
        L = len(self.MHCA_layers)
        shared = [self.cpe.proj.weight, self.cpe.proj.bias] + [t for c in self.crpe.conv_list for t in (c.weight, c.bias)]
        prev = getattr(_FANOUT, "d", None)
        _FANOUT.d = {id(p): list(tcx_autograd.fan_out(p, L)) for p in shared if p.requires_grad} if L > 1 else None
        try:
            for blk in self.MHCA_layers:
                x = blk(x, size)
        finally:
            _FANOUT.d = prev
        return x

    def forward(self, x, size):
        H, W = size
        B = x.shape[0]
        if _recording(x, self.cpe.proj.weight):
            return _as_nchw(self.run_blocks_train(x, size).reshape(B, H, W, -1))
        y = ops.mhca_blocks(x.contiguous().unsqueeze(0), H, W, [list(self.MHCA_layers)])[0]
        return _as_nchw(y.view(B, H, W, -1))


def dpr_generator(drop_path_rate, num_layers, num_stages):
    vals = [v.item() for v in torch.linspace(0, drop_path_rate, sum(num_layers))]
    out, cur = [], 0
    for i in range(num_stages):
        out.append(vals[cur:cur + num_layers[i]])
        cur += num_layers[i]
    return out


# --------------------------------------------------------------------------------------
# IFF: coordinate attention over the four concatenated maps (reference MSTr.py:1270-1348)
# --------------------------------------------------------------------------------------
class silu_sigmoid(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.silu = nn.SiLU(inplace=inplace)


class silu_swish(nn.Module):
    """t * min(SiLU(t+3)/6, 1); evaluated inside the IFF kernel (holder only)."""

    def __init__(self, inplace=True):
        super().__init__()
        self.sigmoid = silu_sigmoid(inplace=inplace)


class CoordAtt(nn.Module):
    def __init__(self, inp, oup, reduction=32):
        super().__init__()
        mip = max(8, inp // reduction)
        self.conv1 = nn.Conv2d(inp, mip, kernel_size=1, stride=1, padding=0)
        self.bn1 = nn.BatchNorm2d(mip)
        self.act = silu_swish()
        self.conv_h = nn.Conv2d(mip, inp, kernel_size=1, stride=1, padding=0)
        self.conv_w = nn.Conv2d(mip, inp, kernel_size=1, stride=1, padding=0)
        self.conv_in_out = nn.Conv2d(inp, oup, kernel_size=1, stride=1, padding=0)

    def nhwc(self, maps):
        """maps: list of NHWC tensors whose channel concatenation is the reference's input."""
        if self.training:
            # training row (MSTr.py:1322-1348): pooling, gating, the four 1x1 convs and BatchNorm+silu_swish are autograd nodes
            x = torch.cat(maps, dim=-1) if len(maps) > 1 else maps[0]
            H = x.shape[1]
            y = tcx_autograd.bn_act(_lin_nhwc(tcx_autograd.coord_pool(x), self.conv1), self.bn1, ops.ACT_SILU_SWISH)
            z = torch.cat([_lin_nhwc(y[:, :H], self.conv_h), _lin_nhwc(y[:, H:], self.conv_w)], dim=1)
            return _lin_nhwc(tcx_autograd.coord_gate(x, z), self.conv_in_out)
        _check_eval_bn(self, maps[0])
        return ops.iff_coordatt(maps, self.conv1.weight, self.conv1.bias, _bn(self.bn1),
                                self.conv_h.weight, self.conv_h.bias, self.conv_w.weight, self.conv_w.bias,
                                self.conv_in_out.weight, self.conv_in_out.bias)

    def forward(self, x):
        return _as_nchw(self.nhwc([_nhwc(x)]))


class MHCA_stage(nn.Module):
    def __init__(self, embed_dim, out_embed_dim, num_layers=1, num_heads=8, mlp_ratio=3, num_path=4,
                 norm_cfg="BN", drop_path_list=[], concat='normal', use_sa=True, sa_ker=7):
        super().__init__()
        if concat != 'coord':
            raise NotImplementedError("only concat='coord' (IFF, the default of the entry scripts) is built")
        self.concat = concat
        self.mhca_blks = nn.ModuleList([
            MHCAEncoder(embed_dim, num_layers, num_heads, mlp_ratio, drop_path_list=drop_path_list)
            for _ in range(num_path)])
        self.InvRes = ResBlock(in_features=embed_dim, out_features=embed_dim, norm_cfg=norm_cfg)
        self.aggregate = CoordAtt(inp=embed_dim * (num_path + 1), oup=out_embed_dim, reduction=16)

    def nhwc(self, stacked):
        """stacked: [P,B,H,W,C] RIPM outputs -> [B,H,W,C_out]."""
        P, B, H, W, C = stacked.shape
        if self.training:
            def branch(i):
                def run():
                    t = self.mhca_blks[i].run_blocks_train(stacked[i].reshape(B, H * W, C), (H, W))
                    return t.reshape(B, H, W, C)
                return run
            # the residual block and the three transformer branches are independent until the IFF concatenation
            maps = _parallel_branches(stacked.device, [lambda: self.InvRes.nhwc(stacked[0])] + [branch(i) for i in range(P)],
                                      inputs=(stacked,))
            return self.aggregate.nhwc(maps)
        # the residual branch only needs path 0 of the RIPM output: it runs on a side stream next to the three
        # transformer branches (a parallel branch of the captured graph)
        main = torch.cuda.current_stream(stacked.device)
        side = _side_stream(stacked.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            res = self.InvRes.nhwc(stacked[0])
        enc = ops.mhca_blocks(stacked.view(P, B, H * W, C), H, W,
                              [list(e.MHCA_layers) for e in self.mhca_blks])
        main.wait_stream(side)
        res.record_stream(main)
        maps = [res] + [enc[i].view(B, H, W, C) for i in range(P)]
        return self.aggregate.nhwc(maps)

    def forward(self, inputs):
        stacked = torch.stack([_nhwc(x) for x in inputs], 0)
        return _as_nchw(self.nhwc(stacked))


class MSViT(nn.Module):
    def __init__(self, image_size, in_dim, key_dim, value_dim, layers, head_count=1, dil_conv=1,
                 token_mlp='mix_skip', MSViT_config=1, concat='normal', use_sa_list=[True, True, False], sa_ker=7):
        super().__init__()
        self.Hs = [56, 28, 14, 7]
        self.Ws = [56, 28, 14, 7]
        # dead 1x1 convs, kept for state_dict/init parity (reference MSTr.py:1567-1570)
        for i in range(4):
            setattr(self, 'conv1_1_s%d' % (i + 1), nn.Conv2d(3 * in_dim[i], in_dim[i], 1))
        num_path, num_layers, num_heads, mlp_ratios = [3, 3, 3], [3, 8, 3], [8, 8, 8], [4, 4, 4]
        dpr = dpr_generator(0.0, num_layers, 3)
        for s in range(3):
            setattr(self, 'patch_embed_stage%d' % (s + 2),
                    Patch_Embed_stage(in_dim[s], num_path=num_path[s], isPool=True, norm_cfg='BN'))
        for s in range(3):
            setattr(self, 'mhca_stage%d' % (s + 2),
                    MHCA_stage(in_dim[s], in_dim[s + 1], num_layers[s], num_heads[s], mlp_ratios[s], num_path[s],
                               norm_cfg='BN', drop_path_list=dpr[s], concat=concat, use_sa=use_sa_list[s],
                               sa_ker=sa_ker))
        self.patch_embed1 = OverlapPatchEmbeddings(image_size, 7, 4, 3, 3, in_dim[0])
        self.cpe = ConvPosEnc(in_dim[0], k=3)  # dead (call commented out in the reference, MSTr.py:1716)
        self.block1 = nn.ModuleList([
            EfficientTransformerBlock(in_dim[0], key_dim[0], value_dim[0], head_count, token_mlp)
            for _ in range(layers[0])])
        self.norm1 = nn.LayerNorm(in_dim[0])

    def nhwc(self, x):
        """[B,3,H,W] image -> four NHWC maps."""
        B = x.shape[0]
        t, H, W = self.patch_embed1(x)
        for blk in self.block1:
            t = blk(t, H, W)
        if _recording(t, self.norm1.weight):
            t = tcx_autograd.layernorm(t, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        else:
            t = ops.layernorm(t, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        cur = t.reshape(B, H, W, -1)
        outs = [cur]
        cut = getattr(self, '_grad_cut', None)
        for s in (2, 3, 4):
            stacked = getattr(self, 'patch_embed_stage%d' % s).nhwc(cur)
            cur = getattr(self, 'mhca_stage%d' % s).nhwc(stacked)
            outs.append(cur)
            if s == 2 and cut is not None and torch.is_grad_enabled() and cur.requires_grad:
                # data-parallel training (runtime.TrainStepGraph): the backward is run in two pieces so that the gradient
                # all-reduce of everything behind this point overlaps the backward of stages 1-2.  The maps that cross the
                # cut continue as leaves; the runner feeds their gradients back into the originals (EARLY_MODULES below).
                origs = list(outs)
                outs = [o.detach().requires_grad_(True) for o in origs]
                cur = outs[-1]
                del cut[:]
                cut.extend(zip(origs, outs))
        return outs

    # the modules in front of the cut above (their gradients are complete only after the second backward piece)
    EARLY_MODULES = ('patch_embed1', 'block1', 'norm1', 'patch_embed_stage2', 'mhca_stage2')

    def forward(self, x):
        return [_as_nchw(m) for m in self.nhwc(x)]


# --------------------------------------------------------------------------------------
# Dual Transformer Bridge (reference MSTr.py:2209-2442)
# --------------------------------------------------------------------------------------
class Scale_reduce(nn.Module):
    def __init__(self, dim, reduction_ratio):
        super().__init__()
        if len(reduction_ratio) != 4:
            raise NotImplementedError("Scale_reduce is built for the 4-scale bridge")
        self.dim = dim
        self.reduction_ratio = reduction_ratio
        self.sr0 = nn.Conv2d(dim, dim, reduction_ratio[3], reduction_ratio[3])
        self.sr1 = nn.Conv2d(dim * 2, dim * 2, reduction_ratio[2], reduction_ratio[2])
        self.sr2 = nn.Conv2d(dim * 5, dim * 5, reduction_ratio[1], reduction_ratio[1])
        self.norm = nn.LayerNorm(dim)

    def args(self):
        return (self.sr0.weight, self.sr0.bias, self.sr1.weight, self.sr1.bias, self.sr2.weight, self.sr2.bias,
                self.norm.weight, self.norm.bias, self.norm.eps)

    def forward(self, x):
        if _recording(x, self.sr0.weight):
            return self._forward_train(x)
        return ops.scale_reduce(x.contiguous(), *self.args())

    def _forward_train(self, x):
        """Training row (MSTr.py:2225-2249): the three strided convs (im2row + GEMM), the channel-group packing and the raw
        stage-4 rows are ONE autograd node on the library's kernels; the LayerNorm is the usual node."""
        packed = tcx_autograd.scale_reduce_pack(x, self.sr0.weight, self.sr0.bias, self.sr1.weight, self.sr1.bias,
                                                self.sr2.weight, self.sr2.bias)
        n = self.norm
        return tcx_autograd.layernorm(packed, n.weight, n.bias, n.eps)


class M_EfficientSelfAtten(nn.Module):
    def __init__(self, dim, head, reduction_ratio):
        super().__init__()
        if head != 1 or dim != 64:
            raise NotImplementedError("bridge attention is built for dim=64, head=1")
        self.head = head
        self.reduction_ratio = reduction_ratio
        self.scale = (dim // head) ** -0.5
        self.q = nn.Linear(dim, dim, bias=True)
        self.kv = nn.Linear(dim, dim * 2, bias=True)
        self.proj = nn.Linear(dim, dim)
        if reduction_ratio is not None:
            self.scale_reduce = Scale_reduce(dim, reduction_ratio)

    def args(self):
        return (self.scale, self.q.weight, self.q.bias, self.kv.weight, self.kv.bias,
                self.proj.weight, self.proj.bias) + self.scale_reduce.args()

    def slots(self):
        """The 14 pointer slots of tcx_bridge_sr_attn_fwd."""
        sr = self.scale_reduce
        return [self.q.weight, self.q.bias, self.kv.weight, self.kv.bias, self.proj.weight, self.proj.bias,
                sr.sr0.weight.reshape(64, -1), sr.sr0.bias, sr.sr1.weight.reshape(128, -1), sr.sr1.bias,
                sr.sr2.weight.reshape(320, -1), sr.sr2.bias, sr.norm.weight, sr.norm.bias]

    def forward(self, x, residual=None):
        if _recording(x, self.q.weight):
            # training row (MSTr.py:2267-2292): Linear / Scale_reduce / attention-core / Linear autograd nodes
            q = tcx_autograd.linear(x, self.q.weight, self.q.bias)
            kv = tcx_autograd.linear(self.scale_reduce(x), self.kv.weight, self.kv.bias)
            return tcx_autograd.linear(tcx_autograd.attn_core(q, kv, self.scale), self.proj.weight, self.proj.bias, residual)
        return ops.bridge_sr_attn(x.contiguous(), *self.args(), residual=residual)


class M_EfficientChannelAtten(nn.Module):
    """Layer-1 bridge attention; its ``scale_reduce`` is dead (reference MSTr.py:2306-2307)."""

    def __init__(self, dim, head, reduction_ratio):
        super().__init__()
        if head != 1:
            raise NotImplementedError("bridge channel attention is built for head=1")
        self.head = head
        self.reduction_ratio = reduction_ratio
        self.scale = (dim // head) ** -0.5
        self.q = nn.Linear(dim, dim, bias=True)
        self.k = nn.Linear(dim, dim, bias=True)
        self.v = nn.Linear(dim, dim, bias=True)
        self.proj = nn.Linear(dim, dim)
        if reduction_ratio is not None:
            self.scale_reduce = Scale_reduce(dim, reduction_ratio)

    def slots(self):
        """The 8 pointer slots of tcx_eff_attn_fwd."""
        return [self.k.weight, self.k.bias, self.q.weight, self.q.bias, self.v.weight, self.v.bias,
                self.proj.weight, self.proj.bias]

    def forward(self, x, residual=None):
        if _recording(x, self.q.weight):
            # training row (MSTr.py:2309-2353): the raw [N, C] -> [C, N] re-reading followed by our token-major layout is a
            # transposed copy; the attention core and the four Linear layers are autograd nodes on the library's kernels
            B, N, C = x.shape
            k, q, v = (tcx_autograd.linear(x, m.weight, m.bias).reshape(B, C, N).transpose(1, 2) for m in (self.k, self.q, self.v))
            return tcx_autograd.linear(tcx_autograd.ea_core(k, q, v), self.proj.weight, self.proj.bias, residual)
        return ops.eff_attn(x.contiguous(), self.k.weight, self.k.bias, self.q.weight, self.q.bias,
                            self.v.weight, self.v.bias, self.proj.weight, self.proj.bias,
                            residual=residual, reinterpret=True)


class BridgLayer_4(nn.Module):
    def __init__(self, dims, head, reduction_ratios, ch_att):
        super().__init__()
        self.norm1 = nn.LayerNorm(dims)
        self.attn = (M_EfficientChannelAtten if ch_att else M_EfficientSelfAtten)(dims, head, reduction_ratios)
        self.norm2 = nn.LayerNorm(dims)
        self.mixffn1 = MixFFN_skip(dims, dims * 4)
        self.mixffn2 = MixFFN_skip(dims * 2, dims * 8)
        self.mixffn3 = MixFFN_skip(dims * 5, dims * 20)
        self.mixffn4 = MixFFN_skip(dims * 8, dims * 32)

    def forward(self, inputs):
        first = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        if _recording(first, self.norm1.weight):
            return self._forward_train(inputs)
        if isinstance(inputs, (list, tuple)):
            # a C_k-channel NHWC pixel is C_k/64 consecutive 64-wide tokens
            inputs = ops.bridge_regroup([_nhwc(c) for c in inputs])
        ch = isinstance(self.attn, M_EfficientChannelAtten)
        return ops.bridge_layer(inputs, self.norm1.weight, self.norm1.bias, self.norm1.eps, ch, self.attn.slots(),
                                self.attn.scale, self.norm2.weight, self.norm2.bias,
                                [m.args() for m in (self.mixffn1, self.mixffn2, self.mixffn3, self.mixffn4)])


def _bridge_tokens_train(maps):
    """NCHW-shaped maps -> [B, Ntok, 64] (MSTr.py:2380-2386): one regroup launch (its gradient: one split launch)."""
    return tcx_autograd.bridge_merge([_nhwc(m) for m in maps])


def _bridge_layer_train(self, inputs):
    """BridgLayer_4.forward (MSTr.py:2373-2409) as autograd nodes.  Both skip connections use the pre-norm pattern of
    LayerNormResFn (the two gradients of x / tx1 meet inside the LayerNorm backward kernel); the attention's output projection
    adds its skip in the GEMM epilogue; the four per-scale slabs come from / go back into the token buffer in one launch each
    (BridgeSplitFn / BridgeMergeFn), so the layer runs no ATen slice, cat, fill or add kernels."""
    x = _bridge_tokens_train(inputs) if isinstance(inputs, (list, tuple)) else inputs
    B, ntok, C = x.shape
    S = ops._bridge_side(ntok)
    n1, n2 = self.norm1, self.norm2
    xn, xres = tcx_autograd.layernorm_res(x, n1.weight, n1.bias, n1.eps)
    tx1 = self.attn(xn, residual=xres)
    tx, tx1res = tcx_autograd.layernorm_res(tx1, n2.weight, n2.bias, n2.eps)
    slabs = tcx_autograd.bridge_split(tx)
    fns = []
    for mlp, hw, slab in zip((self.mixffn1, self.mixffn2, self.mixffn3, self.mixffn4), (S, S // 2, S // 4, S // 8), slabs):
        fns.append(lambda mlp=mlp, hw=hw, slab=slab: mlp(slab, hw, hw))
    # the four per-scale Mix-FFNs are independent (MSTr.py:2394-2402)
    ys = _parallel_branches(x.device, fns, inputs=tuple(slabs))
    return tcx_autograd.bridge_merge(ys, tx1res)


BridgLayer_4._forward_train = _bridge_layer_train


class BridgeBlock_4(nn.Module):
    def __init__(self, dims, head, reduction_ratios, br_ch_att_list):
        super().__init__()
        for i in range(4):
            setattr(self, 'bridge_layer%d' % (i + 1), BridgLayer_4(dims, head, reduction_ratios, br_ch_att_list[i]))

    def tokens(self, x):
        first = x[0] if isinstance(x, (list, tuple)) else x
        if _recording(first, self.bridge_layer1.norm1.weight):
            t = _bridge_tokens_train(x) if isinstance(x, (list, tuple)) else x
            for i in range(4):
                t = getattr(self, 'bridge_layer%d' % (i + 1))(t)
            return t
        if isinstance(x, (list, tuple)):
            x = ops.bridge_regroup([_nhwc(c) for c in x])
        layers = []
        for i in range(4):
            lay = getattr(self, 'bridge_layer%d' % (i + 1))
            layers.append((lay.norm1.weight, lay.norm1.bias, isinstance(lay.attn, M_EfficientChannelAtten), lay.attn.slots(),
                           lay.norm2.weight, lay.norm2.bias,
                           [m.args() for m in (lay.mixffn1, lay.mixffn2, lay.mixffn3, lay.mixffn4)]))
        l0 = self.bridge_layer1
        return ops.bridge_block(x, layers, l0.attn.scale, l0.norm1.eps)

    def forward(self, x):
        t = self.tokens(x)
        B, ntok, C = t.shape
        outs, off = [], 0
        S = ops._bridge_side(ntok)      # 56 at 224x224 (the reference hard-codes 56/28/14/7, MSTr.py:2432-2435)
        if _recording(t, self.bridge_layer1.norm1.weight):
            # training row: the four dense slabs in one launch (NCHW-shaped views of NHWC storage, like the slices below)
            return [m.reshape(B, S >> k, S >> k, -1).permute(0, 3, 1, 2) for k, m in enumerate(tcx_autograd.bridge_split(t))]
        for hw, mult in ((S, 1), (S // 2, 2), (S // 4, 5), (S // 8, 8)):
            n = hw * hw * mult
            outs.append(t[:, off:off + n, :].reshape(B, hw, hw, C * mult).permute(0, 3, 1, 2))
            off += n
        return outs


# --------------------------------------------------------------------------------------
# Top level (reference MSTr.py:2759-2852)
# --------------------------------------------------------------------------------------
class MSTransception(nn.Module):
    def __init__(self, num_classes=9, head_count=8, dil_conv=1, token_mlp_mode="mix_skip", MSViT_config=2,
                 concat='coord', have_bridge='original', use_sa_config=1, sa_ker=7, Stage_3or4=3, inter='res',
                 num_sp=1, br_ch_att_list=[True, False, False, False], image_size=224):
        """``image_size`` is an extension (the reference hard-codes 224 and fails on other sizes, SURVEY.md section 0
        defect 4): any multiple of 32 builds the same parameters and runs, e.g. 256 for BASELINE config 5."""
        super().__init__()
        if image_size % 32:
            raise ValueError("image_size must be a multiple of 32")
        if Stage_3or4 != 3:
            raise NotImplementedError("only Stage_3or4=3 (MSViT) is built")
        if have_bridge in ('sp', 'para'):
            raise NotImplementedError("only have_bridge='original' (BridgeBlock_4) is built")
        dims = [64, 128, 320, 512]
        use_sa_list = [True, True, True, False]
        self.backbone = MSViT(image_size=image_size, in_dim=dims, key_dim=dims, value_dim=dims, layers=[2, 2, 2, 2],
                              head_count=head_count, dil_conv=dil_conv, token_mlp=token_mlp_mode,
                              MSViT_config=MSViT_config, concat=concat, use_sa_list=use_sa_list, sa_ker=sa_ker)
        self.reduction_ratios = [1, 2, 4, 8]
        self.have_bridge = have_bridge
        self.bridge = BridgeBlock_4(64, 1, self.reduction_ratios, br_ch_att_list)
        fs = image_size // 32
        io = [[32, 64, 64, 64], [144, 128, 128, 128], [288, 320, 320, 320], [512, 512, 512, 512]]
        self.decoder_3 = MyDecoderLayer((fs, fs), io[3], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_2 = MyDecoderLayer((fs * 2, fs * 2), io[2], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_1 = MyDecoderLayer((fs * 4, fs * 4), io[1], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_0 = MyDecoderLayer((fs * 8, fs * 8), io[0], head_count, token_mlp_mode, n_class=num_classes,
                                        is_last=True)

    def forward(self, x):
        ops.require_cuda(x)
        # (an nn.DataParallel replica holds its weights as plain attributes: ``parameters()`` is empty there, attribute access works)
        recording = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                                 self.backbone.norm1.weight.requires_grad or self.decoder_0.concat_linear.weight.requires_grad)
        if recording and not self.training:
            raise NotImplementedError(
                "transception_b200: backward through eval-mode BatchNorm (running statistics) is not built; call .train() "
                "for a training step or wrap inference in torch.no_grad()")
        # a 1-channel input is read three times by the stem kernel (reference repeats it, MSTr.py:2828-2829)
        maps = self.backbone.nhwc(x)
        if self.have_bridge != "None":
            if recording:
                # training row: every module below routes to the autograd nodes of transception_b200/autograd.py
                tokens = tcx_autograd.bridge_merge(maps)
            else:
                tokens = ops.bridge_regroup(maps)
            maps = [m.permute(0, 2, 3, 1) for m in self.bridge(tokens)]
        b, _, _, c = maps[3].shape
        t3 = self.decoder_3(maps[3].reshape(b, -1, c))
        t2 = self.decoder_2(t3, maps[2])
        t1 = self.decoder_1(t2, maps[1])
        return self.decoder_0(t1, maps[0])
