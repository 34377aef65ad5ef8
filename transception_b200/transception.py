"""Drop-in mirror of the earlier model variant ``networks/Transception.py::Transception``.

Same constructor signature, attribute names, ``state_dict`` keys (561 tensors) and same-seed initialisation as the
reference, so ``Transception(num_classes=9, head_count=1, dil_conv=1, token_mlp_mode="mix_skip").load_state_dict(
reference_checkpoint, strict=True)`` works.  The forward runs on the sm_100a library only (``ops.py`` -> C ABI); tokens
stay NHWC / ``[B, N, C]`` between kernels and every GEMM operand is fp16 with fp32 accumulation (the variant has a single
back end — the fp16 pipeline).

Built structure: ``MiT_3inception`` (Transception.py:362-551) with ``dil_conv=1``, ``token_mlp='mix_skip'``,
``head_count=1`` and both fusion modes: ``concat='original'`` (1x1 conv over the channel concat) and anything else =
``SK_Block`` (selective-kernel fusion, eval-mode BatchNorm).  Stage 1, the Mix-FFN and the decoder are
the modules of ``mstr.py`` (identical code in the reference: Transception.py:90-186, :892-1008; EffSegformer.py:7-46).
"""
import torch
import torch.nn as nn

from . import ops
from .mstr import (EfficientTransformerBlock, MixFFN_skip, MyDecoderLayer, OverlapPatchEmbeddings)


class OverlapPatchEmbeddings_fuse(nn.Module):
    """Dilated strided conv + LayerNorm (reference EffSegformer.py:117-131)."""

    def __init__(self, img_size=224, patch_size=7, stride=4, padding=1, dilation=1, in_ch=3, dim=768):
        super().__init__()
        self.dim = dim
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_ch, dim, patch_size, stride, padding, dilation)
        self.norm = nn.LayerNorm(dim)

    def slots(self):
        p = self.proj
        return (p.weight, p.bias, self.norm.weight, self.norm.bias, p.kernel_size[0], p.stride[0], p.padding[0], p.dilation[0])

    def forward(self, x):
        """x: NCHW map -> (tokens, H, W), like the reference.  Runs the pair kernel with this branch twice removed: use
        ``MiT_3inception`` for the fused two-branch path."""
        t, H1, W1, _, _ = ops.dual_patch_embed(x.permute(0, 2, 3, 1), self.slots(), self.slots(), self.norm.eps)
        return t[:, :H1 * W1].contiguous(), H1, W1


class FuseEfficientAttention(nn.Module):
    """Efficient attention over the concatenated token list with the reference's raw ``[N, C] -> [C, N]`` re-reading of
    keys / queries / values (reference Transception.py:18-87)."""

    def __init__(self, in_channels, key_channels, value_channels, head_count):
        super().__init__()
        if head_count != 1 or key_channels != in_channels or value_channels != in_channels:
            raise NotImplementedError("FuseEfficientAttention is built for head_count=1, key=value=in channels "
                                      "(the Transception() defaults)")
        self.in_channels, self.key_channels = in_channels, key_channels
        self.head_count, self.value_channels = head_count, value_channels
        self.keys = nn.Linear(in_channels, key_channels, bias=True)
        self.queries = nn.Linear(in_channels, key_channels, bias=True)
        self.values = nn.Linear(in_channels, value_channels, bias=True)
        self.reprojection = nn.Linear(value_channels, in_channels)

    def args(self):
        return (self.keys.weight, self.keys.bias, self.queries.weight, self.queries.bias,
                self.values.weight, self.values.bias, self.reprojection.weight, self.reprojection.bias)

    def forward(self, input_):
        return ops.fuse_eff_attn(input_, *self.args())


class EfficientTransformerBlockFuse(nn.Module):
    """reference Transception.py:188-250 (two-branch token layout)."""

    def __init__(self, in_dim, key_dim, value_dim, head_count=1, token_mlp='mix'):
        super().__init__()
        if token_mlp != 'mix_skip':
            raise NotImplementedError("only token_mlp='mix_skip' is built")
        self.norm1 = nn.LayerNorm(in_dim)
        self.attn = FuseEfficientAttention(in_channels=in_dim, key_channels=key_dim, value_channels=value_dim,
                                           head_count=head_count)
        self.norm2 = nn.LayerNorm(in_dim)
        self.mlp1 = MixFFN_skip(in_dim, int(in_dim * 4))
        self.mlp2 = MixFFN_skip(in_dim, int(in_dim * 4))

    def forward(self, x, nfx1_len, nfx2_len, H1, W1, H2, W2):
        if x.shape[1] != nfx1_len + nfx2_len or nfx1_len != H1 * W1 or nfx2_len != H2 * W2:
            raise NotImplementedError("only the two-branch layout (n1 = H1*W1, n2 = H2*W2 tokens) is built")
        return ops.fuse_block(x, H1, W1, H2, W2, self.norm1.weight, self.norm1.bias, self.norm1.eps, self.attn.args(),
                              self.norm2.weight, self.norm2.bias, self.mlp1.args(), self.mlp2.args())


class SK_Block(nn.Module):
    """Selective-kernel fusion of the two branch maps (reference Transception.py:306-358).  Inside ``MiT_3inception`` it is
    fused with the stage LayerNorm and the nearest upsample (``ops.fuse_merge_sk``); the module itself holds the parameters."""

    def __init__(self, in_ch, out_ch, num_path=3, reduction=16, group=1, L=32):
        super().__init__()
        self.d = max(L, in_ch // reduction)
        self.fc = nn.Linear(in_ch, self.d)
        self.fcs = nn.ModuleList([nn.Linear(self.d, in_ch) for _ in range(num_path)])
        self.softmax = nn.Softmax(dim=0)
        self.conv_bn_ac = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=(1, 1)), nn.ReLU(inplace=True),
                                        nn.BatchNorm2d(out_ch))

    def forward(self, x):
        raise NotImplementedError("SK_Block runs fused inside MiT_3inception.stage (ops.fuse_merge_sk); the standalone "
                                  "list-of-maps call is not built")


class MiT_3inception(nn.Module):
    """reference Transception.py:362-551."""

    def __init__(self, image_size, in_dim, key_dim, value_dim, layers, head_count=1, dil_conv=1, token_mlp='mix_skip',
                 concat='original'):
        super().__init__()
        if not dil_conv:
            raise NotImplementedError("only dil_conv=1 (dilated 3x3 + 1x1 branches) is built")
        self.Hs = [image_size // 4, image_size // 8, image_size // 16, image_size // 32]
        self.Ws = list(self.Hs)
        dilation = 2
        k1, p1 = [7, 3, 3, 3], [3, 0, 0, 0]
        k2, p2 = [1, 1, 1, 1], [0, 0, 0, 0]
        strides = [4, 2, 2, 2]
        # creation order = the reference's (same-seed initialisation draws the RNG in this order)
        self.conv1_1_s1 = nn.Conv2d(2 * in_dim[0], in_dim[0], 1)
        self.conv1_1_s2 = nn.Conv2d(2 * in_dim[1], in_dim[1], 1)
        self.conv1_1_s3 = nn.Conv2d(2 * in_dim[2], in_dim[2], 1)
        self.conv1_1_s4 = nn.Conv2d(2 * in_dim[3], in_dim[3], 1)
        self.patch_embed1 = OverlapPatchEmbeddings(image_size, 7, 4, 3, 3, in_dim[0])
        for s in (1, 2, 3):
            size = image_size // (2 ** (s + 1))
            setattr(self, 'patch_embed%d_1' % (s + 1),
                    OverlapPatchEmbeddings_fuse(size, k1[s], strides[s], p1[s], dilation, in_dim[s - 1], in_dim[s]))
            setattr(self, 'patch_embed%d_2' % (s + 1),
                    OverlapPatchEmbeddings_fuse(size, k2[s], strides[s], p2[s], dilation, in_dim[s - 1], in_dim[s]))
        self.block1 = nn.ModuleList([EfficientTransformerBlock(in_dim[0], key_dim[0], value_dim[0], head_count, token_mlp)
                                     for _ in range(layers[0])])
        self.norm1 = nn.LayerNorm(in_dim[0])
        for s in (1, 2, 3):
            setattr(self, 'block%d' % (s + 1), nn.ModuleList([
                EfficientTransformerBlockFuse(in_dim[s], key_dim[s], value_dim[s], head_count, token_mlp)
                for _ in range(layers[s])]))
            setattr(self, 'norm%d' % (s + 1), nn.LayerNorm(in_dim[s]))
        self.concat = concat
        self.sk_concat2 = SK_Block(in_dim[1], in_dim[1], num_path=2, reduction=16)
        self.sk_concat3 = SK_Block(in_dim[2], in_dim[2], num_path=2, reduction=16)
        self.sk_concat4 = SK_Block(in_dim[3], in_dim[3], num_path=2, reduction=16)

    def stage(self, x_nhwc, s):
        """One inception stage (reference :455-483) on an NHWC map -> NHWC map [B, H/2, W/2, C]."""
        pe1, pe2 = getattr(self, 'patch_embed%d_1' % s), getattr(self, 'patch_embed%d_2' % s)
        t, H1, W1, H2, W2 = ops.dual_patch_embed(x_nhwc, pe1.slots(), pe2.slots(), pe1.norm.eps)
        if (H2, W2) != (self.Hs[s - 1], self.Ws[s - 1]):
            raise RuntimeError("stage %d: branch-2 map %dx%d differs from the stage size %dx%d"
                               % (s, H2, W2, self.Hs[s - 1], self.Ws[s - 1]))
        for blk in getattr(self, 'block%d' % s):
            t = blk(t, H1 * W1, H2 * W2, H1, W1, H2, W2)
        norm, conv = getattr(self, 'norm%d' % s), getattr(self, 'conv1_1_s%d' % s)
        if self.concat == 'original':
            out = ops.fuse_merge(t, H1, W1, H2, W2, norm.weight, norm.bias, norm.eps, conv.weight, conv.bias)
        else:
            sk = getattr(self, 'sk_concat%d' % s)
            if sk.training:
                raise NotImplementedError("SK_Block's BatchNorm runs with running statistics only: call .eval()")
            out = ops.fuse_merge_sk(t, H1, W1, H2, W2, norm.weight, norm.bias, norm.eps, sk.fc, sk.fcs[0], sk.fcs[1],
                                    sk.conv_bn_ac[0], sk.conv_bn_ac[2])
        return out.view(out.shape[0], H2, W2, -1)

    def nhwc(self, x):
        B = x.shape[0]
        t, H, W = self.patch_embed1(x)
        for blk in self.block1:
            t = blk(t, H, W)
        t = ops.layernorm(t, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        maps = [t.view(B, H, W, -1)]
        for s in (2, 3, 4):
            maps.append(self.stage(maps[-1], s))
        return maps

    def forward(self, x):
        return [m.permute(0, 3, 1, 2) for m in self.nhwc(x)]


class Transception(nn.Module):
    """reference Transception.py:1010-1057.  ``image_size`` is an extension (the reference hard-codes 224)."""

    def __init__(self, num_classes=9, head_count=1, dil_conv=1, token_mlp_mode="mix_skip", concat='original', image_size=224):
        super().__init__()
        if image_size % 32:
            raise ValueError("image_size must be a multiple of 32")
        dims, key_dim, value_dim, layers = [[64, 128, 320, 512], [64, 128, 320, 512], [64, 128, 320, 512], [2, 2, 2, 2]]
        self.backbone = MiT_3inception(image_size=image_size, in_dim=dims, key_dim=key_dim, value_dim=value_dim, layers=layers,
                                       head_count=head_count, dil_conv=dil_conv, token_mlp=token_mlp_mode, concat=concat)
        fs = image_size // 32
        io = [[32, 64, 64, 64], [144, 128, 128, 128], [288, 320, 320, 320], [512, 512, 512, 512]]
        self.decoder_3 = MyDecoderLayer((fs, fs), io[3], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_2 = MyDecoderLayer((fs * 2, fs * 2), io[2], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_1 = MyDecoderLayer((fs * 4, fs * 4), io[1], head_count, token_mlp_mode, n_class=num_classes)
        self.decoder_0 = MyDecoderLayer((fs * 8, fs * 8), io[0], head_count, token_mlp_mode, n_class=num_classes,
                                        is_last=True)

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("transception_b200 builds the forward path only: call under torch.no_grad()")
        ops.require_cuda(x)
        maps = self.backbone.nhwc(x)
        b, _, _, c = maps[3].shape
        t3 = self.decoder_3(maps[3].reshape(b, -1, c))
        t2 = self.decoder_2(t3, maps[2])
        t1 = self.decoder_1(t2, maps[1])
        return self.decoder_0(t1, maps[0])
