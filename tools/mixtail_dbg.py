#!/usr/bin/env python
"""Debug aid: fused vs unfused Mix-FFN tail vs a torch fp32 restatement, with mismatch localisation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from transception_b200 import MSTransception, ops  # noqa: E402


def ref(mod, x, hw):
    B, N, C = x.shape
    h = F.linear(x, mod.fc1.weight, mod.fc1.bias)
    C4 = h.shape[-1]
    img = h.transpose(1, 2).reshape(B, C4, hw, hw)
    d = F.conv2d(img, mod.dwconv.dwconv.weight, mod.dwconv.dwconv.bias, padding=1, groups=C4).flatten(2).transpose(1, 2)
    a = F.gelu(F.layer_norm(d + h, (C4,), mod.norm1.weight, mod.norm1.bias, mod.norm1.eps))
    return F.linear(a, mod.fc2.weight, mod.fc2.bias)


def main():
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    net = MSTransception(num_classes=9).eval().cuda()
    cases = [("backbone.mhca_stage2.mhca_blks.0.MHCA_layers.0.mlp", 28, 64, 16),
             ("backbone.mhca_stage3.mhca_blks.0.MHCA_layers.0.mlp", 14, 128, 16),
             ("bridge.bridge_layer1.mixffn2", 28, 128, 16),
             ("backbone.block1.0.mlp", 56, 64, 16)]
    with torch.no_grad():
        for path, hw, C, B in cases:
            mod = net.get_submodule(path)
            x = torch.randn(B, hw * hw, C, device="cuda")
            want = ref(mod, x, hw)
            ys = []
            for flag in (0, 1, 1):
                ops.set_flag("mixtail", flag)
                ys.append(mod(x, hw, hw).clone())
                torch.cuda.synchronize()
            print("%s B=%d: |unfused-ref| %.3e  |fused-ref| %.3e  |fused-fused| %.3e" % (
                path, B, (ys[0] - want).abs().max().item(), (ys[1] - want).abs().max().item(), (ys[1] - ys[2]).abs().max().item()))
            bad = ((ys[1] - want).abs() > 0.05).reshape(-1, C)
            rows = bad.any(1).nonzero().flatten()
            if rows.numel():
                r = rows.cpu()
                print("   bad rows %d of %d; tiles %s; rows-in-tile %s; bad cols of first row %s" % (
                    r.numel(), bad.shape[0], sorted(set((r // 128).tolist()))[:20], sorted(set((r % 128).tolist()))[:40],
                    bad[r[0]].nonzero().flatten().tolist()[:20]))
        # whole model
        x = torch.rand(16, 1, 224, 224, device="cuda")
        outs = []
        for flag in (0, 1):
            ops.set_flag("mixtail", flag)
            outs.append(net(x).float().clone())
            torch.cuda.synchronize()
        print("whole model bs16: max|fused-unfused| %.3e (max |logit| %.3e)" % ((outs[0] - outs[1]).abs().max().item(), outs[0].abs().max().item()))
        per = (outs[0] - outs[1]).abs().flatten(1).max(1).values
        print("   per image:", ["%.1e" % v for v in per.tolist()])
        for fl in (("fork", 0), ("pdl", 0)):
            ops.set_flag(*fl)
            o = net(x).float()
            print("   with %s=%d: max|fused-unfused| %.3e" % (fl[0], fl[1], (outs[0] - o).abs().max().item()))


if __name__ == "__main__":
    main()
