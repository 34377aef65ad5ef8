"""Forward+backward timing of the training-row blocks built so far (SURVEY §8d config 3 shapes, bs16) on one B200:
the drop-in module in autograd mode (library kernels forward and backward) next to the same arithmetic in eager PyTorch
on the same GPU (the oracle restatement under torch autograd, fp32 / TF32 off) — CUDA events, 3 warm-up + 10 timed.

Usage: python tools/train_bench.py [--profile]   (--profile: one iteration of each block between cudaProfilerStart/Stop)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mstr_oracle as O  # noqa: E402  (eager-GPU comparison leg only)


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


def case(name, module, call_args, oracle_fn, x, dy):
    from transception_b200 import ops
    mg = module.cuda().train()
    xg = x.cuda().requires_grad_()
    dyg = dy.cuda()

    def ours():
        for p in mg.parameters():
            p.grad = None
        xg.grad = None
        mg(xg, *call_args).backward(dyg)

    sd = {"m." + k: v.detach().clone().requires_grad_() for k, v in mg.state_dict().items()}
    xe = x.cuda().requires_grad_()

    def eager():
        for v in sd.values():
            v.grad = None
        xe.grad = None
        oracle_fn(sd, xe).backward(dyg)

    n0 = ops.launches()
    ours()
    n_launch = ops.launches() - n0
    t_ours = timeit(ours)
    t_eager = timeit(eager)
    with torch.no_grad():
        t_fwd = timeit(lambda: mg(xg, *call_args))

    def graphed(fn):
        """whole forward+backward captured once and replayed (PyTorch's whole-network capture recipe)"""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return timeit(g.replay)

    t_ours_g = graphed(ours)
    t_eager_g = graphed(eager)
    print("%-36s fwd+bwd: ours %7.3f ms (%3d kernels), graph replay %7.3f ms | eager PyTorch %7.3f ms, graph replay %7.3f ms "
          "| x%.2f (graph x%.2f) | inference fwd %6.3f ms"
          % (name, t_ours, n_launch, t_ours_g, t_eager, t_eager_g, t_eager / t_ours, t_eager_g / t_ours_g, t_fwd), flush=True)
    return ours


def main():
    from networks.MSTr import EfficientTransformerBlock, MHCAEncoder, MixFFN_skip, MyDecoderLayer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    B = 16
    runs = []
    torch.manual_seed(1)
    x = torch.randn(B, 3136, 64, generator=g)
    dy = torch.randn(B, 3136, 64, generator=g) * 1e-3
    runs.append(case("MixFFN_skip 56x56 C64", MixFFN_skip(64, 256), (56, 56),
                     lambda sd, xr: O.mixffn_skip(sd, "m", xr, 56, 56), x, dy))
    runs.append(case("EfficientTransformerBlock 56x56 C64", EfficientTransformerBlock(64, 64, 64, 1, "mix_skip"), (56, 56),
                     lambda sd, xr: O.efficient_block(sd, "m", xr, 56, 56), x, dy))
    x = torch.randn(B, 784, 64, generator=g)
    dy = torch.randn(B, 64, 28, 28, generator=g) * 1e-3
    runs.append(case("MHCAEncoder 28x28 C64 L3", MHCAEncoder(64, 3, 8, 4, [0.0] * 3), ((28, 28),),
                     lambda sd, xr: O.mhca_encoder(sd, "m", xr, 28, 28, 3), x, dy))
    x = torch.randn(B, 196, 128, generator=g)
    dy = torch.randn(B, 128, 14, 14, generator=g) * 1e-3
    runs.append(case("MHCAEncoder 14x14 C128 L8", MHCAEncoder(128, 8, 8, 4, [0.0] * 8), ((14, 14),),
                     lambda sd, xr: O.mhca_encoder(sd, "m", xr, 14, 14, 8), x, dy))
    x = torch.randn(B, 49, 320, generator=g)
    dy = torch.randn(B, 320, 7, 7, generator=g) * 1e-3
    runs.append(case("MHCAEncoder 7x7 C320 L3", MHCAEncoder(320, 3, 8, 4, [0.0] * 3), ((7, 7),),
                     lambda sd, xr: O.mhca_encoder(sd, "m", xr, 7, 7, 3), x, dy))
    x = torch.randn(B, 784, 128, generator=g)
    dy = torch.randn(B, 3136, 64, generator=g) * 1e-3
    dec = MyDecoderLayer((28, 28), [144, 128, 128, 128], 1, "mix_skip", n_class=9)
    x2 = torch.randn(B, 28, 28, 160, generator=g).cuda()
    runs.append(case("MyDecoderLayer 28x28 (decoder_1)", dec, (x2,),
                     lambda sd, xr: O.decoder_layer(sd, "m", xr, x2), x, dy))
    if "--profile" in sys.argv:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        for r in runs:
            r()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
