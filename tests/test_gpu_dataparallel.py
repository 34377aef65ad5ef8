"""nn.DataParallel over the drop-in (trainer.py:110-111 wraps the model when n_gpu > 1): replicate(), one Python thread per
device, per-device auxiliary streams and prepared fp16 weight copies.  Needs two GPUs; skipped otherwise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_dataparallel_inference_equals_single_gpu(cuda_lib):
    _two_gpus()
    from networks.MSTr import MSTransception
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).eval().cuda(0)
    x = (torch.rand(4, 1, 224, 224, generator=torch.Generator().manual_seed(0)) * 2 - 1).cuda(0)
    with torch.no_grad():
        want = net(x)
        dp = torch.nn.DataParallel(net, device_ids=[0, 1])
        got = dp(x)
        again = dp(x)
    assert got.device == want.device and got.shape == want.shape
    assert torch.equal(got, again)
    assert torch.equal(got, want), "max-abs %.3e" % (got - want).abs().max().item()


def test_dataparallel_train_step_equals_sequential_replicas(cuda_lib):
    """One training forward + backward under DataParallel (each replica normalises with the BatchNorm statistics of ITS half of
    the batch, gradients are summed onto device 0) against the same two halves run one after the other on one GPU."""
    _two_gpus()
    from networks.MSTr import MSTransception
    from transception_b200.losses import CeDiceLoss
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(4, 1, 224, 224, generator=gen) * 2 - 1).cuda(0)
    labels = torch.randint(0, 9, (4, 224, 224), generator=gen).cuda(0)
    crit = CeDiceLoss(9)

    def make():
        torch.manual_seed(1234)
        return MSTransception(num_classes=9).train().cuda(0)
    a = make()
    loss_a = crit(torch.nn.DataParallel(a, device_ids=[0, 1])(x), labels)
    loss_a.backward()
    b = make()
    loss_b = crit(torch.cat([b(x[:2]), b(x[2:])]), labels)
    loss_b.backward()
    assert abs(loss_a.item() - loss_b.item()) <= 1e-6 * abs(loss_b.item())
    n = 0
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if q.grad is None:
            # a dead parameter: nn.DataParallel's Broadcast node hands back a zero gradient where a single module leaves None
            assert p.grad is None or not p.grad.any(), k
        else:
            assert p.grad is not None, k
            den = q.grad.norm().item()
            assert (p.grad - q.grad).norm().item() <= 1e-5 * den + 1e-12, k
            n += 1
    assert n > 1000
