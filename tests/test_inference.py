"""Batched volume inference: ``transception_b200.inference`` against the reference's slice-by-slice
loop (/root/reference/utils.py:63-98) restated here with the same model."""
import numpy as np
import pytest
import torch

from transception_b200 import inference as INF


def _reference_loop(image, net, patch_size, device="cuda"):
    """utils.py:65-92 verbatim in behaviour: per slice cubic zoom, ToTensor + Normalize, bs-1 forward,
    argmax(softmax), nearest zoom back.  Pinned to the real utils.test_single_volume by
    tests/test_reference_callers.py::test_reference_loop_restatement_is_pinned_to_utils."""
    from scipy.ndimage import zoom
    from torchvision import transforms
    prediction = np.zeros(image.shape, dtype=np.uint8)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
    for ind in range(image.shape[0]):
        sl = image[ind]
        x, y = sl.shape
        if x != patch_size[0] or y != patch_size[1]:
            sl = zoom(sl, (patch_size[0] / x, patch_size[1] / y), order=3)
        inp = tf(sl).unsqueeze(0).float().to(device)
        with torch.no_grad():
            out = torch.argmax(torch.softmax(net(inp), dim=1), dim=1).squeeze(0).cpu().numpy()
        prediction[ind] = zoom(out, (x / patch_size[0], y / patch_size[1]), order=0) if (x, y) != tuple(patch_size) else out
    return prediction


def test_batches_cover_the_volume():
    assert INF.batches(0, 16) == []
    assert INF.batches(5, 16) == [(0, 5)]
    assert INF.batches(33, 16) == [(0, 16), (16, 32), (32, 33)]


@pytest.mark.parametrize("shape", [(224, 224), (96, 128)])
def test_slice_preparation_matches_the_reference_transform(shape):
    from scipy.ndimage import zoom
    from torchvision import transforms
    sl = np.random.default_rng(0).random(shape, dtype=np.float32)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
    want = sl if shape == (224, 224) else zoom(sl, (224 / shape[0], 224 / shape[1]), order=3)
    want = tf(want).float()[0].numpy()
    got = INF._prepare_slice(sl, (224, 224))
    assert got.dtype == np.float32 and np.array_equal(got, want)
    lab = np.random.default_rng(1).integers(0, 9, (224, 224)).astype(np.uint8)
    back = INF._restore_slice(lab, shape, (224, 224))
    assert back.shape == shape


def test_no_cpu_fallback():
    from transception_b200 import MSTransception
    net = MSTransception(num_classes=9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        INF.predict_volume(np.zeros((2, 224, 224), np.float32), net)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,batch", [((19, 256, 256), 16), ((7, 224, 224), 3)])
def test_predict_volume_equals_reference_loop(cuda_lib, shape, batch):
    from transception_b200 import MSTransception
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9).eval().cuda()
    vol = np.random.default_rng(5).random(shape, dtype=np.float32)
    want = _reference_loop(vol, net, (224, 224))
    got = INF.predict_volume(vol, net, (224, 224), batch=batch)
    assert got.shape == want.shape
    assert np.array_equal(got, want), "mismatching voxels: %d" % int((got != want).sum())
    # the drop-in wrapper with an injected metric (medpy is the reference's choice and is optional here)
    label = np.random.default_rng(6).integers(0, 9, shape).astype(np.float32)
    dice = lambda p, g: 2.0 * (p & g).sum() / max(1, p.sum() + g.sum())
    m = INF.test_single_volume(torch.from_numpy(vol)[None], torch.from_numpy(label)[None], net, classes=9,
                               patch_size=[224, 224], batch=batch, metric_fn=dice)
    assert len(m) == 8 and all(0.0 <= v <= 1.0 for v in m)


@pytest.mark.gpu
def test_argmax_classes(cuda_lib):
    from transception_b200 import ops
    x = torch.randn(3, 9, 37, 41, device="cuda")
    x[0, 2, 0, 0] = x[0, 5, 0, 0] = 50.0           # tie -> first index
    got = ops.argmax_classes(x)
    assert got.dtype == torch.uint8 and torch.equal(got.long(), x.argmax(1)) and got[0, 0, 0].item() == 2
