"""The reference's own callers against the drop-in (SURVEY.md §8b last row): ``trainer.trainer_synapse`` (trainer.py:72-230) and
``utils.test_single_volume`` (utils.py:63-110) are imported UNMODIFIED from /root/reference through a stub harness (tensorboardX,
medpy, SimpleITK, matplotlib, the Synapse dataset; ``Tensor.cuda`` is the identity: this box has no GPU) and executed with the
drop-in ``networks.MSTr.MSTransception``.  The product has no CPU path, so the one thing replaced is the arithmetic of
``forward`` (a test-local stand-in that keeps the signature and the logits shape); everything the callers do TO the model —
``train()/eval()``, ``model(image_batch)``, ``model.parameters()`` into ``optim.SGD``, the per-iteration ``param_group['lr']``
writes, ``state_dict()`` checkpoints, the slice loop of ``test_single_volume`` — runs against the drop-in's surface.
Skipped where /root/reference does not exist (the GPU box)."""
import argparse
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("TRANSCEPTION_REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "trainer.py")), reason="reference checkout not present")


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value.detach()) if torch.is_tensor(value) else float(value), step))

    def add_image(self, *a, **k):
        pass

    def close(self):
        pass


class _Synapse(torch.utils.data.Dataset):
    """Stand-in for datasets.dataset_synapse.Synapse_dataset: same constructor keywords, same sample dictionaries."""

    def __init__(self, base_dir=None, list_dir=None, split="train", img_size=224, norm_x_transform=None, norm_y_transform=None):
        self.split, self.size = split, img_size
        self.rng = np.random.default_rng(0)

    def __len__(self):
        return 4 if self.split == "train" else 1

    def __getitem__(self, i):
        if self.split == "train":
            return {"image": torch.rand(1, self.size, self.size), "label": torch.randint(0, 9, (1, self.size, self.size)).float(),
                    "case_name": "train%d" % i}
        return {"image": torch.rand(3, self.size, self.size), "label": torch.randint(0, 9, (3, self.size, self.size)).float(),
                "case_name": "vol%d" % i}


@pytest.fixture()
def harness(monkeypatch, tmp_path):
    recorded = {"dc": []}
    stubs = {}
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = _Writer
    stubs["tensorboardX"] = tb
    medpy = types.ModuleType("medpy")
    metric = types.ModuleType("medpy.metric")
    binary = types.ModuleType("medpy.metric.binary")

    def dc(pred, gt):
        recorded["dc"].append(np.array(pred, copy=True))
        inter = float(np.logical_and(pred, gt).sum())
        return 2.0 * inter / float(pred.sum() + gt.sum())
    binary.dc, binary.hd95 = dc, (lambda pred, gt: 0.0)
    metric.binary = binary
    medpy.metric = metric
    stubs.update({"medpy": medpy, "medpy.metric": metric, "medpy.metric.binary": binary})
    sitk = types.ModuleType("SimpleITK")

    class _Img:
        def SetSpacing(self, *_):
            pass
    sitk.GetImageFromArray = lambda a: _Img()
    sitk.WriteImage = lambda img, path: recorded.setdefault("written", []).append(path)
    stubs["SimpleITK"] = sitk
    ds_pkg = types.ModuleType("datasets")
    ds_mod = types.ModuleType("datasets.dataset_synapse")
    ds_mod.Synapse_dataset, ds_mod.RandomGenerator = _Synapse, object
    ds_pkg.dataset_synapse = ds_mod
    stubs.update({"datasets": ds_pkg, "datasets.dataset_synapse": ds_mod})
    have_mpl = importlib.util.find_spec("matplotlib") is not None
    if not have_mpl:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        for name in ("figure", "title", "savefig", "plot"):
            setattr(plt, name, lambda *a, **k: None)
        mpl.pyplot = plt
        stubs.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    for k, v in stubs.items():
        monkeypatch.setitem(sys.modules, k, v)
    # import order of an integrated checkout: the drop-in's networks/ package shadows the reference's, everything else
    # (trainer.py, utils.py) is the reference's own file
    monkeypatch.syspath_prepend(REF)
    monkeypatch.syspath_prepend(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for name in [m for m in sys.modules if m in ("utils", "trainer")]:
        sys.modules.pop(name, None)
    import networks.MSTr as dropin
    assert "transception_b200" in dropin.MSTransception.__module__, "networks.MSTr must resolve to the drop-in"
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    utils = importlib.import_module("utils")
    trainer = importlib.import_module("trainer")
    assert os.path.realpath(utils.__file__).startswith(os.path.realpath(REF))
    monkeypatch.setattr(trainer, "plot_result", lambda *a, **k: None)          # reporting helper (pandas + matplotlib): not a model call
    yield types.SimpleNamespace(utils=utils, trainer=trainer, recorded=recorded, tmp=str(tmp_path))
    for name in ("utils", "trainer"):
        sys.modules.pop(name, None)


def _dropin_with_cpu_forward():
    """The drop-in class with ONLY its arithmetic replaced (no CPU path exists in the product): same constructor, parameters,
    state_dict and forward signature; logits [B, classes, H, W] depending on one real parameter so that backward flows."""
    from networks.MSTr import MSTransception
    calls = []

    class Shim(MSTransception):
        def forward(self, x):
            calls.append((tuple(x.shape), self.training))
            if x.size(1) == 1:
                x = x.repeat(1, 3, 1, 1)                      # MSTr.py:2828-2829
            w = self.decoder_0.last_layer.weight              # [classes, 64, 1, 1]
            return x.mean(1, keepdim=True) * w.mean(1).view(1, -1, 1, 1) + self.decoder_0.last_layer.bias.view(1, -1, 1, 1)
    torch.manual_seed(1234)
    return Shim(num_classes=9), calls


def test_trainer_synapse_runs_against_the_dropin(harness):
    model, calls = _dropin_with_cpu_forward()
    before = model.decoder_0.last_layer.bias.detach().clone()
    args = argparse.Namespace(grad_clipping=False, use_scheduler=False, base_lr=0.05, num_classes=9, batch_size=2, n_gpu=1,
                              root_path="", test_path="", list_dir="", img_size=224, num_workers=0, seed=1234, max_epochs=1,
                              eval_interval=1, model_name="dropin", z_spacing=1)
    out = harness.trainer.trainer_synapse(args, model, harness.tmp)
    assert out == "Training Finished!"
    train_calls = [c for c in calls if c[1]]
    assert train_calls == [((2, 1, 224, 224), True)] * 2                                  # 4 samples / batch 2, train mode
    assert [c for c in calls if not c[1]] == [((1, 1, 224, 224), False)] * 3              # test_single_volume: 3 slices, eval mode
    assert not torch.equal(before, model.decoder_0.last_layer.bias.detach())              # optim.SGD(model.parameters()) stepped it
    ckpt = torch.load(os.path.join(harness.tmp, "dropin_epoch_0.pth"))
    from networks.MSTr import MSTransception
    fresh = MSTransception(num_classes=9)
    fresh.load_state_dict(ckpt, strict=True)                                              # the checkpoint the caller wrote
    assert len(ckpt) == 2200


def test_reference_loop_restatement_is_pinned_to_utils(harness):
    """tests/test_inference.py compares the batched predictor with a test-local restatement of utils.py:65-92; this pins that
    restatement to the real function: same stand-in network, same volume, the label map recovered from the masks the real
    function hands to medpy."""
    spec = importlib.util.spec_from_file_location("_test_inference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_inference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _reference_loop = mod._reference_loop
    torch.manual_seed(0)
    net = torch.nn.Conv2d(1, 9, 3, padding=1).eval()
    vol = np.random.default_rng(5).random((4, 96, 128), dtype=np.float32)
    label = np.random.default_rng(6).integers(0, 9, vol.shape).astype(np.float32)
    harness.utils.test_single_volume(torch.from_numpy(vol)[None], torch.from_numpy(label)[None], net, classes=9, patch_size=[224, 224])
    masks = harness.recorded["dc"]
    real = np.zeros(vol.shape, dtype=np.uint8)
    seen = 0
    for i in range(1, 9):                 # calculate_metric_percase calls medpy only when both masks are non-empty
        want_mask = label == i
        if seen < len(masks) and want_mask.sum() > 0:
            real[masks[seen] > 0] = i
            seen += 1
    mine = _reference_loop(vol, net, (224, 224), device="cpu")
    assert seen == len(masks) and seen >= 6
    covered = np.isin(mine, [i for i in range(1, 9)])
    assert np.array_equal(np.where(covered, mine, 0), real)
