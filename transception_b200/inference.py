"""Volume inference of the reference (``utils.test_single_volume``, /root/reference/utils.py:63-110; callers
``trainer.py:25-47``, ``test.py:104-123``) with the slices of a volume batched through the forward.

The reference runs one bs-1 forward per slice and ships the full logits to the host.  Here the slices are resized on the
host exactly as the reference does (``scipy.ndimage.zoom``, order 3 in / order 0 out — host-side, outside the hot path),
stacked into batches of ``batch`` slices, pushed through the model on the GPU, reduced to a uint8 label map on the device
(``ops.argmax_classes``) and only that map crosses PCIe.  Per slice the result is identical to the reference loop run with
the same model: the kernels are batch-invariant (no cross-image reduction) and softmax does not change the arg max.
"""
import numpy as np
import torch

from . import ops


def _prepare_slice(slice_, patch_size):
    """utils.py:69-77: cubic zoom to the patch size, ToTensor + Normalize([0.5], [0.5]), float32."""
    from scipy.ndimage import zoom
    x, y = slice_.shape[0], slice_.shape[1]
    if x != patch_size[0] or y != patch_size[1]:
        slice_ = zoom(slice_, (patch_size[0] / x, patch_size[1] / y), order=3)
    return ((np.asarray(slice_) - 0.5) / 0.5).astype(np.float32)


def _restore_slice(out, shape, patch_size):
    """utils.py:88-91: nearest zoom of the label map back to the slice size."""
    from scipy.ndimage import zoom
    x, y = shape
    if x != patch_size[0] or y != patch_size[1]:
        return zoom(out, (x / patch_size[0], y / patch_size[1]), order=0)
    return out


def batches(n, batch):
    """[(start, stop)] covering range(n) in chunks of ``batch`` (the last one ragged)."""
    return [(s, min(n, s + batch)) for s in range(0, n, batch)]


def predict_volume(image, net, patch_size=(224, 224), batch=16, device=None):
    """image: [D, H, W] (or [H, W]) array / tensor -> integer label volume of the same shape (numpy, image dtype of the
    reference's ``prediction = np.zeros_like(label)`` is applied by the caller)."""
    image = image.squeeze(0).cpu().detach().numpy() if isinstance(image, torch.Tensor) else np.asarray(image)
    single = image.ndim == 2
    vol = image[None] if single else image
    dev = torch.device(device) if device is not None else next(net.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("transception_b200 kernels run on sm_100a CUDA tensors only (got device %s); there is no CPU fallback" % dev)
    net.eval()
    D = vol.shape[0]
    pred = np.zeros(vol.shape, dtype=np.uint8)
    if single:       # utils.py:93-98: a 2-D input goes through as it is (no zoom, no normalisation)
        with torch.no_grad():
            x = torch.from_numpy(vol).unsqueeze(1).float().to(dev)
            return ops.argmax_classes(net(x).float()).cpu().numpy()[0]
    host = torch.empty((min(batch, D), 1, patch_size[0], patch_size[1]), dtype=torch.float32).pin_memory()
    with torch.no_grad():
        for s, e in batches(D, batch):
            for i in range(s, e):
                host[i - s, 0] = torch.from_numpy(_prepare_slice(vol[i], patch_size))
            x = host[:e - s].to(dev, non_blocking=True)
            lab = ops.argmax_classes(net(x).float()).cpu().numpy()
            for i in range(s, e):
                pred[i] = _restore_slice(lab[i - s], vol[i].shape, patch_size)
    return pred


def test_single_volume(image, label, net, classes, patch_size=[256, 256], test_save_path=None, case=None, z_spacing=1,
                       batch=16, metric_fn=None):
    """Drop-in for ``utils.test_single_volume`` (same positional arguments and return value).  ``metric_fn(pred_mask, gt_mask)``
    defaults to the reference's ``calculate_metric_percase`` (utils.py:50-60), which needs ``medpy``."""
    image_np = image.squeeze(0).cpu().detach().numpy() if isinstance(image, torch.Tensor) else np.asarray(image)
    label_np = label.squeeze(0).cpu().detach().numpy() if isinstance(label, torch.Tensor) else np.asarray(label)
    prediction = predict_volume(image_np, net, patch_size, batch).astype(label_np.dtype)
    if metric_fn is None:
        metric_fn = calculate_metric_percase
    metric_list = [metric_fn(prediction == i, label_np == i) for i in range(1, classes)]
    if test_save_path is not None:
        import SimpleITK as sitk      # utils.py:100-109
        for arr, tag in ((image_np, "img"), (prediction, "pred"), (label_np, "gt")):
            itk = sitk.GetImageFromArray(arr.astype(np.float32))
            itk.SetSpacing((1, 1, z_spacing))
            sitk.WriteImage(itk, test_save_path + '/' + case + "_" + tag + ".nii.gz")
    return metric_list


test_single_volume.__test__ = False      # not a pytest test


def calculate_metric_percase(pred, gt):
    """utils.py:50-60 (Dice and HD95 through medpy, as the reference)."""
    from medpy import metric
    pred, gt = pred.copy(), gt.copy()
    pred[pred > 0] = 1
    gt[gt > 0] = 1
    if pred.sum() > 0 and gt.sum() > 0:
        return metric.binary.dc(pred, gt), metric.binary.hd95(pred, gt)
    elif pred.sum() > 0 and gt.sum() == 0:
        return 1, 0
    return 0, 0
