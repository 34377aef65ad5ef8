"""ctypes binding of ``libtransception_sm100.so`` (C ABI: ``include/transception_sm100.h``).

PyTorch is used for device memory (outputs and scratch come from the caching allocator) and for the current
stream; every FLOP of the hot path is executed by the library's kernels.  There is no fallback: if the shared
library is missing, or the tensor is not on an sm_100 CUDA device, the call raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtransception_sm100.so")
MHCA_NP = 24

_lib = None
_lock = threading.Lock()
# Programmatic dependent launch on the CUDA-core kernels of the training row (library flag "pdl_chain", csrc/common.cuh
# tcx_launch_chain): 1 = those kernels carry the attribute like the tensor-core pipeline kernels do.  The environment
# variable TCX_PDL_CHAIN overrides it at load time (measurement aid: A/B runs of bench.py).
PDL_CHAIN = 0
_launches = 0          # C-ABI calls issued (each enqueues >= 1 kernel); see kernel_launch_estimate in bench.py

_vp, _i, _f, _ll, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t
_pp = ctypes.POINTER(ctypes.c_void_p)

_PROTOS = {
    "tcx_version": (ctypes.c_char_p, []),
    "tcx_last_error": (ctypes.c_char_p, []),
    "tcx_device_ok": (_i, []),
    "tcx_set_flag": (_i, [ctypes.c_char_p, _i]),
    "tcx_launch_count": (_ll, []),
    "tcx_profile_enable": (_i, [ctypes.c_char_p]),
    "tcx_profile_read": (_i, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i)]),
    "tcx_profile_read_work": (_i, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_double)]),
    "tcx_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    "tcx_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_f32_to_f16": (_i, [_vp, _vp, _ll, _vp]),
    "tcx_linear_f16_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_prepare_weight_f16": (_i, [_vp, _vp, _ll, _vp]),
    "tcx_prepare_conv_weight_f16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "tcx_forget_weight": (_i, [_vp]),
    "tcx_eff_block_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_eff_block_fwd": (_i, [_vp, _pp, _f, _f, _vp, _i, _i, _i, _i, _vp, _vp]),
    "tcx_fuse_eff_attn_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_fuse_eff_attn_fwd": (_i, [_vp, _pp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_fuse_block_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_fuse_block_fwd": (_i, [_vp, _pp, _f, _f, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_dual_patch_embed_workspace_bytes": (_sz, [_i] * 11),
    "tcx_dual_patch_embed_fwd": (_i, [_vp, _pp, _f, _vp] + [_i] * 11 + [_vp, _vp]),
    "tcx_fuse_merge_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_fuse_merge_fwd": (_i, [_vp, _pp, _f, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_fuse_merge_sk_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_fuse_merge_sk_fwd": (_i, [_vp, _pp, _f, _f, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_seg_loss_workspace_bytes": (_sz, [_i, _i, _ll]),
    "tcx_seg_loss_fwd": (_i, [_vp, _vp, _i, _i, _i, _ll, _i, _f, _f, ctypes.POINTER(_f), _vp, _vp, _vp]),
    "tcx_seg_loss_bwd": (_i, [_vp, _vp, _i, _i, _i, _ll, _i, _f, _f, ctypes.POINTER(_f), _vp, _vp, _vp, _vp]),
    "tcx_argmax_classes_fwd": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
    "tcx_bridge_layer_workspace_bytes": (_sz, [_i, _i]),
    "tcx_bridge_layer_fwd": (_i, [_vp, _pp, _i, _f, _f, _vp, _i, _i, _vp, _vp]),
    "tcx_bridge_block_workspace_bytes": (_sz, [_i, _i]),
    "tcx_bridge_block_fwd": (_i, [_vp, _pp, ctypes.POINTER(_i), _i, _f, _f, _vp, _i, _i, _vp, _vp]),
    "tcx_linear_bn_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _i, _i, _i, _vp]),
    "tcx_patch_embed_ln_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    "tcx_dwconv_tokens_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "tcx_eff_attn_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_eff_attn_fwd": (_i, [_vp, _pp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "tcx_mixffn_skip_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_mixffn_skip_fwd": (_i, [_vp, _pp, _f, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_mb_factor_attn_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_mb_factor_attn_fwd": (_i, [_vp, _pp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_mhca_blocks_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_mhca_blocks_fwd": (_i, [_vp, _vp, _pp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp]),
    "tcx_ripm_dwsep_bn_hs_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "tcx_ripm_dwsep_bn_hs_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_resblock_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_resblock_fwd": (_i, [_vp, _pp, _f, _vp, _i, _i, _i, _i, _vp, _vp]),
    "tcx_iff_coordatt_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_iff_coordatt_fwd": (_i, [_pp, _pp, _f, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_bridge_regroup_fwd": (_i, [_pp, _vp, _i, _i, _vp]),
    "tcx_sum_tensors": (_i, [_pp, _i, _ll, _vp, _vp]),
    "tcx_patch_im2row_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "tcx_final_head_train_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "tcx_final_head_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_final_head_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_bridge_split_fwd": (_i, [_vp, _pp, _i, _i, _vp]),
    "tcx_bridge_merge_fwd": (_i, [_pp, _vp, _vp, _i, _i, _vp]),
    "tcx_scale_reduce_saved_bytes": (_sz, [_i, _i]),
    "tcx_scale_reduce_train_workspace_bytes": (_sz, [_i, _i]),
    "tcx_scale_reduce_train_fwd": (_i, [_vp, _pp, _vp, _i, _i, _vp, _vp, _vp]),
    "tcx_scale_reduce_bwd_workspace_bytes": (_sz, [_i, _i]),
    "tcx_scale_reduce_bwd": (_i, [_vp, _pp, _vp, _vp, _pp, _i, _i, _vp, _vp]),
    "tcx_scale_reduce_workspace_bytes": (_sz, [_i, _i]),
    "tcx_scale_reduce_fwd": (_i, [_vp, _pp, _f, _vp, _i, _i, _vp, _vp]),
    "tcx_bridge_sr_attn_workspace_bytes": (_sz, [_i, _i]),
    "tcx_bridge_sr_attn_fwd": (_i, [_vp, _pp, _f, _f, _vp, _vp, _i, _i, _vp, _vp]),
    "tcx_flash_attn_workspace_bytes": (_sz, [_i, _i]),
    "tcx_flash_attn_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "tcx_flash_attn_f16_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "tcx_bridge_mixffn_workspace_bytes": (_sz, [_i, _i]),
    "tcx_bridge_mixffn_fwd": (_i, [_vp, _vp, _pp, _f, _vp, _i, _i, _vp, _vp]),
    "tcx_concat_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp]),
    "tcx_patch_expand_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "tcx_patch_expand_fwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_final_expand_head_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_final_expand_head_fwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_layernorm_bwd_workspace_bytes": (_sz, [_ll, _i]),
    "tcx_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _ll, _i, _vp, _vp]),
    "tcx_linear_bwd_workspace_bytes": (_sz, [_ll, _i, _i]),
    "tcx_linear_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp]),
    "tcx_wgrad_mn_workspace_bytes": (_sz, [_ll, _i, _i, _i, _i]),
    "tcx_wgrad_mn": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp, _vp]),
    "tcx_mt_chunk": (_i, []),
    "tcx_mt_gather": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "tcx_mt_sqnorm": (_i, [_vp, _vp, _vp, _i, _vp, _f, _vp, _vp]),
    "tcx_mt_sgd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _f, _vp]),
    "tcx_eff_attn_saved_bytes": (_sz, [_i, _i, _i]),
    "tcx_eff_attn_train_fwd": (_i, [_vp, _pp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_eff_attn_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_eff_attn_bwd": (_i, [_vp, _pp, _vp, _vp, _pp, _i, _i, _i, _vp, _vp]),
    "tcx_mb_factor_attn_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_mb_factor_attn_bwd": (_i, [_vp, _vp, _pp, _vp, _i, _vp, _pp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_mb_factor_attn_saved_bytes": (_sz, [_i, _i, _i]),
    "tcx_mb_factor_attn_train_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_mb_factor_attn_train_fwd": (_i, [_vp, _pp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tcx_layernorm_dual_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    "tcx_dwconv_tokens_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_dwconv_tokens_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_attn_core_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_attn_core_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_flash_attn_train_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "tcx_flash_attn_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_flash_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_ea_core_workspace_bytes": (_sz, [_i, _i, _i]),
    "tcx_ea_core_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_ea_core_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tcx_bn_act_train_workspace_bytes": (_sz, [_ll, _i]),
    "tcx_bn_act_train_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _ll, _i, _vp, _vp]),
    "tcx_bn_act_train_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _ll, _i, _vp, _vp]),
    "tcx_dwconv3x3_nhwc_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "tcx_dwconv3x3_nhwc_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "tcx_dwconv3x3_nhwc_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tcx_coord_pool_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_coord_pool_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_coord_gate_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_coord_gate_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "tcx_mixffn_skip_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_mixffn_skip_train_fwd": (_i, [_vp, _pp, _f, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tcx_mixffn_skip_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tcx_mixffn_skip_bwd": (_i, [_vp, _pp, _f, _vp, _vp, _vp, _pp, _i, _i, _i, _i, _i, _vp, _vp]),
}
EXPORTS = tuple(_PROTOS)


def load_library():
    """dlopen the in-tree shared library and declare every prototype. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "transception_b200: %s is missing — build it with `python -m transception_b200.build` "
                    "(there is no CPU/PyTorch fallback)" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _PROTOS.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            lib.tcx_set_flag(b"pdl_chain", int(os.environ.get("TCX_PDL_CHAIN", PDL_CHAIN)))
            _lib = lib
    return _lib


def launches():
    """Kernels enqueued by the library so far (counted inside the library, one per launch)."""
    return int(load_library().tcx_launch_count())


def profile_enable(kernel_name):
    load_library().tcx_profile_enable((kernel_name or "").encode())


def profile_read():
    ms, n = ctypes.c_double(0), ctypes.c_int(0)
    _chk(load_library().tcx_profile_read(ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


def profile_read_work():
    """(total ms, launches, summed algorithmic bytes-or-FLOPs) of the kernel selected by profile_enable."""
    ms, n, w = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_double(0)
    _chk(load_library().tcx_profile_read_work(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)))
    return ms.value, n.value, w.value


def set_flag(name, value):
    return load_library().tcx_set_flag(name.encode(), int(value))


def require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("transception_b200 kernels run on sm_100a CUDA tensors only (got device %s); "
                           "there is no CPU fallback" % x.device)


def _chk(rc):
    global _launches
    _launches += 1
    if rc != 0:
        raise RuntimeError("libtransception_sm100: " + load_library().tcx_last_error().decode())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr16(t):
    if t is None:
        return None
    if t.dtype != torch.float16 or not t.is_contiguous() or not t.is_cuda:
        raise RuntimeError("transception_b200: expected a contiguous fp16 CUDA tensor, got %s %s on %s"
                           % (tuple(t.shape), t.dtype, t.device))
    return t.data_ptr()


_KEEP = __import__("collections").deque(maxlen=512)


def _aligned(t):
    """The kernels read parameters with 16-byte accesses / TMA.  nn.DataParallel's replicas are views into coalesced broadcast
    buffers at arbitrary 4-byte offsets: such a tensor is copied to a fresh (256-byte aligned) allocation for the call.  The
    copy has to stay alive until the kernels that read it are ENQUEUED (afterwards the caching allocator re-uses the block in
    stream order; the library joins its auxiliary streams before a call returns) — a clone dropped before the launch would hand
    its block to the next clone.  _KEEP holds the last few hundred of them, far more than one call takes."""
    if t is None or t.data_ptr() % 16 == 0:
        return t
    c = t.detach().clone()
    _KEEP.append(c)
    return c


def _ptr(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise RuntimeError("transception_b200: expected a contiguous fp32 CUDA tensor, got %s %s contiguous=%s on %s"
                           % (tuple(t.shape), t.dtype, t.is_contiguous(), t.device))
    return _aligned(t).data_ptr()


# ---- prepared (fp16) GEMM weights ----------------------------------------------------------------
# id(tensor) -> (weakref, version, data_ptr, fp16 copy).  A weight is (re)prepared when it is first seen, when its
# version counter moved (in-place update, load_state_dict) or when its storage moved (.cuda(), .to()).
import weakref

_prepared = {}
USE_F16 = True
_raw_gen = 0          # bumped whenever weights are updated through raw pointers (FusedSGD / captured train-step graphs)


class _Prep:
    """Registry entry of one prepared weight: weak reference, the version / storage / raw-update generation the fp16 copy was
    made from, the copy, the patchify-conv permutation (or None), and whether an optimizer keeps the copy current itself."""
    __slots__ = ("ref", "version", "ptr", "w16", "conv", "gen", "tracked")

    def current(self, w):
        return (self.ref() is w and self.version == w._version and self.ptr == w.data_ptr() and
                (self.tracked or self.gen == _raw_gen))


def bump_raw_generation():
    """Weights were just updated in place through raw pointers (no ``_version`` bump): prepared copies that no optimizer table
    keeps current are converted again by the next forward that uses them."""
    global _raw_gen
    _raw_gen += 1


def prepare_weight(w, conv=None):
    """Register an fp16 copy of a GEMM weight matrix with the library (include/transception_sm100.h).
    ``conv=(N, Cin, r)``: a strided patchify conv weight, stored with its K axis permuted to (ky, kx, cin)."""
    if not USE_F16 or w is None or not w.is_cuda:
        return
    base = w._base        # a full-size view (e.g. conv1x1 weight .reshape(C, C)) is tracked through its parameter
    if base is not None and base.data_ptr() == w.data_ptr() and base.numel() == w.numel():
        w = base
    key = id(w)
    ent = _prepared.get(key)
    if ent is not None and ent.current(w):
        return
    lib = load_library()
    if ent is not None and ent.ptr != w.data_ptr():
        lib.tcx_forget_weight(ent.ptr)
    if ent is not None and ent.ref() is w and ent.ptr == w.data_ptr() and ent.w16.numel() == w.numel() and ent.w16.device == w.device:
        w16 = ent.w16         # same storage: convert into the existing copy (its address may be baked into graphs / optimizer tables)
        tracked = ent.tracked
    else:
        w16 = torch.empty(w.numel(), dtype=torch.float16, device=w.device)
        tracked = False
    src = w.detach()
    if not src.is_contiguous():
        raise RuntimeError("transception_b200: weight matrices must be contiguous")
    if conv is not None:
        rc = lib.tcx_prepare_conv_weight_f16(src.data_ptr(), w16.data_ptr(), conv[0], conv[1], conv[2], _stream())
    else:
        rc = lib.tcx_prepare_weight_f16(src.data_ptr(), w16.data_ptr(), w.numel(), _stream())
    if rc != 0:
        raise RuntimeError("libtransception_sm100: " + lib.tcx_last_error().decode())
    ptr = w.data_ptr()

    def _gone(_ref, key=key, ptr=ptr):
        e = _prepared.get(key)
        if e is not None and e.ptr == ptr and e.ref() is None:
            _prepared.pop(key, None)
            if _lib is not None:
                _lib.tcx_forget_weight(ptr)
    e = _Prep()
    e.ref, e.version, e.ptr, e.w16, e.conv, e.gen, e.tracked = weakref.ref(w, _gone), w._version, ptr, w16, conv, _raw_gen, tracked
    _prepared[key] = e


def prepared_copy(w, track=False):
    """(fp16 copy, conv) of a parameter the forward has prepared, or None.  ``conv`` is the (N, Cin, r) of a patchify-conv
    weight whose copy is K-permuted (refreshed with tcx_prepare_conv_weight_f16), else None.  ``track``: the caller (an
    optimizer table) takes over keeping the copy current after its raw-pointer updates."""
    ent = _prepared.get(id(w))
    if ent is None or ent.ref() is not w or ent.ptr != w.data_ptr():
        return None
    if track:
        ent.tracked = True
    return ent.w16, ent.conv


def refresh_prepared(w):
    """Re-convert the prepared copy of ``w`` from its current values on the current stream (a weight changed through a raw
    pointer or ``.data``, which does not move the version counter)."""
    ent = _prepared.get(id(w))
    if ent is None or ent.ref() is not w:
        return
    lib = load_library()
    if ent.conv is not None:
        rc = lib.tcx_prepare_conv_weight_f16(w.data_ptr(), ent.w16.data_ptr(), ent.conv[0], ent.conv[1], ent.conv[2], _stream())
    else:
        rc = lib.tcx_prepare_weight_f16(w.data_ptr(), ent.w16.data_ptr(), w.numel(), _stream())
    if rc != 0:
        raise RuntimeError("libtransception_sm100: " + lib.tcx_last_error().decode())


def invalidate_prepared(model=None):
    """Forget the version stamps of the prepared fp16 copies (of ``model``'s parameters, or all): the next forward converts
    them again.  Needed after in-place writes through ``p.data`` (EMA swaps, hand-written optimizers), which do not bump
    ``_version``; also used to force the conversions INTO a captured forward graph."""
    ids = None if model is None else {id(p) for p in model.parameters()}
    for key, ent in _prepared.items():
        if ids is None or key in ids:
            ent.version = -1


def _table(tensors, mats=()):
    """Host array of device pointers; slots listed in ``mats`` are GEMM weight matrices (prepared on first use)."""
    arr = (ctypes.c_void_p * len(tensors))()
    tensors = [_aligned(t) for t in tensors]       # re-aligned replicas (and their prepared copies) live on in _KEEP
    for i in mats:
        prepare_weight(tensors[i])
    for i, t in enumerate(tensors):
        arr[i] = _ptr(t.detach() if t is not None else None)
    return arr


def _ws(nbytes, like):
    return torch.empty((max(int(nbytes), 16) + 3) // 4, dtype=torch.float32, device=like.device)


def _d(t):
    return t.detach() if t is not None else None


# ------------------------------------------------------------------------------------------------
def layernorm(x, w, b, eps):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    y = torch.empty_like(x)
    C = x.shape[-1]
    _chk(lib.tcx_layernorm_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(b)), _ptr(y), x.numel() // C, C, eps, _stream()))
    return y


LN_DUAL_WIDTHS = (64, 128, 256, 320, 512)


def layernorm_dual(x, w, b, eps):
    """(LN(x) in fp32, the same values in fp16) from one pass — the fp16 twin is the GEMM operand of the training node that
    follows (mixffn_skip_train / mb_factor_attn_train ``xn16=``).  Returns (y, None) for widths the dual kernel is not built for."""
    require_cuda(x)
    C = x.shape[-1]
    if not USE_F16 or C not in LN_DUAL_WIDTHS:
        return layernorm(x, w, b, eps), None
    lib = load_library()
    x = x.contiguous()
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _chk(lib.tcx_layernorm_dual_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(b)), _ptr(y), _ptr16(y16), x.numel() // C, C, eps, _stream()))
    return y, y16


def linear(x, w, b=None, act=0, residual=None):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    K = x.shape[-1]
    N = w.shape[0]
    M = x.numel() // K
    y = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=x.dtype)
    _chk(lib.tcx_linear_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(b)), _ptr(residual), _ptr(y), M, N, K, act, _stream()))
    return y


def to_f16(x):
    """fp32 -> fp16 copy made by the library's own conversion kernel."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _chk(lib.tcx_f32_to_f16(_ptr(x.detach()), _ptr16(y), x.numel(), _stream()))
    return y


def linear_f16(x16, w16, b=None, residual=None, out_f16=False):
    """fp16-operand nn.Linear: x16 [..,K] fp16, w16 [N,K] fp16 -> fp32 (+bias +residual) or fp16 (+bias)."""
    require_cuda(x16)
    lib = load_library()
    K = x16.shape[-1]
    N = w16.shape[0]
    M = x16.numel() // K
    y = torch.empty(x16.shape[:-1] + (N,), device=x16.device, dtype=torch.float16 if out_f16 else torch.float32)
    _chk(lib.tcx_linear_f16_fwd(_ptr16(x16), _ptr16(w16), _ptr(_d(b)), _ptr(residual),
                                _ptr16(y) if out_f16 else _ptr(y), M, N, K, int(out_f16), _stream()))
    return y


def linear_bn_act(x, w, bn_w, bn_b, bn_rm, bn_rv, bn_eps, hardswish=False):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=x.dtype)
    _chk(lib.tcx_linear_bn_act_fwd(_ptr(x), _ptr(_d(w).reshape(N, K)), _ptr(_d(bn_w)), _ptr(_d(bn_b)), _ptr(bn_rm),
                                   _ptr(bn_rv), bn_eps, 2 if hardswish else 0, _ptr(y), M, N, K, _stream()))
    return y


def patch_embed_ln(x, w, b, stride, padding, lnw, lnb, eps):
    require_cuda(x)
    lib = load_library()
    B, Cin, H, W = x.shape
    if tuple(w.shape) != (64, 3, 7, 7) or stride != 4 or padding != 3:
        raise NotImplementedError("patch_embed_ln is built for the stage-1 stem (3->64, 7x7, stride 4, pad 3)")
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    out = torch.empty((B, Ho * Wo, 64), device=x.device, dtype=x.dtype)
    _chk(lib.tcx_patch_embed_ln_fwd(_ptr(x), B, Cin, H, W, _ptr(_d(w)), _ptr(_d(b)), _ptr(_d(lnw)), _ptr(_d(lnb)),
                                    eps, _ptr(out), _stream()))
    return out


def sum_tensors(ts):
    """ts[0] + ts[1] + ... (up to 16 equally shaped fp32 tensors, index order) in one launch."""
    require_cuda(ts[0])
    lib = load_library()
    ts = [t.contiguous() for t in ts]
    out = torch.empty_like(ts[0])
    while len(ts) > 16:                      # deeper fan-outs: fold the first 16 and continue
        head = torch.empty_like(ts[0])
        tab = (ctypes.c_void_p * 16)(*[_ptr(t) for t in ts[:16]])
        _chk(lib.tcx_sum_tensors(tab, 16, head.numel(), _ptr(head), _stream()))
        ts = [head] + ts[16:]
    tab = (ctypes.c_void_p * len(ts))(*[_ptr(t) for t in ts])
    _chk(lib.tcx_sum_tensors(tab, len(ts), out.numel(), _ptr(out), _stream()))
    return out


def patch_embed_conv(x, w, b):
    """The stem conv alone (no LayerNorm): [B, Ho*Wo, 64] tokens."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, Cin, H, W = x.shape
    if tuple(w.shape) != (64, 3, 7, 7):
        raise NotImplementedError("patch_embed_conv is built for the stage-1 stem (3->64, 7x7, stride 4, pad 3)")
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    out = torch.empty((B, Ho * Wo, 64), device=x.device, dtype=x.dtype)
    _chk(lib.tcx_patch_embed_ln_fwd(_ptr(x), B, Cin, H, W, _ptr(_d(w)), _ptr(_d(b)), None, None, 0.0, _ptr(out), _stream()))
    return out


def patch_embed_conv_bwd(x, w, dy):
    """(dw [64, 3, 7, 7], db [64]) of patch_embed_conv: one im2row launch + the Linear weight-gradient kernel."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, Cin, H, W = x.shape
    Ho, Wo = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
    Kp = 148
    patches = torch.empty((B * Ho * Wo, Kp), device=x.device, dtype=torch.float32)
    _chk(lib.tcx_patch_im2row_fwd(_ptr(x), B, Cin, H, W, _ptr(patches), Kp, _stream()))
    wpad = torch.empty((64, Kp), device=x.device, dtype=torch.float32)       # shape only: dx is not requested
    _, dwp, db = linear_bwd(patches, wpad, dy.reshape(-1, 64), need_dx=False)
    return dwp[:, :147].reshape(w.shape), db


def dwconv_tokens(x, H, W, w, b, add_input):
    require_cuda(x)
    lib = load_library()
    B, N, C = x.shape
    y = torch.empty_like(x)
    _chk(lib.tcx_dwconv_tokens_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(b)), _ptr(y), B, H, W, C, int(add_input), _stream()))
    return y


def eff_attn(xn, kw, kb, qw, qb, vw, vb, rw, rb, residual=None, reinterpret=False):
    require_cuda(xn)
    lib = load_library()
    B, N, C = xn.shape
    y = torch.empty_like(xn)
    ws = _ws(lib.tcx_eff_attn_workspace_bytes(B, N, C), xn)
    tab = _table([kw.reshape(C, C), kb, qw.reshape(C, C), qb, vw.reshape(C, C), vb, rw.reshape(C, C), rb], mats=(0, 2, 4, 6))
    _chk(lib.tcx_eff_attn_fwd(_ptr(xn), tab, _ptr(residual), _ptr(y), B, N, C, int(reinterpret), _ptr(ws), _stream()))
    return y


def mixffn_skip(xn, H, W, fc1w, fc1b, dww, dwb, lnw, lnb, eps, fc2w, fc2b, residual=None):
    require_cuda(xn)
    lib = load_library()
    xn = xn.contiguous()
    B, N, C = xn.shape
    C4 = fc1w.shape[0]
    y = torch.empty_like(xn)
    ws = _ws(lib.tcx_mixffn_skip_workspace_bytes(B, N, C4), xn)
    tab = _table([fc1w, fc1b, dww, dwb, lnw, lnb, fc2w, fc2b], mats=(0, 6))
    _chk(lib.tcx_mixffn_skip_fwd(_ptr(xn), tab, eps, _ptr(residual), _ptr(y), B, H, W, C, C4, _ptr(ws), _stream()))
    return y


def mb_factor_attn(xn, H, W, heads, scale, qkvw, qkvb, crpe_w, crpe_b, head_splits, projw, projb, residual=None,
                   keep_ws=False):
    require_cuda(xn)
    lib = load_library()
    B, N, C = xn.shape
    _check_crpe(head_splits, crpe_w, heads)
    y = torch.empty_like(xn)
    ws = _ws(lib.tcx_mb_factor_attn_workspace_bytes(B, N, C), xn)
    tab = _table([qkvw, qkvb, crpe_w[0], crpe_b[0], crpe_w[1], crpe_b[1], crpe_w[2], crpe_b[2], projw, projb])
    _chk(lib.tcx_mb_factor_attn_fwd(_ptr(xn), tab, _ptr(residual), _ptr(y), B, H, W, C, heads, _ptr(ws), _stream()))
    if keep_ws:
        return y, ws
    return y


def _check_crpe(head_splits, crpe_w, heads):
    if list(head_splits) != [2, 3, 3] or heads != 8 or [w.shape[-1] for w in crpe_w] != [3, 5, 7]:
        raise NotImplementedError("conv relative position encoding is built for crpe_window={3:2,5:3,7:3}, 8 heads")


def crpe(q, v, H, W, weights, biases, head_splits):
    raise NotImplementedError("ConvRelPosEnc is fused into the MB attention kernel; call FactorAtt_ConvRelPosEnc")


def _block_params(blk):
    f, m, c = blk.factoratt_crpe, blk.mlp, blk.factoratt_crpe.crpe
    _check_crpe(c.head_splits, [k.weight for k in c.conv_list], f.num_heads)
    cl = c.conv_list
    return [blk.cpe.proj.weight, blk.cpe.proj.bias, blk.norm1.weight, blk.norm1.bias, f.qkv.weight, f.qkv.bias,
            cl[0].weight, cl[0].bias, cl[1].weight, cl[1].bias, cl[2].weight, cl[2].bias, f.proj.weight, f.proj.bias,
            blk.norm2.weight, blk.norm2.bias, m.fc1.weight, m.fc1.bias, m.dwconv.dwconv.weight, m.dwconv.dwconv.bias,
            m.norm1.weight, m.norm1.bias, m.fc2.weight, m.fc2.bias]


def mhca_blocks(x, H, W, branches):
    """x: [G,B,N,C]; branches: G lists of L MHCABlock modules. Returns a new [G,B,N,C] tensor."""
    require_cuda(x)
    lib = load_library()
    G, B, N, C = x.shape
    L = len(branches[0])
    assert len(branches) == G and all(len(b) == L for b in branches)
    flat = []
    for br in branches:
        for blk in br:
            flat.extend(_block_params(blk))
    b0 = branches[0][0]
    x = x.contiguous()
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_mhca_blocks_workspace_bytes(G, B, N, C), x)
    mats = [k * MHCA_NP + j for k in range(G * L) for j in (4, 12, 16, 22)]
    _chk(lib.tcx_mhca_blocks_fwd(_ptr(x), _ptr(y), _table(flat, mats), G, L, B, H, W, C, b0.factoratt_crpe.num_heads,
                                 b0.norm1.eps, b0.mlp.norm1.eps, _ptr(ws), _stream()))
    return y


def ripm_dwsep_bn_hs(x, stride, dww, pww, bn_w, bn_b, bn_rm, bn_rv, bn_eps, out=None):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, C), device=x.device, dtype=x.dtype)
    ws = _ws(lib.tcx_ripm_dwsep_bn_hs_workspace_bytes(B, H, W, C, stride), x)
    _chk(lib.tcx_ripm_dwsep_bn_hs_fwd(_ptr(x), _ptr(_d(dww)), _ptr(_d(pww).reshape(C, C)), _ptr(_d(bn_w)),
                                      _ptr(_d(bn_b)), _ptr(bn_rm), _ptr(bn_rv), bn_eps, _ptr(out), B, H, W, C, stride,
                                      _ptr(ws), _stream()))
    return out


def resblock(x, c1w, bn1, dww, bn2, c2w, bn3):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_resblock_workspace_bytes(B, H, W, C), x)
    tab = _table([c1w.reshape(C, C), *bn1[:4], dww, *bn2[:4], c2w.reshape(C, C), *bn3[:4]])
    _chk(lib.tcx_resblock_fwd(_ptr(x), tab, bn1[4], _ptr(y), B, H, W, C, _ptr(ws), _stream()))
    return y


def iff_coordatt(maps, c1w, c1b, bn, chw, chb, cww, cwb, cow, cob):
    require_cuda(maps[0])
    lib = load_library()
    if len(maps) == 1:   # a single pre-concatenated NHWC map: split channel-wise into 4 dense sources
        B, H, W, C4 = maps[0].shape
        maps = [m.contiguous() for m in maps[0].split(C4 // 4, dim=3)]
    maps = [m.contiguous() for m in maps]
    B, H, W, C = maps[0].shape
    assert H == W and len(maps) == 4
    mip, inp, Cout = c1w.shape[0], 4 * C, cow.shape[0]
    y = torch.empty((B, H, W, Cout), device=maps[0].device, dtype=maps[0].dtype)
    ws = _ws(lib.tcx_iff_coordatt_workspace_bytes(B, H, C, mip), maps[0])
    tab = _table([c1w.reshape(mip, inp), c1b, *bn[:4], chw.reshape(inp, mip), chb, cww.reshape(inp, mip), cwb,
                  cow.reshape(Cout, inp), cob])
    _chk(lib.tcx_iff_coordatt_fwd(_table(maps), tab, bn[4], _ptr(y), B, H, C, mip, Cout, _ptr(ws), _stream()))
    return y


def _bridge_side(ntok):
    # ntok = S^2 * (1 + 2/4 + 5/16 + 8/64) = S^2 * 1.9375
    S = int(round((ntok / 1.9375) ** 0.5))
    if S * S * 31 != ntok * 16:
        raise RuntimeError("bridge: token count %d does not correspond to a 4-scale pyramid" % ntok)
    return S


def bridge_regroup(maps):
    require_cuda(maps[0])
    lib = load_library()
    maps = [m.contiguous() for m in maps]
    B, S = maps[0].shape[0], maps[0].shape[1]
    exp = [(B, S >> k, S >> k, c) for k, c in enumerate((64, 128, 320, 512))]
    if [tuple(m.shape) for m in maps] != exp:
        raise RuntimeError("bridge_regroup: expected NHWC maps %s, got %s" % (exp, [tuple(m.shape) for m in maps]))
    ntok = sum(m.shape[1] * m.shape[2] * m.shape[3] // 64 for m in maps)
    out = torch.empty((B, ntok, 64), device=maps[0].device, dtype=maps[0].dtype)
    _chk(lib.tcx_bridge_regroup_fwd(_table(maps), _ptr(out), B, S, _stream()))
    return out


_BRIDGE_MULT = (1, 2, 5, 8)


def bridge_split(tokens):
    """[B, Ntok, 64] token buffer -> its four dense per-scale slabs [B, (S/2^k)^2, 64*{1,2,5,8}] (one launch)."""
    require_cuda(tokens)
    lib = load_library()
    tokens = tokens.contiguous()
    B, ntok, C = tokens.shape
    S = _bridge_side(ntok)
    outs = [torch.empty((B, (S >> k) ** 2, 64 * m), device=tokens.device, dtype=tokens.dtype) for k, m in enumerate(_BRIDGE_MULT)]
    tab = (ctypes.c_void_p * 4)(*[_ptr(o) for o in outs])
    _chk(lib.tcx_bridge_split_fwd(_ptr(tokens), tab, B, S, _stream()))
    return outs


def bridge_merge(slabs, residual=None):
    """Four per-scale slabs (any shape with B leading and (S/2^k)^2 * 64*{1,2,5,8} elements per image) -> [B, Ntok, 64]
    (+ residual), one launch."""
    require_cuda(slabs[0])
    lib = load_library()
    slabs = [m.contiguous() for m in slabs]
    B = slabs[0].shape[0]
    S = int(round((slabs[0].numel() // (B * 64)) ** 0.5))
    per = [(S >> k) ** 2 * m for k, m in enumerate(_BRIDGE_MULT)]
    if S <= 0 or S % 8 or [m.numel() for m in slabs] != [B * n * 64 for n in per]:
        raise RuntimeError("bridge_merge: expected slabs of %s tokens per image, got %s" % (per, [tuple(m.shape) for m in slabs]))
    out = torch.empty((B, sum(per), 64), device=slabs[0].device, dtype=slabs[0].dtype)
    if residual is not None:
        residual = residual.contiguous()
        if residual.shape != out.shape:
            raise RuntimeError("bridge_merge: residual %s does not match %s" % (tuple(residual.shape), tuple(out.shape)))
    tab = (ctypes.c_void_p * 4)(*[_ptr(m) for m in slabs])
    _chk(lib.tcx_bridge_merge_fwd(tab, _ptr(residual), _ptr(out), B, S, _stream()))
    return out


def scale_reduce_pack_train(x, params):
    """Scale_reduce without its LayerNorm: (packed [B, Nred, 64], saved).  params = [sr0_w, sr0_b, sr1_w, sr1_b, sr2_w, sr2_b]."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, ntok, C = x.shape
    S = _bridge_side(ntok)
    nred = (S // 8) ** 2 * 16
    packed = torch.empty((B, nred, 64), device=x.device, dtype=x.dtype)
    saved = _ws(lib.tcx_scale_reduce_saved_bytes(B, S), x)
    ws = _ws(lib.tcx_scale_reduce_train_workspace_bytes(B, S), x)
    _chk(lib.tcx_scale_reduce_train_fwd(_ptr(x), _table(params), _ptr(packed), B, S, _ptr(saved), _ptr(ws), _stream()))
    return packed, saved


def scale_reduce_pack_bwd(dpacked, saved, params, ntok):
    """(dx [B, Ntok, 64], six parameter gradients) of scale_reduce_pack_train."""
    require_cuda(dpacked)
    lib = load_library()
    dpacked = dpacked.contiguous()
    B = dpacked.shape[0]
    S = _bridge_side(ntok)
    dx = torch.empty((B, ntok, 64), device=dpacked.device, dtype=dpacked.dtype)
    grads = [torch.empty_like(p) for p in params]
    gtab = (ctypes.c_void_p * 6)(*[_ptr(g) for g in grads])
    ws = _ws(lib.tcx_scale_reduce_bwd_workspace_bytes(B, S), dpacked)
    _chk(lib.tcx_scale_reduce_bwd(_ptr(dpacked), _table(params), _ptr(saved), _ptr(dx), gtab, B, S, _ptr(ws), _stream()))
    return dx, grads


def scale_reduce(x, s0w, s0b, s1w, s1b, s2w, s2b, lnw, lnb, eps):
    require_cuda(x)
    lib = load_library()
    B, ntok, C = x.shape
    S = _bridge_side(ntok)
    nred = (S // 8) ** 2 * 8 + (S // 8) ** 2 * 8
    out = torch.empty((B, nred, 64), device=x.device, dtype=x.dtype)
    ws = _ws(lib.tcx_scale_reduce_workspace_bytes(B, S), x)
    tab = _table([s0w.reshape(64, -1), s0b, s1w.reshape(128, -1), s1b, s2w.reshape(320, -1), s2b, lnw, lnb])
    _chk(lib.tcx_scale_reduce_fwd(_ptr(x), tab, eps, _ptr(out), B, S, _ptr(ws), _stream()))
    return out


def bridge_sr_attn(xn, scale, qw, qb, kvw, kvb, pw, pb, s0w, s0b, s1w, s1b, s2w, s2b, lnw, lnb, eps, residual=None):
    require_cuda(xn)
    lib = load_library()
    B, ntok, C = xn.shape
    S = _bridge_side(ntok)
    y = torch.empty_like(xn)
    ws = _ws(lib.tcx_bridge_sr_attn_workspace_bytes(B, S), xn)
    tab = _table([qw, qb, kvw, kvb, pw, pb, s0w.reshape(64, -1), s0b, s1w.reshape(128, -1), s1b,
                  s2w.reshape(320, -1), s2b, lnw, lnb])
    _chk(lib.tcx_bridge_sr_attn_fwd(_ptr(xn), tab, scale, eps, _ptr(residual), _ptr(y), B, S, _ptr(ws), _stream()))
    return y


def flash_attn(q, kv, scale):
    """softmax(q k^T * scale) v for one 64-wide head: q [B,Nq,64], kv [B,Nk,128] (k | v) -> [B,Nq,64]."""
    require_cuda(q)
    lib = load_library()
    q, kv = q.contiguous(), kv.contiguous()
    B, Nq, D = q.shape
    Nk = kv.shape[1]
    if D != 64 or kv.shape[2] != 128 or kv.shape[0] != B:
        raise RuntimeError("flash_attn: expected q [B,Nq,64] and kv [B,Nk,128], got %s %s" % (tuple(q.shape), tuple(kv.shape)))
    out = torch.empty_like(q)
    ws = _ws(lib.tcx_flash_attn_workspace_bytes(B, Nk), q)
    _chk(lib.tcx_flash_attn_fwd(_ptr(q), _ptr(kv), _ptr(out), B, Nq, Nk, scale, _ptr(ws), _stream()))
    return out


def flash_attn_f16(q16, kv16, scale):
    """fp16 form of flash_attn: q16 [B,Nq,64], kv16 [B,Nk,128] (k | v) -> fp16 [B,Nq,64] (tcgen05 kernel)."""
    require_cuda(q16)
    lib = load_library()
    B, Nq, D = q16.shape
    Nk = kv16.shape[1]
    if D != 64 or kv16.shape[2] != 128 or kv16.shape[0] != B:
        raise RuntimeError("flash_attn_f16: expected q [B,Nq,64] and kv [B,Nk,128]")
    out = torch.empty_like(q16)
    ws = _ws(lib.tcx_flash_attn_workspace_bytes(B, Nk), q16)
    _chk(lib.tcx_flash_attn_f16_fwd(_ptr16(q16), _ptr16(kv16), _ptr16(out), B, Nq, Nk, scale, _ptr(ws), _stream()))
    return out


def bridge_mixffn(tx, tx1, mix_args):
    require_cuda(tx)
    lib = load_library()
    B, ntok, C = tx.shape
    S = _bridge_side(ntok)
    y = torch.empty_like(tx)
    flat, eps = [], mix_args[0][6]
    for a in mix_args:
        flat.extend([a[0], a[1], a[2], a[3], a[4], a[5], a[7], a[8]])
    ws = _ws(lib.tcx_bridge_mixffn_workspace_bytes(B, S), tx)
    mats = [8 * k + j for k in range(4) for j in (0, 6)]
    _chk(lib.tcx_bridge_mixffn_fwd(_ptr(tx), _ptr(tx1), _table(flat, mats), eps, _ptr(y), B, S, _ptr(ws), _stream()))
    return y


def _mix_slots(a):
    """MixFFN_skip.args() -> the 8 C-ABI slots {fc1_w,fc1_b,dw_w,dw_b,ln_w,ln_b,fc2_w,fc2_b} and its LN eps."""
    return [a[0], a[1], a[2], a[3], a[4], a[5], a[7], a[8]], a[6]


def eff_block(x, H, W, n1w, n1b, ln_eps, attn_args, n2w, n2b, mix_args):
    """EfficientTransformerBlock.forward (reference MSTr.py:164-173) in one call."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, N, C = x.shape
    kw, kb, qw, qb, vw, vb, rw, rb = attn_args
    mix, mlp_eps = _mix_slots(mix_args)
    slots = [n1w, n1b, kw.reshape(C, C), kb, qw.reshape(C, C), qb, vw.reshape(C, C), vb, rw.reshape(C, C), rb,
             n2w, n2b] + mix
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_eff_block_workspace_bytes(B, N, C), x)
    _chk(lib.tcx_eff_block_fwd(_ptr(x), _table(slots, mats=(2, 4, 6, 8, 12, 18)), ln_eps, mlp_eps, _ptr(y), B, H, W, C, _ptr(ws),
                               _stream()))
    return y


def bridge_layer(x, n1w, n1b, ln_eps, channel_att, attn_slots, scale, n2w, n2b, mix_args_list):
    """BridgLayer_4.forward (reference MSTr.py:2373-2409) on the token buffer in one call."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, ntok, C = x.shape
    S = _bridge_side(ntok)
    attn_slots = list(attn_slots) + [None] * (14 - len(attn_slots))
    slots = [n1w, n1b] + attn_slots + [n2w, n2b]
    for a in mix_args_list:
        slots.extend(_mix_slots(a)[0])
    slots = [_aligned(t) for t in slots]
    mats = [18 + 8 * k + j for k in range(4) for j in (0, 6)]
    if channel_att:
        mats += [2, 4, 6, 8]
    else:
        mats += [2, 4, 6]
        for k, (cin, r) in enumerate(((64, 8), (128, 4), (320, 2))):
            prepare_weight(slots[8 + 2 * k], conv=(cin, cin, r))
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_bridge_layer_workspace_bytes(B, S), x)
    _chk(lib.tcx_bridge_layer_fwd(_ptr(x), _table(slots, mats), int(channel_att), scale, ln_eps, _ptr(y), B, S,
                                  _ptr(ws), _stream()))
    return y


def _bridge_layer_slots(n1w, n1b, channel_att, attn_slots, n2w, n2b, mix_args_list):
    attn_slots = list(attn_slots) + [None] * (14 - len(attn_slots))
    slots = [n1w, n1b] + attn_slots + [n2w, n2b]
    for a in mix_args_list:
        slots.extend(_mix_slots(a)[0])
    slots = [_aligned(t) for t in slots]
    mats = [18 + 8 * k + j for k in range(4) for j in (0, 6)]
    if channel_att:
        mats += [2, 4, 6, 8]
    else:
        mats += [2, 4, 6]
        for k, (cin, r) in enumerate(((64, 8), (128, 4), (320, 2))):
            prepare_weight(slots[8 + 2 * k], conv=(cin, cin, r))
    return slots, mats


def bridge_block(x, layers, scale, ln_eps):
    """BridgeBlock_4.forward on the token buffer: ``layers`` = [(n1w, n1b, channel_att, attn_slots, n2w, n2b, mix_args_list)]."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, ntok, C = x.shape
    S = _bridge_side(ntok)
    slots, mats, flags = [], [], []
    for i, lay in enumerate(layers):
        sl, mt = _bridge_layer_slots(*lay)
        mats += [i * 50 + m for m in mt]
        slots += sl
        flags.append(int(lay[2]))
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_bridge_block_workspace_bytes(B, S), x)
    fl = (ctypes.c_int * len(flags))(*flags)
    _chk(lib.tcx_bridge_block_fwd(_ptr(x), _table(slots, mats), fl, len(layers), scale, ln_eps, _ptr(y), B, S, _ptr(ws),
                                  _stream()))
    return y


def concat_linear(x1, x2, w, b):
    """Linear(cat[x1, x2]) without materialising the concatenation.  ``x2`` may be a per-image slab of a larger buffer
    ([B,N,C2] with dense rows and an arbitrary batch pitch): it is read in place."""
    require_cuda(x1)
    lib = load_library()
    B, N, C1 = x1.shape
    C2 = x2.shape[-1]
    Nout = w.shape[0]
    x1 = x1.contiguous()
    slab = (x2.dim() == 3 and x2.stride(2) == 1 and x2.stride(1) == C2 and x2.stride(0) % 4 == 0 and
            x2.dtype == torch.float32 and x2.is_cuda and x2.data_ptr() % 16 == 0)
    y = torch.empty((B, N, Nout), device=x1.device, dtype=x1.dtype)
    if slab and not x2.is_contiguous():
        _chk(lib.tcx_concat_linear_fwd(_ptr(x1), x2.data_ptr(), _ptr(_d(w)), _ptr(_d(b)), _ptr(y), N, C1, C2, Nout,
                                       B, x2.stride(0), _stream()))
    else:
        _chk(lib.tcx_concat_linear_fwd(_ptr(x1), _ptr(x2.contiguous()), _ptr(_d(w)), _ptr(_d(b)), _ptr(y), B * N, C1, C2,
                                       Nout, 1, 0, _stream()))
    return y


def patch_expand(x, H, W, w, scale, lnw, lnb, eps):
    require_cuda(x)
    lib = load_library()
    B, N, C = x.shape
    c = w.shape[0] // (scale * scale)
    y = torch.empty((B, N * scale * scale, c), device=x.device, dtype=x.dtype)
    ws = _ws(lib.tcx_patch_expand_workspace_bytes(B, H, W, C, scale), x)
    _chk(lib.tcx_patch_expand_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(lnw)), _ptr(_d(lnb)), eps, _ptr(y), B, H, W, C, scale,
                                  _ptr(ws), _stream()))
    return y


def final_expand_head(x, H, W, ew, lnw, lnb, eps, cw, cb):
    require_cuda(x)
    lib = load_library()
    B, N, C = x.shape
    if C != 64:
        raise NotImplementedError("final_expand_head is built for dim 64")
    ncls = cw.shape[0]
    y = torch.empty((B, ncls, 4 * H, 4 * W), device=x.device, dtype=x.dtype)
    ws = _ws(lib.tcx_final_expand_head_workspace_bytes(B, H, W), x)
    _chk(lib.tcx_final_expand_head_fwd(_ptr(x), _ptr(_d(ew)), _ptr(_d(lnw)), _ptr(_d(lnb)), eps,
                                       _ptr(_d(cw).reshape(ncls, 64)), _ptr(_d(cb)), ncls, _ptr(y), B, H, W,
                                       _ptr(ws), _stream()))
    return y


def final_head_train(e, H, W, lnw, lnb, eps, cw, cb):
    """Pixel shuffle x4 + LayerNorm(64) + class head on the expand output e [B, H*W, 1024] -> NCHW logits [B, ncls, 4H, 4W]."""
    require_cuda(e)
    lib = load_library()
    e = e.contiguous()
    B = e.shape[0]
    if e.shape[-1] != 1024 or e.numel() != B * H * W * 1024:
        raise NotImplementedError("final_head_train is built for dim 64 (expand output 16 * 64 wide)")
    ncls = cw.shape[0]
    y = torch.empty((B, ncls, 4 * H, 4 * W), device=e.device, dtype=e.dtype)
    _chk(lib.tcx_final_head_train_fwd(_ptr(e), _ptr(_d(lnw)), _ptr(_d(lnb)), eps, _ptr(_d(cw).reshape(ncls, 64)), _ptr(_d(cb)), ncls,
                                      _ptr(y), B, H, W, _stream()))
    return y


def final_head_bwd(e, dlogits, H, W, lnw, lnb, eps, cw):
    """(de, d ln_w, d ln_b, d cls_w, d cls_b) of final_head_train."""
    require_cuda(e)
    lib = load_library()
    dlogits = dlogits.contiguous()
    B, ncls = dlogits.shape[0], dlogits.shape[1]
    if ncls > 16:
        raise NotImplementedError("final_head_bwd is built for at most 16 classes (got %d)" % ncls)
    de = torch.empty_like(e)
    dlnw, dlnb, dcw, dcb = torch.empty_like(lnw), torch.empty_like(lnb), torch.empty_like(cw), torch.empty(ncls, device=e.device, dtype=e.dtype)
    ws = _ws(lib.tcx_final_head_bwd_workspace_bytes(B, H, W), e)
    _chk(lib.tcx_final_head_bwd(_ptr(e), _ptr(dlogits), _ptr(_d(lnw)), _ptr(_d(lnb)), eps, _ptr(_d(cw).reshape(ncls, 64)), ncls, _ptr(de),
                                _ptr(dlnw), _ptr(dlnb), _ptr(dcw), _ptr(dcb), B, H, W, _ptr(ws), _stream()))
    return de, dlnw, dlnb, dcw, dcb


# ------------------------------------------------------------------------------------------------
# networks/Transception.py variant — fp16 pipeline only
# ------------------------------------------------------------------------------------------------
def fuse_eff_attn(xn, kw, kb, qw, qb, vw, vb, rw, rb, residual=None):
    """FuseEfficientAttention.forward (reference Transception.py:49-87, head_count=1) on tokens [B, N, C]."""
    require_cuda(xn)
    lib = load_library()
    xn = xn.contiguous()
    B, N, C = xn.shape
    y = torch.empty_like(xn)
    ws = _ws(lib.tcx_fuse_eff_attn_workspace_bytes(B, N, C), xn)
    tab = _table([kw, kb, qw, qb, vw, vb, rw, rb], mats=(0, 2, 4, 6))
    _chk(lib.tcx_fuse_eff_attn_fwd(_ptr(xn), tab, _ptr(residual.contiguous() if residual is not None else None), _ptr(y),
                                   B, N, C, _ptr(ws), _stream()))
    return y


def fuse_block(x, H1, W1, H2, W2, n1w, n1b, ln_eps, attn_args, n2w, n2b, mix1_args, mix2_args):
    """EfficientTransformerBlockFuse.forward (reference Transception.py:213-250) on [B, H1*W1 + H2*W2, C] tokens."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, N, C = x.shape
    if N != H1 * W1 + H2 * W2:
        raise NotImplementedError("EfficientTransformerBlockFuse is built for the two-branch token layout (n1 + n2 tokens)")
    m1, eps1 = _mix_slots(mix1_args)
    m2, eps2 = _mix_slots(mix2_args)
    if eps1 != eps2:
        raise NotImplementedError("mlp1 / mlp2 LayerNorm eps differ")
    slots = [n1w, n1b] + list(attn_args) + [n2w, n2b] + m1 + m2
    y = torch.empty_like(x)
    ws = _ws(lib.tcx_fuse_block_workspace_bytes(B, N, C), x)
    _chk(lib.tcx_fuse_block_fwd(_ptr(x), _table(slots, mats=(2, 4, 6, 8, 12, 18, 20, 26)), ln_eps, eps1, _ptr(y), B, H1, W1, H2, W2,
                                C, _ptr(ws), _stream()))
    return y


def dual_patch_embed(x_nhwc, pe1, pe2, ln_eps):
    """Both OverlapPatchEmbeddings_fuse branches of a stage (reference EffSegformer.py:117-131) on an NHWC fp32 map.
    pe = (proj_w [C,Cin,k,k], proj_b, norm_w, norm_b, k, stride, padding, dilation).  Returns (tokens, H1, W1, H2, W2)."""
    require_cuda(x_nhwc)
    lib = load_library()
    x = x_nhwc.contiguous()
    B, H, W, Cin = x.shape
    (w1, b1, nw1, nb1, k1, s1, p1, d1), (w2, b2, nw2, nb2, k2, s2, p2, d2) = pe1, pe2
    if s1 != s2 or d1 != d2:
        raise NotImplementedError("the two patch-merging branches must share stride and dilation")
    w1, w2 = _aligned(w1), _aligned(w2)
    C = w1.shape[0]
    H1, W1 = (H + 2 * p1 - d1 * (k1 - 1) - 1) // s1 + 1, (W + 2 * p1 - d1 * (k1 - 1) - 1) // s1 + 1
    H2, W2 = (H + 2 * p2 - d1 * (k2 - 1) - 1) // s1 + 1, (W + 2 * p2 - d1 * (k2 - 1) - 1) // s1 + 1
    prepare_weight(w1, conv=(C, Cin, k1))
    prepare_weight(w2, conv=(C, Cin, k2))
    tokens = torch.empty((B, H1 * W1 + H2 * W2, C), device=x.device, dtype=torch.float32)
    ws = _ws(lib.tcx_dual_patch_embed_workspace_bytes(B, H, W, Cin, C, k1, k2, s1, p1, p2, d1), x)
    tab = _table([w1, b1, nw1, nb1, w2, b2, nw2, nb2])
    _chk(lib.tcx_dual_patch_embed_fwd(_ptr(x), tab, ln_eps, _ptr(tokens), B, H, W, Cin, C, k1, k2, s1, p1, p2, d1, _ptr(ws),
                                      _stream()))
    return tokens, H1, W1, H2, W2


def fuse_merge(tokens, H1, W1, H2, W2, nw, nb, ln_eps, cw, cb):
    """Stage tail of MiT_3inception (reference Transception.py:462-476, concat='original') -> [B, H2*W2, C] tokens."""
    require_cuda(tokens)
    lib = load_library()
    t = tokens.contiguous()
    B, N, C = t.shape
    out = torch.empty((B, H2 * W2, C), device=t.device, dtype=torch.float32)
    ws = _ws(lib.tcx_fuse_merge_workspace_bytes(B, N, H2 * W2, C), t)
    tab = _table([nw, nb, cw.reshape(C, 2 * C), cb], mats=(2,))
    _chk(lib.tcx_fuse_merge_fwd(_ptr(t), tab, ln_eps, _ptr(out), B, H1, W1, H2, W2, C, _ptr(ws), _stream()))
    return out


def fuse_merge_sk(tokens, H1, W1, H2, W2, nw, nb, ln_eps, fc, fcs0, fcs1, conv, bn):
    """Stage tail with SK_Block fusion (reference Transception.py:477-481, :328-358) -> [B, H2*W2, C] tokens."""
    require_cuda(tokens)
    lib = load_library()
    t = tokens.contiguous()
    B, N, C = t.shape
    d = fc.weight.shape[0]
    out = torch.empty((B, H2 * W2, C), device=t.device, dtype=torch.float32)
    ws = _ws(lib.tcx_fuse_merge_sk_workspace_bytes(B, N, H2 * W2, C), t)
    tab = _table([nw, nb, fc.weight, fc.bias, fcs0.weight, fcs0.bias, fcs1.weight, fcs1.bias, conv.weight.reshape(C, C), conv.bias,
                  bn.weight, bn.bias, bn.running_mean, bn.running_var], mats=(8,))
    _chk(lib.tcx_fuse_merge_sk_fwd(_ptr(t), tab, ln_eps, bn.eps, _ptr(out), B, H1, W1, H2, W2, C, d, _ptr(ws), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# fused training loss
# ------------------------------------------------------------------------------------------------
_LABEL_KIND = {torch.int64: 0, torch.float32: 1, torch.int32: 2, torch.uint8: 3}


def _seg_loss_common(logits, labels, class_w):
    require_cuda(logits)
    require_cuda(labels)
    if logits.dtype != torch.float32 or labels.dtype not in _LABEL_KIND:
        raise TypeError("seg_loss: logits must be float32 and labels int64 / float32 / int32 / uint8 (got %s, %s)"
                        % (logits.dtype, labels.dtype))
    B, K = logits.shape[0], logits.shape[1]
    HW = logits[0, 0].numel()
    if labels.numel() != B * HW:
        raise ValueError("predict %s & target %s shape do not match" % (tuple(logits.shape), tuple(labels.shape)))
    cw = None
    if class_w is not None:
        cw = (ctypes.c_float * K)(*[float(v) for v in class_w])
    return B, K, HW, cw


def seg_loss_fwd(logits, labels, w_ce=0.4, w_dice=0.6, class_w=None, softmax=True):
    """Returns (out, ws): out = device tensor [loss, ce, dice, bad_labels, class_wise_dice...]; ws feeds seg_loss_bwd."""
    lib = load_library()
    logits, labels = logits.contiguous(), labels.contiguous()
    B, K, HW, cw = _seg_loss_common(logits, labels, class_w)
    out = torch.empty(4 + K, device=logits.device, dtype=torch.float32)
    ws = _ws(lib.tcx_seg_loss_workspace_bytes(B, K, HW), logits)
    _chk(lib.tcx_seg_loss_fwd(_ptr(logits), labels.data_ptr(), _LABEL_KIND[labels.dtype], B, K, HW, int(softmax), w_ce, w_dice, cw,
                              _ptr(out), _ptr(ws), _stream()))
    return out, ws


def seg_loss_bwd(logits, labels, ws, grad_out=None, w_ce=0.4, w_dice=0.6, class_w=None, softmax=True):
    lib = load_library()
    logits, labels = logits.contiguous(), labels.contiguous()
    B, K, HW, cw = _seg_loss_common(logits, labels, class_w)
    d = torch.empty_like(logits)
    go = grad_out.contiguous().float() if grad_out is not None else None
    _chk(lib.tcx_seg_loss_bwd(_ptr(logits), labels.data_ptr(), _LABEL_KIND[labels.dtype], B, K, HW, int(softmax), w_ce, w_dice, cw,
                              _ptr(go), _ptr(d), _ptr(ws), _stream()))
    return d


def argmax_classes(logits):
    """[B, K, H, W] fp32 logits -> [B, H, W] uint8 arg max over classes (reference utils.py:86)."""
    require_cuda(logits)
    lib = load_library()
    logits = logits.contiguous()
    B, K = logits.shape[:2]
    out = torch.empty((B,) + tuple(logits.shape[2:]), device=logits.device, dtype=torch.uint8)
    _chk(lib.tcx_argmax_classes_fwd(_ptr(logits), out.data_ptr(), B, K, out[0].numel(), _stream()))
    return out


# ---- training row: backward entries (include/transception_sm100.h, "training row") ------------------------------
def layernorm_bwd(x, w, dy, eps, dres=None):
    """(dx, dw, db) of nn.LayerNorm over the last dim; ``dres`` (optional, shape of x) is added to dx in the same kernel."""
    require_cuda(x)
    lib = load_library()
    x, dy = x.contiguous(), dy.contiguous()
    dres = dres.contiguous() if dres is not None else None
    C = x.shape[-1]
    M = x.numel() // C
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
    ws = _ws(lib.tcx_layernorm_bwd_workspace_bytes(M, C), x)
    _chk(lib.tcx_layernorm_bwd(_ptr(x), _ptr(_d(w)), _ptr(dy), _ptr(dres), eps, _ptr(dx), _ptr(dw), _ptr(db), M, C, _ptr(ws), _stream()))
    return dx, dw, db


def linear_bwd(x, w, dy, need_dx=True, need_dw=True, need_db=True):
    """(dx, dw, db) of y = x w^T + b; x fp32 or fp16 [.., K], w [N, K], dy [.., N]."""
    require_cuda(dy)
    lib = load_library()
    x, dy = x.contiguous(), dy.contiguous()
    N, K = w.shape
    M = dy.numel() // N
    dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if need_dx else None
    dw = torch.empty_like(w) if need_dw else None
    db = torch.empty(N, dtype=torch.float32, device=x.device) if need_db else None
    ws = _ws(lib.tcx_linear_bwd_workspace_bytes(M, N, K), dy)
    x16 = x.dtype == torch.float16
    _chk(lib.tcx_linear_bwd(_ptr16(x) if x16 else _ptr(x), int(x16), _ptr(_d(w)), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), M, N, K,
                            _ptr(ws), _stream()))
    return dx, dw, db


_WGRAD_FMT = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


def wgrad_mn(a, b, alpha=1.0, need_db=False, need_T=False, mask_ch=0):
    """out[z] = alpha * a[z]^T b[z] over the token axis (a [.., T, NL], b [.., T, KL], both fp32 (TF32 MMA), both fp16 or both
    bf16, optional leading batch dim): the MN-major tcgen05 weight-gradient kernel.  Returns (out, outT or None, db or None)."""
    require_cuda(a)
    if a.dtype != b.dtype or a.dtype not in _WGRAD_FMT or not (a.is_contiguous() and b.is_contiguous()):
        raise RuntimeError("wgrad_mn: operands must be contiguous and both fp32, both fp16 or both bf16")
    lib = load_library()
    fmt = _WGRAD_FMT[a.dtype]
    batch = a.shape[0] if a.dim() == 3 else 1
    T, NL, KL = a.shape[-2], a.shape[-1], b.shape[-1]
    shape = (batch, NL, KL) if a.dim() == 3 else (NL, KL)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    outT = torch.empty(shape[:-2] + (KL, NL), dtype=torch.float32, device=a.device) if need_T else None
    db = torch.empty(NL, dtype=torch.float32, device=a.device) if need_db else None
    ws = _ws(lib.tcx_wgrad_mn_workspace_bytes(T, NL, KL, batch, fmt), a)
    _chk(lib.tcx_wgrad_mn(a.data_ptr(), b.data_ptr(), fmt, T, NL, KL, NL, KL, batch, alpha, _ptr(out), _ptr(outT), _ptr(db), mask_ch,
                          _ptr(ws), _stream()))
    return out, outT, db


def mixffn_skip_train(xn, H, W, fc1w, fc1b, dww, dwb, lnw, lnb, eps, fc2w, fc2b, residual=None, xn16=None):
    """Training forward of MixFFN_skip: (y, saved) with ``saved`` the opaque buffer mixffn_skip_bwd consumes."""
    require_cuda(xn)
    lib = load_library()
    xn = xn.contiguous()
    B, N, C = xn.shape
    C4 = fc1w.shape[0]
    y = torch.empty_like(xn)
    saved = _ws(lib.tcx_mixffn_skip_saved_bytes(B, N, C, C4), xn)
    tab = _table([fc1w, fc1b, dww, dwb, lnw, lnb, fc2w, fc2b], mats=(0, 6))
    if xn16 is not None and xn16.shape != xn.shape:
        raise RuntimeError("mixffn_skip_train: xn16 %s does not match xn %s" % (tuple(xn16.shape), tuple(xn.shape)))
    _chk(lib.tcx_mixffn_skip_train_fwd(_ptr(xn), tab, eps, _ptr(residual), _ptr(y), B, H, W, C, C4, _ptr(saved), _ptr16(xn16), _stream()))
    return y, saved


def mixffn_skip_bwd(dy, saved, B, H, W, fc1w, fc1b, dww, dwb, lnw, lnb, eps, fc2w, fc2b, need_dx=True, xn=None):
    """(dxn, [8 parameter gradients in slot order]) of MixFFN_skip.  ``xn``: the fp32 forward input (optional; the saved fp16
    copy is converted when it is absent)."""
    require_cuda(dy)
    lib = load_library()
    dy = dy.contiguous()
    C4, C = fc1w.shape
    params = [fc1w, fc1b, dww, dwb, lnw, lnb, fc2w, fc2b]
    grads = [torch.empty_like(p) for p in params]
    dxn = torch.empty_like(dy) if need_dx else None
    tab = _table(params)
    gtab = (ctypes.c_void_p * 8)(*[_ptr(g) for g in grads])
    ws = _ws(lib.tcx_mixffn_skip_bwd_workspace_bytes(B, H * W, C, C4), dy)
    xn = xn.contiguous() if xn is not None else None
    _chk(lib.tcx_mixffn_skip_bwd(_ptr(dy), tab, eps, _ptr(saved), _ptr(xn), _ptr(dxn), gtab,
                                 B, H, W, C, C4, _ptr(ws), _stream()))
    return dxn, grads


def eff_attn_train(xn, kw, kb, qw, qb, vw, vb, rw, rb, residual=None):
    """Training forward of EfficientAttention on LayerNorm output xn [B, N, C]: (y, saved)."""
    require_cuda(xn)
    lib = load_library()
    xn = xn.contiguous()
    B, N, C = xn.shape
    y = torch.empty_like(xn)
    saved = _ws(lib.tcx_eff_attn_saved_bytes(B, N, C), xn)
    tab = _table([kw, kb, qw, qb, vw, vb, rw, rb], mats=(0, 2, 4, 6))
    _chk(lib.tcx_eff_attn_train_fwd(_ptr(xn), tab, _ptr(residual), _ptr(y), B, N, C, _ptr(saved), _stream()))
    return y, saved


def eff_attn_bwd(dy, saved, kw, kb, qw, qb, vw, vb, rw, rb, need_dx=True):
    """(dxn, [8 parameter gradients in slot order]) of EfficientAttention."""
    require_cuda(dy)
    lib = load_library()
    dy = dy.contiguous()
    B, N, C = dy.shape
    params = [kw, kb, qw, qb, vw, vb, rw, rb]
    grads = [torch.empty_like(p) for p in params]
    dxn = torch.empty_like(dy) if need_dx else None
    tab = _table(params)
    gtab = (ctypes.c_void_p * 8)(*[_ptr(g) for g in grads])
    ws = _ws(lib.tcx_eff_attn_bwd_workspace_bytes(B, N, C), dy)
    _chk(lib.tcx_eff_attn_bwd(_ptr(dy), tab, _ptr(saved), _ptr(dxn), gtab, B, N, C, _ptr(ws), _stream()))
    return dxn, grads


def mb_factor_attn_train(xn, H, W, heads, qkvw, qkvb, crpe_w, crpe_b, head_splits, projw, projb, residual=None, xn16=None):
    """Training forward of FactorAtt_ConvRelPosEnc on the fp16 pipeline: (y, saved) with ``saved`` the opaque buffer that
    mb_factor_attn_bwd(..., saved_f16=True) consumes."""
    require_cuda(xn)
    lib = load_library()
    xn = xn.contiguous()
    B, N, C = xn.shape
    _check_crpe(head_splits, crpe_w, heads)
    y = torch.empty_like(xn)
    saved = _ws(lib.tcx_mb_factor_attn_saved_bytes(B, N, C), xn)
    ws = _ws(lib.tcx_mb_factor_attn_train_workspace_bytes(B, N, C), xn)
    tab = _table([qkvw, qkvb, crpe_w[0], crpe_b[0], crpe_w[1], crpe_b[1], crpe_w[2], crpe_b[2], projw, projb], mats=(0, 8))
    if xn16 is not None and xn16.shape != xn.shape:
        raise RuntimeError("mb_factor_attn_train: xn16 %s does not match xn %s" % (tuple(xn16.shape), tuple(xn.shape)))
    _chk(lib.tcx_mb_factor_attn_train_fwd(_ptr(xn), tab, _ptr(residual), _ptr(y), B, H, W, C, heads, _ptr(saved), _ptr(ws), _ptr16(xn16),
                                          _stream()))
    return y, saved


def mb_factor_attn_bwd(dy, xn, fwd_ws, H, W, heads, qkvw, qkvb, crpe_w, crpe_b, projw, projb, need_dx=True, saved_f16=False):
    """(dxn, [10 parameter gradients in slot order]) of FactorAtt_ConvRelPosEnc; fwd_ws = mb_factor_attn(..., keep_ws=True)[1], or
    the ``saved`` buffer of mb_factor_attn_train with ``saved_f16``."""
    require_cuda(dy)
    lib = load_library()
    dy, xn = dy.contiguous(), xn.contiguous()
    B, N, C = dy.shape
    params = [qkvw, qkvb, crpe_w[0], crpe_b[0], crpe_w[1], crpe_b[1], crpe_w[2], crpe_b[2], projw, projb]
    grads = [torch.empty_like(p) for p in params]
    dxn = torch.empty_like(dy) if need_dx else None
    tab = _table(params)
    gtab = (ctypes.c_void_p * 10)(*[_ptr(g) for g in grads])
    ws = _ws(lib.tcx_mb_factor_attn_bwd_workspace_bytes(B, N, C), dy)
    _chk(lib.tcx_mb_factor_attn_bwd(_ptr(dy), _ptr(xn), tab, _ptr(fwd_ws), int(saved_f16), _ptr(dxn), gtab, B, H, W, C, heads, _ptr(ws),
                                    _stream()))
    return dxn, grads


def dwconv_tokens_bwd(x, H, W, w, dy, add_input, need_dx=True):
    """(dx, dw, db) of y = dw3x3(x) + b (+ x) on tokens x [B, H*W, C]."""
    require_cuda(dy)
    lib = load_library()
    x, dy = x.contiguous(), dy.contiguous()
    B, N, C = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = _ws(lib.tcx_dwconv_tokens_bwd_workspace_bytes(B, H, W, C), x)
    _chk(lib.tcx_dwconv_tokens_bwd(_ptr(x), _ptr(_d(w)), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), B, H, W, C, int(add_input), _ptr(ws),
                                   _stream()))
    return dx, dw, db


def attn_core_bwd(q, kv, dout, scale):
    """(dq, dkv) of out = softmax(q k^T * scale) v with kv = [k | v]."""
    require_cuda(q)
    lib = load_library()
    q, kv, dout = q.contiguous(), kv.contiguous(), dout.contiguous()
    B, Nq, d = q.shape
    Nk = kv.shape[1]
    assert d == 64 and kv.shape[2] == 128
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    ws = _ws(lib.tcx_attn_core_bwd_workspace_bytes(B, Nq, Nk), q)
    _chk(lib.tcx_attn_core_bwd(_ptr(q), _ptr(kv), _ptr(dout), scale, _ptr(dq), _ptr(dkv), B, Nq, Nk, _ptr(ws), _stream()))
    return dq, dkv


def flash_attn_train(q, kv, scale):
    """(out, lse) of softmax(q k^T * scale) v on the tcgen05 flash kernel; lse feeds flash_attn_bwd."""
    require_cuda(q)
    lib = load_library()
    q, kv = q.contiguous(), kv.contiguous()
    B, Nq, D = q.shape
    Nk = kv.shape[1]
    if D != 64 or kv.shape[2] != 128 or kv.shape[0] != B:
        raise RuntimeError("flash_attn: expected q [B,Nq,64] and kv [B,Nk,128], got %s %s" % (tuple(q.shape), tuple(kv.shape)))
    out = torch.empty_like(q)
    lse = torch.empty((B, Nq), dtype=torch.float32, device=q.device)
    ws = _ws(lib.tcx_flash_attn_workspace_bytes(B, Nk), q)
    _chk(lib.tcx_flash_attn_train_fwd(_ptr(q), _ptr(kv), _ptr(out), _ptr(lse), B, Nq, Nk, scale, _ptr(ws), _stream()))
    return out, lse


def flash_attn_bwd(q, kv, out, lse, dout, scale):
    """(dq, dkv) of the bridge attention core on the tcgen05 flash backward kernel."""
    require_cuda(q)
    lib = load_library()
    q, kv, out, dout = q.contiguous(), kv.contiguous(), out.contiguous(), dout.contiguous()
    B, Nq, _ = q.shape
    Nk = kv.shape[1]
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    ws = _ws(lib.tcx_flash_attn_bwd_workspace_bytes(B, Nq, Nk), q)
    _chk(lib.tcx_flash_attn_bwd(_ptr(q), _ptr(kv), _ptr(out), _ptr(lse), _ptr(dout), scale, _ptr(dq), _ptr(dkv), B, Nq, Nk, _ptr(ws),
                                _stream()))
    return dq, dkv


def ea_core(k, q, v):
    """softmax_channels(q) @ (softmax_tokens(k)^T v) on token-major fp32 k, q, v [B, N, C]."""
    require_cuda(k)
    lib = load_library()
    k, q, v = k.contiguous(), q.contiguous(), v.contiguous()
    B, N, C = k.shape
    out = torch.empty_like(k)
    ws = _ws(lib.tcx_ea_core_workspace_bytes(B, N, C), k)
    _chk(lib.tcx_ea_core_fwd(_ptr(k), _ptr(q), _ptr(v), _ptr(out), B, N, C, _ptr(ws), _stream()))
    return out


def ea_core_bwd(k, q, v, dout):
    require_cuda(k)
    lib = load_library()
    k, q, v, dout = k.contiguous(), q.contiguous(), v.contiguous(), dout.contiguous()
    B, N, C = k.shape
    dk, dq, dv = torch.empty_like(k), torch.empty_like(k), torch.empty_like(k)
    ws = _ws(lib.tcx_ea_core_workspace_bytes(B, N, C), k)
    _chk(lib.tcx_ea_core_bwd(_ptr(k), _ptr(q), _ptr(v), _ptr(dout), _ptr(dk), _ptr(dq), _ptr(dv), B, N, C, _ptr(ws), _stream()))
    return dk, dq, dv


ACT_NONE, ACT_HARDSWISH, ACT_SILU_SWISH = 0, 2, 4


def bn_act_train(x, w, b, rm, rv, eps, momentum, act):
    """(y, stat) of BatchNorm2d with batch statistics + activation on NHWC rows x [..., C]; updates rm / rv in place."""
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    stat = torch.empty(2 * C, dtype=torch.float32, device=x.device)
    ws = _ws(lib.tcx_bn_act_train_workspace_bytes(M, C), x)
    _chk(lib.tcx_bn_act_train_fwd(_ptr(x), _ptr(_d(w)), _ptr(_d(b)), _ptr(rm), _ptr(rv), eps, momentum, act, _ptr(y), _ptr(stat), M, C,
                                  _ptr(ws), _stream()))
    return y, stat


def bn_act_train_bwd(x, dy, stat, w, b, act):
    lib = load_library()
    x, dy = x.contiguous(), dy.contiguous()
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    dwdb = torch.empty((2, C), dtype=w.dtype, device=w.device)       # adjacent: the fold kernel writes both in place
    dw, db = dwdb[0], dwdb[1]
    ws = _ws(lib.tcx_bn_act_train_workspace_bytes(M, C), x)
    _chk(lib.tcx_bn_act_train_bwd(_ptr(x), _ptr(dy), _ptr(stat), _ptr(_d(w)), _ptr(_d(b)), act, _ptr(dx), _ptr(dw), _ptr(db), M, C,
                                  _ptr(ws), _stream()))
    return dx, dw, db


def dwconv3x3_nhwc(x, w, stride):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // stride + 1, (W - 1) // stride + 1, C), dtype=x.dtype, device=x.device)
    _chk(lib.tcx_dwconv3x3_nhwc_fwd(_ptr(x), _ptr(_d(w)), _ptr(y), B, H, W, C, stride, _stream()))
    return y


def dwconv3x3_nhwc_bwd(x, w, dy, stride, need_dx=True):
    lib = load_library()
    x, dy = x.contiguous(), dy.contiguous()
    B, H, W, C = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    ws = _ws(lib.tcx_dwconv3x3_nhwc_bwd_workspace_bytes(B, H, W, C, stride), x)
    _chk(lib.tcx_dwconv3x3_nhwc_bwd(_ptr(x), _ptr(_d(w)), _ptr(dy), _ptr(dx), _ptr(dw), B, H, W, C, stride, _ptr(ws), _stream()))
    return dx, dw


def coord_pool(x):
    require_cuda(x)
    lib = load_library()
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty((B, H + W, C), dtype=x.dtype, device=x.device)
    _chk(lib.tcx_coord_pool_fwd(_ptr(x), _ptr(y), B, H, W, C, _stream()))
    return y


def coord_pool_bwd(dy, H, W):
    lib = load_library()
    dy = dy.contiguous()
    B, _, C = dy.shape
    dx = torch.empty((B, H, W, C), dtype=dy.dtype, device=dy.device)
    _chk(lib.tcx_coord_pool_bwd(_ptr(dy), _ptr(dx), B, H, W, C, _stream()))
    return dx


def coord_gate(x, z):
    require_cuda(x)
    lib = load_library()
    x, z = x.contiguous(), z.contiguous()
    B, H, W, C = x.shape
    out = torch.empty_like(x)
    _chk(lib.tcx_coord_gate_fwd(_ptr(x), _ptr(z), _ptr(out), B, H, W, C, _stream()))
    return out


def coord_gate_bwd(x, z, dout):
    lib = load_library()
    x, z, dout = x.contiguous(), z.contiguous(), dout.contiguous()
    B, H, W, C = x.shape
    dx, dz = torch.empty_like(x), torch.empty_like(z)
    _chk(lib.tcx_coord_gate_bwd(_ptr(x), _ptr(z), _ptr(dout), _ptr(dx), _ptr(dz), B, H, W, C, _stream()))
    return dx, dz
