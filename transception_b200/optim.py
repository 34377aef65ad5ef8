"""Fused SGD for the training row: the optimizer step of ``trainer.py:125,148`` on the library's multi-tensor kernels.

``FusedSGD`` takes the arguments of ``torch.optim.SGD`` (the reference constructs ``optim.SGD(model.parameters(), lr=base_lr,
momentum=0.9, weight_decay=0.0001)``) and produces the same update, with three differences in HOW it runs:

* one launch updates every parameter (plus one for the optional gradient-norm clip) instead of ~100 ``multi_tensor_apply``
  launches, and the same kernel refreshes the fp16 GEMM copies the forward keeps (``ops.prepare_weight``), so no per-weight
  conversion kernels run in the next forward;
* the learning rate lives in a device scalar: ``param_group['lr'] = lr_`` (``trainer.py:151-153`` does this every iteration)
  is picked up by ``sync_lr()`` as a 4-byte copy — a captured CUDA graph of the step follows the schedule without re-capture;
* ``gather_grads()`` / ``step(from_flat=True)`` expose the flat gradient bucket for the data-parallel all-reduce: the
  gradients are gathered once into one buffer, reduced in place, and the update reads them from there (no ``torch.cat``, no
  unflatten copy).

Gradient clipping (``max_norm``, the ``clip_grad_norm_`` of the reference's ``--grad_clipping`` option) is applied inside the
update (``g * coef``); ``p.grad`` itself is left unscaled.  There is no CPU path.
"""
import ctypes

import torch

from . import ops


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, *, max_norm=None):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError("FusedSGD: negative hyper-parameter")
        if dampening != 0.0 or nesterov:
            raise NotImplementedError("FusedSGD is built for dampening=0, nesterov=False (the reference's optim.SGD call)")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov))
        if max_norm is not None and len(self.param_groups) != 1:
            raise NotImplementedError("FusedSGD: max_norm needs a single parameter group")
        self.max_norm = max_norm
        self._tables, self._lr_t, self._lr_seen = {}, {}, {}
        self.norm_coef = None           # device [norm, coef] of the last step (max_norm set)
        self._tail_ids = None           # set_bucket_tail: parameters placed at the END of the flat bucket
        self._frozen = False            # freeze_tables: the tables are used as they are (graph capture with late-bound gradient pointers)

    def freeze_tables(self, on=True):
        """While frozen, the device tables built by the last eager step are used as they are, whatever ``p.grad`` currently is:
        a graph capture that contains the gathers / the update records the tables' ADDRESSES, and ``refresh_grad_ptrs()`` writes
        the addresses of the gradient tensors allocated inside that capture into them afterwards."""
        if on and not self._tables:
            raise RuntimeError("FusedSGD.freeze_tables: no tables yet (run one eager step first)")
        if self._frozen and not on:
            self._tables.clear()          # their gradient pointers belong to a captured graph: rebuild from the next eager step
        self._frozen = bool(on)

    def refresh_grad_ptrs(self):
        """Rewrite the gradient-pointer tables in place from the current ``p.grad`` tensors of the SAME parameters (same order)."""
        for tab in self._tables.values():
            used = tab["used"]
            if any(p.grad is None or p.grad.numel() != p.numel() or not p.grad.is_contiguous() or p.grad.dtype != torch.float32
                   for p in used):
                raise RuntimeError("FusedSGD.refresh_grad_ptrs: a parameter of the table has no (contiguous fp32) gradient")
            tab["g"].copy_(torch.tensor([p.grad.data_ptr() for p in used], dtype=torch.int64))
            tab["key"] = tuple((p.data_ptr(), p.grad.data_ptr()) for p in used)
            tab["keep"] = (tab["keep"][0], [p.grad for p in used])

    def set_bucket_tail(self, params):
        """Order the flat gradient bucket as [all other parameters | ``params``] (each part in registration order).  The data-parallel
        runner all-reduces the head while the backward of the tail's layers is still running (``gather_grads(part)``)."""
        self._tail_ids = {id(p) for p in params}
        self._tables.clear()

    # ---- tables ---------------------------------------------------------------------------------------------------------------
    def _group_table(self, gi, group):
        if self._frozen:
            return self._tables.get(gi)
        used = [p for p in group["params"] if p.grad is not None]
        n_head = len(used)
        if self._tail_ids:
            head = [p for p in used if id(p) not in self._tail_ids]
            used = head + [p for p in used if id(p) in self._tail_ids]
            n_head = len(head)
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in used)
        tab = self._tables.get(gi)
        if tab is not None and tab["key"] == key:
            return tab
        if not used:
            return None
        dev = used[0].device
        for p in used:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                raise RuntimeError("FusedSGD: parameters and gradients must be contiguous fp32 CUDA tensors")
        chunk = ops.load_library().tcx_mt_chunk()
        numel = [p.numel() for p in used]
        offs, total = [], 0
        for n in numel:
            offs.append(total)
            total += (n + 3) // 4 * 4                   # 16-byte aligned slices of the flat bucket
        old = tab or {}
        flat = old.get("flat")
        if flat is None or flat.numel() != total:
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
        bufs = []
        for p in used:
            st = self.state[p]
            if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            bufs.append(st["momentum_buffer"])
        w16, special = [], []
        for p in used:
            pc = ops.prepared_copy(p, track=True)
            if pc is None:
                w16.append(0)
            elif pc[1] is not None:
                w16.append(0)
                special.append(p)                       # K-permuted patchify-conv copy: refreshed by its own conversion kernel
            else:
                w16.append(pc[0].data_ptr())
        blocks, head_blocks = [], 0
        for t, n in enumerate(numel):
            blocks.extend((t, c) for c in range((n + chunk - 1) // chunk))
            if t == n_head - 1:
                head_blocks = len(blocks)
        head_elems = offs[n_head] if n_head < len(used) else total

        def i64(vals):
            return torch.tensor(vals, dtype=torch.int64).to(dev)
        tab = {
            "key": key, "used": used, "n": len(used), "total": total, "flat": flat, "special": special,
            "p": i64([p.data_ptr() for p in used]), "g": i64([p.grad.data_ptr() for p in used]),
            "gflat": i64([flat.data_ptr() + 4 * o for o in offs]), "buf": i64([b.data_ptr() for b in bufs]), "w16": i64(w16),
            "numel": i64(numel), "offs": i64(offs), "blocks": torch.tensor(blocks, dtype=torch.int32).to(dev),
            "nblocks": len(blocks), "part": torch.empty(len(blocks), dtype=torch.float32, device=dev),
            "head_blocks": head_blocks, "head_elems": head_elems,
            "keep": (bufs, [p.grad for p in used]),
        }
        if gi not in self._lr_t:
            self._lr_t[gi] = torch.full((), float(group["lr"]), dtype=torch.float32, device=dev)
            self._lr_seen[gi] = float(group["lr"])
        self._tables[gi] = tab
        return tab

    def sync_lr(self):
        """Copy changed ``param_group['lr']`` values into their device scalars (call outside graph capture)."""
        for gi, group in enumerate(self.param_groups):
            if gi in self._lr_t and self._lr_seen[gi] != float(group["lr"]):
                self._lr_t[gi].fill_(float(group["lr"]))
                self._lr_seen[gi] = float(group["lr"])

    # ---- data-parallel bucket ---------------------------------------------------------------------------------------------------
    def gather_grads(self, part=None):
        """Gather the used gradients into the flat bucket (one launch) and return the slice written: the all-reduce operand.
        ``part`` = 0 / 1: only the head / the tail of the bucket (``set_bucket_tail``)."""
        lib = ops.load_library()
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedSGD.gather_grads needs a single parameter group")
        tab = self._group_table(0, self.param_groups[0])
        if tab is None:
            return None
        b0, b1 = 0, tab["nblocks"]
        e0, e1 = 0, tab["total"]
        if part == 0:
            b1, e1 = tab["head_blocks"], tab["head_elems"]
        elif part == 1:
            b0, e0 = tab["head_blocks"], tab["head_elems"]
        if b1 > b0:
            ops._chk(lib.tcx_mt_gather(tab["g"].data_ptr(), tab["numel"].data_ptr(), tab["offs"].data_ptr(),
                                       tab["blocks"].data_ptr() + 8 * b0, b1 - b0, tab["flat"].data_ptr(), ops._stream()))
        return tab["flat"][e0:e1]

    # ---- step -------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, from_flat=False):
        """One SGD update.  ``from_flat``: read the gradients from the flat bucket filled by ``gather_grads()`` (and reduced
        across ranks by the caller) instead of ``p.grad``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = ops.load_library()
        capturing = torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            tab = self._group_table(gi, group)
            if tab is None:
                continue
            if not capturing:
                self.sync_lr()
            gsrc = tab["gflat"] if from_flat else tab["g"]
            coef_ptr = None
            if self.max_norm is not None:
                if self.norm_coef is None or self.norm_coef.device != tab["flat"].device:
                    self.norm_coef = torch.zeros(2, dtype=torch.float32, device=tab["flat"].device)
                ops._chk(lib.tcx_mt_sqnorm(gsrc.data_ptr(), tab["numel"].data_ptr(), tab["blocks"].data_ptr(), tab["nblocks"],
                                           tab["part"].data_ptr(), float(self.max_norm), self.norm_coef.data_ptr(), ops._stream()))
                coef_ptr = self.norm_coef.data_ptr() + 4
            ops._chk(lib.tcx_mt_sgd(tab["p"].data_ptr(), gsrc.data_ptr(), tab["buf"].data_ptr(), tab["w16"].data_ptr(),
                                    tab["numel"].data_ptr(), tab["blocks"].data_ptr(), tab["nblocks"], self._lr_t[gi].data_ptr(),
                                    ctypes.c_void_p(coef_ptr), float(group["momentum"]), float(group["weight_decay"]), ops._stream()))
            for p in tab["special"]:
                ops.refresh_prepared(p)
        if not capturing:
            ops.bump_raw_generation()       # prepared copies outside the tables are stale now
        return loss
