"""CUDA-graph runner for the inference forward.

One eager forward of the default model enqueues several hundred small kernels; replaying them as one CUDA graph
removes the per-launch host cost (streams and graphs instead of a tracing compiler).  The runner owns static
input/output buffers; ``run_host`` is the end-to-end call with pinned host buffers on both sides.
"""
import torch


class GraphRunner:
    """``microbatches`` > 1 splits the batch into equal slices whose forwards are captured on parallel streams: the
    model is a long chain of small, latency-bound kernels, so independent slices overlap on the GPU (eval-mode
    BatchNorm makes every image independent, so the result is identical to the single-chain forward)."""

    def __init__(self, model, batch, in_ch=1, size=224, device="cuda", warmup=3, microbatches=1):
        self.model = model.eval()
        self.device = torch.device(device)
        if batch % microbatches:
            raise ValueError("batch %d does not split into %d micro-batches" % (batch, microbatches))
        self.x = torch.zeros((batch, in_ch, size, size), device=self.device, dtype=torch.float32)
        self.stream = torch.cuda.Stream(device=self.device)
        self.side = [torch.cuda.Stream(device=self.device) for _ in range(microbatches - 1)]
        self.graph = torch.cuda.CUDAGraph()
        self.microbatches = microbatches
        from . import ops
        with torch.no_grad():
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):
                    self._forward()
            self.stream.synchronize()
            n0 = ops.launches()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.y = self._forward()
            self.kernels_per_replay = ops.launches() - n0
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def _forward(self):
        """The forward of the whole batch on the current stream (+ side streams for the extra micro-batches)."""
        if self.microbatches == 1:
            return self.model(self.x)
        per = self.x.shape[0] // self.microbatches
        main = torch.cuda.current_stream(self.device)
        outs = [None] * self.microbatches
        for i, st in enumerate(self.side):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                outs[i + 1] = self.model(self.x[(i + 1) * per:(i + 2) * per])
        outs[0] = self.model(self.x[:per])
        for st in self.side:
            main.wait_stream(st)
        if not hasattr(self, "_ybuf"):
            self._ybuf = torch.empty((self.x.shape[0],) + tuple(outs[0].shape[1:]), device=self.device, dtype=outs[0].dtype)
        for i, o in enumerate(outs):
            self._ybuf[i * per:(i + 1) * per].copy_(o)
        return self._ybuf

    def replay(self):
        """Enqueue one forward on the current stream (input = self.x, output = self.y)."""
        self.graph.replay()

    def run_device(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y

    def run_host(self, x_host, y_host):
        """End-to-end step with pinned host tensors: H2D input copy, graph replay, D2H logits copy.

        The D2H copy (29 MB of fp32 logits at bs16, ~0.5 ms over PCIe) is issued on a separate copy stream from one of
        two device staging buffers, so it overlaps the NEXT step's input copy and forward; within a step the order
        H2D -> forward -> D2H is kept by events.  Call ``drain()`` (or synchronise the device) before reading
        ``y_host`` or stopping a timer."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [torch.empty_like(self.y) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._drained = [None, None]
            self._step = 0
        main = torch.cuda.current_stream(self.device)
        i = self._step & 1
        self.x.copy_(x_host, non_blocking=True)
        self.graph.replay()
        if self._drained[i] is not None:
            main.wait_event(self._drained[i])          # the D2H that last read this staging buffer has finished
        self._stage[i].copy_(self.y, non_blocking=True)
        self._ready[i].record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ready[i])
            y_host.copy_(self._stage[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            self._drained[i] = ev
        self._step += 1
        return y_host

    def drain(self):
        """Make the current stream wait for every outstanding device-to-host copy of run_host."""
        if hasattr(self, "_copy_stream"):
            torch.cuda.current_stream(self.device).wait_stream(self._copy_stream)


class TrainStepGraph:
    """CUDA-graph runner for the training step of ``trainer.py:139-149``: forward (train mode) -> criterion -> backward ->
    [gradient all-reduce] -> [``clip_grad_norm_``] -> ``optimizer.step()``.

    An eager step enqueues several thousand small kernels and is bound by their host launch cost; replayed as graphs the same
    step is about twice as fast.  The step is captured in pieces so that every table the optimizer needs is built OUTSIDE
    capture from the static gradient addresses of the first graph:

    * graph A: forward + loss + backward (the gradient tensors are allocated inside this capture and keep their addresses);
    * with ``transception_b200.optim.FusedSGD``: graph B = the fused clip + SGD update (single process).  With
      ``torch.distributed`` initialised (world > 1) the backward is cut between encoder stages 2 and 3 (``MSViT._grad_cut``):
      graph A1 = forward + loss + backward down to the cut (decoder, bridge, stages 4 and 3: 98 % of the gradient elements),
      gathered into the head of the optimizer's flat bucket and all-reduced (NCCL, average) ASYNCHRONOUSLY while graph A2 = the
      backward of stages 2 and 1 runs; the small tail of the bucket follows, then graph C = the update read from the bucket.  The fused update also refreshes the fp16 GEMM copies the
      forward reads, and takes the learning rate from a device scalar: assign ``param_group['lr']`` as ``trainer.py:151-153``
      does and call ``step()`` — no re-capture;
    * with any other ``torch.optim`` optimizer: graph B = ``clip_grad_norm_`` + ``optimizer.step()`` as written by the caller
      (``shard.GradBucket`` all-reduce before it when world > 1).  ``torch.optim.SGD`` reads ``lr`` as a Python float at capture
      time, so a schedule needs ``recapture()`` there — use ``FusedSGD``.

    ``criterion(outputs, labels)`` must be sync-free (``transception_b200.losses.CeDiceLoss``; the reference's ``DiceLoss``
    reads ``.item()`` per class).  Usage::

        opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        runner = TrainStepGraph(net, CeDiceLoss(9), opt, batch=16, in_ch=1, size=224)
        loss = runner.step(image_batch, label_batch)        # device scalar, valid after the stream reaches it

    ``max_norm`` defaults to None like the reference (``--grad_clipping`` is off by default).
    """

    def __init__(self, model, criterion, optimizer, batch, in_ch=1, size=224, device="cuda", max_norm=None, warmup=3,
                 label_dtype=torch.int64, sample=None, single_graph=False):
        """``sample`` = (images, labels) of the first batch: the warm-up steps before capture are real training steps on it
        (otherwise they run on a zero batch)."""
        import torch.distributed as dist
        from .optim import FusedSGD
        from .shard import GradBucket
        self.model, self.criterion, self.optimizer = model.train(), criterion, optimizer
        self.device = torch.device(device)
        self.fused = isinstance(optimizer, FusedSGD)
        if self.fused and max_norm is not None:
            optimizer.max_norm = max_norm
        self.max_norm = max_norm
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket = GradBucket(model.parameters())
        # split backward / overlapped all-reduce: needs the fused optimizer's bucket and a model with the cut (MSViT)
        bb = getattr(model, "backbone", None)
        import os
        forced = os.environ.get("TCX_FORCE_OVERLAP") == "1" and dist.is_available() and dist.is_initialized()    # measurement aid
        self.overlap = self.fused and (self.world > 1 or forced) and bb is not None and hasattr(bb, "EARLY_MODULES")
        # overlap only, opt-in: the collectives are captured inside ONE graph of the whole step.  Measured at 2 GPUs: the same
        # step time as the separate graphs with eager collectives (21.14 vs 21.19 ms), and a profiler attached to the replay of a
        # graph that contains NCCL kernels hung — so the separate graphs are the default.
        self.single_graph = bool(single_graph)
        if self.overlap:
            early = [p for name in bb.EARLY_MODULES for p in getattr(bb, name).parameters()]
            optimizer.set_bucket_tail(early)
        self.x = torch.zeros((batch, in_ch, size, size), device=self.device, dtype=torch.float32)
        self.labels = torch.zeros((batch, size, size), device=self.device, dtype=label_dtype)
        self.loss = torch.zeros((), device=self.device)
        if sample is not None:
            self.x.copy_(sample[0])
            self.labels.copy_(sample[1])
        self._warmup = warmup
        self.recapture()

    # the phases of trainer.py:139-149
    def _fwd_bwd(self):
        from . import ops
        if not self.fused:
            # a torch optimizer moves the weights without telling the library: convert the fp16 GEMM copies again in every
            # forward (inside the captured graph too — otherwise a replay would keep multiplying with the weights of capture time)
            ops.invalidate_prepared(self.model)
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(self.x), self.labels)
        loss.backward()
        self.loss.copy_(loss.detach())

    def _fwd_bwd1(self):
        """Forward + loss + the backward down to the cut between encoder stages 2 and 3."""
        bb = self.model.backbone
        bb._grad_cut = []
        try:
            self.optimizer.zero_grad(set_to_none=True)
            loss = self.criterion(self.model(self.x), self.labels)
            loss.backward()
            self._cut = list(bb._grad_cut)
        finally:
            bb._grad_cut = None
        self.loss.copy_(loss.detach())

    def _bwd2(self):
        """The rest of the backward: the gradients that arrived at the cut continue into stages 2 and 1."""
        origs = [o for o, _ in self._cut]
        grads = [leaf.grad for _, leaf in self._cut]
        self._cut = None
        torch.autograd.backward(origs, grads)

    def _reduce(self, flat, async_op=False):
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            w.wait()
            w = None
        flat.mul_(1.0 / self.world)
        return w

    def _allreduce(self):
        """Average the gradients over the ranks: the optimizer's flat bucket (fused) or shard.GradBucket."""
        import torch.distributed as dist
        if self.world == 1:
            return
        if self.fused:
            flat = self._flat
            if dist.get_backend() == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.mul_(1.0 / self.world)
        else:
            self.bucket.allreduce()

    def _gather(self):
        self._flat = self.optimizer.gather_grads()

    def _update(self):
        if self.fused:
            self.optimizer.step(from_flat=self.world > 1 or self.overlap)
            return
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_norm, norm_type=2)
        self.optimizer.step()

    def eager_step(self):
        if getattr(self, "_captured", False):
            # an eager zero_grad(set_to_none=True) + backward would allocate NEW gradient tensors, while the graphs keep
            # reading and writing the ones they were captured with
            raise RuntimeError("TrainStepGraph: eager steps after capture break the graphs' static gradient buffers; "
                               "use step() / replay(), or recapture()")
        if self.overlap:
            self._fwd_bwd1()
            self._bwd2()
        else:
            self._fwd_bwd()
        if (self.world > 1 or self.overlap) and self.fused:
            self._gather()
        self._allreduce()
        self._update()

    def _capture(self, fn, prerun=True):
        if prerun:
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def recapture(self):
        """(Re)build the graphs from the current model / optimizer state.  ``max(warmup, 2)`` eager steps (they are the warm-up
        of the capture recipe) plus one more step completed while capturing are REAL training steps on the current batch
        (``steps_done`` counts them); recording itself executes nothing.  Order: capture graph A, replay it so the static
        gradient tensors hold real values, [capture + replay the gather graph, all-reduce], capture the update graph and replay
        it once — every graph is replayed exactly once per step, so captured training equals eager training bit for bit."""
        from . import ops
        self.steps_done = getattr(self, "steps_done", 0)
        self._captured = False
        if self.fused:
            self.optimizer.freeze_tables(False)       # tables frozen for an earlier capture are rebuilt by the eager steps below
        n0 = ops.launches()
        self.eager_step()
        self.kernels_per_step = ops.launches() - n0
        self.first_loss = self.loss.clone()
        for _ in range(max(self._warmup - 1, 1)):
            self.eager_step()
        self.steps_done += max(self._warmup, 2)
        torch.cuda.synchronize(self.device)
        if self.overlap:
            return self._recapture_overlap()
        ga = self._capture(self._fwd_bwd, prerun=False)
        ga.replay()
        graphs = [ga]
        if self.fused:
            # the optimizer's tables (device arrays of parameter / gradient / momentum / fp16-copy pointers) are built here, eagerly,
            # from the static gradient tensors of graph A: nothing is uploaded under capture
            for gi, group in enumerate(self.optimizer.param_groups):
                self.optimizer._group_table(gi, group)
        if self.world > 1 and self.fused:
            gg = self._capture(self._gather, prerun=False)
            gg.replay()
            graphs.append(gg)
        self._graphs = graphs
        if self.world > 1:
            self._allreduce()
        gu = self._capture(self._update, prerun=False)
        gu.replay()
        graphs.append(gu)
        self.steps_done += 1
        if self.fused:
            ops.bump_raw_generation()
        self._captured = True

    def _overlap_step_in_one_graph(self):
        """The whole data-parallel step as it is recorded into ONE graph: the head all-reduce is a side branch of the graph that
        runs beside the backward of stages 2-1 (no graph boundary, so nothing drains at the cut)."""
        cur = torch.cuda.current_stream(self.device)
        side = self._ar_stream
        self._fwd_bwd1()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._flat_head = self.optimizer.gather_grads(0)
            self._reduce(self._flat_head)
        self._bwd2()
        self._flat_tail = self.optimizer.gather_grads(1)
        self._reduce(self._flat_tail)
        cur.wait_stream(side)
        self._update()

    def _recapture_overlap(self):
        """world > 1 with the fused optimizer.  Default: A1 (forward + loss + backward to the cut), A2 (backward of stages 2-1), the
        two bucket gathers and the update as five graphs with eager collectives between them; see ``replay``.  ``single_graph``:
        one graph with the collectives captured inside it — the optimizer's tables keep the structure of the last eager step and
        get the addresses of the gradient tensors allocated by the capture afterwards (FusedSGD.freeze_tables /
        refresh_grad_ptrs); falls back to the five graphs if that capture fails."""
        from . import ops
        if self.single_graph:
            if not hasattr(self, "_ar_stream"):
                self._ar_stream = torch.cuda.Stream(self.device)
            self.optimizer.freeze_tables(True)
            try:
                g = self._capture(self._overlap_step_in_one_graph, prerun=False)
                self.optimizer.refresh_grad_ptrs()
            except Exception as e:  # noqa: BLE001 — e.g. a collective that cannot be captured on this stack
                import warnings
                warnings.warn("TrainStepGraph: capturing the data-parallel step as one graph failed (%s: %s); using separate "
                              "graphs with eager collectives" % (type(e).__name__, str(e)[:200]))
                self.single_graph = False
                g = None
                self.optimizer.freeze_tables(False)
            if g is not None:
                # the tables stay frozen: the graph holds their addresses, they must not be rebuilt while it is in use
                g.replay()
                self._graphs = [g]
                self.steps_done += 1
                ops.bump_raw_generation()
                self._captured = True
                return
        a1 = self._capture(self._fwd_bwd1, prerun=False)
        a2 = self._capture(self._bwd2, prerun=False)       # continues the autograd graph recorded (not executed) by A1's capture
        a1.replay()
        a2.replay()                                         # the static gradient tensors of both pieces hold real values now
        for gi, group in enumerate(self.optimizer.param_groups):
            self.optimizer._group_table(gi, group)
        g1 = self._capture(lambda: setattr(self, "_flat_head", self.optimizer.gather_grads(0)), prerun=False)
        g2 = self._capture(lambda: setattr(self, "_flat_tail", self.optimizer.gather_grads(1)), prerun=False)
        g1.replay()
        g2.replay()
        self._reduce(self._flat_head)
        self._reduce(self._flat_tail)
        gu = self._capture(self._update, prerun=False)
        gu.replay()
        self._graphs = [a1, g1, a2, g2, gu]
        self.steps_done += 1
        ops.bump_raw_generation()
        self._captured = True

    def _replay_overlap(self):
        if self.single_graph:
            self.optimizer.sync_lr()
            self._graphs[0].replay()
            return
        a1, g1, a2, g2, gu = self._graphs
        self.optimizer.sync_lr()
        a1.replay()
        g1.replay()
        work = self._reduce(self._flat_head, async_op=True)      # on the collective's stream, behind A1 + the head gather
        a2.replay()                                              # stages 2-1 backward runs beside it
        g2.replay()
        self._reduce(self._flat_tail)
        if work is not None:
            work.wait()                                          # the current stream waits for the head all-reduce
        gu.replay()

    # the pieces of one data-parallel step, for checks that look at the gradients between them (bench.py)
    def replay_backward(self):
        """Forward + loss + the whole backward: the static ``p.grad`` tensors hold this rank's own gradients afterwards.
        (The piecewise calls need separate graphs: ``single_graph=False``.)"""
        if self.overlap and self.single_graph:
            raise RuntimeError("TrainStepGraph: the step is one graph; build the runner with single_graph=False to replay its pieces")
        if self.overlap:
            self._graphs[0].replay()
            self._graphs[2].replay()
        else:
            self._graphs[0].replay()

    def replay_reduce(self):
        """Gather into the flat bucket and average it over the ranks (no overlap); returns the bucket."""
        if self.overlap:
            self._graphs[1].replay()
            self._graphs[3].replay()
            self._reduce(self._flat_head)
            self._reduce(self._flat_tail)
        elif self.world > 1 and self.fused:
            self._graphs[1].replay()
            self._allreduce()
        return self.optimizer._tables[0]["flat"] if self.fused else None

    def replay_update(self):
        from . import ops
        self._graphs[-1].replay()
        if self.fused:
            ops.bump_raw_generation()

    def replay(self):
        """One training step on the batch currently in ``self.x`` / ``self.labels``."""
        from . import ops
        if self.overlap:
            self._replay_overlap()
            ops.bump_raw_generation()
            self.steps_done += 1
            return
        if self.fused:
            self.optimizer.sync_lr()         # a changed param_group['lr'] reaches the device scalar the update graph reads
        self._graphs[0].replay()
        if self.world > 1:
            if self.fused:
                self._graphs[1].replay()
            self._allreduce()
        self._graphs[-1].replay()
        if self.fused:
            ops.bump_raw_generation()
        self.steps_done += 1

    def step(self, x, labels):
        """Copy a batch in (host pinned or device tensors), run one step, return the device loss scalar."""
        self.x.copy_(x, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.replay()
        return self.loss
