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
    [gradient all-reduce] -> ``clip_grad_norm_`` -> ``optimizer.step()``.

    An eager step enqueues ~6 600 small kernels and is bound by their host launch cost; replayed as a graph the same step is
    about twice as fast.  With one process the whole step is ONE graph; with ``torch.distributed`` initialised (world > 1) the
    step is a forward+loss+backward graph, an eager flat-bucket all-reduce (``shard.GradBucket``: the gradient tensors are
    allocated once inside the first capture, so the bucket reads the same addresses every step) and a clip+optimizer graph.

    ``criterion(outputs, labels)`` must be sync-free (``transception_b200.losses.CeDiceLoss``; the reference's ``DiceLoss``
    reads ``.item()`` per class).  Usage::

        runner = TrainStepGraph(net, CeDiceLoss(9), optimizer, batch=16, in_ch=1, size=224, max_norm=5)
        loss = runner.step(image_batch, label_batch)        # device scalar, valid after the stream reaches it

    The learning-rate schedule of ``trainer.py:151-153`` writes ``param_group['lr']`` on the host; SGD reads it as a Python
    float at capture time, so call ``recapture()`` when it changes (or use a tensor ``lr`` with ``capturable`` optimizers).
    """

    def __init__(self, model, criterion, optimizer, batch, in_ch=1, size=224, device="cuda", max_norm=5.0, warmup=3,
                 label_dtype=torch.int64, sample=None):
        """``sample`` = (images, labels) of the first batch: the warm-up steps before capture are real training steps on it
        (otherwise they run on a zero batch)."""
        import torch.distributed as dist
        from .shard import GradBucket
        self.model, self.criterion, self.optimizer = model.train(), criterion, optimizer
        self.device = torch.device(device)
        self.max_norm = max_norm
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bucket = GradBucket(model.parameters())
        self.x = torch.zeros((batch, in_ch, size, size), device=self.device, dtype=torch.float32)
        self.labels = torch.zeros((batch, size, size), device=self.device, dtype=label_dtype)
        self.loss = torch.zeros((), device=self.device)
        if sample is not None:
            self.x.copy_(sample[0])
            self.labels.copy_(sample[1])
        self._warmup = warmup
        self.recapture()

    # the three phases of trainer.py:139-149
    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(self.x), self.labels)
        loss.backward()
        self.loss.copy_(loss.detach())

    def _update(self):
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_norm, norm_type=2)
        self.optimizer.step()

    def eager_step(self):
        if getattr(self, "_captured", False):
            # an eager zero_grad(set_to_none=True) + backward would allocate NEW gradient tensors, while the graphs keep
            # reading and writing the ones they were captured with
            raise RuntimeError("TrainStepGraph: eager steps after capture break the graphs' static gradient buffers; "
                               "use step() / replay(), or recapture()")
        self._fwd_bwd()
        if self.world > 1:
            self.bucket.allreduce()
        self._update()

    def _capture(self, fn):
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
        """(Re)build the graphs from the current model / optimizer state.  ``max(warmup, 2)`` eager steps plus the one step the
        capture recipe runs on a side stream before recording are REAL training steps on the current batch (``steps_done``
        counts them); recording itself executes nothing."""
        from . import ops
        self.steps_done = getattr(self, "steps_done", 0)
        self._captured = False
        n0 = ops.launches()
        self.eager_step()
        self.kernels_per_step = ops.launches() - n0
        self.first_loss = self.loss.clone()
        for _ in range(max(self._warmup - 1, 1)):
            self.eager_step()
        self.steps_done += max(self._warmup, 2) + 1
        torch.cuda.synchronize(self.device)
        if self.world == 1:
            self._graphs = (self._capture(self.eager_step),)
        else:
            ga = self._capture(self._fwd_bwd)
            self.bucket.allreduce()
            self._graphs = (ga, self._capture(self._update))
        self._captured = True

    def replay(self):
        """One training step on the batch currently in ``self.x`` / ``self.labels``."""
        self._graphs[0].replay()
        if len(self._graphs) > 1:
            self.bucket.allreduce()
            self._graphs[1].replay()
        self.steps_done += 1

    def step(self, x, labels):
        """Copy a batch in (host pinned or device tensors), run one step, return the device loss scalar."""
        self.x.copy_(x, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.replay()
        return self.loss
