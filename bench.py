#!/usr/bin/env python
"""bench.py — headline measurement for the TransCeption hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|forward]

Headline (default, `--mode train`) = BASELINE.json's metric, "images/sec fwd+bwd @224x224 bs16": one step = forward in train mode
+ 0.4 CE + 0.6 Dice + backward + SGD(momentum 0.9, weight decay 1e-4) on one batch of 16 synthetic 224x224 images per GPU
(trainer.py:139-149; the reference's --grad_clipping is off by default, so no clip), seeded random-init weights; N GPUs =
N ranks x 16 images (weak scaling) with one NCCL all-reduce (average) of the flat gradient bucket per step.

Printed JSON (one line, rank 0):
  value          images/s with the batch resident in HBM (the step replayed as CUDA graphs, CUDA-event timed, max over ranks)
  e2e            images/s through the public call TrainStepGraph.step(pinned images, pinned labels): H2D copies + step + D2H loss
  roofline       the kernel with the largest share of the step, timed live with CUDA events; roofline_other_kernels the next ones
  roofline_step  the whole step against SURVEY section 8(d)'s algorithmic FLOPs / bytes
  forward_step   the inference forward (BASELINE configs[1]) at the same batch: device-resident and end to end
  gpu_baseline   the oracle (PyTorch restatement of the reference) on this same B200, TF32 on, eager and CUDA graph
  cpu_baseline   the same oracle on this box's host cores (bounded sample) + bs16 parity of our numbers against it
`--mode forward` prints the round-1 forward line (configs[1]) as the line's own metric.  `--impl reference` times the CPU
implementation of the selected mode as its own arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec fwd @224x224 bs16 (TransCeption MSTransception, fp32 IO)"
BATCH, SIZE, NCLS, IN_CH = 16, 224, 9, 1


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def _make_inputs(torch, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, IN_CH, SIZE, SIZE, generator=g) * 2 - 1


MODEL = "MSTransception"     # --model Transception times the networks/Transception.py variant (non-headline)


def _model_cls():
    import transception_b200
    return getattr(transception_b200, MODEL)


def cpu_reference_time(torch, steps, warmup, budget_s=150.0):
    """Time the CPU oracle forward (the reference's algorithm on host cores). Returns (img/s, cores, sample, ms/step)."""
    if MODEL == "Transception":
        from oracle import transception_oracle as O
    else:
        from oracle import mstr_oracle as O
    MSTransception = _model_cls()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    sd = {k: v for k, v in MSTransception(num_classes=NCLS, image_size=SIZE).state_dict().items()}
    bs = BATCH
    x = _make_inputs(torch, bs, 0)
    with torch.no_grad():
        t0 = time.perf_counter(); O.forward(sd, x); t1 = time.perf_counter() - t0       # warm + size probe
        total = max(1, steps + warmup - 1)
        while bs > 1 and t1 * (bs / BATCH) * total > budget_s:
            bs //= 2
        x = x[:bs].contiguous()
        for _ in range(max(0, warmup - 1)):
            O.forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(sd, x)
        dt = time.perf_counter() - t0
    sample = "%d timed forward(s) of bs%d 224x224 (of the bs16 workload), torch %s CPU fp32, %d threads" % (
        steps, bs, torch.__version__, torch.get_num_threads())
    return bs * steps / dt, cores, sample, dt / steps * 1e3


TRAIN_METRIC = "images/sec fwd+bwd @224x224 bs16 (TransCeption MSTransception train step: forward + 0.4 CE + 0.6 Dice + backward + SGD)"
TRAIN_WORKLOAD = ("TransCeption Synapse 224x224 bs16 train step per GPU (BASELINE configs[2]/[3] shape): forward + 0.4 CE + 0.6 Dice + "
                  "backward + SGD(momentum 0.9, wd 1e-4), 9 classes")
FWD_WORKLOAD = "TransCeption Synapse 224x224 bs16 fp32 forward (BASELINE configs[1]), per-GPU batch 16, 9 class logits"
# SURVEY section 8(d) / BASELINE.md section 2: algorithmic work of the whole model, per image
FWD_GFLOP_PER_IMG = 16.88            # forward; a train step is 3x (backward = dgrad + wgrad)
FWD_MB_PER_IMG_BF16 = 34.0           # activation traffic at fused-kernel boundaries (encoder + bridge), bf16
WEIGHT_MB_BF16 = 72.7                # once per batch


def _train_inputs(torch, batch, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, IN_CH, SIZE, SIZE, generator=g) * 2 - 1
    labels = torch.randint(0, NCLS, (batch, SIZE, SIZE), generator=g)
    return x, labels


def cpu_reference_train_time(torch, steps, warmup, budget_s=150.0):
    """The reference's train step (trainer.py:139-149) as restated by the oracle + torch autograd + torch.optim.SGD on the host
    cores, on a bounded sample (batch halved until the run fits the budget).  Returns (img/s, cores, sample, ms/step)."""
    from oracle import loss_oracle as LO
    from oracle import mstr_oracle as O
    MSTransception = _model_cls()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone())
          for k, v in MSTransception(num_classes=NCLS, image_size=SIZE).state_dict().items()}
    leaves = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.SGD(leaves, lr=0.05, momentum=0.9, weight_decay=1e-4)
    O.BN_TRAIN = True

    def step(x, labels):
        opt.zero_grad(set_to_none=True)
        loss = LO.ce_dice(O.forward(sd, x), labels, NCLS)[0]
        loss.backward()
        opt.step()
        return float(loss.detach())

    bs = 2
    x, labels = _train_inputs(torch, BATCH, 0)
    t0 = time.perf_counter(); step(x[:bs], labels[:bs]); t1 = time.perf_counter() - t0
    total = max(1, steps + warmup - 1)
    while bs < BATCH and t1 * 2 * total < budget_s:
        bs *= 2; t1 *= 2
    for _ in range(max(0, warmup - 1)):
        step(x[:bs], labels[:bs])
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x[:bs], labels[:bs])
    dt = time.perf_counter() - t0
    O.BN_TRAIN = False
    sample = "%d timed train step(s) of bs%d 224x224 (of the bs16 workload), torch %s CPU fp32 autograd + SGD, %d threads" % (
        steps, bs, torch.__version__, torch.get_num_threads())
    return bs * steps / dt, cores, sample, dt / steps * 1e3


def cpu_train_parity(torch, first_loss_ours):
    """bs16 parity inside the cpu_baseline leg: the loss of the FIRST train step (same seeded weights, same synthetic batch)
    from the CPU oracle against the one our step reported.  One CPU forward in train mode (no backward)."""
    from oracle import loss_oracle as LO
    from oracle import mstr_oracle as O
    MSTransception = _model_cls()
    torch.manual_seed(1234)
    sd = {k: v.clone() for k, v in MSTransception(num_classes=NCLS, image_size=SIZE).state_dict().items()}
    x, labels = _train_inputs(torch, BATCH, 0)
    O.BN_TRAIN = True
    try:
        with torch.no_grad():
            ref = float(LO.ce_dice(O.forward(sd, x), labels, NCLS)[0])
    finally:
        O.BN_TRAIN = False
    return {"first_step_loss_oracle_bs16": ref, "first_step_loss_ours_bs16": first_loss_ours,
            "rel_diff": abs(first_loss_ours - ref) / abs(ref), "tolerance": 2e-3}


def gpu_baseline_leg(torch, dev, mode):
    """The honest GPU opponent: the oracle — a plain-PyTorch restatement of the reference — on
    this same B200 with TF32 matmuls/convs ON, eager and replayed as a CUDA graph.  Test/bench infrastructure like cpu_baseline:
    nothing of it is on the product path."""
    from oracle import loss_oracle as LO
    from oracle import mstr_oracle as O
    MSTransception = _model_cls()
    out = {"what": "oracle port (PyTorch ops = the reference's own op sequence) on cuda:0, TF32 on", "batch": BATCH}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        torch.manual_seed(1234)
        sd = {k: v.to(dev) for k, v in MSTransception(num_classes=NCLS, image_size=SIZE).state_dict().items()}
        xh, lh = _train_inputs(torch, BATCH, 0)
        x, labels = xh.to(dev), lh.to(dev)

        def timed(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n

        def graphed(fn):
            st = torch.cuda.Stream(dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream(dev).wait_stream(st)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g.replay

        with torch.no_grad():
            f = lambda: O.forward(sd, x)      # noqa: E731
            ms = timed(f, 10)
            out["forward_eager_ms"], out["forward_eager_img_s"] = ms, BATCH / ms * 1e3
            try:
                ms = timed(graphed(f), 10)
                out["forward_graph_ms"], out["forward_graph_img_s"] = ms, BATCH / ms * 1e3
            except Exception as e:  # noqa: BLE001
                out["forward_graph_error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
        if mode == "train":
            leaves = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
            params = [v for v in leaves.values() if v.requires_grad]
            opt = torch.optim.SGD(params, lr=0.05, momentum=0.9, weight_decay=1e-4)
            O.BN_TRAIN = True

            def step():
                opt.zero_grad(set_to_none=True)
                loss = LO.ce_dice(O.forward(leaves, x), labels, NCLS)[0]
                loss.backward()
                opt.step()
            try:
                ms = timed(step, 5)
                out["train_eager_ms"], out["train_eager_img_s"] = ms, BATCH / ms * 1e3
                try:
                    ms = timed(graphed(step), 5)
                    out["train_graph_ms"], out["train_graph_img_s"] = ms, BATCH / ms * 1e3
                except Exception as e:  # noqa: BLE001
                    out["train_graph_error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
            finally:
                O.BN_TRAIN = False
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        torch.cuda.empty_cache()
    return out


def _traffic_table():
    """DRAM bytes per launch of the profiled kernels from the committed ncu --set full captures (profiles/r02_traffic.json:
    {kernel: {"dram_bytes_per_launch": ..., "source": "profiles/..."}}); absent entries report traffic = null."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p))
    except (OSError, ValueError):
        return {}


def train_step_leg(torch, dev, world, rank, K, W, with_kernel_legs=True):
    """fwd + loss + bwd (+ gradient all-reduce when world > 1) + SGD step at bs16 per GPU, every FLOP of the model, of the loss
    and of the optimizer on the library's kernels (autograd nodes of transception_b200/autograd.py, optim.FusedSGD).  Timed with
    CUDA events over graph replays (runtime.TrainStepGraph: the repo's public training runner)."""
    import torch.distributed as dist
    from transception_b200 import ops
    from transception_b200.losses import CeDiceLoss
    from transception_b200.optim import FusedSGD
    from transception_b200.runtime import TrainStepGraph
    MSTransception = _model_cls()
    torch.manual_seed(1234)
    net = MSTransception(num_classes=NCLS, image_size=SIZE).to(dev).train()
    opt = FusedSGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    xh, lh = _train_inputs(torch, BATCH, rank)
    pdl_chain = ops.set_flag("pdl_chain", 0)      # read the flag (set_flag returns the previous value) ...
    ops.set_flag("pdl_chain", pdl_chain)          # ... and put it back
    runner = TrainStepGraph(net, CeDiceLoss(NCLS), opt, BATCH, IN_CH, SIZE, device=dev, warmup=W, sample=(xh, lh))
    launches = runner.kernels_per_step
    x, labels, loss_buf, run = runner.x, runner.labels, runner.loss, runner.replay
    first_loss = float(runner.first_loss)
    run()
    torch.cuda.synchronize(dev)
    grads_match = weights_match = None
    for _ in range(2):
        run()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    # end to end through the public call: the step's batch comes from pinned host memory and the loss is read back, every step
    xp, lp = xh.pin_memory(), lh.pin_memory()
    lossh = torch.zeros((), dtype=torch.float32).pin_memory()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e2.record()
    for _ in range(K):
        lossh.copy_(runner.step(xp, lp), non_blocking=True)
    e3.record()
    torch.cuda.synchronize(dev)
    e2e_ms = e2.elapsed_time(e3)
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # data-parallel training keeps the replicas identical: every rank's weights after the same number of steps
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        allw = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allw, chk)
        weights_match = all(torch.equal(allw[0], c) for c in allw)
    ms, e2e_ms = t.tolist()
    last_loss = float(lossh)
    if world > 1:
        # the all-reduced gradient must equal the mean of the ranks' own gradients (identical weights,
        # rank-specific data, per-rank BatchNorm statistics) and be identical on every rank afterwards.  The check replays the
        # step's pieces; a runner that holds the whole step as ONE graph (single_graph=True) is replaced by one with separate
        # graphs for it (the timing is over: the optimizer's tables may be re-built)
        overlap_kind = "one graph, collectives captured" if getattr(runner, "single_graph", False) and getattr(runner, "overlap", False) else "separate graphs"
        if getattr(runner, "single_graph", False) and getattr(runner, "overlap", False):
            runner = TrainStepGraph(net, CeDiceLoss(NCLS), opt, BATCH, IN_CH, SIZE, device=dev, warmup=1, sample=(xh, lh), single_graph=False)
        runner.replay_backward()            # forward + loss + backward graph(s): write the static gradient tensors
        torch.cuda.synchronize(dev)
        tab = opt._tables[0]
        with_grad = tab["used"]
        pick = [0, len(with_grad) // 3, 2 * len(with_grad) // 3, len(with_grad) - 1]
        own = torch.cat([with_grad[i].grad.flatten()[:4096].clone() for i in pick])
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        want = torch.stack(gathered).mean(0)
        runner.replay_reduce()              # gather into the flat bucket + all-reduce (average)
        offs = tab["offs"].tolist()
        got = torch.cat([tab["flat"][offs[i]:offs[i] + min(4096, with_grad[i].numel())] for i in pick])
        mean_ok = bool((got - want).abs().max() <= 1e-6 * want.abs().max() + 1e-12)
        chk = torch.stack([tab["flat"].double().sum(), tab["flat"].double().abs().sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        grads_match = mean_ok and all(torch.equal(allc[0], c) for c in allc)
        runner.replay_update()              # SGD update from the reduced bucket
        torch.cuda.synchronize(dev)
    # roofline legs: the kernels with the largest shares of the step, each timed live with CUDA-event pairs on the launching
    # stream over one eager step (serial: no stream forks, no PDL); algorithmic bytes are counted by the library per launch
    legs, traffic = [], _traffic_table()
    try:
        if not with_kernel_legs:
            raise StopIteration
        peaks, peak_kind = _peaks()
        ops.set_flag("fork", 0)
        ops.set_flag("pdl", 0)
        from transception_b200 import mstr as _mstr
        _branch_streams = _mstr.TRAIN_BRANCH_STREAMS
        _mstr.TRAIN_BRANCH_STREAMS = False
        runner._captured = False             # the measurement is over: eager steps (new gradient tensors) retire the graphs
        for kname in ("gemm_tc", "wgrad_tc", "dw_bwd_fused", "ln_bwd_fused", "dwln", "dwconv3x3"):
            ops.profile_enable(kname)
            torch.cuda.synchronize(dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(200_000_000)   # ~0.1 s spin: the host enqueues the step behind it, so brackets see back-to-back kernels
            t0.record()
            runner.eager_step()
            t1.record()
            torch.cuda.synchronize(dev)
            k_ms, k_n, k_work = ops.profile_read_work()
            ops.profile_enable("")
            if k_n and k_work > 0:
                ach = k_work / (k_ms * 1e-3) / 1e9
                tr = traffic.get(kname, {})
                legs.append({"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / peaks["hbm_gbs"], "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                             "algorithmic_bytes_per_launch": k_work / k_n, "launches_per_step": k_n, "avg_launch_ms": k_ms / k_n,
                             "ms_per_step": k_ms, "share_of_step": k_ms / t0.elapsed_time(t1),
                             "share_basis": "event-bracketed launches of one eager, serial (no forks / PDL) step of %.1f ms; event pairs "
                                            "add 2-4 us per launch, so small-kernel numbers are lower bounds" % t0.elapsed_time(t1),
                             "peak_source": peak_kind + " hbm copy"})
        legs.sort(key=lambda r: -r["ms_per_step"])
        if world == 1 and legs:
            # the denominator of the shares: the SAME serial configuration (no stream forks, no PDL, no branch streams) captured
            # as a CUDA graph and replayed — its time is the sum of the step's kernel durations (host enqueue cost removed), the
            # quantity the ncu launch list under profiles/ sums
            serial = TrainStepGraph(net, CeDiceLoss(NCLS), opt, BATCH, IN_CH, SIZE, device=dev, warmup=1, sample=(xh, lh))
            serial.replay()
            torch.cuda.synchronize(dev)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(3):
                serial.replay()
            s1.record()
            torch.cuda.synchronize(dev)
            serial_ms = s0.elapsed_time(s1) / 3
            for leg in legs:
                leg["share_of_step"] = leg["ms_per_step"] / serial_ms
                leg["share_basis"] = ("event-bracketed launches of one eager, serial (no forks / PDL / branch streams) step, over the "
                                      "replay time of that serial step captured as a CUDA graph (%.1f ms = the sum of its kernel "
                                      "durations); event pairs add 2-4 us per launch, so small-kernel numbers are upper bounds of the "
                                      "share" % serial_ms)
            del serial
    except StopIteration:
        legs = []
    except Exception as e:  # noqa: BLE001
        legs = [{"error": "%s: %s" % (type(e).__name__, e)}]
        ops.profile_enable("")
    finally:
        ops.set_flag("fork", 1)
        ops.set_flag("pdl", 1)
        try:
            _mstr.TRAIN_BRANCH_STREAMS = _branch_streams
        except NameError:
            pass
    imgs = world * BATCH * K
    step_s = ms / K * 1e-3
    peaks, peak_kind = _peaks()
    tf = 3 * FWD_GFLOP_PER_IMG * BATCH / step_s / 1e3
    tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    return {"metric": TRAIN_METRIC, "value": imgs / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / K,
            "roofline": legs[0] if legs else None, "roofline_other_kernels": legs[1:],
            "roofline_step": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                              "basis": "SURVEY 8(d): 3 x %.2f GFLOP per image (forward algorithmic FLOPs of the whole model; backward = "
                                       "dgrad + wgrad) x %d images per GPU over the measured step; %s bf16 sustained peak"
                                       % (FWD_GFLOP_PER_IMG, BATCH, peak_kind)},
            "e2e": {"value": imgs / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": xh.numel() * 4 + lh.numel() * 8, "d2h_bytes_per_step": 4,
                    "api": "TrainStepGraph.step(pinned images, pinned labels) + loss read back, every step"},
            "cuda_graph": True, "pdl_chain": pdl_chain, "library_kernels_per_step": launches, "first_loss": first_loss, "last_loss": last_loss,
            "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
            "allreduced_grads_equal_rank_mean_and_identical_across_ranks": grads_match,
            "weights_identical_across_ranks_after_training": weights_match,
            "grad_allreduce": ("backward cut between encoder stages 2 and 3: the gradients behind the cut (98 % of the elements) are "
                               "gathered into the head of one flat fp32 bucket and all-reduced (NCCL, average) asynchronously while "
                               "the backward of stages 2-1 runs; the tail follows; the fused update reads the bucket in place (timed runner: "
                               + overlap_kind + ")" if getattr(runner, "overlap", False) else
                               "gradients gathered into one flat fp32 bucket by one kernel, one NCCL all-reduce (average), the fused "
                               "update reads the bucket in place") if world > 1 else None,
            "dtype": "fp16 storage / fp16+TF32 tensor-core forward, TF32 tensor-core backward (operands read in place), fp32 "
                     "accumulation, gradients, master weights and optimizer state"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.mode == "train":
        ips, cores, sample, ms = cpu_reference_train_time(torch, args.steps, args.warmup)
        metric, workload = TRAIN_METRIC, TRAIN_WORKLOAD
    else:
        ips, cores, sample, ms = cpu_reference_time(torch, args.steps, args.warmup)
        metric, workload = METRIC, FWD_WORKLOAD
    print(json.dumps({"impl": "reference", "metric": metric, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload},
                      "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def forward_leg(torch, dev, world, rank, local, K, W, args, with_kernel_legs):
    """The inference forward (BASELINE configs[1]) at bs16: device-resident graph replay with an L2 flush between steps, end to
    end through GraphRunner.run_host, and (optionally) the per-kernel roofline legs."""
    import torch.distributed as dist
    from transception_b200 import ops
    from transception_b200.runtime import GraphRunner
    MSTransception = _model_cls()
    peaks, peak_kind = _peaks()
    torch.manual_seed(1234)
    net = MSTransception(num_classes=NCLS, image_size=SIZE).eval().to(dev)
    runner = GraphRunner(net, BATCH, IN_CH, SIZE, device=dev, microbatches=args.microbatches)
    x_host = _make_inputs(torch, BATCH, rank).pin_memory()
    y_host = torch.empty((BATCH, NCLS, SIZE, SIZE), dtype=torch.float32).pin_memory()
    runner.x.copy_(x_host)
    kernels_per_replay = runner.kernels_per_replay
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -----------------------------------------------------------
    for _ in range(W):
        runner.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    for s, e in ev:
        flush.zero_()                      # evict L2 (untimed)
        s.record(); runner.replay(); e.record()
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- end to end: pinned host input -> H2D -> forward -> D2H logits ---------------------------
    for _ in range(W):
        runner.run_host(x_host, y_host)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(K):
        runner.run_host(x_host, y_host)
    runner.drain()                         # the last step's D2H copy (copy stream) is inside the timed region
    s1.record()
    barrier()
    e2e_ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel roofline legs, timed live (eager launches, CUDA events on the launching stream) --------
    # work = algorithmic bytes (HBM-bound kernels) or FLOPs (flash) summed by the library over the timed launches
    KERNELS = (("gemm_tc", "hbm"), ("dwln", "hbm"), ("mb_fused16", "hbm"), ("flash_tc", "tensor"), ("flash_ffma", "tensor"))
    if not with_kernel_legs:
        KERNELS = ()
    legs = []
    reps = max(3, min(K, 5))
    ops.set_flag("fork", 0)      # serial kernels for attribution: concurrent branches would share SMs inside a bracket
    ops.set_flag("pdl", 0)
    for kname, bound in KERNELS:
        ops.profile_enable(kname)
        fw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        with torch.no_grad():
            for s0, s1 in fw:
                flush.zero_()
                torch.cuda._sleep(80_000_000)      # ~40 ms spin: the host enqueues the whole forward behind it, so the
                s0.record()                        # event pairs bracket back-to-back kernels, not launch latency
                net(runner.x)
                s1.record()
        torch.cuda.synchronize(dev)
        k_ms, k_n, k_work = ops.profile_read_work()
        ops.profile_enable("")
        if k_n:
            legs.append({"kernel": kname, "bound": bound, "ms": k_ms, "n": k_n, "work": k_work, "per_step_ms": k_ms / reps,
                         "launches_per_step": k_n // reps, "eager_step_ms": sum(a.elapsed_time(b) for a, b in fw) / reps})
    ops.set_flag("fork", 1)
    ops.set_flag("pdl", 1)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    imgs = world * BATCH * K
    out = {"metric": METRIC, "value": imgs / (dev_ms * 1e-3), "unit": "images/s", "ms_per_step": dev_ms / K, "clocks": clocks,
           "e2e": {"value": imgs / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms / K,
                   "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4,
                   "api": "GraphRunner.run_host(pinned x, pinned logits): every step copies its input H2D, replays the forward and "
                          "copies its logits D2H; the D2H runs on a copy stream and overlaps the next step"},
           "gpu_launches": kernels_per_replay * K, "kernels_per_forward": kernels_per_replay,
           "dtype": "f16/tf32 tensor-core MMA, f32 accumulate, f32 IO"}
    step_s = dev_ms / K * 1e-3
    tf = FWD_GFLOP_PER_IMG * BATCH / step_s / 1e3
    gbs = (FWD_MB_PER_IMG_BF16 * BATCH + WEIGHT_MB_BF16) / 1e3 / step_s
    tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    out["roofline_step"] = {"tensor": {"achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak},
                            "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]},
                            "basis": "SURVEY 8(d): %.2f GFLOP and %.1f MB (bf16 fused-kernel boundaries) per image + %.1f MB weights per batch"
                                     % (FWD_GFLOP_PER_IMG, FWD_MB_PER_IMG_BF16, WEIGHT_MB_BF16)}
    traffic = _traffic_table()
    rl = []
    for g in legs:
        if g["bound"] == "hbm":
            ach, peak, unit, src = g["work"] / (g["ms"] * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s", peak_kind + " hbm copy"
        else:
            peak = tpeak
            ach, unit, src = g["work"] / (g["ms"] * 1e-3) / 1e12, "TFLOP/s", peak_kind + " bf16 sustained"
        tr = traffic.get(g["kernel"], {})
        rl.append({"kernel": g["kernel"], "bound": g["bound"], "achieved": ach, "peak": peak, "unit": unit,
                   "frac": ach / peak, "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                   "launches_per_step": g["launches_per_step"],
                   "avg_launch_ms": g["ms"] / g["n"], "ms_per_step": g["per_step_ms"],
                   "share_of_step": g["per_step_ms"] / g["eager_step_ms"],
                   "share_basis": "eager, serial (no fork / PDL) forward of %.2f ms bracketed in the same run; the "
                                  "event pair around every launch adds ~2-4 us, so small-kernel times are upper bounds"
                                  % g["eager_step_ms"],
                   "peak_source": src})
    rl.sort(key=lambda r: -r["ms_per_step"])
    out["kernel_legs"] = rl
    del runner, net
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from transception_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.load_library()
    for fv in args.flag:
        name, _, val = fv.partition("=")
        ops.set_flag(name, int(val))
    K, W = args.steps, max(args.warmup, 3)
    headline = MODEL == "MSTransception" and (SIZE, NCLS, IN_CH) == (224, 9, 1)

    if args.mode == "train":
        sampler = ClockSampler(local) if rank == 0 else None
        tr = train_step_leg(torch, dev, world, rank, K, W, with_kernel_legs=not args.no_legs)
        clocks = sampler.stop() if sampler else None
        fwd = None
        if not args.no_forward:
            try:
                fwd = forward_leg(torch, dev, world, rank, local, max(5, min(K, 20)), 3, args, with_kernel_legs=False)
                fwd.pop("clocks", None)
            except Exception as e:  # noqa: BLE001
                fwd = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            line = {"metric": tr["metric"] if headline else tr["metric"] + " [non-headline workload: %s %dx%d %d classes]" % (MODEL, SIZE, SIZE, NCLS),
                    "value": tr["value"], "unit": "images/s", "n_gpus": world, "steps": K,
                    "warmup": W, "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": tr["dtype"], "data": "synthetic",
                    "config": {"workload": TRAIN_WORKLOAD if headline else "TransCeption %s %dx%d bs16 train step, %d classes" % (MODEL, SIZE, SIZE, NCLS),
                               "l2": "step footprint (activations saved for backward, > 2 GB) exceeds L2",
                               "timing": "CUDA events around K steps; max over ranks",
                               "graph": ("forward+loss+backward graph and fused-SGD graph" if world == 1 else
                                         "forward+loss+backward-to-the-cut graph, bucket-head gather graph, async NCCL all-reduce beside the "
                                         "stage 2-1 backward graph, bucket-tail gather graph + all-reduce, fused-SGD graph")},
                    "clocks": clocks, "e2e": tr["e2e"], "gpu_launches": tr["library_kernels_per_step"] * K,
                    **({"flags": list(args.flag)} if args.flag else {}),
                    "roofline": tr["roofline"], "roofline_other_kernels": tr["roofline_other_kernels"], "roofline_step": tr["roofline_step"],
                    "train": {k: tr[k] for k in ("cuda_graph", "pdl_chain", "library_kernels_per_step", "first_loss", "last_loss", "grad_allreduce",
                                                   "peak_memory_gb", "allreduced_grads_equal_rank_mean_and_identical_across_ranks",
                                                   "weights_identical_across_ranks_after_training")},
                    "forward_step": fwd}
            if world == 1 and not args.no_gpu_baseline:
                try:
                    line["gpu_baseline"] = gpu_baseline_leg(torch, dev, "train")
                except Exception as e:  # noqa: BLE001
                    line["gpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            if not args.no_cpu and world == 1:
                ips, cores, sample, _ = cpu_reference_train_time(torch, 1, 1, budget_s=40.0)
                line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
                if headline:
                    line["cpu_baseline"]["parity_bs16"] = cpu_train_parity(torch, tr["first_loss"])
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    fwd = forward_leg(torch, dev, world, rank, local, K, W, args, with_kernel_legs=True)
    if rank == 0:
        line = {"metric": METRIC, "value": fwd["value"], "unit": "images/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": fwd["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": fwd["dtype"], "data": "synthetic",
                "config": {"workload": FWD_WORKLOAD if headline else "TransCeption %s %dx%d bs16 fp32 forward, %d class logits [non-headline]" % (MODEL, SIZE, SIZE, NCLS),
                           "l2": "256 MiB memset between timed steps (untimed); step footprint > L2",
                           "timing": "per-step CUDA events on the replay stream, summed; max over ranks",
                           "graph": "whole forward replayed as one CUDA graph"},
                "clocks": fwd["clocks"], "e2e": fwd["e2e"], "gpu_launches": fwd["gpu_launches"], "roofline_step": fwd["roofline_step"]}
        rl = fwd["kernel_legs"]
        if rl:
            line["roofline"] = rl[0]                 # the kernel with the largest share of the step
            line["roofline_other_kernels"] = rl[1:]
        if world == 1 and not args.no_gpu_baseline:
            try:
                line["gpu_baseline"] = gpu_baseline_leg(torch, dev, "forward")
            except Exception as e:  # noqa: BLE001
                line["gpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if not args.no_cpu and world == 1:
            ips, cores, sample, _ = cpu_reference_time(torch, 3, 1, budget_s=40.0)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    global SIZE, NCLS, IN_CH, METRIC, MODEL
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mode", default="train", choices=["forward", "train"],
                    help="train = BASELINE.json's metric (fwd+bwd bs16 train step, the headline); forward = configs[1]")
    ap.add_argument("--no-forward", action="store_true", help="train mode: skip the forward_step leg")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the gpu_baseline leg (oracle on the same GPU)")
    ap.add_argument("--no-legs", action="store_true", help="train mode: skip the per-kernel roofline legs (quick A/B runs)")
    ap.add_argument("--flag", action="append", default=[], metavar="NAME=VALUE",
                    help="library back-end switch for A/B runs (tcx_set_flag), e.g. --flag pdl=0; recorded in the line as `flags`")
    ap.add_argument("--size", type=int, default=SIZE, help="input side (default 224 = the headline workload; 256 = config 5)")
    ap.add_argument("--classes", type=int, default=NCLS)
    ap.add_argument("--in-ch", type=int, default=IN_CH)
    ap.add_argument("--microbatches", type=int, default=1,
                    help="split each rank's batch of 16 into this many slices captured on parallel streams")
    ap.add_argument("--model", default=MODEL, choices=["MSTransception", "Transception"],
                    help="MSTransception = the headline workload; Transception = the networks/Transception.py variant")
    args = ap.parse_args()
    if (args.size, args.classes, args.in_ch, args.model) != (SIZE, NCLS, IN_CH, MODEL):
        SIZE, NCLS, IN_CH, MODEL = args.size, args.classes, args.in_ch, args.model
        METRIC = "images/sec fwd @%dx%d bs16 (TransCeption %s, %d classes, fp32 IO) [non-headline workload]" % (
            SIZE, SIZE, MODEL, NCLS)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
