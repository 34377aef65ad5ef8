"""CPU tests of the host side: C-ABI surface, drop-in mirror, loud failure without CUDA, world-size-2 sharding."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "transception_sm100.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tcx_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from transception_b200 import ops
    assert os.path.exists(ops.LIB_PATH), "build the library first: python -m transception_b200.build"
    lib = ctypes.CDLL(ops.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes binding declares a prototype for every export, and nothing that is not in the header
    assert sorted(ops.EXPORTS) == names
    lib.tcx_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.tcx_version()


def test_mirror_state_dict_surface():
    from networks.MSTr import MSTransception
    torch.manual_seed(1234)
    net = MSTransception(num_classes=9)
    sd = net.state_dict()
    assert len(sd) == 2200
    assert sum(p.numel() for p in net.parameters()) == 47316553
    # aliases created by the shared cpe/crpe modules
    a = "backbone.mhca_stage2.mhca_blks.0.cpe.proj.weight"
    b = "backbone.mhca_stage2.mhca_blks.0.MHCA_layers.1.cpe.proj.weight"
    assert sd[a].data_ptr() == sd[b].data_ptr()
    for dead in ("backbone.conv1_1_s1.weight", "backbone.cpe.proj.weight",
                 "bridge.bridge_layer1.attn.scale_reduce.sr0.weight", "backbone.block1.0.mlp.norm2.weight"):
        assert dead in sd


def test_constructor_contract():
    from networks.MSTr import MSTransception
    for kw in (dict(Stage_3or4=4), dict(have_bridge="sp"), dict(have_bridge="para"), dict(concat="cbam")):
        with pytest.raises(NotImplementedError):
            MSTransception(num_classes=9, **kw)
    net = MSTransception(num_classes=2, br_ch_att_list=[False, True, False, False])
    assert type(net.bridge.bridge_layer2.attn).__name__ == "M_EfficientChannelAtten"
    assert net.decoder_0.last_layer.weight.shape[0] == 2


def _kernel_body(src, name):
    """Text of the body of ``__global__ ... name(...) { ... }`` (every definition concatenated)."""
    out = []
    for m in re.finditer(r"__global__[^;{}]*?\b%s\s*\(" % re.escape(name), src):
        i, depth = m.end() - 1, 0
        while True:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            if depth == 0:
                break
            i += 1
        b = src.index("{", i)
        assert src[i + 1:b].strip() == "", name
        j, depth = b, 0
        while True:
            depth += {"{": 1, "}": -1}.get(src[j], 0)
            if depth == 0:
                break
            j += 1
        out.append(src[b + 1:j])
    return out


def test_every_kernel_launched_with_the_pdl_attribute_waits_for_its_predecessor():
    """A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start before the kernel in front of it
    has finished; it must execute griddepcontrol.wait before touching anything that kernel wrote.  Static check over the
    sources: kernels launched through tcx_launch_chain start with PDL_TOP() (trigger + wait as the first statement), kernels
    launched through tcx_launch_pdl contain a pdl_wait()."""
    csrc = os.path.join(ROOT, "transception_b200", "csrc")
    texts = {f: open(os.path.join(csrc, f)).read() for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh"))}
    everything = "\n".join(texts.values())
    chain = set(re.findall(r"tcx_launch_chain\(\s*([A-Za-z_][A-Za-z0-9_]*)", everything))
    pdl = set(re.findall(r"tcx_launch_pdl\(\s*([A-Za-z_][A-Za-z0-9_]*)", everything))
    for helper_param in ("kernel", "void"):     # the helpers' own declarations: tcx_launch_pdl(void (*kernel)(KArgs...), ...)
        chain.discard(helper_param)
        pdl.discard(helper_param)
    assert len(chain) >= 60 and len(pdl) >= 15, (len(chain), len(pdl))
    for name in sorted(chain):
        bodies = _kernel_body(everything, name)
        assert bodies, "no definition found for %s" % name
        for body in bodies:
            assert body.lstrip().startswith("PDL_TOP();"), "%s is launched with the PDL attribute but does not start with PDL_TOP()" % name
    for name in sorted(pdl):
        bodies = _kernel_body(everything, name)
        assert bodies, "no definition found for %s" % name
        for body in bodies:
            assert "pdl_wait()" in body or "PDL_TOP()" in body, "%s is launched with the PDL attribute but never waits" % name
    # and nothing is launched the plain way from a file whose kernels were converted (a plain launch after a converted kernel
    # would be correct, but the list above is meant to be complete)
    for f in ("bwd.cu", "bwd_enc.cu", "bwd_mb.cu", "bwd_mix.cu", "flash_bwd.cu", "loss.cu", "optim.cu", "misc.cu", "elementwise.cu"):
        assert "<<<" not in texts[f], f


def test_no_cpu_fallback():
    from networks.MSTr import MSTransception
    net = MSTransception(num_classes=9).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 224, 224))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "transception_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, "%s mentions the oracle" % f


def test_shard_range():
    from transception_b200 import shard
    assert shard.shard_range(128, 3, 8) == (48, 64)
    assert [shard.shard_range(32, r, 2) for r in range(2)] == [(0, 16), (16, 32)]
    with pytest.raises(ValueError):
        shard.shard_range(30, 0, 4)
    with pytest.raises(ValueError):
        shard.shard_range(32, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transception_b200 import shard
    g = torch.Generator().manual_seed(0)
    x = torch.rand(8, 1, 4, 4, generator=g)                # same global batch on every rank
    mine = shard.shard_batch(x, rank, world)
    y = mine * 2 + 1                                       # stand-in for the per-image forward (no collective)
    full = shard.gather_rows(y, world)
    ms = shard.max_over_ranks([10.0 + rank, 5.0 - rank])
    q.put((rank, torch.equal(full, x * 2 + 1), ms))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ms in res:
        assert ok, "rank %d: gathered shards differ from the unsharded result" % rank
        assert ms == [11.0, 5.0]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transception_b200.shard import GradBucket
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 4), (5,), (2, 2, 2), (7,))]
    for i, p in enumerate(params[:3]):                         # the last parameter is "dead": no gradient on any rank
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = GradBucket(params).allreduce()
    want = [(1 + world) / 2.0 * (i + 1) for i in range(3)]     # mean over ranks of (rank+1)*(i+1)
    ok = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(params[:3], want)) and params[3].grad is None
    q.put((rank, ok, n))
    dist.destroy_process_group()


def test_world_size_2_gloo_gradient_allreduce():
    """one all-reduce (average) over the flat bucket of used-parameter gradients; dead parameters keep grad None."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, n in res:
        assert ok, "rank %d: averaged gradients wrong" % rank
        assert n == 12 + 5 + 8


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs beside ours): under torchrun only rank 0 works and prints; the other
    ranks exit 0 silently.  Rank 0's line carries the keys of the contract with `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-300:], r.stderr[-300:])
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--mode", "forward", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
