#!/usr/bin/env python
"""Generate tests/golden/*.pt from the REAL reference (authoring container only) — TEST INFRASTRUCTURE.

    python oracle/make_golden.py

Imports /root/reference/networks/MSTr.py through oracle/ref_shim.py (two shims), loads the seeded +
perturbed weights of oracle/fixtures.py STRICTLY into the reference model (which also proves state_dict
compatibility of the drop-in mirror: 2 200 keys), runs the reference's own modules on seeded inputs and stores
their outputs (subsampled to keep the fixtures small) together with summary statistics.  The fixtures pin
oracle/mstr_oracle.py (tests/test_oracle.py, CPU) and the CUDA path (tests/test_gpu_parity.py::test_golden_*).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import fixtures as FX  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

from oracle.cases import CASES, flatten_out as _flatten_out  # noqa: E402


def main():
    R = ref_shim.load_reference()
    mirror = FX.seeded_model(perturb=True)
    sd = mirror.state_dict()
    torch.manual_seed(FX.MODEL_SEED)
    ref = R.MSTransception(num_classes=9).eval()
    # same-seed init parity of the drop-in mirror (before perturbation) — every one of the 2 200 tensors
    plain = FX.seeded_model(perturb=False).state_dict()
    ref_sd = ref.state_dict()
    assert list(plain.keys()) == list(ref_sd.keys()), "state_dict key order differs from the reference"
    bad = [k for k in ref_sd if not torch.equal(plain[k], ref_sd[k])]
    assert not bad, "same-seed init differs for %d tensors, e.g. %s" % (len(bad), bad[:3])
    ref.load_state_dict(sd, strict=True)
    os.makedirs(OUT, exist_ok=True)
    golden = {"meta": {"torch": torch.__version__, "n_keys": len(ref_sd), "model_seed": FX.MODEL_SEED,
                       "perturb_seed": FX.PERTURB_SEED, "reference_commit": "0c7ee13"}}
    with torch.no_grad():
        for name, path, mk, _ in CASES:
            args = mk()
            out = _flatten_out(ref.get_submodule(path)(*args))
            golden[name] = {"path": path, "sub": [FX.subsample(o) for o in out], "stats": [FX.stats(o) for o in out],
                            "shape": [tuple(o.shape) for o in out]}
            print("%-22s %-70s %s" % (name, path, [tuple(o.shape) for o in out]))
        # whole model, both input-channel forms (configs[0]: bs2 single-slice forward)
        for cin, bs in ((1, 2), (3, 1)):
            x = FX.image(bs, cin, seed=0)
            xin = x.repeat(1, 3, 1, 1) if cin == 1 else x
            enc = ref.backbone(xin)
            br = ref.bridge(enc)
            logits = ref(x)
            golden["model_c%d" % cin] = {
                "logits_sub": logits[:, :, ::4, ::4].clone(), "logits_stats": FX.stats(logits),
                "argmax_sub": logits.argmax(1)[:, ::2, ::2].to(torch.uint8).clone(),
                "enc_sub": [FX.subsample(m) for m in enc], "enc_stats": [FX.stats(m) for m in enc],
                "bridge_sub": [FX.subsample(m) for m in br], "bridge_stats": [FX.stats(m) for m in br]}
            print("model_c%d logits stats %s" % (cin, FX.stats(logits).tolist()))
        # the survey's un-perturbed fingerprints re-derived from the live reference
        ref.load_state_dict(plain, strict=True)
        x = FX.image(2, 1, seed=0)
        lg = ref(x)
        golden["fingerprint"] = {"x_sum": x.sum().item(), "param_sum": sum(p.double().sum().item() for p in ref.parameters()),
                                 "logits_stats": FX.stats(lg), "logits_00": lg[0, :, 0, 0].clone(),
                                 "logits_1_111_57": lg[1, :, 111, 57].clone()}
        print("fingerprint", golden["fingerprint"]["x_sum"], golden["fingerprint"]["param_sum"], lg[0, :, 0, 0].tolist())
    path = os.path.join(OUT, "mstr_golden.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
