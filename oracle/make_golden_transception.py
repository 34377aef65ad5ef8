#!/usr/bin/env python
"""Generate tests/golden/transception_golden.pt from the REAL reference variant (authoring container only) — TEST
INFRASTRUCTURE.

    python oracle/make_golden_transception.py

Imports /root/reference/networks/Transception.py through oracle/ref_shim.py, proves same-seed initialisation and
state_dict-key parity of the drop-in mirror (561 tensors), loads the perturbed weights STRICTLY into the reference model,
runs the reference's own modules on seeded inputs and stores (subsampled) outputs + summary statistics.  The fixtures pin
oracle/transception_oracle.py (tests/test_oracle_transception.py, CPU) and the CUDA path (tests/test_gpu_transception.py).
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import fixtures as FX  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.cases import flatten_out  # noqa: E402
from oracle.cases_transception import CASES, MODEL_SEED, seeded_model  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref_shim.load_reference()
    R = importlib.import_module("_tcx_ref_networks.Transception")
    sd = seeded_model(perturb=True).state_dict()
    plain = seeded_model(perturb=False).state_dict()
    torch.manual_seed(MODEL_SEED)
    ref = R.Transception(num_classes=9).eval()
    ref_sd = ref.state_dict()
    assert list(plain.keys()) == list(ref_sd.keys()), "state_dict key order differs from the reference"
    bad = [k for k in ref_sd if not torch.equal(plain[k], ref_sd[k])]
    assert not bad, "same-seed init differs for %d tensors, e.g. %s" % (len(bad), bad[:3])
    ref.load_state_dict(sd, strict=True)
    golden = {"meta": {"torch": torch.__version__, "n_keys": len(ref_sd), "model_seed": MODEL_SEED,
                       "perturb_seed": FX.PERTURB_SEED, "reference_commit": "0c7ee13"}}
    with torch.no_grad():
        for name, path, mk, _ in CASES:
            out = flatten_out(ref.get_submodule(path)(*mk()))
            golden[name] = {"path": path, "sub": [FX.subsample(o) for o in out], "stats": [FX.stats(o) for o in out],
                            "shape": [tuple(o.shape) for o in out]}
            print("%-20s %-32s %s" % (name, path, [tuple(o.shape) for o in out]))
        for cin, bs in ((1, 2), (3, 1)):
            x = FX.image(bs, cin, seed=0)
            enc = ref.backbone(x.repeat(1, 3, 1, 1) if cin == 1 else x)
            logits = ref(x)
            golden["model_c%d" % cin] = {
                "logits_sub": logits[:, :, ::4, ::4].clone(), "logits_stats": FX.stats(logits),
                "argmax_sub": logits.argmax(1)[:, ::2, ::2].to(torch.uint8).clone(),
                "enc_sub": [FX.subsample(m) for m in enc], "enc_stats": [FX.stats(m) for m in enc]}
            print("model_c%d logits stats %s" % (cin, FX.stats(logits).tolist()))
        # selective-kernel fusion (concat='sk'): same parameters, other stage tail
        ref_sk = R.Transception(num_classes=9, concat='sk').eval()
        ref_sk.load_state_dict(sd, strict=True)
        x = FX.image(1, 1, seed=0)
        enc = ref_sk.backbone(x.repeat(1, 3, 1, 1))
        logits = ref_sk(x)
        golden["model_sk"] = {"logits_sub": logits[:, :, ::4, ::4].clone(), "logits_stats": FX.stats(logits),
                              "argmax_sub": logits.argmax(1)[:, ::2, ::2].to(torch.uint8).clone(),
                              "enc_sub": [FX.subsample(m) for m in enc], "enc_stats": [FX.stats(m) for m in enc]}
        print("model_sk logits stats %s" % FX.stats(logits).tolist())
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "transception_golden.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
