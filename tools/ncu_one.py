#!/usr/bin/env python
"""Run one library op a few times (for `ncu --set full -k regex:<kernel>` captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transception_b200 import ops  # noqa: E402

ops.load_library()
what = sys.argv[1]
r = lambda *s: torch.randn(*s, device="cuda")
if what == "linear":
    M, N, K = (int(a) for a in sys.argv[2:5])
    res = len(sys.argv) > 5
    x, w, b, rr = r(M, K), r(N, K), r(N), r(M, N)
    fn = lambda: ops.linear(x, w, b, residual=rr if res else None)
elif what == "linear16":
    M, N, K, o16 = (int(a) for a in sys.argv[2:6])
    x, w, b, rr = r(M, K).half(), r(N, K).half(), r(N), r(M, N)
    fn = lambda: ops.linear_f16(x, w, b, residual=None if o16 else rr, out_f16=bool(o16))
elif what == "flash":
    q, kv = r(16, 6076, 64), r(16, 784, 128)
    fn = lambda: ops.flash_attn(q, kv, 0.125)
elif what == "flash16":
    q, kv = r(16, 6076, 64).half(), r(16, 784, 128).half()
    fn = lambda: ops.flash_attn_f16(q, kv, 0.125)
elif what == "mixffn":
    B, hw, C = (int(a) for a in sys.argv[2:5])
    x = r(B, hw * hw, C)
    C4 = 4 * C
    args = (r(C4, C) * C ** -0.5, r(C4), r(C4, 1, 3, 3), r(C4), r(C4), r(C4), 1e-5, r(C, C4) * C4 ** -0.5, r(C))
    fn = lambda: ops.mixffn_skip(x, hw, hw, *args, residual=x)
for _ in range(4):
    fn()
torch.cuda.synchronize()
