#!/usr/bin/env python
"""Single-GPU check of `cfg_ddim_update_p2p` (csrc/elementwise.cu): the two ranks of a CFG-branch pair are played by two
streams of one device whose receive slots / flag words point at each other.  Each step's result must be bit-identical to
`cfg_ddim_update` on the concatenated `[e_u; e_c]` batch.  Run as its own process (tests/test_cfg_p2p_gpu.py does): the
kernel's bounded wait traps if the partner never arrives, which would poison the caller's CUDA context.

    python tools/cfg_p2p_loopback.py [--steps 5]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seervideoldm_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

dev = torch.device("cuda", 0)
ok = True
for (b, C, F2, cond_f, H) in ((1, 4, 3, 0, 8), (2, 4, 11, 1, 32), (3, 4, 10, 2, 32)):
    g = torch.Generator().manual_seed(b * 100 + F2)
    n = b * C * F2 * H * H
    npad = (n + 31) // 32 * 32
    recv = [torch.zeros(2 * npad, device=dev) for _ in range(2)]            # [rank][slot * npad ...]
    flag = [torch.zeros(32, dtype=torch.int32, device=dev) for _ in range(2)]
    cnt = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    x = torch.randn(b, C, F2, H, H, generator=g).to(dev)
    worst = 0
    for step in range(1, args.steps + 1):
        eps = torch.randn(2 * b, C, F2 + cond_f, H, H, generator=g).to(dev)
        coef = [float(v) for v in torch.rand(4, generator=g) * 0.9 + 0.05]
        want_prev, want_x0 = ops.cfg_ddim_update(eps, x, cond_f, True, 7.5, *coef)
        halves = [eps[:b].contiguous(), eps[b:].contiguous()]
        torch.cuda.synchronize()
        s = step & 1
        got = [None, None]
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                got[r] = ops.cfg_ddim_update_p2p(halves[r], r, recv[1 - r][s * npad:(s + 1) * npad], recv[r][s * npad:(s + 1) * npad],
                                                 flag[1 - r][16 * s:16 * s + 1], flag[r][16 * s:16 * s + 1], cnt[r], step, x,
                                                 cond_f, 7.5, *coef)
        torch.cuda.synchronize()
        for r in (0, 1):
            same = bool(torch.equal(got[r][0], want_prev)) and bool(torch.equal(got[r][1], want_x0))
            ok = ok and same
            worst += 0 if same else 1
        x = want_prev
    print(f"b={b} C={C} F2={F2} cond_f={cond_f} {H}x{H}: {args.steps} steps, both branches bit-identical to cfg_ddim_update: {worst == 0}")
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
