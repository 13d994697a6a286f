"""UltraHonk sumcheck kernels at size: one arithmetic-relation round (plain and Rep3) and one partially_evaluate over
40 polynomials, CUDA-event timed, with the algorithmic bytes / products they imply.
usage: python tools/time_sumcheck.py [log_n ...]   (also the command profiled under ncu: `... 20 once`)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from co_snarks_b200 import binding as B

args = [a for a in sys.argv[1:] if a != "once"]
once = "once" in sys.argv
sizes = [int(a) for a in args] or [20]
ctx = B.Context(0)
rng = np.random.Generator(np.random.PCG64(3))


def rnd(n):  # n random Montgomery-form elements (any 254-bit pattern below r is a valid element)
    a = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64) << np.uint64(1)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for lg in sizes:
    n = 1 << lg
    reps = 1 if once else 10
    d_beta = ctx.alloc(n * 32)
    ctx.sumcheck_gate_separator(B.CS_BN254, rnd(lg), d_beta)
    wit = ("w_l", "w_r", "w_o", "w_4", "w_l_shift", "w_4_shift")
    sel = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith")
    plain = {nm: ctx.to_device(rnd(n)) for nm in wit + sel}
    shared = dict({nm: ctx.to_device(rnd(2 * n)) for nm in wit}, **{nm: plain[nm] for nm in sel})
    res = {}
    res["arith_round_plain_ms"] = timed(lambda: ctx.sumcheck_arith_round(B.CS_BN254, B.CS_PLAIN, 0, plain, n, d_beta, 2), reps)
    res["arith_round_rep3_ms"] = timed(lambda: ctx.sumcheck_arith_round(B.CS_BN254, B.CS_REP3, 0, shared, n, d_beta, 2), reps)
    # 13 polynomials x 2 rows x 32 B (+ the scaling factor) per edge; Rep3: 6 of them are 64-B shares
    res["arith_round_plain_gbs"] = (13 * 2 * 32 + 32) * (n / 2) / (res["arith_round_plain_ms"] * 1e-3) / 1e9
    res["arith_round_rep3_gbs"] = ((7 * 32 + 6 * 64) * 2 + 32) * (n / 2) / (res["arith_round_rep3_ms"] * 1e-3) / 1e9
    # fold: 40 public polynomials (the AllEntities count of UltraHonk), 32 B in per row, 16 B out per row
    k = 40
    big_in = [ctx.to_device(rnd(n)) for _ in range(k)]
    big_out = [ctx.alloc(n // 2 * 32) for _ in range(k)]
    u = rnd(1)[0]

    def fold():
        ctx.sumcheck_fold(B.CS_BN254, big_in, big_out, False, n, u)
        ctx.synchronize()
    res["fold_40_polys_ms"] = timed(fold, reps)
    res["fold_gbs"] = k * n * 48 / (res["fold_40_polys_ms"] * 1e-3) / 1e9
    out["2p%d" % lg] = {a: round(b, 4) for a, b in res.items()}
    for p in list(plain.values()) + [shared[nm] for nm in wit] + big_in + big_out + [d_beta]:
        ctx.free(p)
print(json.dumps(out))
