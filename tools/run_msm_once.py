"""Runs a few standalone 2^k G1 (and optionally G2) MSMs -- the command profiled under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from co_snarks_b200 import binding as B

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
group = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
skew = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0  # fraction of scalars set to 1 (witness-like skew)
n = 1 << lg
ctx = B.Context(0)
rng = np.random.Generator(np.random.PCG64(4))
def rnd(n):
    a = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64) << np.uint64(1)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    return a
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
if group == 0:
    gen = B.ints_to_limbs(B.to_mont_ints([1, 2], Q, 4), 4).reshape(-1)
else:
    g2 = [10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930, 4082367875863433681332203403145435568316851327593401208105741076214120093531]
    gen = B.ints_to_limbs(B.to_mont_ints(g2, Q, 4), 4).reshape(-1)
pts = ctx.fixed_base_mul(B.CS_BN254, group, gen, rnd(n), montgomery=False)
import time
ctx.synchronize()
_t0 = time.perf_counter()
bases = ctx.bases_upload(B.CS_BN254, group, pts, wb)
ctx.synchronize()
print("bases_upload_ms", round((time.perf_counter() - _t0) * 1e3, 1), "(H2D of n points + per-window table expansion)")
sc = rnd(n)
if skew > 0:
    ones = rng.random(n) < skew
    sc[ones] = 0
    sc[ones, 0] = 1
d = ctx.to_device(sc)
ctx.msm_profile(True)
for _ in range(reps):
    out, inf = ctx.msm(bases, d, n=n, montgomery=False, device=True)
    print("stage_ms", [round(x, 3) for x in ctx.msm_stage_ms()])
