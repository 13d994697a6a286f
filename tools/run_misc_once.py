"""One pass over the smaller kernel families at 2^20 elements -- the command profiled under ncu for the kernels that the
Groth16 bench does not launch: batched VM opcodes, masks / F::rand / share_rep3, lincomb (Shamir), poly eval."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np

from co_snarks_b200 import binding as B

n = 1 << 20
ctx = B.Context(0)
lib = ctx.lib
rng = np.random.Generator(np.random.PCG64(5))


def rnd(m):
    a = rng.integers(0, 2 ** 63, size=(m, 4), dtype=np.uint64) << np.uint64(1)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


dx, dy, dp, do = ctx.to_device(rnd(2 * n)), ctx.to_device(rnd(2 * n)), ctx.to_device(rnd(n)), ctx.alloc(n * 64)
for op, y in ((B.R3B_ADD, dy), (B.R3B_MUL_PUBLIC, dp), (B.R3B_ADD_PUBLIC, dp)):
    ctx.rep3_batch(B.CS_BN254, op, 0, dx, y, do, n)
seed = bytes(range(32))
ctx.rep3_masks_device(B.CS_BN254, seed, 0, seed[::-1], 0, n, do)
ctx._check(lib.cs_fr_rand_device(ctx.h, B.CS_BN254, seed, C.c_uint64(0), C.c_void_p(do), n))
w = np.ones(4, dtype=np.uint64)
ins = (C.c_void_p * 3)(C.c_void_p(dx), C.c_void_p(dy), C.c_void_p(dp))
wts = rnd(3)
ctx._check(lib.cs_vec_lincomb(ctx.h, B.CS_BN254, ins, B._ptr(wts), 3, n, C.c_void_p(do)))
ctx.synchronize()
print("ok")
