"""One process, two GPUs: cs_rep3_mul_vec_reshare on GPU 0 storing the reshared halves into a buffer on GPU 1 over
NVLink -- the command profiled under ncu for the fused product + peer-store kernel (single process, so ncu can follow
it).  usage: python tools/run_mul_vec_peer_once.py [log_n] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from co_snarks_b200 import binding as B

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
ctx0, ctx1 = B.Context(0), B.Context(1)
# in-process peer nets: connecting them enables peer access between the two devices
nets = [B.Net.peer(ctx0, 0, 2), B.Net.peer(ctx1, 1, 2)]
for x in nets:
    x.connect_local(nets)
g = np.random.Generator(np.random.PCG64(9))
sh = g.integers(0, 2 ** 63, size=(2, 2 * n, 4), dtype=np.uint64)
sh[..., 3] &= np.uint64((1 << 60) - 1)
d_a, d_b = ctx0.to_device(sh[0]), ctx0.to_device(sh[1])
d_out = ctx0.alloc(n * 64)
d_peer = ctx1.alloc(n * 64)   # the "next party's" vector, in GPU 1's HBM
seed = bytes(range(32))
prf = (seed, 0, seed[::-1], 0, 12)
torch.cuda.set_device(0)
for peer in (d_peer, None):
    ms = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ctx0.rep3_mul_vec_reshare(B.CS_BN254, d_a, d_b, n, prf, d_out, peer)
        ctx0.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print("peer_store" if peer else "local_only", "2^%d" % lg, "ms", [round(x, 3) for x in ms],
          "payload_GBs", round(n * 32 / (min(ms) * 1e-3) / 1e9, 1) if peer else None)
for x in nets:
    x.free()
