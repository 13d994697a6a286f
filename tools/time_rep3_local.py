"""Times one party's Rep3 local phase (cs_groth16_rep3_local) on one GPU: pageable vs pinned host buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from co_snarks_b200 import binding as B
from co_snarks_b200.rep3 import random_field_limbs
from workloads.synth_groth16 import SynthGroth16

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = B.Context(0)
syn = SynthGroth16(ctx, lg, valid=False)
pk = syn.make_key()
n = 1 << lg
rng = np.random.Generator(np.random.PCG64(1))
fes = lambda k: random_field_limbs(rng, k)
nw = syn.m - syn.ni
shares = np.concatenate([fes(nw), fes(nw)], axis=1)
m1, m2 = fes(n), fes(n)
r_sh, s_sh = fes(2), fes(2)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.int64).pin_memory()
    t.numpy().view(np.uint64)[:] = a
    return t.numpy().view(np.uint64), t
for label, (sh, a1, a2, keep) in (("pageable", (shares, m1, m2, None)),
                                  ("pinned", (pin(shares)[0], pin(m1)[0], pin(m2)[0], None))):
    # keep pinned tensors alive
    if label == "pinned":
        ps, pm1, pm2 = pin(shares), pin(m1), pin(m2)
        sh, a1, a2 = ps[0], pm1[0], pm2[0]
    for _ in range(3):
        pk.rep3_local(0, syn.public_inputs, sh, a1, a2, r_sh, s_sh)
    t0 = time.perf_counter()
    for _ in range(5):
        pk.rep3_local(0, syn.public_inputs, sh, a1, a2, r_sh, s_sh)
    print(label, "rep3_local ms:", (time.perf_counter() - t0) / 5 * 1e3)
    t0 = time.perf_counter()
    for _ in range(5):
        pk.rep3_local(0, syn.public_inputs, sh, None, None, r_sh, s_sh)
    print(label, "rep3_local (no masks) ms:", (time.perf_counter() - t0) / 5 * 1e3)
t0 = time.perf_counter()
for _ in range(5):
    pk.prove_plain(syn.public_inputs, syn.private_witness, r_sh[0:1], s_sh[0:1])
print("plain prove ms:", (time.perf_counter() - t0) / 5 * 1e3)
# host-side protocol arithmetic: ~11 G1 scalar muls + adds through the C ABI
g = syn.points["delta_g1"][0]
t0 = time.perf_counter()
for _ in range(11):
    B.point_scalar_mul(ctx.lib, 0, 0, g, r_sh[0])
print("11 host G1 scalar muls ms:", (time.perf_counter() - t0) * 1e3)
