"""Times the device Plonk prover (plain driver) on the synthetic workload and checks the proof with the oracle's
pairing verifier.  usage: time_plonk.py [log_n ...]   (default 16 18 20).  One JSON line."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from co_snarks_b200 import binding as B
from workloads.synth_plonk import SynthPlonk

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
stream = torch.cuda.Stream()
ctx = B.Context(0, stream=stream.cuda_stream)
out = {}
for lg in sizes:
    t0 = time.time()
    syn = SynthPlonk(ctx, lg)
    pk = syn.make_key()
    setup_s = time.time() - t0
    rng = random.Random(5)
    bl = B.ints_to_limbs(B.to_mont_ints([rng.randrange(R) for _ in range(11)], R, 4), 4)
    ms = []
    for i in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pts, evs = pk.prove_plain(syn.public_inputs, syn.private_witness, bl)  # host buffers in, proof out (synchronous)
        ms.append((time.perf_counter() - t0) * 1e3)
    ok = None
    if os.environ.get("CS_PLONK_VERIFY", "1") == "1":
        from helpers import Conv, plonk_proof_from_device
        from oracle import plonk as OP
        from oracle.fields import BN254
        from oracle.pairing_bn254 import pairing_product_is_one
        proof = plonk_proof_from_device(Conv("bn254"), pts, evs)
        ok = bool(OP.verify(BN254, syn.vk_ints(), proof, syn.full_witness[1:syn.n_public + 1], pairing_product_is_one))
    cpu = None
    if os.environ.get("CS_PLONK_CPU", "0") == "1" and lg <= int(os.environ.get("CS_PLONK_CPU_MAX_LG", "18")):
        # CPU baseline of the Plonk row: the C/OpenMP restatement (oracle/c/plonk.inc) on the host cores, same key
        from oracle.c import run as OC
        cpu_s = OC.time_plonk(syn.key, syn.public_inputs, syn.private_witness, bl, reps=1)
        cpu = {"seconds_per_proof": round(cpu_s, 3), "cores": OC.lib().oracle_num_threads(), "kind": "port"}
    t = sum(ms[2:]) / len(ms[2:])
    out["2p%d" % lg] = {"ms_per_proof": round(t, 3), "proofs_per_s": round(1e3 / t, 2), "verified": ok,
                        "n_additions": syn.n_additions, "setup_s": round(setup_s, 1), "launches": ctx.launch_count(),
                        "cpu_baseline": cpu}
    pk.free()
print(json.dumps(out))
