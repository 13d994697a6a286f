"""One plain Plonk proof on the synthetic circuit -- the command profiled under ncu (launch list)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from co_snarks_b200 import binding as B
from workloads.synth_plonk import SynthPlonk
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
ctx = B.Context(0)
syn = SynthPlonk(ctx, lg)
pk = syn.make_key()
rng = random.Random(5)
bl = B.ints_to_limbs(B.to_mont_ints([rng.randrange(R) for _ in range(11)], R, 4), 4)
print("SETUP_LAUNCHES", ctx.launch_count())
pk.prove_plain(syn.public_inputs, syn.private_witness, bl)
print("TOTAL_LAUNCHES", ctx.launch_count())
