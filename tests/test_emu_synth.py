"""CPU-only: the synthetic-key generator (workloads/synth_groth16.py) yields keys whose proofs pass the
same pairing check as the reference's snarkjs fixtures -- here on the emulated kernels at 2^4."""
import random

import kernel_checks as K
from helpers import Conv
from oracle.pairing_bn254 import groth16_verify
from workloads.synth_groth16 import SynthGroth16


def test_emu_synth_key_proves_and_verifies(emu_ctx):
    cv = Conv("bn254")
    syn = SynthGroth16(emu_ctx, 4)
    pk = syn.make_key()
    rng = random.Random(3)
    r_, s_ = rng.randrange(cv.r), rng.randrange(cv.r)
    A, Bp, Cp = pk.prove_plain(syn.public_inputs, syn.private_witness, cv.fr([r_]), cv.fr([s_]))
    proof = (cv.pt1(A), cv.pt2(Bp), cv.pt1(Cp))
    vk = syn.vk_ints()
    assert groth16_verify(vk, syn.witness[1:2], proof)
    assert not groth16_verify(vk, [syn.witness[1] + 1], proof)
    pk.free()
