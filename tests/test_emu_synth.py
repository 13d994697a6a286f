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


def test_emu_edge_circuits(emu_ctx):
    """Edge shapes the reference's length checks allow: no private witness at all (nw = 0: every MSM over the
    witness is empty), and a single-constraint circuit (domain 4, mostly padding) -- proofs must verify."""
    cv = Conv("bn254")
    r = cv.r
    # x * 1 = x  with x public:  variables (1, x), no private witness
    sys_pub_only = ([[(1, 1)]], [[(1, 0)]], [[(1, 1)]], [1, 12345], 2)
    # a * b = c  with c public, a and b private
    sys_mul = ([[(1, 2)]], [[(1, 3)]], [[(1, 1)]], [1, 33, 3, 11], 2)
    for r1cs in (sys_pub_only, sys_mul):
        syn = SynthGroth16(emu_ctx, 0, r1cs=r1cs)
        pk = syn.make_key()
        rng = random.Random(9)
        A, Bp, Cp = pk.prove_plain(syn.public_inputs, syn.private_witness, cv.fr([rng.randrange(r)]), cv.fr([rng.randrange(r)]))
        proof = (cv.pt1(A), cv.pt2(Bp), cv.pt1(Cp))
        assert groth16_verify(syn.vk_ints(), syn.witness[1:2], proof)
        assert not groth16_verify(syn.vk_ints(), [syn.witness[1] + 1], proof)
        pk.free()
