"""CPU-only: the Python oracle against the committed golden vectors (tests/golden/make_golden.py):
the reference's Plonk round-1 known answers, the snarkjs proofs/verification keys of the reference's
Groth16 fixtures, and -- when /root/reference is mounted -- the raw fixture files themselves."""
import json
import os
import random

import pytest

from helpers import golden_groth16, gp1, gp2, ih, load_golden
from oracle import formats as F
from oracle import groth16 as OG
from oracle.ec import g1 as og1
from oracle.fields import BLS12_381, BN254, CURVES, roots_of_unity
from oracle.ntt import ifft
from oracle.pairing_bn254 import groth16_verify

REF = "/root/reference"


@pytest.mark.parametrize("name", ["multiplier2", "poseidon"])
def test_snarkjs_proof_verifies_and_oracle_proof_matches_golden(name):
    z, m, w, g = golden_groth16(name)
    vk = OG.vk_from_zkey(z)
    public = [ih(x) for x in g["public"]]
    sp = g["snarkjs_proof"]
    snark = (gp1(sp["a"]), gp2(sp["b"]), gp1(sp["c"]))
    assert groth16_verify(vk, public, snark)                      # co-groth16/src/lib.rs:72-91
    assert not groth16_verify(vk, [public[0] + 1] + public[1:], snark)
    ni = m["num_instance_variables"]
    assert OG.witness_map_plain(m, w[:ni], w[ni:], z["r"], 28) == [ih(x) for x in g["h"]]
    for pr in g["oracle_proofs"]:
        proof = OG.prove_plain(z, m, w, ih(pr["r"]), ih(pr["s"]))
        assert F.proof_to_json(*proof) == pr["json"]
        assert groth16_verify(vk, public, proof)                  # co-groth16/src/lib.rs:40-69


def test_rep3_emulation_equals_plain_for_summed_randomness():
    z, m, w, g = golden_groth16("multiplier2")
    proof, r_tot, s_tot = OG.prove_rep3(z, m, w, random.Random(7))
    assert proof == OG.prove_plain(z, m, w, r_tot, s_tot)          # all parties open the same proof
    assert groth16_verify(OG.vk_from_zkey(z), [ih(x) for x in g["public"]], proof)


@pytest.mark.parametrize("curve,name", [("bn254", "multiplier2"), ("bls12_381", "poseidon")])
def test_plonk_round1_kat(curve, name):
    """iNTT + MSM pinned bit-for-bit on the reference's known answers (co-plonk/src/round1.rs:351-371,397-417)."""
    g = load_golden("plonk_round1_%s_%s" % (curve, name))
    c = CURVES[curve]
    n = g["domain_size"]
    _, roots = roots_of_unity(c.r)
    gen = roots[n.bit_length() - 1]
    assert gen == ih(g["group_gen"])
    G = og1(c)
    p_tau = [gp1(P) for P in g["p_tau"]]
    for wire, exp in zip(g["wires"], g["expected_commitments"]):
        poly = ifft([ih(x) for x in wire["buffer"]], gen, c.r)
        assert poly == [ih(x) for x in wire["poly"]]
        blinded = [ih(x) for x in wire["blinded"]]
        assert G.msm(p_tau[:len(blinded)], blinded) == gp1(exp)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_file_parsers_agree_with_golden():
    base = REF + "/test_vectors/Groth16/bn254/poseidon/"
    z = F.read_groth16_zkey(base + "circuit.zkey")
    m = F.zkey_matrices(z)
    _, w = F.read_wtns(base + "witness.wtns")
    zg, mg, wg, g = golden_groth16("poseidon")
    assert w == wg and m["a"] == mg["a"] and m["b"] == mg["b"]
    for k in ("a_query", "b_g2_query", "h_query", "l_query", "alpha_g1", "delta_g2"):
        assert z[k] == zg[k]
    vk = F.read_vk_json(base + "verification_key.json")
    assert vk["ic"] == z["ic"] and vk["gamma_g2"] == z["gamma_g2"]
    pts = F.read_bn254_crs_g1(REF + "/co-noir/co-noir-common/src/crs/bn254_g1.dat", 4)
    assert pts[0] == (1, 2) and all(og1(BN254).on_curve(P) for P in pts)
