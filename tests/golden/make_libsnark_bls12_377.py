"""Pins the oracle's LibSnarkReduction restatement (oracle/groth16.py: witness_map_libsnark, following
co-groth16/src/groth16/reduction.rs:241-342) on a fixture the REFERENCE holds: the Penumbra `output` circuit under
test_vectors/Groth16/bls12_377 -- what `proof_libsnark_penumbra_output_bls12_377` (co-groth16/src/lib.rs:231-298) runs:
ProvingKey / VerifyingKey / Matrix files in arkworks' uncompressed serialisation, witness.wtns, plain_prove with
LibSnarkReduction, verify under circuit.vk.

Run here (needs /root/reference; pure Python, a few minutes):   python tests/golden/make_libsnark_bls12_377.py
Writes tests/golden/libsnark_bls12_377_penumbra_output.json.gz: the matrices, the witness, the verification key, the
SHA-256 of the h coefficients the oracle computes, and the proof it assembles for fixed (r, s) from the reference's
proving key -- which tests/test_oracle_golden.py re-verifies with the BLS12-377 pairing (and re-derives h)."""
import gzip
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import groth16 as OG
from oracle import pairing_bls12_377 as P

CIRCUIT = sys.argv[1] if len(sys.argv) > 1 else "penumbra_output"   # also: penumbra_spend, penumbra_delegator_vote
WRITE = CIRCUIT == "penumbra_output"                                # the other two are only verified, not stored
SRC = "/root/reference/test_vectors/Groth16/bls12_377/" + CIRCUIT
OUT = os.path.join(ROOT, "tests", "golden", "libsnark_bls12_377_%s.json.gz" % CIRCUIT)


def h_digest(h):
    return hashlib.sha256(b"".join(int(x).to_bytes(32, "little") for x in h)).hexdigest()


def read_wtns_lenient(path):
    """The Penumbra witness files carry section ids / lengths of 0 (they were not written by snarkjs); the layout is
    otherwise the wtns one: header, [id u32, len u64], n8, prime, nVars, [id, len], nVars values."""
    import struct
    d = open(path, "rb").read()
    assert d[:4] == b"wtns"
    off = 12 + 12
    (n8,) = struct.unpack_from("<I", d, off)
    off += 4
    prime = int.from_bytes(d[off:off + n8], "little")
    off += n8
    (nv,) = struct.unpack_from("<I", d, off)
    off += 4 + 12
    assert len(d) - off == nv * n8
    return prime, [int.from_bytes(d[off + k * n8:off + (k + 1) * n8], "little") for k in range(nv)]


def main():
    t0 = time.time()
    pk = P.read_pk(open(os.path.join(SRC, "circuit.pk"), "rb").read())
    vk, _ = P.read_vk(open(os.path.join(SRC, "circuit.vk"), "rb").read())
    a, b, c = (P.read_matrix(open(os.path.join(SRC, n + ".bin"), "rb").read()) for n in "abc")
    r, w = read_wtns_lenient(os.path.join(SRC, "witness.wtns"))
    assert r == P.R
    ni = len(pk["b_g1_query"]) - len(pk["l_query"])          # lib.rs:262-264
    nw = len(pk["a_query"]) - len(pk["b_g1_query"]) + len(pk["l_query"])
    m = {"num_constraints": len(a), "num_instance_variables": ni, "num_witness_variables": nw, "a": a, "b": b, "c": c}
    assert len(w) == ni + nw
    pub, wit = w[:ni], w[ni:]
    print("parsed: %d constraints, %d instance, %d witness variables, h_query %d (%.1f s)" % (len(a), ni, nw, len(pk["h_query"]), time.time() - t0))
    h = OG.witness_map_libsnark(m, pub, wit, P.R)
    print("witness map done (%.1f s), h digest %s" % (time.time() - t0, h_digest(h)))
    G1, G2 = P.g1(), P.g2()
    r_rand, s_rand = 0x1234567890abcdef1234567890abcdef % P.R, 0xfedcba0987654321fedcba0987654321 % P.R
    v = pk["vk"]
    inputs = pub[1:]
    A = OG._calc_coeff(G1, G1.mul(pk["delta_g1"], r_rand), pk["a_query"], v["alpha_g1"], inputs, wit)
    B1 = OG._calc_coeff(G1, G1.mul(pk["delta_g1"], s_rand), pk["b_g1_query"], pk["beta_g1"], inputs, wit)
    B2 = OG._calc_coeff(G2, G2.mul(v["delta_g2"], s_rand), pk["b_g2_query"], v["beta_g2"], inputs, wit)
    L = G1.msm(pk["l_query"], wit)
    H = G1.msm(pk["h_query"], h)
    print("MSMs done (%.1f s)" % (time.time() - t0))
    C = G1.add(G1.mul(A, s_rand), G1.mul(B1, r_rand))
    C = G1.add(C, G1.neg(G1.mul(pk["delta_g1"], r_rand * s_rand % P.R)))
    C = G1.add(G1.add(C, L), H)
    ok = P.groth16_verify(vk, inputs, (A, B2, C))
    print("proof verifies under the reference's circuit.vk:", ok)
    assert ok, "the oracle's LibSnark proof does not verify"
    if not WRITE:
        return
    enc = lambda x: [str(c) for c in x] if isinstance(x, tuple) and isinstance(x[0], int) else [[str(c) for c in y] for y in x]
    out = {"source": SRC, "r": str(P.R), "matrices": {"a": [[(str(cf), ix) for cf, ix in row] for row in a],
                                                       "b": [[(str(cf), ix) for cf, ix in row] for row in b],
                                                       "c": [[(str(cf), ix) for cf, ix in row] for row in c]},
           "num_instance_variables": ni, "num_witness_variables": nw, "witness": [str(x) for x in w],
           "vk": {"alpha_g1": enc(vk["alpha_g1"]), "beta_g2": enc(vk["beta_g2"]), "gamma_g2": enc(vk["gamma_g2"]),
                  "delta_g2": enc(vk["delta_g2"]), "ic": [enc(p) for p in vk["ic"]]},
           "rs": [str(r_rand), str(s_rand)], "h_sha256": h_digest(h),
           "proof": {"a": enc(A), "b": enc(B2), "c": enc(C)}}
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
