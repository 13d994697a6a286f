"""Plonk round 1 restatement (oracle; test infrastructure only) -- exists to PIN the oracle's
NTT + MSM against the reference's bit-exact known-answer tests
(co-plonk/src/round1.rs:331-420), which run the Plain driver with deterministic blinders b[i] = i.

Follows co-plonk/src/round1.rs:109-134 (compute_single_wire_poly), :137-189, :191-224
(calculate_additions), lib.rs:138-180 (get_witness, blind_coefficients), types.rs:76-100 (domains
with snarkjs roots), :110-126 (witness[0] := 0).
"""
from .fields import roots_of_unity
from .ntt import ifft
from .ec import g1 as _g1


def round1_commitments(z, full_witness, blinders=None):
    curve = z["curve"]
    r = curve.r
    n = z["domain_size"]
    _, roots = roots_of_unity(r)
    gen = roots[n.bit_length() - 1]
    if blinders is None:
        blinders = list(range(11))  # Round1Challenges::deterministic (round1.rs:95-104)
    npub = z["n_public"]
    public_inputs = [0] + list(full_witness[1:npub + 1])  # types.rs:118-120
    witness = list(full_witness[npub + 1:])
    additions = []

    def get_witness(idx):  # lib.rs:138-160
        if idx <= npub:
            return public_inputs[idx]
        if idx < z["n_vars"] - z["n_additions"]:
            return witness[idx - npub - 1]
        if idx < z["n_vars"]:
            return additions[idx + z["n_additions"] - z["n_vars"]]
        raise ValueError("corrupted witness")

    for s1, s2, f1, f2 in z["additions"]:  # round1.rs:191-224
        additions.append((get_witness(s1) * f1 + get_witness(s2) * f2) % r)

    G1 = _g1(curve)
    out = []
    for wire_map, blind in ((z["map_a"], blinders[0:2]), (z["map_b"], blinders[2:4]), (z["map_c"], blinders[4:6])):
        buf = [get_witness(i) for i in wire_map] + [0] * (n - len(wire_map))
        poly = ifft(buf, gen, r)
        # blind_coefficients (lib.rs:163-178): poly[i] -= rev(blind)[i]; then append rev(blind)
        rev = list(reversed(blind))
        for i, c in enumerate(rev):
            poly[i] = (poly[i] - c) % r
        poly += rev
        out.append(G1.msm(z["p_tau"][:len(poly)], poly))
    return out
