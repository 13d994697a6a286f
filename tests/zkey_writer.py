"""Writes snarkjs-format Groth16 .zkey / .wtns files from the golden fixtures (test helper), so the ingest
path (cs_groth16_pk_from_zkey / cs_wtns_read) is testable where /root/reference is not mounted.  Layout per
SURVEY.md 8c (probed on the reference's test vectors)."""
import struct


def _le(v, n):
    return int(v).to_bytes(n, "little")


def write_zkey(path, z, m, n8=32):
    q, r = z["q"], z["r"]
    Rq, Rr = pow(2, 8 * n8, q), pow(2, 8 * n8, r)

    def g1(P):
        return b"\0" * (2 * n8) if P is None else _le(P[0] * Rq % q, n8) + _le(P[1] * Rq % q, n8)

    def g2(P):
        if P is None:
            return b"\0" * (4 * n8)
        return b"".join(_le(c * Rq % q, n8) for c in (P[0][0], P[0][1], P[1][0], P[1][1]))

    ni = m["num_instance_variables"]
    nc = m["num_constraints"]
    coeffs = []
    for mat, rows in ((0, m["a"]), (1, m["b"])):
        for row, ents in enumerate(rows):
            for cf, sig in ents:
                coeffs.append((mat, row, sig, cf))
    for k in range(ni):  # the public-input rows snarkjs appends to A
        coeffs.append((0, nc + k, k, 1))
    secs = {
        1: struct.pack("<I", 1),
        2: struct.pack("<I", n8) + _le(q, n8) + struct.pack("<I", n8) + _le(r, n8) +
           struct.pack("<III", z["n_vars"], z["n_public"], z["domain_size"]) +
           g1(z["alpha_g1"]) + g1(z["beta_g1"]) + g2(z["beta_g2"]) + g2(z["gamma_g2"]) + g1(z["delta_g1"]) + g2(z["delta_g2"]),
        3: b"".join(g1(P) for P in z["ic"]),
        4: struct.pack("<I", len(coeffs)) + b"".join(struct.pack("<III", a, b, c) + _le(v * Rr * Rr % r, n8) for a, b, c, v in coeffs),
        5: b"".join(g1(P) for P in z["a_query"]),
        6: b"".join(g1(P) for P in z["b_g1_query"]),
        7: b"".join(g2(P) for P in z["b_g2_query"]),
        8: b"".join(g1(P) for P in z["l_query"]),
        9: b"".join(g1(P) for P in z["h_query"]),
        10: b"",
    }
    with open(path, "wb") as f:
        f.write(b"zkey" + struct.pack("<II", 1, len(secs)))
        for t in sorted(secs):
            f.write(struct.pack("<IQ", t, len(secs[t])) + secs[t])


def write_wtns(path, r, values, n8=32):
    s1 = struct.pack("<I", n8) + _le(r, n8) + struct.pack("<I", len(values))
    s2 = b"".join(_le(v, n8) for v in values)
    with open(path, "wb") as f:
        f.write(b"wtns" + struct.pack("<II", 2, 2))
        f.write(struct.pack("<IQ", 1, len(s1)) + s1)
        f.write(struct.pack("<IQ", 2, len(s2)) + s2)


def write_plonk_zkey(path, z):
    """snarkjs Plonk .zkey (protocol 2) from the oracle-style dict of tests/helpers.golden_plonk (layout: SURVEY.md 8c,
    oracle/formats.read_plonk_zkey).  Section 13 needs the Lagrange coefficients as well as the evaluations."""
    q, r = z["q"], z["r"]
    n8q, n8r = (q.bit_length() + 63) // 64 * 8, (r.bit_length() + 63) // 64 * 8
    Rq, Rr = pow(2, 8 * n8q, q), pow(2, 8 * n8r, r)
    fr = lambda v: _le(v * Rr % r, n8r)

    def g1(P):
        return b"\0" * (2 * n8q) if P is None else _le(P[0] * Rq % q, n8q) + _le(P[1] * Rq % q, n8q)

    def g2(P):
        return b"".join(_le(c * Rq % q, n8q) for c in (P[0][0], P[0][1], P[1][0], P[1][1]))

    def poly(P):
        return b"".join(fr(v) for v in P["coeffs"]) + b"".join(fr(v) for v in P["evals"])
    secs = {
        1: struct.pack("<I", 2),
        2: struct.pack("<I", n8q) + _le(q, n8q) + struct.pack("<I", n8r) + _le(r, n8r) +
           struct.pack("<IIIII", z["n_vars"], z["n_public"], z["domain_size"], z["n_additions"], z["n_constraints"]) +
           fr(z["k1"]) + fr(z["k2"]) + b"".join(g1(z["vk_" + k]) for k in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3")) + g2(z["x2"]),
        3: b"".join(struct.pack("<II", a, b) + fr(f1) + fr(f2) for a, b, f1, f2 in z["additions"]),
        4: b"".join(struct.pack("<I", v) for v in z["map_a"]),
        5: b"".join(struct.pack("<I", v) for v in z["map_b"]),
        6: b"".join(struct.pack("<I", v) for v in z["map_c"]),
        7: poly(z["qm"]), 8: poly(z["ql"]), 9: poly(z["qr"]), 10: poly(z["qo"]), 11: poly(z["qc"]),
        12: poly(z["s1"]) + poly(z["s2"]) + poly(z["s3"]),
        13: b"".join(poly(P) for P in z["lagrange"]),
        14: b"".join(g1(P) for P in z["p_tau"]),
    }
    with open(path, "wb") as f:
        f.write(b"zkey" + struct.pack("<II", 1, len(secs)))
        for t in sorted(secs):
            f.write(struct.pack("<IQ", t, len(secs[t])) + secs[t])
