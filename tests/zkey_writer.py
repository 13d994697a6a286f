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
