"""Readers for the snarkjs / circom / Barretenberg file formats on the path (oracle; test infra).

The reference parses these with the un-vendored crate `taceo-circom-types 0.3.1`
(Cargo.lock:4799; used at co-circom/co-circom/src/bin/co-circom.rs:1005-1016).  Layouts follow
the snarkjs binary format and were probed on the fixtures under test_vectors/ (SURVEY.md 8c).
All field elements returned here are canonical python ints.
"""
import json
import struct
from .fields import curve_by_q, inv


def _sections(data, magic):
    assert data[:4] == magic, "bad magic"
    _version, nsec = struct.unpack_from("<II", data, 4)
    off = 12
    secs = {}
    for _ in range(nsec):
        typ, ln = struct.unpack_from("<IQ", data, off)
        off += 12
        secs.setdefault(typ, (off, ln))
        off += ln
    return secs


def _int_le(b):
    return int.from_bytes(b, "little")


class _Rd:
    def __init__(self, data, off):
        self.d, self.o = data, off

    def u32(self):
        (v,) = struct.unpack_from("<I", self.d, self.o)
        self.o += 4
        return v

    def raw(self, n):
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def int(self, n):
        return _int_le(self.raw(n))


def _mont_out(x, Rinv, p):
    return x * Rinv % p


def _g1(rd, n8q, Rinv, q):
    x = _mont_out(rd.int(n8q), Rinv, q)
    y = _mont_out(rd.int(n8q), Rinv, q)
    return None if x == 0 and y == 0 else (x, y)


def _g2(rd, n8q, Rinv, q):
    c = [_mont_out(rd.int(n8q), Rinv, q) for _ in range(4)]
    return None if not any(c) else ((c[0], c[1]), (c[2], c[3]))


def read_wtns(path):
    """witness.wtns: section 1 header (n8, r, nVars), section 2 = nVars canonical LE values."""
    data = open(path, "rb").read()
    secs = _sections(data, b"wtns")
    rd = _Rd(data, secs[1][0])
    n8 = rd.u32()
    r = rd.int(n8)
    nvars = rd.u32()
    rd = _Rd(data, secs[2][0])
    vals = [rd.int(n8) for _ in range(nvars)]
    return r, vals


def read_groth16_zkey(path):
    """Groth16 .zkey -> dict.  Points are affine canonical ints (None = infinity); `coeffs` is the
    list of (matrix, row, signal, value) of section 4 with value converted from R^2-Montgomery."""
    data = open(path, "rb").read()
    secs = _sections(data, b"zkey")
    rd = _Rd(data, secs[1][0])
    assert rd.u32() == 1, "not a groth16 zkey"
    rd = _Rd(data, secs[2][0])
    n8q = rd.u32()
    q = rd.int(n8q)
    n8r = rd.u32()
    r = rd.int(n8r)
    curve = curve_by_q(q)
    assert curve.r == r
    Rq_inv = inv(pow(2, 8 * n8q, q), q)
    Rr_inv = inv(pow(2, 8 * n8r, r), r)
    n_vars, n_public, domain_size = rd.u32(), rd.u32(), rd.u32()
    z = dict(curve=curve, q=q, r=r, n8q=n8q, n8r=n8r, n_vars=n_vars, n_public=n_public,
             domain_size=domain_size)
    z["alpha_g1"] = _g1(rd, n8q, Rq_inv, q)
    z["beta_g1"] = _g1(rd, n8q, Rq_inv, q)
    z["beta_g2"] = _g2(rd, n8q, Rq_inv, q)
    z["gamma_g2"] = _g2(rd, n8q, Rq_inv, q)
    z["delta_g1"] = _g1(rd, n8q, Rq_inv, q)
    z["delta_g2"] = _g2(rd, n8q, Rq_inv, q)

    def pts(sec, fn, sz):
        off, ln = secs[sec]
        rd = _Rd(data, off)
        return [fn(rd, n8q, Rq_inv, q) for _ in range(ln // sz)]

    z["ic"] = pts(3, _g1, 2 * n8q)
    rd = _Rd(data, secs[4][0])
    ncoef = rd.u32()
    coeffs = []
    Rr_inv2 = Rr_inv * Rr_inv % r
    for _ in range(ncoef):
        m, row, sig = rd.u32(), rd.u32(), rd.u32()
        coeffs.append((m, row, sig, rd.int(n8r) * Rr_inv2 % r))
    z["coeffs"] = coeffs
    z["a_query"] = pts(5, _g1, 2 * n8q)
    z["b_g1_query"] = pts(6, _g1, 2 * n8q)
    z["b_g2_query"] = pts(7, _g2, 4 * n8q)
    z["l_query"] = pts(8, _g1, 2 * n8q)  # private variables only (n_vars - n_public - 1)
    z["h_query"] = pts(9, _g1, 2 * n8q)
    return z


def zkey_matrices(z):
    """ConstraintMatrices as the reference builds them from zkey section 4
    (fields used: co-groth16/src/lib.rs:262-272).  The section holds the A and B rows incl. the
    n_public+1 "public input" rows; the reference's matrices exclude those rows
    (num_constraints = domain rows - ...), and reduction.rs:111-113 re-inserts them."""
    n_pub1 = z["n_public"] + 1
    rows = {0: {}, 1: {}}
    max_row = -1
    for m, row, sig, val in z["coeffs"]:
        rows[m].setdefault(row, []).append((val, sig))
        max_row = max(max_row, row)
    n_rows = max_row + 1
    num_constraints = n_rows - n_pub1
    a = [rows[0].get(i, []) for i in range(num_constraints)]
    b = [rows[1].get(i, []) for i in range(num_constraints)]
    # sanity: the trailing rows of A are exactly  1 * signal k  for k = 0..n_public
    for k in range(n_pub1):
        assert rows[0].get(num_constraints + k, []) == [(1, k)], "unexpected public row"
        assert rows[1].get(num_constraints + k, []) == []
    return dict(a=a, b=b, num_constraints=num_constraints, num_instance_variables=n_pub1,
                num_witness_variables=z["n_vars"] - n_pub1)


def read_plonk_zkey(path, full=True):
    """Plonk .zkey (protocol 2; parsed in the reference by taceo-circom-types' plonk::Zkey): header with k1, k2
    and the selector / permutation commitments, additions, wire maps, the selector, sigma and Lagrange
    polynomials (n coefficients followed by 4n evaluations over the extended domain each) and the p_tau points.
    Fields used by the prover: co-plonk/src/round1.rs:109-224, round2.rs:99-160, round3.rs:300-420,
    round4.rs:143-144, round5.rs:120-260."""
    data = open(path, "rb").read()
    secs = _sections(data, b"zkey")
    rd = _Rd(data, secs[1][0])
    assert rd.u32() == 2, "not a plonk zkey"
    rd = _Rd(data, secs[2][0])
    n8q = rd.u32()
    q = rd.int(n8q)
    n8r = rd.u32()
    r = rd.int(n8r)
    curve = curve_by_q(q)
    Rq_inv = inv(pow(2, 8 * n8q, q), q)
    Rr_inv = inv(pow(2, 8 * n8r, r), r)
    z = dict(curve=curve, q=q, r=r, n8q=n8q, n8r=n8r)
    for k in ("n_vars", "n_public", "domain_size", "n_additions", "n_constraints"):
        z[k] = rd.u32()
    z["k1"] = rd.int(n8r) * Rr_inv % r
    z["k2"] = rd.int(n8r) * Rr_inv % r
    for k in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3"):
        z["vk_" + k] = _g1(rd, n8q, Rq_inv, q)
    z["x2"] = _g2(rd, n8q, Rq_inv, q)
    rd = _Rd(data, secs[3][0])
    adds = []
    for _ in range(z["n_additions"]):
        s1, s2 = rd.u32(), rd.u32()
        f1 = rd.int(n8r) * Rr_inv % r
        f2 = rd.int(n8r) * Rr_inv % r
        adds.append((s1, s2, f1, f2))
    z["additions"] = adds
    for name, sec in (("map_a", 4), ("map_b", 5), ("map_c", 6)):
        rd = _Rd(data, secs[sec][0])
        z[name] = [rd.u32() for _ in range(z["n_constraints"])]
    if full:
        n = z["domain_size"]

        def poly(rd):
            co = [rd.int(n8r) * Rr_inv % r for _ in range(n)]
            ev = [rd.int(n8r) * Rr_inv % r for _ in range(4 * n)]
            return dict(coeffs=co, evals=ev)
        for name, sec in (("qm", 7), ("ql", 8), ("qr", 9), ("qo", 10), ("qc", 11)):
            z[name] = poly(_Rd(data, secs[sec][0]))
        rd = _Rd(data, secs[12][0])
        z["s1"], z["s2"], z["s3"] = poly(rd), poly(rd), poly(rd)
        rd = _Rd(data, secs[13][0])
        z["lagrange"] = [poly(rd) for _ in range(max(1, z["n_public"]))]
        assert rd.o == secs[13][0] + secs[13][1], "lagrange section size"
    off, ln = secs[14]
    rd = _Rd(data, off)
    z["p_tau"] = [_g1(rd, n8q, Rq_inv, q) for _ in range(ln // (2 * n8q))]
    return z


def read_plonk_vk_json(path):
    d = json.load(open(path))
    p1 = lambda a: None if int(a[2]) == 0 else (int(a[0]), int(a[1]))
    vk = dict(n_public=d["nPublic"], power=d["power"], k1=int(d["k1"]), k2=int(d["k2"]), w=int(d["w"]))
    for k in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"):
        vk[k.lower()] = p1(d[k])
    x = d["X_2"]
    vk["x2"] = ((int(x[0][0]), int(x[0][1])), (int(x[1][0]), int(x[1][1])))
    return vk


def read_plonk_proof_json(path):
    d = json.load(open(path))
    p1 = lambda a: None if int(a[2]) == 0 else (int(a[0]), int(a[1]))
    out = {k.lower(): p1(d[k]) for k in ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")}
    for k in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"):
        out[k] = int(d[k])
    return out


def plonk_proof_to_json(proof, curve_name="bn128"):
    """PlonkProof as snarkjs writes it (round5.rs:50-70)."""
    pt = lambda P: ["0", "1", "0"] if P is None else [str(P[0]), str(P[1]), "1"]
    d = {}
    for k, name in (("a", "A"), ("b", "B"), ("c", "C"), ("z", "Z"), ("t1", "T1"), ("t2", "T2"), ("t3", "T3")):
        d[name] = pt(proof[k])
    for k in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"):
        d[k] = str(proof[k])
    d["Wxi"], d["Wxiw"] = pt(proof["wxi"]), pt(proof["wxiw"])
    d["protocol"], d["curve"] = "plonk", curve_name
    return d


def read_vk_json(path):
    """snarkjs verification_key.json -> dict of affine points (canonical ints)."""
    j = json.load(open(path))

    def p1(a):
        return (int(a[0]), int(a[1]))

    def p2(a):
        return ((int(a[0][0]), int(a[0][1])), (int(a[1][0]), int(a[1][1])))

    return dict(alpha_g1=p1(j["vk_alpha_1"]), beta_g2=p2(j["vk_beta_2"]), gamma_g2=p2(j["vk_gamma_2"]),
                delta_g2=p2(j["vk_delta_2"]), ic=[p1(x) for x in j["IC"]], n_public=j["nPublic"])


def read_proof_json(path):
    j = json.load(open(path))
    A = (int(j["pi_a"][0]), int(j["pi_a"][1]))
    B = ((int(j["pi_b"][0][0]), int(j["pi_b"][0][1])), (int(j["pi_b"][1][0]), int(j["pi_b"][1][1])))
    C = (int(j["pi_c"][0]), int(j["pi_c"][1]))
    return A, B, C


def proof_to_json(A, B, C, curve_name="bn128"):
    """CircomGroth16Proof JSON with decimal strings (co-circom.rs:1055-1066; layout of
    test_vectors/Groth16/bn254/multiplier2/circom.proof).  Key order is fixed so that byte
    comparison between two producers is meaningful."""
    d = {
        "pi_a": [str(A[0]), str(A[1]), "1"],
        "pi_b": [[str(B[0][0]), str(B[0][1])], [str(B[1][0]), str(B[1][1])], ["1", "0"]],
        "pi_c": [str(C[0]), str(C[1]), "1"],
        "protocol": "groth16",
        "curve": curve_name,
    }
    return json.dumps(d, separators=(",", ":"))


def read_bn254_crs_g1(path, n, offset=0):
    """co-noir-common/src/crs/parse.rs:93-101,154-158: 64 B/point, x then y, big-endian canonical."""
    with open(path, "rb") as f:
        f.seek(64 * offset)
        raw = f.read(64 * n)
    return [(int.from_bytes(raw[i:i + 32], "big"), int.from_bytes(raw[i + 32:i + 64], "big"))
            for i in range(0, len(raw), 64)]
