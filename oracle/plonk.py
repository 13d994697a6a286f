"""Plonk prover + verifier restatement (oracle; TEST INFRASTRUCTURE ONLY -- never imported by the product).

The plain driver's path through co-plonk (the snarkjs Plonk prover), function by function:
  transcript   co-plonk/src/types.rs:140-190 (Keccak-256 over big-endian field/point bytes)
  round 1      co-plonk/src/round1.rs:108-320 (wire polynomials, additions, blinding lib.rs:163-178)
  round 2      co-plonk/src/round2.rs:95-250   (beta, gamma, permutation polynomial z)
  round 3      co-plonk/src/round3.rs:20-560   (alpha, quotient t split in three, the ap/bp/cp/zp blinding
                                                bookkeeping and the mul4vec expansion, restated literally)
  round 4      co-plonk/src/round4.rs:100-165  (xi, evaluations)
  round 5      co-plonk/src/round5.rs:78-340   (v, linearisation r, opening polynomials W_xi, W_xiw)
  verifier     co-plonk/src/plonk.rs:28-245
Pinned against the reference's known-answer tests with deterministic blinders b[i] = i
(round1.rs:331-420, round2.rs:280-310, round3.rs:589-634, round4.rs:181-248, round5.rs:369-409), the
transcript KAT (types.rs:201-236) and the verifier-challenge KAT (plonk.rs:251-310); see
tests/test_oracle_golden.py.  All values are canonical python ints; points are affine tuples or None.
"""
from .ec import g1 as _g1
from .fields import inv, roots_of_unity
from .ntt import fft, ifft

# ---------------------------------------------------------------------------------------------
# Keccak-256 (original padding 0x01, as sha3::Keccak256) -- hashlib only has the NIST variant.
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(A):
    for rc in _RC:
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= rc
    return A


def keccak256(data):
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        blk = msg[off:off + rate]
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = _keccak_f(A)
    out = b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


class Transcript:
    """types.rs:140-190: scalars and coordinates are hashed as fixed-width BIG-endian canonical values; the
    point at infinity as 2 * byte_len zero bytes; challenge = digest read big-endian, reduced mod r."""

    def __init__(self, curve):
        self.curve = curve
        self.buf = bytearray()
        self.qlen = (curve.q.bit_length() + 7) // 8
        self.rlen = (curve.r.bit_length() + 7) // 8

    def add_scalar(self, s):
        self.buf += int(s).to_bytes(self.rlen, "big")

    def add_point(self, P):
        if P is None:
            self.buf += bytes(2 * self.qlen)
        else:
            self.buf += int(P[0]).to_bytes(self.qlen, "big") + int(P[1]).to_bytes(self.qlen, "big")

    def get_challenge(self):
        return int.from_bytes(keccak256(bytes(self.buf)), "big") % self.curve.r


# ---------------------------------------------------------------------------------------------
def domains(curve, n):
    """types.rs:76-111: snarkjs roots; (w_n, w_4n, roots[2])."""
    _, roots = roots_of_unity(curve.r)
    pw = n.bit_length() - 1
    return roots[pw], roots[pw + 2], roots[2]


def _fft_ext(poly, n4, gen4, r):
    return fft(list(poly) + [0] * (n4 - len(poly)), gen4, r)


def _blind(poly, coeff_rev, r):
    """lib.rs:163-178 blind_coefficients: poly[i] -= rev(coeff)[i]; append rev(coeff)."""
    rev = list(reversed(coeff_rev))
    for i, c in enumerate(rev):
        poly[i] = (poly[i] - c) % r
    return poly + rev


def _eval(poly, x, r):
    acc = 0
    for c in reversed(poly):
        acc = (acc * x + c) % r
    return acc


def lagrange_evaluations(power, n_public, xi, w, r):
    """lib.rs:181-207 calculate_lagrange_evaluations -> (l[max(1, n_public)], xi^n)."""
    xin = xi
    for _ in range(power):
        xin = xin * xin % r
    n = 1 << power
    zh = (xin - 1) % r
    ls, wi = [], 1
    for _ in range(max(1, n_public)):
        ls.append(wi * zh % r * inv(n * (xi - wi) % r, r) % r)
        wi = wi * w % r
    return ls, xin


def _div_by_zerofier1(p, beta, r):
    """round5.rs:78-93 with n = 1: in-place synthetic division by (X - beta), dropping the last entry."""
    ib = inv(beta, r)
    p = list(p)
    p[0] = p[0] * (-ib) % r
    for i in range(1, len(p)):
        p[i] = (p[i - 1] - p[i]) * ib % r
    return p[:-1]


def prove(z, full_witness, blinders=None, trace=None):
    """Plonk::plain_prove (lib.rs:271-281).  blinders: the 11 round-1 field elements b[0..11)
    (Round1Challenges; deterministic() = [0..11) in the reference's tests).  Returns the PlonkProof as a dict."""
    curve = z["curve"]
    r = curve.r
    n = z["domain_size"]
    n4 = 4 * n
    power = n.bit_length() - 1
    w_n, w_4n, w_4 = domains(curve, n)
    b = list(range(11)) if blinders is None else [int(x) % r for x in blinders]
    npub = z["n_public"]
    G1 = _g1(curve)
    ptau = z["p_tau"]
    commit = lambda poly: G1.msm(ptau[:len(poly)], poly)

    # ---- init round + round 1 (round1.rs)
    public0 = [0] + [int(x) % r for x in full_witness[1:npub + 1]]  # types.rs:118-120
    witness = [int(x) % r for x in full_witness[npub + 1:]]
    additions = []

    def get_witness(idx):  # lib.rs:138-160
        if idx <= npub:
            return public0[idx]
        if idx < z["n_vars"] - z["n_additions"]:
            return witness[idx - npub - 1]
        if idx < z["n_vars"]:
            return additions[idx + z["n_additions"] - z["n_vars"]]
        raise ValueError("Cannot index into witness %d" % idx)

    for s1, s2, f1, f2 in z["additions"]:
        additions.append((get_witness(s1) * f1 + get_witness(s2) * f2) % r)
    buf, poly, ev = {}, {}, {}
    for k, wire_map, bl in (("a", z["map_a"], b[0:2]), ("b", z["map_b"], b[2:4]), ("c", z["map_c"], b[4:6])):
        buf[k] = [get_witness(i) for i in wire_map] + [0] * (n - len(wire_map))
        p = ifft(buf[k], w_n, r)
        ev[k] = _fft_ext(p, n4, w_4n, r)  # evaluations of the UNBLINDED polynomial (round1.rs:124-131)
        poly[k] = _blind(p, bl, r)
    proof = {k: commit(poly[k]) for k in ("a", "b", "c")}
    public_inputs = public0[1:]  # round1.rs:32-43

    # ---- round 2 (round2.rs)
    t = Transcript(curve)
    for k in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3"):
        t.add_point(z["vk_" + k])
    for v in public_inputs:
        t.add_scalar(v)
    for k in ("a", "b", "c"):
        t.add_point(proof[k])
    beta = t.get_challenge()
    t = Transcript(curve)
    t.add_scalar(beta)
    gamma = t.get_challenge()
    k1, k2 = z["k1"], z["k2"]
    num, den, wi = [], [], 1
    for i in range(n):
        a_, b_, c_ = buf["a"][i], buf["b"][i], buf["c"][i]
        bw = beta * wi % r
        num.append((a_ + bw + gamma) * (b_ + k1 * bw + gamma) % r * (c_ + k2 * bw + gamma) % r)
        den.append((a_ + beta * z["s1"]["evals"][4 * i] + gamma) * (b_ + beta * z["s2"]["evals"][4 * i] + gamma) % r
                   * (c_ + beta * z["s3"]["evals"][4 * i] + gamma) % r)
        wi = wi * w_n % r
    # array_prod_mul (mpc/plain.rs:199-247): running products; the denominators' are inverted
    buffer_z, pn, pd = [], 1, 1
    for i in range(n):
        pn, pd = pn * num[i] % r, pd * den[i] % r
        if pd == 0:
            raise ZeroDivisionError("Cannot invert zero")
        buffer_z.append(pn * inv(pd, r) % r)
    buffer_z = buffer_z[-1:] + buffer_z[:-1]  # rotate_right(1)
    pz = ifft(buffer_z, w_n, r)
    ev["z"] = _fft_ext(pz, n4, w_4n, r)
    poly["z"] = _blind(pz, b[6:9], r)
    proof["z"] = commit(poly["z"])

    # ---- round 3 (round3.rs)
    t = Transcript(curve)
    t.add_scalar(beta)
    t.add_scalar(gamma)
    t.add_point(proof["z"])
    alpha = t.get_challenge()
    alpha2 = alpha * alpha % r
    z1 = [0, (-1 + w_4) % r, (-2) % r, (-1 - w_4) % r]
    z2 = [0, (-2 * w_4) % r, 4, (2 * w_4) % r]
    z3 = [0, (2 + 2 * w_4) % r, (-8) % r, (2 - 2 * w_4) % r]
    L0, lag = z["lagrange"][0]["evals"], z["lagrange"]
    tv, tzv = [], []
    wi = 1
    for i in range(n4):
        a_, b_, c_, z_ = ev["a"][i], ev["b"][i], ev["c"][i], ev["z"][i]
        zw_ = ev["z"][(n4 + 4 + i) % n4]
        ap = (b[1] + b[0] * wi) % r
        bp = (b[3] + b[2] * wi) % r
        cp = (b[5] + b[4] * wi) % r
        w2 = wi * wi % r
        zp = (b[6] * w2 + b[7] * wi + b[8]) % r
        ww = wi * w_n % r
        zwp = (b[6] * ww * ww + b[7] * ww + b[8]) % r
        m = i % 4
        a_b, a_bp, ap_b, ap_bp = a_ * b_ % r, a_ * bp % r, ap * b_ % r, ap * bp % r
        a0 = (a_bp + ap_b) % r
        if m:
            a0 = (a0 + ap_bp * z1[m]) % r
        qm, ql, qr, qo, qc = (z[k]["evals"][i] for k in ("qm", "ql", "qr", "qo", "qc"))
        e1 = (a_b * qm + a_ * ql + b_ * qr + c_ * qo) % r
        e1z = (a0 * qm + ap * ql + bp * qr + cp * qo) % r
        pi = 0
        for j in range(len(lag)):  # all of zkey.lagrange, as the reference does (round3.rs:372-376)
            pi = (pi - buf["a"][j] * lag[j]["evals"][i]) % r
        e1 = (e1 + pi + qc) % r
        bw = beta * wi % r

        def mul4(a, b_2, c, d, ap_, bp_, cp_, dp_):
            # mul4vec + mul4vec_post (round3.rs:20-108): (a + ap Z)(b + bp Z)(c + cp Z)(d + dp Z) reduced with the
            # extended-domain evaluations of Z_H, Z_H^2, Z_H^3 (z1, z2, z3) -> (value, blinding part)
            ab, abp, apb, apbp = a * b_2 % r, a * bp_ % r, ap_ * b_2 % r, ap_ * bp_ % r
            cd, cdp, cpd, cpdp = c * d % r, c * dp_ % r, cp_ * d % r, cp_ * dp_ % r
            rr = ab * cd % r
            x0 = (apb * cd + abp * cd + ab * cpd + ab * cdp) % r
            x1 = (apbp * cd + apb * cpd + apb * cdp + abp * cpd + abp * cdp + ab * cpdp) % r
            x2 = (abp * cpdp + apb * cpdp + apbp * cdp + apbp * cpd) % r
            x3 = apbp * cpdp % r
            rz = x0
            if m:
                rz = (x0 + x1 * z1[m] + x2 * z2[m] + x3 * z3[m]) % r
            return rr, rz
        e2, e2z = mul4((a_ + bw + gamma) % r, (b_ + bw * k1 + gamma) % r, (c_ + bw * k2 + gamma) % r, z_, ap, bp, cp, zp)
        e3, e3z = mul4((a_ + z["s1"]["evals"][i] * beta + gamma) % r, (b_ + z["s2"]["evals"][i] * beta + gamma) % r,
                       (c_ + z["s3"]["evals"][i] * beta + gamma) % r, zw_, ap, bp, cp, zwp)
        e4 = (z_ - 1) * L0[i] % r * alpha2 % r
        e4z = zp * L0[i] % r * alpha2 % r
        tv.append((e1 + e2 * alpha - e3 * alpha + e4) % r)
        tzv.append((e1z + e2z * alpha - e3z * alpha + e4z) % r)
        wi = wi * w_4n % r
    ct = ifft(tv, w_4n, r)
    for i in range(n):
        ct[i] = (-ct[i]) % r
    for i in range(n, n4):
        ct[i] = (ct[i - n] - ct[i]) % r  # sequential: uses the updated ct[i - n] (division by Z_H)
    ctz = ifft(tzv, w_4n, r)
    tf = [(x + y) % r for x, y in zip(ct, ctz)]
    t1 = tf[:n] + [b[9]]
    t2 = tf[n:2 * n]
    t2[0] = (t2[0] - b[9]) % r
    t2.append(b[10])
    t3 = tf[2 * n:2 * n + n + 6]
    t3[0] = (t3[0] - b[10]) % r
    proof["t1"], proof["t2"], proof["t3"] = commit(t1), commit(t2), commit(t3)

    # ---- round 4 (round4.rs)
    t = Transcript(curve)
    t.add_scalar(alpha)
    for k in ("t1", "t2", "t3"):
        t.add_point(proof[k])
    xi = t.get_challenge()
    xiw = xi * w_n % r
    proof["eval_a"], proof["eval_b"], proof["eval_c"] = (_eval(poly[k], xi, r) for k in ("a", "b", "c"))
    proof["eval_zw"] = _eval(poly["z"], xiw, r)
    proof["eval_s1"] = _eval(z["s1"]["coeffs"], xi, r)
    proof["eval_s2"] = _eval(z["s2"]["coeffs"], xi, r)

    # ---- round 5 (round5.rs)
    t = Transcript(curve)
    for v in (xi, proof["eval_a"], proof["eval_b"], proof["eval_c"], proof["eval_s1"], proof["eval_s2"], proof["eval_zw"]):
        t.add_scalar(v)
    v = [t.get_challenge()]
    for i in range(1, 5):
        v.append(v[i - 1] * v[0] % r)
    ls, xin = lagrange_evaluations(power, npub, xi, w_n, r)
    zh = (xin - 1) % r
    eval_pi = 0
    for val, l in zip(public_inputs, ls):
        eval_pi = (eval_pi - l * val) % r
    ea, eb, ec, ezw, es1, es2 = (proof[k] for k in ("eval_a", "eval_b", "eval_c", "eval_zw", "eval_s1", "eval_s2"))
    betaxi = beta * xi % r
    e2 = (ea + betaxi + gamma) * (eb + betaxi * k1 + gamma) % r * (ec + betaxi * k2 + gamma) % r * alpha % r
    e3 = (ea + beta * es1 + gamma) * (eb + beta * es2 + gamma) % r * ezw % r * alpha % r
    e4 = alpha2 * ls[0] % r
    e24 = (e2 + e4) % r
    ln = n + 6
    pr = [0] * ln
    for i in range(n):
        pr[i] = (z["qm"]["coeffs"][i] * (ea * eb % r) + z["ql"]["coeffs"][i] * ea + z["qr"]["coeffs"][i] * eb
                 + z["qo"]["coeffs"][i] * ec + z["qc"]["coeffs"][i] - z["s3"]["coeffs"][i] * (e3 * beta % r)) % r
    for i, c in enumerate(poly["z"]):
        pr[i] = (pr[i] + c * e24) % r
    xin2 = xin * xin % r
    tmp = [0] * ln
    for i, c in enumerate(t3):
        tmp[i] = c * xin2 % r
    for i, c in enumerate(t2):
        tmp[i] = (tmp[i] + c * xin) % r
    for i, c in enumerate(t1):
        tmp[i] = (tmp[i] + c) % r
    for i in range(ln):
        pr[i] = (pr[i] - tmp[i] * zh) % r
    r0 = (eval_pi - e3 * (ec + gamma) - e4) % r
    pr[0] = (pr[0] + r0) % r
    res = list(pr)
    for k, vv in (("a", v[0]), ("b", v[1]), ("c", v[2])):
        for i, c in enumerate(poly[k]):
            res[i] = (res[i] + c * vv) % r
    for i in range(n):
        res[i] = (res[i] + v[3] * z["s1"]["coeffs"][i] + v[4] * z["s2"]["coeffs"][i]) % r
    res[0] = (res[0] - v[0] * ea - v[1] * eb - v[2] * ec - v[3] * es1 - v[4] * es2) % r
    wxi = _div_by_zerofier1(res, xi, r)
    pz2 = list(poly["z"])
    pz2[0] = (pz2[0] - ezw) % r
    wxiw = _div_by_zerofier1(pz2, xiw, r)
    proof["wxi"], proof["wxiw"] = commit(wxi), commit(wxiw)
    if trace is not None:
        trace.update(beta=beta, gamma=gamma, alpha=alpha, xi=xi, v=v, buffers=buf, polys=poly, evals=ev,
                     t1=t1, t2=t2, t3=t3, r=pr, wxi=wxi, wxiw=wxiw, additions=additions)
    return proof


# ---------------------------------------------------------------------------------------------
def verifier_challenges(curve, vk, proof, public_inputs):
    """plonk.rs:28-103."""
    r = curve.r
    t = Transcript(curve)
    for k in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3"):
        t.add_point(vk[k])
    for p in public_inputs:
        t.add_scalar(int(p) % r)
    for k in ("a", "b", "c"):
        t.add_point(proof[k])
    beta = t.get_challenge()
    t = Transcript(curve)
    t.add_scalar(beta)
    gamma = t.get_challenge()
    t = Transcript(curve)
    t.add_scalar(beta)
    t.add_scalar(gamma)
    t.add_point(proof["z"])
    alpha = t.get_challenge()
    t = Transcript(curve)
    t.add_scalar(alpha)
    for k in ("t1", "t2", "t3"):
        t.add_point(proof[k])
    xi = t.get_challenge()
    t = Transcript(curve)
    for k in (None, "eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"):
        t.add_scalar(xi if k is None else proof[k])
    v = [t.get_challenge()]
    for i in range(1, 5):
        v.append(v[i - 1] * v[0] % r)
    t = Transcript(curve)
    t.add_point(proof["wxi"])
    t.add_point(proof["wxiw"])
    u = t.get_challenge()
    return dict(alpha=alpha, beta=beta, gamma=gamma, xi=xi, v=v, u=u)


def verify(curve, vk, proof, public_inputs, pairing_product_is_one):
    """Plonk::verify (plonk.rs:110-245).  `pairing_product_is_one` = the curve's pairing check
    (oracle.pairing_bn254.pairing_product_is_one for BN254)."""
    r = curve.r
    if vk["n_public"] != len(public_inputs):
        raise ValueError("Invalid number of public inputs")
    public_inputs = [int(x) % r for x in public_inputs]
    ch = verifier_challenges(curve, vk, proof, public_inputs)
    alpha, beta, gamma, xi, v, u = (ch[k] for k in ("alpha", "beta", "gamma", "xi", "v", "u"))
    w_n, _, _ = domains(curve, 1 << vk["power"])
    ls, xin = lagrange_evaluations(vk["power"], vk["n_public"], xi, w_n, r)
    pi = 0
    for val, l in zip(public_inputs, ls):
        pi = (pi - l * val) % r
    ea, eb, ec, ezw, es1, es2 = (proof[k] for k in ("eval_a", "eval_b", "eval_c", "eval_zw", "eval_s1", "eval_s2"))
    G1 = _g1(curve)
    J = G1.to_jac
    mul = lambda P, k: G1.jmul(J(P), k % r)
    add = G1.jadd
    neg = lambda Jp: (Jp[0], (-Jp[1]) % curve.q, Jp[2])
    e2 = alpha * alpha % r * ls[0] % r
    e3a = (ea + es1 * beta + gamma) % r
    e3b = (eb + es2 * beta + gamma) % r
    e3c = (ec + gamma) % r
    e3 = e3a * e3b % r * e3c % r * ezw % r * alpha % r
    r0 = (pi - e2 - e3) % r
    d1 = add(add(add(add(mul(vk["qm"], ea * eb), mul(vk["ql"], ea)), mul(vk["qr"], eb)), mul(vk["qo"], ec)), J(vk["qc"]))
    betaxi = beta * xi % r
    d2a = (ea + betaxi + gamma) * (eb + betaxi * vk["k1"] + gamma) % r * (ec + betaxi * vk["k2"] + gamma) % r * alpha % r
    d2 = mul(proof["z"], d2a + e2 + u)
    d3 = mul(vk["s3"], e3a * e3b % r * (alpha * beta % r * ezw % r))
    d4 = add(add(J(proof["t1"]), mul(proof["t2"], xin)), mul(proof["t3"], xin * xin))
    d4 = G1.jmul(d4, (xin - 1) % r)
    d = add(add(d1, d2), add(neg(d3), neg(d4)))
    e = mul(curve.g1, v[0] * ea + v[1] * eb + v[2] * ec + v[3] * es1 + v[4] * es2 + u * ezw - r0)
    f = add(add(add(add(add(d, mul(proof["a"], v[0])), mul(proof["b"], v[1])), mul(proof["c"], v[2])),
                mul(vk["s1"], v[3])), mul(vk["s2"], v[4]))
    s = u * xi % r * w_n % r
    a1 = add(J(proof["wxi"]), mul(proof["wxiw"], u))
    b1 = add(add(add(mul(proof["wxi"], xi), mul(proof["wxiw"], s)), neg(e)), f)
    A1, B1 = G1.to_affine(a1), G1.to_affine(b1)
    # e(A1, X_2) == e(B1, G2)   <=>   e(-A1, X_2) * e(B1, G2) == 1
    return pairing_product_is_one([(G1.neg(A1), vk["x2"]), (B1, curve.g2)])
