"""BLS12-377 optimal-ate pairing product check and the arkworks containers of the reference's BLS12-377 fixtures
(oracle; test infrastructure only).

Role: pin the oracle's LibSnarkReduction restatement (co-groth16/src/groth16/reduction.rs:241-342) on the only
fixtures the reference holds for it, the Penumbra keys under test_vectors/Groth16/bls12_377 -- the acceptance
criterion of `proof_libsnark_penumbra_bls12_377` (co-groth16/src/lib.rs:231-285) is that the proof verifies under
circuit.vk, a pairing equation over BLS12-377.

Curve: E / Fq : y^2 = x^3 + 1, x = 0x8508C00000000001, r = x^4 - x^2 + 1, q = (x - 1)^2 r / 3 + x.
Fq2 = Fq[u] / (u^2 + 5); Fq12 = Fq2[w] / (w^6 - u).  The twist E' : y^2 = x^3 + 1/u is of D type: (x', y') in E'(Fq2)
maps to (x' w^2, y' w^3) in E(Fq12).  The loop count x is positive.
"""
X_PARAM = 0x8508C00000000001
R = X_PARAM ** 4 - X_PARAM ** 2 + 1
Q = ((X_PARAM - 1) ** 2 * R) // 3 + X_PARAM
NONRES = -5  # u^2
ZERO2, ONE2 = (0, 0), (1, 0)
ONE12 = (ONE2,) + (ZERO2,) * 5
B_G1 = 1
B_G2 = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)  # 1 / u


class _F2:
    zero, one = ZERO2, ONE2
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
    @staticmethod
    def mul(a, b): return ((a[0] * b[0] + NONRES * a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
    @staticmethod
    def sqr(a): return _F2.mul(a, a)
    @staticmethod
    def neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
    @staticmethod
    def small(k, a): return (k * a[0] % Q, k * a[1] % Q)
    @staticmethod
    def inv(a):
        n = pow((a[0] * a[0] - NONRES * a[1] * a[1]) % Q, Q - 2, Q)
        return (a[0] * n % Q, (-a[1]) * n % Q)
    @staticmethod
    def is_zero(a): return a[0] % Q == 0 and a[1] % Q == 0


F2 = _F2


class _F1:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % Q)
    sub = staticmethod(lambda a, b: (a - b) % Q)
    mul = staticmethod(lambda a, b: a * b % Q)
    sqr = staticmethod(lambda a: a * a % Q)
    neg = staticmethod(lambda a: (-a) % Q)
    small = staticmethod(lambda k, a: k * a % Q)
    inv = staticmethod(lambda a: pow(a, Q - 2, Q))
    is_zero = staticmethod(lambda a: a % Q == 0)


def _group(F, b):
    from .ec import Group
    return Group(F, b)


def g1():
    return _group(_F1, B_G1)


def g2():
    return _group(_F2, B_G2)


def _mul_u(a):  # (a0 + a1 u) u = -5 a1 + a0 u
    return (NONRES * a[1] % Q, a[0] % Q)


def f12_mul(a, b):
    t = [ZERO2] * 11
    for i in range(6):
        ai = a[i]
        if ai == ZERO2:
            continue
        for j in range(6):
            bj = b[j]
            if bj == ZERO2:
                continue
            t[i + j] = F2.add(t[i + j], F2.mul(ai, bj))
    out = list(t[:6])
    for k in range(6, 11):
        if t[k] != ZERO2:
            out[k - 6] = F2.add(out[k - 6], _mul_u(t[k]))  # w^6 = u
    return tuple(out)


def f12_pow(a, e):
    res = ONE12
    for bit in bin(e)[2:]:
        res = f12_mul(res, res)
        if bit == "1":
            res = f12_mul(res, a)
    return res


def _line(R1, R2, P):
    """Line through the untwisted R1, R2 (tangent if equal) at P in G1:
    l(P) = y_P - m x_P w + (m x1 - y1) w^3 with m the slope on the twist.  -> (line, R1 + R2 on the twist)."""
    x1, y1 = R1
    x2, y2 = R2
    xp, yp = P
    if x1 != x2:
        m = F2.mul(F2.sub(y2, y1), F2.inv(F2.sub(x2, x1)))
    elif y1 == y2:
        m = F2.mul(F2.small(3, F2.sqr(x1)), F2.inv(F2.small(2, y1)))
    else:  # vertical line x_P - x1 w^2
        return ((xp % Q, 0), ZERO2, F2.neg(x1), ZERO2, ZERO2, ZERO2), None
    x3 = F2.sub(F2.sub(F2.sqr(m), x1), x2)
    y3 = F2.sub(F2.mul(m, F2.sub(x1, x3)), y1)
    line = ((yp % Q, 0), F2.neg(F2.small(xp, m)), ZERO2, F2.sub(F2.mul(m, x1), y1), ZERO2, ZERO2)
    return line, (x3, y3)


def miller_loop(Qt, P):
    if Qt is None or P is None:
        return ONE12
    Rp, f = Qt, ONE12
    for i in range(X_PARAM.bit_length() - 2, -1, -1):
        ln, R2 = _line(Rp, Rp, P)
        f = f12_mul(f12_mul(f, f), ln)
        Rp = R2
        if X_PARAM & (1 << i):
            ln, R2 = _line(Rp, Qt, P)
            f = f12_mul(f, ln)
            Rp = R2
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def pairing_product_is_one(pairs):
    f = ONE12
    for P, Qt in pairs:
        f = f12_mul(f, miller_loop(Qt, P))
    return f12_pow(f, FINAL_EXP) == ONE12


def groth16_verify(vk, public_inputs, proof):
    """ark_groth16::Groth16::verify: e(A, B) = e(alpha, beta) e(sum_i x_i ic_i, gamma) e(C, delta)."""
    G1, G2 = g1(), g2()
    A, B, C = proof
    if A is None or C is None or B is None or not G1.on_curve(A) or not G1.on_curve(C) or not G2.on_curve(B):
        return False
    acc = G1.to_jac(vk["ic"][0])
    for s, P in zip(public_inputs, vk["ic"][1:]):
        acc = G1.jadd(acc, G1.jmul(G1.to_jac(P), int(s) % R))
    L = G1.to_affine(acc)
    return pairing_product_is_one([(G1.neg(A), B), (vk["alpha_g1"], vk["beta_g2"]), (L, vk["gamma_g2"]), (C, vk["delta_g2"])])


# ---- arkworks CanonicalSerialize, uncompressed (ark-serialize 0.5): field elements little-endian, an affine point is
# x || y with the two flag bits (bit 7: y is the larger root, bit 6: infinity) in the top bits of y's last byte
def _fq(b):
    return int.from_bytes(b, "little")


def read_g1(buf, off):
    x = _fq(buf[off:off + 48])
    yb = bytearray(buf[off + 48:off + 96])
    inf = bool(yb[47] & 0x40)
    yb[47] &= 0x3f
    return (None if inf else (x, _fq(yb))), off + 96


def read_g2(buf, off):
    x = (_fq(buf[off:off + 48]), _fq(buf[off + 48:off + 96]))
    yb = bytearray(buf[off + 96:off + 192])
    inf = bool(yb[95] & 0x40)
    yb[95] &= 0x3f
    return (None if inf else (x, (_fq(yb[:48]), _fq(yb[48:])))), off + 192


def _read_vec(buf, off, rd):
    n = int.from_bytes(buf[off:off + 8], "little")
    off += 8
    out = []
    for _ in range(n):
        v, off = rd(buf, off)
        out.append(v)
    return out, off


def read_vk(buf, off=0):
    vk = {}
    vk["alpha_g1"], off = read_g1(buf, off)
    vk["beta_g2"], off = read_g2(buf, off)
    vk["gamma_g2"], off = read_g2(buf, off)
    vk["delta_g2"], off = read_g2(buf, off)
    vk["ic"], off = _read_vec(buf, off, read_g1)
    return vk, off


def read_pk(buf):
    """ark_groth16::ProvingKey: vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query."""
    pk = {}
    pk["vk"], off = read_vk(buf, 0)
    pk["beta_g1"], off = read_g1(buf, off)
    pk["delta_g1"], off = read_g1(buf, off)
    for name, rd in (("a_query", read_g1), ("b_g1_query", read_g1), ("b_g2_query", read_g2), ("h_query", read_g1), ("l_query", read_g1)):
        pk[name], off = _read_vec(buf, off, rd)
    assert off == len(buf), "trailing bytes in the proving key"
    return pk


def read_matrix(buf):
    """ark_relations Matrix<F> = Vec<Vec<(F, usize)>>: rows of (coefficient, column)."""
    off = 8
    rows = []
    for _ in range(int.from_bytes(buf[:8], "little")):
        k = int.from_bytes(buf[off:off + 8], "little")
        off += 8
        row = []
        for _ in range(k):
            row.append((int.from_bytes(buf[off:off + 32], "little"), int.from_bytes(buf[off + 32:off + 40], "little")))
            off += 40
        rows.append(row)
    assert off == len(buf)
    return rows
