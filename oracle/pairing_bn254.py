"""BN254 optimal-ate pairing (oracle; test infrastructure only).

Used to check Groth16 proofs against the snarkjs verification keys, i.e. the same acceptance
criterion as the reference's tests (co-groth16/src/lib.rs:40-69, verifier.rs:17-30 which defer to
ark-groth16's pairing check):  e(A,B) = e(alpha,beta) * e(sum pub_i IC_i, gamma) * e(C, delta).

Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u; elements are 6-tuples of Fq2 tuples.  The twist maps
(x', y') in E'(Fq2) to (x' w^2, y' w^3) in E(Fq12).
"""
from .fields import BN254
from .ec import Fq2Ops, g1 as _g1, g2 as _g2

Q = BN254.q
R = BN254.r
F2 = Fq2Ops(Q)
XI = (9, 1)
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE = 63

ZERO2 = (0, 0)
ONE12 = ((1, 0),) + (ZERO2,) * 5


def _mul_xi(a):
    # (a0 + a1 u)(9 + u) = 9a0 - a1 + (a0 + 9 a1) u
    return ((9 * a[0] - a[1]) % Q, (a[0] + 9 * a[1]) % Q)


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
            out[k - 6] = F2.add(out[k - 6], _mul_xi(t[k]))
    return tuple(out)


def f12_pow(a, e):
    res = ONE12
    for bit in bin(e)[2:]:
        res = f12_mul(res, res)
        if bit == "1":
            res = f12_mul(res, a)
    return res


def _line(R1, R2, P):
    """Line through psi(R1), psi(R2) (tangent if equal) evaluated at P in G1; returns (line, R1+R2).
    Sparse Fq12: -yp + (m xp) w + (y1 - m x1) w^3."""
    x1, y1 = R1
    x2, y2 = R2
    xp, yp = P
    if x1 != x2:
        m = F2.mul(F2.sub(y2, y1), F2.inv(F2.sub(x2, x1)))
    elif y1 == y2:
        m = F2.mul(F2.small(3, F2.sqr(x1)), F2.inv(F2.small(2, y1)))
    else:
        # vertical line: xp - x1 w^2
        return ((xp % Q, 0), ZERO2, F2.neg(x1), ZERO2, ZERO2, ZERO2), None
    x3 = F2.sub(F2.sub(F2.sqr(m), x1), x2)
    y3 = F2.sub(F2.mul(m, F2.sub(x1, x3)), y1)
    line = (((-yp) % Q, 0), F2.small(xp, m), ZERO2, F2.sub(y1, F2.mul(m, x1)), ZERO2, ZERO2)
    return line, (x3, y3)


def _frob2(a):
    """Fq2 Frobenius = conjugation."""
    return (a[0], (-a[1]) % Q)


# psi^-1 o Frobenius o psi on the twist: (x, y) -> (conj(x) * xi^((q-1)/3), conj(y) * xi^((q-1)/2))
def _f2_pow(a, e):
    res = (1, 0)
    for bit in bin(e)[2:]:
        res = F2.sqr(res)
        if bit == "1":
            res = F2.mul(res, a)
    return res


_G2X = _f2_pow(XI, (Q - 1) // 3)
_G2Y = _f2_pow(XI, (Q - 1) // 2)


def _twist_frob(Pt):
    return (F2.mul(_frob2(Pt[0]), _G2X), F2.mul(_frob2(Pt[1]), _G2Y))


def miller_loop(Qt, P):
    """Qt affine in G2 (Fq2 coords), P affine in G1; None -> identity."""
    if Qt is None or P is None:
        return ONE12
    Rp = Qt
    f = ONE12
    for i in range(LOG_ATE, -1, -1):
        ln, R2 = _line(Rp, Rp, P)
        f = f12_mul(f12_mul(f, f), ln)
        Rp = R2
        if ATE_LOOP_COUNT & (1 << i):
            ln, R2 = _line(Rp, Qt, P)
            f = f12_mul(f, ln)
            Rp = R2
    Q1 = _twist_frob(Qt)
    Q2 = _twist_frob(Q1)
    nQ2 = (Q2[0], F2.neg(Q2[1]))
    ln, Rp = _line(Rp, Q1, P)
    f = f12_mul(f, ln)
    ln, _ = _line(Rp, nQ2, P)
    f = f12_mul(f, ln)
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs [(P in G1, Q in G2)]."""
    f = ONE12
    for P, Qt in pairs:
        f = f12_mul(f, miller_loop(Qt, P))
    return f12_pow(f, FINAL_EXP) == ONE12


def groth16_verify(vk, public_inputs, proof):
    """vk: dict(alpha_g1, beta_g2, gamma_g2, delta_g2, ic[list]); public_inputs excludes the leading 1;
    proof: (A in G1, B in G2, C in G1)."""
    G1, G2 = _g1(BN254), _g2(BN254)
    A, B, C = proof
    for P in (A, C):
        if P is None or not G1.on_curve(P):
            return False
    if B is None or not G2.on_curve(B):
        return False
    acc = G1.to_jac(vk["ic"][0])
    for s, P in zip(public_inputs, vk["ic"][1:]):
        acc = G1.jadd(acc, G1.jmul(G1.to_jac(P), int(s) % R))
    L = G1.to_affine(acc)
    return pairing_product_is_one([
        (G1.neg(A), B), (vk["alpha_g1"], vk["beta_g2"]), (L, vk["gamma_g2"]), (C, vk["delta_g2"])])
