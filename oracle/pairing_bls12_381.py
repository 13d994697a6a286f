"""BLS12-381 optimal-ate pairing product check (oracle; test infrastructure only).

Same role as pairing_bn254.py for the reference's BLS12-381 fixtures (test_vectors/{Groth16,Plonk}/bls12_381):
the acceptance criterion of co-groth16/src/lib.rs:93-160 and co-plonk/src/plonk.rs:110-245 is a pairing equation.

Fq12 = Fq2[w]/(w^6 - xi), xi = 1 + u.  The twist E': y^2 = x^3 + 4 xi is of M type: (x', y') in E'(Fq2) maps to
(x' / w^2, y' / w^3) in E(Fq12).  Lines are scaled by w^3 (and a sign), factors that the final exponentiation
removes because their squares lie in Fq2.  The loop runs over |x| = 0xd201000000010000; x < 0 only inverts the
result, which a "product == 1" check does not see.
"""
from .fields import BLS12_381
from .ec import Fq2Ops, g1 as _g1, g2 as _g2

Q = BLS12_381.q
R = BLS12_381.r
F2 = Fq2Ops(Q)
ATE_LOOP_COUNT = 0xd201000000010000
LOG_ATE = 62  # bits below the leading one

ZERO2 = (0, 0)
ONE12 = ((1, 0),) + (ZERO2,) * 5


def _mul_xi(a):
    # (a0 + a1 u)(1 + u) = a0 - a1 + (a0 + a1) u
    return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)


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
    """Line through the untwisted R1, R2 (tangent if equal) at P in G1, times -w^3:
    (y1 - m x1) + (m xp) w^2 - yp w^3 with m the slope on the twist.  Returns (line, R1 + R2 on the twist)."""
    x1, y1 = R1
    x2, y2 = R2
    xp, yp = P
    if x1 != x2:
        m = F2.mul(F2.sub(y2, y1), F2.inv(F2.sub(x2, x1)))
    elif y1 == y2:
        m = F2.mul(F2.small(3, F2.sqr(x1)), F2.inv(F2.small(2, y1)))
    else:
        # vertical line xp - x1 / w^2, times w^2: -x1 + xp w^2
        return (F2.neg(x1), ZERO2, (xp % Q, 0), ZERO2, ZERO2, ZERO2), None
    x3 = F2.sub(F2.sub(F2.sqr(m), x1), x2)
    y3 = F2.sub(F2.mul(m, F2.sub(x1, x3)), y1)
    line = (F2.sub(y1, F2.mul(m, x1)), ZERO2, F2.small(xp, m), ((-yp) % Q, 0), ZERO2, ZERO2)
    return line, (x3, y3)


def miller_loop(Qt, P):
    """Qt affine in G2 (Fq2 coordinates on the twist), P affine in G1; None -> identity."""
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
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs [(P in G1, Q in G2)]."""
    f = ONE12
    for P, Qt in pairs:
        f = f12_mul(f, miller_loop(Qt, P))
    return f12_pow(f, FINAL_EXP) == ONE12


def groth16_verify(vk, public_inputs, proof):
    """Same contract as pairing_bn254.groth16_verify."""
    G1, G2 = _g1(BLS12_381), _g2(BLS12_381)
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
