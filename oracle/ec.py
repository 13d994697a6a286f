"""Short-Weierstrass group arithmetic for G1 (over Fq) and G2 (over Fq2 = Fq[u]/(u^2+1)).

Oracle / test infrastructure only.  Affine points are `(x, y)` tuples or `None` (infinity); G2
coordinates are `(c0, c1)` tuples.  The MSM result is mathematically unique, so any exact
restatement equals `taceo_ark_algebra::msm::msm_unchecked` (call sites:
co-groth16/src/mpc/rep3.rs:124-132, co-groth16/src/groth16.rs:194) after `into_affine`.
"""


class Fq1Ops:
    """Base-field ops on python ints."""

    def __init__(self, q):
        self.q = q
        self.zero, self.one = 0, 1

    def add(self, a, b): return (a + b) % self.q
    def sub(self, a, b): return (a - b) % self.q
    def mul(self, a, b): return a * b % self.q
    def sqr(self, a): return a * a % self.q
    def neg(self, a): return (-a) % self.q
    def inv(self, a): return pow(a, self.q - 2, self.q)
    def is_zero(self, a): return a % self.q == 0
    def small(self, k, a): return k * a % self.q


class Fq2Ops:
    """Fq2 = Fq[u]/(u^2+1) on (c0, c1) tuples (BN254 and BLS12-381 both use u^2 = -1)."""

    def __init__(self, q):
        self.q = q
        self.zero, self.one = (0, 0), (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.q, (a[1] + b[1]) % self.q)
    def sub(self, a, b): return ((a[0] - b[0]) % self.q, (a[1] - b[1]) % self.q)

    def mul(self, a, b):
        q = self.q
        return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def sqr(self, a): return self.mul(a, a)
    def neg(self, a): return ((-a[0]) % self.q, (-a[1]) % self.q)

    def inv(self, a):
        q = self.q
        n = pow(a[0] * a[0] + a[1] * a[1], q - 2, q)
        return (a[0] * n % q, (-a[1]) * n % q)

    def is_zero(self, a): return a[0] % self.q == 0 and a[1] % self.q == 0
    def small(self, k, a): return (k * a[0] % self.q, k * a[1] % self.q)


class Group:
    """y^2 = x^3 + b with a = 0.  Jacobian internally."""

    def __init__(self, F, b):
        self.F, self.b = F, b

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.is_zero(F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    # --- Jacobian (X, Y, Z); infinity is Z == 0
    def to_jac(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jdbl(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.small(2, F.sub(F.sub(F.sqr(F.add(X, B)), A), C))
        E = F.small(3, A)
        X3 = F.sub(F.sqr(E), F.small(2, D))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.small(8, C))
        Z3 = F.small(2, F.mul(Y, Z))
        return (X3, Y3, Z3)

    def jadd(self, P, Q):
        F = self.F
        if F.is_zero(P[2]):
            return Q
        if F.is_zero(Q[2]):
            return P
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1, Z2Z2 = F.sqr(Z1), F.sqr(Z2)
        U1, U2 = F.mul(X1, Z2Z2), F.mul(X2, Z1Z1)
        S1, S2 = F.mul(Y1, F.mul(Z2, Z2Z2)), F.mul(Y2, F.mul(Z1, Z1Z1))
        if U1 == U2:
            if S1 == S2:
                return self.jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.small(2, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def jmul(self, J, k):
        F = self.F
        R = (F.one, F.one, F.zero)
        if k == 0:
            return R
        for bit in bin(k)[2:]:
            R = self.jdbl(R)
            if bit == "1":
                R = self.jadd(R, J)
        return R

    def mul(self, P, k):
        return self.to_affine(self.jmul(self.to_jac(P), k))

    def msm(self, points, scalars, c=None):
        """Sum_i scalars[i] * points[i] over min(len) pairs (msm_unchecked chops to the shorter
        slice, co-noir-common/src/honk_curve.rs:33-34).  Pippenger with unsigned c-bit windows."""
        n = min(len(points), len(scalars))
        F = self.F
        inf = (F.one, F.one, F.zero)
        if n == 0:
            return None
        if n < 8:
            acc = inf
            for P, s in zip(points[:n], scalars[:n]):
                if P is not None and s:
                    acc = self.jadd(acc, self.jmul(self.to_jac(P), s))
            return self.to_affine(acc)
        if c is None:
            c = max(2, min(16, n.bit_length() - 2))
        nbits = max(int(s).bit_length() for s in scalars[:n]) or 1
        nwin = (nbits + c - 1) // c
        jac = [None if P is None else self.to_jac(P) for P in points[:n]]
        total = inf
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jdbl(total)
            buckets = {}
            sh = w * c
            mask = (1 << c) - 1
            for J, s in zip(jac, scalars[:n]):
                d = (int(s) >> sh) & mask
                if d and J is not None:
                    b = buckets.get(d)
                    buckets[d] = J if b is None else self.jadd(b, J)
            run, acc, prev = inf, inf, None
            for d in sorted(buckets, reverse=True):
                if prev is not None:
                    # acc += run * (prev - d)
                    gap = prev - d
                    acc = self.jadd(acc, self.jmul(run, gap)) if gap > 1 else self.jadd(acc, run)
                run = self.jadd(run, buckets[d])
                prev = d
            if prev is not None:
                acc = self.jadd(acc, self.jmul(run, prev)) if prev > 1 else self.jadd(acc, run)
            total = self.jadd(total, acc)
        return self.to_affine(total)


def g1(curve):
    return Group(Fq1Ops(curve.q), curve.b)


def g2(curve):
    return Group(Fq2Ops(curve.q), curve.b2)
