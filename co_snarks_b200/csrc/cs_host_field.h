// Host-side field / curve arithmetic (64-bit limbs, unsigned __int128) for the handful of
// single-element operations the prover does outside the GPU kernels: finishing an MSM (XYZZ ->
// affine), `scalar_mul_public_point_hs`, `add_assign_points_public_hs`, the tiny public-input MSM
// (`query[1..=pub]`, co-groth16/src/groth16.rs:194), the final sums of
// `create_proof_with_assignment` (groth16.rs:297-337) and the root-of-unity derivation
// (groth16.rs:60-100).  This is product code, not the oracle, and it is not a fallback for any
// kernel: nothing vector-sized runs here.  Same Montgomery representation and memory layout as the
// device types (little-endian limbs), so values move with memcpy.
#pragma once
#include <stdint.h>
#include <string.h>

namespace cs { namespace host {

typedef unsigned __int128 u128;

template <class P>
struct HFp {
  static constexpr int N = P::N / 2;  // 64-bit limbs
  uint64_t l[N];

  static uint64_t modl(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
  static uint64_t m0inv() {  // -p^-1 mod 2^64 from the 32-bit constant (one Newton step)
    uint64_t p0 = modl(0);
    uint64_t x = (uint64_t)(uint32_t)(0u - P::M0);  // p^-1 mod 2^32
    x *= 2 - p0 * x;                                 // p^-1 mod 2^64
    return 0 - x;
  }
  static HFp zero() { HFp r; memset(r.l, 0, sizeof(r.l)); return r; }
  static HFp one() {
    HFp r;
    for (int i = 0; i < N; i++) r.l[i] = (uint64_t)P::one(2 * i) | ((uint64_t)P::one(2 * i + 1) << 32);
    return r;
  }
  static HFp r2() {
    HFp r;
    for (int i = 0; i < N; i++) r.l[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32);
    return r;
  }
  static HFp from_u64(uint64_t v) { HFp r = zero(); r.l[0] = v; return r.to_mont(); }

  bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
  bool operator==(const HFp& b) const { return memcmp(l, b.l, sizeof(l)) == 0; }
  bool operator!=(const HFp& b) const { return !(*this == b); }

  static bool geq_mod(const uint64_t* a) {
    for (int i = N - 1; i >= 0; i--) {
      uint64_t m = modl(i);
      if (a[i] != m) return a[i] > m;
    }
    return true;
  }
  static void sub_mod(uint64_t* a) {
    u128 br = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a[i] - modl(i) - br;
      a[i] = (uint64_t)d;
      br = (d >> 64) & 1;
    }
  }
  friend HFp operator+(const HFp& a, const HFp& b) {
    HFp r;
    u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  friend HFp operator-(const HFp& a, const HFp& b) {
    HFp r;
    u128 br = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a.l[i] - b.l[i] - br;
      r.l[i] = (uint64_t)d;
      br = (d >> 64) & 1;
    }
    if (br) {
      u128 c = 0;
      for (int i = 0; i < N; i++) { c += (u128)r.l[i] + modl(i); r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  HFp neg() const { return is_zero() ? *this : zero() - *this; }
  HFp dbl() const { return *this + *this; }
  friend HFp operator*(const HFp& a, const HFp& b) {
    // CIOS
    uint64_t t[N + 2];
    memset(t, 0, sizeof(t));
    const uint64_t inv = m0inv();
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) {
        c += (u128)a.l[j] * b.l[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N] = (uint64_t)c;
      t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * inv;
      c = (u128)m * modl(0) + t[0];
      c >>= 64;
      for (int j = 1; j < N; j++) {
        c += (u128)m * modl(j) + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N - 1] = (uint64_t)c;
      t[N] = t[N + 1] + (uint64_t)(c >> 64);
    }
    HFp r;
    memcpy(r.l, t, sizeof(r.l));
    if (t[N] || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  HFp sqr() const { return *this * *this; }
  HFp to_mont() const { return *this * r2(); }
  HFp from_mont() const { HFp o = zero(); o.l[0] = 1; return *this * o; }
  // this^e for a little-endian multi-limb exponent
  HFp pow(const uint64_t* e, int nlimbs) const {
    HFp res = one(), base = *this;
    for (int i = 0; i < nlimbs; i++)
      for (int b = 0; b < 64; b++) {
        if ((e[i] >> b) & 1) res = res * base;
        base = base.sqr();
      }
    return res;
  }
  HFp inverse() const {
    uint64_t e[N];
    for (int i = 0; i < N; i++) e[i] = modl(i);
    e[0] -= 2;  // moduli are odd and > 2
    return pow(e, N);
  }
};

template <class P>
struct HFp2 {
  typedef HFp<P> F;
  F c0, c1;
  static HFp2 zero() { return HFp2{F::zero(), F::zero()}; }
  static HFp2 one() { return HFp2{F::one(), F::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const HFp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  bool operator!=(const HFp2& b) const { return !(*this == b); }
  friend HFp2 operator+(const HFp2& a, const HFp2& b) { return HFp2{a.c0 + b.c0, a.c1 + b.c1}; }
  friend HFp2 operator-(const HFp2& a, const HFp2& b) { return HFp2{a.c0 - b.c0, a.c1 - b.c1}; }
  friend HFp2 operator*(const HFp2& a, const HFp2& b) {
    F v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;
    return HFp2{v0 - v1, (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1};
  }
  HFp2 sqr() const { return *this * *this; }
  HFp2 neg() const { return HFp2{c0.neg(), c1.neg()}; }
  HFp2 dbl() const { return HFp2{c0.dbl(), c1.dbl()}; }
  HFp2 inverse() const {
    F n = (c0.sqr() + c1.sqr()).inverse();
    return HFp2{c0 * n, (c1 * n).neg()};
  }
};

// Points: affine with (0,0) = infinity, XYZZ accumulators (same conventions as cs_curve.cuh).
template <class F>
struct HAffine {
  F x, y;
  bool is_inf() const { return x.is_zero() && y.is_zero(); }
  static HAffine inf() { return HAffine{F::zero(), F::zero()}; }
};

template <class F>
struct HXyzz {
  F x, y, zz, zzz;
  bool is_inf() const { return zz.is_zero(); }
  static HXyzz inf() { return HXyzz{F::zero(), F::zero(), F::zero(), F::zero()}; }
  static HXyzz from_affine(const HAffine<F>& p) {
    if (p.is_inf()) return inf();
    return HXyzz{p.x, p.y, F::one(), F::one()};
  }
};

template <class F>
HXyzz<F> hdbl(const HXyzz<F>& p) {
  if (p.is_inf() || p.y.is_zero()) return HXyzz<F>::inf();
  F U = p.y.dbl(), V = U.sqr(), W = U * V, S = p.x * V, X2 = p.x.sqr(), M = X2.dbl() + X2;
  HXyzz<F> r;
  r.x = M.sqr() - S.dbl();
  r.y = M * (S - r.x) - W * p.y;
  r.zz = V * p.zz;
  r.zzz = W * p.zzz;
  return r;
}

template <class F>
HXyzz<F> hadd(const HXyzz<F>& a, const HXyzz<F>& q) {
  if (q.is_inf()) return a;
  if (a.is_inf()) return q;
  F U1 = a.x * q.zz, U2 = q.x * a.zz, S1 = a.y * q.zzz, S2 = q.y * a.zzz;
  F Pp = U2 - U1, R = S2 - S1;
  if (Pp.is_zero()) return R.is_zero() ? hdbl(a) : HXyzz<F>::inf();
  F PP = Pp.sqr(), PPP = Pp * PP, Q = U1 * PP;
  HXyzz<F> r;
  r.x = R.sqr() - PPP - Q.dbl();
  r.y = R * (Q - r.x) - S1 * PPP;
  r.zz = a.zz * q.zz * PP;
  r.zzz = a.zzz * q.zzz * PPP;
  return r;
}

template <class F>
HXyzz<F> hneg(const HXyzz<F>& a) {
  HXyzz<F> r = a;
  r.y = r.y.neg();
  return r;
}

// k * p for a canonical little-endian scalar of `nlimbs` 64-bit limbs
template <class F>
HXyzz<F> hmul(const HXyzz<F>& p, const uint64_t* k, int nlimbs) {
  HXyzz<F> r = HXyzz<F>::inf();
  for (int i = nlimbs - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      r = hdbl(r);
      if ((k[i] >> b) & 1) r = hadd(r, p);
    }
  return r;
}

template <class F>
HAffine<F> haffine(const HXyzz<F>& p) {
  if (p.is_inf()) return HAffine<F>::inf();
  F zi = p.zzz.inverse();
  F zzi = (zi * p.zz).sqr();
  return HAffine<F>{p.x * zzi, p.y * zi};
}

}}  // namespace cs::host
