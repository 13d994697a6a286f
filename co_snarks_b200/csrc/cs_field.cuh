// Prime-field arithmetic in Montgomery form on 32-bit limbs (device side).
//
// Replaces the arkworks `Fp<MontBackend<_, N>>` arithmetic that every reference function on the
// hot path bottoms out in (field elements are `[u64; N]` little-endian Montgomery limbs with
// R = 2^(64 N); a `[u64; N]` is bit-identical to our `uint32_t[2 N]`, so buffers cross the C ABI
// without conversion).  SURVEY.md 8(a): "F_r element = 32 B ... Montgomery form, R = 2^256".
//
// mul(): word-serial Montgomery product with the even/odd accumulator split, so every
// 32x32->64 partial product is one IMAD.WIDE in a single carry chain (no carry-save fix-ups).
// All results are fully reduced to [0, p), which the exact equality tests in the point formulas need.
//
// Attribution: the even/odd-accumulator word-serial product below (mul_n, cmad_n, madc_n_rshift,
// mad_n_redc and the way mul_inline drives them) follows the structure and helper naming of
// Supranational's sppark, ff/mont_t.cuh (Copyright Supranational LLC, Apache License 2.0,
// https://github.com/supranational/sppark); see NOTICE at the repository root.  The product-scanning
// squaring, the conversions and everything else in this file are this repository's own.
#pragma once
#include "cs_prims.cuh"

namespace cs {

// ---- generic N-limb helpers ---------------------------------------------------------------------
template <int N>
CS_D void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  CS_UNROLL
  for (int j = 0; j < N; j += 2) {
    acc[j] = mul_lo(a[j], bi);
    acc[j + 1] = mul_hi(a[j], bi);
  }
}

// acc[0..N) += (a[0], a[2], ...) * bi   (pairs (lo,hi) land on (j, j+1)); carry-out stays in CC.
template <int N>
CS_D void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  acc[0] = mad_lo_cc(a[0], bi, acc[0]);
  acc[1] = madc_hi_cc(a[0], bi, acc[1]);
  CS_UNROLL
  for (int j = 2; j < N; j += 2) {
    acc[j] = madc_lo_cc(a[j], bi, acc[j]);
    acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]);
  }
}

// odd = (odd >> 64) + (a[0], a[2], ...) * bi + CC      (no carry-out possible)
template <int N>
CS_D void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
  CS_UNROLL
  for (int j = 0; j < N - 2; j += 2) {
    odd[j] = madc_lo_cc(a[j], bi, odd[j + 2]);
    odd[j + 1] = madc_hi_cc(a[j], bi, odd[j + 3]);
  }
  odd[N - 2] = madc_lo_cc(a[N - 2], bi, 0);
  odd[N - 1] = madc_hi(a[N - 2], bi, 0);
}

// One row of the interleaved Montgomery product.  X is the accumulator aligned with bit 0 of the
// running value T, Y the one aligned 32 bits up (T = X + Y * 2^32); the roles swap every row.
template <class P, bool FIRST>
CS_D void mad_n_redc(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
  constexpr int N = P::N;
  if (FIRST) {
    mul_n<N>(Y, a + 1, bi);
    mul_n<N>(X, a, bi);
  } else {
    X[0] = add_cc(X[0], Y[1]);
    madc_n_rshift<N>(Y, a + 1, bi);
    cmad_n<N>(X, a, bi);
    Y[N - 1] = addc(Y[N - 1], 0);
  }
  uint32_t mi = mul_lo(X[0], P::M0);
  uint32_t mod[N];
  CS_UNROLL
  for (int i = 0; i < N; i++) mod[i] = P::mod(i);
  cmad_n<N>(Y, mod + 1, mi);
  cmad_n<N>(X, mod, mi);
  Y[N - 1] = addc(Y[N - 1], 0);
}

template <class P>
struct Fp {
  static constexpr int N = P::N;
  uint32_t l[N];

  // ---- constants
  static CS_D Fp zero() { Fp r; CS_UNROLL for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
  static CS_D Fp one() { Fp r; CS_UNROLL for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }
  static CS_D Fp r2() { Fp r; CS_UNROLL for (int i = 0; i < N; i++) r.l[i] = P::r2(i); return r; }

  CS_D bool is_zero() const {
    uint32_t o = 0;
    CS_UNROLL
    for (int i = 0; i < N; i++) o |= l[i];
    return o == 0;
  }
  CS_D bool operator==(const Fp& b) const {
    uint32_t o = 0;
    CS_UNROLL
    for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
    return o == 0;
  }
  CS_D bool operator!=(const Fp& b) const { return !(*this == b); }

  // r = (r >= p) ? r - p : r
  CS_D void final_sub() {
    uint32_t t[N];
    t[0] = sub_cc(l[0], P::mod(0));
    CS_UNROLL
    for (int i = 1; i < N; i++) t[i] = subc_cc(l[i], P::mod(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if l < p
    CS_UNROLL
    for (int i = 0; i < N; i++) l[i] = borrow ? l[i] : t[i];
  }

  friend CS_D Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
    CS_UNROLL
    for (int i = 1; i < N; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    // p has at least one spare top bit (254/255/381-bit moduli), so no carry-out here
    r.final_sub();
    return r;
  }
  friend CS_D Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
    CS_UNROLL
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = subc(0, 0);
    r.l[0] = add_cc(r.l[0], borrow & P::mod(0));
    CS_UNROLL
    for (int i = 1; i < N; i++) r.l[i] = addc_cc(r.l[i], borrow & P::mod(i));
    return r;
  }
  CS_D Fp neg() const { return is_zero() ? *this : (zero() - *this); }
  CS_D Fp dbl() const { return *this + *this; }

  // Out-of-line on purpose: one ~230-instruction copy per kernel keeps the point formulas (10-42
  // multiplications each) inside the instruction cache; operands travel by value in registers.
  friend CS_D Fp operator*(const Fp& a, const Fp& b) { return mul_ool(a, b); }
  static CS_DN Fp mul_ool(Fp a, Fp b) { return mul_inline(a, b); }
  static CS_D Fp mul_inline(const Fp& a, const Fp& b) {
    uint32_t even[N], odd[N];
    mad_n_redc<P, true>(even, odd, a.l, b.l[0]);
    mad_n_redc<P, false>(odd, even, a.l, b.l[1]);
    CS_UNROLL
    for (int i = 2; i < N; i += 2) {
      mad_n_redc<P, false>(even, odd, a.l, b.l[i]);
      mad_n_redc<P, false>(odd, even, a.l, b.l[i + 1]);
    }
    // result = even + (odd >> 32)
    Fp r;
    r.l[0] = add_cc(even[0], odd[1]);
    CS_UNROLL
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = addc(even[N - 1], 0);
    r.final_sub();
    return r;
  }
  // Squaring: product-scanning (Comba) form so that the 28 symmetric cross products a_i a_j (i < j) are
  // computed once and doubled: 36 + 64 (reduction) IMAD.WIDE instead of 128.  Each column keeps a 3-word
  // accumulator; the cross-product sum is doubled in a second 3-word register set before it is merged.
  CS_D Fp sqr() const { return sqr_ool(*this); }
  static CS_DN Fp sqr_ool(Fp a) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    uint32_t m[N];
    Fp r;
    CS_UNROLL
    for (int k = 0; k < 2 * N - 1; k++) {
      // cross terms 2 * sum_{i < j, i + j = k} a_i a_j
      uint32_t t0 = 0, t1 = 0, t2 = 0;
      CS_UNROLL
      for (int i = 0; i < N; i++) {
        int j = k - i;
        if (j > i && j < N) {
          t0 = mad_lo_cc(a.l[i], a.l[j], t0);
          t1 = madc_hi_cc(a.l[i], a.l[j], t1);
          t2 = addc(t2, 0);
        }
      }
      t2 = (t2 << 1) | (t1 >> 31);
      t1 = (t1 << 1) | (t0 >> 31);
      t0 = t0 << 1;
      c0 = add_cc(c0, t0);
      c1 = addc_cc(c1, t1);
      c2 = addc(c2, t2);
      if ((k & 1) == 0) {  // square term a_{k/2}^2
        c0 = mad_lo_cc(a.l[k / 2], a.l[k / 2], c0);
        c1 = madc_hi_cc(a.l[k / 2], a.l[k / 2], c1);
        c2 = addc(c2, 0);
      }
      // reduction terms sum_i m_i p_{k-i}
      CS_UNROLL
      for (int i = 0; i < N; i++) {
        int j = k - i;
        if (i < k && i < N && j >= 0 && j < N && (k < N ? i < k : true)) {
          if (k < N || i >= k - N + 1) {
            c0 = mad_lo_cc(m[i], P::mod(j), c0);
            c1 = madc_hi_cc(m[i], P::mod(j), c1);
            c2 = addc(c2, 0);
          }
        }
      }
      if (k < N) {
        m[k] = mul_lo(c0, P::M0);
        c0 = mad_lo_cc(m[k], P::mod(0), c0);
        c1 = madc_hi_cc(m[k], P::mod(0), c1);
        c2 = addc(c2, 0);
      } else {
        r.l[k - N] = c0;
      }
      c0 = c1; c1 = c2; c2 = 0;
    }
    r.l[N - 1] = c0;
    r.final_sub();
    return r;
  }

  // a b + c d with ONE Montgomery reduction (product scanning: both partial-product sets and the reduction terms of
  // a column meet in one 3-word accumulator): 2 N^2 + N^2 wide multiply-adds instead of the 4 N^2 of two products.
  // Needs 2 p < 2^(32 N), true for every modulus here (254 / 255 / 381 bits in 256 / 256 / 384).
  static CS_D Fp dot2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return dot2_ool(a, b, c, d); }
  static CS_DN Fp dot2_ool(Fp a, Fp b, Fp c, Fp d) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    uint32_t m[N];
    Fp r;
    CS_UNROLL
    for (int k = 0; k < 2 * N - 1; k++) {
      CS_UNROLL
      for (int i = 0; i < N; i++) {
        const int j = k - i;
        if (j >= 0 && j < N) {
          c0 = mad_lo_cc(a.l[i], b.l[j], c0);
          c1 = madc_hi_cc(a.l[i], b.l[j], c1);
          c2 = addc(c2, 0);
          c0 = mad_lo_cc(c.l[i], d.l[j], c0);
          c1 = madc_hi_cc(c.l[i], d.l[j], c1);
          c2 = addc(c2, 0);
        }
      }
      // reduction terms m_i p_j, i + j = k, of the rows already determined (j >= 1)
      CS_UNROLL
      for (int i = 0; i < N; i++) {
        const int j = k - i;
        if (j >= 1 && j < N) {
          c0 = mad_lo_cc(m[i], P::mod(j), c0);
          c1 = madc_hi_cc(m[i], P::mod(j), c1);
          c2 = addc(c2, 0);
        }
      }
      if (k < N) {
        m[k] = mul_lo(c0, P::M0);
        c0 = mad_lo_cc(m[k], P::mod(0), c0);
        c1 = madc_hi_cc(m[k], P::mod(0), c1);
        c2 = addc(c2, 0);
      } else {
        r.l[k - N] = c0;
      }
      c0 = c1; c1 = c2; c2 = 0;
    }
    r.l[N - 1] = c0;
    r.final_sub();
    return r;
  }

  // Montgomery <-> canonical
  CS_D Fp to_mont() const { return (*this) * r2(); }
  CS_D Fp from_mont() const {
    Fp o = zero();
    o.l[0] = 1;
    return (*this) * o;
  }

  // a^(p-2); only used off the per-proof path (table precomputation)
  CS_D Fp inverse() const {
    Fp res = one();
    Fp base = *this;
    for (int i = 0; i < N; i++) {
      uint32_t e = P::mod(i);
      if (i == 0) e -= 2;  // all supported moduli have mod[0] >= 2 (they are odd and > 2)
      for (int b = 0; b < 32; b++) {
        if ((e >> b) & 1) res = res * base;
        base = base.sqr();
      }
    }
    return res;
  }
};

}  // namespace cs
