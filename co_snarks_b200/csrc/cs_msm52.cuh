// MSM bucket accumulation on the FP64 pipe (BN254 G1): k_msm_accum0 with the mixed XYZZ addition carried out
// on 5 x 52-bit limbs (cs_field52.cuh).  Same slices, same order, same result as k_msm_accum0 -- the accumulated
// point is converted back to the 8 x 32-bit Montgomery form at the end of the slice, so everything downstream
// (fold levels, bucket reduction) is untouched and parity with the oracle stays bit-exact on the affine result.
//
// Table: unchanged geometry (W windows x n affine points, 64 bytes each), but the coordinates are stored in the
// radix-2^260 Montgomery form (canonical, 8 x 32-bit words), so that loading a point is pure bit slicing -- no
// per-point field product.  cs_bases_upload converts the table once (k_msm_table_to_m260).
//
// Lazy ranges (units of p; every multiplicand must be < 8, cs_field52.cuh):
//   X1 < 6, Y1 < 4, ZZ, ZZZ < 1.2 on entry.   U2, S2 < 1.02.   P = U2 - X1 + 6p < 7.02;  R = +-S2 - Y1 + {4,6}p < 6.
//   PP < 1.77, PPP < 1.2, Q < 1.17, R^2 < 1.57.   X3 = R^2 - PPP - 2Q + 4p < 5.57.   D = Q - X3 + 6p < 7.17.
//   Y3 = R D - Y1 PPP + 2p < 3.68.   ZZ' = ZZ PP < 1.04, ZZZ' = ZZZ PPP < 1.03.   The invariants close.
// P == 0 (mod p) -- the only place the special cases of the group law (P + P, P + (-P)) enter -- is detected by
// a one-word filter; a hit (always for true special cases, 2^-29 otherwise) re-runs that one addition with the
// exact 32-bit formulas.
#pragma once
#include "cs_curve.cuh"
#include "cs_msm.cuh"
#include "cs_params.cuh"
#include "cs_params52.cuh"
#include "cs_field52.cuh"

namespace cs {

struct Acc52 {
  I52 x, y, zz, zzz;
};

// 8 x 32-bit words -> 5 x 52-bit limbs (pure bit slicing; no change of Montgomery radix)
CS_D I52 f52_slice_words(const uint32_t* w) {
  const uint64_t w0 = w[0] | ((uint64_t)w[1] << 32), w1 = w[2] | ((uint64_t)w[3] << 32);
  const uint64_t w2 = w[4] | ((uint64_t)w[5] << 32), w3 = w[6] | ((uint64_t)w[7] << 32);
  I52 v;
  v.l[0] = w0 & F52_MASK;
  v.l[1] = ((w0 >> 52) | (w1 << 12)) & F52_MASK;
  v.l[2] = ((w1 >> 40) | (w2 << 24)) & F52_MASK;
  v.l[3] = ((w2 >> 28) | (w3 << 36)) & F52_MASK;
  v.l[4] = w3 >> 16;
  return v;
}
CS_D void f52_pack_words(const I52& v, uint32_t* w) {  // value < 2^256
  const uint64_t w0 = v.l[0] | (v.l[1] << 52), w1 = (v.l[1] >> 12) | (v.l[2] << 40);
  const uint64_t w2 = (v.l[2] >> 24) | (v.l[3] << 28), w3 = (v.l[3] >> 36) | (v.l[4] << 16);
  w[0] = (uint32_t)w0; w[1] = (uint32_t)(w0 >> 32); w[2] = (uint32_t)w1; w[3] = (uint32_t)(w1 >> 32);
  w[4] = (uint32_t)w2; w[5] = (uint32_t)(w2 >> 32); w[6] = (uint32_t)w3; w[7] = (uint32_t)(w3 >> 32);
}

// r = s (+-1) * a - b + K p
template <class P52, int K>
CS_D I52 f52_pm_sub(const I52& a, bool neg_a, const I52& b) {
  I52 r;
  int64_t carry = 0;
  CS_UNROLL
  for (int k = 0; k < 5; k++) {
    const int64_t av = (int64_t)a.l[k];
    const int64_t v = (neg_a ? -av : av) - (int64_t)b.l[k] + (int64_t)P52::kp(K, k) + carry;
    r.l[k] = (uint64_t)v & F52_MASK;
    carry = v >> 52;
  }
  return r;
}
// r = a - b - 2 c + K p
template <class P52, int K>
CS_D I52 f52_sub_sub2(const I52& a, const I52& b, const I52& c) {
  I52 r;
  int64_t carry = 0;
  CS_UNROLL
  for (int k = 0; k < 5; k++) {
    const int64_t v = (int64_t)a.l[k] - (int64_t)b.l[k] - 2 * (int64_t)c.l[k] + (int64_t)P52::kp(K, k) + carry;
    r.l[k] = (uint64_t)v & F52_MASK;
    carry = v >> 52;
  }
  return r;
}

// acc += (+-) p with p = table point (M260 canonical; never the point at infinity: those entries are dropped
// before the sort) and acc != infinity.  Returns false -- leaving acc untouched -- when P may be 0 (mod p): the
// caller then redoes its whole slice with the exact 32-bit formulas (accum_slice_exact).
template <class P52>
CS_D bool madd52(Acc52& a, const I52& x2i, const I52& y2i, bool negate) {
  const D52 x2 = f52_to_double(x2i), y2 = f52_to_double(y2i);
  const D52 zz = f52_to_double(a.zz), zzz = f52_to_double(a.zzz);
  const I52 U2 = f52_mul<P52>(x2, zz);
  const I52 S2 = f52_mul<P52>(y2, zzz);
  const I52 P = f52_sub<P52, 6>(U2, a.x);
  if (f52_maybe_zero_mod_p<P52>(P)) return false;
  const I52 R = negate ? f52_pm_sub<P52, 6>(S2, true, a.y) : f52_pm_sub<P52, 4>(S2, false, a.y);
  const D52 Pd = f52_to_double(P), Rd = f52_to_double(R);
  const I52 PP = f52_sqr<P52>(Pd);
  const D52 PPd = f52_to_double(PP);
  const I52 PPP = f52_mul<P52>(Pd, PPd);
  const D52 PPPd = f52_to_double(PPP);
  const I52 Q = f52_mul<P52>(f52_to_double(a.x), PPd);
  const I52 R2 = f52_sqr<P52>(Rd);
  const I52 X3 = f52_sub_sub2<P52, 4>(R2, PPP, Q);
  const I52 D = f52_sub<P52, 6>(Q, X3);
  const I52 RD = f52_mul<P52>(Rd, f52_to_double(D));
  const I52 YP = f52_mul<P52>(f52_to_double(a.y), PPPd);
  a.x = X3;
  a.y = f52_sub<P52, 2>(RD, YP);
  a.zz = f52_mul<P52>(zz, PPd);
  a.zzz = f52_mul<P52>(zzz, PPPd);
  return true;
}

// The exact path for one slice: the table points back to the 32-bit canonical Montgomery (R = 2^256) form and
// the proven formulas of cs_curve.cuh (P + P, P + (-P) and infinity handled there).  Reached when a slice holds a
// true special case, and for one slice in ~2^24 otherwise.  Scalar arguments only: nothing of the fast path's
// state has its address taken.
template <class P52, class P32>
CS_DN void accum_slice_exact(const Affine<Fp<P32>>* __restrict__ table, const uint32_t* __restrict__ sorted, uint32_t beg,
                             uint32_t end, Xyzz<Fp<P32>>* __restrict__ out) {
  typedef Fp<P32> F;
  Xyzz<F> acc = Xyzz<F>::inf();
  for (uint32_t k = beg; k < end; k++) {
    const uint32_t e = sorted[k];
    Affine<F> q = table[e & ~MSM_SIGN];
    Affine<F> p;
    p.x = f52_to_fp<P52, P32>(f52_slice_words(q.x.l));
    p.y = f52_to_fp<P52, P32>(f52_slice_words(q.y.l));
    madd(acc, p, (e & MSM_SIGN) != 0);
  }
  *out = acc;
}

// table (Montgomery R = 2^256, canonical) -> (Montgomery R = 2^260, canonical), in place; once per upload
template <class P52, class P32>
CS_GLOBAL void k_msm_table_to_m260(uint32_t* __restrict__ coords, size_t ncoords) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncoords) return;
  Fp<P32> x;
  CS_UNROLL
  for (int k = 0; k < 8; k++) x.l[k] = coords[i * 8 + k];
  I52 v = f52_from_fp<P52, P32>(x);  // x 2^260, < 2p
  // canonical: subtract p once if needed
  I52 t;
  int64_t carry = 0;
  CS_UNROLL
  for (int k = 0; k < 5; k++) {
    const int64_t d = (int64_t)v.l[k] - (int64_t)P52::kp(1, k) + carry;
    t.l[k] = (uint64_t)d & F52_MASK;
    carry = d >> 52;
  }
  if (carry == 0) v = t;  // no borrow: v >= p
  uint32_t w[8];
  f52_pack_words(v, w);
  CS_UNROLL
  for (int k = 0; k < 8; k++) coords[i * 8 + k] = w[k];
}

// k_msm_accum0 on the FP64 pipe.  Same arguments and slice order; `table` holds M260 coordinates.
template <class P52, class P32, int MINB>
CS_GLOBAL void __launch_bounds__(128, MINB) k_msm_accum0_f52(const Affine<Fp<P32>>* __restrict__ table,
                                                             const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ count,
                                                             const uint32_t* __restrict__ start,
                                                             const uint32_t* __restrict__ sstart0, uint32_t nb1, uint32_t S,
                                                             const uint32_t* __restrict__ order,
                                                             const uint32_t* __restrict__ order_b,
                                                             Xyzz<Fp<P32>>* __restrict__ part0) {
  typedef Fp<P32> F;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= sstart0[nb1]) return;
  uint32_t s = order[t];
  uint32_t b = order_b[t];
  uint32_t j = s - sstart0[b];
  uint32_t beg = start[b] + j * S;
  uint32_t end = start[b] + count[b];
  if (end > beg + S) end = beg + S;
  // the first entry starts the accumulator; from then on acc is never infinity on the fast path (a cancellation
  // P + (-P) shows up as P == 0 mod p and takes the exact path)
  uint32_t e = sorted[beg];
  Affine<F> p = table[e & ~MSM_SIGN];
  Acc52 acc;
  acc.x = f52_slice_words(p.x.l);
  {
    const I52 y = f52_slice_words(p.y.l);
    I52 z;
    CS_UNROLL
    for (int k = 0; k < 5; k++) { z.l[k] = 0; acc.zz.l[k] = acc.zzz.l[k] = P52::one(k); }
    acc.y = (e & MSM_SIGN) ? f52_sub<P52, 1>(z, y) : y;  // p - y: y != 0 (b != 0, odd group order)
  }
  bool ok = true;
  if (beg + 1 < end) {
    e = sorted[beg + 1];
    p = table[e & ~MSM_SIGN];
  }
  for (uint32_t k = beg + 1; k < end; k++) {
    const uint32_t e_cur = e;
    const I52 x2 = f52_slice_words(p.x.l), y2 = f52_slice_words(p.y.l);
    if (k + 1 < end) {  // prefetch the next point while this addition runs
      e = sorted[k + 1];
      p = table[e & ~MSM_SIGN];
    }
    ok = madd52<P52>(acc, x2, y2, (e_cur & MSM_SIGN) != 0);
    if (!ok) break;
  }
  if (!ok) {
    accum_slice_exact<P52, P32>(table, sorted, beg, end, part0 + s);
    return;
  }
  Xyzz<F> out;
  out.x = f52_to_fp<P52, P32>(acc.x);
  out.y = f52_to_fp<P52, P32>(acc.y);
  out.zz = f52_to_fp<P52, P32>(acc.zz);
  out.zzz = f52_to_fp<P52, P32>(acc.zzz);
  part0[s] = out;
}

// host-side dispatch used by msm_enqueue: only BN254 G1 has the FP64 path
template <class F>
int msm_accum0_f52(const Affine<F>*, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, uint32_t,
                   const uint32_t*, const uint32_t*, Xyzz<F>*, uint32_t, cudaStream_t) {
  return fail(-4, "msm: the FP64 accumulation exists for BN254 G1 only");
}
template <>
inline int msm_accum0_f52<Fp<Bn254Fq>>(const Affine<Fp<Bn254Fq>>* table, const uint32_t* sorted, const uint32_t* count,
                                       const uint32_t* start, const uint32_t* sstart0, uint32_t nb1, uint32_t S,
                                       const uint32_t* order, const uint32_t* order_b, Xyzz<Fp<Bn254Fq>>* part0,
                                       uint32_t max_s0, cudaStream_t st) {
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("CS_ACCUM0_F52_MINB"); minb = e ? atoi(e) : 3; }
#define CS_ACC52(M)                                                                                                  \
  CS_LAUNCH(k_msm_accum0_f52<Bn254Fq52 COMMA Bn254Fq COMMA M>, ceil_div(max_s0, 128), 128, 0, st, table, sorted, count, \
            start, sstart0, nb1, S, order, order_b, part0)
  switch (minb) {
    case 2: CS_ACC52(2); break;
    case 4: CS_ACC52(4); break;
    default: CS_ACC52(3); break;
  }
#undef CS_ACC52
  return 0;
}

}  // namespace cs
