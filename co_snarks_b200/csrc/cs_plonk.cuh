// Kernels of the Plonk prover (plain driver): everything co-plonk's rounds do per element, per prefix or
// per polynomial runs here on device-resident vectors; the host only hashes the transcript.
//
// Reference: co-circom/co-plonk/src/round1.rs (wire polynomials, additions), round2.rs:95-190 (z),
// round3.rs:20-108,240-520 (quotient with the blinding bookkeeping), round5.rs:78-280 (linearisation,
// division by X - xi), mpc/plain.rs:199-247 (array_prod_mul = running products).
#pragma once
#include "cs_common.cuh"
#include "cs_field.cuh"
#include "cs_ntt.cuh"

namespace cs {

// w^k from the table of the first half of the powers (w^(half) = -1)
template <class FrP>
CS_D Fp<FrP> root_pow(const uint32_t* __restrict__ tw, uint32_t half, uint32_t k) {
  if (k < half) return ld_fr<FrP>(tw + (size_t)k * FrP::N);
  return Fp<FrP>::zero() - ld_fr<FrP>(tw + (size_t)(k - half) * FrP::N);
}

// x^e by square-and-multiply over pw[j] = x^(2^j)
template <class FrP>
CS_D Fp<FrP> table_pow(const uint32_t* __restrict__ pw, uint64_t e) {
  Fp<FrP> acc = Fp<FrP>::one();
  for (uint32_t j = 0; (e >> j) != 0; j++)
    if ((e >> j) & 1) acc = acc * ld_fr<FrP>(pw + (size_t)j * FrP::N);
  return acc;
}

// additions of one dependency level (round1.rs:191-224): w[first_add + k] = w[s1] f1 + w[s2] f2
// `batch` components per signal (1 plain, 2 Rep3 share {a, b}: mul_with_public / add act per component)
template <class FrP>
CS_GLOBAL void k_plonk_additions(const uint32_t* __restrict__ order, uint32_t lo, uint32_t hi,
                                 const uint32_t* __restrict__ ids, const uint32_t* __restrict__ factors,
                                 uint32_t first_add, uint32_t batch, uint32_t* __restrict__ w) {
  uint32_t t = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hi) return;
  uint32_t k = order[t];
  Fp<FrP> f1 = ld_fr<FrP>(factors + (size_t)(2 * k) * FrP::N), f2 = ld_fr<FrP>(factors + (size_t)(2 * k + 1) * FrP::N);
  for (uint32_t c = 0; c < batch; c++) {
    Fp<FrP> a = ld_fr<FrP>(w + ((size_t)ids[2 * k] * batch + c) * FrP::N) * f1;
    Fp<FrP> b = ld_fr<FrP>(w + ((size_t)ids[2 * k + 1] * batch + c) * FrP::N) * f2;
    st_fr<FrP>(w + ((size_t)(first_add + k) * batch + c) * FrP::N, a + b);
  }
}

// buffer[i] = w[map[i]] for i < nc, 0 up to n (round1.rs:118-124)
template <class FrP>
CS_GLOBAL void k_plonk_gather(const uint32_t* __restrict__ map, uint32_t nc, uint32_t n, uint32_t batch,
                               const uint32_t* __restrict__ w, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (uint32_t c = 0; c < batch; c++) {
    Fp<FrP> v = Fp<FrP>::zero();
    if (i < nc) v = ld_fr<FrP>(w + ((size_t)map[i] * batch + c) * FrP::N);
    st_fr<FrP>(out + ((size_t)i * batch + c) * FrP::N, v);
  }
}

struct PlonkBlind { uint32_t v[6][8]; };  // count (<= 3) coefficients x batch (<= 2) components
// blind_coefficients (lib.rs:163-178): poly[i] -= rev[i], poly[n + i] = rev[i]   (one thread per element)
template <class FrP>
CS_GLOBAL void k_plonk_blind(uint32_t* __restrict__ poly, uint32_t n, uint32_t batch, PlonkBlind rev, uint32_t count) {
  uint32_t t = threadIdx.x;
  if (t >= count * batch) return;
  uint32_t i = t / batch, cix = t - i * batch;
  Fp<FrP> c;
  for (int l = 0; l < FrP::N; l++) c.l[l] = rev.v[t][l];
  uint32_t* lo = poly + ((size_t)i * batch + cix) * FrP::N;
  st_fr<FrP>(lo, ld_fr<FrP>(lo) - c);
  st_fr<FrP>(poly + ((size_t)(n + i) * batch + cix) * FrP::N, c);
}

// component `comp` of an interleaved share vector as a contiguous vector (the additive share a party works on)
template <class FrP>
CS_GLOBAL void k_extract_component(const uint32_t* __restrict__ in, uint32_t n, uint32_t batch, uint32_t comp,
                                   uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_fr<FrP>(out + (size_t)i * FrP::N, ld_fr<FrP>(in + ((size_t)i * batch + comp) * FrP::N));
}

// Challenges, blinders and the small constants of rounds 2-3, by value (Montgomery limbs)
struct PlonkConsts {
  uint32_t b[11][8];
  uint32_t beta[8], gamma[8], alpha[8], alpha2[8], k1[8], k2[8];
  uint32_t z1[4][8], z2[4][8], z3[4][8];
};
template <class FrP>
CS_D Fp<FrP> cload(const uint32_t* p) {
  Fp<FrP> r;
  CS_UNROLL
  for (int l = 0; l < FrP::N; l++) r.l[l] = p[l];
  return r;
}

// round2.rs:99-160: num_i = (a + beta w^i + gamma)(b + k1 beta w^i + gamma)(c + k2 beta w^i + gamma),
//                   den_i = (a + beta s1(w^i) + gamma)(b + beta s2 + gamma)(c + beta s3 + gamma)
// sigma evaluations are read from the 4n-point tables at stride 4; w^i = w4^(4 i).
template <class FrP>
CS_GLOBAL void k_plonk_numden(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                               const uint32_t* __restrict__ c, const uint32_t* __restrict__ s1,
                               const uint32_t* __restrict__ s2, const uint32_t* __restrict__ s3,
                               const uint32_t* __restrict__ tw4, uint32_t n, PlonkConsts K,
                               uint32_t* __restrict__ num, uint32_t* __restrict__ den) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F beta = cload<FrP>(K.beta), gamma = cload<FrP>(K.gamma);
  F av = ld_fr<FrP>(a + (size_t)i * NW), bv = ld_fr<FrP>(b + (size_t)i * NW), cv = ld_fr<FrP>(c + (size_t)i * NW);
  F bw = beta * root_pow<FrP>(tw4, 2 * n, 4 * i);
  F nn = (av + bw + gamma) * (bv + cload<FrP>(K.k1) * bw + gamma) * (cv + cload<FrP>(K.k2) * bw + gamma);
  F dd = (av + beta * ld_fr<FrP>(s1 + (size_t)(4 * i) * NW) + gamma) *
         (bv + beta * ld_fr<FrP>(s2 + (size_t)(4 * i) * NW) + gamma) *
         (cv + beta * ld_fr<FrP>(s3 + (size_t)(4 * i) * NW) + gamma);
  st_fr<FrP>(num + (size_t)i * NW, nn);
  st_fr<FrP>(den + (size_t)i * NW, dd);
}

// ---- scans over field elements: OP = 0 product, 1 sum; REV scans from the end (suffix) ----------------
// Three phases: per-block inclusive scan (SCAN_ITEMS consecutive elements per thread) + block totals,
// scan of the block totals (one block), apply.
constexpr uint32_t SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
template <class FrP, int OP>
CS_D Fp<FrP> scan_op(const Fp<FrP>& x, const Fp<FrP>& y) { return OP == 0 ? x * y : x + y; }
template <class FrP, int OP>
CS_D Fp<FrP> scan_id() { return OP == 0 ? Fp<FrP>::one() : Fp<FrP>::zero(); }

template <class FrP, int OP>
CS_GLOBAL void k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int rev,
                            uint32_t* __restrict__ totals) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  CS_DYN_SMEM(uint32_t, sm);  // SCAN_THREADS elements
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  F v[SCAN_ITEMS];
  F run = scan_id<FrP, OP>();
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    uint32_t p = base + k;
    F x = scan_id<FrP, OP>();
    if (p < n) x = ld_fr<FrP>(in + (size_t)(rev ? n - 1 - p : p) * NW);
    run = scan_op<FrP, OP>(run, x);
    v[k] = run;
  }
  st_fr<FrP>(sm + (size_t)threadIdx.x * NW, run);
  __syncthreads();
  // Hillis-Steele over the per-thread totals
  for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {
    F t = ld_fr<FrP>(sm + (size_t)threadIdx.x * NW);
    F o = scan_id<FrP, OP>();
    bool has = threadIdx.x >= d;
    if (has) o = ld_fr<FrP>(sm + (size_t)(threadIdx.x - d) * NW);
    __syncthreads();
    if (has) st_fr<FrP>(sm + (size_t)threadIdx.x * NW, scan_op<FrP, OP>(o, t));
    __syncthreads();
  }
  F prefix = scan_id<FrP, OP>();
  if (threadIdx.x > 0) prefix = ld_fr<FrP>(sm + (size_t)(threadIdx.x - 1) * NW);
  for (uint32_t k = 0; k < SCAN_ITEMS; k++) {
    uint32_t p = base + k;
    if (p < n) st_fr<FrP>(out + (size_t)(rev ? n - 1 - p : p) * NW, scan_op<FrP, OP>(prefix, v[k]));
  }
  if (threadIdx.x == SCAN_THREADS - 1) st_fr<FrP>(totals + (size_t)blockIdx.x * NW, ld_fr<FrP>(sm + (size_t)threadIdx.x * NW));
}

// exclusive scan of the block totals in place (single block, serial per thread chunk + Hillis-Steele)
template <class FrP, int OP>
CS_GLOBAL void k_scan_totals(uint32_t* __restrict__ totals, uint32_t nb) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  CS_DYN_SMEM(uint32_t, sm);
  const uint32_t per = (nb + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = threadIdx.x * per;
  F run = scan_id<FrP, OP>();
  for (uint32_t k = lo; k < lo + per && k < nb; k++) run = scan_op<FrP, OP>(run, ld_fr<FrP>(totals + (size_t)k * NW));
  st_fr<FrP>(sm + (size_t)threadIdx.x * NW, run);
  __syncthreads();
  for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
    F t = ld_fr<FrP>(sm + (size_t)threadIdx.x * NW);
    F o = scan_id<FrP, OP>();
    bool has = threadIdx.x >= d;
    if (has) o = ld_fr<FrP>(sm + (size_t)(threadIdx.x - d) * NW);
    __syncthreads();
    if (has) st_fr<FrP>(sm + (size_t)threadIdx.x * NW, scan_op<FrP, OP>(o, t));
    __syncthreads();
  }
  F acc = scan_id<FrP, OP>();
  if (threadIdx.x > 0) acc = ld_fr<FrP>(sm + (size_t)(threadIdx.x - 1) * NW);
  for (uint32_t k = lo; k < lo + per && k < nb; k++) {
    F x = ld_fr<FrP>(totals + (size_t)k * NW);
    st_fr<FrP>(totals + (size_t)k * NW, acc);  // exclusive
    acc = scan_op<FrP, OP>(acc, x);
  }
}

template <class FrP, int OP>
CS_GLOBAL void k_scan_apply(uint32_t* __restrict__ out, uint32_t n, int rev, const uint32_t* __restrict__ totals) {
  constexpr int NW = FrP::N;
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t blk = p / SCAN_TILE;
  if (blk == 0) return;
  size_t idx = rev ? n - 1 - p : p;
  st_fr<FrP>(out + idx * NW, scan_op<FrP, OP>(ld_fr<FrP>(totals + (size_t)blk * NW), ld_fr<FrP>(out + idx * NW)));
}

// buffer_z rotated right by one (round2.rs:166-167): z[(i + 1) mod n] = Pnum[i] * (1 / Pden[n-1]) * Sden[i + 1]
// with Pnum the running products of num, Sden the suffix products of den (Sden[n] = 1): the running
// quotient num_0..i / den_0..i with ONE field inversion for the whole vector.
template <class FrP>
CS_GLOBAL void k_plonk_zbuf(const uint32_t* __restrict__ pnum, const uint32_t* __restrict__ sden,
                            const uint32_t* __restrict__ inv_total, uint32_t n, uint32_t* __restrict__ z) {
  constexpr int NW = FrP::N;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<FrP> v = ld_fr<FrP>(pnum + (size_t)i * NW) * ld_fr<FrP>(inv_total);
  if (i + 1 < n) v = v * ld_fr<FrP>(sden + (size_t)(i + 1) * NW);
  st_fr<FrP>(z + (size_t)((i + 1) % n) * NW, v);
}

// (A + Ap Z)(B + Bp Z)(C + Cp Z)(D + Dp Z) on the extended domain: value and blinding part
// (mul4vec + mul4vec_post, round3.rs:20-108)
template <class FrP>
CS_D void plonk_mul4(const Fp<FrP>& a, const Fp<FrP>& b, const Fp<FrP>& c, const Fp<FrP>& d, const Fp<FrP>& ap,
                     const Fp<FrP>& bp, const Fp<FrP>& cp, const Fp<FrP>& dp, uint32_t m, const PlonkConsts& K,
                     Fp<FrP>& r, Fp<FrP>& rz) {
  typedef Fp<FrP> F;
  F ab = a * b, abp = a * bp, apb = ap * b, apbp = ap * bp;
  F cd = c * d, cdp = c * dp, cpd = cp * d, cpdp = cp * dp;
  r = ab * cd;
  rz = apb * cd + abp * cd + ab * cpd + ab * cdp;
  if (m) {
    F x1 = apbp * cd + apb * cpd + apb * cdp + abp * cpd + abp * cdp + ab * cpdp;
    F x2 = abp * cpdp + apb * cpdp + apbp * cdp + apbp * cpd;
    F x3 = apbp * cpdp;
    rz = rz + x1 * cload<FrP>(K.z1[m]) + x2 * cload<FrP>(K.z2[m]) + x3 * cload<FrP>(K.z3[m]);
  }
}

struct PlonkQuotIn {
  const uint32_t *a, *b, *c, *z;                      // 4n evaluations of the unblinded a, b, c, z
  const uint32_t *qm, *ql, *qr, *qo, *qc, *s1, *s2, *s3;  // 4n evaluations from the key
  const uint32_t* lagrange;                            // nlag x 4n
  const uint32_t* buf_a;                               // n (public-input rows come first)
  const uint32_t* tw4;                                 // w4^k, k < 2n
};
// compute_t's per-point loop (round3.rs:300-520) for the plain driver: t and tz at every point of the 4n domain
template <class FrP>
CS_GLOBAL void k_plonk_quotient(PlonkQuotIn in, uint32_t n, uint32_t nlag, PlonkConsts K,
                                 uint32_t* __restrict__ t_out, uint32_t* __restrict__ tz_out) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  const uint32_t n4 = 4 * n;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uint32_t m = i & 3;
  F w = root_pow<FrP>(in.tw4, 2 * n, i);
  F a = ld_fr<FrP>(in.a + (size_t)i * NW), b = ld_fr<FrP>(in.b + (size_t)i * NW), c = ld_fr<FrP>(in.c + (size_t)i * NW);
  F z = ld_fr<FrP>(in.z + (size_t)i * NW);
  F zw = ld_fr<FrP>(in.z + (size_t)((i + 4) % n4) * NW);
  F ap = cload<FrP>(K.b[1]) + cload<FrP>(K.b[0]) * w;
  F bp = cload<FrP>(K.b[3]) + cload<FrP>(K.b[2]) * w;
  F cp = cload<FrP>(K.b[5]) + cload<FrP>(K.b[4]) * w;
  F b6 = cload<FrP>(K.b[6]), b7 = cload<FrP>(K.b[7]), b8 = cload<FrP>(K.b[8]);
  F zp = b6 * w.sqr() + b7 * w + b8;
  F ww = root_pow<FrP>(in.tw4, 2 * n, (i + 4) % n4);  // w * w_n
  F zwp = b6 * ww.sqr() + b7 * ww + b8;
  F a_b = a * b, a_bp = a * bp, ap_b = ap * b;
  F a0 = a_bp + ap_b;
  if (m) a0 = a0 + ap * bp * cload<FrP>(K.z1[m]);
  F qm = ld_fr<FrP>(in.qm + (size_t)i * NW), ql = ld_fr<FrP>(in.ql + (size_t)i * NW);
  F qr = ld_fr<FrP>(in.qr + (size_t)i * NW), qo = ld_fr<FrP>(in.qo + (size_t)i * NW);
  F e1 = a_b * qm + a * ql + b * qr + c * qo;
  F e1z = a0 * qm + ap * ql + bp * qr + cp * qo;
  F pi = F::zero();
  for (uint32_t j = 0; j < nlag; j++)
    pi = pi - ld_fr<FrP>(in.buf_a + (size_t)j * NW) * ld_fr<FrP>(in.lagrange + ((size_t)j * n4 + i) * NW);
  e1 = e1 + pi + ld_fr<FrP>(in.qc + (size_t)i * NW);
  F beta = cload<FrP>(K.beta), gamma = cload<FrP>(K.gamma), alpha = cload<FrP>(K.alpha);
  F bw = beta * w;
  F e2, e2z, e3, e3z;
  plonk_mul4<FrP>(a + bw + gamma, b + bw * cload<FrP>(K.k1) + gamma, c + bw * cload<FrP>(K.k2) + gamma, z, ap, bp, cp, zp,
                  m, K, e2, e2z);
  plonk_mul4<FrP>(a + ld_fr<FrP>(in.s1 + (size_t)i * NW) * beta + gamma, b + ld_fr<FrP>(in.s2 + (size_t)i * NW) * beta + gamma,
                  c + ld_fr<FrP>(in.s3 + (size_t)i * NW) * beta + gamma, zw, ap, bp, cp, zwp, m, K, e3, e3z);
  F l0a2 = ld_fr<FrP>(in.lagrange + (size_t)i * NW) * cload<FrP>(K.alpha2);
  F e4 = (z - F::one()) * l0a2;
  F e4z = zp * l0a2;
  st_fr<FrP>(t_out + (size_t)i * NW, e1 + (e2 - e3) * alpha + e4);
  st_fr<FrP>(tz_out + (size_t)i * NW, e1z + (e2z - e3z) * alpha + e4z);
}

// round3.rs:522-556: division of t by Z_H in coefficient form (a chain of four per residue class mod n), plus
// the blinding part tz, split into t1 | t2 | t3 with the b9, b10 adjustments.  t1, t2: n + 1 entries, t3: n + 6.
template <class FrP>
CS_GLOBAL void k_plonk_tsplit(const uint32_t* __restrict__ ct, const uint32_t* __restrict__ ctz, uint32_t n,
                               PlonkConsts K, uint32_t* __restrict__ t1, uint32_t* __restrict__ t2,
                               uint32_t* __restrict__ t3) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  F c0 = F::zero() - ld_fr<FrP>(ct + (size_t)j * NW);
  F c1 = c0 - ld_fr<FrP>(ct + (size_t)(j + n) * NW);
  F c2 = c1 - ld_fr<FrP>(ct + (size_t)(j + 2 * n) * NW);
  F v1 = c0 + ld_fr<FrP>(ctz + (size_t)j * NW);
  F v2 = c1 + ld_fr<FrP>(ctz + (size_t)(j + n) * NW);
  F v3 = c2 + ld_fr<FrP>(ctz + (size_t)(j + 2 * n) * NW);
  F b9 = cload<FrP>(K.b[9]), b10 = cload<FrP>(K.b[10]);
  if (j == 0) {
    v2 = v2 - b9;
    v3 = v3 - b10;
    st_fr<FrP>(t1 + (size_t)n * NW, b9);
    st_fr<FrP>(t2 + (size_t)n * NW, b10);
  }
  st_fr<FrP>(t1 + (size_t)j * NW, v1);
  st_fr<FrP>(t2 + (size_t)j * NW, v2);
  st_fr<FrP>(t3 + (size_t)j * NW, v3);
  if (j < 6) {  // t3 takes n + 6 entries of the last quarter
    F c3 = c2 - ld_fr<FrP>(ct + (size_t)(j + 3 * n) * NW);
    st_fr<FrP>(t3 + (size_t)(n + j) * NW, c3 + ld_fr<FrP>(ctz + (size_t)(j + 3 * n) * NW));
  }
}

// round5.rs:120-260: the numerator of W_xi, res = r(X) + v0 a + v1 b + v2 c + v3 s1 + v4 s2 - constants,
// with r(X) = ab qm + a ql + b qr + c qo + qc - e3 beta s3 + e24 z - zh (t1 + xi^n t2 + xi^2n t3) + r0.
struct PlonkLinIn {
  const uint32_t *qm, *ql, *qr, *qo, *qc, *s1, *s2, *s3;  // n coefficients each
  const uint32_t *pa, *pb, *pc;                             // n + 2
  const uint32_t* pz;                                       // n + 3
  const uint32_t *t1, *t2, *t3;                             // n + 1, n + 1, n + 6
};
struct PlonkLinW {
  uint32_t ab[8], ea[8], eb[8], ec[8], e3beta[8], e24[8], zh[8], xin[8], xin2[8], v[5][8], c0[8];
};
template <class FrP>
CS_GLOBAL void k_plonk_wxi_numerator(PlonkLinIn in, PlonkLinW W, uint32_t n, int pub, uint32_t* __restrict__ out) {
  // pub = 0: this party holds additive shares only and contributes no public polynomial / constant terms
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n + 6) return;
  F acc = F::zero();
  if (i < n && pub) {
    acc = ld_fr<FrP>(in.qm + (size_t)i * NW) * cload<FrP>(W.ab) + ld_fr<FrP>(in.ql + (size_t)i * NW) * cload<FrP>(W.ea) +
          ld_fr<FrP>(in.qr + (size_t)i * NW) * cload<FrP>(W.eb) + ld_fr<FrP>(in.qo + (size_t)i * NW) * cload<FrP>(W.ec) +
          ld_fr<FrP>(in.qc + (size_t)i * NW) - ld_fr<FrP>(in.s3 + (size_t)i * NW) * cload<FrP>(W.e3beta) +
          ld_fr<FrP>(in.s1 + (size_t)i * NW) * cload<FrP>(W.v[3]) + ld_fr<FrP>(in.s2 + (size_t)i * NW) * cload<FrP>(W.v[4]);
  }
  if (i < n + 3) acc = acc + ld_fr<FrP>(in.pz + (size_t)i * NW) * cload<FrP>(W.e24);
  F tt = ld_fr<FrP>(in.t3 + (size_t)i * NW) * cload<FrP>(W.xin2);
  if (i < n + 1) tt = tt + ld_fr<FrP>(in.t2 + (size_t)i * NW) * cload<FrP>(W.xin) + ld_fr<FrP>(in.t1 + (size_t)i * NW);
  acc = acc - tt * cload<FrP>(W.zh);
  if (i < n + 2)
    acc = acc + ld_fr<FrP>(in.pa + (size_t)i * NW) * cload<FrP>(W.v[0]) + ld_fr<FrP>(in.pb + (size_t)i * NW) * cload<FrP>(W.v[1]) +
          ld_fr<FrP>(in.pc + (size_t)i * NW) * cload<FrP>(W.v[2]);
  if (i == 0 && pub) acc = acc + cload<FrP>(W.c0);
  st_fr<FrP>(out + (size_t)i * NW, acc);
}

// Division by (X - x) (round5.rs:78-93 with n = 1): q_i = -x^-(i+1) * sum_{j<=i} p_j x^j, as
// scale by x^j -> running sums -> scale by -x^-(i+1).  pw = table of base^(2^k); out[i] = in[i] * (neg ? -1 : 1) *
// base^(i + shift).  `sub0` (nullable) is subtracted from in[0] first (the "- eval" constant of W_xiw).
template <class FrP>
CS_GLOBAL void k_scale_by_powers(const uint32_t* __restrict__ in, const uint32_t* __restrict__ pw, uint32_t shift,
                                 int neg, const uint32_t* __restrict__ sub0, uint32_t n, uint32_t* __restrict__ out) {
  typedef Fp<FrP> F;
  constexpr int NW = FrP::N;
  constexpr uint32_t CH = 8;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = t * CH;
  if (lo >= n) return;
  F base = ld_fr<FrP>(pw);
  F p = table_pow<FrP>(pw, (uint64_t)lo + shift);
  if (neg) p = F::zero() - p;
  for (uint32_t i = lo; i < lo + CH && i < n; i++) {
    F x = ld_fr<FrP>(in + (size_t)i * NW);
    if (sub0 && i == 0) x = x - ld_fr<FrP>(sub0);
    st_fr<FrP>(out + (size_t)i * NW, x * p);
    p = p * base;
  }
}

}  // namespace cs
