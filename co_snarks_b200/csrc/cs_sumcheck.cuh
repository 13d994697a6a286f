// UltraHonk sumcheck, prover-side kernels ("sumcheck is a new kernel family").
//
// Replaces, for the rows of one sumcheck round (n = current round size):
//  * GateSeparatorPolynomial::new                     co-noir/ultrahonk/src/decider/types.rs:53-67
//        beta_products[j] = prod_{i : bit i of j} beta_i
//  * partially_evaluate_init / partially_evaluate_inplace
//        co-noir/co-ultrahonk/src/co_decider/co_sumcheck/co_sumcheck_prover.rs:33-97
//        out[i] = p[2i] + (p[2i+1] - p[2i]) * u          on public values and, per component, on Rep3 shares
//  * SumcheckRound::compute_univariate_inner's edge loop (co_sumcheck_round.rs:261-305: extend_edges :53-72 +
//    fold_and_filter with the scaling factor beta_products[(edge >> 1) * periodicity]) for the
//    UltraArithmeticRelation (relations/ultra_arithmetic_relation.rs:87-241, plain ultrahonk/src/decider/relations/
//    ultra_arithmetic_relation.rs): the two sub-relation accumulators r0 (6 evaluations, a half share under Rep3) and
//    r1 (5 evaluations, a Rep3 share) summed over all edges of the round.
//
// One thread per edge; for every evaluation point k = 0..6 the thread re-reads its two rows (L1/L2 hits), extends them
// to k by repeated addition of the edge's slope, evaluates the relation, and the 32 lanes' contributions meet in a
// warp-shuffle tree before one lane adds them to the block's accumulators in shared memory.  Block results go to
// `partial`, a second launch sums them.  Under Rep3 the per-element random masks of local_mul_vec
// (rep3/arithmetic.rs:132-146) are replaced by ONE zero share per evaluation, added by the host entry point: the
// masks only ever reach the protocol through this sum.
#pragma once
#include "cs_common.cuh"
#include "cs_field.cuh"
#include "cs_curve.cuh"
#include "cs_ntt.cuh"  // ld_fr / st_fr

namespace cs {

constexpr int SC_MAX_PARTIAL = 7;   // MAX_PARTIAL_RELATION_LENGTH (co-noir-common/src/constants.rs)
constexpr int SC_R0_LEN = 6, SC_R1_LEN = 5;
constexpr int SC_SLOTS = SC_R0_LEN + 2 * SC_R1_LEN;  // r0[k] | r1[k].a | r1[k].b

template <class FrP>
CS_GLOBAL void k_sc_gate_separator(const uint32_t* __restrict__ betas, uint32_t log_n, uint32_t* __restrict__ out) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >> log_n) return;
  Fp<FrP> acc = Fp<FrP>::one();
  for (uint32_t i = 0; i < log_n; i++)
    if ((j >> i) & 1) acc = acc * ld_fr<FrP>(betas + i * FrP::N);
  st_fr<FrP>(out + j * FrP::N, acc);
}

constexpr int SC_FOLD_MAX = 64;  // polynomials per launch
struct ScFoldArgs {
  const uint32_t* in[SC_FOLD_MAX];
  uint32_t* out[SC_FOLD_MAX];
};
// grid.y = polynomial; comps = 1 (public) or 2 (Rep3 share: both components, mul_with_public)
template <class FrP>
CS_GLOBAL void k_sc_fold(ScFoldArgs a, uint32_t comps, size_t half, Fp<FrP> u, int pad_zero) {
  const uint32_t* __restrict__ in = a.in[blockIdx.y];
  uint32_t* __restrict__ out = a.out[blockIdx.y];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (; i < half * comps; i += step) {
    const size_t e = i / comps, c = i % comps;
    Fp<FrP> lo = ld_fr<FrP>(in + ((2 * e) * comps + c) * FrP::N), hi = ld_fr<FrP>(in + ((2 * e + 1) * comps + c) * FrP::N);
    st_fr<FrP>(out + i * FrP::N, lo + (hi - lo) * u);
    // "if poly.len() < 2 { poly.push(zero) }" (co_sumcheck_prover.rs:75-77,91-93)
    if (pad_zero) st_fr<FrP>(out + (comps + i) * FrP::N, Fp<FrP>::zero());
  }
}

struct ScArithPolys {
  const uint32_t *w_l, *w_r, *w_o, *w_4, *w_l_shift, *w_4_shift;     // public values or Rep3 shares
  const uint32_t *q_m, *q_l, *q_r, *q_o, *q_4, *q_c, *q_arith;       // public
};

// value at k of the edge (row 2e, row 2e + 1) of a public polynomial: Univariate::extend_from, length 2
template <class FrP>
CS_D Fp<FrP> sc_ext(const uint32_t* __restrict__ p, size_t e, uint32_t comps, uint32_t c, int k) {
  Fp<FrP> a = ld_fr<FrP>(p + ((2 * e) * comps + c) * FrP::N);
  if (k == 0) return a;
  Fp<FrP> b = ld_fr<FrP>(p + ((2 * e + 1) * comps + c) * FrP::N);
  Fp<FrP> d = b - a;
  for (int j = 1; j < k; j++) b = b + d;
  return b;
}

template <class P>
CS_D Fp<P> warp_sum(Fp<P> v) {
  CS_UNROLL
  for (unsigned d = 16; d > 0; d >>= 1) v = v + shfl_down(v, d);
  return v;
}

// SH: witness polynomials are Rep3 shares.  partial: [gridDim.x][SC_SLOTS] field elements.
template <class FrP, bool SH>
CS_GLOBAL void __launch_bounds__(128) k_sc_arith_round(ScArithPolys p, size_t n_edges,
                                                       const uint32_t* __restrict__ beta_products, size_t periodicity,
                                                       int party, Fp<FrP> neg_half, uint32_t* __restrict__ partial) {
  typedef Fp<FrP> F;
  constexpr uint32_t WC = SH ? 2 : 1;
  __shared__ __align__(16) uint32_t acc_sm[4][SC_SLOTS][FrP::N];
  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
  for (uint32_t k = t; k < 4 * SC_SLOTS * FrP::N; k += blockDim.x) (&acc_sm[0][0][0])[k] = 0;
  __syncthreads();
  const size_t e = (size_t)blockIdx.x * blockDim.x + t;
  const bool live = e < n_edges;
  const size_t ee = live ? e : 0;
  const F scaling = live ? ld_fr<FrP>(beta_products + ee * periodicity * FrP::N) : F::zero();
  const F one = F::one(), two = one + one, three = two + one;
  for (int k = 0; k < SC_MAX_PARTIAL - 1; k++) {
    const F q_arith = sc_ext<FrP>(p.q_arith, ee, 1, 0, k);
    const F q_m = sc_ext<FrP>(p.q_m, ee, 1, 0, k);
    const F wl_a = sc_ext<FrP>(p.w_l, ee, WC, 0, k), w4_a = sc_ext<FrP>(p.w_4, ee, WC, 0, k);
    F wl_b = F::zero(), w4_b = F::zero();
    if (SH) { wl_b = sc_ext<FrP>(p.w_l, ee, WC, 1, k); w4_b = sc_ext<FrP>(p.w_4, ee, WC, 1, k); }
    // ---- r0 = q_arith * [ -1/2 (q_arith - 3) q_m w_l w_r + q_l w_l + q_r w_r + q_o w_o + q_4 w_4 + q_c
    //                       + (q_arith - 1) w_4_shift ] * scaling            (ultra_arithmetic_relation.rs:87-176)
    F c0;
    {
      const F wr_a = sc_ext<FrP>(p.w_r, ee, WC, 0, k);
      F mul;
      if (SH) {
        const F wr_b = sc_ext<FrP>(p.w_r, ee, WC, 1, k);
        mul = F::dot2(wl_a, wr_a + wr_b, wl_b, wr_a);  // local_mul_vec without the mask, one reduction for both products
      } else {
        mul = wl_a * wr_a;
      }
      F tmp = mul * q_m;
      tmp = tmp * (q_arith - three);
      tmp = tmp * neg_half;
      // mul_with_public_to_half_share = public * share.a (co-noir-common/src/mpc/rep3.rs:90-95)
      tmp = tmp + sc_ext<FrP>(p.q_l, ee, 1, 0, k) * wl_a + sc_ext<FrP>(p.q_r, ee, 1, 0, k) * wr_a;
      tmp = tmp + sc_ext<FrP>(p.q_o, ee, 1, 0, k) * sc_ext<FrP>(p.w_o, ee, WC, 0, k) + sc_ext<FrP>(p.q_4, ee, 1, 0, k) * w4_a;
      if (!SH || party == 0) tmp = tmp + sc_ext<FrP>(p.q_c, ee, 1, 0, k);  // add_assign_public_half_share: party 0 only
      tmp = tmp + (q_arith - one) * sc_ext<FrP>(p.w_4_shift, ee, WC, 0, k);
      tmp = tmp * q_arith;
      c0 = tmp * scaling;
    }
    c0 = warp_sum(c0);
    // ---- r1 = (w_l + w_4 - w_l_shift + q_m) (q_arith - 2)(q_arith - 1) q_arith * scaling     (:177-241)
    F c1a = F::zero(), c1b = F::zero();
    if (k < SC_R1_LEN) {
      const F f = (q_arith - two) * (q_arith - one) * q_arith * scaling;
      F ta = wl_a + w4_a - sc_ext<FrP>(p.w_l_shift, ee, WC, 0, k);
      if (!SH || party == 0) ta = ta + q_m;  // add_public: party 0's a, party 1's b (rep3/arithmetic.rs:41-48)
      c1a = ta * f;
      if (SH) {
        F tb = wl_b + w4_b - sc_ext<FrP>(p.w_l_shift, ee, WC, 1, k);
        if (party == 1) tb = tb + q_m;
        c1b = tb * f;
      }
      c1a = warp_sum(c1a);
      if (SH) c1b = warp_sum(c1b);
    }
    if (lane == 0) {
      st_fr<FrP>(acc_sm[wid][k], c0);
      if (k < SC_R1_LEN) {
        st_fr<FrP>(acc_sm[wid][SC_R0_LEN + 2 * k], c1a);
        st_fr<FrP>(acc_sm[wid][SC_R0_LEN + 2 * k + 1], c1b);
      }
    }
  }
  __syncthreads();
  if (t < SC_SLOTS) {
    F s = ld_fr<FrP>(acc_sm[0][t]);
    for (uint32_t w = 1; w < (blockDim.x >> 5); w++) s = s + ld_fr<FrP>(acc_sm[w][t]);
    st_fr<FrP>(partial + ((size_t)blockIdx.x * SC_SLOTS + t) * FrP::N, s);
  }
}

// out[slot] = sum over blocks of partial[block][slot]; one block per slot
template <class FrP>
CS_GLOBAL void k_sc_sum_partials(const uint32_t* __restrict__ partial, size_t nblocks, uint32_t* __restrict__ out) {
  typedef Fp<FrP> F;
  __shared__ __align__(16) uint32_t sm[8][FrP::N];
  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5, slot = blockIdx.x;
  F s = F::zero();
  for (size_t b = t; b < nblocks; b += blockDim.x) s = s + ld_fr<FrP>(partial + (b * SC_SLOTS + slot) * FrP::N);
  s = warp_sum(s);
  if (lane == 0) st_fr<FrP>(sm[wid], s);
  __syncthreads();
  if (t == 0) {
    F r = ld_fr<FrP>(sm[0]);
    for (uint32_t w = 1; w < (blockDim.x >> 5); w++) r = r + ld_fr<FrP>(sm[w]);
    st_fr<FrP>(out + slot * FrP::N, r);
  }
}

}  // namespace cs
