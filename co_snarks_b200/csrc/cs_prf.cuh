// On-device Rep3 mask generation: ChaCha keystream -> from_be_bytes_mod_order -> a - b.
//
// Replaces `Rep3Rand::masking_field_elements_vec` (mpc-core/src/protocols/rep3/rngs.rs:137-156) with
// RngType = rand_chacha::ChaCha12Rng (mpc-core/src/lib.rs:13): for each of the two correlated streams the
// seed is the ChaCha key, the block counter is 64-bit starting at 0, the stream id is 0 and 12 rounds
// are run; element i takes keystream words [pos + 8i, pos + 8i + 8), reads them as 32 big-endian bytes
// and reduces mod r.  The host keeps the ChaCha12Rng objects (seed + word position) and advances them
// by 8n words per call, so masks never cross PCIe.
#pragma once
#include "cs_common.cuh"
#include "cs_field.cuh"
#include "cs_ntt.cuh"  // st_fr

namespace cs {

CS_D uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define CS_QR(a, b, c, d)                     \
  a += b; d = rotl32(d ^ a, 16);              \
  c += d; b = rotl32(b ^ c, 12);              \
  a += b; d = rotl32(d ^ a, 8);               \
  c += d; b = rotl32(b ^ c, 7);

CS_D void chacha_block(const uint32_t* key, uint64_t counter, uint32_t rounds, uint32_t* out) {
  uint32_t s[16], x[16];
  s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
  CS_UNROLL
  for (int i = 0; i < 8; i++) s[4 + i] = key[i];
  s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = 0; s[15] = 0;
  CS_UNROLL
  for (int i = 0; i < 16; i++) x[i] = s[i];
  for (uint32_t r = 0; r < rounds; r += 2) {
    CS_QR(x[0], x[4], x[8], x[12]) CS_QR(x[1], x[5], x[9], x[13]) CS_QR(x[2], x[6], x[10], x[14]) CS_QR(x[3], x[7], x[11], x[15])
    CS_QR(x[0], x[5], x[10], x[15]) CS_QR(x[1], x[6], x[11], x[12]) CS_QR(x[2], x[7], x[8], x[13]) CS_QR(x[3], x[4], x[9], x[14])
  }
  CS_UNROLL
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

CS_D uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

// 8 keystream words starting at word position `w` -> field element (Montgomery) of the big-endian value mod r
template <class FrP>
CS_D Fp<FrP> prf_field_element(const uint32_t* key, uint64_t w, uint32_t rounds) {
  uint32_t blk[16], words[8];
  uint64_t b0 = w >> 4;
  uint32_t off = (uint32_t)(w & 15);
  chacha_block(key, b0, rounds, blk);
  uint32_t got = 0;
  for (uint32_t k = off; k < 16 && got < 8; k++) words[got++] = blk[k];
  if (got < 8) {
    chacha_block(key, b0 + 1, rounds, blk);
    for (uint32_t k = 0; got < 8; k++) words[got++] = blk[k];
  }
  // bytes of the keystream read big-endian: least significant limb = byte-swapped last word
  Fp<FrP> v;
  CS_UNROLL
  for (int j = 0; j < 8; j++) v.l[j] = bswap32(words[7 - j]);
  // v < 2^256 is not reduced.  R^2 * v with R^2 as the multiplicand keeps every row of the word-serial
  // product inside its bounds (the unreduced operand is consumed one limb at a time) and the result
  // (R^2 v + m r) / R < 2 r is brought to [0, r) by the final subtraction: v R mod r.
  return Fp<FrP>::r2() * v;
}

// 16 keystream words starting at `w` -> field element of the 64-byte big-endian value mod r.  A 64-byte draw
// is uniform up to 2^-256 -- the draw used for random SHARES (arithmetic::rand), where the 32-byte reduction of
// the mask vectors (statistical distance ~0.04 for BN254) would weaken the hiding of opened masked values.
// mont(hi 2^256 + lo) = mont(hi) * R + mont(lo), and a Montgomery product with R^2 multiplies by R.
template <class FrP>
CS_D Fp<FrP> prf_field_element_wide(const uint32_t* key, uint64_t w, uint32_t rounds) {
  Fp<FrP> hi = prf_field_element<FrP>(key, w, rounds), lo = prf_field_element<FrP>(key, w + 8, rounds);
  return hi * Fp<FrP>::r2() + lo;
}

// out[i] = F(stream1, pos1 + 8 i) - F(stream2, pos2 + 8 i)      (rngs.rs:103-106,137-156)
template <class FrP>
CS_GLOBAL void k_rep3_masks(const uint32_t* __restrict__ keys /* 16 words: key1 | key2 */, uint64_t pos1,
                            uint64_t pos2, uint32_t rounds, size_t n, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k1[8], k2[8];
  CS_UNROLL
  for (int j = 0; j < 8; j++) { k1[j] = keys[j]; k2[j] = keys[8 + j]; }
  Fp<FrP> a = prf_field_element<FrP>(k1, pos1 + 8 * i, rounds);
  Fp<FrP> b = prf_field_element<FrP>(k2, pos2 + 8 * i, rounds);
  st_fr<FrP>(out + i * FrP::N, a - b);
}

struct PrfKey1 { uint32_t k[8]; };  // one ChaCha key, passed by value

// ---- uniform field elements by rejection sampling (ark-ff Fp::rand) on private sub-streams -----------------
// ChaCha12Rng::from_seed(seed) with set_stream(stream): state words 14, 15 carry the 64-bit stream id, the block
// counter starts at 0.  Every element owns one stream, so n elements are drawn in parallel and each draw is exact
// rejection sampling: 8 keystream words -> 4 u64 limbs, top limb masked to the modulus width, accepted iff < r;
// the accepted limbs ARE the Montgomery representation (as in ark-ff).
CS_D void chacha_block_stream(const uint32_t* key, uint64_t counter, uint64_t stream, uint32_t rounds, uint32_t* out) {
  uint32_t s[16], x[16];
  s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
  CS_UNROLL
  for (int i = 0; i < 8; i++) s[4 + i] = key[i];
  s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = (uint32_t)stream; s[15] = (uint32_t)(stream >> 32);
  CS_UNROLL
  for (int i = 0; i < 16; i++) x[i] = s[i];
  for (uint32_t r = 0; r < rounds; r += 2) {
    CS_QR(x[0], x[4], x[8], x[12]) CS_QR(x[1], x[5], x[9], x[13]) CS_QR(x[2], x[6], x[10], x[14]) CS_QR(x[3], x[7], x[11], x[15])
    CS_QR(x[0], x[5], x[10], x[15]) CS_QR(x[1], x[6], x[11], x[12]) CS_QR(x[2], x[7], x[8], x[13]) CS_QR(x[3], x[4], x[9], x[14])
  }
  CS_UNROLL
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

template <class FrP>
CS_D Fp<FrP> prf_fr_rand_stream(const uint32_t* key, uint64_t stream, uint32_t rounds, uint32_t modulus_bits) {
  static_assert(FrP::N == 8, "256-bit scalar fields");
  const uint32_t top_mask = modulus_bits >= 256 ? 0xffffffffu : (0xffffffffu >> (256 - modulus_bits));
  for (uint64_t counter = 0;; counter++) {
    uint32_t blk[16];
    chacha_block_stream(key, counter, stream, rounds, blk);
    for (int half = 0; half < 2; half++) {
      Fp<FrP> v;
      CS_UNROLL
      for (int j = 0; j < 8; j++) v.l[j] = blk[8 * half + j];
      v.l[7] &= top_mask;
      // v < r ?  (subtract with borrow)
      uint32_t t = sub_cc(v.l[0], FrP::mod(0));
      CS_UNROLL
      for (int j = 1; j < 8; j++) t = subc_cc(v.l[j], FrP::mod(j));
      const uint32_t borrow = subc(0, 0);
      (void)t;
      if (borrow) return v;
    }
  }
}

// n uniform elements (Montgomery): element i from stream stream_base + i
template <class FrP>
CS_GLOBAL void k_fr_rand(PrfKey1 key, uint64_t stream_base, uint32_t rounds, uint32_t bits, size_t n, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_fr<FrP>(out + i * FrP::N, prf_fr_rand_stream<FrP>(key.k, stream_base + i, rounds, bits));
}

// share_field_elements (mpc-core/src/protocols/rep3.rs:281-293, co-circom-types/src/lib.rs:279-382) on the device:
// a, b uniform, c = val - a - b; party 0 holds (a, c), party 1 (b, a), party 2 (c, b).  The witness never leaves
// the GPU it was produced on; each party's share vector is written where that party will read it.
template <class FrP>
CS_GLOBAL void k_share_rep3(PrfKey1 key, uint32_t rounds, uint32_t bits, const uint32_t* __restrict__ wit, size_t n,
                            uint32_t* __restrict__ out0, uint32_t* __restrict__ out1, uint32_t* __restrict__ out2) {
  constexpr int NW = FrP::N;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fp<FrP> a = prf_fr_rand_stream<FrP>(key.k, 2 * i, rounds, bits);
  const Fp<FrP> b = prf_fr_rand_stream<FrP>(key.k, 2 * i + 1, rounds, bits);
  const Fp<FrP> c = ld_fr<FrP>(wit + i * NW) - a - b;
  st_fr<FrP>(out0 + (2 * i) * NW, a); st_fr<FrP>(out0 + (2 * i + 1) * NW, c);
  st_fr<FrP>(out1 + (2 * i) * NW, b); st_fr<FrP>(out1 + (2 * i + 1) * NW, a);
  st_fr<FrP>(out2 + (2 * i) * NW, c); st_fr<FrP>(out2 + (2 * i + 1) * NW, b);
}

struct PrfKeys { uint32_t k[16]; };  // key1 | key2, passed by value

// mul_vec on large vectors as ONE kernel (mpc-core/src/protocols/rep3/arithmetic.rs:132-160):
//   z_i = a_i * b_i + mask_i          local_mul_vec, masks drawn in registers (rngs.rs:103-106)
//   out[i].a = z_i                    this party's half of the new share
//   next_out[i].b = z_i               reshare_vec: the store goes straight into the NEXT party's share vector
//                                     (peer memory over NVLink; `next_out` is an IPC-mapped pointer)
// After all three parties' kernels have completed, every `out` holds complete (a, b) shares -- no staging
// buffer, no pack/unpack pass, and the transfer overlaps the arithmetic element by element.
template <class FrP>
CS_GLOBAL void k_rep3_mul_vec_reshare(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, PrfKeys keys,
                                      uint64_t pos1, uint64_t pos2, uint32_t rounds, size_t n,
                                      uint32_t* __restrict__ out, uint32_t* __restrict__ next_out) {
  constexpr int NW = FrP::N;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<FrP> aa = ld_fr<FrP>(a + (2 * i) * NW), ab = ld_fr<FrP>(a + (2 * i + 1) * NW);
  Fp<FrP> ba = ld_fr<FrP>(b + (2 * i) * NW), bb = ld_fr<FrP>(b + (2 * i + 1) * NW);
  Fp<FrP> z = Fp<FrP>::dot2(aa, ba + bb, ab, ba);  // one reduction for both products
  if (rounds) {
    Fp<FrP> m1 = prf_field_element<FrP>(keys.k, pos1 + 8 * i, rounds);
    Fp<FrP> m2 = prf_field_element<FrP>(keys.k + 8, pos2 + 8 * i, rounds);
    z = z + (m1 - m2);
  }
  st_fr<FrP>(out + (2 * i) * NW, z);
  if (next_out) st_fr<FrP>(next_out + (2 * i + 1) * NW, z);
}

// b-halves arriving through a staging buffer (the NCCL send/recv baseline of the same step): out[i].b = recv[i]
template <class FrP>
CS_GLOBAL void k_rep3_set_b(const uint32_t* __restrict__ recv, size_t n, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_fr<FrP>(out + (2 * i + 1) * FrP::N, ld_fr<FrP>(recv + i * FrP::N));
}

// raw keystream words (test hook: RFC 7539 block vector with rounds = 20)
static CS_GLOBAL void k_chacha_keystream(const uint32_t* __restrict__ key, uint64_t first_block, uint32_t rounds,
                                         uint32_t nblocks, uint32_t* __restrict__ out) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  uint32_t k[8], blk[16];
  for (int j = 0; j < 8; j++) k[j] = key[j];
  chacha_block(k, first_block + b, rounds, blk);
  for (int j = 0; j < 16; j++) out[(size_t)b * 16 + j] = blk[j];
}

}  // namespace cs
