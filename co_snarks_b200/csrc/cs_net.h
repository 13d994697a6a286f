// Party-to-party transport and Rep3 correlated randomness (host side of the library).
//
// cs_net   = mpc_net::Network (mpc-net/src/lib.rs:34-63) reduced to id / send / recv, either over caller
//            callbacks or over per-party mailboxes in GPU memory reached through CUDA IPC (NVLink peer copies).
// Rep3Net  = Rep3NetworkExt (mpc-core/src/protocols/rep3/network.rs:30-79): reshare / broadcast on a 3-mesh.
// cs_rep3_state = Rep3State's Rep3Rand (rep3.rs:43-75, rngs.rs:86-156): two ChaCha12Rng streams.
#pragma once
#include <stdint.h>
#include <string.h>
#include "cs_common.cuh"
#include "cs_host_field.h"
#include "../../include/cosnarks_gpu.h"

namespace cs {

constexpr int NET_MAX_PARTIES = 8;
constexpr uint32_t NET_CHUNK = 65536;  // payload bytes per mailbox slot (point-sized messages use a few hundred; vectors stream in 64 KB chunks)
constexpr uint32_t NET_SLOTS = 8;     // slots per (receiver, sender) channel

// one slot: payload, then the header that makes it valid (written by a second, later copy)
struct NetSlot {
  uint8_t payload[NET_CHUNK];
  uint64_t seq;  // chunk number + 1
  uint64_t len;  // payload bytes in this chunk
};
// channel (owner <- from): slots written by `from`; ack written by `from` too: how many of the OWNER's chunks
// to `from` have been consumed there (credit for the owner's sends)
struct NetChannel {
  NetSlot slot[NET_SLOTS];
  uint64_t ack;
  uint64_t pad[7];
};

}  // namespace cs

struct cs_net {
  int id = 0, n = 0;
  uint64_t bytes_sent = 0;
  bool is_cb = false;
  cs_net_callbacks cb{};
  // peer mailboxes
  int device = 0;
  cudaStream_t st = nullptr;
  cs::NetChannel* d_box = nullptr;                       // own mailbox: n channels
  cs::NetChannel* peer_box[cs::NET_MAX_PARTIES] = {};    // the other parties' mailboxes as seen from here
  bool peer_ipc[cs::NET_MAX_PARTIES] = {};
  cs::NetSlot* h_send = nullptr;   // pinned staging
  cs::NetSlot* h_recv = nullptr;
  uint64_t* h_ack = nullptr;       // pinned: [n] acks being sent, [n] acks read back
  uint64_t send_seq[cs::NET_MAX_PARTIES] = {}, recv_seq[cs::NET_MAX_PARTIES] = {}, acked[cs::NET_MAX_PARTIES] = {};
  bool connected = false;
};

namespace cs {

// ---- ChaCha12Rng (rand_chacha 0.3): key = seed, 64-bit block counter from 0, stream 0; `pos` counts 32-bit words
struct HostChaCha {
  uint8_t seed[32];
  uint32_t key[8];
  uint64_t pos = 0;
  void init(const uint8_t* s, uint64_t p) {
    memcpy(seed, s, 32);
    for (int i = 0; i < 8; i++) key[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    pos = p;
  }
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  void block(uint64_t counter, uint32_t* out) const {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = 0; s[15] = 0;
    uint32_t x[16];
    memcpy(x, s, sizeof(x));
#define CS_HQR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < 6; r++) {
      CS_HQR(0, 4, 8, 12) CS_HQR(1, 5, 9, 13) CS_HQR(2, 6, 10, 14) CS_HQR(3, 7, 11, 15)
      CS_HQR(0, 5, 10, 15) CS_HQR(1, 6, 11, 12) CS_HQR(2, 7, 8, 13) CS_HQR(3, 4, 9, 14)
    }
#undef CS_HQR
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
  }
  uint32_t next_u32() {
    uint32_t blk[16];
    block(pos >> 4, blk);
    return blk[(pos++) & 15];
  }
  void words(uint32_t* out, int k) { for (int i = 0; i < k; i++) out[i] = next_u32(); }
  // ark-ff Fp::rand: N u64 limbs from the rng, the top limb masked to the modulus width, rejection sampling;
  // the accepted limbs are the element's internal (Montgomery) representation.
  template <class FrP>
  void fr_rand(uint64_t* out, unsigned modulus_bits) {
    typedef host::HFp<FrP> HR;
    for (;;) {
      for (int i = 0; i < HR::N; i++) { uint64_t lo = next_u32(); uint64_t hi = next_u32(); out[i] = lo | (hi << 32); }
      const unsigned shave = 64 * HR::N - modulus_bits;
      if (shave) out[HR::N - 1] &= (~0ull) >> shave;
      if (!HR::geq_mod(out)) return;
    }
  }
  // from_be_bytes_mod_order over the next 32 keystream bytes -> Montgomery element
  template <class FrP>
  host::HFp<FrP> fr_be_mod_order() {
    typedef host::HFp<FrP> HR;
    uint32_t w[8];
    words(w, 8);
    // keystream bytes read big-endian: most significant byte = first byte of word 0
    HR v = HR::zero();
    for (int j = 0; j < 8; j++) {
      uint32_t x = w[7 - j];
      uint32_t sw = (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
      if (j & 1) v.l[j / 2] |= (uint64_t)sw << 32; else v.l[j / 2] = sw;
    }
    // v < 2^256, not reduced: v * R^2 / R = v R mod r after the product's final subtraction (one more may be needed)
    HR r = HR::r2() * v;
    return r;
  }
  // rand 0.8 Standard for [u8; 32]: every byte is `next_u32() as u8`
  void gen_seed(uint8_t* out) { for (int i = 0; i < 32; i++) out[i] = (uint8_t)next_u32(); }
};

}  // namespace cs

struct cs_rep3_state {
  int id = 0;
  cs::HostChaCha rng1, rng2;  // own stream, previous party's stream
};

namespace cs {

// Rep3NetworkExt (rep3/network.rs:30-79) on a cs_net of three parties
struct Rep3Net {
  cs_net* net;
  int id, next, prev;
  explicit Rep3Net(cs_net* n) : net(n), id(n->id), next((n->id + 1) % 3), prev((n->id + 2) % 3) {}
  int send_next(const void* d, size_t b) { return cs_net_send(net, next, d, b); }
  int send_prev(const void* d, size_t b) { return cs_net_send(net, prev, d, b); }
  int recv_prev(void* d, size_t b) { return cs_net_recv(net, prev, d, b); }
  int recv_next(void* d, size_t b) { return cs_net_recv(net, next, d, b); }
  // reshare: send to next, receive from prev
  int reshare(const void* mine, void* from_prev, size_t b) {
    CS_TRY(send_next(mine, b));
    return recv_prev(from_prev, b);
  }
};

}  // namespace cs
