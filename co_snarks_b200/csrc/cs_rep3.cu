// cs_net (mpc_net::Network replaced on-box) and cs_rep3_state (Rep3State's correlated randomness).
// See include/cosnarks_gpu.h for the contract and cs_net.h for the mailbox layout.
#include <chrono>
#include <thread>
#if !defined(CS_EMU)
#include <sys/random.h>
#endif
#include <stdio.h>
#include "cs_lib.cuh"
#include "cs_net.h"

using namespace cs;

namespace {

double net_timeout_s() {
  static double t = -1;
  if (t < 0) {
    const char* e = getenv("CS_NET_TIMEOUT_S");
    t = e ? atof(e) : 120.0;
    if (t <= 0) t = 120.0;
  }
  return t;
}

struct Deadline {
  std::chrono::steady_clock::time_point end;
  unsigned spins = 0;
  Deadline() : end(std::chrono::steady_clock::now() + std::chrono::microseconds((long long)(net_timeout_s() * 1e6))) {}
  bool expired() {
    if (++spins > 64) std::this_thread::yield();  // co-operate when parties share cores / one GPU
    return (spins & 255) == 0 && std::chrono::steady_clock::now() > end;
  }
};

#if defined(CS_EMU)
constexpr cudaMemcpyKind kAny = cudaMemcpyHostToDevice;  // the emulation's memcpy ignores the kind
#else
constexpr cudaMemcpyKind kAny = cudaMemcpyDefault;
#endif

// One chunk towards `to` if a credit is free: 1 = sent, 0 = would block, < 0 = error.
int try_send_chunk(cs_net* net, int to, const uint8_t* data, size_t bytes, size_t& off) {
  NetChannel* dst = net->peer_box[to] + net->id;  // my channel inside the receiver's mailbox
  const uint64_t c = net->send_seq[to];
  // credit: at most NET_SLOTS chunks in flight towards `to`
  if (c - net->acked[to] >= NET_SLOTS) {
    CS_CUDA(cudaMemcpyAsync(&net->h_ack[net->n + to], &net->d_box[to].ack, 8, cudaMemcpyDeviceToHost, net->st));
    CS_CUDA(cudaStreamSynchronize(net->st));
    net->acked[to] = net->h_ack[net->n + to];
    if (c - net->acked[to] >= NET_SLOTS) return 0;
  }
  const size_t len = bytes - off < NET_CHUNK ? bytes - off : NET_CHUNK;
  NetSlot* s = &dst->slot[c % NET_SLOTS];
  if (len) memcpy(net->h_send->payload, data + off, len);
  net->h_send->seq = c + 1;
  net->h_send->len = len;
  // payload first, then the header that publishes it: two stream-ordered copies into the peer's HBM
  if (len) CS_CUDA(cudaMemcpyAsync(s->payload, net->h_send->payload, len, kAny, net->st));
  CS_CUDA(cudaMemcpyAsync(&s->seq, &net->h_send->seq, 16, kAny, net->st));
  CS_CUDA(cudaStreamSynchronize(net->st));
  net->send_seq[to] = c + 1;
  off += len;
  return 1;
}

// The next chunk from `from` if it has arrived: 1 = received, 0 = nothing yet, < 0 = error.
int try_recv_chunk(cs_net* net, int from, uint8_t* data, size_t bytes, size_t& off) {
  NetChannel* ch = net->d_box + from;
  const uint64_t c = net->recv_seq[from];
  NetSlot* s = &ch->slot[c % NET_SLOTS];
  const size_t want = bytes - off < NET_CHUNK ? bytes - off : NET_CHUNK;
  // the header first: it is the last thing the sender wrote, so a valid header implies a complete payload;
  // fetching both in one copy could pair a fresh header with stale payload bytes
  CS_CUDA(cudaMemcpyAsync(&net->h_recv->seq, &s->seq, 16, cudaMemcpyDeviceToHost, net->st));
  CS_CUDA(cudaStreamSynchronize(net->st));
  if (net->h_recv->seq != c + 1) return 0;
  if (net->h_recv->len != want)
    return fail(CS_ERR_STATE, "cs_net: party %d expected %zu bytes from party %d, got %llu", net->id, want, from,
                (unsigned long long)net->h_recv->len);
  if (want) {
    CS_CUDA(cudaMemcpyAsync(net->h_recv->payload, s->payload, want, cudaMemcpyDeviceToHost, net->st));
    CS_CUDA(cudaStreamSynchronize(net->st));
    memcpy(data + off, net->h_recv->payload, want);
  }
  net->recv_seq[from] = c + 1;
  off += want;
  // acknowledge into the sender's mailbox (its channel for me): frees one of its credits
  net->h_ack[from] = c + 1;
  CS_CUDA(cudaMemcpyAsync(&net->peer_box[from][net->id].ack, &net->h_ack[from], 8, kAny, net->st));
  return 1;
}

int peer_send(cs_net* net, int to, const uint8_t* data, size_t bytes) {
  size_t off = 0;
  bool first = true;
  Deadline dl;
  while (first || off < bytes) {
    int rc = try_send_chunk(net, to, data, bytes, off);
    if (rc < 0) return rc;
    if (rc) { first = false; dl = Deadline(); continue; }
    if (dl.expired()) return fail(CS_ERR_STATE, "cs_net: party %d timed out waiting for credits from party %d", net->id, to);
  }
  return 0;
}

int peer_recv(cs_net* net, int from, uint8_t* data, size_t bytes) {
  size_t off = 0;
  bool first = true;
  Deadline dl;
  while (first || off < bytes) {
    int rc = try_recv_chunk(net, from, data, bytes, off);
    if (rc < 0) return rc;
    if (rc) { first = false; dl = Deadline(); continue; }
    if (dl.expired()) return fail(CS_ERR_STATE, "cs_net: party %d timed out waiting for a message from party %d", net->id, from);
  }
  return 0;
}

// Send to one party and receive from another with both directions making progress chunk by chunk: an all-to-all of
// messages larger than the credit window (NET_SLOTS * NET_CHUNK) cannot dead-lock on everybody sending first.
int peer_sendrecv(cs_net* net, int to, const uint8_t* sdata, size_t sbytes, int from, uint8_t* rdata, size_t rbytes) {
  size_t soff = 0, roff = 0;
  bool sfirst = true, rfirst = true;
  Deadline dl;
  while (sfirst || rfirst || soff < sbytes || roff < rbytes) {
    bool progress = false;
    if (sfirst || soff < sbytes) {
      int rc = try_send_chunk(net, to, sdata, sbytes, soff);
      if (rc < 0) return rc;
      if (rc) { sfirst = false; progress = true; }
    }
    if (rfirst || roff < rbytes) {
      int rc = try_recv_chunk(net, from, rdata, rbytes, roff);
      if (rc < 0) return rc;
      if (rc) { rfirst = false; progress = true; }
    }
    if (progress) dl = Deadline();
    else if (dl.expired())
      return fail(CS_ERR_STATE, "cs_net: party %d timed out in sendrecv (to %d: %zu of %zu, from %d: %zu of %zu bytes)", net->id, to,
                  soff, sbytes, from, roff, rbytes);
  }
  return 0;
}

}  // namespace

extern "C" {

int cs_os_random(uint8_t* out, size_t bytes) {
  if (!out) return fail(CS_ERR_ARG, "cs_os_random: NULL argument");
#if defined(CS_EMU)
  FILE* f = fopen("/dev/urandom", "rb");
  if (!f || fread(out, 1, bytes, f) != bytes) { if (f) fclose(f); return fail(CS_ERR_STATE, "cs_os_random: /dev/urandom unavailable"); }
  fclose(f);
#else
  size_t off = 0;
  while (off < bytes) {
    ssize_t k = getrandom(out + off, bytes - off, 0);
    if (k < 0) return fail(CS_ERR_STATE, "cs_os_random: getrandom failed");
    off += (size_t)k;
  }
#endif
  return 0;
}

int cs_net_from_callbacks(int id, int n_parties, const cs_net_callbacks* cb, cs_net** out) {
  if (!cb || !cb->send || !cb->recv || !out) return fail(CS_ERR_ARG, "cs_net_from_callbacks: NULL argument");
  if (n_parties < 2 || n_parties > NET_MAX_PARTIES || id < 0 || id >= n_parties)
    return fail(CS_ERR_ARG, "cs_net_from_callbacks: bad party id %d of %d", id, n_parties);
  cs_net* n = new cs_net();
  n->id = id; n->n = n_parties; n->is_cb = true; n->cb = *cb; n->connected = true;
  *out = n;
  return 0;
}

int cs_net_peer_create(cs_ctx* ctx, int id, int n_parties, cs_net** out) {
  if (!ctx || !out) return fail(CS_ERR_ARG, "cs_net_peer_create: NULL argument");
  if (n_parties < 2 || n_parties > NET_MAX_PARTIES || id < 0 || id >= n_parties)
    return fail(CS_ERR_ARG, "cs_net_peer_create: bad party id %d of %d", id, n_parties);
  CS_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<cs_net> n(new cs_net());
  n->id = id; n->n = n_parties; n->device = ctx->device;
  CS_CUDA(cudaStreamCreateWithFlags(&n->st, cudaStreamNonBlocking));
  // cudaMalloc (not a pool): the allocation must be exportable through CUDA IPC
  CS_CUDA(cudaMalloc((void**)&n->d_box, sizeof(NetChannel) * n_parties));
  CS_CUDA(cudaMemsetAsync(n->d_box, 0, sizeof(NetChannel) * n_parties, n->st));
  CS_CUDA(cudaStreamSynchronize(n->st));
  CS_CUDA(cudaMallocHost((void**)&n->h_send, sizeof(NetSlot)));
  CS_CUDA(cudaMallocHost((void**)&n->h_recv, sizeof(NetSlot)));
  CS_CUDA(cudaMallocHost((void**)&n->h_ack, 8 * 2 * n_parties));
  memset(n->h_ack, 0, 8 * 2 * n_parties);
  *out = n.release();
  return 0;
}

int cs_net_peer_handle(cs_net* net, uint8_t* out_handle64) {
  if (!net || net->is_cb || !out_handle64) return fail(CS_ERR_ARG, "cs_net_peer_handle: not a peer-mailbox net");
  cudaIpcMemHandle_t h;
  CS_CUDA(cudaSetDevice(net->device));
  CS_CUDA(cudaIpcGetMemHandle(&h, net->d_box));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(out_handle64, &h, 64);
  return 0;
}

int cs_net_peer_connect(cs_net* net, const uint8_t* handles) {
  if (!net || net->is_cb || !handles) return fail(CS_ERR_ARG, "cs_net_peer_connect: bad argument");
  CS_CUDA(cudaSetDevice(net->device));
  for (int p = 0; p < net->n; p++) {
    if (p == net->id) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * p, 64);
    void* ptr = nullptr;
    CS_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    net->peer_box[p] = (NetChannel*)ptr;
    net->peer_ipc[p] = true;
  }
  net->connected = true;
  return 0;
}

int cs_net_peer_connect_local(cs_net* net, cs_net* const* peers) {
  if (!net || net->is_cb || !peers) return fail(CS_ERR_ARG, "cs_net_peer_connect_local: bad argument");
  CS_CUDA(cudaSetDevice(net->device));
  for (int p = 0; p < net->n; p++) {
    if (p == net->id) continue;
    if (!peers[p] || peers[p]->is_cb || peers[p]->n != net->n || peers[p]->id != p)
      return fail(CS_ERR_ARG, "cs_net_peer_connect_local: peer %d is not the mailbox net of party %d", p, p);
#if !defined(CS_EMU)
    if (peers[p]->device != net->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[p]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(CS_ERR_CUDA, "cs_net_peer_connect_local: no peer access %d -> %d (%s)", net->device, peers[p]->device,
                    cudaGetErrorString(e));
      cudaGetLastError();
    }
#endif
    net->peer_box[p] = peers[p]->d_box;
    net->peer_ipc[p] = false;
  }
  net->connected = true;
  return 0;
}

int cs_net_send(cs_net* net, int to, const void* data, size_t bytes) {
  if (!net || (!data && bytes)) return fail(CS_ERR_ARG, "cs_net_send: NULL argument");
  if (to < 0 || to >= net->n || to == net->id) return fail(CS_ERR_ARG, "cs_net_send: bad destination %d", to);
  net->bytes_sent += bytes;
  if (net->is_cb) {
    int rc = net->cb.send(net->cb.user, to, data, bytes);
    return rc ? fail(CS_ERR_STATE, "cs_net_send: transport callback failed (%d)", rc) : 0;
  }
  if (!net->connected) return fail(CS_ERR_STATE, "cs_net_send: mailbox net is not connected");
  CS_CUDA(cudaSetDevice(net->device));
  return peer_send(net, to, (const uint8_t*)data, bytes);
}

int cs_net_recv(cs_net* net, int from, void* data, size_t bytes) {
  if (!net || (!data && bytes)) return fail(CS_ERR_ARG, "cs_net_recv: NULL argument");
  if (from < 0 || from >= net->n || from == net->id) return fail(CS_ERR_ARG, "cs_net_recv: bad source %d", from);
  if (net->is_cb) {
    int rc = net->cb.recv(net->cb.user, from, data, bytes);
    return rc ? fail(CS_ERR_STATE, "cs_net_recv: transport callback failed (%d)", rc) : 0;
  }
  if (!net->connected) return fail(CS_ERR_STATE, "cs_net_recv: mailbox net is not connected");
  CS_CUDA(cudaSetDevice(net->device));
  return peer_recv(net, from, (uint8_t*)data, bytes);
}

int cs_net_sendrecv(cs_net* net, int to, const void* sdata, size_t sbytes, int from, void* rdata, size_t rbytes) {
  if (!net || (!sdata && sbytes) || (!rdata && rbytes)) return fail(CS_ERR_ARG, "cs_net_sendrecv: NULL argument");
  if (to < 0 || to >= net->n || to == net->id) return fail(CS_ERR_ARG, "cs_net_sendrecv: bad destination %d", to);
  if (from < 0 || from >= net->n || from == net->id) return fail(CS_ERR_ARG, "cs_net_sendrecv: bad source %d", from);
  if (net->is_cb) {  // the callback transport queues its sends (mpc_net::Network::send does not wait for the receiver)
    CS_TRY(cs_net_send(net, to, sdata, sbytes));
    return cs_net_recv(net, from, rdata, rbytes);
  }
  if (!net->connected) return fail(CS_ERR_STATE, "cs_net_sendrecv: mailbox net is not connected");
  net->bytes_sent += sbytes;
  CS_CUDA(cudaSetDevice(net->device));
  return peer_sendrecv(net, to, (const uint8_t*)sdata, sbytes, from, (uint8_t*)rdata, rbytes);
}

uint64_t cs_net_bytes_sent(const cs_net* net) { return net ? net->bytes_sent : 0; }

void cs_net_free(cs_net* net) {
  if (!net) return;
  if (!net->is_cb) {
    cudaSetDevice(net->device);
    if (net->st) cudaStreamSynchronize(net->st);
    for (int p = 0; p < net->n; p++)
      if (net->peer_ipc[p] && net->peer_box[p]) cudaIpcCloseMemHandle(net->peer_box[p]);
    if (net->d_box) cudaFree(net->d_box);
    if (net->h_send) cudaFreeHost(net->h_send);
    if (net->h_recv) cudaFreeHost(net->h_recv);
    if (net->h_ack) cudaFreeHost(net->h_ack);
    if (net->st) cudaStreamDestroy(net->st);
  }
  delete net;
}

// ---- Rep3State -----------------------------------------------------------------------------------
int cs_rep3_state_create(cs_net* net, cs_rep3_state** out) {
  if (!net || !out) return fail(CS_ERR_ARG, "cs_rep3_state_create: NULL argument");
  if (net->n != 3) return fail(CS_ERR_ARG, "cs_rep3_state_create: Rep3 needs a 3-party net");
  uint8_t seed1[32], seed2[32];
  CS_TRY(cs_os_random(seed1, 32));  // ChaCha12Rng::from_entropy -> seed1 (rep3.rs:57,71-72)
  Rep3Net rn(net);
  CS_TRY(rn.reshare(seed1, seed2, 32));  // seed2 = net.reshare(seed1) (rep3.rs:73)
  cs_rep3_state* st = new cs_rep3_state();
  st->id = net->id;
  st->rng1.init(seed1, 0);
  st->rng2.init(seed2, 0);
  *out = st;
  return 0;
}

int cs_rep3_state_from_seeds(int party, const uint8_t* own, uint64_t pos_own, const uint8_t* prev, uint64_t pos_prev,
                             cs_rep3_state** out) {
  if (!own || !prev || !out) return fail(CS_ERR_ARG, "cs_rep3_state_from_seeds: NULL argument");
  if (party < 0 || party > 2) return fail(CS_ERR_ARG, "cs_rep3_state_from_seeds: party must be 0..2");
  cs_rep3_state* st = new cs_rep3_state();
  st->id = party;
  st->rng1.init(own, pos_own);
  st->rng2.init(prev, pos_prev);
  *out = st;
  return 0;
}

int cs_rep3_state_fork(cs_rep3_state* st, cs_rep3_state** out) {
  if (!st || !out) return fail(CS_ERR_ARG, "cs_rep3_state_fork: NULL argument");
  uint8_t s1[32], s2[32];
  st->rng1.gen_seed(s1);  // Rep3Rand::fork -> random_seeds (rngs.rs:99-103,233-237)
  st->rng2.gen_seed(s2);
  return cs_rep3_state_from_seeds(st->id, s1, 0, s2, 0, out);
}

int cs_rep3_state_prf(const cs_rep3_state* st, cs_rep3_prf* out) {
  if (!st || !out) return fail(CS_ERR_ARG, "cs_rep3_state_prf: NULL argument");
  memcpy(out->seed1, st->rng1.seed, 32);
  memcpy(out->seed2, st->rng2.seed, 32);
  out->word_pos1 = st->rng1.pos;
  out->word_pos2 = st->rng2.pos;
  out->rounds = 12;
  return 0;
}

int cs_rep3_state_advance(cs_rep3_state* st, uint64_t nwords) {
  if (!st) return fail(CS_ERR_ARG, "cs_rep3_state_advance: NULL argument");
  st->rng1.pos += nwords;
  st->rng2.pos += nwords;
  return 0;
}

int cs_rep3_state_rand(cs_rep3_state* st, cs_curve curve, uint64_t* out_share) {
  if (!st || !out_share) return fail(CS_ERR_ARG, "cs_rep3_state_rand: NULL argument");
  switch ((int)curve) {
    case CS_BN254:
      st->rng1.fr_rand<Bn254Fr>(out_share, 254);
      st->rng2.fr_rand<Bn254Fr>(out_share + 4, 254);
      return 0;
#if defined(CS_ENABLE_BLS12_381)
    case CS_BLS12_381:
      st->rng1.fr_rand<Bls381Fr>(out_share, 255);
      st->rng2.fr_rand<Bls381Fr>(out_share + 4, 255);
      return 0;
#endif
    default: return fail(CS_ERR_ARG, "unsupported curve");
  }
}

void cs_rep3_state_free(cs_rep3_state* st) { delete st; }

}  // extern "C"
