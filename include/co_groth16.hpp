// co_groth16.hpp -- C++17 host-side mirror of the reference's Groth16 prover interface, over the C ABI of
// libcosnarks_gpu.so (include/cosnarks_gpu.h).  Header-only.
//
// The reference's host code is Rust (no toolchain in this image), so the host side above the ABI is
// written in C++ with the reference's names, argument meaning and error behaviour:
//
//   reference (co-circom/co-groth16/src)                      here (namespace co_groth16)
//   ------------------------------------------------------    ----------------------------------------
//   ark_groth16::ProvingKey + ConstraintMatrices (lib.rs)      ProvingKey  (device-resident, cs_groth16_pk)
//   co_circom_types::SharedWitness<F, S>                       SharedWitness<Share>
//   trait R1CSToQAP / CircomReduction (reduction.rs:27-193)    CircomReduction::witness_map_from_matrices
//   trait CircomGroth16Prover (mpc.rs:22-138)                  PlainGroth16Driver / Rep3Groth16Driver
//   CoGroth16<P, T>::prove_inner (groth16.rs:125-177)          CoGroth16<Driver>::prove_inner
//   Groth16::plain_prove (groth16.rs:484-490)                  Groth16::plain_prove
//   Rep3CoGroth16::prove (groth16.rs:360-379)                  Rep3CoGroth16::prove(net0, net1, pk, witness)
//   mpc_net::Network (mpc-net/src/lib.rs:34-63)                mpc_net::Network (id/send/recv)
//   mpc_net::local::LocalNetwork::new_3_parties (local.rs)     mpc_net::LocalNetwork::new_3_parties
//   Rep3State / Rep3Rand (rep3.rs:43-128, rngs.rs:86-156)      Rep3State (cs_rep3_state: two ChaCha12 streams, OS-entropy seeds)
//   eyre::Result / bail!                                       std::runtime_error with the same messages
//
// tests/cpp/test_co_groth16.cpp drives this exactly like tests/tests/circom/e2e_tests/rep3.rs drives
// the reference (three party threads over LocalNetwork, all proofs equal).
#pragma once
#include <array>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include "cosnarks_gpu.h"

namespace co_groth16 {

using Fr = std::array<uint64_t, 4>;   // BN254 scalar, Montgomery limbs (ark Fr layout)
using G1 = std::array<uint64_t, 8>;   // affine x || y, Montgomery; all-zero = infinity
using G2 = std::array<uint64_t, 16>;  // affine x.c0 x.c1 y.c0 y.c1

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("cosnarks_gpu: ") + cs_last_error());
}

struct Context {
  cs_ctx* h = nullptr;
  explicit Context(int device = 0) { check(cs_ctx_create(device, nullptr, &h)); }
  ~Context() { cs_ctx_destroy(h); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
};

// ---- single-point / single-element helpers (latency-only work, as in the reference) ----------------
inline G1 operator+(const G1& a, const G1& b) { G1 r; check(cs_point_add(CS_BN254, CS_G1, a.data(), b.data(), r.data())); return r; }
inline G2 operator+(const G2& a, const G2& b) { G2 r; check(cs_point_add(CS_BN254, CS_G2, a.data(), b.data(), r.data())); return r; }
inline G1 operator-(const G1& a) { G1 r; check(cs_point_neg(CS_BN254, CS_G1, a.data(), r.data())); return r; }
inline G1 operator*(const G1& p, const Fr& s) { G1 r; check(cs_point_scalar_mul(CS_BN254, CS_G1, p.data(), s.data(), r.data())); return r; }
inline G2 operator*(const G2& p, const Fr& s) { G2 r; check(cs_point_scalar_mul(CS_BN254, CS_G2, p.data(), s.data(), r.data())); return r; }
inline Fr fr_mul(const Fr& a, const Fr& b) { Fr r; check(cs_fr_mul(CS_BN254, a.data(), b.data(), r.data())); return r; }
inline Fr fr_add(const Fr& a, const Fr& b) { Fr r; check(cs_fr_add(CS_BN254, a.data(), b.data(), r.data())); return r; }
inline Fr fr_sub(const Fr& a, const Fr& b) { Fr r; check(cs_fr_sub(CS_BN254, a.data(), b.data(), r.data())); return r; }
inline Fr fr_from_canonical(const Fr& c) { Fr r; check(cs_fr_to_mont(CS_BN254, c.data(), r.data(), 1)); return r; }

// ---- data carriers -----------------------------------------------------------------------------
struct Rep3PrimeFieldShare { Fr a, b; };               // rep3/arithmetic/types.rs:21-28
template <class Share>
struct SharedWitness {                                  // co-circom-types/src/lib.rs:207-219
  std::vector<Fr> public_inputs;                        // includes the leading 1
  std::vector<Share> witness;
};
struct Proof { G1 a; G2 b; G1 c; };

// ark_groth16::ProvingKey + ConstraintMatrices, uploaded once (cs_groth16_pk_create)
struct ProvingKey {
  cs_groth16_pk* h = nullptr;
  size_t num_instance_variables = 0, num_witness_variables = 0;
  G1 delta_g1{};
  ProvingKey(Context& ctx, const cs_groth16_key_desc& d) {
    check(cs_groth16_pk_create(ctx.h, &d, &h));
    num_instance_variables = d.num_instance_variables;
    num_witness_variables = d.num_witness_variables;
    std::memcpy(delta_g1.data(), d.delta_g1, sizeof(G1));
  }
  ~ProvingKey() { cs_groth16_pk_free(h); }
  ProvingKey(const ProvingKey&) = delete;
  size_t domain_size() const { return cs_groth16_domain_size(h); }
};

// the length checks of prove_inner (groth16.rs:134-149), same messages
template <class Share>
inline void check_witness_lengths(const ProvingKey& pk, const SharedWitness<Share>& w) {
  if (w.public_inputs.size() != pk.num_instance_variables)
    throw std::runtime_error("amount of public inputs does not match with provided constraint system! Expected " +
                             std::to_string(pk.num_instance_variables) + ", but got " + std::to_string(w.public_inputs.size()));
  if (w.witness.size() != pk.num_witness_variables)
    throw std::runtime_error("amount of private witness variables does not match with provided constraint system! Expected " +
                             std::to_string(pk.num_witness_variables) + ", but got " + std::to_string(w.witness.size()));
}

}  // namespace co_groth16

// ---- mpc_net::Network + LocalNetwork -------------------------------------------------------------
namespace mpc_net {

struct Network {                                         // mpc-net/src/lib.rs:34-63
  virtual ~Network() = default;
  virtual size_t id() const = 0;
  virtual void send(size_t to, const std::vector<uint8_t>& data) = 0;
  virtual std::vector<uint8_t> recv(size_t from) = 0;
};

// In-process 3-party mesh with per-pair FIFO queues (mpc-net/src/local.rs:22-64 uses crossbeam channels).
class LocalNetwork : public Network {
  struct Chan {
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::vector<uint8_t>> q;
  };
  struct Mesh { Chan ch[3][3]; };
  std::shared_ptr<Mesh> mesh_;
  size_t id_;
  LocalNetwork(std::shared_ptr<Mesh> m, size_t id) : mesh_(std::move(m)), id_(id) {}

 public:
  static std::array<std::unique_ptr<LocalNetwork>, 3> new_3_parties() {
    auto m = std::make_shared<Mesh>();
    return {std::unique_ptr<LocalNetwork>(new LocalNetwork(m, 0)), std::unique_ptr<LocalNetwork>(new LocalNetwork(m, 1)),
            std::unique_ptr<LocalNetwork>(new LocalNetwork(m, 2))};
  }
  size_t id() const override { return id_; }
  void send(size_t to, const std::vector<uint8_t>& data) override {
    Chan& c = mesh_->ch[id_][to];
    { std::lock_guard<std::mutex> lk(c.m); c.q.push_back(data); }
    c.cv.notify_one();
  }
  std::vector<uint8_t> recv(size_t from) override {
    Chan& c = mesh_->ch[from][id_];
    std::unique_lock<std::mutex> lk(c.m);
    c.cv.wait(lk, [&] { return !c.q.empty(); });
    auto v = std::move(c.q.front());
    c.q.pop_front();
    return v;
  }
};

}  // namespace mpc_net

namespace co_groth16 {

// ---- Rep3 protocol state (mpc-core/src/protocols/rep3) -----------------------------------------
struct PartyID {                                         // rep3/id.rs:9-47
  size_t v;
  size_t next() const { return (v + 1) % 3; }
  size_t prev() const { return (v + 2) % 3; }
};

template <class T>
inline std::vector<uint8_t> bytes_of(const T& x) { std::vector<uint8_t> b(sizeof(T)); std::memcpy(b.data(), &x, sizeof(T)); return b; }
template <class T>
inline T from_bytes(const std::vector<uint8_t>& b) {
  if (b.size() != sizeof(T)) throw std::runtime_error("During execution of mul_vec in MPC: Invalid number of elements received");
  T x; std::memcpy(&x, b.data(), sizeof(T)); return x;
}
// Rep3NetworkExt (rep3/network.rs:30-79)
template <class T> inline T reshare(mpc_net::Network& net, const T& x) {
  PartyID id{net.id()};
  net.send(id.next(), bytes_of(x));
  return from_bytes<T>(net.recv(id.prev()));
}
template <class T> inline std::pair<T, T> broadcast(mpc_net::Network& net, const T& x) {
  PartyID id{net.id()};
  net.send(id.next(), bytes_of(x));
  net.send(id.prev(), bytes_of(x));
  T p = from_bytes<T>(net.recv(id.prev()));
  T n = from_bytes<T>(net.recv(id.next()));
  return {p, n};
}

// uniform field element from the OS entropy pool (PlainGroth16Driver::rand = thread_rng, mpc/plain.rs:23-26):
// rejection sampling on 254-bit draws, like ark-ff's Fp::rand; the accepted limbs are taken as the Montgomery form
inline Fr fr_rand() {
  static const Fr R = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  for (;;) {
    Fr v;
    check(cs_os_random(reinterpret_cast<uint8_t*>(v.data()), 32));
    v[3] &= (1ull << 62) - 1;
    for (int i = 3; i >= 0; i--) {
      if (v[i] < R[i]) return v;
      if (v[i] > R[i]) break;
    }
  }
}

// mpc_net::Network -> the library's transport handle (cs_net over callbacks).  `send` queues, `recv` blocks.
struct NetAdapter {
  mpc_net::Network& net;
  cs_net* h = nullptr;
  explicit NetAdapter(mpc_net::Network& n) : net(n) {
    cs_net_callbacks cb{this, &NetAdapter::send_cb, &NetAdapter::recv_cb};
    check(cs_net_from_callbacks((int)n.id(), 3, &cb, &h));
  }
  ~NetAdapter() { cs_net_free(h); }
  NetAdapter(const NetAdapter&) = delete;
  static int send_cb(void* u, int to, const void* data, size_t bytes) {
    try {
      auto* b = static_cast<const uint8_t*>(data);
      static_cast<NetAdapter*>(u)->net.send((size_t)to, std::vector<uint8_t>(b, b + bytes));
      return 0;
    } catch (...) { return -1; }
  }
  static int recv_cb(void* u, int from, void* data, size_t bytes) {
    try {
      auto v = static_cast<NetAdapter*>(u)->net.recv((size_t)from);
      if (v.size() != bytes) return -2;
      std::memcpy(data, v.data(), bytes);
      return 0;
    } catch (...) { return -1; }
  }
};

// Rep3State (rep3.rs:43-75): Rep3Rand's two ChaCha12 streams live inside the library.  new(): seed1 from the OS
// entropy pool (ChaCha12Rng::from_entropy), seed2 = net.reshare(seed1) -- exactly setup_prf.
struct Rep3State {
  cs_rep3_state* h = nullptr;
  explicit Rep3State(NetAdapter& net) { check(cs_rep3_state_create(net.h, &h)); }
  ~Rep3State() { cs_rep3_state_free(h); }
  Rep3State(const Rep3State&) = delete;
  Rep3PrimeFieldShare rand() {                                                                   // arithmetic.rs:357-360
    Rep3PrimeFieldShare s;
    uint64_t ab[8];
    check(cs_rep3_state_rand(h, CS_BN254, ab));
    std::memcpy(s.a.data(), ab, 32);
    std::memcpy(s.b.data(), ab + 4, 32);
    return s;
  }
};

// ---- provers ----------------------------------------------------------------------------------------
struct Groth16 {
  // Groth16::plain_prove::<CircomReduction>(pkey, matrices, witness)  (groth16.rs:484-490); r, s as drawn by
  // PlainGroth16Driver::rand (mpc/plain.rs:23-26) unless injected
  static Proof plain_prove(Context& ctx, ProvingKey& pk, const SharedWitness<Fr>& w, const Fr* r = nullptr, const Fr* s = nullptr) {
    check_witness_lengths(pk, w);
    Fr rr = r ? *r : fr_rand(), ss = s ? *s : fr_rand();
    Proof p;
    check(cs_groth16_prove_plain(ctx.h, pk.h, w.public_inputs[0].data(), w.witness.empty() ? nullptr : w.witness[0].data(),
                                 rr.data(), ss.data(), p.a.data(), p.b.data(), p.c.data()));
    return p;
  }
};

struct Rep3CoGroth16 {
  // Rep3CoGroth16::prove::<N, CircomReduction>(net0, net1, &pkey, &matrices, witness)  (groth16.rs:360-379):
  // state0 = Rep3State::new(net0), state1 = state0.fork(0), then prove_inner -- the local phase on this party's GPU
  // and create_proof_with_assignment's two network legs (groth16.rs:296-337), all inside cs_groth16_rep3_prove.
  static Proof prove(Context& ctx, mpc_net::Network& net0, mpc_net::Network& net1, ProvingKey& pk,
                     const SharedWitness<Rep3PrimeFieldShare>& w, Rep3PrimeFieldShare* out_r = nullptr,
                     Rep3PrimeFieldShare* out_s = nullptr) {
    check_witness_lengths(pk, w);
    NetAdapter n0(net0), n1(net1);
    Rep3State state0(n0);
    Proof p;
    uint64_t rs[16];
    check(cs_groth16_rep3_prove(ctx.h, pk.h, n0.h, n1.h, state0.h, w.public_inputs[0].data(),
                                w.witness.empty() ? nullptr : w.witness[0].a.data(), nullptr, p.a.data(), p.b.data(),
                                p.c.data(), rs));
    if (out_r) { std::memcpy(out_r->a.data(), rs, 32); std::memcpy(out_r->b.data(), rs + 4, 32); }
    if (out_s) { std::memcpy(out_s->a.data(), rs + 8, 32); std::memcpy(out_s->b.data(), rs + 12, 32); }
    return p;
  }
};

}  // namespace co_groth16
