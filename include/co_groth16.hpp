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
//   Rep3State / Rep3Rand (rep3.rs:43-128, rngs.rs:86-156)      Rep3State (two correlated PRF streams)
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
#include <random>
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

// Rep3Rand (rngs.rs:86-156): rng1 = own stream, rng2 = previous party's stream, seeds exchanged once
// (rep3.rs:71-110).  The reference uses ChaCha12; the correlation structure is what the protocol needs.
struct Rep3State {
  PartyID id;
  std::mt19937_64 rng1, rng2;
  Rep3State(mpc_net::Network& net, uint64_t seed) : id{net.id()} {
    uint64_t own = seed * 4 + net.id();
    uint64_t prev = reshare(net, own);
    rng1.seed(own);
    rng2.seed(prev);
  }
  static Fr draw(std::mt19937_64& g) {  // uniform 253-bit value (< r), taken as a canonical integer
    Fr c{g(), g(), g(), g() & ((1ull << 61) - 1)};
    return fr_from_canonical(c);
  }
  std::pair<Fr, Fr> random_fes() { Fr a = draw(rng1); Fr b = draw(rng2); return {a, b}; }      // rngs.rs:109-113
  Rep3PrimeFieldShare rand() { auto ab = random_fes(); return {ab.first, ab.second}; }          // arithmetic.rs:357-360
  Fr masking_field_element() { auto ab = random_fes(); return fr_sub(ab.first, ab.second); }    // rngs.rs:103-106
  std::vector<Fr> masking_field_elements_vec(size_t n) {                                         // rngs.rs:137-156
    std::vector<Fr> out(n);
    for (auto& x : out) x = masking_field_element();
    return out;
  }
  G1 masking_ec_element(const G1& generator) {                                                   // rngs.rs:177-186
    auto ab = random_fes();
    return generator * ab.first + (-(generator * ab.second));
  }
};

// ---- R1CSToQAP ------------------------------------------------------------------------------------
struct CircomReduction {                                 // groth16/reduction.rs:73-193
  // plain driver: h as field elements
  static std::vector<Fr> witness_map_from_matrices(Context& ctx, ProvingKey& pk, const std::vector<Fr>& public_inputs,
                                                   const std::vector<Fr>& private_witness) {
    std::vector<Fr> h(pk.domain_size());
    check(cs_groth16_witness_map(ctx.h, pk.h, CS_PLAIN, 0, public_inputs[0].data(),
                                 private_witness.empty() ? nullptr : private_witness[0].data(), nullptr, nullptr, h[0].data()));
    return h;
  }
  // Rep3 driver: half shares of h; consumes two mask vectors from the party's state, in the order
  // reduction.rs:160 and :182 do
  static std::vector<Fr> witness_map_from_matrices(Context& ctx, ProvingKey& pk, Rep3State& state,
                                                   const std::vector<Fr>& public_inputs,
                                                   const std::vector<Rep3PrimeFieldShare>& private_witness) {
    const size_t n = pk.domain_size();
    auto m1 = state.masking_field_elements_vec(n), m2 = state.masking_field_elements_vec(n);
    std::vector<Fr> h(n);
    check(cs_groth16_witness_map(ctx.h, pk.h, CS_REP3, (int)state.id.v, public_inputs[0].data(),
                                 private_witness.empty() ? nullptr : private_witness[0].a.data(), m1[0].data(), m2[0].data(),
                                 h[0].data()));
    return h;
  }
};

// ---- provers ----------------------------------------------------------------------------------------
struct Groth16 {
  // Groth16::plain_prove::<CircomReduction>(pkey, matrices, witness)  (groth16.rs:484-490); r, s as drawn by
  // PlainGroth16Driver::rand (mpc/plain.rs:23-26) unless injected
  static Proof plain_prove(Context& ctx, ProvingKey& pk, const SharedWitness<Fr>& w, const Fr* r = nullptr, const Fr* s = nullptr) {
    check_witness_lengths(pk, w);
    std::random_device rd;
    std::mt19937_64 g(((uint64_t)rd() << 32) ^ rd());
    Fr rr = r ? *r : Rep3State::draw(g), ss = s ? *s : Rep3State::draw(g);
    Proof p;
    check(cs_groth16_prove_plain(ctx.h, pk.h, w.public_inputs[0].data(), w.witness.empty() ? nullptr : w.witness[0].data(),
                                 rr.data(), ss.data(), p.a.data(), p.b.data(), p.c.data()));
    return p;
  }
};

struct Rep3CoGroth16 {
  // Rep3CoGroth16::prove::<N, CircomReduction>(net0, net1, &pkey, &matrices, witness)  (groth16.rs:360-379).
  // The local phase (witness map on shares + the five MSMs, groth16.rs:151-294) is one call on this party's
  // GPU; the rest is create_proof_with_assignment's tail (groth16.rs:296-337) on single points.
  static Proof prove(Context& ctx, mpc_net::Network& net0, mpc_net::Network& net1, ProvingKey& pk,
                     const SharedWitness<Rep3PrimeFieldShare>& w, uint64_t seed, const G1& g1_generator,
                     Rep3PrimeFieldShare* out_r = nullptr, Rep3PrimeFieldShare* out_s = nullptr) {
    check_witness_lengths(pk, w);
    Rep3State state0(net0, seed);
    const size_t n = pk.domain_size();
    auto m1 = state0.masking_field_elements_vec(n), m2 = state0.masking_field_elements_vec(n);
    Rep3PrimeFieldShare r = state0.rand(), s = state0.rand();            // groth16.rs:157
    G1 g_a, g1_b, l_acc, h_acc;
    G2 g2_b;
    check(cs_groth16_rep3_local(ctx.h, pk.h, (int)state0.id.v, w.public_inputs[0].data(),
                                w.witness.empty() ? nullptr : w.witness[0].a.data(), m1[0].data(), m2[0].data(), r.a.data(),
                                s.a.data(), g_a.data(), g1_b.data(), g2_b.data(), l_acc.data(), h_acc.data()));
    // rs = local_mul_vec([r], [s]) (groth16.rs:297; ops.rs:69-76)
    Fr rs = fr_add(fr_add(fr_add(fr_mul(r.a, s.a), fr_mul(r.a, s.b)), fr_mul(r.b, s.a)), state0.masking_field_element());
    G1 r_s_delta_g1 = pk.delta_g1 * rs;
    // network round (groth16.rs:305-308): open_half_point(g_a) on net0 | scalar_mul(g1_b, r) on net1
    auto bc = broadcast(net0, g_a);
    G1 g_a_opened = g_a + bc.first + bc.second;                           // pointshare.rs:152-155
    G1 g1_b_prev = reshare(net1, g1_b);                                   // mpc/rep3.rs:158-160
    G1 r_g1_b = g1_b * r.a + g1_b_prev * r.a + g1_b * r.b + state0.masking_ec_element(g1_generator);  // pointshare/ops.rs:95-102
    G1 g_c = g_a_opened * s.a + r_g1_b + (-r_s_delta_g1) + l_acc + h_acc;  // groth16.rs:314-322
    auto bc_c = broadcast(net0, g_c);                                     // groth16.rs:325-328
    auto bc_b = broadcast(net1, g2_b);
    if (out_r) *out_r = r;
    if (out_s) *out_s = s;
    return Proof{g_a_opened, g2_b + bc_b.first + bc_b.second, g_c + bc_c.first + bc_c.second};
  }
};

}  // namespace co_groth16
