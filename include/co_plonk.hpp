// co_plonk.hpp -- C++17 host-side mirror of the reference's Plonk prover interface, over the C ABI of
// libcosnarks_gpu.so (include/cosnarks_gpu.h).  Header-only; builds on co_groth16.hpp (types, mpc_net).
//
//   reference (co-circom/co-plonk/src)                         here (namespace co_plonk)
//   -------------------------------------------------------    -------------------------------------------
//   circom_types::plonk::Zkey (from_reader)                    Zkey (device-resident, cs_plonk_pk[_from_zkey])
//   circom_types::plonk::PlonkProof (round5.rs:50-70)          PlonkProof
//   Keccak256Transcript (types.rs:140-190)                     Keccak256Transcript
//   Round1Challenges::{random, deterministic} (round1.rs)      Round1Challenges
//   Plonk::plain_prove (lib.rs:271-281)                        Plonk::plain_prove
//   Rep3CoPlonk::prove(nets, zkey, witness) (lib.rs:222-240)   Rep3CoPlonk::prove(net, zkey, witness)
//   CoPlonk::prove_inner rounds 1-5 (lib.rs:80-115)            the step sequence inside Rep3CoPlonk::prove
//   PlonkProofError (lib.rs:40-69)                             std::runtime_error with the library's message
//
// tests/cpp/test_co_plonk.cpp drives it like co-plonk's own tests: deterministic blinders reproduce the
// round 1-5 known answers, and three party threads over LocalNetwork open the same proof.
#pragma once
#include "co_groth16.hpp"

namespace co_plonk {

using co_groth16::check;
using co_groth16::Context;
using co_groth16::Fr;
using co_groth16::G1;
using co_groth16::PartyID;
using co_groth16::Rep3PrimeFieldShare;
using co_groth16::SharedWitness;
using co_groth16::operator+;  // point addition on the raw limb arrays (ADL cannot find it: G1 is a std::array)

struct PlonkProof {
  G1 a, b, c, z, t1, t2, t3;
  Fr eval_a, eval_b, eval_c, eval_s1, eval_s2, eval_zw;
  G1 wxi, wxiw;
  bool operator==(const PlonkProof& o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};

struct Zkey {
  cs_plonk_pk* h = nullptr;
  size_t n_public = 0, n_witness = 0, domain_size = 0;
  std::array<G1, 8> vk{};  // Qm Ql Qr Qo Qc S1 S2 S3
  Zkey(Context& ctx, const cs_plonk_key_desc& d) { check(cs_plonk_pk_create(ctx.h, &d, &h)); info(); }
  Zkey(Context& ctx, const std::string& path) { check(cs_plonk_pk_from_zkey(ctx.h, path.c_str(), &h, nullptr, nullptr)); info(); }
  ~Zkey() { cs_plonk_pk_free(h); }
  Zkey(const Zkey&) = delete;

 private:
  void info() { check(cs_plonk_pk_info(h, &n_public, &n_witness, &domain_size, vk[0].data())); }
};

// types.rs:140-190
class Keccak256Transcript {
  std::vector<uint8_t> buf_;
  static void put_be(std::vector<uint8_t>& b, const uint64_t* limbs, int n) {
    for (int i = n - 1; i >= 0; i--)
      for (int s = 56; s >= 0; s -= 8) b.push_back((uint8_t)(limbs[i] >> s));
  }

 public:
  void add_scalar(const Fr& mont) {
    Fr c;
    check(cs_fr_from_mont(CS_BN254, mont.data(), c.data(), 1));
    put_be(buf_, c.data(), 4);
  }
  void add_point(const G1& p) {  // the point at infinity hashes as zero bytes (types.rs:168-176)
    G1 c;
    check(cs_fq_from_mont(CS_BN254, p.data(), c.data(), 2));
    put_be(buf_, c.data(), 4);
    put_be(buf_, c.data() + 4, 4);
  }
  Fr get_challenge() {  // from_be_bytes_mod_order of the digest
    uint8_t d[32];
    check(cs_keccak256(buf_.data(), buf_.size(), d));
    // reduce the 256-bit value: split as hi * 2^128 + lo, both < r, and recombine in the field
    Fr lo{}, hi{};
    for (int i = 0; i < 16; i++) {
      lo[(15 - i) / 8] |= (uint64_t)d[16 + i] << (8 * ((15 - i) % 8));
      hi[(15 - i) / 8] |= (uint64_t)d[i] << (8 * ((15 - i) % 8));
    }
    Fr two128{0, 0, 1, 0};
    Fr lo_m = co_groth16::fr_from_canonical(lo), hi_m = co_groth16::fr_from_canonical(hi), t_m = co_groth16::fr_from_canonical(two128);
    return co_groth16::fr_add(co_groth16::fr_mul(hi_m, t_m), lo_m);
  }
};

// the eleven blinding scalars of round 1 (round1.rs:45-104)
struct Round1Challenges {
  std::array<Fr, 11> b;
  static Round1Challenges deterministic() {  // b[i] = i, the setting of the reference's known-answer tests
    Round1Challenges c;
    for (uint64_t i = 0; i < 11; i++) c.b[i] = co_groth16::fr_from_canonical(Fr{i, 0, 0, 0});
    return c;
  }
  static Round1Challenges random() {
    Round1Challenges c;
    for (auto& x : c.b) x = co_groth16::fr_rand();  // OS entropy, rejection-sampled
    return c;
  }
};

inline void check_witness_lengths(const Zkey& zkey, size_t n_pub, size_t n_wit) {
  if (n_pub != zkey.n_public + 1 || n_wit != zkey.n_witness)
    throw std::runtime_error("witness does not match the circuit: expected " + std::to_string(zkey.n_public + 1) + " public and " +
                             std::to_string(zkey.n_witness) + " private values, got " + std::to_string(n_pub) + " and " +
                             std::to_string(n_wit));
}

inline PlonkProof assemble(const G1* pts, const Fr* evs) {
  return PlonkProof{pts[0], pts[1], pts[2], pts[3], pts[4], pts[5], pts[6], evs[0], evs[1], evs[2], evs[3], evs[4], evs[5], pts[7], pts[8]};
}

struct Plonk {
  // Plonk::plain_prove(zkey, private_witness)  (lib.rs:271-281)
  static PlonkProof plain_prove(Context& ctx, Zkey& zkey, const SharedWitness<Fr>& w, const Round1Challenges* ch = nullptr) {
    check_witness_lengths(zkey, w.public_inputs.size(), w.witness.size());
    Round1Challenges c = ch ? *ch : Round1Challenges::random();
    G1 pts[9];
    Fr evs[6];
    check(cs_plonk_prove_plain(ctx.h, zkey.h, w.public_inputs[0].data(), w.public_inputs.size(),
                               w.witness.empty() ? nullptr : w.witness[0].data(), w.witness.size(), c.b[0].data(), pts[0].data(),
                               evs[0].data()));
    return assemble(pts, evs);
  }
};

struct Rep3CoPlonk {
  // The same through the library's own party driver (cs_plonk_rep3_prove: step sequence, transcript and openings in
  // C++ inside libcosnarks_gpu.so) over any mpc_net::Network via the NetAdapter callbacks; correlated randomness =
  // Rep3State::new over `net`.  `blinders` == nullptr: Round1Challenges::random (eleven T::rand shares).
  static PlonkProof prove_in_library(Context& ctx, mpc_net::Network& net, Zkey& zkey, const SharedWitness<Rep3PrimeFieldShare>& w,
                                     const std::array<Rep3PrimeFieldShare, 11>* blinders = nullptr) {
    check_witness_lengths(zkey, w.public_inputs.size(), w.witness.size());
    co_groth16::NetAdapter adapter(net);
    co_groth16::Rep3State state(adapter);
    struct Session {
      cs_plonk_rep3* h = nullptr;
      ~Session() { cs_plonk_rep3_free(h); }
    } s;
    check(cs_plonk_rep3_create(ctx.h, zkey.h, (int)net.id(), &s.h));
    // nothing connected: a-halves and opened vectors travel through `net` (parties on different hosts)
    G1 pts[9];
    Fr evs[6];
    check(cs_plonk_rep3_prove(s.h, adapter.h, state.h, w.public_inputs[0].data(), w.public_inputs.size(),
                              w.witness.empty() ? nullptr : w.witness[0].a.data(), w.witness.size(),
                              blinders ? (*blinders)[0].a.data() : nullptr, pts[0].data(), evs[0].data()));
    return assemble(pts, evs);
  }

  // Rep3CoPlonk::prove(nets, zkey, witness)  (lib.rs:222-240).  One call per party (thread / process); `blinders`
  // = this party's shares of the eleven round-1 scalars (Round1Challenges::random draws them with T::rand).
  // The parties' GPUs exchange first-layer products through the arena of the next party (same-process pointer
  // here; cs_ipc_export / cs_ipc_open across processes), everything else goes through `net`.
  static PlonkProof prove(Context& ctx, mpc_net::Network& net, Zkey& zkey, const SharedWitness<Rep3PrimeFieldShare>& w,
                          const std::array<Rep3PrimeFieldShare, 11>& blinders, uint64_t seed) {
    check_witness_lengths(zkey, w.public_inputs.size(), w.witness.size());
    const PartyID id{net.id()};
    const size_t n = zkey.domain_size;
    // correlated ChaCha streams: own seed, previous party's seed (rep3.rs:71-110)
    cs_rep3_prf prf;
    std::memset(&prf, 0, sizeof(prf));
    // seed1 from the OS entropy pool (ChaCha12Rng::from_entropy, rep3.rs:57); `seed` is kept in the signature for
    // source compatibility with round-1 callers and is not used
    (void)seed;
    std::array<uint64_t, 4> own;
    co_groth16::check(cs_os_random(reinterpret_cast<uint8_t*>(own.data()), 32));
    auto prev = co_groth16::reshare(net, own);
    std::memcpy(prf.seed1, own.data(), 32);
    std::memcpy(prf.seed2, prev.data(), 32);
    prf.rounds = 12;
    struct Session {
      cs_plonk_rep3* h = nullptr;
      ~Session() { cs_plonk_rep3_free(h); }
    } s;
    check(cs_plonk_rep3_create(ctx.h, zkey.h, (int)id.v, &s.h));
    // hand the arena to the previous party, take the next party's
    void* arena = nullptr;
    size_t slot_bytes = 0;
    unsigned n_slots = 0;
    check(cs_plonk_rep3_arena(s.h, &arena, &slot_bytes, &n_slots));
    net.send(id.prev(), co_groth16::bytes_of((uintptr_t)arena));
    check(cs_plonk_rep3_connect(s.h, (void*)co_groth16::from_bytes<uintptr_t>(net.recv(id.next()))));
    void *d_out = nullptr, *d_in = nullptr;
    check(cs_plonk_rep3_io(s.h, &d_out, &d_in));

    auto barrier = [&] {  // every party's products are in place before anyone reads them
      check(cs_ctx_synchronize(ctx.h));
      co_groth16::broadcast(net, (uint8_t)1);
    };
    auto open_points = [&](G1* p, int k) {  // open_point_vec_g1 (mpc/rep3.rs:122-138)
      for (int i = 0; i < k; i++) {
        auto o = co_groth16::broadcast(net, p[i]);
        p[i] = p[i] + o.first + o.second;
      }
    };
    auto open_scalars = [&](Fr* v, int k) {  // open_vec
      for (int i = 0; i < k; i++) {
        auto o = co_groth16::broadcast(net, v[i]);
        v[i] = co_groth16::fr_add(co_groth16::fr_add(v[i], o.first), o.second);
      }
    };
    auto open_device_vector = [&](size_t m) {  // sum of the parties' additive vectors at d_out -> d_in
      std::vector<uint64_t> mine(4 * m);
      check(cs_memcpy_d2h(ctx.h, mine.data(), d_out, m * 32));
      PartyID pid{net.id()};
      std::vector<uint8_t> raw((const uint8_t*)mine.data(), (const uint8_t*)mine.data() + m * 32);
      net.send(pid.next(), raw);
      net.send(pid.prev(), raw);
      std::vector<uint8_t> a = net.recv(pid.prev()), b = net.recv(pid.next());
      if (a.size() != m * 32 || b.size() != m * 32)
        throw std::runtime_error("During execution of open_vec in MPC: Invalid number of elements received");
      void *da = nullptr, *db = nullptr;
      check(cs_dev_alloc(ctx.h, m * 32, &da));
      check(cs_dev_alloc(ctx.h, m * 32, &db));
      check(cs_memcpy_h2d(ctx.h, da, a.data(), m * 32));
      check(cs_memcpy_h2d(ctx.h, db, b.data(), m * 32));
      check(cs_vec_add(ctx.h, CS_BN254, (const uint64_t*)d_out, (const uint64_t*)da, (uint64_t*)d_in, m));
      check(cs_vec_add(ctx.h, CS_BN254, (const uint64_t*)d_in, (const uint64_t*)db, (uint64_t*)d_in, m));
      check(cs_ctx_synchronize(ctx.h));
      cs_dev_free(ctx.h, da);
      cs_dev_free(ctx.h, db);
    };
    auto step = [&](int st, const uint64_t* in, uint64_t* out) { check(cs_plonk_rep3_step(s.h, st, in, out)); };

    G1 pts[9];
    // ---- round 1
    check(cs_plonk_rep3_round1(s.h, &prf, w.public_inputs[0].data(), w.public_inputs.size(),
                               w.witness.empty() ? nullptr : w.witness[0].a.data(), w.witness.size(), blinders[0].a.data(),
                               pts[0].data()));
    open_points(pts, 3);
    // ---- round 2
    Keccak256Transcript t;
    for (const G1& p : zkey.vk) t.add_point(p);
    for (size_t i = 1; i < w.public_inputs.size(); i++) t.add_scalar(w.public_inputs[i]);
    for (int i = 0; i < 3; i++) t.add_point(pts[i]);
    const Fr beta = t.get_challenge();
    t = Keccak256Transcript();
    t.add_scalar(beta);
    const Fr gamma = t.get_challenge();
    Fr bg[2] = {beta, gamma};
    step(CS_PLONK_R3_ROUND2_A, bg[0].data(), nullptr); barrier();
    step(CS_PLONK_R3_ROUND2_B, nullptr, nullptr); barrier();
    step(CS_PLONK_R3_ROUND2_C, nullptr, nullptr);
    open_device_vector(2 * n + 1);
    step(CS_PLONK_R3_ROUND2_D, nullptr, nullptr); barrier();
    step(CS_PLONK_R3_ROUND2_E, nullptr, nullptr); barrier();
    step(CS_PLONK_R3_ROUND2_F, nullptr, nullptr);
    open_device_vector(n);
    step(CS_PLONK_R3_ROUND2_G, nullptr, pts[3].data());
    open_points(pts + 3, 1);
    // ---- round 3
    t = Keccak256Transcript();
    t.add_scalar(beta);
    t.add_scalar(gamma);
    t.add_point(pts[3]);
    const Fr alpha = t.get_challenge();
    step(CS_PLONK_R3_ROUND3_A, alpha.data(), nullptr); barrier();
    step(CS_PLONK_R3_ROUND3_B, nullptr, pts[4].data());
    open_points(pts + 4, 3);
    // ---- round 4
    t = Keccak256Transcript();
    t.add_scalar(alpha);
    for (int i = 4; i < 7; i++) t.add_point(pts[i]);
    const Fr xi = t.get_challenge();
    Fr ev[6];  // partial a b c zw | public s1 s2
    step(CS_PLONK_R3_ROUND4, xi.data(), ev[0].data());
    open_scalars(ev, 4);
    const Fr ea = ev[0], eb = ev[1], ec = ev[2], ezw = ev[3], es1 = ev[4], es2 = ev[5];
    // ---- round 5
    t = Keccak256Transcript();
    for (const Fr& x : {xi, ea, eb, ec, es1, es2, ezw}) t.add_scalar(x);
    const Fr v0 = t.get_challenge();
    Fr in5[8] = {xi, v0, ea, eb, ec, es1, es2, ezw};
    step(CS_PLONK_R3_ROUND5, in5[0].data(), pts[7].data());
    open_points(pts + 7, 2);
    Fr evs[6] = {ea, eb, ec, es1, es2, ezw};
    return assemble(pts, evs);
  }
};

}  // namespace co_plonk
