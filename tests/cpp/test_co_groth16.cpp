// Drives include/co_groth16.hpp the way the reference's own tests drive co-groth16:
//   co-groth16/src/lib.rs:40-69 (plain prove), tests/tests/circom/e2e_tests/rep3.rs:36-137 (three party
//   threads over LocalNetwork::new_3_parties(), two networks per party, all parties return the same proof).
// Expected values come from a binary fixture written by tests/test_cpp_mirror.py from the golden vectors
// (oracle proof for fixed (r, s)).  Exit code 0 = all checks passed.
#include <cstdio>
#include <fstream>
#include <thread>
#include "co_groth16.hpp"

using namespace co_groth16;

struct Reader {
  std::ifstream f;
  explicit Reader(const char* p) : f(p, std::ios::binary) { if (!f) throw std::runtime_error("cannot open fixture"); }
  uint64_t u64() { uint64_t v; f.read((char*)&v, 8); return v; }
  template <class T> std::vector<T> vec() {
    uint64_t n = u64();
    std::vector<T> v(n);
    if (n) f.read((char*)v.data(), n * sizeof(T));
    return v;
  }
};

#define EXPECT(cond, msg) do { if (!(cond)) { std::fprintf(stderr, "FAIL: %s\n", msg); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  Reader rd(argv[1]);
  cs_groth16_key_desc d{};
  d.curve = CS_BN254;
  d.num_constraints = rd.u64(); d.num_instance_variables = rd.u64(); d.num_witness_variables = rd.u64();
  auto a_rp = rd.vec<uint32_t>(); auto a_col = rd.vec<uint32_t>(); auto a_cf = rd.vec<uint64_t>();
  auto b_rp = rd.vec<uint32_t>(); auto b_col = rd.vec<uint32_t>(); auto b_cf = rd.vec<uint64_t>();
  d.a_row_ptr = a_rp.data(); d.a_col = a_col.data(); d.a_coeff = a_cf.data(); d.a_nnz = a_col.size();
  d.b_row_ptr = b_rp.data(); d.b_col = b_col.data(); d.b_coeff = b_cf.data(); d.b_nnz = b_col.size();
  auto alpha = rd.vec<uint64_t>(), beta1 = rd.vec<uint64_t>(), beta2 = rd.vec<uint64_t>(), delta1 = rd.vec<uint64_t>(), delta2 = rd.vec<uint64_t>();
  d.alpha_g1 = alpha.data(); d.beta_g1 = beta1.data(); d.beta_g2 = beta2.data(); d.delta_g1 = delta1.data(); d.delta_g2 = delta2.data();
  auto aq = rd.vec<uint64_t>(), b1q = rd.vec<uint64_t>(), b2q = rd.vec<uint64_t>(), lq = rd.vec<uint64_t>(), hq = rd.vec<uint64_t>();
  d.a_query = aq.data(); d.a_query_len = aq.size() / 8;
  d.b_g1_query = b1q.data(); d.b_g1_query_len = b1q.size() / 8;
  d.b_g2_query = b2q.data(); d.b_g2_query_len = b2q.size() / 16;
  d.l_query = lq.data(); d.l_query_len = lq.size() / 8;
  d.h_query = hq.data(); d.h_query_len = hq.size() / 8;
  auto pub = rd.vec<Fr>(), wit = rd.vec<Fr>();
  auto rs = rd.vec<Fr>();               // r, s (Montgomery)
  auto exp_a = rd.vec<uint64_t>(), exp_b = rd.vec<uint64_t>(), exp_c = rd.vec<uint64_t>();
  std::vector<Rep3PrimeFieldShare> shares[3] = {rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>()};
  auto gen = rd.vec<uint64_t>();
  (void)gen;

  // ---- Groth16::plain_prove with injected (r, s) == oracle proof bytes
  Context ctx(0);
  ProvingKey pk(ctx, d);
  SharedWitness<Fr> w{pub, wit};
  Proof p = Groth16::plain_prove(ctx, pk, w, &rs[0], &rs[1]);
  EXPECT(std::memcmp(p.a.data(), exp_a.data(), sizeof(G1)) == 0, "plain proof A");
  EXPECT(std::memcmp(p.b.data(), exp_b.data(), sizeof(G2)) == 0, "plain proof B");
  EXPECT(std::memcmp(p.c.data(), exp_c.data(), sizeof(G1)) == 0, "plain proof C");
  // fresh randomness still yields a well-formed (different) proof
  Proof q = Groth16::plain_prove(ctx, pk, w);
  EXPECT(std::memcmp(q.a.data(), p.a.data(), sizeof(G1)) != 0, "fresh r must change A");

  // ---- error behaviour of prove_inner's length checks (groth16.rs:134-149)
  try {
    SharedWitness<Fr> bad{pub, std::vector<Fr>(wit.begin(), wit.end() - 1)};
    Groth16::plain_prove(ctx, pk, bad);
    EXPECT(false, "length mismatch must fail");
  } catch (const std::runtime_error& e) {
    EXPECT(std::string(e.what()).find("amount of private witness variables does not match") != std::string::npos, "error message");
  }

  // ---- Rep3CoGroth16::prove: three parties, two LocalNetworks each (rep3.rs:57-72)
  auto nets0 = mpc_net::LocalNetwork::new_3_parties();
  auto nets1 = mpc_net::LocalNetwork::new_3_parties();
  Proof proofs[3];
  Rep3PrimeFieldShare rsh[3], ssh[3];
  std::string errs[3];
  std::vector<std::thread> th;
  for (int i = 0; i < 3; i++)
    th.emplace_back([&, i] {
      try {
        Context c(0);
        ProvingKey k(c, d);
        SharedWitness<Rep3PrimeFieldShare> sw{pub, shares[i]};
        proofs[i] = Rep3CoGroth16::prove(c, *nets0[i], *nets1[i], k, sw, &rsh[i], &ssh[i]);
      } catch (const std::exception& e) { errs[i] = e.what(); }
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < 3; i++) EXPECT(errs[i].empty(), errs[i].c_str());
  EXPECT(std::memcmp(&proofs[0], &proofs[1], sizeof(Proof)) == 0 && std::memcmp(&proofs[0], &proofs[2], sizeof(Proof)) == 0,
         "all parties must return the same proof");
  // masks cancel on opening: the MPC proof is the plain proof for r = sum r_i.a, s = sum s_i.a
  Fr r_tot = fr_add(fr_add(rsh[0].a, rsh[1].a), rsh[2].a), s_tot = fr_add(fr_add(ssh[0].a, ssh[1].a), ssh[2].a);
  Proof plain = Groth16::plain_prove(ctx, pk, w, &r_tot, &s_tot);
  EXPECT(std::memcmp(&plain, &proofs[0], sizeof(Proof)) == 0, "Rep3 proof == plain proof for summed randomness");
  std::printf("co_groth16.hpp: plain, error-path and 3-party Rep3 checks passed\n");
  return 0;
}
