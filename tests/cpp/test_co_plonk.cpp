// Drives include/co_plonk.hpp the way co-plonk's own tests drive the reference:
//   co-plonk/src/round{1..5}.rs tests (Round1Challenges::deterministic -> known-answer proof elements),
//   lib.rs:300-325 (plain_prove with fresh blinders), tests/tests/circom/e2e_tests/rep3.rs (three party threads over
//   LocalNetwork::new_3_parties(), all parties return the same proof).
// The zkey is a snarkjs-format file (ingested by cs_plonk_pk_from_zkey); expected values come from a binary
// fixture written by tests/test_cpp_mirror.py from the golden vectors.  Exit code 0 = all checks passed.
#include <cstdio>
#include <fstream>
#include <thread>
#include "co_plonk.hpp"

using namespace co_plonk;

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
  if (argc < 3) return 2;
  const std::string zkey_path = argv[1];
  Reader rd(argv[2]);
  auto pub = rd.vec<Fr>(), wit = rd.vec<Fr>();
  auto exp_pts = rd.vec<G1>();
  auto exp_evs = rd.vec<Fr>();
  std::vector<Rep3PrimeFieldShare> shares[3] = {rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>()};
  std::vector<Rep3PrimeFieldShare> bsh[3] = {rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>(), rd.vec<Rep3PrimeFieldShare>()};
  EXPECT(exp_pts.size() == 9 && exp_evs.size() == 6, "fixture layout");
  const PlonkProof expected = assemble(exp_pts.data(), exp_evs.data());

  // ---- Plonk::plain_prove with the deterministic blinders == the reference's known answers
  Context ctx(0);
  Zkey zkey(ctx, zkey_path);
  EXPECT(zkey.n_public + 1 == pub.size() && zkey.n_witness == wit.size(), "zkey header");
  SharedWitness<Fr> w{pub, wit};
  Round1Challenges det = Round1Challenges::deterministic();
  PlonkProof p = Plonk::plain_prove(ctx, zkey, w, &det);
  EXPECT(p == expected, "plain proof with deterministic blinders");
  PlonkProof q = Plonk::plain_prove(ctx, zkey, w);
  EXPECT(!(q == p), "fresh blinders must change the proof");
  try {
    SharedWitness<Fr> bad{pub, std::vector<Fr>(wit.begin(), wit.end() - 1)};
    Plonk::plain_prove(ctx, zkey, bad);
    EXPECT(false, "length mismatch must fail");
  } catch (const std::runtime_error& e) {
    EXPECT(std::string(e.what()).find("witness does not match the circuit") != std::string::npos, "error message");
  }
  try {
    Zkey missing(ctx, zkey_path + ".does-not-exist");
    EXPECT(false, "missing zkey must fail");
  } catch (const std::runtime_error& e) {
    EXPECT(std::string(e.what()).find("cannot open") != std::string::npos, "open error message");
  }

  // ---- Rep3CoPlonk::prove: three parties over LocalNetwork, blinders = shares of the deterministic ones
  auto nets = mpc_net::LocalNetwork::new_3_parties();
  PlonkProof proofs[3];
  std::string errs[3];
  std::vector<std::thread> th;
  for (int i = 0; i < 3; i++)
    th.emplace_back([&, i] {
      try {
        Context c(0);
        Zkey k(c, zkey_path);
        SharedWitness<Rep3PrimeFieldShare> sw{pub, shares[i]};
        std::array<Rep3PrimeFieldShare, 11> b;
        for (int j = 0; j < 11; j++) b[j] = bsh[i][j];
        proofs[i] = Rep3CoPlonk::prove(c, *nets[i], k, sw, b, 2000 + i);
      } catch (const std::exception& e) { errs[i] = e.what(); }
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < 3; i++) EXPECT(errs[i].empty(), errs[i].c_str());
  EXPECT(proofs[0] == proofs[1] && proofs[0] == proofs[2], "all parties must return the same proof");
  EXPECT(proofs[0] == expected, "Rep3 proof == plain proof for the summed blinders (= the known answers)");
  // ---- the same through the library's own party driver (cs_plonk_rep3_prove) over the callback transport
  // (argv[3] = "library-driver" selects this section: it is exercised on its own on the GPU)
  if (argc <= 3) {
    std::printf("co_plonk.hpp: plain, error-path and 3-party Rep3 checks passed\n");
    return 0;
  }
  auto nets2 = mpc_net::LocalNetwork::new_3_parties();
  PlonkProof proofs2[3];
  std::vector<std::thread> th2;
  for (int i = 0; i < 3; i++)
    th2.emplace_back([&, i] {
      try {
        Context c(0);
        Zkey k(c, zkey_path);
        SharedWitness<Rep3PrimeFieldShare> sw{pub, shares[i]};
        std::array<Rep3PrimeFieldShare, 11> b;
        for (int j = 0; j < 11; j++) b[j] = bsh[i][j];
        proofs2[i] = Rep3CoPlonk::prove_in_library(c, *nets2[i], k, sw, &b);
      } catch (const std::exception& e) { errs[i] = e.what(); }
    });
  for (auto& t : th2) t.join();
  for (int i = 0; i < 3; i++) EXPECT(errs[i].empty(), errs[i].c_str());
  EXPECT(proofs2[0] == proofs2[1] && proofs2[0] == proofs2[2], "library driver: all parties must return the same proof");
  EXPECT(proofs2[0] == expected, "library driver: Rep3 proof == the known answers");
  std::printf("co_plonk.hpp: plain, error-path and 3-party Rep3 checks passed\n");
  return 0;
}
