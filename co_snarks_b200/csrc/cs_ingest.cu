// snarkjs .zkey / .wtns ingest straight into the device layout.
//
// The reference parses these with taceo-circom-types (`Groth16ZKey::from_reader`, `Witness::from_reader`;
// co-circom/co-circom/src/bin/co-circom.rs:1005-1016) into arkworks structs and converts again per proof.
// The binary formats already store what the kernels want: points as affine little-endian MONTGOMERY limbs
// (all-zero = infinity) and one (matrix, row, signal, value) record per non-zero coefficient, so sections
// 5-9 are handed to cs_groth16_pk_create without touching them and section 4 becomes the CSR arrays.
// Layout facts (probed on test_vectors/): header section 2 = n8q, q, n8r, r, nVars, nPublic,
// domainSize, alpha1, beta1, beta2, gamma2, delta1, delta2; section 4 values are in R^2-Montgomery form;
// the A rows nConstraints .. nConstraints+nPublic are the public-input rows that
// groth16/reduction.rs:111-113 re-inserts, so they are dropped from the matrices here.
#include <stdio.h>
#include <fstream>
#include <map>
#include "cs_lib.cuh"
#include "cs_net.h"

using namespace cs;

namespace {

struct Sections {
  std::vector<uint8_t> data;
  std::map<uint32_t, std::pair<size_t, size_t>> sec;  // type -> (offset, length)
};

int read_file(const char* path, const char* magic, Sections& s) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return fail(CS_ERR_ARG, "cannot open %s", path);
  std::streamsize sz = f.tellg();
  f.seekg(0);
  s.data.resize((size_t)sz);
  if (sz && !f.read((char*)s.data.data(), sz)) return fail(CS_ERR_ARG, "cannot read %s", path);
  if (sz < 12 || memcmp(s.data.data(), magic, 4) != 0) return fail(CS_ERR_ARG, "%s: bad magic (expected '%s')", path, magic);
  uint32_t nsec;
  memcpy(&nsec, s.data.data() + 8, 4);
  size_t off = 12;
  for (uint32_t i = 0; i < nsec; i++) {
    if (off + 12 > s.data.size()) return fail(CS_ERR_ARG, "%s: truncated section table", path);
    uint32_t typ;
    uint64_t len;
    memcpy(&typ, s.data.data() + off, 4);
    memcpy(&len, s.data.data() + off + 4, 8);
    off += 12;
    if (len > s.data.size() - off) return fail(CS_ERR_ARG, "%s: section %u exceeds the file", path, typ);
    if (!s.sec.count(typ)) s.sec[typ] = {off, (size_t)len};
    off += len;
  }
  return 0;
}

uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// BN254 / BLS12-381 from the base-field modulus bytes
int detect_curve(const uint8_t* q, uint32_t n8q) {
  auto matches = [&](auto P, uint32_t nlimbs32) {
    if (n8q != nlimbs32 * 4) return false;
    for (uint32_t i = 0; i < nlimbs32; i++) if (rd32(q + 4 * i) != decltype(P)::mod((int)i)) return false;
    return true;
  };
  if (matches(Bn254Fq{}, 8)) return CS_BN254;
  if (matches(Bls381Fq{}, 12)) return CS_BLS12_381;
  return -1;
}

template <class FrP>
void coeff_from_r2(const uint8_t* src, uint64_t* dst) {
  // stored x R^2 -> x R: one Montgomery reduction (multiply by 1)
  host::HFp<FrP> v;
  memcpy(v.l, src, sizeof(v.l));
  host::HFp<FrP> r = v.from_mont();
  memcpy(dst, r.l, sizeof(r.l));
}

}  // namespace

extern "C" {

int cs_groth16_pk_from_zkey(cs_ctx* ctx, const char* path, int window_bits, cs_groth16_pk** out, size_t* out_n_public) {
  if (!ctx || !path || !out) return fail(CS_ERR_ARG, "cs_groth16_pk_from_zkey: NULL argument");
  Sections z;
  CS_TRY(read_file(path, "zkey", z));
  for (uint32_t t = 1; t <= 9; t++)
    if (!z.sec.count(t)) return fail(CS_ERR_ARG, "%s: section %u missing", path, t);
  const uint8_t* d = z.data.data();
  if (rd32(d + z.sec[1].first) != 1) return fail(CS_ERR_ARG, "%s: not a Groth16 zkey (protocol %u)", path, rd32(d + z.sec[1].first));
  const uint8_t* h = d + z.sec[2].first;
  const size_t hlen = z.sec[2].second;
  // sizes come from the file: validate each before it is used to form a pointer
  if (hlen < 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  uint32_t n8q = rd32(h);
  if (n8q != 32 && n8q != 48) return fail(CS_ERR_ARG, "%s: unsupported base field size %u", path, n8q);
  if (hlen < 4 + (size_t)n8q + 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  const uint8_t* q = h + 4;
  uint32_t n8r = rd32(q + n8q);
  if (n8r != 32) return fail(CS_ERR_ARG, "%s: unexpected scalar field size %u", path, n8r);
  if (hlen < 4 + (size_t)n8q + 4 + n8r + 12) return fail(CS_ERR_ARG, "%s: header section too short", path);
  const uint8_t* p = q + n8q + 4 + n8r;
  int curve = detect_curve(q, n8q);
  if (curve < 0) return fail(CS_ERR_ARG, "%s: unsupported curve (base field of %u bytes)", path, n8q);
#if !defined(CS_ENABLE_BLS12_381)
  if (curve != CS_BN254) return fail(CS_ERR_ARG, "%s: BLS12-381 support is not compiled in", path);
#endif
  if (n8r != 32) return fail(CS_ERR_ARG, "%s: unexpected scalar field size %u", path, n8r);
  uint32_t n_vars = rd32(p), n_public = rd32(p + 4), domain = rd32(p + 8);
  p += 12;
  const size_t g1 = 2 * n8q, g2 = 4 * n8q;
  const uint8_t *alpha1 = p, *beta1 = p + g1, *beta2 = p + 2 * g1, *delta1 = p + 2 * g1 + 2 * g2, *delta2 = delta1 + g1;
  if ((size_t)(delta2 + g2 - h) > z.sec[2].second) return fail(CS_ERR_ARG, "%s: header section too short", path);
  // ---- section 4 -> CSR (A, B) without the trailing public-input rows
  const uint8_t* c = d + z.sec[4].first;
  uint32_t ncoef = rd32(c);
  const size_t rec = 12 + n8r;
  if (4 + (size_t)ncoef * rec > z.sec[4].second) return fail(CS_ERR_ARG, "%s: coefficient section too short", path);
  uint32_t max_row = 0;
  for (uint32_t i = 0; i < ncoef; i++) {
    const uint8_t* r = c + 4 + (size_t)i * rec;
    if (rd32(r) > 1) return fail(CS_ERR_ARG, "%s: coefficient %u names matrix %u", path, i, rd32(r));
    if (rd32(r + 8) >= n_vars) return fail(CS_ERR_ARG, "%s: coefficient %u names signal %u >= nVars", path, i, rd32(r + 8));
    if (rd32(r + 4) > max_row) max_row = rd32(r + 4);
  }
  if (ncoef && max_row + 1 < n_public + 1) return fail(CS_ERR_ARG, "%s: fewer rows than public inputs", path);
  const size_t ni = (size_t)n_public + 1;
  if (n_vars < ni) return fail(CS_ERR_ARG, "%s: nVars %u is smaller than nPublic + 1", path, n_vars);
  const size_t nc = ncoef ? (size_t)max_row + 1 - ni : 0;
  std::vector<uint32_t> rp[2], col[2];
  std::vector<uint64_t> cf[2];
  for (int m = 0; m < 2; m++) rp[m].assign(nc + 1, 0);
  for (uint32_t i = 0; i < ncoef; i++) {
    const uint8_t* r = c + 4 + (size_t)i * rec;
    uint32_t m = rd32(r), row = rd32(r + 4);
    if (row < nc) rp[m][row + 1]++;
  }
  for (int m = 0; m < 2; m++) {
    for (size_t k = 0; k < nc; k++) rp[m][k + 1] += rp[m][k];
    col[m].resize(rp[m][nc]);
    cf[m].resize((size_t)rp[m][nc] * 4);
  }
  std::vector<uint32_t> fill[2] = {std::vector<uint32_t>(rp[0].begin(), rp[0].end() - (nc ? 1 : 0)),
                                   std::vector<uint32_t>(rp[1].begin(), rp[1].end() - (nc ? 1 : 0))};
  for (uint32_t i = 0; i < ncoef; i++) {
    const uint8_t* r = c + 4 + (size_t)i * rec;
    uint32_t m = rd32(r), row = rd32(r + 4);
    if (row >= nc) continue;
    uint32_t pos = fill[m][row]++;
    col[m][pos] = rd32(r + 8);
    if (curve == CS_BN254) coeff_from_r2<Bn254Fr>(r + 12, &cf[m][(size_t)pos * 4]);
    else coeff_from_r2<Bls381Fr>(r + 12, &cf[m][(size_t)pos * 4]);
  }
  // ---- descriptor pointing INTO the file image for every point array (copied only if the section
  // happens to start at an address that is not 8-byte aligned: offsets in the format are multiples of 4)
  std::vector<std::vector<uint64_t>> realigned;
  auto aligned = [&](const uint8_t* src, size_t bytes) -> const uint64_t* {
    if (((uintptr_t)src & 7) == 0) return (const uint64_t*)src;
    realigned.emplace_back((bytes + 7) / 8);
    memcpy(realigned.back().data(), src, bytes);
    return realigned.back().data();
  };
  cs_groth16_key_desc k;
  memset(&k, 0, sizeof(k));
  k.curve = (cs_curve)curve;
  k.num_constraints = nc;
  k.num_instance_variables = ni;
  k.num_witness_variables = n_vars - ni;
  static const uint32_t zero_rp[1] = {0};
  k.a_row_ptr = nc ? rp[0].data() : zero_rp; k.a_col = col[0].data(); k.a_coeff = cf[0].data(); k.a_nnz = col[0].size();
  k.b_row_ptr = nc ? rp[1].data() : zero_rp; k.b_col = col[1].data(); k.b_coeff = cf[1].data(); k.b_nnz = col[1].size();
  k.alpha_g1 = aligned(alpha1, g1); k.beta_g1 = aligned(beta1, g1); k.beta_g2 = aligned(beta2, g2);
  k.delta_g1 = aligned(delta1, g1); k.delta_g2 = aligned(delta2, g2);
  auto pts = [&](uint32_t t, size_t sz, const uint64_t*& ptr, size_t& len) {
    ptr = aligned(d + z.sec[t].first, z.sec[t].second);
    len = z.sec[t].second / sz;
  };
  pts(5, g1, k.a_query, k.a_query_len);
  pts(6, g1, k.b_g1_query, k.b_g1_query_len);
  pts(7, g2, k.b_g2_query, k.b_g2_query_len);
  pts(8, g1, k.l_query, k.l_query_len);
  pts(9, g1, k.h_query, k.h_query_len);
  if (k.a_query_len != n_vars || k.h_query_len != domain)
    return fail(CS_ERR_ARG, "%s: section sizes disagree with the header (A %zu vs nVars %u, H %zu vs domain %u)", path,
                k.a_query_len, n_vars, k.h_query_len, domain);
  k.window_bits = window_bits;
  if (out_n_public) *out_n_public = n_public;
  return cs_groth16_pk_create(ctx, &k, out);
}

// Plonk .zkey (protocol 2; circom_types::plonk::Zkey::from_reader, co-circom.rs:1053-1060).  Every field element
// in the file is already in Montgomery form and every polynomial is stored as `coefficients | 4n evaluations`,
// so the sections are handed to cs_plonk_pk_create as they lie; only the additions (interleaved ids and
// factors) and the Lagrange evaluations (interleaved with their coefficients) are regrouped.
//   2 header: n8q q n8r r nVars nPublic domainSize nAdditions nConstraints k1 k2 Qm Ql Qr Qo Qc S1 S2 S3 X_2
//   3 additions  4-6 wire maps  7-11 qm ql qr qo qc  12 sigma1..3  13 Lagrange  14 powers of tau
int cs_plonk_pk_from_zkey(cs_ctx* ctx, const char* path, cs_plonk_pk** out, size_t* out_n_public, size_t* out_n_witness) {
  if (!ctx || !path || !out) return fail(CS_ERR_ARG, "cs_plonk_pk_from_zkey: NULL argument");
  Sections z;
  CS_TRY(read_file(path, "zkey", z));
  for (uint32_t t = 1; t <= 14; t++)
    if (!z.sec.count(t)) return fail(CS_ERR_ARG, "%s: section %u missing", path, t);
  const uint8_t* d = z.data.data();
  if (rd32(d + z.sec[1].first) != 2) return fail(CS_ERR_ARG, "%s: not a Plonk zkey (protocol %u)", path, rd32(d + z.sec[1].first));
  const uint8_t* h = d + z.sec[2].first;
  const size_t hlen = z.sec[2].second;
  if (hlen < 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  const uint32_t n8q = rd32(h);
  if (n8q != 32 && n8q != 48) return fail(CS_ERR_ARG, "%s: unsupported base field size %u", path, n8q);
  if (hlen < 4 + (size_t)n8q + 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  const uint8_t* q = h + 4;
  const uint32_t n8r = rd32(q + n8q);
  if (n8r != 32) return fail(CS_ERR_ARG, "%s: unexpected scalar field size %u", path, n8r);
  if (hlen < 4 + (size_t)n8q + 4 + n8r + 20) return fail(CS_ERR_ARG, "%s: header section too short", path);
  const uint8_t* p = q + n8q + 4 + n8r;
  const int curve = detect_curve(q, n8q);
  if (curve < 0) return fail(CS_ERR_ARG, "%s: unsupported curve (base field of %u bytes)", path, n8q);
#if !defined(CS_ENABLE_BLS12_381)
  if (curve != CS_BN254) return fail(CS_ERR_ARG, "%s: BLS12-381 support is not compiled in", path);
#endif
  if (n8r != 32) return fail(CS_ERR_ARG, "%s: unexpected scalar field size %u", path, n8r);
  cs_plonk_key_desc k;
  memset(&k, 0, sizeof(k));
  k.curve = (cs_curve)curve;
  k.n_vars = rd32(p); k.n_public = rd32(p + 4); k.domain_size = rd32(p + 8); k.n_additions = rd32(p + 12);
  k.n_constraints = rd32(p + 16);
  p += 20;
  const size_t g1 = 2 * (size_t)n8q, g2 = 4 * (size_t)n8q;
  if ((size_t)(p + 2 * n8r + 8 * g1 + g2 - h) > z.sec[2].second) return fail(CS_ERR_ARG, "%s: header section too short", path);
  std::vector<std::vector<uint64_t>> realigned;
  auto aligned = [&](const uint8_t* src, size_t bytes) -> const uint64_t* {
    if (((uintptr_t)src & 7) == 0) return (const uint64_t*)src;
    realigned.emplace_back((bytes + 7) / 8);
    memcpy(realigned.back().data(), src, bytes);
    return realigned.back().data();
  };
  auto aligned32 = [&](const uint8_t* src, size_t bytes) -> const uint32_t* { return (const uint32_t*)aligned(src, bytes); };
  k.k1_mont = aligned(p, n8r);
  k.k2_mont = aligned(p + n8r, n8r);
  k.vk_points = aligned(p + 2 * n8r, 8 * g1);
  const size_t n = k.domain_size, na = k.n_additions, nc = k.n_constraints;
  const size_t nlag = k.n_public ? k.n_public : 1;
  const size_t poly = 5 * n * n8r;  // coefficients + 4n evaluations
  auto need = [&](uint32_t t, size_t bytes) -> int {
    if (z.sec[t].second < bytes) return fail(CS_ERR_ARG, "%s: section %u holds %zu bytes, header implies %zu", path, t, z.sec[t].second, bytes);
    return 0;
  };
  CS_TRY(need(3, na * (8 + 2 * (size_t)n8r)));
  for (uint32_t t = 4; t <= 6; t++) CS_TRY(need(t, nc * 4));
  for (uint32_t t = 7; t <= 11; t++) CS_TRY(need(t, poly));
  CS_TRY(need(12, 3 * poly));
  CS_TRY(need(13, nlag * poly));
  // additions: (u32 id1, u32 id2, f1, f2) records
  std::vector<uint32_t> add_ids(2 * na + 2);
  std::vector<uint64_t> add_f(8 * na + 8);
  const uint8_t* a = d + z.sec[3].first;
  for (size_t i = 0; i < na; i++, a += 8 + 2 * n8r) {
    add_ids[2 * i] = rd32(a);
    add_ids[2 * i + 1] = rd32(a + 4);
    memcpy(&add_f[8 * i], a + 8, 2 * (size_t)n8r);
  }
  k.additions_ids = add_ids.data();
  k.additions_factors = add_f.data();
  k.map_a = aligned32(d + z.sec[4].first, nc * 4);
  k.map_b = aligned32(d + z.sec[5].first, nc * 4);
  k.map_c = aligned32(d + z.sec[6].first, nc * 4);
  for (int i = 0; i < 5; i++) {
    const uint8_t* s0 = d + z.sec[7 + i].first;
    k.q_coeffs[i] = aligned(s0, n * n8r);
    k.q_evals[i] = aligned(s0 + n * n8r, 4 * n * n8r);
  }
  for (int i = 0; i < 3; i++) {
    const uint8_t* s0 = d + z.sec[12].first + (size_t)i * poly;
    k.s_coeffs[i] = aligned(s0, n * n8r);
    k.s_evals[i] = aligned(s0 + n * n8r, 4 * n * n8r);
  }
  std::vector<uint64_t> lag(nlag * 4 * n * 4);
  for (size_t j = 0; j < nlag; j++)
    memcpy(&lag[j * 4 * n * 4], d + z.sec[13].first + j * poly + n * n8r, 4 * n * n8r);
  k.lagrange_evals = lag.data();
  k.p_tau = aligned(d + z.sec[14].first, z.sec[14].second);
  k.n_p_tau = z.sec[14].second / g1;
  if ((size_t)k.n_vars < (size_t)k.n_additions + k.n_public + 1)
    return fail(CS_ERR_ARG, "%s: nVars %zu is smaller than nAdditions + nPublic + 1", path, (size_t)k.n_vars);
  if (out_n_public) *out_n_public = k.n_public;
  if (out_n_witness) *out_n_witness = (size_t)k.n_vars - k.n_additions - k.n_public - 1;
  return cs_plonk_pk_create(ctx, &k, out);
}

// Barretenberg / Ignition CRS file (co-noir/co-noir-common/src/crs/parse.rs:93-101,154-158; bn254_g1.dat):
// 64 bytes per G1 point, x then y, BIG-endian canonical -- byte-swapped into little-endian limbs, converted to
// Montgomery on the host and uploaded as a base set for cs_msm (the UltraHonk commitment MSMs).
int cs_bases_from_crs_file(cs_ctx* ctx, const char* path, size_t offset, size_t n, int window_bits, cs_bases** out) {
  if (!ctx || !path || !out) return fail(CS_ERR_ARG, "cs_bases_from_crs_file: NULL argument");
  if (n == 0) return fail(CS_ERR_ARG, "cs_bases_from_crs_file: empty base set");
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return fail(CS_ERR_ARG, "cannot open %s", path);
  const size_t sz = (size_t)f.tellg();
  if ((offset + n) * 64 > sz) return fail(CS_ERR_ARG, "%s: holds %zu points, %zu + %zu requested", path, sz / 64, offset, n);
  std::vector<uint8_t> raw(n * 64);
  f.seekg((std::streamoff)(offset * 64));
  if (!f.read((char*)raw.data(), (std::streamsize)raw.size())) return fail(CS_ERR_ARG, "cannot read %s", path);
  std::vector<uint64_t> canon(n * 8), mont(n * 8);
  for (size_t i = 0; i < 2 * n; i++)      // each 32-byte coordinate: reverse the bytes
    for (int l = 0; l < 4; l++) {
      uint64_t v = 0;
      for (int b = 0; b < 8; b++) v = (v << 8) | raw[i * 32 + (3 - l) * 8 + b];
      canon[i * 4 + l] = v;
    }
  CS_TRY(cs_fq_to_mont(CS_BN254, canon.data(), mont.data(), 2 * n));
  return cs_bases_upload(ctx, CS_BN254, CS_G1, mont.data(), n, window_bits, out);
}

int cs_wtns_read(const char* path, cs_curve curve, uint64_t* out_mont, size_t capacity, size_t* out_n) {
  if (!path || !out_n) return fail(CS_ERR_ARG, "cs_wtns_read: NULL argument");
  Sections w;
  CS_TRY(read_file(path, "wtns", w));
  if (!w.sec.count(1) || !w.sec.count(2)) return fail(CS_ERR_ARG, "%s: missing section", path);
  const uint8_t* h = w.data.data() + w.sec[1].first;
  if (w.sec[1].second < 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  uint32_t n8 = rd32(h);
  if (n8 != 32) return fail(CS_ERR_ARG, "%s: unexpected field size %u", path, n8);
  if (w.sec[1].second < 4 + (size_t)n8 + 4) return fail(CS_ERR_ARG, "%s: header section too short", path);
  uint32_t nvars = rd32(h + 4 + n8);
  {  // the witness must live in the scalar field of the curve it is converted for
    bool ok = true;
    for (int i = 0; i < 8; i++) {
      const uint32_t want = curve == CS_BN254 ? Bn254Fr::mod(i) : Bls381Fr::mod(i);
      ok = ok && rd32(h + 4 + 4 * i) == want;
    }
    if (curve != CS_BN254 && curve != CS_BLS12_381) return fail(CS_ERR_ARG, "cs_wtns_read: unsupported curve id %d", (int)curve);
    if (!ok) return fail(CS_ERR_ARG, "%s: the witness prime is not the scalar field of %s", path, curve == CS_BN254 ? "BN254" : "BLS12-381");
  }
  if ((size_t)nvars * n8 > w.sec[2].second) return fail(CS_ERR_ARG, "%s: values section too short", path);
  *out_n = nvars;
  if (!out_mont) return 0;  // size query
  if (capacity < nvars) return fail(CS_ERR_ARG, "cs_wtns_read: buffer holds %zu elements, file has %u", capacity, nvars);
  // canonical little-endian -> Montgomery (ark Fr representation)
  std::vector<uint64_t> tmp((size_t)nvars * 4);
  memcpy(tmp.data(), w.data.data() + w.sec[2].first, (size_t)nvars * 32);
  return cs_fr_to_mont(curve, tmp.data(), out_mont, nvars);
}

}  // extern "C"

// ---- CompressedRep3SharedWitness (bincode 1 over the serde derives; see include/cosnarks_gpu.h) ------------------
namespace {

struct Cursor {
  const uint8_t* p;
  size_t left;
  bool ok = true;
  bool take(void* dst, size_t k) {
    if (k > left) { ok = false; return false; }
    if (dst) memcpy(dst, p, k);
    p += k; left -= k;
    return true;
  }
  uint64_t u64() { uint64_t v = 0; take(&v, 8); return v; }
  uint32_t u32() { uint32_t v = 0; take(&v, 4); return v; }
};

// bytes(ark-compressed Vec<T>): u64 byte length, then u64 element count, then count * elem_bytes
bool read_ark_vec(Cursor& c, size_t elem_bytes, std::vector<uint8_t>& out, size_t& count) {
  const uint64_t blen = c.u64();
  if (!c.ok || blen > c.left || blen < 8) return c.ok = false;
  Cursor in{c.p, (size_t)blen};
  count = (size_t)in.u64();
  if (count > (blen - 8) / elem_bytes || count * elem_bytes != blen - 8) return c.ok = false;
  out.assign(in.p, in.p + count * elem_bytes);
  c.take(nullptr, (size_t)blen);
  return true;
}

// SeededType<Vec<F>, ChaCha12Rng>: 0 Shares(bytes(Vec<F>)) | 1 Seed([u8; 32], usize) -> canonical or Montgomery limbs
template <class FrP>
bool read_seeded(Cursor& c, unsigned bits, std::vector<uint64_t>& mont) {
  typedef host::HFp<FrP> HR;
  const uint32_t variant = c.u32();
  if (!c.ok) return false;
  if (variant == 0) {
    std::vector<uint8_t> raw;
    size_t n = 0;
    if (!read_ark_vec(c, 32, raw, n)) return false;
    mont.resize(n * 4);
    for (size_t i = 0; i < n; i++) {
      HR v;
      memcpy(v.l, &raw[32 * i], 32);
      if (HR::geq_mod(v.l)) return c.ok = false;  // ark rejects non-canonical encodings
      HR m = v.to_mont();
      memcpy(&mont[4 * i], m.l, 32);
    }
    return true;
  }
  if (variant != 1) return c.ok = false;
  uint8_t seed[32];
  if (!c.take(seed, 32)) return false;
  const uint64_t len = c.u64();
  if (!c.ok || len > (1ull << 32)) return c.ok = false;
  HostChaCha rng;
  rng.init(seed, 0);
  mont.resize((size_t)len * 4);
  for (size_t i = 0; i < len; i++) rng.template fr_rand<FrP>(&mont[4 * i], bits);  // expand_vec: len x F::rand (rep3.rs:181-196)
  return true;
}

template <class FrP>
int rep3_witness_read_t(const std::vector<uint8_t>& file, const char* path, unsigned bits, uint64_t* out_public, size_t pub_cap,
                        uint64_t* out_shares, size_t sh_cap, size_t* n_pub, size_t* n_wit, cs_share_kind* kind) {
  typedef host::HFp<FrP> HR;
  Cursor c{file.data(), file.size()};
  std::vector<uint8_t> raw;
  size_t np = 0;
  if (!read_ark_vec(c, 32, raw, np)) return fail(CS_ERR_ARG, "%s: malformed public inputs", path);
  const uint32_t variant = c.u32();
  if (!c.ok || variant > 3) return fail(CS_ERR_ARG, "%s: unknown share variant", path);
  std::vector<uint64_t> a, b;
  size_t nw = 0;
  bool replicated = variant < 2;
  if (variant == 0) {
    std::vector<uint8_t> sh;
    if (!read_ark_vec(c, 64, sh, nw)) return fail(CS_ERR_ARG, "%s: malformed replicated shares", path);
    a.resize(nw * 8);  // interleaved a || b directly
    for (size_t i = 0; i < 2 * nw; i++) {
      HR v;
      memcpy(v.l, &sh[32 * i], 32);
      if (HR::geq_mod(v.l)) return fail(CS_ERR_ARG, "%s: non-canonical field element", path);
      HR m = v.to_mont();
      memcpy(&a[4 * i], m.l, 32);
    }
  } else if (variant == 1) {
    if (!read_seeded<FrP>(c, bits, a) || !read_seeded<FrP>(c, bits, b)) return fail(CS_ERR_ARG, "%s: malformed seeded shares", path);
    if (a.size() != b.size()) return fail(CS_ERR_ARG, "Lengths of shares do not match");  // rep3.rs:266-268
    nw = a.size() / 4;
    std::vector<uint64_t> il(nw * 8);
    for (size_t i = 0; i < nw; i++) { memcpy(&il[8 * i], &a[4 * i], 32); memcpy(&il[8 * i + 4], &b[4 * i], 32); }
    a.swap(il);
  } else if (variant == 2) {
    std::vector<uint8_t> sh;
    if (!read_ark_vec(c, 32, sh, nw)) return fail(CS_ERR_ARG, "%s: malformed additive shares", path);
    a.resize(nw * 4);
    for (size_t i = 0; i < nw; i++) {
      HR v;
      memcpy(v.l, &sh[32 * i], 32);
      if (HR::geq_mod(v.l)) return fail(CS_ERR_ARG, "%s: non-canonical field element", path);
      HR m = v.to_mont();
      memcpy(&a[4 * i], m.l, 32);
    }
  } else {
    if (!read_seeded<FrP>(c, bits, a)) return fail(CS_ERR_ARG, "%s: malformed seeded additive shares", path);
    nw = a.size() / 4;
  }
  if (c.left != 0) return fail(CS_ERR_ARG, "%s: %zu trailing bytes", path, c.left);
  *n_pub = np;
  *n_wit = nw;
  if (kind) *kind = replicated ? CS_REP3 : CS_PLAIN;
  if (!out_public && !out_shares) return 0;  // size query
  const size_t per = replicated ? 2 : 1;
  if (!out_public || !out_shares || pub_cap < np || sh_cap < nw * per) return fail(CS_ERR_ARG, "cs_rep3_witness_read: output buffers too small");
  for (size_t i = 0; i < np; i++) {
    HR v;
    memcpy(v.l, &raw[32 * i], 32);
    if (HR::geq_mod(v.l)) return fail(CS_ERR_ARG, "%s: non-canonical field element", path);
    HR m = v.to_mont();
    memcpy(&out_public[4 * i], m.l, 32);
  }
  memcpy(out_shares, a.data(), nw * per * 32);
  return 0;
}

}  // namespace

extern "C" {

int cs_rep3_witness_read(const char* path, cs_curve curve, uint64_t* out_public, size_t public_capacity, uint64_t* out_shares,
                         size_t shares_capacity_elems, size_t* out_n_public, size_t* out_n_witness, cs_share_kind* out_kind) {
  if (!path || !out_n_public || !out_n_witness) return fail(CS_ERR_ARG, "cs_rep3_witness_read: NULL argument");
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return fail(CS_ERR_ARG, "cannot open %s", path);
  const std::streamsize sz = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> data((size_t)sz);
  if (sz && !f.read((char*)data.data(), sz)) return fail(CS_ERR_ARG, "cannot read %s", path);
  switch ((int)curve) {
    case CS_BN254:
      return rep3_witness_read_t<Bn254Fr>(data, path, 254, out_public, public_capacity, out_shares, shares_capacity_elems,
                                          out_n_public, out_n_witness, out_kind);
#if defined(CS_ENABLE_BLS12_381)
    case CS_BLS12_381:
      return rep3_witness_read_t<Bls381Fr>(data, path, 255, out_public, public_capacity, out_shares, shares_capacity_elems,
                                           out_n_public, out_n_witness, out_kind);
#endif
    default: return fail(CS_ERR_ARG, "unsupported curve");
  }
}

int cs_rep3_replicate_additive(cs_net* net, const uint64_t* h_additive, size_t n, uint64_t* h_out_shares) {
  if (!net || (n && (!h_additive || !h_out_shares))) return fail(CS_ERR_ARG, "cs_rep3_replicate_additive: NULL argument");
  if (net->n != 3) return fail(CS_ERR_ARG, "cs_rep3_replicate_additive: Rep3 needs a 3-party net");
  std::vector<uint64_t> prev(n * 4);
  Rep3Net rn(net);
  CS_TRY(rn.send_next(h_additive, n * 32));
  CS_TRY(rn.recv_prev(prev.data(), n * 32));
  for (size_t i = 0; i < n; i++) {
    memcpy(&h_out_shares[8 * i], &h_additive[4 * i], 32);
    memcpy(&h_out_shares[8 * i + 4], &prev[4 * i], 32);
  }
  return 0;
}

}  // extern "C"
