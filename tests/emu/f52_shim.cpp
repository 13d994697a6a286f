// Host-side test shim for cs_field52.cuh (TEST INFRASTRUCTURE): the FP64-pipe Montgomery product compiled with
// fesetround-based rounding-mode emulation, exported for ctypes.
#include "cs_emu.h"
#include "cs_params.cuh"
#include "cs_params52.cuh"
#include "cs_field52.cuh"
using namespace cs;
extern "C" {
void f52_mul_fq(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  I52 x, y; for (int k = 0; k < 5; k++) { x.l[k] = a[k]; y.l[k] = b[k]; }
  I52 r = f52_mul<Bn254Fq52>(f52_to_double(x), f52_to_double(y));
  for (int k = 0; k < 5; k++) out[k] = r.l[k];
}
void f52_sqr_fq(const uint64_t* a, uint64_t* out) {
  I52 x; for (int k = 0; k < 5; k++) x.l[k] = a[k];
  I52 r = f52_sqr<Bn254Fq52>(f52_to_double(x));
  for (int k = 0; k < 5; k++) out[k] = r.l[k];
}
void f52_sub4_fq(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  I52 x, y; for (int k = 0; k < 5; k++) { x.l[k] = a[k]; y.l[k] = b[k]; }
  I52 r = f52_sub<Bn254Fq52, 4>(x, y);
  for (int k = 0; k < 5; k++) out[k] = r.l[k];
}
void f52_roundtrip_fq(const uint32_t* in8, uint64_t* mid5, uint32_t* out8) {
  Fp<Bn254Fq> x; for (int k = 0; k < 8; k++) x.l[k] = in8[k];
  I52 m = f52_from_fp<Bn254Fq52, Bn254Fq>(x);
  for (int k = 0; k < 5; k++) mid5[k] = m.l[k];
  Fp<Bn254Fq> y = f52_to_fp<Bn254Fq52, Bn254Fq>(m);
  for (int k = 0; k < 8; k++) out8[k] = y.l[k];
}
int f52_maybe_zero_fq(const uint64_t* a) {
  I52 x; for (int k = 0; k < 5; k++) x.l[k] = a[k];
  return f52_maybe_zero_mod_p<Bn254Fq52>(x) ? 1 : 0;
}
}
