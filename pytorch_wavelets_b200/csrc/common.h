// common.h -- shared host/device definitions for libb200wave (sm_100a).
//
// The kernel bodies in tile_kernels.h are written once and compiled twice:
//   * by nvcc for sm_100a (the product: pytorch_wavelets_b200/csrc/b200wave.cu), and
//   * by g++ as a block/thread-loop emulation (tests/emu/, CPU tests of the index logic only).
// B200W_FOR_THREADS / B200W_SYNC express "every thread of the CTA runs this phase, then barrier".
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/b200wave.h"

#ifdef __CUDACC__
#define B200W_HD __host__ __device__ __forceinline__
#define B200W_D __device__ __forceinline__
#else
#define B200W_HD inline
#define B200W_D inline
#endif

#ifdef __CUDACC__
#define B200W_FOR_THREADS(tid, NT) { const int tid = (int)threadIdx.x;
#define B200W_END_THREADS }
#define B200W_SYNC() __syncthreads()
#ifdef B200W_F64   /* k_f64.cu compiles the generic kernels once more with the element type redefined to double */
#define B200W_MUL(a, b) __dmul_rn((a), (b))
#define B200W_ADD(a, b) __dadd_rn((a), (b))
#define B200W_SUB(a, b) __dsub_rn((a), (b))
#define B200W_SQRT(a) __dsqrt_rn(a)
#define B200W_DIV(a, b) __ddiv_rn((a), (b))
#else
#define B200W_MUL(a, b) __fmul_rn((a), (b))
#define B200W_ADD(a, b) __fadd_rn((a), (b))
#define B200W_SUB(a, b) __fsub_rn((a), (b))
#define B200W_SQRT(a) __fsqrt_rn(a)
#define B200W_DIV(a, b) __fdiv_rn((a), (b))
#endif
#else
#define B200W_FOR_THREADS(tid, NT) for (int tid = 0; tid < (NT); ++tid) {
#define B200W_END_THREADS }
#define B200W_SYNC() ((void)0)
#define B200W_MUL(a, b) ((a) * (b))
#define B200W_ADD(a, b) ((a) + (b))
#define B200W_SUB(a, b) ((a) - (b))
#define B200W_SQRT(a) sqrtf(a)
#define B200W_DIV(a, b) ((a) / (b))
#endif

namespace b200w {

constexpr int kMaxTaps = B200W_MAX_TAPS;
constexpr float kInvSqrt2 = (float)0.70710678118654752440;   // (a cast, not a suffix: k_f64.cu redefines the type)

// Filter taps travel as kernel parameters (constant bank): with a compile-time tap index the FFMA
// takes its coefficient straight from c[0x0][..]; with a runtime index it is one LDC.
struct alignas(16) Taps {
  float t[kMaxTaps];
};

// Boundary extension: index of the extended signal -> index in [0,N), or -1 meaning "zero".
// Same closed forms as the oracle (reference utils.py:146-163 reflect; dwt/lowlevel.py:28-88 mypad;
// :135-141 periodization pre-extension).
B200W_HD int ext_index(int i, int N, int mode) {
  if ((unsigned)i < (unsigned)N) return i;
  int p, r;
  switch (mode) {
    case B200W_MODE_SYMMETRIC:
      p = 2 * N;
      r = i % p;
      if (r < 0) r += p;
      return r < N ? r : p - 1 - r;
    case B200W_MODE_REFLECT:
      if (N == 1) return 0;
      p = 2 * N - 2;
      r = i % p;
      if (r < 0) r += p;
      return r < N ? r : p - r;
    case B200W_MODE_PERIODIC:
      r = i % N;
      if (r < 0) r += N;
      return r;
    case B200W_MODE_PERIODIZATION:
      p = N + (N & 1);
      r = i % p;
      if (r < 0) r += p;
      return r < N ? r : N - 1;
    default:
      return -1;
  }
}

// DTCWT level-1 extension: symmetric, or zero padding (reference dtcwt/lowlevel.py:75-79)
B200W_HD int sym_or_zero(int i, int N, int sym) {
  return sym ? ext_index(i, N, B200W_MODE_SYMMETRIC) : (((unsigned)i < (unsigned)N) ? i : -1);
}

B200W_HD int floordiv2(int a) { return a >> 1; }  // arithmetic shift == floor division by 2
B200W_HD int imax(int a, int b) { return a > b ? a : b; }
B200W_HD int imin(int a, int b) { return a < b ? a : b; }

// ---- parameter blocks (plain data, passed by value) ------------------------------------------

struct AfbParams {  // K1
  const float* x; long long xps; int xpitch;
  float* ll; long long llps; int llpitch;
  float* highs;
  int planes, H, W, Ho, Wo, Lw, Lh, mode;
  int tiles_x, tiles_y;
  Taps fw_lo, fw_hi, fh_lo, fh_hi;
  int hipitch;        // experiments only (streaming kernel): row pitch of the band-pass planes, 0 = Wo
  // the W-pass taps once more as interleaved {low-pass, high-pass} pairs: one aligned 64-bit constant load feeds a
  // packed FMA (pairs built from two separate arrays cost two extra uniform moves per FFMA2)
  alignas(16) float fwp[2 * kMaxTaps];
};

struct SfbParams {  // K2
  const float* ll; long long llps; int llpitch;
  const float* highs;
  float* y; long long yps; int ypitch;
  int planes, Hc, Wc, Ho, Wo, Lh, Lw, mode;
  int tiles_x, tiles_y;
  Taps gh_lo, gh_hi, gw_lo, gw_hi;
};

struct DtParams {  // K3..K7
  const float* in; long long inps; int inpitch;   // x (forward) / ll (inverse, may be null)
  float* out; long long outps; int outpitch;      // ll (forward) / y (inverse)
  float* highs;                                    // band-pass tensor (output forward, input inverse); may be null
  long long hs[6];                                 // element strides n,c,o,row,col,ri
  float* z; float* dre; float* dim;                // scat outputs
  int N, C, H, W;                                  // forward: input dims; inverse: dims of the ll / quad grid
  int L0, L1;                                      // level-1 filter lengths, or L0 = m for q-shift
  int sym;                                         // 1 symmetric extension, 0 zero padding
  float magbias, magbias2;
  int tiles_x, tiles_y;
  Taps f0, f1, f2, f3;                             // level 1: f0=h0/g0, f1=h1/g1; q-shift: f0=*0a f1=*1a f2=*0b f3=*1b
  // q-shift forward: interleaved pairs {f2[j], f0[j]} (low-pass trees b, a) and {f3[j], f1[j]} (high-pass trees)
  alignas(16) float qlo[2 * kMaxTaps];
  alignas(16) float qhi[2 * kMaxTaps];
};

}  // namespace b200w
