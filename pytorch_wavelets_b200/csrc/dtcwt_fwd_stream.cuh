#pragma once
#include "stream_common.cuh"

namespace b200w {
namespace fast {

// fast_dtcwt.cuh -- streaming DTCWT forward kernels (included inside namespace b200w::fast by
// fast_kernels.cuh).  Same machinery as the DWT kernel: per-warp strip, cp.async ring (StripLoader),
// 128/64-bit conflict-free LDS for the pass along W, rotating register window for the pass along H,
// taps from the constant bank, q2c (or the ScatLayer magnitude) applied in registers before the store.

// store one complex number (re, im) of orientation slot `o` for this lane
__device__ __forceinline__ void store_cplx(float* hq, long long so, long long sr, bool vec, int o, float re, float im) {
  float* q = hq + o * so;
  if (vec) __stcs(reinterpret_cast<float2*>(q), make_float2(re, im));
  else { __stcs(q, re); __stcs(q + sr, im); }
}

// q2c of one real subband quad (a b / c d) into orientation slots o1 (w1) and o2 (w2)
__device__ __forceinline__ void q2c_emit(float a, float b, float c, float d, float* hq, long long so, long long sr,
                                         bool vec, int o1, int o2) {
  a = __fmul_rn(a, kInvSqrt2); b = __fmul_rn(b, kInvSqrt2);
  c = __fmul_rn(c, kInvSqrt2); d = __fmul_rn(d, kInvSqrt2);
  store_cplx(hq, so, sr, vec, o1, __fsub_rn(a, d), __fadd_rn(b, c));
  store_cplx(hq, so, sr, vec, o2, __fadd_rn(a, d), __fsub_rn(b, c));
}

// sqrt of a non-negative finite sum of squares, branch-free: MUFU.RSQ plus one Newton step on the residual (an exact FMA),
// with tiny / zero arguments rescaled by selects.  The IEEE routine (__fsqrt_rn) carries a range-check branch and a slow-path
// call per use; twelve of them per stage serialise twelve dependent MUFU chains, which is what bounded the ScatLayer
// kernel (profiles/r02_notes.md).  The result is the correctly rounded root except for rare 1-ulp ties -- far inside the
// 1e-5 parity tolerance against the reference's torch.sqrt.
// SAFE = false: the argument is known to be >= magbias^2 >= 1e-30 (the usual case, magbias = 1e-2): no rescaling, no zero test.
template <bool SAFE>
__device__ __forceinline__ float sqrt_nonneg(float s) {
  if (!SAFE) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
    const float g = s * y;
    return fmaf(fmaf(-g, g, s), 0.5f * y, g);
  }
  const bool tiny = s < 1e-30f;
  const float t = tiny ? s * 18446744073709551616.f : s;          // x 2^64 (exact)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
  float g = t * y;
  const float h = 0.5f * y;
  const float e = fmaf(-g, g, t);
  g = fmaf(e, h, g);
  g = tiny ? g * 2.3283064365386963e-10f : g;                     // x 2^-32
  return (s <= 0.f) ? 0.f : g;   // (NaN propagates)
}

// ScatLayer epilogue for one subband quad: smoothed magnitudes of w1 / w2 (+ re/r, im/r when DERIV: the tensors the
// backward pass needs -- a separate instantiation, so the inference kernel carries no division code)
template <bool DERIV, bool SAFE>
__device__ __forceinline__ void scat_emit(float a, float b, float c, float d, const DtParams& p, float* z1, float* z2,
                                          long long dbase, long long ostride, int o1, int o2) {
  a = __fmul_rn(a, kInvSqrt2); b = __fmul_rn(b, kInvSqrt2);
  c = __fmul_rn(c, kInvSqrt2); d = __fmul_rn(d, kInvSqrt2);
  const float re[2] = {__fsub_rn(a, d), __fadd_rn(a, d)};
  const float im[2] = {__fadd_rn(b, c), __fsub_rn(b, c)};
  const int os[2] = {o1, o2};
  float* const zq[2] = {z1, z2};                      // this lane's element of the magnitude planes of o1, o2
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float rr = __fmul_rn(re[k], re[k]), ii = __fmul_rn(im[k], im[k]);
    const float r = sqrt_nonneg<SAFE>(__fadd_rn(__fadd_rn(rr, ii), p.magbias2));
    __stcs(zq[k], __fsub_rn(r, p.magbias));
    if (DERIV) {
      __stcs(p.dre + dbase + os[k] * ostride, __fdiv_rn(re[k], r));
      __stcs(p.dim + dbase + os[k] * ostride, __fdiv_rn(im[k], r));
    }
  }
}

// ================================================================================================
// K3 / K7 fast: DTCWT level-1 forward (odd filter lengths L0, L1), optional ScatLayer epilogue.
//   strip = 64 columns per warp (2 per lane); stage = 2 image rows = one row of 2x2 quads.
// ================================================================================================
template <int L0, int L1>
struct J1Cfg {
  static constexpr int M0 = L0 / 2, M1 = L1 / 2, M = (M0 > M1) ? M0 : M1;
  static constexpr int HLA = (M + 3) / 4 * 4;
  static constexpr int SW = HLA + 64 + HLA;
  static constexpr int OFFX = HLA - M;              // staged index of column (c - M) for the lane's first column
  static constexpr int NX = OFFX + 2 * M + 2;       // floats a lane needs per row, from its aligned 8-byte base
  static constexpr int NV2 = (NX + 1) / 2;          // ... as 64-bit loads
  static constexpr int WR = 2 * M + 2;              // register window rows
  static constexpr int UNR = M + 1;                 // window period in stages
  static constexpr int PRO = M;
  static constexpr int NS = 4;
  static constexpr int NFIX = (2 * 2 * HLA + 31) / 32;
  static constexpr int SMEM_BYTES = (NS * 2 * SW + 2 * NS) * 4;
  using Loader = StripLoader<2, SW, NS, NFIX>;
};

// One stage of the two passes.  The outputs of the column pass (one 2x2 quad per band and column pair) are handed back
// in registers: the epilogue (stores / q2c / ScatLayer magnitudes) does not depend on the window position U, so it lives
// once in the kernel's main loop instead of once per unrolled copy of the stage (instruction-cache footprint).
struct J1Quads { float vll[2][2], vlh[2][2], vhl[2][2], vhh[2][2]; };   // [dr][o]

template <int L0, int L1, int U>
__device__ __forceinline__ void j1_stage(const DtParams& p, const float* s0, float2 (&w)[J1Cfg<L0, L1>::WR][2],
                                         bool emit, J1Quads& v) {
  using C = J1Cfg<L0, L1>;
  constexpr int WR = C::WR;
  // row pass on the two staged rows; window entries are {low-pass, high-pass} pairs (packed FMA where both filters
  // have a tap on the sample, scalar FMA on the longer filter's outer taps -- same products, same order)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float x[2 * C::NV2];
#pragma unroll
    for (int q = 0; q < C::NV2; ++q) {
      const float2 t = *reinterpret_cast<const float2*>(s0 + r * C::SW + 2 * q);
      x[2 * q] = t.x; x[2 * q + 1] = t.y;
    }
    const int S = (2 * U + r) % WR;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 2 * C::M + 1; ++i) {
        const int j0 = i - (C::M - C::M0), j1 = i - (C::M - C::M1);
        const float xv = x[C::OFFX + o + i];
        const bool in0 = (j0 >= 0 && j0 < L0), in1 = (j1 >= 0 && j1 < L1);
        if (in0 && in1) acc = ffma2_s(xv, make_float2(p.f0.t[in0 ? j0 : 0], p.f1.t[in1 ? j1 : 0]), acc);
        else if (in0) acc.x = fmaf(p.f0.t[in0 ? j0 : 0], xv, acc.x);
        else if (in1) acc.y = fmaf(p.f1.t[in1 ? j1 : 0], xv, acc.y);
      }
      w[S][o] = acc;
    }
  }
  if (emit) {
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float2 ac = make_float2(0.f, 0.f), bd = make_float2(0.f, 0.f);   // {ll, hl}, {lh, hh}
#pragma unroll
        for (int j = 0; j < L0; ++j) {
          const int sl = (2 * U + dr - C::M - C::M0 + j + 4 * WR) % WR;
          ac = ffma2_s(p.f0.t[j], w[sl][o], ac);
        }
#pragma unroll
        for (int j = 0; j < L1; ++j) {
          const int sl = (2 * U + dr - C::M - C::M1 + j + 4 * WR) % WR;
          bd = ffma2_s(p.f1.t[j], w[sl][o], bd);
        }
        v.vll[dr][o] = ac.x; v.vlh[dr][o] = bd.x; v.vhl[dr][o] = ac.y; v.vhh[dr][o] = bd.y;
      }
  }
}

template <int L0, int L1, int U>
__device__ __forceinline__ void j1_dispatch(int uu, const DtParams& p, const float* s0,
                                            float2 (&w)[J1Cfg<L0, L1>::WR][2], bool emit, J1Quads& v) {
  if constexpr (U < J1Cfg<L0, L1>::UNR) {
    if (uu == U) j1_stage<L0, L1, U>(p, s0, w, emit, v);
    else j1_dispatch<L0, L1, U + 1>(uu, p, s0, w, emit, v);
  }
}

// SCAT: 0 = DTCWT level 1 (q2c), 1 = ScatLayer magnitudes, 2 = ScatLayer magnitudes + derivative tensors
#ifndef B200W_J1_MINB
#define B200W_J1_MINB 1
#endif
// (an explicit minBlocks of 1 is not neutral: ptxas then spends registers freely -- fwd_j2plus 156 -> 176, fwd_j1 96 -> 124 --
// so the plain form is used unless a cap is asked for)
#if B200W_J1_MINB > 1
#define B200W_J1_LB __launch_bounds__(32, B200W_J1_MINB)
#else
#define B200W_J1_LB __launch_bounds__(32)
#endif
template <int L0, int L1, int SCAT>
__global__ void B200W_J1_LB fwd_j1_stream(const __grid_constant__ DtParams p, int n_strips, int n_chunks,
                                                    int CH /* quad rows per chunk */) {
  using C = J1Cfg<L0, L1>;
  extern __shared__ __align__(16) float ring[];
  const int lane = threadIdx.x;
  long long item = blockIdx.x;
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);
  const int n = plane / p.C, ch = plane - n * p.C;

  const int c0 = strip * 64;
  const int qy0 = chunk * CH;
  const int qy1 = imin(qy0 + CH, p.H >> 1);
  const int n_stage = (qy1 - qy0) + C::PRO;
  const int ncols = imin(64, p.W - c0);

  typename C::Loader ld;
  ld.init(ring, p.in + (long long)plane * p.inps, p.inps, 1, p.H, p.W, p.inpitch,
          p.sym ? B200W_MODE_SYMMETRIC : B200W_MODE_ZERO,
          c0 - C::HLA, C::HLA + ncols + C::M, 2 * qy0 - C::M, n_stage, lane);
  ld.prologue();

  float2 w[C::WR][2];
#pragma unroll
  for (int j = 0; j < C::WR; ++j) { w[j][0] = w[j][1] = make_float2(0.f, 0.f); }

  const bool colvalid = (c0 + 2 * lane) < p.W;
  const int h2 = p.H >> 1, w2 = p.W >> 1;
  float* ll_ptr = SCAT ? nullptr : p.out + (long long)plane * p.outps + (long long)(2 * qy0) * p.outpitch + c0 + 2 * lane;
  float* hq = nullptr;
  bool vec = false;
  if (!SCAT && p.highs) {
    hq = p.highs + n * p.hs[0] + ch * p.hs[1] + (long long)qy0 * p.hs[3] + (long long)((c0 >> 1) + lane) * p.hs[4];
    vec = (p.hs[5] == 1) && ((p.hs[4] & 1) == 0) && ((p.hs[3] & 1) == 0) && ((p.hs[2] & 1) == 0) &&
          ((p.hs[1] & 1) == 0) && ((p.hs[0] & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.highs) & 7) == 0);
  }
  // scat: z is (N,7,C,h2,w2), dre/dim (N,6,C,h2,w2); orientation stride = C*h2*w2
  const long long ostride = (long long)p.C * h2 * w2;
  const long long zplane = ((long long)n * 7 * p.C + ch) * h2 * w2;
  const long long dplane = ((long long)n * 6 * p.C + ch) * h2 * w2;
  long long zoff = (long long)qy0 * w2 + (c0 >> 1) + lane;
  // ScatLayer: this lane's element of the seven output planes (avg-pooled low-pass + six magnitudes), advanced by one
  // output row per stage (pointer increments instead of seven 64-bit address computations per stage)
  float* zp[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) zp[k] = SCAT ? p.z + zplane + zoff + k * ostride : nullptr;
  const bool tiny_bias = !(p.magbias2 >= 1e-30f);     // magbias = 0 (or denormal): the root needs its zero / tiny handling

  int uu = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    const float* stage = ld.acquire(t);
    ld.issue(t + C::NS - 1);
    const bool emit = (t >= C::PRO);
    J1Quads v;
    j1_dispatch<L0, L1, 0>(uu, p, stage + 2 * lane, w, emit, v);
    uu = (uu + 1 == C::UNR) ? 0 : uu + 1;
    if (emit) {
      if (colvalid) {
        if (!SCAT) {
          store2(ll_ptr, v.vll[0][0], v.vll[0][1], 2, false);
          store2(ll_ptr + p.outpitch, v.vll[1][0], v.vll[1][1], 2, false);
          if (p.highs) {
            const long long so = p.hs[2], sr = p.hs[5];
            q2c_emit(v.vlh[0][0], v.vlh[0][1], v.vlh[1][0], v.vlh[1][1], hq, so, sr, vec, 0, 5);  // lh -> 15, 165
            q2c_emit(v.vhh[0][0], v.vhh[0][1], v.vhh[1][0], v.vhh[1][1], hq, so, sr, vec, 1, 4);  // hh -> 45, 135
            q2c_emit(v.vhl[0][0], v.vhl[0][1], v.vhl[1][0], v.vhl[1][1], hq, so, sr, vec, 2, 3);  // hl -> 75, 105
          }
        } else {
          float s = __fadd_rn(v.vll[0][0], v.vll[0][1]);
          s = __fadd_rn(s, v.vll[1][0]);
          s = __fadd_rn(s, v.vll[1][1]);
          __stcs(zp[0], __fmul_rn(s, 0.25f));
          const long long db = dplane + zoff;
          if (!tiny_bias) {
            scat_emit<SCAT == 2, false>(v.vlh[0][0], v.vlh[0][1], v.vlh[1][0], v.vlh[1][1], p, zp[1], zp[6], db, ostride, 0, 5);
            scat_emit<SCAT == 2, false>(v.vhh[0][0], v.vhh[0][1], v.vhh[1][0], v.vhh[1][1], p, zp[2], zp[5], db, ostride, 1, 4);
            scat_emit<SCAT == 2, false>(v.vhl[0][0], v.vhl[0][1], v.vhl[1][0], v.vhl[1][1], p, zp[3], zp[4], db, ostride, 2, 3);
          } else {
            scat_emit<SCAT == 2, true>(v.vlh[0][0], v.vlh[0][1], v.vlh[1][0], v.vlh[1][1], p, zp[1], zp[6], db, ostride, 0, 5);
            scat_emit<SCAT == 2, true>(v.vhh[0][0], v.vhh[0][1], v.vhh[1][0], v.vhh[1][1], p, zp[2], zp[5], db, ostride, 1, 4);
            scat_emit<SCAT == 2, true>(v.vhl[0][0], v.vhl[0][1], v.vhl[1][0], v.vhl[1][1], p, zp[3], zp[4], db, ostride, 2, 3);
          }
        }
      }
      ll_ptr += 2 * p.outpitch;
      hq += p.hs[3];
      zoff += (p.W >> 1);
      if (SCAT) {
#pragma unroll
        for (int k = 0; k < 7; ++k) zp[k] += (p.W >> 1);
      }
    }
  }
  cp_async_wait<0>();
}

template <int L0, int L1, int SCAT>
inline int launch_j1_stream(const DtParams& p, cudaStream_t stream) {
  using C = J1Cfg<L0, L1>;
  if (!aligned_plane(p.in, p.inps, p.inpitch)) return kNoFastPath;
  if (!SCAT && (p.outpitch & 1)) return kNoFastPath;
  const int n_strips = (p.W + 63) / 64;
  const long long planes = (long long)p.N * p.C;
  int n_chunks, CH;
  static ConcCache conc_cache;
  const int conc = resident_warps_dev(conc_cache, fwd_j1_stream<L0, L1, SCAT>, C::SMEM_BYTES);
  // (per-chunk overhead of 4 rows: twice the chunk count of the round-1 calibration measured 1.5 % faster on configs[2])
  pick_chunks(planes * n_strips, p.H >> 1, 8, 4, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  fwd_j1_stream<L0, L1, SCAT><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

template <int SCAT>
inline int try_launch_j1_any(const DtParams& p, cudaStream_t stream) {
  if (!SCAT && !p.highs) return kNoFastPath;  // skip_hps: low-pass only, generic kernel
  if ((long long)p.N * p.C == 0) return 0;
  if (p.L0 == 5 && p.L1 == 7) return launch_j1_stream<5, 7, SCAT>(p, stream);   // near_sym_a analysis
  if (p.L0 == 7 && p.L1 == 5) return launch_j1_stream<7, 5, SCAT>(p, stream);   // near_sym_a synthesis (backward)
  if (p.L0 == 9 && p.L1 == 7) return launch_j1_stream<9, 7, SCAT>(p, stream);   // antonini
  if (p.L0 == 5 && p.L1 == 3) return launch_j1_stream<5, 3, SCAT>(p, stream);   // legall
  if (p.L0 == 13 && p.L1 == 19) return launch_j1_stream<13, 19, SCAT>(p, stream);  // near_sym_b
  if (!SCAT) {  // synthesis filter pairs: the backward pass of the level-1 inverse
    if (p.L0 == 7 && p.L1 == 9) return launch_j1_stream<7, 9, 0>(p, stream);    // antonini
    if (p.L0 == 3 && p.L1 == 5) return launch_j1_stream<3, 5, 0>(p, stream);    // legall
    if (p.L0 == 19 && p.L1 == 13) return launch_j1_stream<19, 13, 0>(p, stream);  // near_sym_b
  }
  return kNoFastPath;
}
int try_launch_fwd_j1(const DtParams& p, cudaStream_t stream) { return try_launch_j1_any<0>(p, stream); }
int try_launch_scat_j1(const DtParams& p, cudaStream_t stream) {
  return (p.dre != nullptr) ? try_launch_j1_any<2>(p, stream) : try_launch_j1_any<1>(p, stream);
}

// ================================================================================================
// K4 fast: DTCWT level >= 2 forward, q-shift filters of even length MQ.
//   lane = one output complex column q: 4 input columns -> 2 half-resolution columns;
//   strip = 32 q = 128 input columns; stage = 4 input rows = 2 half-resolution rows = 1 quad row.
//   taps: f0=h0a f1=h1a f2=h0b f3=h1b (stored).
// ================================================================================================
#ifndef B200W_FWDJ2_NS
#define B200W_FWDJ2_NS 4   /* ring depth: 2 -> 1.207 ms, 3 -> 1.137 ms, 4 -> 1.125 ms (DTCWT forward, configs[2]) */
#endif
template <int MQ>
struct J2Cfg {
  static constexpr int HL = MQ - 2;
  static constexpr int HLA = (HL + 3) / 4 * 4;
  static constexpr int SW = HLA + 128 + HLA;
  static constexpr int OFF = HLA - HL;
  static constexpr int NX = OFF + 2 * MQ;
  static constexpr int NV = (NX + 3) / 4;
  static constexpr int WR = 2 * MQ;
  static constexpr int UNR = MQ / 2;
  static constexpr int PRO = (MQ - 2) / 2;
  static constexpr int NS = B200W_FWDJ2_NS;
  static constexpr int NFIX = (4 * 2 * HLA + 31) / 32;
  static constexpr int SMEM_BYTES = (NS * 4 * SW + 2 * NS) * 4;
  using Loader = StripLoader<4, SW, NS, NFIX>;
};

template <int MQ, int U>
__device__ __forceinline__ void j2_stage(const DtParams& p, const float* s0, float2 (&wl)[2 * MQ],
                                         float2 (&wh)[2 * MQ], bool emit, bool want_hi, float*& ll_ptr, float*& hq,
                                         bool qvalid, bool vec) {
  using C = J2Cfg<MQ>;
  constexpr int WR = C::WR;
  static_assert(C::OFF % 2 == 0, "sample pairs must be register pairs");
  // Row pass: each tap multiplies the (even, odd) sample pair by a tap pair in one packed FMA.
  //   wl[S] = {Ya(h0b), Yb(h0a)} = the low-pass interleave (a, b);  wh[S] = {Ya(h1b), Yb(h1a)} -- the high-pass
  //   interleave is (b, a), i.e. wh[S] read back swapped.
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float x[4 * C::NV];
#pragma unroll
    for (int q = 0; q < C::NV; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(s0 + r * C::SW + 4 * q);
      x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    float2 l = make_float2(0.f, 0.f), h = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MQ; ++j) {
      const float2 xv = make_float2(x[C::OFF + 2 * j], x[C::OFF + 2 * j + 1]);
      l = ffma2(xv, make_float2(p.qlo[2 * j], p.qlo[2 * j + 1]), l);  // {Ya with h0b, Yb with h0a}
      h = ffma2(xv, make_float2(p.qhi[2 * j], p.qhi[2 * j + 1]), h);  // {Ya with h1b, Yb with h1a}
    }
    const int S = (4 * U + r) % WR;
    wl[S] = l;
    wh[S] = h;
  }
  if (emit) {
    float vll[2][2], vlh[2][2], vhl[2][2], vhh[2][2];  // [half-res row 0/1][half-res col 0/1]
    // Column pass, both half-resolution columns of a band in one packed FMA (tap broadcast).  Pairs built from wh
    // come out column-swapped (see above).
    float2 ll0 = make_float2(0.f, 0.f), ll1 = ll0, lh0 = ll0, lh1 = ll0, hl0 = ll0, hl1 = ll0, hh0 = ll0, hh1 = ll0;
#pragma unroll
    for (int j = 0; j < MQ; ++j) {
      const int sa = (4 * U + 4 + 2 * j) % WR, sb = (4 * U + 5 + 2 * j) % WR;
      ll0 = ffma2_s(p.f2.t[j], wl[sa], ll0);   // ll[2q]   = Ya(h0b) on lo
      ll1 = ffma2_s(p.f0.t[j], wl[sb], ll1);   // ll[2q+1] = Yb(h0a)
      lh0 = ffma2_s(p.f1.t[j], wl[sb], lh0);   // lh[2q]   = Yb(h1a)   (high-pass interleave)
      lh1 = ffma2_s(p.f3.t[j], wl[sa], lh1);   // lh[2q+1] = Ya(h1b)
      hl0 = ffma2_s(p.f2.t[j], wh[sa], hl0);
      hl1 = ffma2_s(p.f0.t[j], wh[sb], hl1);
      hh0 = ffma2_s(p.f1.t[j], wh[sb], hh0);
      hh1 = ffma2_s(p.f3.t[j], wh[sa], hh1);
    }
    vll[0][0] = ll0.x; vll[0][1] = ll0.y; vll[1][0] = ll1.x; vll[1][1] = ll1.y;
    vlh[0][0] = lh0.x; vlh[0][1] = lh0.y; vlh[1][0] = lh1.x; vlh[1][1] = lh1.y;
    vhl[0][0] = hl0.y; vhl[0][1] = hl0.x; vhl[1][0] = hl1.y; vhl[1][1] = hl1.x;
    vhh[0][0] = hh0.y; vhh[0][1] = hh0.x; vhh[1][0] = hh1.y; vhh[1][1] = hh1.x;
    if (qvalid) {
      store2(ll_ptr, vll[0][0], vll[0][1], 2, false);
      store2(ll_ptr + p.outpitch, vll[1][0], vll[1][1], 2, false);
      if (want_hi) {
        const long long so = p.hs[2], sr = p.hs[5];
        q2c_emit(vlh[0][0], vlh[0][1], vlh[1][0], vlh[1][1], hq, so, sr, vec, 0, 5);
        q2c_emit(vhh[0][0], vhh[0][1], vhh[1][0], vhh[1][1], hq, so, sr, vec, 1, 4);
        q2c_emit(vhl[0][0], vhl[0][1], vhl[1][0], vhl[1][1], hq, so, sr, vec, 2, 3);
      }
    }
    ll_ptr += 2 * p.outpitch;
    hq += p.hs[3];
  }
}

template <int MQ, int U>
__device__ __forceinline__ void j2_dispatch(int uu, const DtParams& p, const float* s0, float2 (&wl)[2 * MQ],
                                            float2 (&wh)[2 * MQ], bool emit, bool want_hi, float*& ll_ptr,
                                            float*& hq, bool qvalid, bool vec) {
  if constexpr (U < J2Cfg<MQ>::UNR) {
    if (uu == U) j2_stage<MQ, U>(p, s0, wl, wh, emit, want_hi, ll_ptr, hq, qvalid, vec);
    else j2_dispatch<MQ, U + 1>(uu, p, s0, wl, wh, emit, want_hi, ll_ptr, hq, qvalid, vec);
  }
}

#ifndef B200W_FWDJ2_MINB
#define B200W_FWDJ2_MINB 1
#endif
// (an explicit minBlocks of 1 is not neutral: ptxas then spends registers freely -- fwd_j2plus 156 -> 176, fwd_j1 96 -> 124 --
// so the plain form is used unless a cap is asked for)
#if B200W_FWDJ2_MINB > 1
#define B200W_FWDJ2_LB __launch_bounds__(32, B200W_FWDJ2_MINB)
#else
#define B200W_FWDJ2_LB __launch_bounds__(32)
#endif
template <int MQ>
__global__ void B200W_FWDJ2_LB fwd_j2plus_stream(const __grid_constant__ DtParams p, int n_strips,
                                                        int n_chunks, int CH /* quad rows per chunk */) {
  using C = J2Cfg<MQ>;
  extern __shared__ __align__(16) float ring[];
  const int lane = threadIdx.x;
  long long item = blockIdx.x;
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);
  const int n = plane / p.C, ch = plane - n * p.C;

  const int q0 = strip * 32;
  const int Q = p.W >> 2;
  const int qy0 = chunk * CH;
  const int qy1 = imin(qy0 + CH, p.H >> 2);
  const int n_stage = (qy1 - qy0) + C::PRO;
  const int nq = imin(32, Q - q0);

  typename C::Loader ld;
  ld.init(ring, p.in + (long long)plane * p.inps, p.inps, 1, p.H, p.W, p.inpitch, B200W_MODE_SYMMETRIC, 4 * q0 - C::HLA,
          C::HLA + 4 * nq + C::HL, 4 * qy0 + 2 - MQ, n_stage, lane);
  ld.prologue();

  float2 wl[C::WR], wh[C::WR];
#pragma unroll
  for (int j = 0; j < C::WR; ++j) { wl[j] = wh[j] = make_float2(0.f, 0.f); }

  const bool qvalid = (q0 + lane) < Q;
  const bool want_hi = (p.highs != nullptr);
  float* ll_ptr = p.out + (long long)plane * p.outps + (long long)(2 * qy0) * p.outpitch + 2 * (q0 + lane);
  float* hq = nullptr;
  bool vec = false;
  if (want_hi) {
    hq = p.highs + n * p.hs[0] + ch * p.hs[1] + (long long)qy0 * p.hs[3] + (long long)(q0 + lane) * p.hs[4];
    vec = (p.hs[5] == 1) && ((p.hs[4] & 1) == 0) && ((p.hs[3] & 1) == 0) && ((p.hs[2] & 1) == 0) &&
          ((p.hs[1] & 1) == 0) && ((p.hs[0] & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.highs) & 7) == 0);
  }

  int uu = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    const float* stage = ld.acquire(t);
    ld.issue(t + C::NS - 1);
    j2_dispatch<MQ, 0>(uu, p, stage + 4 * lane, wl, wh, t >= C::PRO, want_hi, ll_ptr, hq, qvalid, vec);
    uu = (uu + 1 == C::UNR) ? 0 : uu + 1;
  }
  cp_async_wait<0>();
}

template <int MQ>
inline int launch_j2_stream(const DtParams& p, cudaStream_t stream) {
  using C = J2Cfg<MQ>;
  if (!aligned_plane(p.in, p.inps, p.inpitch)) return kNoFastPath;
  if (p.outpitch & 1) return kNoFastPath;
  const int n_strips = ((p.W >> 2) + 31) / 32;
  const long long planes = (long long)p.N * p.C;
  int n_chunks, CH;
  static ConcCache conc_cache;
  const int conc = resident_warps_dev(conc_cache, fwd_j2plus_stream<MQ>, C::SMEM_BYTES);
  pick_chunks(planes * n_strips, p.H >> 2, 4, 3, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  fwd_j2plus_stream<MQ><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

int try_launch_fwd_j2plus(const DtParams& p, cudaStream_t stream) {
  if ((long long)p.N * p.C == 0) return 0;
  if (p.L0 == 10) return launch_j2_stream<10>(p, stream);  // qshift_a, qshift_06
  if (p.L0 == 14) return launch_j2_stream<14>(p, stream);  // qshift_b
  if (p.L0 == 16) return launch_j2_stream<16>(p, stream);  // qshift_c
  if (p.L0 == 18) return launch_j2_stream<18>(p, stream);  // qshift_d
  return kNoFastPath;
}


}  // namespace fast
}  // namespace b200w
