// fast_kernels.cuh -- specialised streaming kernels for the headline configurations (sm_100a).
//
// Design (shared by all kernels in this file): one WARP owns a vertical strip of the plane and
// marches down it.  Input rows are staged into a small per-warp shared-memory ring with cp.async
// (16-byte, L1-bypassing; boundary columns/rows are remapped or zero-filled element-wise), several
// stages ahead of the compute so HBM latency is covered by bytes in flight rather than by occupancy.
// The pass along W reads each lane's window from the ring with aligned 128-bit LDS (conflict-free:
// consecutive lanes read consecutive 16-byte words); the pass along H never touches memory: the
// last L row-filtered rows live in a register window that is shifted as the warp advances (the
// stage loop is unrolled by the window period so the shift is pure register renaming).  Filter taps
// are kernel parameters: every FFMA takes its coefficient from the constant bank.  Warps are fully
// independent (only __syncwarp), so there are no CTA barriers anywhere.
//
// Accumulation order is identical to the generic tile kernels / the oracle (stored-tap order, FMA),
// so the two paths produce bit-identical results (tests/test_gpu_parity.py::test_generic_and_auto_paths_agree).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace b200w {
namespace fast {

constexpr int kNoFastPath = 1;
static int g_force_generic = 0;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Boundary handling stays out of line: the hot loop must fit the instruction cache (an inlined
// ext_index drags three integer modulo sequences per call into every unrolled stage).
// one 4-float chunk that touches the image border (or an unaligned source): element-wise remap
__device__ __noinline__ void load_chunk_cold(float* d, const float* src_row, int gc, int W, int mode) {
#pragma unroll 1
  for (int e = 0; e < 4; ++e) {
    const int g = ext_index(gc + e, W, mode);
    if (g < 0) d[e] = 0.f;
    else cp_async4(d + e, src_row + g);
  }
}
__device__ __noinline__ int ext_index_cold(int i, int N, int mode) { return ext_index(i, N, mode); }

// store two adjacent outputs of one lane; nv = how many of them are inside the row (0..2)
__device__ __forceinline__ void store2(float* ptr, float v0, float v1, int nv, bool stream) {
  if (nv == 2 && ((reinterpret_cast<uintptr_t>(ptr) & 7) == 0)) {
    if (stream) __stcs(reinterpret_cast<float2*>(ptr), make_float2(v0, v1));
    else *reinterpret_cast<float2*>(ptr) = make_float2(v0, v1);
  } else {
    if (nv > 0) { if (stream) __stcs(ptr, v0); else ptr[0] = v0; }
    if (nv > 1) { if (stream) __stcs(ptr + 1, v1); else ptr[1] = v1; }
  }
}

// ================================================================================================
// K1 fast: DWT analysis level, Lw == Lh == L (even), modes zero / symmetric / reflect / periodic.
//   strip = 64 output columns per warp (2 per lane) = 128 input columns + (L-2) halo;
//   stage = 2 input rows = 1 output row.
// ================================================================================================
template <int L>
struct AfbCfg {
  static constexpr int HLA = ((L - 2) + 3) / 4 * 4;  // left halo rounded up to 16 bytes
  static constexpr int SW = HLA + 128;               // staged floats per row
  static constexpr int OFF = HLA - (L - 2);          // lane window offset inside its aligned read
  static constexpr int NX = OFF + L + 2;             // floats a lane needs per row
  static constexpr int NV = (NX + 3) / 4;            // ... as 128-bit loads
  static constexpr int CPR = SW / 4;                 // 16-byte chunks per staged row
  static constexpr int NCH = (2 * CPR + 31) / 32;    // chunks per lane per stage
  static constexpr int NS = 4;                       // ring depth (stages of 2 rows)
  static constexpr int NFIX = (4 * (HLA + L) + 31) / 32;  // border fix-ups per lane per stage (2 rows x 2 sides)
  static constexpr int SMEM_BYTES = NS * 2 * SW * 4;
  static constexpr int PRO = (L - 2) / 2;            // prologue stages before the first output row
  static constexpr int UNR = L / 2;                  // window period: stage copies in the unrolled loop
};

// one stage of compute: row pass on the two staged rows into window slots (2U, 2U+1) mod L, then (if emit)
// the column pass reading tap j from slot (2U+2+j) mod L, and the stores.  U is the position inside the
// window period, so every window index is a compile-time constant: the window never moves.
template <int L, int U>
__device__ __forceinline__ void afb_stage(const AfbParams& p, const float* s0, float (&wl)[L][2], float (&wh)[L][2],
                                          bool emit, float*& ll_ptr, float*& hi_ptr, long long band, int llpitch,
                                          int Wo, int nv) {
  using C = AfbCfg<L>;
  float xa[4 * C::NV], xb[4 * C::NV];
#pragma unroll
  for (int q = 0; q < C::NV; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(s0 + 4 * q);
    const float4 b = *reinterpret_cast<const float4*>(s0 + C::SW + 4 * q);
    xa[4 * q] = a.x; xa[4 * q + 1] = a.y; xa[4 * q + 2] = a.z; xa[4 * q + 3] = a.w;
    xb[4 * q] = b.x; xb[4 * q + 1] = b.y; xb[4 * q + 2] = b.z; xb[4 * q + 3] = b.w;
  }
  constexpr int SA = (2 * U) % L, SB = (2 * U + 1) % L;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    float la = 0.f, ha = 0.f, lb = 0.f, hb = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const float f0 = p.fw_lo.t[j], f1 = p.fw_hi.t[j];
      la = fmaf(f0, xa[C::OFF + 2 * o + j], la);
      ha = fmaf(f1, xa[C::OFF + 2 * o + j], ha);
      lb = fmaf(f0, xb[C::OFF + 2 * o + j], lb);
      hb = fmaf(f1, xb[C::OFF + 2 * o + j], hb);
    }
    wl[SA][o] = la; wh[SA][o] = ha;
    wl[SB][o] = lb; wh[SB][o] = hb;
  }
  if (emit) {
    float all[2], alh[2], ahl[2], ahh[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const float f0 = p.fh_lo.t[j], f1 = p.fh_hi.t[j];
        constexpr int base = 2 * U + 2;
        a0 = fmaf(f0, wl[(base + j) % L][o], a0);
        a1 = fmaf(f1, wl[(base + j) % L][o], a1);
        a2 = fmaf(f0, wh[(base + j) % L][o], a2);
        a3 = fmaf(f1, wh[(base + j) % L][o], a3);
      }
      all[o] = a0; alh[o] = a1; ahl[o] = a2; ahh[o] = a3;
    }
    store2(ll_ptr, all[0], all[1], nv, false);
    store2(hi_ptr, alh[0], alh[1], nv, true);
    store2(hi_ptr + band, ahl[0], ahl[1], nv, true);
    store2(hi_ptr + 2 * band, ahh[0], ahh[1], nv, true);
    ll_ptr += llpitch;
    hi_ptr += Wo;
  }
}

template <int L, int U>
__device__ __forceinline__ void afb_stage_dispatch(int uu, const AfbParams& p, const float* s0, float (&wl)[L][2],
                                                   float (&wh)[L][2], bool emit, float*& ll_ptr, float*& hi_ptr,
                                                   long long band, int llpitch, int Wo, int nv) {
  if constexpr (U < AfbCfg<L>::UNR) {
    if (uu == U) afb_stage<L, U>(p, s0, wl, wh, emit, ll_ptr, hi_ptr, band, llpitch, Wo, nv);
    else afb_stage_dispatch<L, U + 1>(uu, p, s0, wl, wh, emit, ll_ptr, hi_ptr, band, llpitch, Wo, nv);
  }
}

template <int L>
__global__ void __launch_bounds__(32) afb2d_stream(const __grid_constant__ AfbParams p, int n_strips, int n_chunks,
                                                   int CH) {
  using C = AfbCfg<L>;
  extern __shared__ __align__(16) float ring[];  // this warp's staging ring: NS stages x 2 rows x SW floats
  const int lane = threadIdx.x;
  long long item = blockIdx.x;                    // one warp per CTA: no intra-CTA load imbalance
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);

  const int k0 = strip * 64;
  const int ky0 = chunk * CH;
  const int ky1 = imin(ky0 + CH, p.Ho);
  const int n_stage = (ky1 - ky0) + C::PRO;
  const int c_a = 2 * k0 - C::HLA;
  const int r_begin = 2 * ky0 - (L - 2);
  const int nvalid = imin(64, p.Wo - k0);
  const int need_cols = C::HLA + 2 * nvalid;  // staged columns that feed a valid output
  const int H = p.H, W = p.W, mode = p.mode, xpitch = p.xpitch;
  const float* xp = p.x + (long long)plane * p.xps;

  // ---- static per-lane schedule (computed once; only the source rows change from stage to stage) ----
  // (a) border fix-ups: staged columns outside the image are filled from the staged copy of the column the
  //     boundary extension maps them to (or with zeros) after the stage has landed -- two shared-memory
  //     accesses instead of an element-wise global gather.  Possible when that source column is staged too.
  const int nleft = imin(imax(0, -c_a), need_cols);
  const int sr0 = imax(W - c_a, 0);                   // first staged column right of the image
  const int nright = imax(0, need_cols - sr0);
  const int nb_row = nleft + nright;
  int fix_dst[C::NFIX], fix_src[C::NFIX];
  bool bad = (2 * nb_row > 32 * C::NFIX);
#pragma unroll
  for (int q = 0; q < C::NFIX; ++q) {
    fix_dst[q] = -1;
    fix_src[q] = -1;
    const int e = lane + 32 * q;
    if (e < 2 * nb_row) {
      const int rr = (e >= nb_row) ? 1 : 0;
      const int idx = e - rr * nb_row;
      const int sidx = (idx < nleft) ? idx : sr0 + (idx - nleft);
      const int g = ext_index(c_a + sidx, W, mode);
      fix_dst[q] = rr * C::SW + sidx;
      if (g >= 0) {
        const int ss = g - c_a;
        if (ss < 0 || ss >= need_cols) bad = true;
        fix_src[q] = rr * C::SW + ss;
      }
    }
  }
  const bool use_cold = __any_sync(0xffffffffu, bad);   // e.g. 'periodic': the source is in another strip
  const bool any_fix = (nb_row > 0) && !use_cold;
  // (b) copy schedule: kind 0 nothing, 1 aligned 16-byte cp.async, 2 element-wise (only when use_cold)
  int c_soff[C::NCH], c_gcol[C::NCH], c_kind[C::NCH];
#pragma unroll
  for (int k = 0; k < C::NCH; ++k) {
    const int ch = lane + 32 * k;
    const int rr = (ch >= C::CPR) ? 1 : 0;
    const int cc = ch - rr * C::CPR;
    const int gc = c_a + 4 * cc;
    c_soff[k] = rr * C::SW + 4 * cc;
    c_gcol[k] = gc;
    int kind = 0;
    if (ch < 2 * C::CPR && 4 * cc < need_cols) {
      const bool inside = (gc >= 0 && gc + 3 < W);
      const bool partly = (gc + 3 >= 0 && gc < W);
      // a chunk straddling the right edge may be read whole: the row pitch covers it (launcher guarantees)
      if (inside || (partly && gc >= 0 && gc + 3 < xpitch && !use_cold)) kind = 1;
      else if (use_cold) kind = 2;
    }
    c_kind[k] = kind | (rr << 2);
  }

  auto issue = [&](int t) {
    if (t < n_stage) {
      float* dst = ring + (t & (C::NS - 1)) * (2 * C::SW);
      const int r0 = r_begin + 2 * t;
      const int gr0 = ((unsigned)r0 < (unsigned)H) ? r0 : ext_index_cold(r0, H, mode);
      const int gr1 = ((unsigned)(r0 + 1) < (unsigned)H) ? r0 + 1 : ext_index_cold(r0 + 1, H, mode);
      const float* src0 = xp + (long long)gr0 * xpitch;
      const float* src1 = xp + (long long)gr1 * xpitch;
#pragma unroll
      for (int k = 0; k < C::NCH; ++k) {
        const int kind = c_kind[k] & 3;
        if (kind != 0) {
          const bool rr = (c_kind[k] & 4) != 0;
          const int gr = rr ? gr1 : gr0;
          const float* src = rr ? src1 : src0;
          float* d = dst + c_soff[k];
          if (gr < 0) *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
          else if (kind == 1) cp_async16(d, src + c_gcol[k]);
          else load_chunk_cold(d, src, c_gcol[k], W, mode);
        }
      }
    }
    cp_async_commit();
  };

#pragma unroll 1
  for (int t = 0; t < C::NS - 1; ++t) issue(t);

  float wl[L][2], wh[L][2];
#pragma unroll
  for (int j = 0; j < L; ++j) { wl[j][0] = wl[j][1] = wh[j][0] = wh[j][1] = 0.f; }

  const long long band = (long long)p.Ho * p.Wo;
  float* ll_ptr = p.ll + (long long)plane * p.llps + (long long)ky0 * p.llpitch + k0 + 2 * lane;
  float* hi_ptr = p.highs + (long long)plane * 3 * band + (long long)ky0 * p.Wo + k0 + 2 * lane;
  const int nv = imax(0, imin(2, p.Wo - (k0 + 2 * lane)));
  const int llpitch = p.llpitch, Wo = p.Wo;

  int uu = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    cp_async_wait<C::NS - 2>();
    __syncwarp();
    float* stage = ring + (t & (C::NS - 1)) * (2 * C::SW);
    if (any_fix) {
#pragma unroll
      for (int q = 0; q < C::NFIX; ++q)
        if (fix_dst[q] >= 0) stage[fix_dst[q]] = (fix_src[q] >= 0) ? stage[fix_src[q]] : 0.f;
      __syncwarp();
    }
    issue(t + C::NS - 1);
    afb_stage_dispatch<L, 0>(uu, p, stage + 4 * lane, wl, wh, t >= C::PRO, ll_ptr, hi_ptr, band, llpitch, Wo, nv);
    uu = (uu + 1 == C::UNR) ? 0 : uu + 1;
  }
  cp_async_wait<0>();
}

template <int L>
inline int launch_afb_stream(const AfbParams& p, cudaStream_t stream) {
  using C = AfbCfg<L>;
  // aligned 16-byte staging needs an aligned source; anything else takes the generic kernel
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0) && (p.xpitch % 4 == 0) && (p.xps % 4 == 0);
  if (!vec_ok) return kNoFastPath;
  const int n_strips = (p.Wo + 63) / 64;
  // enough independent warps to fill the machine several times over; otherwise split the rows
  const long long want = 148LL * 32 * 3;
  const long long base = (long long)p.planes * n_strips;
  int n_chunks = 1;
  if (base < want) {
    n_chunks = (int)((want + base - 1) / (base > 0 ? base : 1));
    const int max_chunks = (p.Ho + 15) / 16;
    if (n_chunks > max_chunks) n_chunks = max_chunks;
    if (n_chunks < 1) n_chunks = 1;
  }
  const int CH = (p.Ho + n_chunks - 1) / n_chunks;
  n_chunks = (p.Ho + CH - 1) / CH;
  const long long blocks = base * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  afb2d_stream<L><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

inline int try_launch_afb(const AfbParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
  if (p.Lw != p.Lh || p.mode == B200W_MODE_PERIODIZATION) return kNoFastPath;
  if (p.planes == 0) return 0;
  switch (p.Lw) {
    case 2: return launch_afb_stream<2>(p, stream);
    case 4: return launch_afb_stream<4>(p, stream);
    case 6: return launch_afb_stream<6>(p, stream);
    case 8: return launch_afb_stream<8>(p, stream);
    case 10: return launch_afb_stream<10>(p, stream);
    case 12: return launch_afb_stream<12>(p, stream);
    case 16: return launch_afb_stream<16>(p, stream);
    default: return kNoFastPath;
  }
}

inline int try_launch_fwd_j1(const DtParams&, cudaStream_t) { return kNoFastPath; }
inline int try_launch_fwd_j2plus(const DtParams&, cudaStream_t) { return kNoFastPath; }

}  // namespace fast
}  // namespace b200w
