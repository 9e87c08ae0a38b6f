// stream_common.cuh -- shared device/host helpers of the streaming kernels (sm_100a): cp.async staging ring
// (StripLoader), packed FFMA2, row-chunk cost model.  The kernels themselves live in afb_stream.cuh, sfb_stream.cuh,
// dtcwt_fwd_stream.cuh, dtcwt_inv_stream.cuh (one translation unit each) and dwt_pyramid.cuh (fused pyramid).
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
#include "fast_api.h"

#ifndef B200W_EXP
#define B200W_EXP 0
#endif

namespace b200w {
namespace fast {


// experiment switches (-D at build time, see _build.build(extra_flags=...)): L2 prefetch qualifier of the staging copies, streaming stores
#ifndef B200W_CPASYNC_L2
#define B200W_CPASYNC_L2 0
#endif
#ifndef B200W_STREAM_STORES
#define B200W_STREAM_STORES 1
#endif
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
#if B200W_CPASYNC_L2 == 256
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#elif B200W_CPASYNC_L2 == 128
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
// the same with the destination already a 32-bit shared-window address (the hot path keeps those in registers:
// converting a generic pointer costs three instructions on sm_100 every time)
__device__ __forceinline__ void cp_async16_s(unsigned s, const float* gsrc) {
#if B200W_CPASYNC_L2 == 256
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#elif B200W_CPASYNC_L2 == 128
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4_s(unsigned s, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float lds_s(unsigned s) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(s) : "memory");
  return v;
}
__device__ __forceinline__ void sts_s(unsigned s, float v) {
  asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(s), "f"(v) : "memory");
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

// ---- bulk async copies (the TMA engine's 1-D form, SASS UBLKCP) completing on an mbarrier ---------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(float* smem_dst, const float* gsrc, unsigned bytes, unsigned bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  // try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so a long
  // hint keeps waiting warps out of the issue slots instead of spinning
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680)
      : "memory");
}

// Boundary handling stays out of line: the hot loop must fit the instruction cache (an inlined
// ext_index drags three integer modulo sequences per call into every unrolled stage).
static __device__ __noinline__ int ext_index_cold(int i, int N, int mode) { return ext_index(i, N, mode); }

// General (rare) stage load: rows outside the image (remapped or zero-filled) and, when `elementwise`,
// border columns gathered element by element.  Used for the few stages that touch the top/bottom border
// and for extension modes whose source column is not inside the strip (e.g. 'periodic').
static __device__ __noinline__ void load_stage_general(float* dst, int rows, int rpp, int sw, int cpr, const float* plane,
                                                long long ps, int nplanes, int r0, int H, int W, int pitch, int mode,
                                                int c_a, int need_cols, int elementwise, int lane) {
  for (int ch = lane; ch < rows * cpr; ch += 32) {
    const int v = ch / cpr;
    const int cc = ch - v * cpr;
    const int g = v / rpp;
    const int rr = v - g * rpp;
    if (4 * cc >= need_cols || g >= nplanes) continue;
    const int gr = ext_index(r0 + rr, H, mode);
    float* d = dst + v * sw + 4 * cc;
    const int gc = c_a + 4 * cc;
    if (gr < 0) {
      *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float* src = plane + (long long)g * ps + (long long)gr * pitch;
    const bool inside = (gc >= 0 && gc + 3 < W);
    if (inside || (!elementwise && gc >= 0 && gc < W && gc + 3 < pitch)) {
      cp_async16(d, src + gc);
    } else if (elementwise) {
      for (int e = 0; e < 4; ++e) {
        const int gg = ext_index(gc + e, W, mode);
        if (gg < 0) d[e] = 0.f;
        else cp_async4(d + e, src + gg);
      }
    }
  }
}

// ================================================================================================
// StripLoader: per-warp staging of a vertical strip, ROWS image rows per stage, ring of NS stages.
//   - static per-lane copy schedule (aligned 16-byte cp.async.cg), computed once;
//   - columns outside the image are filled after landing from the staged copy of the column the
//     extension maps them to ("fix-ups": one LDS + one STS per border element);
//   - rows outside the image / exotic modes go through load_stage_general (out of line).
// The caller guarantees: plane base 16-byte aligned, pitch % 4 == 0, c_a % 4 == 0, SW % 4 == 0.
// ================================================================================================
template <int ROWS, int SW, int NS, int NFIX, int RPP = ROWS, bool GSRC = false>
struct StripLoader {
  // GSRC: border elements whose source column lies outside the staged strip (wrap-around modes) are fetched
  // from global memory with the stage; without it such a strip falls back to the element-wise general loader.
  // ROWS "virtual" rows per stage = (ROWS / RPP) planes x RPP image rows: a warp working on a narrow
  // remainder strip stages the same rows of several planes at once (lanes are split between planes).
  static constexpr int CPR = SW / 4;
  static constexpr int NCH = (ROWS * CPR + 31) / 32;
  static constexpr int STAGE = ROWS * SW;  // floats per stage
  static constexpr int SMEM_FLOATS = NS * STAGE;

  float* ring;
  unsigned ring_s;       // the ring's shared-window address
  int slot_i, slot_a;    // ring slot of the next issue() / acquire() (stage t lives in slot t % NS)
  const float* plane;
  long long ps;
  int nplanes, H, W, pitch, mode, c_a, need_cols, r_begin, n_stage, lane;
  bool use_cold, any_fix;
  int c_soff[NCH];  // staged offset of the chunk this lane copies (-1: none)
  int c_goff[NCH];  // its source offset from the stage's first row (plane stride and pitch folded in when G == 1)
  int fix_dst[NFIX], fix_src[NFIX];

  __device__ __forceinline__ void init(float* ring_, const float* plane_, long long ps_, int nplanes_, int H_, int W_,
                                       int pitch_, int mode_, int c_a_, int need_cols_, int r_begin_, int n_stage_,
                                       int lane_) {
    ring = ring_; plane = plane_; ps = ps_; nplanes = nplanes_; H = H_; W = W_; pitch = pitch_; mode = mode_;
    ring_s = (unsigned)__cvta_generic_to_shared(ring_);
    slot_i = slot_a = 0;
    c_a = c_a_; need_cols = need_cols_; r_begin = r_begin_; n_stage = n_stage_; lane = lane_;
    const int nleft = imin(imax(0, -c_a), need_cols);
    const int sr0 = imax(W - c_a, 0);  // first staged column right of the image
    const int nright = imax(0, need_cols - sr0);
    const int nb_row = nleft + nright;
    const int vrows = nplanes * RPP;
    bool bad = (vrows * nb_row > 32 * NFIX);
#pragma unroll
    for (int q = 0; q < NFIX; ++q) {
      fix_dst[q] = -1;
      fix_src[q] = -1;
      const int e = lane + 32 * q;
      if (e < vrows * nb_row) {
        const int v = e / (nb_row > 0 ? nb_row : 1);
        const int idx = e - v * nb_row;
        const int sidx = (idx < nleft) ? idx : sr0 + (idx - nleft);
        const int g = ext_index(c_a + sidx, W, mode);
        fix_dst[q] = v * SW + sidx;
        if (g >= 0) {
          const int ss = g - c_a;
          // the mirrored / wrapped source column is staged too: patch from shared memory after landing;
          // otherwise ('periodic', 'periodization': it is at the other end of the row) fetch it from global
          // memory together with the stage -- encoded as -(column) - 2
          if (ss >= 0 && ss < need_cols && g < W) fix_src[q] = v * SW + ss;
          else if (GSRC) fix_src[q] = -g - 2;
          else bad = true;
        }
      }
    }
    use_cold = __any_sync(0xffffffffu, bad);
    any_fix = (nb_row > 0) && !use_cold;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ch = lane + 32 * k;
      const int v = ch / CPR;
      const int cc = ch - v * CPR;
      const int g = v / RPP;
      const int rr = v - g * RPP;
      const int gc = c_a + 4 * cc;
      // a chunk straddling the right edge is read whole: the row pitch covers it
      const bool on = (ch < ROWS * CPR) && (g < nplanes) && (4 * cc < need_cols) && (gc >= 0) && (gc < W) &&
                      (gc + 3 < pitch);
      c_soff[k] = on ? (v * SW + 4 * cc) : -1;
      c_goff[k] = (int)((long long)g * ps + (long long)rr * pitch + gc);  // launcher keeps this below 2^31
    }
  }

  __device__ __forceinline__ void issue(int t) {
    const int slot = slot_i;
    slot_i = (slot_i + 1 == NS) ? 0 : slot_i + 1;
    if (t < n_stage) {
      const int r0 = r_begin + RPP * t;
      if (!use_cold && r0 >= 0 && r0 + RPP <= H) {
        const unsigned dst_s = ring_s + slot * (STAGE * 4);
        const float* src = plane + (long long)r0 * pitch;
#pragma unroll
        for (int k = 0; k < NCH; ++k)
          if (c_soff[k] >= 0) cp_async16_s(dst_s + 4 * c_soff[k], src + c_goff[k]);
        if (GSRC && any_fix) {
          float* dst = ring + slot * STAGE;
#pragma unroll
          for (int q = 0; q < NFIX; ++q)
            if (fix_dst[q] >= 0 && fix_src[q] <= -2) {
              const int v = fix_dst[q] / SW;
              const int g = v / RPP, rr = v - g * RPP;
              cp_async4(dst + fix_dst[q], src + (long long)g * ps + (long long)rr * pitch + (-fix_src[q] - 2));
            }
        }
      } else {
        float* dst = ring + slot * STAGE;
        load_stage_general(dst, ROWS, RPP, SW, CPR, plane, ps, nplanes, r0, H, W, pitch, mode, c_a, need_cols,
                           use_cold ? 1 : 0, lane);
        if (GSRC && any_fix) {  // border elements whose source is elsewhere in the (remapped) row
#pragma unroll
          for (int q = 0; q < NFIX; ++q)
            if (fix_dst[q] >= 0 && fix_src[q] <= -2) {
              const int v = fix_dst[q] / SW;
              const int g = v / RPP, rr = v - g * RPP;
              const int gr = ext_index_cold(r0 + rr, H, mode);
              if (gr < 0) dst[fix_dst[q]] = 0.f;
              else cp_async4(dst + fix_dst[q], plane + (long long)g * ps + (long long)gr * pitch + (-fix_src[q] - 2));
            }
        }
      }
    }
    cp_async_commit();
  }

  __device__ __forceinline__ void prologue() {
#pragma unroll 1
    for (int t = 0; t < NS - 1; ++t) issue(t);
  }

  // wait for stage t, make it visible to the warp, patch the border columns; returns the stage base
  __device__ __forceinline__ float* acquire(int t) {
    cp_async_wait<NS - 2>();
    __syncwarp();
    const int slot = slot_a;
    slot_a = (slot_a + 1 == NS) ? 0 : slot_a + 1;
    if (any_fix) {
      const unsigned st_s = ring_s + slot * (STAGE * 4);
#pragma unroll
      for (int q = 0; q < NFIX; ++q)
        if (fix_dst[q] >= 0 && fix_src[q] >= -1)
          sts_s(st_s + 4 * fix_dst[q], (fix_src[q] >= 0) ? lds_s(st_s + 4 * fix_src[q]) : 0.f);
      __syncwarp();
    }
    return ring + slot * STAGE;
  }
};

// store two adjacent outputs of one lane; nv = how many of them are inside the row (0..2)
__device__ __forceinline__ void store2(float* ptr, float v0, float v1, int nv, bool stream) {
  stream = stream && (B200W_STREAM_STORES != 0);
  if (nv == 2 && ((reinterpret_cast<uintptr_t>(ptr) & 7) == 0)) {
    if (stream) __stcs(reinterpret_cast<float2*>(ptr), make_float2(v0, v1));
    else *reinterpret_cast<float2*>(ptr) = make_float2(v0, v1);
  } else {
    if (nv > 0) { if (stream) __stcs(ptr, v0); else ptr[0] = v0; }
    if (nv > 1) { if (stream) __stcs(ptr + 1, v1); else ptr[1] = v1; }
  }
}

// Packed fp32 FMA (Blackwell FFMA2): d = a * b + c on both halves, each an IEEE fma -- the same roundings as two
// scalar fmaf, in one issue slot.  ptxas folds a duplicated scalar ({x, x}) into the broadcast operand form and
// takes tap pairs / scalars straight from uniform registers, so the pairs cost no extra moves.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 ffma2_s(float x, float2 b, float2 c) { return ffma2(make_float2(x, x), b, c); }

// Resident warps of a one-warp-per-CTA kernel on the whole GPU (cached per kernel by the caller).
template <class K>
inline int resident_warps(K kernel, int smem_bytes, int threads = 32) {
  int per_sm = 0, dev = 0, sms = 148;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = imax(1, 16 / (threads / 32));
  }
  if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * (threads / 32) * sms;
}

// Per-device cache of resident_warps (one per call site): the occupancy query is not free, and a process may
// drive several devices with different SM counts / carve-outs.
struct ConcCache { int v[64] = {}; };
template <class K>
inline int resident_warps_dev(ConcCache& c, K kernel, int smem_bytes, int threads = 32) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return resident_warps(kernel, smem_bytes, threads);
  int v = c.v[dev];
  if (v == 0) { v = resident_warps(kernel, smem_bytes, threads); c.v[dev] = v; }
  return v;
}

// How many row-chunks to split each (plane, strip) march into.  Cost model, in units of one output row of one
// warp: every chunk pays `pro` extra rows (halo + pipeline fill + schedule set-up); the grid drains through `conc`
// resident warps; the last wave leaves the machine partly idle for about half a chunk.  Calibrated on B200
// against K1/K3/K4 timings at 1..10 chunks (profiles/r01_notes.md).
inline void pick_chunks(long long base_items, int rows_out, int min_rows, int pro, int conc, int* n_chunks, int* CH) {
  const int max_chunks = (rows_out + min_rows - 1) / min_rows;
  double best = 0.0;
  int best_ch = (rows_out + min_rows - 1) / min_rows * min_rows, best_nc = 1, last_ch = -1;
  if (best_ch < min_rows) best_ch = min_rows;
  for (int nc = 1; nc <= max_chunks && nc <= 64; ++nc) {
    int ch = (rows_out + nc - 1) / nc;
    ch = (ch + min_rows - 1) / min_rows * min_rows;
    if (ch == last_ch) continue;
    last_ch = ch;
    const int n = (rows_out + ch - 1) / ch;
    const double work = (double)base_items * (rows_out + (double)n * pro) / (conc > 0 ? conc : 1);
    const double cost = work + 0.5 * (ch + pro);
    if (nc == 1 || cost < best) { best = cost; best_ch = ch; best_nc = n; }
  }
#ifdef B200W_CHUNK_SCALE_NUM   /* experiments: scale the chosen chunk count (variant builds only) */
  {
    int nc = best_nc * B200W_CHUNK_SCALE_NUM / B200W_CHUNK_SCALE_DEN;
    if (nc < 1) nc = 1;
    if (nc > max_chunks) nc = max_chunks;
    int ch = (rows_out + nc - 1) / nc;
    ch = (ch + min_rows - 1) / min_rows * min_rows;
    best_ch = ch;
    best_nc = (rows_out + ch - 1) / ch;
  }
#endif
  *CH = best_ch;
  *n_chunks = best_nc;
}

inline bool aligned_plane(const void* base, long long ps, int pitch) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (pitch % 4 == 0) && (ps % 4 == 0);
}

// A narrow last strip can have mirrored border columns whose source lies LEFT of its staged window (Wo = 257: the
// strip holds one output column, the mirror needs columns up to 6 to its left).  Instead of dropping to the
// element-wise loader, stage `sh` more columns on the left (a multiple of 4, as far as the row stride allows).
static __device__ __noinline__ int widen_left(int c_a, int need, int W, int mode, int room) {
  if (c_a + need <= W || c_a <= 0) return 0;
  const int g = ext_index(c_a + need - 1, W, mode);  // source of the farthest border column (mirror modes)
  if (g < 0 || g >= c_a) return 0;
  const int sh = (c_a - g + 3) & ~3;
  return (sh <= c_a && sh <= room) ? sh : 0;  // room: what the widest reading lane leaves of the row stride
}

}  // namespace fast
}  // namespace b200w
