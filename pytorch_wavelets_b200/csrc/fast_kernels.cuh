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

#ifndef B200W_EXP
#define B200W_EXP 0
#endif

namespace b200w {
namespace fast {

constexpr int kNoFastPath = 1;
static int g_force_generic = 0;
static long long g_tune_want = 0;
static int g_tune_hipitch = 0;
static int g_tune_balanced = 0;

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
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// Boundary handling stays out of line: the hot loop must fit the instruction cache (an inlined
// ext_index drags three integer modulo sequences per call into every unrolled stage).
__device__ __noinline__ int ext_index_cold(int i, int N, int mode) { return ext_index(i, N, mode); }

// General (rare) stage load: rows outside the image (remapped or zero-filled) and, when `elementwise`,
// border columns gathered element by element.  Used for the few stages that touch the top/bottom border
// and for extension modes whose source column is not inside the strip (e.g. 'periodic').
__device__ __noinline__ void load_stage_general(float* dst, int rows, int rpp, int sw, int cpr, const float* plane,
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
  // Alternative staging (compile with -DB200W_BULK_STAGING=1): each staged row is ONE bulk async copy
  // (cp.async.bulk, the TMA engine's 1-D form, SASS UBLKCP) of its in-image part, completing on a per-stage
  // mbarrier -- no per-lane address generation, no LDGSTS traffic.  Measured on B200 it LOSES to the per-lane
  // 16-byte cp.async schedule at this granularity (544-byte rows, 2 KB stages per warp): DWT level 1
  // 2.31 vs 1.91 ms, level 2 0.81 vs 0.58 ms (profiles/r01_notes.md), so it is off by default.
#ifndef B200W_BULK_STAGING
#define B200W_BULK_STAGING 0
#endif
  static constexpr bool kBulk = (B200W_BULK_STAGING != 0) && (ROWS == RPP) && (ROWS <= 4);
  static constexpr int SMEM_FLOATS = NS * STAGE + 2 * NS;  // ring + NS mbarriers (8 bytes each)

  unsigned bar0;
  int b_soff, b_goff, b_bytes;
  bool bulk_on;
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
    {  // bulk path: the in-image part [cstart, cend) of every staged row, rounded to 16 bytes inside the pitch
      const int cstart = imax(0, c_a);
      const int cend = imin(c_a + (need_cols + 3) / 4 * 4, (W + 3) / 4 * 4);
      b_soff = cstart - c_a;
      b_goff = cstart;
      b_bytes = (cend - cstart) * 4;
      bulk_on = kBulk && !use_cold && (cend > cstart);
      bar0 = (unsigned)__cvta_generic_to_shared(ring + NS * STAGE);
      if (bulk_on) {
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NS; ++i) mbar_init(bar0 + 8 * i, 1);
          mbar_fence_init();
        }
      }
      __syncwarp();
    }
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

  __device__ __forceinline__ void issue_bulk(int t) {
    if (t < n_stage) {
      float* dst = ring + (t % NS) * STAGE;
      const unsigned bar = bar0 + 8 * (t % NS);
      const int r0 = r_begin + RPP * t;
      int mine = -1, nvalid = 0;
#pragma unroll
      for (int v = 0; v < ROWS; ++v) {
        const int r = r0 + v;
        const int g = ((unsigned)r < (unsigned)H) ? r : ext_index_cold(r, H, mode);
        if (g >= 0) ++nvalid;
        else for (int i = lane; i < SW; i += 32) dst[v * SW + i] = 0.f;  // zero padding above / below the image
        if (lane == v) mine = g;
      }
      fence_proxy_async();  // earlier generic-proxy writes to this slot (fix-ups, zero fill) before the async writes
      if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)(nvalid * b_bytes));
      if (lane < ROWS && mine >= 0)
        bulk_copy_g2s(dst + lane * SW + b_soff, plane + (long long)mine * pitch + b_goff, (unsigned)b_bytes, bar);
    }
  }

  __device__ __forceinline__ void issue(int t) {
    if (kBulk && bulk_on) { issue_bulk(t); return; }
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
    if (kBulk && bulk_on) {
      __syncwarp();  // every lane is done with the slot the next issue() will overwrite
      mbar_wait(bar0 + 8 * (t % NS), (unsigned)((t / NS) & 1));
      float* stage = ring + (t % NS) * STAGE;
      if (any_fix) {
#pragma unroll
        for (int q = 0; q < NFIX; ++q)
          if (fix_dst[q] >= 0 && fix_src[q] >= -1) stage[fix_dst[q]] = (fix_src[q] >= 0) ? stage[fix_src[q]] : 0.f;
        __syncwarp();
      }
      return stage;
    }
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

// How many row-chunks to split each (plane, strip) march into.  Cost model, in units of one output row of one
// warp: every chunk pays `pro` extra rows (halo + pipeline fill + schedule set-up); the grid drains through `conc`
// resident warps; the last wave leaves the machine partly idle for about half a chunk.  Calibrated on B200
// against K1/K3/K4 timings at 1..10 chunks (profiles/r01_notes.md).
inline void pick_chunks(long long base_items, int rows_out, int min_rows, int pro, int conc, int* n_chunks, int* CH) {
  if (g_tune_want > 0) {  // experiments: the old fixed-target rule
    int nc = 1;
    if (base_items < g_tune_want) nc = (int)((g_tune_want + base_items - 1) / (base_items > 0 ? base_items : 1));
    int ch = (rows_out + nc - 1) / (nc > 0 ? nc : 1);
    ch = (ch + min_rows - 1) / min_rows * min_rows;
    if (ch < min_rows) ch = min_rows;
    *CH = ch;
    *n_chunks = (rows_out + ch - 1) / ch;
    return;
  }
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
  *CH = best_ch;
  *n_chunks = best_nc;
}

inline bool aligned_plane(const void* base, long long ps, int pitch) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (pitch % 4 == 0) && (ps % 4 == 0);
}

// A narrow last strip can have mirrored border columns whose source lies LEFT of its staged window (Wo = 257: the
// strip holds one output column, the mirror needs columns up to 6 to its left).  Instead of dropping to the
// element-wise loader, stage `sh` more columns on the left (a multiple of 4, as far as the row stride allows).
__device__ __noinline__ int widen_left(int c_a, int need, int W, int mode, int room) {
  if (c_a + need <= W || c_a <= 0) return 0;
  const int g = ext_index(c_a + need - 1, W, mode);  // source of the farthest border column (mirror modes)
  if (g < 0 || g >= c_a) return 0;
  const int sh = (c_a - g + 3) & ~3;
  return (sh <= c_a && sh <= room) ? sh : 0;  // room: what the widest reading lane leaves of the row stride
}

// ================================================================================================
// K1 fast: DWT analysis level, Lw == Lh == L (even), modes zero / symmetric / reflect / periodic.
//   strip = 64 output columns per warp (2 per lane) = 128 input columns + (L-2) halo;
//   stage = 2 input rows = 1 output row.
// ================================================================================================
// XM: 0 = zero / symmetric / reflect (border sources inside the strip), 1 = periodic, 2 = periodization
// (wrap-around modes: border columns come from the other end of the row, fetched with the stage)
template <int L, int PW = 32, int HSM = 2, int XM = 0>
struct AfbCfg {
  static constexpr bool PER = (XM == 2);
  // PW = column pairs per plane handled by one warp: 32 -> the warp owns one 64-column strip of one plane;
  // PW < 32 -> a narrow remainder strip, the warp's lanes are split over G = 32/PW planes.
  static constexpr int G = 32 / PW;
  // out[k] = sum_j f[j] xe[2k + j - PL]: PL = L-2, or L/2-1 for periodization (reference afb1d :134-154)
  static constexpr int PL = PER ? (L - 1 - L / 2) : (L - 2);
  static constexpr int HLA = (PL + 3) / 4 * 4;       // left halo rounded up to 16 bytes
  static constexpr int OFF = HLA - PL;               // lane window offset inside its aligned read
  static constexpr int NX = OFF + L + 2;             // floats a lane needs per row
  static constexpr int NV = (NX + 3) / 4;            // ... as 128-bit loads
  static constexpr int SW = 4 * (PW - 1) + 4 * NV;   // staged floats per row (per plane)
  static constexpr int RH = L - 2 - PL;              // columns needed right of the last output's 2k+1
  // half-stages (2 input rows = 1 output row) per stage: the largest of 4, 2, 1 dividing the window period
  static constexpr int HS0 = ((L / 2) % 4 == 0) ? 4 : (((L / 2) % 2 == 0) ? 2 : 1);
  static constexpr int HS = (HS0 < HSM) ? HS0 : HSM;
  static constexpr int RPS = 2 * HS;                 // image rows per stage
#ifndef B200W_AFB_NS
#define B200W_AFB_NS 3
#endif
  static constexpr int NS = (HS == 4) ? 2 : ((HS == 2) ? B200W_AFB_NS : 4);  // ring depth in stages
  static constexpr int NFIX = (PW == 32) ? (RPS * 2 * (HLA + L) + 31) / 32 : 8;  // border fix-ups per lane per stage
  static constexpr int SMEM_BYTES = (NS * RPS * G * SW + 2 * NS) * 4;
  static constexpr int PRO = (L - 2) / 2;            // prologue half-stages before the first output row
  static constexpr int UNR = L / 2;                  // window period in half-stages
  static constexpr int UNS = UNR / HS;               // ... in stages: copies of the stage body
  using Loader = StripLoader<RPS * G, SW, NS, NFIX, RPS, (XM != 0)>;
};

// ---- where a finished output row goes -------------------------------------------------------------------------
// DirectOut: straight from registers to global memory (each lane stores its two adjacent columns of each band).
struct DirectOut {
  float* ll_ptr; float* hi_ptr; long long band; int llpitch, Wo, nv;
  // Rows of the reference's contiguous outputs start at any 4-byte phase (Wo = 259: every other row is 8-byte
  // misaligned).  `par` bit b = this row of plane b (0 = ll, 1..3 = the band-pass planes) is misaligned; the value is
  // the same in every lane (lanes are 8 bytes apart), and it flips with `flip` from row to row.  One bit test per
  // store replaces an address test + divergence bookkeeping around each of them.
  unsigned par, flip;
  __device__ __forceinline__ void init_parity() {
    par = (unsigned)((reinterpret_cast<uintptr_t>(ll_ptr) >> 2) & 1) |
          (unsigned)((reinterpret_cast<uintptr_t>(hi_ptr) >> 2) & 1) << 1 |
          (unsigned)((reinterpret_cast<uintptr_t>(hi_ptr + band) >> 2) & 1) << 2 |
          (unsigned)((reinterpret_cast<uintptr_t>(hi_ptr + 2 * band) >> 2) & 1) << 3;
    flip = (unsigned)(llpitch & 1) | ((Wo & 1) ? 14u : 0u);
  }
  __device__ __forceinline__ void pair(float* ptr, float v0, float v1, unsigned odd, bool stream) {
    stream = stream && (B200W_STREAM_STORES != 0);
    if (!odd) {
      if (stream) __stcs(reinterpret_cast<float2*>(ptr), make_float2(v0, v1));
      else *reinterpret_cast<float2*>(ptr) = make_float2(v0, v1);
    } else {
      if (stream) { __stcs(ptr, v0); __stcs(ptr + 1, v1); }
      else { ptr[0] = v0; ptr[1] = v1; }
    }
  }
  __device__ __forceinline__ void row(float2 lo0, float2 lo1, float2 hi0, float2 hi1) {
    // {column low-pass of (l, h)} = {ll, band 1}; {column high-pass} = {band 0, band 2}   (reference order lh, hl, hh)
    if (nv == 2) {
      pair(ll_ptr, lo0.x, lo1.x, par & 1u, false);
      pair(hi_ptr, hi0.x, hi1.x, par & 2u, true);
      pair(hi_ptr + band, lo0.y, lo1.y, par & 4u, true);
      pair(hi_ptr + 2 * band, hi0.y, hi1.y, par & 8u, true);
    } else if (nv == 1) {   // the last column of an odd-width plane
      ll_ptr[0] = lo0.x;
      hi_ptr[0] = hi0.x;
      hi_ptr[band] = lo0.y;
      hi_ptr[2 * band] = hi0.y;
    }
    par ^= flip;
    ll_ptr += llpitch;
    hi_ptr += Wo;
  }
};

// one half-stage of compute: row pass on the two staged rows into window slots (2U, 2U+1) mod L, then (if emit)
// the column pass reading tap j from slot (2U+2+j) mod L, and the output row.  U is the position inside the
// window period, so every window index is a compile-time constant: the window never moves.
template <int L, int PW, int HSM, int XM, int U, class Out>
__device__ __forceinline__ void afb_stage(const AfbParams& p, const float* s0, float2 (&w)[L][2], bool emit, Out& out) {
  using C = AfbCfg<L, PW, HSM, XM>;
  float xa[4 * C::NV], xb[4 * C::NV];
#pragma unroll
  for (int q = 0; q < C::NV; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(s0 + 4 * q);
    const float4 b = *reinterpret_cast<const float4*>(s0 + C::SW + 4 * q);
    xa[4 * q] = a.x; xa[4 * q + 1] = a.y; xa[4 * q + 2] = a.z; xa[4 * q + 3] = a.w;
    xb[4 * q] = b.x; xb[4 * q + 1] = b.y; xb[4 * q + 2] = b.z; xb[4 * q + 3] = b.w;
  }
  // window entries are {row-lowpass, row-highpass} pairs: one packed FMA per tap feeds both
  constexpr int SA = (2 * U) % L, SB = (2 * U + 1) % L;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    float2 ra = make_float2(0.f, 0.f), rb = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const float2 f = make_float2(p.fw_lo.t[j], p.fw_hi.t[j]);
      ra = ffma2_s(xa[C::OFF + 2 * o + j], f, ra);
      rb = ffma2_s(xb[C::OFF + 2 * o + j], f, rb);
    }
    w[SA][o] = ra;
    w[SB][o] = rb;
  }
  if (emit) {
    float2 lo[2], hi[2];  // column low-pass of {l, h} -> {ll, hl}; column high-pass -> {lh, hh}
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < L; ++j) {
        constexpr int base = 2 * U + 2;
        a0 = ffma2_s(p.fh_lo.t[j], w[(base + j) % L][o], a0);
        a1 = ffma2_s(p.fh_hi.t[j], w[(base + j) % L][o], a1);
      }
      lo[o] = a0; hi[o] = a1;
    }
    out.row(lo[0], lo[1], hi[0], hi[1]);
  }
}

template <int L, int PW, int HSM, int XM, int V, class Out>
__device__ __forceinline__ void afb_stage_dispatch(int vv, const AfbParams& p, const float* s0, float2 (&w)[L][2],
                                                   int h0, int h_emit_end, Out& out) {
  using C = AfbCfg<L, PW, HSM, XM>;
  if constexpr (V < C::UNS) {
    if (vv == V) {
      // h0 = index of this stage's first half-stage; output rows are emitted for PRO <= h < h_emit_end
      afb_stage<L, PW, HSM, XM, C::HS * V>(p, s0, w, h0 >= C::PRO && h0 < h_emit_end, out);
      if constexpr (C::HS >= 2)
        afb_stage<L, PW, HSM, XM, C::HS * V + 1>(p, s0 + 2 * C::SW, w, h0 + 1 >= C::PRO && h0 + 1 < h_emit_end, out);
      if constexpr (C::HS == 4) {
        afb_stage<L, PW, HSM, XM, C::HS * V + 2>(p, s0 + 4 * C::SW, w, h0 + 2 >= C::PRO && h0 + 2 < h_emit_end, out);
        afb_stage<L, PW, HSM, XM, C::HS * V + 3>(p, s0 + 6 * C::SW, w, h0 + 3 >= C::PRO && h0 + 3 < h_emit_end, out);
      }
    } else {
      afb_stage_dispatch<L, PW, HSM, XM, V + 1>(vv, p, s0, w, h0, h_emit_end, out);
    }
  }
}

// strip0 / n_strips: the 64-column strips this launch covers (PW == 32), or the single remainder strip
// starting at output column k_rem (PW < 32, n_strips == 1).
template <int L, int PW, int MINB, int HSM, int XM>
__global__ void __launch_bounds__(32, MINB) afb2d_stream(const __grid_constant__ AfbParams p, int n_strips, int n_chunks,
                                                   int CH, int k_rem, int swid) {
  using C = AfbCfg<L, PW, HSM, XM>;
  extern __shared__ __align__(16) float ring[];  // this warp's staging ring
  const int lane = threadIdx.x;
  long long item = blockIdx.x;                    // one warp per CTA: no intra-CTA load imbalance
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int pgroup = (int)(item / n_chunks);
  const int g = lane / PW, jp = lane % PW;        // plane within the group, column pair within the plane
  const int plane0 = pgroup * C::G;
  const int nplanes = imin(C::G, p.planes - plane0);
  const int plane = plane0 + g;

  // swid = output columns per strip (even, <= 64)
  const int k0 = (PW == 32) ? strip * swid : k_rem;
  const int ky0 = chunk * CH;
  const int ky1 = imin(ky0 + CH, p.Ho);
  const int n_half = (ky1 - ky0) + C::PRO;             // half-stages: PRO of warm-up, then one output row each
  const int n_stage = (n_half + C::HS - 1) / C::HS;
  const int nvalid = imin((PW == 32) ? swid : 2 * PW, p.Wo - k0);

  const int sh = (XM == 0 && PW == 32) ? widen_left(2 * k0 - C::HLA, C::HLA + 2 * nvalid + C::RH, p.W, p.mode,
                                                         C::SW - 4 * C::NV - 4 * ((nvalid + 1) / 2 - 1)) : 0;
  typename C::Loader ld;
  ld.init(ring, p.x + (long long)plane0 * p.xps, p.xps, nplanes, p.H, p.W, p.xpitch, p.mode, 2 * k0 - C::HLA - sh,
          C::HLA + 2 * nvalid + C::RH + sh, 2 * ky0 - C::PL, n_stage, lane);
  ld.prologue();

  float2 w[L][2];
#pragma unroll
  for (int j = 0; j < L; ++j) { w[j][0] = w[j][1] = make_float2(0.f, 0.f); }

  DirectOut out;
  const int hipitch = p.hipitch > 0 ? p.hipitch : p.Wo;
  out.band = (long long)p.Ho * hipitch;
  out.ll_ptr = p.ll + (long long)plane * p.llps + (long long)ky0 * p.llpitch + k0 + 2 * jp;
  out.hi_ptr = p.highs + (long long)plane * 3 * out.band + (long long)ky0 * hipitch + k0 + 2 * jp;
  out.nv = (g < nplanes) ? imax(0, imin(2, k0 + nvalid - (k0 + 2 * jp))) : 0;
  out.llpitch = p.llpitch;
  out.Wo = hipitch;
  out.init_parity();
  const int lane_off = g * (C::RPS * C::SW) + ((sh > 0 && out.nv == 0) ? 0 : 4 * jp + sh);

  int vv = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    const float* stage = ld.acquire(t);
    ld.issue(t + C::NS - 1);
    afb_stage_dispatch<L, PW, HSM, XM, 0>(vv, p, stage + lane_off, w, C::HS * t, n_half, out);
    vv = (vv + 1 == C::UNS) ? 0 : vv + 1;
  }
  cp_async_wait<0>();
}

template <int L, int PW, int MINB, int HSM, int XM = 0>
inline void launch_afb_kernel(const AfbParams& p, cudaStream_t stream, long long blocks, int n_strips, int n_chunks,
                              int CH, int k_rem) {
  using C = AfbCfg<L, PW, HSM, XM>;
  // strips are 64 columns wide; splitting the columns evenly over the strips instead (g_tune_balanced) was
  // measured 5 % slower (more row segments that straddle 128-byte lines)
  const int swid = g_tune_balanced ? (p.Wo + 2 * n_strips - 1) / (2 * n_strips) * 2 : 64;
  afb2d_stream<L, PW, MINB, HSM, XM><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH, k_rem, swid);
}

template <int L, int PW>
inline int launch_afb_part(const AfbParams& p, cudaStream_t stream, int n_strips, int k_rem) {
  constexpr int G = 32 / PW;
  const long long groups = ((long long)p.planes + G - 1) / G;
  int n_chunks, CH;
  static const int conc = resident_warps(afb2d_stream<L, PW, 1, 2, 0>, AfbCfg<L, PW, 2, 0>::SMEM_BYTES);
  pick_chunks(groups * n_strips, p.Ho, 16, (L - 2) / 2 + 8, conc, &n_chunks, &CH);
  const long long blocks = groups * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return B200W_ESIZE;
  if (p.mode == B200W_MODE_PERIODIZATION) {
    launch_afb_kernel<L, PW, 1, 2, 2>(p, stream, blocks, n_strips, n_chunks, CH, k_rem);
    return 0;
  }
  if (p.mode == B200W_MODE_PERIODIC) {
    launch_afb_kernel<L, PW, 1, 2, 1>(p, stream, blocks, n_strips, n_chunks, CH, k_rem);
    return 0;
  }
  // measured and not kept: register caps (__launch_bounds__(32, 18..32): a little faster on the small levels, 15 % slower
  // on the large one) and 8-row stages (1.97 vs 1.91 ms) -- profiles/r01_notes.md
  launch_afb_kernel<L, PW, 1, 2>(p, stream, blocks, n_strips, n_chunks, CH, k_rem);
  return 0;
}

template <int L>
inline int launch_afb_stream(const AfbParams& p, cudaStream_t stream) {
  // aligned 16-byte staging needs an aligned source; anything else takes the generic kernel
  if (!aligned_plane(p.x, p.xps, p.xpitch)) return kNoFastPath;
  // Every 64-column strip (including a narrow last one) is an ordinary warp item.  Measured on B200: the
  // kernel is latency/occupancy-bound, not issue-bound, so a mostly-idle last strip costs almost nothing,
  // while packing it across planes (AfbCfg<L, PW<32>, kept for reference) needs a second launch that is
  // slower than what it saves (profiles/r01_notes.md).
  const int n_strips = (p.Wo + 63) / 64;
  return launch_afb_part<L, 32>(p, stream, n_strips, 0);
}

inline int try_launch_afb(const AfbParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
  if (p.Lw != p.Lh) return kNoFastPath;
  if (p.planes == 0) return 0;
  switch (p.Lw) {
    case 2: return launch_afb_stream<2>(p, stream);
    case 4: return launch_afb_stream<4>(p, stream);
    case 6: return launch_afb_stream<6>(p, stream);
    case 8: return launch_afb_stream<8>(p, stream);
    case 10: return launch_afb_stream<10>(p, stream);
    case 12: return launch_afb_stream<12>(p, stream);
    case 14: return launch_afb_stream<14>(p, stream);
    case 16: return launch_afb_stream<16>(p, stream);
    case 18: return launch_afb_stream<18>(p, stream);
    case 20: return launch_afb_stream<20>(p, stream);
    default: return kNoFastPath;
  }
}

#include "fast_dtcwt.cuh"
#include "fast_inverse.cuh"

}  // namespace fast
}  // namespace b200w
