// dwt_pyramid.cuh -- fused multi-level DWT analysis (sm_100a): all J levels of DWTForward in ONE launch, no
// inter-level low-pass in HBM, inputs staged by the TMA engine (cp.async.bulk, SASS UBLKCP.S.G), outputs written by
// the TMA engine (cp.async.bulk.global.shared::cta, SASS UBLKCP.G.S) from shared-memory staging rings laid out
// at the global phase.  Design, index rules and the shared-memory plan: pyramid_plan.h.
//
// Arithmetic = afb2d_stream's (stream_common.cuh / afb_stream.cuh): out[k] = sum_j f[j] xe[2k + j - (L-2)] with
// the taps in stored order and packed IEEE FMAs (FFMA2), W pass then H pass -- bit-identical to the per-level
// kernels and to the oracle.
#pragma once
#include "pyramid_plan.h"
#include "stream_common.cuh"

namespace b200w {
namespace fast {

template <int L>
struct PyrCfg {
  static constexpr int NC = kPyrNC;
  static constexpr int PL = L - 2;
  static constexpr int PRO = PL / 2;
  static constexpr int UNR = L / 2;
  static constexpr int HS = pyr_hs(L);      // output rows per stage
  static constexpr int RS = 2 * HS;         // extended input rows per stage
  static constexpr int HALO = pyr_halo(L);
  static constexpr int NX = 2 * NC + PL;    // floats a lane reads per staged row
  static_assert(HS >= 4 && HS % 2 == 0 && HS % UNR == 0, "stage = whole window periods, even");
};

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_store_s2g(float* gdst, unsigned ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_load_g2s(unsigned sdst, const float* gsrc, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sdst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

// 128-byte phase (in floats) of a band's first element in global memory: staging index = (position + phase) % cap
__device__ __forceinline__ int pyr_phase(const float* band_base) {
  return (int)((reinterpret_cast<uintptr_t>(band_base) >> 2) & 31);
}
// global base of band b of level l for this plane (b == 3: the final low-pass)
__device__ __forceinline__ float* pyr_band_base(const PyrParams& p, int l, int b, int plane) {
  const PyrLevel& v = p.lv[l];
  const long long band = (long long)v.Ho * v.Wo;
  if (b == 3) return p.yl + (long long)plane * band;
  return p.highs[l] + ((long long)plane * 3 + b) * band;
}

// ================================================================================================
// producer warp: streams the extended rows of the input plane into the level-0 ring, kPyrNSlot slots of HS rows
// ================================================================================================
template <int L>
__device__ __forceinline__ void pyr_producer(const PyrParams& p, int plane, float* smem, unsigned bar0, int lane) {
  using C = PyrCfg<L>;
  const PyrLevel& v = p.lv[0];
  const float* src_plane = p.x + (long long)plane * p.xps;
  const int n_slots = 2 * v.n_stage;
  const unsigned row_bytes = (unsigned)v.W * 4u;
  float* ring = smem + v.in_off;
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
#pragma unroll 1
  for (int q = 0; q < n_slots; ++q) {
    const int slot = q % kPyrNSlot;
    const int use = q / kPyrNSlot;
    if (use > 0) mbar_wait(bar0 + 8 * (v.bar_in + v.n_in + slot), (unsigned)((use - 1) & 1));
    // row j of the slot holds extended row e = q*HS + j - PL
    int src = -1;
    if (lane < C::HS) src = ext_index(q * C::HS + lane - C::PL, v.H, p.mode);
    const unsigned valid = __ballot_sync(0xffffffffu, src >= 0);
    unsigned zero_rows = ((1u << C::HS) - 1u) & ~valid;
    while (zero_rows) {   // zero padding above / below the image: plain stores, published by the arrive below
      const int j = __ffs(zero_rows) - 1;
      zero_rows &= zero_rows - 1;
      float* dst = ring + (slot * C::HS + j) * v.in_pitch + C::HALO;
      for (int i = lane * 4; i < v.W; i += 128) *reinterpret_cast<float4*>(dst + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const unsigned full = bar0 + 8 * (v.bar_in + slot);
    if (lane == 0) mbar_arrive_expect_tx(full, (unsigned)__popc(valid) * row_bytes);
    __syncwarp();
    if (src >= 0)
      bulk_load_g2s(ring_s + 4u * (unsigned)((slot * C::HS + lane) * v.in_pitch + C::HALO),
                    src_plane + (long long)src * p.xpitch, row_bytes, full);
  }
}

// ================================================================================================
// writer warp: flushes finished staging groups of every level to HBM with bulk stores
// ================================================================================================
template <int L>
__device__ __forceinline__ void pyr_writer(const PyrParams& p, int plane, float* smem, unsigned bar0, int lane) {
  using C = PyrCfg<L>;
  int next_g[kPyrMaxLevels];
#pragma unroll
  for (int l = 0; l < kPyrMaxLevels; ++l) next_g[l] = 0;
  int prev_l = -1, prev_g = 0;      // last event whose bulk reads are not yet known to be complete
  int remaining = 0;
  for (int l = 0; l < p.J; ++l) remaining += p.lv[l].n_stage;
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);

  while (remaining > 0) {
    bool any = false;
#pragma unroll
    for (int l = 0; l < kPyrMaxLevels; ++l) {
      if (l >= p.J) break;
      const PyrLevel& v = p.lv[l];
      const int g = next_g[l];
      if (g >= v.n_stage) continue;
      if (!__all_sync(0xffffffffu, mbar_test(bar0 + 8 * (v.bar_out + g % kPyrNGO), (unsigned)((g / kPyrNGO) & 1)))) continue;
      any = true;
      // rows [k0, k1) of the level = stream positions [s0, s1) of every band
      const int k0 = imax(0, g * C::HS - C::PRO), k1 = imin(v.Ho, (g + 1) * C::HS - C::PRO);
      const int s0 = k0 * v.Wo, s1 = k1 * v.Wo;
      // lanes 0..3: the 16-byte aligned middle of one band each; lanes 8..31: head / tail elements
      const int b = (lane < 4) ? lane : (lane - 8) / 6;
      if (lane != 4 && lane != 5 && lane != 6 && lane != 7 && b < v.nbands) {
        float* gb = pyr_band_base(p, l, b, plane);
        const int a = pyr_phase(gb);
        const int hd = imin(s1 - s0, (4 - ((s0 + a) & 3)) & 3);
        const int m0 = s0 + hd;
        const int m1 = m0 + ((s1 - m0) & ~3);
        const unsigned st_s = smem_s + 4u * (unsigned)(v.st_off + b * v.st_cap);
        if (lane < 4) {
          const int n = m1 - m0;
          if (n > 0) {
            const int i0 = (m0 + a) % v.st_cap;
            const int first = imin(n, v.st_cap - i0);
            bulk_store_s2g(gb + m0, st_s + 4u * (unsigned)i0, 4u * (unsigned)first);
            if (n > first) bulk_store_s2g(gb + m0 + first, st_s, 4u * (unsigned)(n - first));
          }
        } else {
          const int e = (lane - 8) % 6;
          int pos = -1;
          if (e < 3) { if (e < hd) pos = s0 + e; }
          else if (e - 3 < s1 - m1) pos = m1 + (e - 3);
          if (pos >= 0) gb[pos] = lds_s(st_s + 4u * (unsigned)((pos + a) % v.st_cap));
        }
      }
      bulk_commit();
      // everything but this event's own bulk group has been read out of shared memory: release the previous event
      bulk_wait_read<1>();
      __syncwarp();
      if (prev_l >= 0 && lane == 0) mbar_arrive(bar0 + 8 * (p.lv[prev_l].bar_out + kPyrNGO + prev_g % kPyrNGO));
      prev_l = l; prev_g = g;
      next_g[l] = g + 1;
      --remaining;
    }
    if (!any) {
      if (prev_l >= 0) {
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (p.lv[prev_l].bar_out + kPyrNGO + prev_g % kPyrNGO));
        prev_l = -1;
      } else {
        __nanosleep(64);
      }
    }
  }
  bulk_wait<0>();   // all bulk stores of this plane are complete before the CTA retires
}

// ================================================================================================
// level worker
// ================================================================================================
template <int L>
__device__ __forceinline__ void pyr_rowpass(const PyrParams& p, const float* row, float2 (&dst)[PyrCfg<L>::NC]) {
  using C = PyrCfg<L>;
  float x[C::NX];
#pragma unroll
  for (int q = 0; q < C::NX / 2; ++q) {
    const float2 v = *reinterpret_cast<const float2*>(row + 2 * q);
    x[2 * q] = v.x; x[2 * q + 1] = v.y;
  }
#pragma unroll
  for (int o = 0; o < C::NC; ++o) {
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < L; ++j) r = ffma2_s(x[2 * o + j], make_float2(p.fw_lo.t[j], p.fw_hi.t[j]), r);
    dst[o] = r;
  }
}

template <int L>
struct PyrEmit {   // where the finished rows of one lane go
  unsigned st_s[4];     // staging band bases (shared-window byte addresses)
  int sidx[4];          // staging float index of (current row, c0) per band
  int cap, Wo, nv, c0, nb;
  unsigned nr_s;        // next level's ring: byte address of (row 0, column c0); 0 for the last level
  int nr_pitch4, nr_rows, nr_slot;   // pitch in bytes, ring depth, slot of the current row
  int Wn, mode;         // next level's input width (= Wo), extension mode (for the halo copies)
  bool edge;

  __device__ __forceinline__ void put(int b, int i, float v) const {
    int e = sidx[b] + i;
    if (e >= cap) e -= cap;
    sts_s(st_s[b] + 4u * (unsigned)e, v);
  }
  // lo[o] = {ll, hl}, hi[o] = {lh, hh} of column c0 + o
  __device__ __forceinline__ void row(const float2 (&lo)[kPyrNC], const float2 (&hi)[kPyrNC]) {
    using C = PyrCfg<L>;
#pragma unroll
    for (int o = 0; o < kPyrNC; ++o) {
      if (o < nv) {
        put(0, o, hi[o].x);
        put(1, o, lo[o].y);
        put(2, o, hi[o].y);
        if (nb == 4) put(3, o, lo[o].x);
      }
    }
    if (nr_s != 0) {
      const unsigned rb = nr_s + (unsigned)(nr_slot * nr_pitch4);
#pragma unroll
      for (int o = 0; o < kPyrNC; ++o)
        if (o < nv) sts_s(rb + 4u * o, lo[o].x);
      if (edge) {   // copies into the halo cells the extension maps onto this column (next level's W-pass border)
#pragma unroll
        for (int o = 0; o < kPyrNC; ++o) {
          if (o < nv) {
            const int c = c0 + o;
            if (mode == B200W_MODE_SYMMETRIC) {
              if (c < C::PL) sts_s(rb + 4u * o - 4u * (unsigned)(2 * c + 1), lo[o].x);             // cell -1-c
              if (c >= Wn - (L - 1)) sts_s(rb + 4u * o + 4u * (unsigned)(2 * (Wn - c) - 1), lo[o].x);  // cell 2Wn-1-c
            } else if (mode == B200W_MODE_REFLECT) {
              if (c >= 1 && c <= C::PL) sts_s(rb + 4u * o - 4u * (unsigned)(2 * c), lo[o].x);        // cell -c
              if (c <= Wn - 2 && c >= Wn - L) sts_s(rb + 4u * o + 4u * (unsigned)(2 * (Wn - 1 - c)), lo[o].x);  // 2Wn-2-c
            }
          }
        }
      }
      nr_slot = (nr_slot + 1 == nr_rows) ? 0 : nr_slot + 1;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      sidx[b] += Wo;
      if (sidx[b] >= cap) sidx[b] -= cap;
    }
  }
};

template <int L>
__device__ __forceinline__ void pyr_worker(const PyrParams& p, int plane, int lvl, int wl, int lane, float* smem,
                                           unsigned bar0) {
  using C = PyrCfg<L>;
  const PyrLevel& v = p.lv[lvl];
  const bool last = (lvl == p.J - 1);
  const int c0 = C::NC * (32 * wl + lane);
  const int nv = imax(0, imin(C::NC, v.Wo - c0));
  const int rd = (nv > 0) ? (C::HALO + 2 * c0 - C::PL) : C::HALO;   // lane's read offset inside a ring row
  const float* in_ring = smem + v.in_off;
  const float* zero_row = smem + p.zero_off;
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);

  PyrEmit<L> em;
  em.cap = v.st_cap; em.Wo = v.Wo; em.nv = nv; em.c0 = c0; em.nb = v.nbands; em.mode = p.mode;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    em.st_s[b] = smem_s + 4u * (unsigned)(v.st_off + (b < v.nbands ? b : 0) * v.st_cap);
    em.sidx[b] = (b < v.nbands) ? (c0 + pyr_phase(pyr_band_base(p, lvl, b, plane))) % v.st_cap : 0;
  }
  em.nr_s = 0; em.nr_pitch4 = 0; em.nr_rows = 1; em.nr_slot = 0; em.Wn = v.Wo; em.edge = false;
  int nx_bar = 0, nx_n = 1, nx_warps_first = 0;
  if (!last) {
    const PyrLevel& u = p.lv[lvl + 1];
    em.nr_s = smem_s + 4u * (unsigned)(u.in_off + C::HALO + c0);
    em.nr_pitch4 = 4 * u.in_pitch;
    em.nr_rows = u.in_rows;
    // this warp's columns map onto halo cells if they touch either end of the row
    const int cw0 = C::NC * 32 * wl, cw1 = imin(cw0 + C::NC * 32, v.Wo);
    em.edge = (p.mode != B200W_MODE_ZERO) && (cw0 <= C::PL || cw1 > v.Wo - L);
    nx_bar = u.bar_in; nx_n = u.n_in;
  }
  (void)nx_warps_first;
  // level 0: does this warp read the left / right halo cells of the staged input rows?
  const int cw0 = C::NC * 32 * wl, cw1 = imin(cw0 + C::NC * 32, v.Wo);
  const bool patch_l = (lvl == 0) && (p.mode != B200W_MODE_ZERO) && (2 * cw0 - C::PL < 0);
  const bool patch_r = (lvl == 0) && (p.mode != B200W_MODE_ZERO) && (2 * (cw1 - 1) + 1 >= v.W);
  // lane -> one halo cell of a staged row: lanes [0, PL) the left cells -1..-PL, lanes [PL, PL + L - 1) the right ones
  int patch_dst = -1, patch_src = 0;
  if (patch_l && lane < C::PL) {
    const int j = lane + 1;
    patch_dst = C::HALO - j;
    patch_src = C::HALO + ((p.mode == B200W_MODE_SYMMETRIC) ? j - 1 : j);
  } else if (patch_r && lane >= C::PL && lane < C::PL + L - 1) {
    const int i = lane - C::PL;
    patch_dst = C::HALO + v.W + i;
    patch_src = C::HALO + ((p.mode == B200W_MODE_SYMMETRIC) ? v.W - 1 - i : v.W - 2 - i);
  }
  const bool do_patch = patch_l || patch_r;

  float2 w[L][C::NC];
#pragma unroll
  for (int j = 0; j < L; ++j)
#pragma unroll
    for (int o = 0; o < C::NC; ++o) w[j][o] = make_float2(0.f, 0.f);

  int g_seen = 0, g_rel = 0;   // levels >= 1: input groups waited for / released so far
  const int prev_stages = (lvl > 0) ? p.lv[lvl - 1].n_stage : 0;

#pragma unroll 1
  for (int t = 0; t < v.n_stage; ++t) {
    // ---- room for this stage's output group ---------------------------------------------------------------
    if (t >= kPyrNGO) mbar_wait(bar0 + 8 * (v.bar_out + kPyrNGO + t % kPyrNGO), (unsigned)((t / kPyrNGO - 1) & 1));
    if (!last && t >= nx_n) mbar_wait(bar0 + 8 * (nx_bar + nx_n + t % nx_n), (unsigned)((t / nx_n - 1) & 1));
    // ---- inputs ----------------------------------------------------------------------------------------------
    const float* slot_rows = nullptr;   // level 0: first row of the current input slot
    if (lvl > 0) {
      const int g_need = imin(pyr_group_of_row(pyr_stage_max_row(t, C::RS, C::PL, v.H, p.mode), C::HS, C::PRO),
                              prev_stages - 1);
      while (g_seen <= g_need) {
        mbar_wait(bar0 + 8 * (v.bar_in + g_seen % v.n_in), (unsigned)((g_seen / v.n_in) & 1));
        ++g_seen;
      }
    }
#pragma unroll
    for (int hh = 0; hh < C::HS; ++hh) {
      if (lvl == 0 && (hh == 0 || hh == C::HS / 2)) {
        const int q = 2 * t + (hh ? 1 : 0);
        if (hh) {   // done with the first slot of the stage
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (v.bar_in + v.n_in + (q - 1) % kPyrNSlot));
        }
        const int slot = q % kPyrNSlot;
        mbar_wait(bar0 + 8 * (v.bar_in + slot), (unsigned)((q / kPyrNSlot) & 1));
        slot_rows = in_ring + slot * C::HS * v.in_pitch;
        if (do_patch) {
          if (patch_dst >= 0) {
            const unsigned a = (unsigned)__cvta_generic_to_shared(slot_rows);
#pragma unroll
            for (int j = 0; j < C::HS; ++j)
              sts_s(a + 4u * (unsigned)(j * v.in_pitch + patch_dst), lds_s(a + 4u * (unsigned)(j * v.in_pitch + patch_src)));
          }
          __syncwarp();
        }
      }
      const float *r0, *r1;
      if (lvl == 0) {
        const int j = 2 * hh - ((hh >= C::HS / 2) ? C::HS : 0);
        r0 = slot_rows + j * v.in_pitch + rd;
        r1 = r0 + v.in_pitch;
      } else {
        const int e = t * C::RS + 2 * hh - C::PL;
        const int s0 = ((unsigned)e < (unsigned)v.H) ? e : ext_index_cold(e, v.H, p.mode);
        const int s1 = ((unsigned)(e + 1) < (unsigned)v.H) ? e + 1 : ext_index_cold(e + 1, v.H, p.mode);
        r0 = (s0 >= 0 ? in_ring + (s0 % v.in_rows) * v.in_pitch : zero_row) + rd;
        r1 = (s1 >= 0 ? in_ring + (s1 % v.in_rows) * v.in_pitch : zero_row) + rd;
      }
      pyr_rowpass<L>(p, r0, w[(2 * hh) % L]);
      pyr_rowpass<L>(p, r1, w[(2 * hh + 1) % L]);
      const int k = t * C::HS + hh - C::PRO;
      if (k >= 0 && k < v.Ho) {
        float2 lo[C::NC], hi[C::NC];
#pragma unroll
        for (int o = 0; o < C::NC; ++o) {
          float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < L; ++j) {
            a0 = ffma2_s(p.fh_lo.t[j], w[(2 * hh + 2 + j) % L][o], a0);
            a1 = ffma2_s(p.fh_hi.t[j], w[(2 * hh + 2 + j) % L][o], a1);
          }
          lo[o] = a0; hi[o] = a1;
        }
        em.row(lo, hi);
      }
    }
    // ---- hand the inputs back, publish the outputs -----------------------------------------------------------
    fence_proxy_async();   // this lane's staging stores become visible to the bulk-store engine
    __syncwarp();
    if (lane == 0) {
      if (lvl == 0) mbar_arrive(bar0 + 8 * (v.bar_in + v.n_in + (2 * t + 1) % kPyrNSlot));
      mbar_arrive(bar0 + 8 * (v.bar_out + t % kPyrNGO));
      if (!last) mbar_arrive(bar0 + 8 * (nx_bar + t % nx_n));
    }
    if (lvl > 0) {
      const int lo_row = pyr_stage_release_bound(t, C::RS, C::PL, v.H, L);
      while (g_rel < prev_stages && pyr_group_end(g_rel, C::HS, C::PRO) <= lo_row) {
        if (lane == 0) mbar_arrive(bar0 + 8 * (v.bar_in + v.n_in + g_rel % v.n_in));
        ++g_rel;
      }
    }
  }
}

// ================================================================================================
// the kernel: one CTA per plane; warp 0 producer, warp 1 writer, the rest level workers
// ================================================================================================
template <int L, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) dwt_pyramid(const __grid_constant__ PyrParams p) {
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.x;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(smem);

  // zero the rings (zero-padding halos, cells nobody writes), then the barriers
  {
    const int n4 = p.smem_bytes / 16;
    float4* s4 = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < n4; i += blockDim.x) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (tid == 0) {
    for (int l = 0; l < p.J; ++l) {
      const PyrLevel& v = p.lv[l];
      const int n_up = (l == 0) ? 1 : p.lv[l - 1].nwarps;       // arrivals that fill an input slot / group
      for (int i = 0; i < v.n_in; ++i) {
        mbar_init(bar0 + 8 * (v.bar_in + i), (unsigned)n_up);
        mbar_init(bar0 + 8 * (v.bar_in + v.n_in + i), (unsigned)v.nwarps);
      }
      for (int i = 0; i < kPyrNGO; ++i) {
        mbar_init(bar0 + 8 * (v.bar_out + i), (unsigned)v.nwarps);
        mbar_init(bar0 + 8 * (v.bar_out + kPyrNGO + i), 1u);
      }
    }
    mbar_fence_init();
  }
  fence_proxy_async();   // the zero fill (generic proxy) is ordered before the TMA writes into the same rings
  __syncthreads();

  if (warp == 0) {
    pyr_producer<L>(p, plane, smem, bar0, lane);
  } else if (warp == 1) {
    pyr_writer<L>(p, plane, smem, bar0, lane);
  } else {
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < kPyrMaxLevels; ++l)
      if (l < p.J && warp >= p.lv[l].warp0) lvl = l;
    pyr_worker<L>(p, plane, lvl, warp - p.lv[lvl].warp0, lane, smem, bar0);
  }
}

template <int L>
inline int launch_pyramid(const PyrParams& p, cudaStream_t stream) {
  if (p.planes <= 0) return 0;
  static int smem_set[64] = {};
  int dev = 0;
  (void)cudaGetDevice(&dev);
  const bool small = p.threads <= 256;
  if (dev < 0 || dev >= 64 || !(smem_set[dev] & (small ? 1 : 2))) {
    cudaError_t e = small ? cudaFuncSetAttribute(dwt_pyramid<L, 256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                          : cudaFuncSetAttribute(dwt_pyramid<L, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return kNoFastPath; }
    if (dev >= 0 && dev < 64) smem_set[dev] |= (small ? 1 : 2);
  }
  if (small) dwt_pyramid<L, 256, 2><<<(unsigned)p.planes, p.threads, p.smem_bytes, stream>>>(p);
  else dwt_pyramid<L, 512, 1><<<(unsigned)p.planes, p.threads, p.smem_bytes, stream>>>(p);
  return 0;
}

}  // namespace fast
}  // namespace b200w
