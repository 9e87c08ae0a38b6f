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

// ---- optional wait-time instrumentation (variant builds only: -DB200W_PYR_PROF) -----------------------------------
#ifdef B200W_PYR_PROF
__device__ unsigned long long g_pyr_prof[64];   // [role (0 producer, 1 writer, 2 + level)][8 counters]
#define PYR_T0() const long long pyr_t0_ = clock64()
#define PYR_ACC(k) prof[k] += (unsigned long long)(clock64() - pyr_t0_)
#define PYR_DECL() unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long pyr_life_ = clock64()
#define PYR_FLUSH(role, lane)                                                                              \
  do {                                                                                                     \
    prof[7] = (unsigned long long)(clock64() - pyr_life_);                                                 \
    if ((lane) == 0)                                                                                       \
      for (int k_ = 0; k_ < 8; ++k_) atomicAdd(&g_pyr_prof[8 * (role) + k_], prof[k_]);                    \
  } while (0)
#else
#define PYR_T0() do {} while (0)
#define PYR_ACC(k) do {} while (0)
#define PYR_DECL() do {} while (0)
#define PYR_FLUSH(role, lane) do {} while (0)
#endif

constexpr int kPyrNSlotC = B200W_PYR_NSLOT;   // the shape the kernel is compiled for (plan_pyramid_best only offers it)

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
// wait used by the lightly loaded roles (producer, levels >= 1): poll + sleep instead of the hardware try_wait loop, whose
// wake-ups (any barrier traffic on the SM) make the idle warps spin through the issue slots the level-1 warps need
__device__ __forceinline__ void mbar_wait_relaxed(unsigned bar, unsigned parity) {
  while (!mbar_test(bar, parity)) __nanosleep(200);
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

// 16-byte phase (in floats) of a band's first element in global memory: staging index = (position + phase) % cap, so
// 16-byte aligned runs in global memory are 16-byte aligned in the staging ring (what the bulk copies need)
__device__ __forceinline__ int pyr_phase(const float* band_base) {
  return (int)((reinterpret_cast<uintptr_t>(band_base) >> 2) & 3);
}
// global base of band b of level l for this plane (b == 3: the final low-pass)
__device__ __forceinline__ float* pyr_band_base(const PyrParams& p, int l, int b, int plane) {
  const PyrLevel& v = p.lv[l];
  const long long band = (long long)v.Ho * v.Wo;
  if (b == 3) return p.yl + (long long)plane * v.Ho * p.ll_pitch;
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
  PYR_DECL();
#pragma unroll 1
  for (int q = 0; q < n_slots; ++q) {
    const int slot = q % kPyrNSlotC;
    const int use = q / kPyrNSlotC;
    if (use > 0) { PYR_T0(); mbar_wait(bar0 + 8 * (v.bar_in + v.n_in + slot), (unsigned)((use - 1) & 1)); PYR_ACC(0); }
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
  PYR_FLUSH(0, lane);
}

// ================================================================================================
// writer warp: flushes finished staging groups of every level to HBM with bulk stores
// ================================================================================================
template <int L, int SPLIT>
__device__ __forceinline__ void pyr_writer(const PyrParams& p, int plane, float* smem, unsigned bar0, int lane) {
  using C = PyrCfg<L>;
  const int SG = C::HS / SPLIT;             // rows per staging group
  const int R = kPyrNGO * SG;                    // rows in a staging ring
  // lanes 0..3: the 16-byte aligned middle of one band each; lanes 8..31: one head / tail element of a band each
  const bool bulk_lane = lane < 4;
  const int myb = bulk_lane ? lane : (lane >= 8 ? (lane - 8) / 6 : 4);
  const int mye = (lane >= 8) ? (lane - 8) % 6 : 0;
  float* gb[kPyrMaxLevels];      // this lane's band in global memory, per level
  unsigned stb[kPyrMaxLevels];   // ... and its staging ring
  int ph[kPyrMaxLevels], next_g[kPyrMaxLevels], rslot[kPyrMaxLevels];
  int remaining = 0;
#pragma unroll
  for (int l = 0; l < kPyrMaxLevels; ++l) {
    next_g[l] = 0; rslot[l] = 0; gb[l] = nullptr; stb[l] = 0; ph[l] = 0;
    if (l < p.J) {
      remaining += p.lv[l].n_stage * SPLIT;
      if (myb < p.lv[l].nbands) {
        gb[l] = pyr_band_base(p, l, myb, plane);
        ph[l] = pyr_phase(gb[l]);
        stb[l] = (unsigned)__cvta_generic_to_shared(smem) + 4u * (unsigned)(p.lv[l].st_off + myb * p.lv[l].st_cap);   // band 3 follows three st_cap bands
      }
    }
  }
  int prev_bar = -1;      // out_empty barrier of the last event whose bulk reads are not yet known to be complete
  PYR_DECL();

  while (remaining > 0) {
    bool any = false;
#pragma unroll
    for (int l = 0; l < kPyrMaxLevels; ++l) {
      if (l >= p.J) break;
      const PyrLevel& v = p.lv[l];
      const int g = next_g[l];
      if (g >= v.n_stage * SPLIT) continue;
      const int bar = v.bar_out + g % kPyrNGO;
      if (!__all_sync(0xffffffffu, mbar_test(bar0 + 8 * bar, (unsigned)((g / kPyrNGO) & 1)))) continue;
      any = true;
      PYR_T0();
      // rows [k0, k1) of the level = stream positions [s0, s1) of every band; the group starts at ring row rslot
      const int k0 = imax(0, g * SG - C::PRO), k1 = imin(v.Ho, (g + 1) * SG - C::PRO);
      if (k1 > k0 && gb[l] != nullptr) {
        const int rowlen = (myb == 3) ? p.ll_pitch : v.Wo;     // the low-pass rows carry their pitch
        const int cap = (myb == 3) ? v.st_cap_ll : v.st_cap;
        const int s0 = k0 * rowlen, s1 = k1 * rowlen, a = ph[l];
        const int hd = imin(s1 - s0, (4 - ((s0 + a) & 3)) & 3);
        const int m0 = s0 + hd;
        const int m1 = m0 + ((s1 - m0) & ~3);
        int i0 = rslot[l] * rowlen + a;               // ring index of position s0 (may exceed the ring by < 4)
        if (bulk_lane) {
          const int n = m1 - m0;
          if (n > 0) {
            int i = i0 + hd;
            if (i >= cap) i -= cap;
            const int first = imin(n, cap - i);
            bulk_store_s2g(gb[l] + m0, stb[l] + 4u * (unsigned)i, 4u * (unsigned)first);
            if (n > first) bulk_store_s2g(gb[l] + m0 + first, stb[l], 4u * (unsigned)(n - first));
          }
        } else {
          int pos = -1;
          if (mye < 3) { if (mye < hd) pos = s0 + mye; }
          else if (mye - 3 < s1 - m1) pos = m1 + (mye - 3);
          if (pos >= 0) {
            int i = i0 + (pos - s0);
            while (i >= cap) i -= cap;
            gb[l][pos] = lds_s(stb[l] + 4u * (unsigned)i);
          }
        }
      }
      if (k1 > k0) { rslot[l] += k1 - k0; if (rslot[l] >= R) rslot[l] -= R; }
      bulk_commit();
      // everything but this event's own bulk group has been read out of shared memory: release the previous event
      bulk_wait_read<1>();
      __syncwarp();
      if (prev_bar >= 0 && lane == 0) mbar_arrive(bar0 + 8 * prev_bar);
      prev_bar = bar + kPyrNGO;
      next_g[l] = g + 1;
      --remaining;
      PYR_ACC(0);   // time spent flushing events
#ifdef B200W_PYR_PROF
      prof[1] += 1;
#endif
    }
    if (!any) {
      if (prev_bar >= 0) {
        PYR_T0();
        bulk_wait_read<0>();
        PYR_ACC(2);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * prev_bar);
        prev_bar = -1;
      } else {
        __nanosleep(100);
      }
    }
  }
  { PYR_T0(); bulk_wait<0>(); PYR_ACC(3); }   // all bulk stores of this plane are complete before the CTA retires
  PYR_FLUSH(1, lane);
}

// ================================================================================================
// level worker
// ================================================================================================
// Shared-memory accesses of the hot loop are volatile asm on shared-window addresses: they keep their program order
// relative to the mbarrier waits / arrives (also volatile asm), and the compiler never has to prove an address is shared.
__device__ __forceinline__ float2 lds64_s(unsigned s) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(s));
  return v;
}
// ... with a compile-time byte offset folded into the instruction (no address arithmetic per access)
template <int OFF>
__device__ __forceinline__ float2 lds64_so(unsigned s) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];\n" : "=f"(v.x), "=f"(v.y) : "r"(s), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts_so(unsigned s, float v) {
  asm volatile("st.shared.f32 [%0+%2], %1;\n" ::"r"(s), "f"(v), "n"(OFF) : "memory");
}
__device__ __forceinline__ uint2 lds64u_s(unsigned s) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(s));
  return v;
}
__device__ __forceinline__ void stsu_s(unsigned s, unsigned v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(s), "r"(v) : "memory"); }

template <int L>
__device__ __forceinline__ void pyr_rowpass(const PyrParams& p, unsigned row_s, float2 (&dst)[PyrCfg<L>::NC]) {
  using C = PyrCfg<L>;
  float x[C::NX];
  {
    float2 v;
#define PYR_LD(Q) if constexpr (Q < C::NX / 2) { v = lds64_so<8 * (Q)>(row_s); x[2 * (Q)] = v.x; x[2 * (Q) + 1] = v.y; }
    PYR_LD(0) PYR_LD(1) PYR_LD(2) PYR_LD(3) PYR_LD(4) PYR_LD(5) PYR_LD(6) PYR_LD(7) PYR_LD(8) PYR_LD(9) PYR_LD(10) PYR_LD(11)
#undef PYR_LD
    static_assert(C::NX / 2 <= 12, "unrolled loads");
  }
#pragma unroll
  for (int o = 0; o < C::NC; ++o) {
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < L; ++j) r = ffma2_s(x[2 * o + j], make_float2(p.fw[2 * j], p.fw[2 * j + 1]), r);
    dst[o] = r;
  }
}

// Where the finished rows of one lane go.  Staging: band b's ring holds R = kPyrNGO * HS rows; row k sits at row slot
// k % R, shifted by the band's global 128-byte phase, so only the last slot's tail wraps to the ring start.
template <int L>
struct PyrEmit {
  unsigned st_s[4];     // staging band bases (shared-window byte addresses)
  unsigned so[4];       // byte offset of (current row slot, c0) inside the band's ring, phase included
  unsigned capb, wob;   // ring size and row advance in bytes (band-pass bands)
  unsigned capb3, wob3; // ... of the low-pass band of the last level (rows of ll_pitch floats)
  int nv, nb, kslot, R;
  unsigned nr_s;        // next level's ring: byte address of (row 0, column c0); 0 for the last level
  unsigned nr_off, nr_pitch4, nr_bytes;   // byte offset of the current row, row pitch, ring size
  int dup[kPyrNC];      // byte delta from a column's cell to the halo cell the extension maps onto it (0: none)

  __device__ __forceinline__ void putn(int b, const float (&v)[kPyrNC], bool wrap_slot) {
    const unsigned a = st_s[b] + so[b];
    if (!wrap_slot) {
      if (nv > 0) sts_so<0>(a, v[0]);
      if (nv > 1) sts_so<4>(a, v[1]);
      if constexpr (kPyrNC > 2) { if (nv > 2) sts_so<8>(a, v[kPyrNC > 2 ? 2 : 0]); }
    } else {
#pragma unroll
      for (int i = 0; i < kPyrNC; ++i) {
        const unsigned cb = (b == 3) ? capb3 : capb;
        unsigned oi = so[b] + 4u * i;
        if (oi >= cb) oi -= cb;
        if (nv > i) sts_s(st_s[b] + oi, v[i]);
      }
    }
  }
  // lo[o] = {ll, hl}, hi[o] = {lh, hh} of column c0 + o
  __device__ __forceinline__ void row(const float2 (&lo)[kPyrNC], const float2 (&hi)[kPyrNC]) {
    float v0[kPyrNC], v1[kPyrNC], v2[kPyrNC], v3[kPyrNC];
#pragma unroll
    for (int i = 0; i < kPyrNC; ++i) { v0[i] = hi[i].x; v1[i] = lo[i].y; v2[i] = hi[i].y; v3[i] = lo[i].x; }
    const bool wrap_slot = (kslot == R - 1);
    putn(0, v0, wrap_slot);
    putn(1, v1, wrap_slot);
    putn(2, v2, wrap_slot);
    if (nb == 4) putn(3, v3, wrap_slot);
    // the slot after the last one is slot 0
#pragma unroll
    for (int b = 0; b < 3; ++b) so[b] = so[b] + wob - (wrap_slot ? capb : 0u);
    so[3] = so[3] + wob3 - (wrap_slot ? capb3 : 0u);
    kslot = wrap_slot ? 0 : kslot + 1;
    if (nr_s != 0) {
      const unsigned rb = nr_s + nr_off;
      if (nv > 0) sts_so<0>(rb, v3[0]);
      if (nv > 1) sts_so<4>(rb, v3[1]);
      if constexpr (kPyrNC > 2) { if (nv > 2) sts_so<8>(rb, v3[kPyrNC > 2 ? 2 : 0]); }
#pragma unroll
      for (int i = 0; i < kPyrNC; ++i)
        if (dup[i] != 0) sts_s(rb + 4u * i + (unsigned)dup[i], v3[i]);   // the W-pass border of the next level
      nr_off += nr_pitch4;
      if (nr_off >= nr_bytes) nr_off = 0;
    }
  }
};

template <int L, int SPLIT>
__device__ __forceinline__ void pyr_worker(const PyrParams& p, int plane, int lvl, int wl, int warp, int lane,
                                           float* smem, unsigned bar0) {
  using C = PyrCfg<L>;
  const PyrLevel& v = p.lv[lvl];
  const bool last = (lvl == p.J - 1);
  const int c0 = C::NC * (32 * wl + lane);
  const int nv = imax(0, imin(C::NC, v.Wo - c0));
  const unsigned rd4 = 4u * (unsigned)((nv > 0) ? (C::HALO + 2 * c0 - C::PL) : C::HALO);   // read offset in a ring row
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned ring_s = smem_s + 4u * (unsigned)v.in_off;
  const unsigned zero_s = smem_s + 4u * (unsigned)p.zero_off;
  const unsigned tab_s = smem_s + 4u * (unsigned)(p.tab_off + 32 * warp);   // this warp's row-address table

  PyrEmit<L> em;
  em.capb = 4u * (unsigned)v.st_cap; em.wob = 4u * (unsigned)v.Wo; em.nv = nv; em.nb = v.nbands;
  em.capb3 = 4u * (unsigned)v.st_cap_ll; em.wob3 = 4u * (unsigned)p.ll_pitch;
  em.kslot = 0; em.R = kPyrNGO * (C::HS / SPLIT);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    em.st_s[b] = smem_s + 4u * (unsigned)(v.st_off + (b < v.nbands ? b : 0) * v.st_cap);
    em.so[b] = (b < v.nbands) ? 4u * (unsigned)(c0 + pyr_phase(pyr_band_base(p, lvl, b, plane))) : 0u;
  }
  em.nr_s = 0; em.nr_off = 0; em.nr_pitch4 = 0; em.nr_bytes = 1;
#pragma unroll
  for (int i = 0; i < C::NC; ++i) em.dup[i] = 0;
  int nx_bar = 0, nx_n = 1;
  if (!last) {
    const PyrLevel& u = p.lv[lvl + 1];
    em.nr_s = smem_s + 4u * (unsigned)(u.in_off + C::HALO + c0);
    em.nr_pitch4 = 4u * (unsigned)u.in_pitch;
    em.nr_bytes = em.nr_pitch4 * (unsigned)u.in_rows;
    nx_bar = u.bar_in; nx_n = u.n_in;
    // halo cells of the next level's rows that the extension maps onto this lane's columns (the plan guarantees
    // Wo >= 2L - 2, so a column feeds at most one halo cell)
    const int Wn = v.Wo;
#pragma unroll
    for (int i = 0; i < C::NC; ++i) {
      const int c = c0 + i;
      int cell = c;   // target cell index (column coordinates); == c means none
      if (i < nv) {
        if (p.mode == B200W_MODE_SYMMETRIC) {
          if (c < C::PL) cell = -1 - c;
          else if (c >= Wn - (L - 1)) cell = 2 * Wn - 1 - c;
        } else if (p.mode == B200W_MODE_REFLECT) {
          if (c >= 1 && c <= C::PL) cell = -c;
          else if (c <= Wn - 2 && c >= Wn - L) cell = 2 * Wn - 2 - c;
        }
      }
      em.dup[i] = 4 * (cell - c);
    }
  }
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

  PYR_DECL();
  int g_seen = 0, g_rel = 0;   // levels >= 1: input groups waited for / released so far
  const int prev_stages = (lvl > 0) ? p.lv[lvl - 1].n_stage : 0;
  const int pitch4 = 4 * v.in_pitch;
  constexpr int split = SPLIT, nslot = kPyrNSlotC;

#pragma unroll 1
  for (int t = 0; t < v.n_stage; ++t) {
    // ---- the stage's RS row addresses, one per lane, into this warp's table ----------------------------------
    if (lane < C::RS) {
      unsigned a;
      if (lvl == 0) {
        const int slot = (2 * t + (lane >= C::HS ? 1 : 0)) % nslot;
        a = ring_s + (unsigned)((slot * C::HS + (lane >= C::HS ? lane - C::HS : lane)) * pitch4);
      } else {
        const int e = t * C::RS + lane - C::PL;
        const int src = ((unsigned)e < (unsigned)v.H) ? e : ext_index_cold(e, v.H, p.mode);
        a = (src >= 0) ? ring_s + (unsigned)((src % v.in_rows) * pitch4) : zero_s;
      }
      stsu_s(tab_s + 4u * lane, a);
    }
    // ---- room for this stage's rows in the next level's ring (the staging groups are claimed inside the stage) ---
    if (!last && t >= nx_n) { PYR_T0(); mbar_wait(bar0 + 8 * (nx_bar + nx_n + t % nx_n), (unsigned)((t / nx_n - 1) & 1)); PYR_ACC(2); }
    // ---- inputs ----------------------------------------------------------------------------------------------
    if (lvl > 0) {
      const int g_need = imin(pyr_group_of_row(pyr_stage_max_row(t, C::RS, C::PL, v.H, p.mode), C::HS, C::PRO),
                              prev_stages - 1);
      PYR_T0();
      while (g_seen <= g_need) {
        mbar_wait_relaxed(bar0 + 8 * (v.bar_in + g_seen % v.n_in), (unsigned)((g_seen / v.n_in) & 1));
        ++g_seen;
      }
      PYR_ACC(0);
    }
    __syncwarp();   // the table
    const int k_first = t * C::HS - C::PRO;
#pragma unroll
    for (int hh = 0; hh < C::HS; ++hh) {
      if (hh == 0 || (split == 2 && hh == C::HS / 2)) {   // a new staging group: its ring rows must have been flushed
        const int sg = t * split + (hh ? 1 : 0);
        if (sg >= kPyrNGO) {
          PYR_T0();
          mbar_wait(bar0 + 8 * (v.bar_out + kPyrNGO + sg % kPyrNGO), (unsigned)((sg / kPyrNGO - 1) & 1));
          PYR_ACC(1);
        }
      }
      if (lvl == 0 && (hh == 0 || hh == C::HS / 2)) {
        const int q = 2 * t + (hh ? 1 : 0);
        if (hh) {   // done with the first slot of the stage
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (v.bar_in + v.n_in + (q - 1) % nslot));
        }
        const int slot = q % nslot;
        { PYR_T0(); mbar_wait(bar0 + 8 * (v.bar_in + slot), (unsigned)((q / nslot) & 1)); PYR_ACC(0); }
        if (do_patch) {
          PYR_T0();
          if (patch_dst >= 0) {
            const unsigned a = ring_s + (unsigned)(slot * C::HS * pitch4);
#pragma unroll
            for (int j = 0; j < C::HS; ++j)
              sts_s(a + (unsigned)(j * pitch4 + 4 * patch_dst), lds_s(a + (unsigned)(j * pitch4 + 4 * patch_src)));
          }
          __syncwarp();
          PYR_ACC(4);
        }
      }
      const uint2 ra = lds64u_s(tab_s + 8u * hh);
      pyr_rowpass<L>(p, ra.x + rd4, w[(2 * hh) % L]);
      pyr_rowpass<L>(p, ra.y + rd4, w[(2 * hh + 1) % L]);
      const int k = k_first + hh;
      if (k >= 0 && k < v.Ho) {
        PYR_T0();
        float2 lo[C::NC], hi[C::NC];
#pragma unroll
        for (int o = 0; o < C::NC; ++o) {
          float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < L; ++j) {
            a0 = ffma2_s(p.fh_lo[j], w[(2 * hh + 2 + j) % L][o], a0);
            a1 = ffma2_s(p.fh_hi[j], w[(2 * hh + 2 + j) % L][o], a1);
          }
          lo[o] = a0; hi[o] = a1;
        }
        em.row(lo, hi);
        PYR_ACC(5);
      }
      if (split == 2 && hh == C::HS / 2 - 1) {   // a staging group ends inside the stage: publish it
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (v.bar_out + (t * split) % kPyrNGO));
      }
    }
    // ---- hand the inputs back, publish the outputs -----------------------------------------------------------
#ifdef B200W_PYR_PROF
    const long long pyr_tf_ = clock64();
#endif
    fence_proxy_async();   // this lane's staging stores become visible to the bulk-store engine
    __syncwarp();
#ifdef B200W_PYR_PROF
    prof[3] += (unsigned long long)(clock64() - pyr_tf_);
#endif
    if (lane == 0) {
      if (lvl == 0) mbar_arrive(bar0 + 8 * (v.bar_in + v.n_in + (2 * t + 1) % nslot));
      mbar_arrive(bar0 + 8 * (v.bar_out + (t * split + split - 1) % kPyrNGO));
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
  PYR_FLUSH(2 + lvl, lane);
}

// ================================================================================================
// the kernel: one CTA per plane; warp 0 producer, warp 1 writer, the rest level workers
// ================================================================================================
// SPLIT: staging groups per worker stage (2 = half-stage groups: the single-level CTA then fits 4 per SM; the multi-level
// CTAs keep whole-stage groups -- measured 0.70 vs 0.89 ms per 268 Mpix at 1024^2 with half-stage groups)
template <int L, int MAXT, int MINB, int SPLIT>
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
    pyr_writer<L, SPLIT>(p, plane, smem, bar0, lane);
  } else {
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < kPyrMaxLevels; ++l)
      if (l < p.J && warp >= p.lv[l].warp0) lvl = l;
    pyr_worker<L, SPLIT>(p, plane, lvl, warp - p.lv[lvl].warp0, warp, lane, smem, bar0);
  }
}

template <int L, int MAXT, int MINB, int SPLIT>
inline int launch_pyramid_v(const PyrParams& p, cudaStream_t stream, int slot) {
  static int smem_set[64] = {};
  int dev = 0;
  (void)cudaGetDevice(&dev);
  (void)slot;
  if (dev < 0 || dev >= 64 || !smem_set[dev]) {
    if (cudaFuncSetAttribute(dwt_pyramid<L, MAXT, MINB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      (void)cudaGetLastError();
      return kNoFastPath;
    }
    if (dev >= 0 && dev < 64) smem_set[dev] = 1;
  }
  dwt_pyramid<L, MAXT, MINB, SPLIT><<<(unsigned)p.planes, p.threads, p.smem_bytes, stream>>>(p);
  return 0;
}

// three instantiations by CTA size (pyr_ctas_for_threads in pyramid_plan.h mirrors the MINB values)
template <int L>
inline int launch_pyramid(const PyrParams& p, cudaStream_t stream) {
  if (p.planes <= 0) return 0;
  if (p.threads <= 160) {
    if (p.split != pyr_split(L)) return kNoFastPath;
    return launch_pyramid_v<L, 160, pyr_minb_small(L), pyr_split(L)>(p, stream, 0);
  }
  if (p.split != 1) return kNoFastPath;
  if (p.threads <= 256) return launch_pyramid_v<L, 256, 2, 1>(p, stream, 1);
  return launch_pyramid_v<L, 512, 1, 1>(p, stream, 2);
}

}  // namespace fast
}  // namespace b200w
