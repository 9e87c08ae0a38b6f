// pyramid_plan.h -- host-side planning (and the index rules shared with the device code) of the fused DWT
// pyramid kernel (dwt_pyramid.cuh): ONE launch computes all J analysis levels of DWTForward
// (reference dwt/transform2d.py:68-74 = J x AFB2D, dwt/lowlevel.py:336-347) with no inter-level low-pass in HBM.
//
// One CTA per plane, warp-specialised dataflow (every arrow is a shared-memory ring guarded by mbarriers):
//
//   producer warp --TMA bulk row copies--> input ring --> level-1 warps --ll rows--> ring --> level-2 warps --> ...
//                                                             |                                  |
//                                                             +--- band-pass rows --> staging ---+--> writer warp
//                                                                                      (linear, at the global phase)  --TMA bulk stores--> HBM
//
// * level workers: a lane owns NC = 3 adjacent output columns; the pass along W reads the staged rows with 64-bit
//   LDS, the pass along H runs in a register window (same FMA order as afb2d_stream / the oracle: bit-identical);
// * rows are handed on in GROUPS = the output rows one worker stage produces (HS rows); stage s of a level
//   consumes extended rows [s*RS - PL, (s+1)*RS - PL) of its input, RS = 2*HS, PL = L-2;
// * the band-pass planes of a level are contiguous (Ho*Wo floats), so a group of rows is one contiguous run in
//   HBM: the workers deposit it in a staging ring laid out at the run's 128-byte phase and the writer flushes it
//   with cp.async.bulk (16-byte aligned middle) plus at most 3+3 scalar head/tail elements.
#pragma once
#include "common.h"

namespace b200w {

constexpr int kPyrMaxLevels = 4;
constexpr int kPyrMaxTaps = 16;
#ifndef B200W_PYR_NC
#define B200W_PYR_NC 3
#endif
constexpr int kPyrNC = B200W_PYR_NC;      // output columns per lane
constexpr int kPyrNGO = 2;     // staging groups per level (ring depth)
// the shared-memory shape the kernel is compiled for (and the only one the plan offers): input-ring slots of half a
// level-1 stage each, staging groups per worker stage
#ifndef B200W_PYR_SPLIT
#define B200W_PYR_SPLIT 2   /* half-stage staging groups: 53 KB per single-level CTA -> 4 CTAs per SM (r02_notes.md) */
#endif
#ifndef B200W_PYR_NSLOT
#define B200W_PYR_NSLOT 4
#endif
constexpr int kPyrAuxWarps = 2;  // warp 0 = producer, warp 1 = writer

struct PyrLevel {
  int H, W, Ho, Wo;             // input / output size of the level
  int n_stage;                  // worker stages = output groups
  int warp0, nwarps;            // compute warps [warp0, warp0 + nwarps) of the CTA
  int in_off, in_pitch, in_rows;  // input ring: float offset in dynamic smem, row pitch (floats), depth (rows)
  int n_in;                     // barriers per direction on the input ring: slots (level 0) / groups (levels >= 1)
  int bar_in;                   // index of in_full[0]; in_empty[0] = bar_in + n_in
  int st_off, st_cap, nbands;   // staging: float offset of band 0, floats per band (multiple of 4), 3 (+1: final ll)
  int st_cap_ll;                // last level: floats of the low-pass band's ring (rows of ll_pitch floats)
  int bar_out;                  // index of out_full[0]; out_empty[0] = bar_out + kPyrNGO
};

struct PyrParams {
  const float* x; long long xps; int xpitch;
  float* yl;
  float* highs[kPyrMaxLevels];
  int planes, J, mode, L;
  int nslot;                    // input-ring slots (half a level-1 stage each): 3 or 4
  int split;                    // staging groups per worker stage (a group = HS / split output rows): 1 or 2
  int ll_pitch;                 // row pitch (floats) of the final low-pass in global memory (>= its width)
  int zero_off;                 // a row of zeros (floats) for zero-padding rows of levels >= 1
  int tab_off;                  // per-warp tables of the current stage's row addresses (32 words per warp)
  int n_bars, smem_bytes, threads;
  PyrLevel lv[kPyrMaxLevels];
  // taps: the W-pass pairs {low-pass, high-pass} interleaved (one 64-bit uniform load feeds a packed FMA), the
  // H-pass taps as two scalar arrays
  alignas(16) float fw[2 * kPyrMaxTaps];
  alignas(16) float fh_lo[kPyrMaxTaps];
  alignas(16) float fh_hi[kPyrMaxTaps];
};

// ---- compile-time shape of a worker stage for filter length L ---------------------------------------------
constexpr int pyr_hs(int L) {     // half-stages (= output rows) per stage: a multiple of the window period L/2, even, >= 4
  int m = 1;
  while ((L / 2) * m < 4 || (((L / 2) * m) & 1)) ++m;
  return (L / 2) * m;
}
constexpr int pyr_halo(int L) { return (L - 2 + 3) / 4 * 4; }   // left pad of a ring row (floats)
// staging groups per worker stage: half-stage groups where a half stage is an even number of rows (its ring then wraps on
// a 16-byte boundary for every width), whole-stage groups otherwise
constexpr int pyr_split(int L) { return (pyr_hs(L) % 4 == 0) ? B200W_PYR_SPLIT : 1; }

// ---- group / stage index rules (host plan simulation and device code use the same functions) -----------------
// rows of group g of a level with Ho output rows: [g*HS - PRO, (g+1)*HS - PRO) clipped to [0, Ho)
B200W_HD int pyr_group_end(int g, int HS, int PRO) { return (g + 1) * HS - PRO; }   // exclusive, unclipped
B200W_HD int pyr_group_of_row(int r, int HS, int PRO) { return (r + PRO) / HS; }

// largest source row stage t of a consumer reads (input height H): rows e in [t*RS - PL, (t+1)*RS - PL) -> ext(e)
B200W_HD int pyr_stage_max_row(int t, int RS, int PL, int H, int mode) {
  const int e0 = t * RS - PL, e1 = e0 + RS - 1;
  int hi = imin(e1, H - 1);
  if (e0 < 0) {                       // mirrored rows above the image: the most negative one reaches farthest
    const int g = ext_index(e0, H, mode);
    if (g > hi) hi = g;
  }
  if (hi < 0) hi = 0;
  return hi;
}
// every row below this bound is dead once stage t is done (the bottom mirror reaches back to H - L + 1)
B200W_HD int pyr_stage_release_bound(int t, int RS, int PL, int H, int L) {
  return imin((t + 1) * RS - PL, H - L + 1);
}

// ---- the plan -----------------------------------------------------------------------------------------------
// Returns 0 and fills p (everything except pointers and taps), or 1 when the fused kernel does not apply.
// nslot / split / extra_in: shared-memory shape (see plan_pyramid_best); ll_pitch: row pitch of the final low-pass (0 = its width)
inline int plan_pyramid(PyrParams& p, int planes, int H, int W, int J, int L, int mode, long long xps, int xpitch,
                        const void* x, int max_smem_bytes, int nslot = 4, int split = 1, int extra_in = 1,
                        int ll_pitch = 0) {
  if (J < 1 || J > kPyrMaxLevels) return 1;
  if (L < 2 || (L & 1) || L > 12) return 1;   // (longer filters: the register window no longer fits 128 registers)
  if (mode != B200W_MODE_ZERO && mode != B200W_MODE_SYMMETRIC && mode != B200W_MODE_REFLECT) return 1;
  // TMA bulk row copies: 16-byte aligned source rows of a multiple of 16 bytes
  if ((W & 3) || (xpitch & 3) || (xps & 3) || (reinterpret_cast<uintptr_t>(x) & 15)) return 1;
  const int HS = pyr_hs(L), RS = 2 * HS, PL = L - 2, PRO = PL / 2, HALO = pyr_halo(L);
  p.planes = planes; p.J = J; p.mode = mode; p.L = L;
  p.nslot = nslot; p.split = split;
  if (nslot < 3 || nslot > 4 || (split != 1 && split != 2) || (HS % split)) return 1;
  int warp = kPyrAuxWarps;
  int h = H, w = W;
  int nbar = 0;
  int off = 0;                                   // floats, after the barrier block (added at the end)
  for (int l = 0; l < J; ++l) {
    PyrLevel& v = p.lv[l];
    if (h < L || w < L) return 1;                // one reflection must cover the halo
    if (l > 0 && w < 2 * L - 2) return 1;        // ... and a column feeds at most one halo cell of the next level
    v.H = h; v.W = w;
    v.Ho = (h + L - 1) / 2; v.Wo = (w + L - 1) / 2;
    v.n_stage = (v.Ho + PRO + HS - 1) / HS;
    const int lanes = (v.Wo + kPyrNC - 1) / kPyrNC;
    v.nwarps = (lanes + 31) / 32;
    v.warp0 = warp; warp += v.nwarps;
    // ring row: HALO left pad, the row, right halo; a lane with at least one valid column reads up to
    // HALO + 2*(c0 + NC - 1) + 1 with c0 <= Wo - 1  ->  pitch >= HALO + 2*(Wo + NC - 2) + 2
    v.in_pitch = (HALO + 2 * (v.Wo + kPyrNC - 2) + 2 + 3) / 4 * 4;
    if (v.in_pitch < HALO + w + L - 1) v.in_pitch = (HALO + w + L - 1 + 3) / 4 * 4;
    if (l == 0) {
      v.n_in = nslot;
      v.in_rows = nslot * HS;
    } else {
      // groups of the previous level that must be resident at once: simulate the consumer's stage sequence
      const PyrLevel& u = p.lv[l - 1];
      int need = 1, rel = 0;
      for (int t = 0; t < v.n_stage; ++t) {
        const int g_need = imin(pyr_group_of_row(pyr_stage_max_row(t, RS, PL, h, mode), HS, PRO), u.n_stage - 1);
        if (g_need - rel + 1 > need) need = g_need - rel + 1;
        const int lo = pyr_stage_release_bound(t, RS, PL, h, L);
        while (rel < u.n_stage && pyr_group_end(rel, HS, PRO) <= lo) ++rel;
      }
      v.n_in = need + extra_in;                  // (+1: the producing level can work a group ahead)
      v.in_rows = v.n_in * HS;
    }
    v.in_off = off; off += v.in_rows * v.in_pitch;
    v.bar_in = nbar; nbar += 2 * v.n_in;
    v.nbands = (l == J - 1) ? 4 : 3;
    v.st_cap = kPyrNGO * (HS / split) * v.Wo;    // whole rows
    v.st_cap_ll = 0;
    if (l == J - 1) {
      p.ll_pitch = (ll_pitch > 0) ? ll_pitch : v.Wo;
      if (p.ll_pitch < v.Wo) return 1;
      v.st_cap_ll = kPyrNGO * (HS / split) * p.ll_pitch;
    }
    if ((v.st_cap & 3) || (v.st_cap_ll & 3)) return 1;   // the rings must wrap on a 16-byte boundary
    v.st_off = off; off += 3 * v.st_cap + v.st_cap_ll;
    v.bar_out = nbar; nbar += 2 * kPyrNGO;
    h = v.Ho; w = v.Wo;
  }
  int maxpitch = 0;
  for (int l = 0; l < J; ++l) maxpitch = imax(maxpitch, p.lv[l].in_pitch);
  p.zero_off = off; off += (maxpitch + 2 * kPyrNC + 8 + 3) / 4 * 4;
  p.tab_off = off; off += 32 * warp;
  p.n_bars = nbar;
  const int bar_floats = (2 * nbar + 31) / 32 * 32;   // 8 bytes each, block rounded to 128 bytes
  for (int l = 0; l < J; ++l) { p.lv[l].in_off += bar_floats; p.lv[l].st_off += bar_floats; }
  p.zero_off += bar_floats;
  p.tab_off += bar_floats;
  p.smem_bytes = (off + bar_floats) * 4;
  p.threads = 32 * warp;
  if (p.threads > 512 || p.smem_bytes > max_smem_bytes) return 1;
  p.x = static_cast<const float*>(x); p.xps = xps; p.xpitch = xpitch;
  return 0;
}

// CTAs per SM the kernel instantiation for `threads` is compiled for (__launch_bounds__ in dwt_pyramid.cuh)
// (a 5-warp CTA -- one level of up to 96 lanes -- runs 4 per SM in 96 registers for filters up to 8 taps; longer
// filters spill under that cap and stay at 3)
#ifndef B200W_PYR_MINB_SMALL
#define B200W_PYR_MINB_SMALL 4
#endif
constexpr int pyr_minb_small(int L) { return (L <= 8) ? B200W_PYR_MINB_SMALL : (B200W_PYR_MINB_SMALL < 3 ? B200W_PYR_MINB_SMALL : 3); }
inline int pyr_ctas_for_threads(int threads, int L) { return threads <= 160 ? pyr_minb_small(L) : (threads <= 256 ? 2 : 1); }

// Picks the shared-memory shape that lets the most CTAs share an SM (up to what the register budget of the matching
// kernel instantiation allows), preferring the deeper rings among equals.
inline int plan_pyramid_best(PyrParams& p, int planes, int H, int W, int J, int L, int mode, long long xps, int xpitch,
                             const void* x, int max_smem_bytes, int sm_smem_bytes, int ll_pitch = 0) {
  // 4 input slots; half-stage staging groups for the single-level CTA (<= 160 threads), whole-stage groups otherwise.
  // Measured (profiles/r02_notes.md; level 1 of 4096 x 512^2, db4): whole-stage groups 1.67 ms at 3 CTAs/SM, 3 slots +
  // half-stage groups 1.59 ms, 4 slots + half-stage groups 1.55 ms at 4 CTAs/SM; the 512-thread 3-level CTA is slower with
  // half-stage groups (0.89 vs 0.70 ms per 268 Mpix at 1024^2); an 80-register 8-warp form for a third 3-level CTA spills.
  PyrParams best;
  int best_ctas = 0;
  for (int split = 1; split <= 2; ++split) {
    PyrParams q;
    if (plan_pyramid(q, planes, H, W, J, L, mode, xps, xpitch, x, max_smem_bytes, B200W_PYR_NSLOT, split, 1, ll_pitch))
      continue;
    if (split != ((q.threads <= 160) ? pyr_split(L) : 1)) continue;   // the shape each kernel instantiation is compiled for
    const int by_smem = sm_smem_bytes / (q.smem_bytes + 1024);
    const int ctas = imin(by_smem, pyr_ctas_for_threads(q.threads, L));
    if (ctas > best_ctas) { best_ctas = ctas; best = q; }
  }
  if (best_ctas == 0) return 1;
  p = best;
  return 0;
}

}  // namespace b200w
