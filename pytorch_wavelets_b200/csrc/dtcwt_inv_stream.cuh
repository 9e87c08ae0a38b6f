#pragma once
#include "stream_common.cuh"

namespace b200w {
namespace fast {

// ================================================================================================
// K5 fast: DTCWT level-1 inverse (reference INV_J1.forward / inv_j1, transform_funcs.py:152-184).
//   y = R1(C1(hh) + C0(hl)) + R0(C1(lh) + C0(ll)),  C/R = col/row filter with g0 (L0 taps) / g1 (L1 taps).
// The passes commute, so the kernel runs the W pass on staged rows first and the H pass in a register
// window:  A = R1(hh) + R0(lh),  B = R1(hl) + R0(ll),  y = C1(A) + C0(B).
//   lane  = one complex column q (output columns 2q, 2q+1); strip = 32 complex columns;
//   stage = one complex row i (quad rows 2i, 2i+1): the six orientation rows + two ll rows are staged with
//           16-byte cp.async, c2q'd once per complex sample into three real band rows in shared memory
//           (symmetric extension = copy of the mirrored band sample, row extension = mirrored complex row with
//           the row parity swapped), then filtered.
// Requires the default band-pass layout along the row (re/im adjacent, columns contiguous) and 16-byte
// aligned rows; everything else takes the generic kernel.
// ================================================================================================
template <int HLA, int NS, int MS>
struct QuadStager;

#ifndef B200W_INVJ1_NS
#define B200W_INVJ1_NS 3
#endif
#ifndef B200W_INVJ1_MINB
#define B200W_INVJ1_MINB 1
#endif
// (an explicit minBlocks of 1 is not neutral: ptxas then spends registers freely -- fwd_j2plus 156 -> 176, fwd_j1 96 -> 124 --
// so the plain form is used unless a cap is asked for)
#if B200W_INVJ1_MINB > 1
#define B200W_INVJ1_LB __launch_bounds__(32, B200W_INVJ1_MINB)
#else
#define B200W_INVJ1_LB __launch_bounds__(32)
#endif

template <int L0, int L1>
struct I1Cfg {
  static constexpr int M0 = L0 / 2, M1 = L1 / 2, M = (M0 > M1) ? M0 : M1;
  static constexpr int HLA = (M + 3) / 4 * 4;
  static constexpr int SW = HLA + 64 + HLA;
  static constexpr int CPR = SW / 4;
  static constexpr int OFFX = HLA - M;
  static constexpr int NX = OFFX + 2 * M + 2;
  static constexpr int NV2 = (NX + 1) / 2;
  static constexpr int MS = (M + 1) / 2;             // complex rows of context above / below
  static constexpr int WR = 4 * MS + 2;              // quad rows held in the register window
  static constexpr int UNR = WR / 2;                 // window period in stages
  static constexpr int PRO = 2 * MS;
  static constexpr int NS = B200W_INVJ1_NS;
  static constexpr int SMEM_BYTES = QuadStager<HLA, NS, MS>::SMEM_FLOATS * 4;
};

template <int L0, int L1, int U>
__device__ __forceinline__ void i1_stage(const DtParams& p, const float* band, const float* llrow, bool has_hi,
                                         bool has_ll, float2 (&wA)[I1Cfg<L0, L1>::WR],
                                         float2 (&wB)[I1Cfg<L0, L1>::WR], bool emit, float*& y_ptr, bool colvalid) {
  using C = I1Cfg<L0, L1>;
  constexpr int WR = C::WR;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float xlh[2 * C::NV2], xhl[2 * C::NV2], xhh[2 * C::NV2], xll[2 * C::NV2];
#pragma unroll
    for (int q = 0; q < C::NV2; ++q) {
      const float2 a = *reinterpret_cast<const float2*>(band + (0 * 2 + rr) * C::SW + 2 * q);
      const float2 b = *reinterpret_cast<const float2*>(band + (1 * 2 + rr) * C::SW + 2 * q);
      const float2 c = *reinterpret_cast<const float2*>(band + (2 * 2 + rr) * C::SW + 2 * q);
      const float2 d = *reinterpret_cast<const float2*>(llrow + rr * C::SW + 2 * q);
      xlh[2 * q] = a.x; xlh[2 * q + 1] = a.y;
      xhl[2 * q] = b.x; xhl[2 * q + 1] = b.y;
      xhh[2 * q] = c.x; xhh[2 * q + 1] = c.y;
      xll[2 * q] = d.x; xll[2 * q + 1] = d.y;
    }
    const int S = (2 * U + rr) % WR;
    float a2[2], b2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float r1hh = 0.f, r1hl = 0.f, r0lh = 0.f, r0ll = 0.f;
#pragma unroll
      for (int j = 0; j < L1; ++j) {
        const int ix = C::OFFX + e + (C::M - C::M1) + j;
        r1hh = fmaf(p.f1.t[j], xhh[ix], r1hh);
        r1hl = fmaf(p.f1.t[j], xhl[ix], r1hl);
      }
#pragma unroll
      for (int j = 0; j < L0; ++j) {
        const int ix = C::OFFX + e + (C::M - C::M0) + j;
        r0lh = fmaf(p.f0.t[j], xlh[ix], r0lh);
        r0ll = fmaf(p.f0.t[j], xll[ix], r0ll);
      }
      // A = R1(hh) + R0(lh);  B = R1(hl) + R0(ll)   (absent inputs contribute exact zeros)
      a2[e] = has_hi ? __fadd_rn(r1hh, r0lh) : 0.f;
      b2[e] = has_hi ? (has_ll ? __fadd_rn(r1hl, r0ll) : r1hl) : r0ll;
    }
    wA[S] = make_float2(a2[0], a2[1]);
    wB[S] = make_float2(b2[0], b2[1]);
  }
  if (emit) {
    // column pass: one packed FMA per tap covers the lane's two output columns
#pragma unroll
    for (int dr = 0; dr < 2; ++dr) {
      float2 a = make_float2(0.f, 0.f), b = a;
#pragma unroll
      for (int j = 0; j < L1; ++j) a = ffma2_s(p.f1.t[j], wA[(2 * U - 2 * C::MS + dr - C::M1 + j + 4 * WR) % WR], a);
#pragma unroll
      for (int j = 0; j < L0; ++j) b = ffma2_s(p.f0.t[j], wB[(2 * U - 2 * C::MS + dr - C::M0 + j + 4 * WR) % WR], b);
      const float o0 = has_hi ? __fadd_rn(a.x, b.x) : b.x, o1 = has_hi ? __fadd_rn(a.y, b.y) : b.y;
      if (colvalid) store2(y_ptr + dr * p.outpitch, o0, o1, 2, false);
    }
    y_ptr += 2 * p.outpitch;
  }
}

template <int L0, int L1, int U>
__device__ __forceinline__ void i1_dispatch(int uu, const DtParams& p, const float* band, const float* llrow,
                                            bool has_hi, bool has_ll, float2 (&wA)[I1Cfg<L0, L1>::WR],
                                            float2 (&wB)[I1Cfg<L0, L1>::WR], bool emit, float*& y_ptr,
                                            bool colvalid) {
  if constexpr (U < I1Cfg<L0, L1>::UNR) {
    if (uu == U) i1_stage<L0, L1, U>(p, band, llrow, has_hi, has_ll, wA, wB, emit, y_ptr, colvalid);
    else i1_dispatch<L0, L1, U + 1>(uu, p, band, llrow, has_hi, has_ll, wA, wB, emit, y_ptr, colvalid);
  }
}

// ------------------------------------------------------------------------------------------------
// QuadStager: shared front end of the two DTCWT inverse kernels.  Per stage = one complex row ic:
//   ring rows 0..5 = the six orientation rows (complex, as stored), rows 6,7 = ll rows 2ic, 2ic+1;
//   after landing, c2q turns the complex rows into three real band rows x two quad rows in `bandbuf`
//   (row index b*2 + rr, b: 0 lh, 1 hl, 2 hh), and the out-of-image columns of all eight real rows
//   are patched from their mirror (or zero).  Out-of-image complex rows are the mirrored row with the
//   row parity swapped.  Staged window: columns [c0 - HLA, c0 + 64 + HLA).
// ------------------------------------------------------------------------------------------------
template <int HLA, int NS, int MS>
struct QuadStager {
  static constexpr int SW = HLA + 64 + HLA;
  static constexpr int CPR = SW / 4;
  static constexpr int VR = 8;
  static constexpr int STAGE = VR * SW;
  static constexpr int BAND = 6 * SW;
  static constexpr int NCH = (VR * CPR + 31) / 32;
  static constexpr int NHALO = HLA / 2;
  static constexpr int NFIX = (8 * 2 * HLA + 31) / 32;
  static constexpr int SMEM_FLOATS = NS * STAGE + BAND;

  float* ring;
  float* bandbuf;
  const float* hbase;
  const float* llp;
  long long hs2, hs3;
  int inpitch, H, W, h2, sym, i0, n_stage, lane;
  bool has_hi, has_ll, any_fix;
  int c_soff[NCH], c_gcol[NCH];
  int fix_dst[NFIX], fix_src[NFIX];

  __device__ __forceinline__ void init(float* smem, const DtParams& p, int plane, int c0, int need_cols, int i0_,
                                       int n_stage_, int lane_) {
    ring = smem;
    bandbuf = smem + NS * STAGE;
    lane = lane_;
    i0 = i0_;
    n_stage = n_stage_;
    H = p.H; W = p.W; h2 = p.H >> 1;
    has_hi = (p.highs != nullptr);
    has_ll = (p.in != nullptr);
    const int n = plane / p.C, ch = plane - n * p.C;
    hbase = has_hi ? p.highs + n * p.hs[0] + ch * p.hs[1] : nullptr;
    llp = has_ll ? p.in + (long long)plane * p.inps : nullptr;
    hs2 = p.hs[2]; hs3 = p.hs[3];
    inpitch = p.inpitch;
    sym = has_hi ? p.sym : 1;   // low-pass-only path ignores `mode` (reference transform_funcs.py:159)
    const int c_a = c0 - HLA;
    // zero ring + band buffer once (absent inputs / never-copied columns must read as zeros)
    for (int i = lane; i < SMEM_FLOATS; i += 32) smem[i] = 0.f;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int chn = lane + 32 * k;
      const int v = chn / CPR;
      const int cc = chn - v * CPR;
      const int gc = c_a + 4 * cc;
      const bool on = (chn < VR * CPR) && (4 * cc < need_cols) && (gc >= 0) && (gc + 3 < W) &&
                      ((v < 6) ? has_hi : has_ll);
      c_soff[k] = on ? v * SW + 4 * cc : -1;
      c_gcol[k] = gc;
    }
    const int nleft = imin(imax(0, -c_a), need_cols);
    const int sr0 = imax(W - c_a, 0);
    const int nright = imax(0, need_cols - sr0);
    const int nb_row = nleft + nright;
    bool bad = false;
#pragma unroll
    for (int q = 0; q < NFIX; ++q) {
      fix_dst[q] = -1;
      fix_src[q] = -1;
      const int e = lane + 32 * q;
      if (e < 8 * nb_row) {
        const int v = e / (nb_row > 0 ? nb_row : 1);   // 0..5 band rows, 6..7 ll rows
        const int idx = e - v * nb_row;
        const int sidx = (idx < nleft) ? idx : sr0 + (idx - nleft);
        const int g = sym_or_zero(c_a + sidx, W, sym);
        fix_dst[q] = v * SW + sidx;
        if (g >= 0) {
          const int ss = g - c_a;
          if (ss < 0 || ss >= need_cols) bad = true;
          fix_src[q] = v * SW + ss;
        }
      }
    }
    any_fix = (nb_row > 0) && !__any_sync(0xffffffffu, bad);
  }

  __device__ __forceinline__ void issue(int t) {
    if (t < n_stage) {
      float* dst = ring + (t % NS) * STAGE;
      const int ic = i0 - MS + t;                    // complex row of this stage (may be outside the image)
      int ir = ic;
      if (ic < 0) ir = -1 - ic; else if (ic >= h2) ir = 2 * h2 - 1 - ic;
      const bool row_ok = (ic >= 0 && ic < h2) || (sym && ir >= 0 && ir < h2);
      const int r0 = sym_or_zero(2 * ic, H, sym), r1 = sym_or_zero(2 * ic + 1, H, sym);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        if (c_soff[k] < 0) continue;
        const int v = c_soff[k] / SW;
        float* d = dst + c_soff[k];
        if (v < 6) {
          if (row_ok) cp_async16(d, hbase + (long long)v * hs2 + (long long)ir * hs3 + c_gcol[k]);
          else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          const int rq = (v == 6) ? r0 : r1;
          if (rq >= 0) cp_async16(d, llp + (long long)rq * inpitch + c_gcol[k]);
          else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  }

  __device__ __forceinline__ void prologue() {
#pragma unroll 1
    for (int t = 0; t < NS - 1; ++t) issue(t);
  }

  // wait for stage t, c2q it into bandbuf, patch borders; returns the ring stage (ll rows at rows 6,7)
  __device__ __forceinline__ float* acquire(int t) {
    cp_async_wait<NS - 2>();
    __syncwarp();
    float* stage = ring + (t % NS) * STAGE;
    const int ic = i0 - MS + t;
    const bool swap = (ic < 0 || ic >= h2);          // extended rows: row parity swapped
    if (has_hi) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        int qc;                                      // staged complex column (float offset 2*qc)
        if (part == 0) qc = NHALO + lane;
        else if (lane < NHALO) qc = lane;
        else if (lane < 2 * NHALO) qc = NHALO + 32 + (lane - NHALO);
        else break;
        const float* sp = stage + 2 * qc;
        const float2 w0 = *reinterpret_cast<const float2*>(sp + 0 * SW);
        const float2 w1 = *reinterpret_cast<const float2*>(sp + 1 * SW);
        const float2 w2 = *reinterpret_cast<const float2*>(sp + 2 * SW);
        const float2 w3 = *reinterpret_cast<const float2*>(sp + 3 * SW);
        const float2 w4 = *reinterpret_cast<const float2*>(sp + 4 * SW);
        const float2 w5 = *reinterpret_cast<const float2*>(sp + 5 * SW);
        // band <- (w1, w2): lh <- (o0, o5), hl <- (o2, o3), hh <- (o1, o4)   (reference transform_funcs.py:91-93)
        const float2 p1[3] = {w0, w2, w1}, p2[3] = {w5, w3, w4};
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const float a_ = __fmul_rn(__fadd_rn(p1[b].x, p2[b].x), kInvSqrt2);   // (row 0, col 0)
          const float b_ = __fmul_rn(__fadd_rn(p1[b].y, p2[b].y), kInvSqrt2);   // (row 0, col 1)
          const float c_ = __fmul_rn(__fsub_rn(p1[b].y, p2[b].y), kInvSqrt2);   // (row 1, col 0)
          const float d_ = __fmul_rn(__fsub_rn(p2[b].x, p1[b].x), kInvSqrt2);   // (row 1, col 1)
          float* o0 = bandbuf + (b * 2 + (swap ? 1 : 0)) * SW + 2 * qc;
          float* o1 = bandbuf + (b * 2 + (swap ? 0 : 1)) * SW + 2 * qc;
          *reinterpret_cast<float2*>(o0) = make_float2(a_, b_);
          *reinterpret_cast<float2*>(o1) = make_float2(c_, d_);
        }
      }
      __syncwarp();
    }
    if (any_fix) {
#pragma unroll
      for (int q = 0; q < NFIX; ++q) {
        if (fix_dst[q] < 0) continue;
        const bool is_ll = fix_dst[q] >= 6 * SW;
        float* basep = is_ll ? stage : bandbuf;      // ll rows live in the ring at virtual rows 6,7
        if (is_ll ? has_ll : has_hi) basep[fix_dst[q]] = (fix_src[q] >= 0) ? basep[fix_src[q]] : 0.f;
      }
      __syncwarp();
    }
    return stage;
  }
};

inline bool quad_inputs_ok(const DtParams& p) {
  if (p.highs) {
    // band-pass rows must be plain complex rows: re/im adjacent, columns contiguous, 16-byte aligned
    if (p.hs[5] != 1 || p.hs[4] != 2) return false;
    if ((p.hs[0] | p.hs[1] | p.hs[2] | p.hs[3]) & 3) return false;
    if (reinterpret_cast<uintptr_t>(p.highs) & 15) return false;
  }
  if (p.in && !aligned_plane(p.in, p.inps, p.inpitch)) return false;
  return (p.W & 3) == 0;
}

template <int L0, int L1>
__global__ void B200W_INVJ1_LB inv_j1_stream(const __grid_constant__ DtParams p, int n_strips, int n_chunks,
                                                    int CH /* complex rows per chunk */) {
  using C = I1Cfg<L0, L1>;
  using QS = QuadStager<C::HLA, C::NS, C::MS>;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x;
  long long item = blockIdx.x;
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);

  const int c0 = strip * 64;                         // first output (= quad-domain) column of the strip
  const int i0 = chunk * CH;
  const int i1 = imin(i0 + CH, p.H >> 1);
  const int n_stage = (i1 - i0) + C::PRO;
  const int ncols = imin(64, p.W - c0);

  QS qs;
  qs.init(smem, p, plane, c0, C::HLA + ncols + C::M, i0, n_stage, lane);
  qs.prologue();

  float2 wA[C::WR], wB[C::WR];
#pragma unroll
  for (int j = 0; j < C::WR; ++j) { wA[j] = wB[j] = make_float2(0.f, 0.f); }

  const bool colvalid = (c0 + 2 * lane) < p.W;
  float* y_ptr = p.out + (long long)plane * p.outps + (long long)(2 * i0) * p.outpitch + c0 + 2 * lane;

  int uu = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    const float* stage = qs.acquire(t);
    qs.issue(t + C::NS - 1);
    i1_dispatch<L0, L1, 0>(uu, p, qs.bandbuf + 2 * lane, stage + 6 * C::SW + 2 * lane, qs.has_hi, qs.has_ll, wA, wB,
                           t >= C::PRO, y_ptr, colvalid);
    uu = (uu + 1 == C::UNR) ? 0 : uu + 1;
    __syncwarp();   // band buffer is rewritten by the next stage's c2q
  }
  cp_async_wait<0>();
}

template <int L0, int L1>
inline int launch_i1_stream(const DtParams& p, cudaStream_t stream) {
  using C = I1Cfg<L0, L1>;
  if (!quad_inputs_ok(p) || (p.outpitch & 1) || p.W < 2 * C::HLA) return kNoFastPath;
  const int n_strips = (p.W + 63) / 64;
  const long long planes = (long long)p.N * p.C;
  int n_chunks, CH;
  static ConcCache conc_cache;
  const int conc = resident_warps_dev(conc_cache, inv_j1_stream<L0, L1>, C::SMEM_BYTES);
  pick_chunks(planes * n_strips, p.H >> 1, 8, 8, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  inv_j1_stream<L0, L1><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

// (a 128-column form -- two complex columns per lane, band rows interleaved by filter so that the pass along W is packed
// FFMA2 as well: -40 % shared-memory wavefronts, -30 % instructions -- was built and measured in round 2: 164-205
// registers leave 9-12 warps per SM instead of 20, and the level is 10-13 % slower; profiles/r02_notes.md)
int try_launch_inv_j1(const DtParams& p, cudaStream_t stream) {
  if ((long long)p.N * p.C == 0) return 0;
  if (p.L0 == 7 && p.L1 == 5) return launch_i1_stream<7, 5>(p, stream);   // near_sym_a synthesis
  if (p.L0 == 5 && p.L1 == 7) return launch_i1_stream<5, 7>(p, stream);   // near_sym_a analysis (backward of fwd)
  if (p.L0 == 7 && p.L1 == 9) return launch_i1_stream<7, 9>(p, stream);   // antonini synthesis
  if (p.L0 == 3 && p.L1 == 5) return launch_i1_stream<3, 5>(p, stream);   // legall synthesis
  if (p.L0 == 19 && p.L1 == 13) return launch_i1_stream<19, 13>(p, stream);  // near_sym_b synthesis
  if (p.L0 == 9 && p.L1 == 7) return launch_i1_stream<9, 7>(p, stream);   // antonini analysis (backward of fwd)
  if (p.L0 == 5 && p.L1 == 3) return launch_i1_stream<5, 3>(p, stream);   // legall analysis
  if (p.L0 == 13 && p.L1 == 19) return launch_i1_stream<13, 19>(p, stream);  // near_sym_b analysis
  return kNoFastPath;
}

// ================================================================================================
// K6 fast: DTCWT level >= 2 inverse (reference INV_J2PLUS.forward / inv_j2plus, transform_funcs.py:279-307).
//   y = R_H(C_H(hh) + C_L(hl)) + R_L(C_H(lh) + C_L(ll)),  C/R = col/row interpolating q-shift filters
//   (colifilt / rowifilt, dtcwt/lowlevel.py:154-239): out[4t+s] = sum_{j<m2} f_s[j] x[sym(2(t+j) + o_s - m2)],
//   low call (ha,hb) = (g0b,g0a), high call (ha,hb) = (g1b,g1a) with the high-pass phase table.
// W pass first on the staged rows (A = R_H(hh) + R_L(lh), B = R_H(hl) + R_L(ll), 4 output columns per lane),
// H pass in the register window (y = C_H(A) + C_L(B), 4 output rows per stage).  Same front end as K5.
//   taps: f0=g0a f1=g1a f2=g0b f3=g1b (stored).
// ================================================================================================
template <int M2, bool HP>
struct IfPhase {  // dtcwt/lowlevel.py:169-186
  static constexpr int par(int s) { return (M2 % 2 == 0) ? (s >= 2 ? 1 : 0) : (s < 2 ? 1 : 0); }
  static constexpr int off(int s) {
    return (M2 % 2 == 0) ? (HP ? (s ^ 1) : s) : (HP ? (2 - (s & 1)) : (1 + (s & 1)));
  }
  static constexpr int omin = (M2 % 2 == 0) ? 0 : 1;
  static constexpr int omax = (M2 % 2 == 0) ? 3 : 2;
};

#ifndef B200W_INVJ2_NS
#define B200W_INVJ2_NS 3
#endif
template <int MQ>
struct I2Cfg {
  static constexpr int M2 = MQ / 2;
  static constexpr int OMIN = IfPhase<M2, false>::omin, OMAX = IfPhase<M2, false>::omax;
  static constexpr int HL = M2 - OMIN;               // input columns needed left of 2u
  static constexpr int HR = M2 + OMAX - 3;           // ... right of 2u+1
  static constexpr int HMAX = (HL > HR) ? HL : HR;
  static constexpr int HLA = (HMAX + 3) / 4 * 4;
  static constexpr int SW = HLA + 64 + HLA;
  static constexpr int OFFX = HLA - HL;
  static constexpr int NX = OFFX + HL + 2 + HR;
  static constexpr int NV2 = (NX + 1) / 2;
  static constexpr int MS = (HMAX + 1) / 2;
  static constexpr int WR = 4 * MS + 2;
  static constexpr int UNR = WR / 2;
  static constexpr int PRO = 2 * MS;
  static constexpr int NS = B200W_INVJ2_NS;
  static constexpr int SMEM_BYTES = QuadStager<HLA, NS, MS>::SMEM_FLOATS * 4;
};

template <int MQ, int U>
__device__ __forceinline__ void i2_stage(const DtParams& p, const float* band, const float* llrow, bool has_hi,
                                         bool has_ll, float2 (&wA)[I2Cfg<MQ>::WR][2], float2 (&wB)[I2Cfg<MQ>::WR][2],
                                         bool emit, float*& y_ptr, bool colvalid, bool vec4) {
  using C = I2Cfg<MQ>;
  using PL = IfPhase<C::M2, false>;
  using PH = IfPhase<C::M2, true>;
  constexpr int WR = C::WR, M2 = C::M2;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float xlh[2 * C::NV2], xhl[2 * C::NV2], xhh[2 * C::NV2], xll[2 * C::NV2];
#pragma unroll
    for (int q = 0; q < C::NV2; ++q) {
      const float2 a = *reinterpret_cast<const float2*>(band + (0 * 2 + rr) * C::SW + 2 * q);
      const float2 b = *reinterpret_cast<const float2*>(band + (1 * 2 + rr) * C::SW + 2 * q);
      const float2 c = *reinterpret_cast<const float2*>(band + (2 * 2 + rr) * C::SW + 2 * q);
      const float2 d = *reinterpret_cast<const float2*>(llrow + rr * C::SW + 2 * q);
      xlh[2 * q] = a.x; xlh[2 * q + 1] = a.y;
      xhl[2 * q] = b.x; xhl[2 * q + 1] = b.y;
      xhh[2 * q] = c.x; xhh[2 * q + 1] = c.y;
      xll[2 * q] = d.x; xll[2 * q + 1] = d.y;
    }
    const int S = (2 * U + rr) % WR;
    float a4[4], b4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      // s even -> ha, s odd -> hb;  low: ha=g0b(f2) hb=g0a(f0);  high: ha=g1b(f3) hb=g1a(f1)
      const float* fl = (s & 1) ? p.f0.t : p.f2.t;
      const float* fh = (s & 1) ? p.f1.t : p.f3.t;
      float rh_hh = 0.f, rh_hl = 0.f, rl_lh = 0.f, rl_ll = 0.f;
#pragma unroll
      for (int j = 0; j < M2; ++j) {
        const int ih = C::OFFX + 2 * j + (PH::off(s) - C::OMIN);
        const int il = C::OFFX + 2 * j + (PL::off(s) - C::OMIN);
        const float ch = fh[2 * j + PH::par(s)], cl = fl[2 * j + PL::par(s)];
        rh_hh = fmaf(ch, xhh[ih], rh_hh);
        rh_hl = fmaf(ch, xhl[ih], rh_hl);
        rl_lh = fmaf(cl, xlh[il], rl_lh);
        rl_ll = fmaf(cl, xll[il], rl_ll);
      }
      a4[s] = has_hi ? __fadd_rn(rh_hh, rl_lh) : 0.f;
      b4[s] = has_hi ? (has_ll ? __fadd_rn(rh_hl, rl_ll) : rh_hl) : rl_ll;
    }
    wA[S][0] = make_float2(a4[0], a4[1]); wA[S][1] = make_float2(a4[2], a4[3]);
    wB[S][0] = make_float2(b4[0], b4[1]); wB[S][1] = make_float2(b4[2], b4[3]);
  }
  if (emit) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {   // output row 4i + s; one packed FMA per tap covers two of the lane's four columns
      const float* fl = (s & 1) ? p.f0.t : p.f2.t;
      const float* fh = (s & 1) ? p.f1.t : p.f3.t;
      float2 o[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float2 a = make_float2(0.f, 0.f), b = a;
#pragma unroll
        for (int j = 0; j < M2; ++j) {
          const int sh = (2 * U - 2 * C::MS + 2 * j + PH::off(s) - M2 + 4 * WR) % WR;
          const int sl = (2 * U - 2 * C::MS + 2 * j + PL::off(s) - M2 + 4 * WR) % WR;
          a = ffma2_s(fh[2 * j + PH::par(s)], wA[sh][c], a);
          b = ffma2_s(fl[2 * j + PL::par(s)], wB[sl][c], b);
        }
        o[c] = has_hi ? make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)) : b;
      }
      if (colvalid) {
        float* q = y_ptr + s * p.outpitch;
        if (vec4) *reinterpret_cast<float4*>(q) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        else { q[0] = o[0].x; q[1] = o[0].y; q[2] = o[1].x; q[3] = o[1].y; }
      }
    }
    y_ptr += 4 * p.outpitch;
  }
}

template <int MQ, int U>
__device__ __forceinline__ void i2_dispatch(int uu, const DtParams& p, const float* band, const float* llrow,
                                            bool has_hi, bool has_ll, float2 (&wA)[I2Cfg<MQ>::WR][2],
                                            float2 (&wB)[I2Cfg<MQ>::WR][2], bool emit, float*& y_ptr, bool colvalid,
                                            bool vec4) {
  if constexpr (U < I2Cfg<MQ>::UNR) {
    if (uu == U) i2_stage<MQ, U>(p, band, llrow, has_hi, has_ll, wA, wB, emit, y_ptr, colvalid, vec4);
    else i2_dispatch<MQ, U + 1>(uu, p, band, llrow, has_hi, has_ll, wA, wB, emit, y_ptr, colvalid, vec4);
  }
}

#ifndef B200W_INVJ2_MINB
#define B200W_INVJ2_MINB 1
#endif
// (an explicit minBlocks of 1 is not neutral: ptxas then spends registers freely -- fwd_j2plus 156 -> 176, fwd_j1 96 -> 124 --
// so the plain form is used unless a cap is asked for)
#if B200W_INVJ2_MINB > 1
#define B200W_INVJ2_LB __launch_bounds__(32, B200W_INVJ2_MINB)
#else
#define B200W_INVJ2_LB __launch_bounds__(32)
#endif
template <int MQ>
__global__ void B200W_INVJ2_LB inv_j2plus_stream(const __grid_constant__ DtParams p, int n_strips,
                                                        int n_chunks, int CH /* complex rows per chunk */) {
  using C = I2Cfg<MQ>;
  using QS = QuadStager<C::HLA, C::NS, C::MS>;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x;
  long long item = blockIdx.x;
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);

  const int c0 = strip * 64;                         // first input (quad-domain) column of the strip
  const int i0 = chunk * CH;
  const int i1 = imin(i0 + CH, p.H >> 1);
  const int n_stage = (i1 - i0) + C::PRO;
  const int ncols = imin(64, p.W - c0);

  QS qs;
  qs.init(smem, p, plane, c0, C::HLA + ncols + C::HR, i0, n_stage, lane);
  qs.prologue();

  float2 wA[C::WR][2], wB[C::WR][2];
#pragma unroll
  for (int j = 0; j < C::WR; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) { wA[j][c] = make_float2(0.f, 0.f); wB[j][c] = make_float2(0.f, 0.f); }

  const bool colvalid = (c0 + 2 * lane) < p.W;
  float* y_ptr = p.out + (long long)plane * p.outps + (long long)(4 * i0) * p.outpitch + 2 * c0 + 4 * lane;
  const bool vec4 = ((p.outpitch & 3) == 0) && ((p.outps & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

  int uu = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    const float* stage = qs.acquire(t);
    qs.issue(t + C::NS - 1);
    i2_dispatch<MQ, 0>(uu, p, qs.bandbuf + 2 * lane, stage + 6 * C::SW + 2 * lane, qs.has_hi, qs.has_ll, wA, wB,
                       t >= C::PRO, y_ptr, colvalid, vec4);
    uu = (uu + 1 == C::UNR) ? 0 : uu + 1;
    __syncwarp();
  }
  cp_async_wait<0>();
}

template <int MQ>
inline int launch_i2_stream(const DtParams& p, cudaStream_t stream) {
  using C = I2Cfg<MQ>;
  if (!quad_inputs_ok(p) || p.W < 2 * C::HLA) return kNoFastPath;
  const int n_strips = (p.W + 63) / 64;
  const long long planes = (long long)p.N * p.C;
  int n_chunks, CH;
  static ConcCache conc_cache;
  const int conc = resident_warps_dev(conc_cache, inv_j2plus_stream<MQ>, C::SMEM_BYTES);
  pick_chunks(planes * n_strips, p.H >> 1, 8, 8, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  inv_j2plus_stream<MQ><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

int try_launch_inv_j2plus(const DtParams& p, cudaStream_t stream) {
  if ((long long)p.N * p.C == 0) return 0;
  if (p.L0 == 10) return launch_i2_stream<10>(p, stream);  // qshift_a, qshift_06
  if (p.L0 == 14) return launch_i2_stream<14>(p, stream);  // qshift_b
  if (p.L0 == 16) return launch_i2_stream<16>(p, stream);  // qshift_c
  if (p.L0 == 18) return launch_i2_stream<18>(p, stream);  // qshift_d
  return kNoFastPath;
}


}  // namespace fast
}  // namespace b200w
