// fast_inverse.cuh -- streaming DWT synthesis kernel (included inside namespace b200w::fast).
//
// One warp owns 64 coefficient columns (= 128 output columns) of one plane and marches down the
// coefficient rows.  The four subband rows (ll, lh, hl, hh) of each coefficient row are staged in a
// per-warp shared-memory ring with 4-byte cp.async (the band-pass tensors the caller hands in are
// contiguous with odd widths, so 16-byte alignment cannot be assumed).  The pass along W runs first on
// the staged rows (64-bit conflict-free LDS), the pass along H runs in a rotating register window --
// the two passes commute exactly in real arithmetic; in fp32 the result differs from the H-then-W order
// of the generic kernel / oracle by rounding only (<= 1e-6 relative, tests use the 1e-5 tolerance).
//   y[2c+ph] = sum_{i<L/2} a[c+i] g[L-2-2i+ph]   (non-periodization synthesis, reference sfb1d :263-267)

template <int L>
struct SfbCfg {
  static constexpr int HALF = L / 2;
  static constexpr int SWB = 96;                                  // staged floats per band row (3 x 32 lanes)
  static constexpr int KR = (HALF % 2 == 0) ? 2 : 1;              // coefficient rows per stage
  static constexpr int UNS = HALF / KR;                           // window period in stages
  static constexpr int NS = 3;
  static constexpr int STAGE = KR * 4 * SWB;                      // floats per stage
  static constexpr int SMEM_BYTES = NS * STAGE * 4;
  static constexpr int NVB = (HALF + 1 + 1) / 2;                  // 64-bit loads per band row per lane
  static_assert(HALF - 1 <= 32, "halo must fit the third 32-lane copy");
};

// one coefficient row: W pass into window slot U, then (if emit) the H pass for the output row pair.
// Packed FMA throughout: along W a coefficient times the (even, odd)-phase tap pair gives both output columns it
// feeds; along H a tap times a window column pair gives two adjacent outputs of one row.
// PER (periodization): the same sums over the periodic extension of the coefficients give y[(n' + L/2 - 1) mod 2K]
// (reference sfb1d :252-261 re-indexed: 2k + j - (L/2 - 1) = n  <=>  n' = n - L/2 + 1 with n' = 2c + phase,
// a[(c + i) mod K]); only the staging (wrapped rows / columns) and the store positions differ.
template <int L, int U, bool PER>
__device__ __forceinline__ void sfb_row(const SfbParams& p, const float* srow, float2 (&wP)[L / 2][2],
                                        float2 (&wQ)[L / 2][2], bool emit, float*& y_ptr, int ypitch, int nv4,
                                        bool row1_ok, bool vec4, const int (&ncol)[4], int nr0, int nr1) {
  using C = SfbCfg<L>;
  constexpr int HALF = C::HALF;
  float a[4][2 * C::NVB];  // [band][window]
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int q = 0; q < C::NVB; ++q) {
      const float2 v = *reinterpret_cast<const float2*>(srow + b * C::SWB + 2 * q);
      a[b][2 * q] = v.x; a[b][2 * q + 1] = v.y;
    }
  // W pass: P = S(ll; gw_lo) + S(hl; gw_hi), Q = S(lh; gw_lo) + S(hh; gw_hi)   (bands: 0 ll, 1 lh, 2 hl, 3 hh)
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    float2 s_ll = make_float2(0.f, 0.f), s_lh = s_ll, s_hl = s_ll, s_hh = s_ll;   // {phase 0, phase 1}
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      const float2 g0 = make_float2(p.gw_lo.t[L - 2 - 2 * i], p.gw_lo.t[L - 1 - 2 * i]);
      const float2 g1 = make_float2(p.gw_hi.t[L - 2 - 2 * i], p.gw_hi.t[L - 1 - 2 * i]);
      s_ll = ffma2_s(a[0][e + i], g0, s_ll);
      s_lh = ffma2_s(a[1][e + i], g0, s_lh);
      s_hl = ffma2_s(a[2][e + i], g1, s_hl);
      s_hh = ffma2_s(a[3][e + i], g1, s_hh);
    }
    wP[U][e] = make_float2(__fadd_rn(s_ll.x, s_hl.x), __fadd_rn(s_ll.y, s_hl.y));
    wQ[U][e] = make_float2(__fadd_rn(s_lh.x, s_hh.x), __fadd_rn(s_lh.y, s_hh.y));
  }
  if (emit) {
    float2 o[2][2];  // [output row of the pair][column pair]
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          const int sl = (U + 1 + i) % HALF;
          s0 = ffma2_s(p.gh_lo.t[L - 2 - 2 * i + ph], wP[sl][e], s0);
          s1 = ffma2_s(p.gh_hi.t[L - 2 - 2 * i + ph], wQ[sl][e], s1);
        }
        o[ph][e] = make_float2(__fadd_rn(s0.x, s1.x), __fadd_rn(s0.y, s1.y));
      }
    if constexpr (PER) {
      // y_ptr = plane base; nr0 / nr1 = rotated output rows of the pair (-1: outside the requested output),
      // ncol[] = rotated output columns of the lane's four values
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const int nr = ph ? nr1 : nr0;
        if (nr < 0) continue;
        float* q = y_ptr + (long long)nr * ypitch;
        if (ncol[0] >= 0) q[ncol[0]] = o[ph][0].x;
        if (ncol[1] >= 0) q[ncol[1]] = o[ph][0].y;
        if (ncol[2] >= 0) q[ncol[2]] = o[ph][1].x;
        if (ncol[3] >= 0) q[ncol[3]] = o[ph][1].y;
      }
    } else {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        if (ph == 1 && !row1_ok) break;
        float* q = y_ptr + ph * ypitch;
        if (vec4 && nv4 == 4) {
          *reinterpret_cast<float4*>(q) = make_float4(o[ph][0].x, o[ph][0].y, o[ph][1].x, o[ph][1].y);
        } else {
          if (0 < nv4) q[0] = o[ph][0].x;
          if (1 < nv4) q[1] = o[ph][0].y;
          if (2 < nv4) q[2] = o[ph][1].x;
          if (3 < nv4) q[3] = o[ph][1].y;
        }
      }
      y_ptr += 2 * ypitch;
    }
  }
}

template <int L, int V, bool PER>
__device__ __forceinline__ void sfb_stage_dispatch(int vv, const SfbParams& p, const float* stage,
                                                   float2 (&wP)[L / 2][2], float2 (&wQ)[L / 2][2], int rho0,
                                                   int rho_end, int m0, float*& y_ptr, int ypitch, int nv4, bool vec4,
                                                   const int (&ncol)[4]) {
  using C = SfbCfg<L>;
  if constexpr (V < C::UNS) {
    if (vv == V) {
#pragma unroll
      for (int r = 0; r < C::KR; ++r) {
        const int rho = rho0 + r;                         // coefficient row index relative to the chunk start
        const bool emit = (rho >= C::HALF - 1) && (rho < rho_end);
        const int n0 = 2 * (m0 + rho - (C::HALF - 1));    // first output row of the pair (before rotation if PER)
        int nr0 = -1, nr1 = -1;
        if (PER && emit) {
          const int N = 2 * p.Hc;
          nr0 = n0 + C::HALF - 1; if (nr0 >= N) nr0 -= N;
          nr1 = n0 + C::HALF;     if (nr1 >= N) nr1 -= N;
          if (nr0 >= p.Ho) nr0 = -1;
          if (nr1 >= p.Ho) nr1 = -1;
        }
        if (r == 0)
          sfb_row<L, C::KR * V, PER>(p, stage, wP, wQ, emit, y_ptr, ypitch, nv4, n0 + 1 < p.Ho, vec4, ncol, nr0, nr1);
        else
          sfb_row<L, C::KR * V + (C::KR - 1), PER>(p, stage + 4 * C::SWB, wP, wQ, emit, y_ptr, ypitch, nv4, n0 + 1 < p.Ho,
                                                   vec4, ncol, nr0, nr1);
      }
    } else {
      sfb_stage_dispatch<L, V + 1, PER>(vv, p, stage, wP, wQ, rho0, rho_end, m0, y_ptr, ypitch, nv4, vec4, ncol);
    }
  }
}

template <int L, bool PER = false>
__global__ void __launch_bounds__(32) sfb2d_stream(const __grid_constant__ SfbParams p, int n_strips, int n_chunks,
                                                   int CH /* output row pairs per chunk */) {
  using C = SfbCfg<L>;
  extern __shared__ __align__(16) float ring[];
  const int lane = threadIdx.x;
  long long item = blockIdx.x;
  const int strip = (int)(item % n_strips);
  item /= n_strips;
  const int chunk = (int)(item % n_chunks);
  const int plane = (int)(item / n_chunks);

  const int c0 = strip * 64;                       // first coefficient column (= pair index) of the strip
  const int npairs_h = PER ? p.Hc : (p.Ho + 1) >> 1;
  const int m0 = chunk * CH;
  const int m1 = imin(m0 + CH, npairs_h);
  const int n_rows = (m1 - m0) + C::HALF - 1;      // coefficient rows m0 .. m1-1+HALF-1
  const int n_stage = (n_rows + C::KR - 1) / C::KR;

  // zero the ring once: positions that are never copied (columns beyond Wc, absent band-passes) must read 0
  for (int i = lane; i < C::NS * C::STAGE; i += 32) ring[i] = 0.f;
  __syncwarp();

  const long long band = (long long)p.Hc * p.Wc;
  const float* bptr[4];
  int bpitch[4];
  bptr[0] = p.ll + (long long)plane * p.llps;
  bpitch[0] = p.llpitch;
#pragma unroll
  for (int b = 1; b < 4; ++b) {
    bptr[b] = p.highs ? p.highs + ((long long)plane * 3 + (b - 1)) * band : nullptr;
    bpitch[b] = p.Wc;
  }
  // the three 32-lane column copies of a band row: coefficient columns c0 + lane + {0, 32, 64}; PER wraps them
  int colw[3];
  bool okc[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int cidx = c0 + lane + 32 * j;
    const bool in_lanes = (j < 2) || (lane < C::HALF - 1);
    okc[j] = in_lanes && (PER ? (cidx < p.Wc + C::HALF - 1) : (cidx < p.Wc));
    colw[j] = PER ? cidx % p.Wc : cidx;
  }

  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + 4 * lane;
  int slot_i = 0;
  auto issue = [&](int t) {
    const int slot = slot_i;
    slot_i = (slot_i + 1 == C::NS) ? 0 : slot_i + 1;
    if (t < n_stage) {
      const unsigned dst = ring_s + slot * (C::STAGE * 4);
#pragma unroll
      for (int r = 0; r < C::KR; ++r) {
        int k = m0 + C::KR * t + r;
        bool row_ok = (C::KR * t + r < n_rows);
        if (PER) k %= p.Hc; else row_ok = row_ok && (k < p.Hc);
        if (row_ok) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (bptr[b] == nullptr) continue;
            const float* src = bptr[b] + (long long)k * bpitch[b];
            const unsigned d = dst + (r * 4 + b) * (C::SWB * 4);
            if (okc[0]) cp_async4_s(d, src + colw[0]);
            if (okc[1]) cp_async4_s(d + 128, src + colw[1]);
            if (okc[2]) cp_async4_s(d + 256, src + colw[2]);
          }
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int t = 0; t < C::NS - 1; ++t) issue(t);

  float2 wP[C::HALF][2], wQ[C::HALF][2];
#pragma unroll
  for (int j = 0; j < C::HALF; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) { wP[j][c] = make_float2(0.f, 0.f); wQ[j][c] = make_float2(0.f, 0.f); }

  const int col0 = 2 * c0 + 4 * lane;
  float* y_ptr = PER ? p.y + (long long)plane * p.yps
                     : p.y + (long long)plane * p.yps + (long long)(2 * m0) * p.ypitch + col0;
  const int nv4 = imax(0, imin(4, p.Wo - col0));
  int ncol[4] = {-1, -1, -1, -1};
  if (PER) {
    const int N = 2 * p.Wc;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int c = col0 + q + C::HALF - 1;
      if (c >= N) c -= N;
      ncol[q] = (col0 + q < N && c < p.Wo) ? c : -1;
    }
  }
  const bool vec4 = ((p.ypitch & 3) == 0) && ((p.yps & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);

  int vv = 0, slot_a = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    cp_async_wait<C::NS - 2>();
    __syncwarp();
    issue(t + C::NS - 1);
    const float* stage = ring + slot_a * C::STAGE + 2 * lane;
    slot_a = (slot_a + 1 == C::NS) ? 0 : slot_a + 1;
    sfb_stage_dispatch<L, 0, PER>(vv, p, stage, wP, wQ, C::KR * t, n_rows, m0, y_ptr, p.ypitch, nv4, vec4, ncol);
    vv = (vv + 1 == C::UNS) ? 0 : vv + 1;
  }
  cp_async_wait<0>();
}

template <int L, bool PER>
inline int launch_sfb_stream_m(const SfbParams& p, cudaStream_t stream) {
  using C = SfbCfg<L>;
  const int npairs_w = PER ? p.Wc : (p.Wo + 1) >> 1;
  const int npairs_h = PER ? p.Hc : (p.Ho + 1) >> 1;
  const int n_strips = (npairs_w + 63) / 64;
  int n_chunks, CH;
  static const int conc = resident_warps(sfb2d_stream<L, PER>, C::SMEM_BYTES);
  pick_chunks((long long)p.planes * n_strips, npairs_h, 16, L / 2 + 8, conc, &n_chunks, &CH);
  const long long blocks = (long long)p.planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  sfb2d_stream<L, PER><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

template <int L>
inline int launch_sfb_stream(const SfbParams& p, cudaStream_t stream) {
  if (p.mode == B200W_MODE_PERIODIZATION) return launch_sfb_stream_m<L, true>(p, stream);
  return launch_sfb_stream_m<L, false>(p, stream);
}

inline int try_launch_sfb(const SfbParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
  if (p.Lw != p.Lh) return kNoFastPath;
  if (p.planes == 0) return 0;
  switch (p.Lw) {
    case 2: return launch_sfb_stream<2>(p, stream);
    case 4: return launch_sfb_stream<4>(p, stream);
    case 6: return launch_sfb_stream<6>(p, stream);
    case 8: return launch_sfb_stream<8>(p, stream);
    case 10: return launch_sfb_stream<10>(p, stream);
    case 12: return launch_sfb_stream<12>(p, stream);
    case 14: return launch_sfb_stream<14>(p, stream);
    case 16: return launch_sfb_stream<16>(p, stream);
    case 18: return launch_sfb_stream<18>(p, stream);
    case 20: return launch_sfb_stream<20>(p, stream);
    default: return kNoFastPath;
  }
}

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
  static constexpr int NS = 3;
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
__global__ void __launch_bounds__(32) inv_j1_stream(const __grid_constant__ DtParams p, int n_strips, int n_chunks,
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
  static const int conc = resident_warps(inv_j1_stream<L0, L1>, C::SMEM_BYTES);
  pick_chunks(planes * n_strips, p.H >> 1, 8, 8, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  inv_j1_stream<L0, L1><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

inline int try_launch_inv_j1(const DtParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
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
  static constexpr int NS = 3;
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

template <int MQ>
__global__ void __launch_bounds__(32) inv_j2plus_stream(const __grid_constant__ DtParams p, int n_strips,
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
  static const int conc = resident_warps(inv_j2plus_stream<MQ>, C::SMEM_BYTES);
  pick_chunks(planes * n_strips, p.H >> 1, 8, 8, conc, &n_chunks, &CH);
  const long long blocks = planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  inv_j2plus_stream<MQ><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

inline int try_launch_inv_j2plus(const DtParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
  if ((long long)p.N * p.C == 0) return 0;
  if (p.L0 == 10) return launch_i2_stream<10>(p, stream);  // qshift_a, qshift_06
  if (p.L0 == 14) return launch_i2_stream<14>(p, stream);  // qshift_b
  if (p.L0 == 16) return launch_i2_stream<16>(p, stream);  // qshift_c
  if (p.L0 == 18) return launch_i2_stream<18>(p, stream);  // qshift_d
  return kNoFastPath;
}
