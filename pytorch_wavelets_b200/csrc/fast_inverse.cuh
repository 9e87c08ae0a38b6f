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

// one coefficient row: W pass into window slot U, then (if emit) the H pass for the output row pair
template <int L, int U>
__device__ __forceinline__ void sfb_row(const SfbParams& p, const float* srow, float (&wP)[L / 2][4],
                                        float (&wQ)[L / 2][4], bool emit, float*& y_ptr, int ypitch, int nv4,
                                        bool row1_ok, bool vec4) {
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
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      float s_ll = 0.f, s_lh = 0.f, s_hl = 0.f, s_hh = 0.f;
#pragma unroll
      for (int i = 0; i < HALF; ++i) {
        const float g0 = p.gw_lo.t[L - 2 - 2 * i + ph], g1 = p.gw_hi.t[L - 2 - 2 * i + ph];
        s_ll = fmaf(a[0][e + i], g0, s_ll);
        s_lh = fmaf(a[1][e + i], g0, s_lh);
        s_hl = fmaf(a[2][e + i], g1, s_hl);
        s_hh = fmaf(a[3][e + i], g1, s_hh);
      }
      wP[U][2 * e + ph] = __fadd_rn(s_ll, s_hl);
      wQ[U][2 * e + ph] = __fadd_rn(s_lh, s_hh);
    }
  if (emit) {
    float o[2][4];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          const int sl = (U + 1 + i) % HALF;
          s0 = fmaf(wP[sl][c], p.gh_lo.t[L - 2 - 2 * i + ph], s0);
          s1 = fmaf(wQ[sl][c], p.gh_hi.t[L - 2 - 2 * i + ph], s1);
        }
        o[ph][c] = __fadd_rn(s0, s1);
      }
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      if (ph == 1 && !row1_ok) break;
      float* q = y_ptr + ph * ypitch;
      if (vec4 && nv4 == 4) {
        *reinterpret_cast<float4*>(q) = make_float4(o[ph][0], o[ph][1], o[ph][2], o[ph][3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nv4) q[c] = o[ph][c];
      }
    }
    y_ptr += 2 * ypitch;
  }
}

template <int L, int V>
__device__ __forceinline__ void sfb_stage_dispatch(int vv, const SfbParams& p, const float* stage,
                                                   float (&wP)[L / 2][4], float (&wQ)[L / 2][4], int rho0,
                                                   int rho_end, int m0, float*& y_ptr, int ypitch, int nv4, bool vec4) {
  using C = SfbCfg<L>;
  if constexpr (V < C::UNS) {
    if (vv == V) {
      {
        const int rho = rho0;                             // coefficient row index relative to the chunk start
        const bool emit = (rho >= C::HALF - 1) && (rho < rho_end);
        const int n0 = 2 * (m0 + rho - (C::HALF - 1));    // first output row of the pair
        sfb_row<L, C::KR * V>(p, stage, wP, wQ, emit, y_ptr, ypitch, nv4, n0 + 1 < p.Ho, vec4);
      }
      if constexpr (C::KR == 2) {
        const int rho = rho0 + 1;
        const bool emit = (rho >= C::HALF - 1) && (rho < rho_end);
        const int n0 = 2 * (m0 + rho - (C::HALF - 1));
        sfb_row<L, C::KR * V + 1>(p, stage + 4 * C::SWB, wP, wQ, emit, y_ptr, ypitch, nv4, n0 + 1 < p.Ho, vec4);
      }
    } else {
      sfb_stage_dispatch<L, V + 1>(vv, p, stage, wP, wQ, rho0, rho_end, m0, y_ptr, ypitch, nv4, vec4);
    }
  }
}

template <int L>
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
  const int npairs_h = (p.Ho + 1) >> 1;
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
  bptr[0] = p.ll + (long long)plane * p.llps + c0;
  bpitch[0] = p.llpitch;
#pragma unroll
  for (int b = 1; b < 4; ++b) {
    bptr[b] = p.highs ? p.highs + ((long long)plane * 3 + (b - 1)) * band + c0 : nullptr;
    bpitch[b] = p.Wc;
  }
  const bool ok0 = (c0 + lane) < p.Wc;
  const bool ok1 = (c0 + 32 + lane) < p.Wc;
  const bool ok2 = (lane < C::HALF - 1) && ((c0 + 64 + lane) < p.Wc);

  auto issue = [&](int t) {
    if (t < n_stage) {
      float* dst = ring + (t % C::NS) * C::STAGE;
#pragma unroll
      for (int r = 0; r < C::KR; ++r) {
        const int k = m0 + C::KR * t + r;
        if (k < p.Hc && C::KR * t + r < n_rows) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (bptr[b] == nullptr) continue;
            const float* src = bptr[b] + (long long)k * bpitch[b];
            float* d = dst + (r * 4 + b) * C::SWB;
            if (ok0) cp_async4(d + lane, src + lane);
            if (ok1) cp_async4(d + 32 + lane, src + 32 + lane);
            if (ok2) cp_async4(d + 64 + lane, src + 64 + lane);
          }
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int t = 0; t < C::NS - 1; ++t) issue(t);

  float wP[C::HALF][4], wQ[C::HALF][4];
#pragma unroll
  for (int j = 0; j < C::HALF; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) { wP[j][c] = 0.f; wQ[j][c] = 0.f; }

  const int col0 = 2 * c0 + 4 * lane;
  float* y_ptr = p.y + (long long)plane * p.yps + (long long)(2 * m0) * p.ypitch + col0;
  const int nv4 = imax(0, imin(4, p.Wo - col0));
  const bool vec4 = ((p.ypitch & 3) == 0) && ((p.yps & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);

  int vv = 0;
#pragma unroll 1
  for (int t = 0; t < n_stage; ++t) {
    cp_async_wait<C::NS - 2>();
    __syncwarp();
    issue(t + C::NS - 1);
    const float* stage = ring + (t % C::NS) * C::STAGE + 2 * lane;
    sfb_stage_dispatch<L, 0>(vv, p, stage, wP, wQ, C::KR * t, n_rows, m0, y_ptr, p.ypitch, nv4, vec4);
    vv = (vv + 1 == C::UNS) ? 0 : vv + 1;
  }
  cp_async_wait<0>();
}

template <int L>
inline int launch_sfb_stream(const SfbParams& p, cudaStream_t stream) {
  using C = SfbCfg<L>;
  const int n_strips = (((p.Wo + 1) >> 1) + 63) / 64;
  int n_chunks, CH;
  pick_chunks((long long)p.planes * n_strips, (p.Ho + 1) >> 1, 16, &n_chunks, &CH);
  const long long blocks = (long long)p.planes * n_strips * n_chunks;
  if (blocks <= 0) return 0;
  if (blocks > 2147483647LL) return kNoFastPath;
  sfb2d_stream<L><<<(unsigned)blocks, 32, C::SMEM_BYTES, stream>>>(p, n_strips, n_chunks, CH);
  return 0;
}

inline int try_launch_sfb(const SfbParams& p, cudaStream_t stream) {
  if (g_force_generic) return kNoFastPath;
  if (p.Lw != p.Lh || p.mode == B200W_MODE_PERIODIZATION) return kNoFastPath;
  if (p.planes == 0) return 0;
  switch (p.Lw) {
    case 2: return launch_sfb_stream<2>(p, stream);
    case 4: return launch_sfb_stream<4>(p, stream);
    case 6: return launch_sfb_stream<6>(p, stream);
    case 8: return launch_sfb_stream<8>(p, stream);
    case 10: return launch_sfb_stream<10>(p, stream);
    case 12: return launch_sfb_stream<12>(p, stream);
    case 16: return launch_sfb_stream<16>(p, stream);
    default: return kNoFastPath;
  }
}
