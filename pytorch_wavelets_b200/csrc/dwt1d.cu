// dwt1d.cu -- 1-D DWT analysis / synthesis levels (sm_100a): AFB1D / SFB1D of the reference (dwt/lowlevel.py:368-424,
// 697-743 = afb1d / sfb1d along the last dimension of (N, C, L) signals), the callers' side of the 2-D banks
// (SURVEY.md 8(f) rank 4).  HBM-bound and tiny per element: one CTA stages a contiguous segment of one signal (with
// the boundary extension resolved by index arithmetic while staging) in shared memory with coalesced loads, every thread
// produces one output position of both bands, stores are coalesced.  Accumulation order = the oracle's (stored tap
// order, first term a product, then fused multiply-adds), so the analysis is bit-identical to it.
#include <cuda_runtime.h>

#include "launch_params.h"

namespace b200w {

constexpr int kT1 = 256;   // outputs per CTA (= threads)

struct Afb1dParams {
  const float* x; long long xpitch;
  float* lo; float* hi;
  int rows, N, K, L, mode, tiles;
  Taps f0, f1;
};

struct Sfb1dParams {
  const float* lo; const float* hi;
  float* y;
  int rows, K, Nout, L, mode, tiles;
  Taps g0, g1;
};

// out[k] = sum_j f[j] xe[2k + j - pl]  (reference afb1d :134-168)
__global__ void __launch_bounds__(kT1) afb1d_rows(const __grid_constant__ Afb1dParams p) {
  extern __shared__ __align__(16) float seg[];
  const int tile = blockIdx.x % p.tiles, row = blockIdx.x / p.tiles;
  const int k0 = tile * kT1;
  const int pl = (p.mode == B200W_MODE_PERIODIZATION) ? (p.L - 1 - p.L / 2) : (p.L - 2);
  const int nseg = 2 * kT1 + p.L;                       // extended positions [2*k0 - pl, 2*k0 - pl + nseg)
  const float* xr = p.x + (long long)row * p.xpitch;
  for (int i = threadIdx.x; i < nseg; i += kT1) {
    const int g = ext_index(2 * k0 - pl + i, p.N, p.mode);
    seg[i] = (g >= 0) ? xr[g] : 0.f;
  }
  __syncthreads();
  const int k = k0 + threadIdx.x;
  if (k >= p.K) return;
  const float* s = seg + 2 * threadIdx.x;
  float a0 = B200W_MUL(p.f0.t[0], s[0]), a1 = B200W_MUL(p.f1.t[0], s[0]);
  for (int j = 1; j < p.L; ++j) {
    a0 = fmaf(p.f0.t[j], s[j], a0);
    a1 = fmaf(p.f1.t[j], s[j], a1);
  }
  p.lo[(long long)row * p.K + k] = a0;
  p.hi[(long long)row * p.K + k] = a1;
}

// y[n] = sum_k lo[k] g0[s - 2k] + sum_k hi[k] g1[s - 2k], s = n + off  (reference sfb1d :252-267)
__global__ void __launch_bounds__(kT1) sfb1d_rows(const __grid_constant__ Sfb1dParams p) {
  extern __shared__ __align__(16) float seg[];
  const int tile = blockIdx.x % p.tiles, row = blockIdx.x / p.tiles;
  const int n0 = tile * kT1;
  const bool per = (p.mode == B200W_MODE_PERIODIZATION);
  const int off = per ? (p.L / 2 - 1) : (p.L - 2);
  // coefficient indices needed by the tile: k in [floor((n0 + off - L + 2) / 2), floor((n0 + kT1 - 1 + off) / 2)]
  const int kb = floordiv2(n0 + off - p.L + 2);
  const int nk = kT1 / 2 + p.L / 2 + 2;
  float* slo = seg;
  float* shi = seg + nk;
  const float* lo = p.lo + (long long)row * p.K;
  const float* hi = p.hi ? p.hi + (long long)row * p.K : nullptr;
  for (int i = threadIdx.x; i < nk; i += kT1) {
    int k = kb + i;
    bool ok = true;
    if (per) { k %= p.K; if (k < 0) k += p.K; }
    else ok = (k >= 0 && k < p.K);
    slo[i] = ok ? lo[k] : 0.f;
    shi[i] = (ok && hi) ? hi[k] : 0.f;
  }
  __syncthreads();
  const int n = n0 + threadIdx.x;
  if (n >= p.Nout) return;
  const int s = n + off;
  const int kmin = floordiv2(s - p.L + 2), kmax = floordiv2(s);
  float a0 = 0.f, a1 = 0.f;
  bool first = true;
  for (int k = kmin; k <= kmax; ++k) {
    if (!per && (k < 0 || k >= p.K)) continue;          // the oracle skips them too (they do not start the sum)
    const int t = s - 2 * k;
    const float vl = slo[k - kb], vh = shi[k - kb];
    if (first) { a0 = B200W_MUL(vl, p.g0.t[t]); a1 = B200W_MUL(vh, p.g1.t[t]); first = false; }
    else { a0 = fmaf(vl, p.g0.t[t], a0); a1 = fmaf(vh, p.g1.t[t], a1); }
  }
  p.y[(long long)row * p.Nout + n] = B200W_ADD(a0, a1);
}

}  // namespace b200w

using namespace b200w;

extern "C" {

int b200w_dwt_afb1d(const float* x, long long x_pitch, int rows, int N, float* lo, float* hi, const float* f0,
                    const float* f1, int L, int mode, void* stream) {
  if (!dwt_mode_ok(mode)) return B200W_EMODE;
  if (!x || !lo || !hi) return B200W_EARG;
  if (rows < 0 || N < 1 || x_pitch < N) return B200W_ESIZE;
  if (L < 2) return B200W_EFILTER;
  Afb1dParams p;
  int rc;
  if ((rc = set_taps(p.f0, f0, L)) || (rc = set_taps(p.f1, f1, L))) return rc;
  p.x = x; p.xpitch = x_pitch; p.lo = lo; p.hi = hi;
  p.rows = rows; p.N = N; p.K = coeff_len(N, L, mode); p.L = L; p.mode = mode;
  p.tiles = cdiv(p.K, kT1);
  const long long blocks = (long long)rows * p.tiles;
  if (blocks == 0) return B200W_OK;
  if (!grid_ok(blocks)) return B200W_ESIZE;
  afb1d_rows<<<(unsigned)blocks, kT1, (2 * kT1 + L) * sizeof(float), (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200W_OK : B200W_ECUDA;
}

int b200w_dwt_sfb1d(const float* lo, const float* hi, int rows, int K, float* y, int Nout, const float* g0,
                    const float* g1, int L, int mode, void* stream) {
  if (!dwt_mode_ok(mode)) return B200W_EMODE;
  if (!lo || !y) return B200W_EARG;
  if (rows < 0 || K < 1) return B200W_ESIZE;
  if (L < 2) return B200W_EFILTER;
  if (Nout < 1 || Nout > rec_len(K, L, mode)) return B200W_ESIZE;
  Sfb1dParams p;
  int rc;
  if ((rc = set_taps(p.g0, g0, L)) || (rc = set_taps(p.g1, g1, L))) return rc;
  p.lo = lo; p.hi = hi; p.y = y;
  p.rows = rows; p.K = K; p.Nout = Nout; p.L = L; p.mode = mode;
  p.tiles = cdiv(Nout, kT1);
  const long long blocks = (long long)rows * p.tiles;
  if (blocks == 0) return B200W_OK;
  if (!grid_ok(blocks)) return B200W_ESIZE;
  sfb1d_rows<<<(unsigned)blocks, kT1, 2 * (kT1 / 2 + L / 2 + 2) * sizeof(float), (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200W_OK : B200W_ECUDA;
}

}  // extern "C"
