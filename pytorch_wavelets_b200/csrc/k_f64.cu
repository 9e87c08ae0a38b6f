// k_f64.cu -- float64 instantiation of the generic per-level kernels (sm_100a).
//
// The reference computes in the default dtype, so torch.float64 modules and inputs work there
// (dwt/lowlevel.py:972 `torch.get_default_dtype()`, dtcwt/lowlevel.py:67; its tests run both precisions,
// tests/test_dwt.py:143-160, tests/test_dtcwt.py:116-135).  The B200 fast paths (streaming and pyramid kernels, packed
// FFMA2) are float32-only by design; double precision takes the generic tile kernels and the 1-D row kernels, compiled
// here a second time with the element type redefined: same source, same index logic, same accumulation order, IEEE
// double arithmetic (fma / __dmul_rn / __dsqrt_rn).  Entry points: the C ABI names with an `_f64` suffix, same
// arguments with `double` in place of `float` (include/b200wave.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define B200W_F64 1
#define float double
#define fmaf fma
#define b200w b200w_f64   /* the namespace: the double-precision parameter blocks must not collide with the float ones */
/* the prototypes of include/b200wave.h become the prototypes of the _f64 entry points defined in this unit */
#define b200w_dwt_afb2d b200w_dwt_afb2d_f64
#define b200w_dwt_sfb2d b200w_dwt_sfb2d_f64
#define b200w_dwt_afb1d b200w_dwt_afb1d_f64
#define b200w_dwt_sfb1d b200w_dwt_sfb1d_f64
#define b200w_dtcwt_fwd_j1 b200w_dtcwt_fwd_j1_f64
#define b200w_dtcwt_fwd_j2plus b200w_dtcwt_fwd_j2plus_f64
#define b200w_dtcwt_inv_j1 b200w_dtcwt_inv_j1_f64
#define b200w_dtcwt_inv_j2plus b200w_dtcwt_inv_j2plus_f64
#define b200w_scat_j1 b200w_scat_j1_f64

#include "dwt1d.cu"   /* afb1d_rows / sfb1d_rows + b200w_dwt_afb1d_f64 / b200w_dwt_sfb1d_f64 (pulls in launch_params.h) */

namespace b200w {
constexpr int NT64 = 256;
extern __shared__ __align__(16) float g_smem64[];

__global__ void __launch_bounds__(NT64) k64_afb2d_tile(const __grid_constant__ AfbParams p) { afb2d_tile<NT64>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_sfb2d_tile(const __grid_constant__ SfbParams p) { sfb2d_tile<NT64>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_fwd_j1_tile(const __grid_constant__ DtParams p) { fwd_j1_tile<NT64, false>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_scat_j1_tile(const __grid_constant__ DtParams p) { fwd_j1_tile<NT64, true>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_fwd_j2plus_tile(const __grid_constant__ DtParams p) { fwd_j2plus_tile<NT64>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_inv_j1_tile(const __grid_constant__ DtParams p) { inv_j1_tile<NT64>(p, blockIdx.x, g_smem64); }
__global__ void __launch_bounds__(NT64) k64_inv_j2plus_tile(const __grid_constant__ DtParams p) { inv_j2plus_tile<NT64>(p, blockIdx.x, g_smem64); }

template <class K, class P>
static int launch_tile64(K kernel, const P& p, long long blocks, int smem_elems, void* stream) {
  if (blocks == 0) return B200W_OK;
  const size_t bytes = (size_t)smem_elems * sizeof(float);
  if (bytes > 227 * 1024) return B200W_EFILTER;
  if (bytes > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return B200W_ECUDA;
  }
  kernel<<<(unsigned)blocks, NT64, bytes, (cudaStream_t)stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? B200W_OK : B200W_ECUDA;
}
}  // namespace b200w

extern "C" {

int b200w_dwt_afb2d(const float* x, long long x_plane_stride, int x_pitch, float* ll, long long ll_plane_stride,
                    int ll_pitch, float* highs, int planes, int H, int W, const float* fw_lo, const float* fw_hi,
                    int Lw, const float* fh_lo, const float* fh_hi, int Lh, int mode, void* stream) {
  AfbParams p;
  const int rc = build_afb(p, x, x_plane_stride, x_pitch, ll, ll_plane_stride, ll_pitch, highs, planes, H, W, fw_lo,
                           fw_hi, Lw, fh_lo, fh_hi, Lh, mode);
  if (rc) return rc;
  return launch_tile64(k64_afb2d_tile, p, (long long)planes * p.tiles_x * p.tiles_y, afb_smem_floats(Lw, Lh), stream);
}

int b200w_dwt_sfb2d(const float* ll, long long ll_plane_stride, int ll_pitch, const float* highs, float* y,
                    long long y_plane_stride, int y_pitch, int planes, int Hc, int Wc, int Ho, int Wo,
                    const float* gh_lo, const float* gh_hi, int Lh, const float* gw_lo, const float* gw_hi, int Lw,
                    int mode, void* stream) {
  SfbParams p;
  const int rc = build_sfb(p, ll, ll_plane_stride, ll_pitch, highs, y, y_plane_stride, y_pitch, planes, Hc, Wc, Ho,
                           Wo, gh_lo, gh_hi, Lh, gw_lo, gw_hi, Lw, mode);
  if (rc) return rc;
  return launch_tile64(k64_sfb2d_tile, p, (long long)planes * p.tiles_x * p.tiles_y, sfb_smem_floats(Lh, Lw), stream);
}

int b200w_dtcwt_fwd_j1(const float* x, long long x_plane_stride, int x_pitch, float* ll, long long ll_plane_stride,
                       int ll_pitch, float* highs, const long long hs[6], int N, int C, int H, int W,
                       const float* h0, int L0, const float* h1, int L1, int mode, void* stream) {
  DtParams p;
  const int rc = build_fwd_j1(p, x, x_plane_stride, x_pitch, ll, ll_plane_stride, ll_pitch, highs, hs, N, C, H, W, h0,
                              L0, h1, L1, mode);
  if (rc) return rc;
  return launch_tile64(k64_fwd_j1_tile, p, (long long)N * C * p.tiles_x * p.tiles_y, fwdj1_smem_floats(L0, L1), stream);
}

int b200w_dtcwt_fwd_j2plus(const float* x, long long x_plane_stride, int x_pitch, float* ll,
                           long long ll_plane_stride, int ll_pitch, float* highs, const long long hs[6], int N,
                           int C, int H, int W, const float* h0a, const float* h1a, const float* h0b,
                           const float* h1b, int m, void* stream) {
  DtParams p;
  const int rc = build_fwd_j2plus(p, x, x_plane_stride, x_pitch, ll, ll_plane_stride, ll_pitch, highs, hs, N, C, H, W,
                                  h0a, h1a, h0b, h1b, m);
  if (rc) return rc;
  return launch_tile64(k64_fwd_j2plus_tile, p, (long long)N * C * p.tiles_x * p.tiles_y, fwdj2_smem_floats(m), stream);
}

int b200w_dtcwt_inv_j1(const float* ll, long long ll_plane_stride, int ll_pitch, const float* highs,
                       const long long hs[6], float* y, long long y_plane_stride, int y_pitch, int N, int C, int H,
                       int W, const float* g0, int L0, const float* g1, int L1, int mode, void* stream) {
  DtParams p;
  const int rc = build_inv_j1(p, ll, ll_plane_stride, ll_pitch, highs, hs, y, y_plane_stride, y_pitch, N, C, H, W, g0,
                              L0, g1, L1, mode);
  if (rc) return rc;
  return launch_tile64(k64_inv_j1_tile, p, (long long)N * C * p.tiles_x * p.tiles_y, invj1_smem_floats(L0, L1), stream);
}

int b200w_dtcwt_inv_j2plus(const float* ll, long long ll_plane_stride, int ll_pitch, const float* highs,
                           const long long hs[6], float* y, long long y_plane_stride, int y_pitch, int N, int C,
                           int H, int W, const float* g0a, const float* g1a, const float* g0b, const float* g1b,
                           int m, void* stream) {
  DtParams p;
  const int rc = build_inv_j2plus(p, ll, ll_plane_stride, ll_pitch, highs, hs, y, y_plane_stride, y_pitch, N, C, H, W,
                                  g0a, g1a, g0b, g1b, m);
  if (rc) return rc;
  return launch_tile64(k64_inv_j2plus_tile, p, (long long)N * C * p.tiles_x * p.tiles_y, invj2_smem_floats(m), stream);
}

int b200w_scat_j1(const float* x, float* z, float* dre_dr, float* dim_dr, int N, int C, int H, int W,
                  const float* h0, int L0, const float* h1, int L1, int mode, float magbias, void* stream) {
  DtParams p;
  const int rc = build_scat_j1(p, x, z, dre_dr, dim_dr, N, C, H, W, h0, L0, h1, L1, mode, magbias);
  if (rc) return rc;
  return launch_tile64(k64_scat_j1_tile, p, (long long)N * C * p.tiles_x * p.tiles_y, fwdj1_smem_floats(L0, L1), stream);
}

}  // extern "C"
