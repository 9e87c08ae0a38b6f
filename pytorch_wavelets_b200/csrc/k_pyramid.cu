// k_pyramid.cu -- translation unit of the fused DWT pyramid kernel (dwt_pyramid.cuh, sm_100a)
#include "dwt_pyramid.cuh"

namespace b200w {
namespace fast {

static int max_optin_smem() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) {
      (void)cudaGetLastError();
      v = 227 * 1024;
    }
    cache[dev] = v;
  }
  return cache[dev];
}

static int sm_smem() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 228 * 1024;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess || v <= 0) {
      (void)cudaGetLastError();
      v = 228 * 1024;
    }
    cache[dev] = v;
  }
  return cache[dev];
}

int plan_dwt_pyramid(PyrParams& p, const float* x, long long xps, int xpitch, int planes, int H, int W, int J, int L,
                     int mode, int ll_pitch) {
  return plan_pyramid_best(p, planes, H, W, J, L, mode, xps, xpitch, x, max_optin_smem(), sm_smem(), ll_pitch)
             ? kNoFastPath : 0;
}

int launch_dwt_pyramid(const PyrParams& p, cudaStream_t stream) {
  switch (p.L) {
    case 2: return launch_pyramid<2>(p, stream);
    case 4: return launch_pyramid<4>(p, stream);
    case 6: return launch_pyramid<6>(p, stream);
    case 8: return launch_pyramid<8>(p, stream);
    case 10: return launch_pyramid<10>(p, stream);
    case 12: return launch_pyramid<12>(p, stream);
    default: return kNoFastPath;
  }
}

}  // namespace fast
}  // namespace b200w

#ifdef B200W_PYR_PROF
// variant builds only (tools/pyr_prof.py): read and clear the wait-time counters of dwt_pyramid
extern "C" int b200w_debug_pyr_prof(unsigned long long* out) {
  unsigned long long z[64] = {};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out, b200w::fast::g_pyr_prof, sizeof(z)) != cudaSuccess) return -2;
  if (cudaMemcpyToSymbol(b200w::fast::g_pyr_prof, z, sizeof(z)) != cudaSuccess) return -3;
  return 0;
}
#endif
