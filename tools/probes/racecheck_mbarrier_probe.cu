// racecheck_mbarrier_probe.cu -- does compute-sanitizer's racecheck model mbarrier arrive / try_wait (inline PTX) as
// synchronisation?  Two warps, one shared buffer, three correct hand-offs:
//   A: warp 0 stores, __syncwarp, lane 0 mbarrier.arrive;  warp 1 waits (try_wait.parity loop), then loads
//   B: the same with mbarrier.test_wait polling on the consumer side
//   C: control -- the same data flow ordered by __syncthreads (must be clean)
// The PTX memory model orders A and B (arrive = release.cta, successful wait = acquire.cta).  If racecheck reports
// hazards for A / B and none for C, its reports on dwt_pyramid (which uses exactly these hand-offs, plus TMA complete_tx)
// are a modelling gap of the tool, not races.   nvcc -arch=sm_100a -lineinfo -o rc_probe racecheck_mbarrier_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_try(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void probe(float* out, int which) {
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ float buf[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned b0 = (unsigned)__cvta_generic_to_shared(&bars[0]);
  if (threadIdx.x == 0) { mbar_init(b0, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  if (which == 2) {                       // C: __syncthreads
    if (warp == 0) buf[lane] = (float)lane;
    __syncthreads();
    if (warp == 1) out[lane] = buf[31 - lane];
    return;
  }
  if (warp == 0) {
    buf[lane] = (float)lane;
    __syncwarp();
    if (lane == 0) mbar_arrive(b0);
  } else {
    if (which == 0) mbar_wait_try(b0, 0);
    else while (!mbar_test(b0, 0)) __nanosleep(50);
    out[lane] = buf[31 - lane];
  }
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  float* d; cudaMalloc(&d, 32 * sizeof(float));
  probe<<<1, 64>>>(d, which);
  cudaError_t e = cudaDeviceSynchronize();
  float h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("case %d: %s, out[0] = %g (expect 31)\n", which, cudaGetErrorString(e), h[0]);
  return 0;
}
