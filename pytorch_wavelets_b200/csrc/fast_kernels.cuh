// fast_kernels.cuh -- specialised streaming kernels for the headline configurations (sm_100a).
// (stub: fast paths are added incrementally; every try_launch_* returns kNoFastPath when no
// specialisation applies and the caller falls back to the generic tile kernels.)
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace b200w {
namespace fast {

constexpr int kNoFastPath = 1;
static int g_force_generic = 0;

inline int try_launch_afb(const AfbParams&, cudaStream_t) { return kNoFastPath; }
inline int try_launch_fwd_j1(const DtParams&, cudaStream_t) { return kNoFastPath; }
inline int try_launch_fwd_j2plus(const DtParams&, cudaStream_t) { return kNoFastPath; }

}  // namespace fast
}  // namespace b200w
