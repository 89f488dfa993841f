// kernels_tensor.cuh — tensor-line specialised kernels (placeholder while the generic path is validated)
#pragma once
#include "common.cuh"
namespace sse {
struct TensorPlan { int ok = 0; int has_nodal = 0; int has_fluxdiff = 0; };
inline void tensor_plan_build(TensorPlan& tp, const sse_config&, const sse_arrays&, const Ops&) { tp.ok = 0; }
template <class F> inline int32_t tensor_plan_upload(TensorPlan&, F) { return SSE_OK; }
template <int D, int NC> inline cudaError_t tensor_set_attrs(const TensorPlan&) { return cudaSuccess; }
template <int D, int NC> inline void tensor_launch_nodal(const TensorPlan&, const Ops&, const Geo&, const Law&, int, const double*, double*, double*, long long, int, cudaStream_t) {}
template <int D, int NC> inline void tensor_launch_fluxdiff(const TensorPlan&, const Ops&, const Geo&, const Law&, long long, long long, const double*, const double*, double*, int, cudaStream_t) {}
}
