// handle.h — the state behind an sse_handle, shared by sse_b200.cu (single-GPU entry points) and comm.cu (multi-GPU driver)
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/sse_b200.h"
#include "common.cuh"
#include "kernels_tensor.cuh"
#include "ct_api.h"

namespace sse {
int32_t fail(int32_t code, const char* fmt, ...);     // records the message for sse_last_error_string and returns code
}
#define CU(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) return sse::fail(SSE_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// multi-GPU state of a handle (comm.cu): one NCCL rank per handle; facet-trace halos only
struct sse_comm {
    void* nccl = nullptr;           // ncclComm_t
    int rank = 0, world = 1;
    bool planned = false;
    cudaStream_t s_comm = nullptr;  // halo transfers run here, next to the interior elements on the handle's stream
    cudaEvent_t e_packed = nullptr, e_done = nullptr, e_sent = nullptr;
    std::vector<sse_handle*> peers; // one process, several handles, no NCCL (sse_comm_init_local): peers[r] = handle of rank r
    std::vector<int> nbr_rank;
    std::vector<long long> send_count, recv_count;     // facet nodes per neighbour (segments of the packed buffers, in order)
    long long n_interior = 0;       // elements [0, n_interior) read no ghost facet; [n_interior, N_e) do
};

struct sse_handle {
    sse_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> owned;       // device allocations freed in sse_destroy
    sse::Ops ops;
    sse::Geo geo;
    sse::Law law;
    sse::TensorPlan tp;             // tensor-line specialisation (kernels_tensor.cuh); tp.ok == 0 -> generic only
    sse::CtPlan ct;                 // compile-time-sized kernels (kernels_ct.cuh): Euler on p = 3, 4 ModalTensor tets
    sse::DenseDev dense{};          // dense all-pairs flux differencing (ModalMulti / NodalMulti Euler): dense.ok == 0 -> not used
    size_t smem_dense = 0;
    int variant = 1;
    int project = 0;                // 0 none, 1 nodal, 2 general entropy projection
    int second_order = 0;
    double *u_q = nullptr, *u_f = nullptr, *q_q = nullptr, *q_f = nullptr;
    size_t smem_nodal = 0, smem_time = 0, smem_aux = 0;
    int threads = 128;
    int sm_count = 148;
    long long launches = 0;         // kernels launched through this handle (sse_launch_count; bench.py's gpu_launches)
    // halo
    long long n_send = 0;
    long long* d_send_idx = nullptr;
    double *d_send = nullptr, *d_recv = nullptr;
    int halo_vars = 0;
    sse_comm comm;
    // host-buffer residual (sse_rhs_host): highest local face neighbour of every element (from mapP), device staging
    // states, copy streams and events, all created on first use
    std::vector<long long> nbr_hi;
    std::vector<long long> nbr;     // up to N_fac distinct local face neighbours per element (-1: none); empty if some element has more
    std::vector<char> nbr_ghost;    // element reads at least one ghost facet slot (halo-adjacent)
    int plan_chunks = 0;            // cached schedule of sse_rhs_host for this many ranges
    std::vector<int> plan_order, plan_ready;
    double *h2d_u = nullptr, *d2h_du = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> events;
    // CUDA-graph replay of sse_step_ck54 (sse_set_graph_mode)
    struct Graph {
        bool on = false;
        cudaStream_t stream = nullptr;
        cudaEvent_t e_in = nullptr, e_out = nullptr;
        cudaGraphExec_t exec = nullptr;
        double *u = nullptr, *tmp = nullptr, *dudt = nullptr;
        double dt = 0.0;
        int variant = -1;
        long long launches = 0;
    } graph;
    // non-finite / non-physical state detection (SSE_ERR_NONFINITE): device flag set by the kernels, read at sse_synchronize
    volatile int* h_flag = nullptr; // mapped, page-locked; Geo.flag is its device alias
};

// entry points of sse_b200.cu used by the multi-GPU driver, and of comm.cu used by the single-GPU entry points
namespace sse {
int32_t check_flag(sse_handle* h);
int32_t dist_rhs(sse_handle* h, const double* d_u, double* d_dudt, RkStage rk);
int32_t dist_rhs_host(sse_handle* h, const double* h_u, double* h_dudt, double* d_u, double* d_du, int32_t chunks);
int32_t dist_allreduce_sum(sse_handle* h, double* d_buf, int n);        // no-op without a communicator
void comm_release(sse_handle* h);
int32_t pass_b_stage(sse_handle* h, double* d_dudt, int64_t first, int64_t count, RkStage rk, cudaEvent_t mid = nullptr);
void make_range_plan_general(const std::vector<std::pair<long long, long long>>& ranges, long long ne, int nfac,
                             const std::vector<long long>& nbr, const std::vector<long long>& nbr_hi, std::vector<int>& ready);
}
