// Shared host/device helpers for the h2gcn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/h2gcn_b200.h"

namespace h2 {

#ifndef H2_GATHER_WARPS
#define H2_GATHER_WARPS 8
#endif
constexpr int kWarpsPerCta = H2_GATHER_WARPS;   // CSR gather kernel: warps (= rows) per CTA
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kNumSms = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return H2_ERR_CUDA;
}

#define H2_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return h2::cuda_fail(e__, #call); \
    } while (0)

#define H2_REQUIRE(cond, code, ...)   \
    do {                              \
        if (!(cond)) {                \
            h2::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

// after a <<<>>> launch
#define H2_LAUNCHED(name)                                         \
    do {                                                          \
        h2::g_launches.fetch_add(1, std::memory_order_relaxed);   \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return h2::cuda_fail(e__, name);  \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- plan (schedule of the fused round) ----------------------------------------------------------------------
struct PlanHost {
    uint32_t magic;
    int32_t n_rows;
    int32_t n_hops;
    int32_t cta_threshold;  // virtual rows with more stored entries than this are processed by a whole CTA
    int64_t n_vrows;        // n_rows * n_hops
    int64_t n_cta_rows;     // leading entries of perm[] handled one-per-CTA
    int64_t total_nnz;
    int64_t max_row_nnz;
};
constexpr uint32_t kPlanMagic = 0x48324731u;  // "H2G1"

struct PlanCounts {  // device mirror written by the plan kernels
    int64_t n_cta_rows;
    int64_t total_nnz;
    int64_t max_row_nnz;
};

}  // namespace h2
