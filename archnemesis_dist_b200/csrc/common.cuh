// common.cuh -- shared helpers for libansb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ansb200.h"

struct ansb200_table {
    double *K;    // resident copy, plane-major [NP*NT][NWAVE][NG][NGAS] (NULL when stored as float32)
    double *lnK;  // same layout: log(K), -inf for K==0, NaN for K<0 (NULL when stored as float32)
    float *Kf;    // float32 storage variants (ansb200_table_create_ex): K as float32 -- lossless for .kta data,
    float *lnKf;  // which are float32 on disk -- and optionally ln K as float32 (lossy, ~4e-6 relative in k)
    int NWAVE, NG, NP, NT, NGAS;
    int storage;  // ANSB200_TABLE_F64 / _K32 / _F32
};

// What the kernels see of a table: exactly one pointer of each pair is set (uniform over the launch).
struct AnsTab {
    const double *lnK, *K;
    const float *lnKf, *Kf;
};
inline AnsTab ans_tab(const ansb200_table *t) { return AnsTab{t->lnK, t->K, t->lnKf, t->Kf}; }

void ansb200_set_error(const char *fmt, ...);

#define ANS_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            ansb200_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                              cudaGetErrorString(err__));                                         \
            return ANSB200_ECUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define ANS_REQUIRE(cond, ...)                                                                    \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            ansb200_set_error(__VA_ARGS__);                                                       \
            return ANSB200_EINVAL;                                                                \
        }                                                                                         \
    } while (0)

#define ANS_LAUNCH_CHECK()                                                                        \
    do {                                                                                          \
        cudaError_t err__ = cudaGetLastError();                                                   \
        if (err__ != cudaSuccess) {                                                               \
            ansb200_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                              cudaGetErrorString(err__));                                         \
            return ANSB200_ECUDA;                                                                 \
        }                                                                                         \
    } while (0)

static inline int ans_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
