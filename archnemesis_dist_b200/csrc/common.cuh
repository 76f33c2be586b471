// common.cuh -- shared helpers for libansb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ansb200.h"

struct ansb200_table {
    double *K;    // resident copy, plane-major [NP*NT][NWAVE][NG][NGAS]
    double *lnK;  // same layout: log(K), -inf for K==0, NaN for K<0
    int NWAVE, NG, NP, NT, NGAS;
};

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
