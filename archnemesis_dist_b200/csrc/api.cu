// api.cu -- error plumbing and the table handle of libansb200.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void ansb200_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *ansb200_last_error(void) { return g_err; }
extern "C" int ansb200_version(void) { return 100; }

// ln K tabulated once per table: positive entries get log(K); zero -> -inf; negative -> NaN so
// that one non-finite corner routes the interpolation to the reference's linear / zero branches.
__global__ void ans_table_log_kernel(const double *__restrict__ K, double *__restrict__ lnK, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double v = K[i];
        lnK[i] = v > 0.0 ? log(v) : (v == 0.0 ? -INFINITY : NAN);
    }
}

extern "C" int ansb200_table_create(const double *K, int is_device, int NWAVE, int NG, int NP, int NT, int NGAS,
                                    ansb200_table **out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(K && out, "table_create: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NG > 0 && NP >= 2 && NT >= 2 && NGAS > 0, "table_create: bad shape");
    ANS_REQUIRE(NG <= ANSB200_MAX_NG, "table_create: NG=%d exceeds %d", NG, ANSB200_MAX_NG);
    size_t n = (size_t)NWAVE * NG * NP * NT * NGAS;
    ansb200_table *t = new ansb200_table();
    t->NWAVE = NWAVE; t->NG = NG; t->NP = NP; t->NT = NT; t->NGAS = NGAS;
    t->K = nullptr; t->lnK = nullptr;
    if (cudaMalloc(&t->K, n * sizeof(double)) != cudaSuccess || cudaMalloc(&t->lnK, n * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        if (t->K) cudaFree(t->K);
        delete t;
        ansb200_set_error("table_create: cudaMalloc of 2 x %zu bytes failed", n * sizeof(double));
        return ANSB200_ENOMEM;
    }
    cudaError_t e = cudaMemcpyAsync(t->K, K, n * sizeof(double),
                                    is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
        cudaFree(t->K); cudaFree(t->lnK); delete t;
        ansb200_set_error("table_create: copy failed: %s", cudaGetErrorString(e));
        return ANSB200_ECUDA;
    }
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ans_table_log_kernel<<<blocks, 256, 0, stream>>>(t->K, t->lnK, n);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(t->K); cudaFree(t->lnK); delete t;
        ansb200_set_error("table_create: log kernel launch failed: %s", cudaGetErrorString(e));
        return ANSB200_ECUDA;
    }
    *out = t;
    return ANSB200_OK;
}

extern "C" int ansb200_table_destroy(ansb200_table *t)
{
    if (!t) return ANSB200_OK;
    cudaFree(t->K);
    cudaFree(t->lnK);
    delete t;
    return ANSB200_OK;
}

extern "C" int ansb200_table_shape(const ansb200_table *t, int *NWAVE, int *NG, int *NP, int *NT, int *NGAS)
{
    ANS_REQUIRE(t, "table_shape: null handle");
    if (NWAVE) *NWAVE = t->NWAVE;
    if (NG) *NG = t->NG;
    if (NP) *NP = t->NP;
    if (NT) *NT = t->NT;
    if (NGAS) *NGAS = t->NGAS;
    return ANSB200_OK;
}

extern "C" const double *ansb200_table_k(const ansb200_table *t) { return t ? t->K : nullptr; }
extern "C" const double *ansb200_table_lnk(const ansb200_table *t) { return t ? t->lnK : nullptr; }
