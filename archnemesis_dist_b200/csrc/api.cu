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

// Table upload: the reference layout K[NWAVE,NG,NP,NT,NGAS] is re-ordered to plane-major
// [NP*NT][NWAVE][NG][NGAS] and ln K is tabulated beside it: positive entries get log(K); zero -> -inf;
// negative -> NaN so that one non-finite corner routes the interpolation to the reference's linear /
// zero branches.  One thread per output element (coalesced writes, strided one-off reads).
__global__ void ans_table_log_kernel(const double *__restrict__ Kin, double *__restrict__ K, double *__restrict__ lnK,
                                     int NWAVE, int NG, int NPT, int NGAS)
{
    const size_t n = (size_t)NWAVE * NG * NPT * NGAS;
    const size_t pairs = (size_t)NWAVE * NG;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const size_t gas = i % NGAS;
        const size_t rest = i / NGAS;
        const size_t pair = rest % pairs;       // wave*NG + g
        const size_t pt = rest / pairs;         // ip*NT + it
        const double v = Kin[(pair * NPT + pt) * NGAS + gas];
        K[i] = v;
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
    double *staging = nullptr;
    if (cudaMalloc(&t->K, n * sizeof(double)) != cudaSuccess || cudaMalloc(&t->lnK, n * sizeof(double)) != cudaSuccess ||
        (!is_device && cudaMalloc(&staging, n * sizeof(double)) != cudaSuccess)) {
        cudaGetLastError();
        if (t->K) cudaFree(t->K);
        if (t->lnK) cudaFree(t->lnK);
        delete t;
        ansb200_set_error("table_create: cudaMalloc of %d x %zu bytes failed", is_device ? 2 : 3, n * sizeof(double));
        return ANSB200_ENOMEM;
    }
    const double *src = K;
    cudaError_t e = cudaSuccess;
    if (!is_device) {
        e = cudaMemcpyAsync(staging, K, n * sizeof(double), cudaMemcpyHostToDevice, stream);
        src = staging;
    }
    if (e == cudaSuccess) {
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        ans_table_log_kernel<<<blocks, 256, 0, stream>>>(src, t->K, t->lnK, NWAVE, NG, NP * NT, NGAS);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && staging) {
        e = cudaStreamSynchronize(stream);   // one-off upload: wait, then release the staging copy
        cudaFree(staging);
        staging = nullptr;
    }
    if (e != cudaSuccess) {
        cudaFree(t->K); cudaFree(t->lnK); if (staging) cudaFree(staging); delete t;
        ansb200_set_error("table_create: upload failed: %s", cudaGetErrorString(e));
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
