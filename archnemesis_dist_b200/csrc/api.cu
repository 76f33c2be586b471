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
template <typename TK, typename TL>
__global__ void ans_table_log_kernel(const double *__restrict__ Kin, TK *__restrict__ K, TL *__restrict__ lnK,
                                     int NWAVE, int NG, int NPT, int NGAS, unsigned long long *__restrict__ inexact)
{
    const size_t n = (size_t)NWAVE * NG * NPT * NGAS;
    const size_t pairs = (size_t)NWAVE * NG;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (; i < n; i += stride) {
        const size_t gas = i % NGAS;
        const size_t rest = i / NGAS;
        const size_t pair = rest % pairs;       // wave*NG + g
        const size_t pt = rest / pairs;         // ip*NT + it
        const double v = Kin[(pair * NPT + pt) * NGAS + gas];
        K[i] = (TK)v;
        if ((double)(TK)v != v && v == v) ++bad;       // a value float32 cannot hold: the variant would not be lossless
        lnK[i] = (TL)(v > 0.0 ? log(v) : (v == 0.0 ? -INFINITY : NAN));
    }
    if (bad && inexact) atomicAdd(inexact, bad);
}

// Storage variants: F64 keeps K and ln K as float64 (2 x 8 bytes per entry); K32 keeps K as float32 -- the values of a
// .kta / .lta file ARE float32 (Spectroscopy_0.py:2849), so this is lossless and is refused otherwise -- and ln K as
// float64 (12 bytes); F32 keeps both as float32 (8 bytes): the "FP32 k-interp variant" of BASELINE.json, lossy in
// ln K (~6e-8 relative, i.e. ~4e-6 relative in k at ln k ~ -50), reported separately.
extern "C" int ansb200_table_create_ex(const double *K, int is_device, int NWAVE, int NG, int NP, int NT, int NGAS,
                                       int storage, ansb200_table **out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(K && out, "table_create: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NG > 0 && NP >= 2 && NT >= 2 && NGAS > 0, "table_create: bad shape");
    ANS_REQUIRE(NG <= ANSB200_MAX_NG, "table_create: NG=%d exceeds %d", NG, ANSB200_MAX_NG);
    ANS_REQUIRE(storage >= ANSB200_TABLE_F64 && storage <= ANSB200_TABLE_F32, "table_create: unknown storage %d", storage);
    size_t n = (size_t)NWAVE * NG * NP * NT * NGAS;
    ansb200_table *t = new ansb200_table();
    t->NWAVE = NWAVE; t->NG = NG; t->NP = NP; t->NT = NT; t->NGAS = NGAS; t->storage = storage;
    t->K = nullptr; t->lnK = nullptr; t->Kf = nullptr; t->lnKf = nullptr;
    double *staging = nullptr;
    unsigned long long *inexact = nullptr;
    const size_t kbytes = n * (storage == ANSB200_TABLE_F64 ? 8 : 4), lbytes = n * (storage == ANSB200_TABLE_F32 ? 4 : 8);
    void *kp = nullptr, *lp = nullptr;
    if (cudaMalloc(&kp, kbytes) != cudaSuccess || cudaMalloc(&lp, lbytes) != cudaSuccess ||
        cudaMalloc((void **)&inexact, sizeof(unsigned long long)) != cudaSuccess ||
        (!is_device && cudaMalloc(&staging, n * sizeof(double)) != cudaSuccess)) {
        cudaGetLastError();
        if (kp) cudaFree(kp);
        if (lp) cudaFree(lp);
        if (inexact) cudaFree(inexact);
        delete t;
        ansb200_set_error("table_create: cudaMalloc of %zu bytes failed", kbytes + lbytes + (is_device ? 0 : n * 8));
        return ANSB200_ENOMEM;
    }
    if (storage == ANSB200_TABLE_F64) t->K = (double *)kp; else t->Kf = (float *)kp;
    if (storage == ANSB200_TABLE_F32) t->lnKf = (float *)lp; else t->lnK = (double *)lp;
    const double *src = K;
    cudaError_t e = cudaMemsetAsync(inexact, 0, sizeof(unsigned long long), stream);
    if (e == cudaSuccess && !is_device) {
        e = cudaMemcpyAsync(staging, K, n * sizeof(double), cudaMemcpyHostToDevice, stream);
        src = staging;
    }
    if (e == cudaSuccess) {
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        const int NPT = NP * NT;
        if (storage == ANSB200_TABLE_F64)
            ans_table_log_kernel<double, double><<<blocks, 256, 0, stream>>>(src, t->K, t->lnK, NWAVE, NG, NPT, NGAS, nullptr);
        else if (storage == ANSB200_TABLE_K32)
            ans_table_log_kernel<float, double><<<blocks, 256, 0, stream>>>(src, t->Kf, t->lnK, NWAVE, NG, NPT, NGAS, inexact);
        else
            ans_table_log_kernel<float, float><<<blocks, 256, 0, stream>>>(src, t->Kf, t->lnKf, NWAVE, NG, NPT, NGAS, inexact);
        e = cudaGetLastError();
    }
    unsigned long long nbad = 0;
    if (e == cudaSuccess && (staging || storage != ANSB200_TABLE_F64)) {
        if (storage != ANSB200_TABLE_F64)
            e = cudaMemcpyAsync(&nbad, inexact, sizeof(nbad), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // one-off upload: wait, then release the staging copy
    }
    if (staging) cudaFree(staging);
    cudaFree(inexact);
    if (e != cudaSuccess || (storage == ANSB200_TABLE_K32 && nbad)) {
        cudaFree(kp); cudaFree(lp); delete t;
        if (e != cudaSuccess) {
            ansb200_set_error("table_create: upload failed: %s", cudaGetErrorString(e));
            return ANSB200_ECUDA;
        }
        ansb200_set_error("table_create: %llu table values are not float32 numbers: the K32 storage would not be lossless", nbad);
        return ANSB200_EINVAL;
    }
    *out = t;
    return ANSB200_OK;
}

extern "C" int ansb200_table_create(const double *K, int is_device, int NWAVE, int NG, int NP, int NT, int NGAS,
                                    ansb200_table **out, void *stream_)
{
    return ansb200_table_create_ex(K, is_device, NWAVE, NG, NP, NT, NGAS, ANSB200_TABLE_F64, out, stream_);
}

extern "C" int ansb200_table_destroy(ansb200_table *t)
{
    if (!t) return ANSB200_OK;
    cudaFree(t->K);
    cudaFree(t->lnK);
    cudaFree(t->Kf);
    cudaFree(t->lnKf);
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
