// kdist.cu -- k-distributions from a monochromatic absorption spectrum (generation of k-tables).
//
// Reference: the per-bin tail of calc_ktable_chunk (archnemesis/Spectroscopy_0.py:3619-3660).  For every spectral bin
// of the k-table the reference masks the line-by-line grid (`(wavecalc >= vbinmin) & (wavecalc <= vbinmax)`, a pass
// over the whole grid per bin), argsorts the absorption coefficients inside the bin, forms the cumulative
// distribution g_i = cumsum(ils_i * delv) / sum(ils * delv) in that order (ils = 1 without an instrument function)
// and reads the k-coefficients off at the g-ordinates with np.interp(G_ORD, g_sorted, k_sorted).
//
// Here: the host turns the masks into index ranges [lo, hi) of the (ascending) grid; one CTA per bin loads the
// bin's coefficients (and weights) into shared memory, sorts them with a bitonic network (keys padded with +inf to
// a power of two; the weight travels with its key), scans the weights, and NG threads do np.interp's search and
// slope form.  Memory traffic is the bin's points once; the sort is n log^2 n / 2 compare-exchanges in shared memory.
// Bins beyond the shared-memory capacity (KD_MAX_N / KD_MAX_NW points) are refused (ANSB200_EINVAL): the host sorts
// those with the library's radix sort.
//
// Without weights g_i = (i+1)/n exactly; the reference accumulates delv n times (np.cumsum) and divides by a
// pairwise sum, which differs from that by ~n ulp -- far below the 1e-9 bar, and documented in the tests.
#include <float.h>
#include <math.h>
#include "common.cuh"

constexpr int KD_THREADS = 1024;
constexpr int KD_MAX_N = 16384;       // points per bin, unweighted (128 KB of keys)
constexpr int KD_MAX_NW = 8192;       // points per bin, weighted (keys + weights)

template <bool WEIGHTED>
__global__ void __launch_bounds__(KD_THREADS)
ans_kdist_kernel(const double *__restrict__ kabs, const double *__restrict__ w, const int32_t *__restrict__ lo,
                 const int32_t *__restrict__ hi, const int64_t *__restrict__ woff, const double *__restrict__ g_ord,
                 int NG, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char kd_smem[];
    __shared__ double s_part[KD_THREADS / 32];
    __shared__ double s_total;
    const int b = blockIdx.x;
    const int i0 = lo[b], n = hi[b] - i0;
    int NP2 = 32;
    while (NP2 < n) NP2 <<= 1;
    double *key = reinterpret_cast<double *>(kd_smem);
    double *wt = key + NP2;                                  // (weighted only)
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < NP2; i += nthr) {
        key[i] = i < n ? kabs[i0 + i] : INFINITY;
        if (WEIGHTED) wt[i] = i < n ? w[woff[b] + i] : 0.0;
    }
    __syncthreads();
    // bitonic sort, ascending
    for (int k = 2; k <= NP2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (NP2 >> 1); t += nthr) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // lower index of the pair
                const int p = i | j;
                const bool up = (i & k) == 0;
                const double a = key[i], c = key[p];
                if ((a > c) == up) {
                    key[i] = c;
                    key[p] = a;
                    if (WEIGHTED) {
                        const double wa = wt[i];
                        wt[i] = wt[p];
                        wt[p] = wa;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (WEIGHTED) {
        // inclusive scan of the weights in sorted order: contiguous chunk per thread, warp scan, scan of the warp sums
        const int per = (NP2 + nthr - 1) / nthr;
        const int a0 = min(tid * per, NP2), a1 = min(a0 + per, NP2);
        double s = 0.0;
        for (int i = a0; i < a1; ++i) s += wt[i];
        double incl = s;
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) s_part[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            double v = lane < (nthr >> 5) ? s_part[lane] : 0.0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += up;
            }
            if (lane < (nthr >> 5)) s_part[lane] = v;               // inclusive sums of the warps
            if (lane == 31) s_total = v;
        }
        __syncthreads();
        double run = (incl - s) + (wid > 0 ? s_part[wid - 1] : 0.0);       // exclusive prefix of this thread's chunk
        for (int i = a0; i < a1; ++i) {
            run += wt[i];
            wt[i] = run;
        }
        __syncthreads();
    }
    // np.interp(G_ORD, g_sorted, k_sorted): g_i = c_i / total (weighted) or (i + 1) / n
    for (int ig = tid; ig < NG; ig += nthr) {
        const double g = g_ord[ig];
        const double tot = WEIGHTED ? s_total : (double)n;
        auto gval = [&](int i) { return WEIGHTED ? wt[i] / tot : (double)(i + 1) / tot; };
        double r;
        if (g <= gval(0)) {
            r = key[0];
        } else if (g >= gval(n - 1)) {
            r = key[n - 1];
        } else {
            int a = 0, c = n - 1;                            // gval(a) <= g < gval(c)
            while (c - a > 1) {
                const int m = (a + c) >> 1;
                if (gval(m) <= g) a = m; else c = m;
            }
            const double xa = gval(a), xc = gval(a + 1);
            if (g == xa) {
                r = key[a];
            } else {
                const double slope = (key[a + 1] - key[a]) / (xc - xa);
                r = slope * (g - xa) + key[a];
            }
        }
        out[(size_t)b * NG + ig] = r;
    }
}

extern "C" int ansb200_kdist_capacity(int weighted) { return weighted ? KD_MAX_NW : KD_MAX_N; }

extern "C" int ansb200_kdist(const double *kabs, const double *w, const int32_t *lo, const int32_t *hi,
                             const int64_t *woff, int NBIN, int max_n, const double *g_ord, int NG, double *out,
                             void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(kabs && lo && hi && g_ord && out, "kdist: null pointer");
    ANS_REQUIRE(NBIN > 0 && NG > 0 && max_n > 0, "kdist: bad shape");
    ANS_REQUIRE(!w || woff, "kdist: weights need their per-bin offsets");
    const int cap = w ? KD_MAX_NW : KD_MAX_N;
    ANS_REQUIRE(max_n <= cap, "kdist: a bin of %d points exceeds the shared-memory sort (%d)", max_n, cap);
    int NP2 = 32;
    while (NP2 < max_n) NP2 <<= 1;
    const size_t smem = (size_t)NP2 * 8 * (w ? 2 : 1);
    if (w) {
        ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_kdist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        ans_kdist_kernel<true><<<(unsigned)NBIN, KD_THREADS, smem, stream>>>(kabs, w, lo, hi, woff, g_ord, NG, out);
    } else {
        ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_kdist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        ans_kdist_kernel<false><<<(unsigned)NBIN, KD_THREADS, smem, stream>>>(kabs, w, lo, hi, woff, g_ord, NG, out);
    }
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
