// project.cu -- layer-space gradients -> state vector (map2pro + map2xvec).
//
// Reference: archnemesis/ForwardModel_0.py:5319-5383 (map2pro) and :5387-5424 (map2xvec).  Both are
// linear, so the host folds D[LAYINC[j,path],:] (DAM / DTE / DCO per parameter, with the
// reference's parameter-inclusion rule :699-702) and xmap into one matrix per path,
//     M[path, k*NLAYMAX + j, x] = sum_pro D_k[LAYINC[j,path], pro] * xmap[x, k, pro],
// and this kernel is the skinny FP64 product out[w,path,:] = dspec[w,path,:] . M[path].
// FP64 FMA pipe; nothing here is shaped for tensor cores (0.5 GFLOP at 4000 x 1000 x 60).
#include "common.cuh"

constexpr int PJ_BM = 32, PJ_BN = 64, PJ_BK = 16;

__global__ void __launch_bounds__(256)
ans_project_kernel(const double *__restrict__ dspec, const double *__restrict__ M, int NWAVE, int E, int NPATH, int NX,
                   double *__restrict__ out)
{
    __shared__ double sA[PJ_BK][PJ_BM + 1];
    __shared__ double sBm[PJ_BK][PJ_BN];
    const int ipath = blockIdx.z;
    const int w0 = blockIdx.x * PJ_BM, x0 = blockIdx.y * PJ_BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16
    const double *Mp = M + (size_t)ipath * E * NX;
    double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int e0 = 0; e0 < E; e0 += PJ_BK) {
        for (int t = threadIdx.x; t < PJ_BM * PJ_BK; t += 256) {
            const int r = t / PJ_BK, c = t - r * PJ_BK;
            const int w = w0 + r, e = e0 + c;
            sA[c][r] = (w < NWAVE && e < E) ? dspec[((size_t)w * NPATH + ipath) * E + e] : 0.0;
        }
        for (int t = threadIdx.x; t < PJ_BK * PJ_BN; t += 256) {
            const int r = t / PJ_BN, c = t - r * PJ_BN;
            const int e = e0 + r, x = x0 + c;
            sBm[r][c] = (e < E && x < NX) ? Mp[(size_t)e * NX + x] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PJ_BK; ++kk) {
            const double a0 = sA[kk][ty * 2], a1 = sA[kk][ty * 2 + 1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double b = sBm[kk][tx + 16 * c];
                acc[0][c] += a0 * b;
                acc[1][c] += a1 * b;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int w = w0 + ty * 2 + r;
        if (w >= NWAVE) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int x = x0 + tx + 16 * c;
            if (x < NX) out[((size_t)w * NPATH + ipath) * NX + x] = acc[r][c];
        }
    }
}

extern "C" int ansb200_jacobian_project(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAYMAX,
                                        int NPATH, int NX, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(dspec && M && out, "jacobian_project: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NPAR > 0 && NLAYMAX > 0 && NPATH > 0 && NX > 0, "jacobian_project: bad shape");
    ANS_REQUIRE(NPATH <= 65535, "jacobian_project: NPATH too large");
    dim3 grid(ans_div_up(NWAVE, PJ_BM), ans_div_up(NX, PJ_BN), NPATH);
    ans_project_kernel<<<grid, 256, 0, stream>>>(dspec, M, NWAVE, NPAR * NLAYMAX, NPATH, NX, out);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
