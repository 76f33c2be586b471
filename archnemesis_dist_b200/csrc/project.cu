// project.cu -- layer-space gradients -> state vector (map2pro + map2xvec).
//
// Reference: archnemesis/ForwardModel_0.py:5319-5383 (map2pro) and :5387-5424 (map2xvec).  Both are
// linear, so the host folds D[LAYINC[j,path],:] (DAM / DTE / DCO per parameter, with the
// reference's parameter-inclusion rule :699-702) and xmap into one matrix per path,
//     M[path, k*NLAYMAX + j, x] = sum_pro D_k[LAYINC[j,path], pro] * xmap[x, k, pro],
// and this kernel is the skinny FP64 product out[w,path,:] = dspec[w,path,:] . M[path].
// FP64 FMA pipe (0.5 GFLOP at 4000 x 1000 x 60, 61 GFLOP for 64 limb paths of 200 layers, before the zero blocks of M are
// skipped); when M has its usual structure the sparse kernel further down is used instead.
#include "common.cuh"

// Tiling: a CTA owns 128 wavenumbers x 64 state-vector columns of one path and walks E = NPAR*NLAYMAX in
// chunks of 16, double-buffered in shared memory.  Warp w owns the 8 columns [8w, 8w+8); lane l owns rows
// 4l..4l+3, so a thread keeps a 4 x 8 micro-tile (32 FP64 accumulators) and per k reads its four A values
// (two 128-bit loads, conflict-free) and the warp's eight B values (broadcast).  M is block-sparse -- a
// state-vector element maps to the layers of ONE parameter (temperature, one gas, ...), so in a chunk of 16
// rows of M most column groups are all zero: each warp tests its 16 x 8 block of the staged B tile
// (ballot, warp-uniform) and skips the 512 FMAs per thread when it is empty.  At config 2 that leaves
// ~1/NPAR of the dense flops; the dense rate is FP64-FMA-pipe bound.
constexpr int PJ_BM = 128, PJ_BN = 64, PJ_BK = 16, PJ_THREADS = 256;

__global__ void __launch_bounds__(PJ_THREADS, 2)
ans_project_kernel(const double *__restrict__ dspec, const double *__restrict__ M, int NWAVE, int E, int NPATH, int NX,
                   double *__restrict__ out, size_t mstride, const int32_t *__restrict__ chunks, int nchunks, int listed)
{
    __shared__ __align__(16) double sA[2][PJ_BK][PJ_BM];      // [k][row]
    __shared__ __align__(16) double sB[2][PJ_BK][PJ_BN];      // [k][col]
    const int ipath = blockIdx.z;
    const int w0 = blockIdx.x * PJ_BM, x0 = blockIdx.y * PJ_BN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *Mp = M + (size_t)ipath * mstride;       // (0: one layer-space matrix for every path)
    double acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

    // staging: A tile = 128 rows x 16 k (thread t: row t/2, 8 consecutive k -> one 64-byte run of dspec);
    // B tile = 16 k x 64 columns (thread t: k = t/16, 4 consecutive columns)
    const int ar = threadIdx.x >> 1, ak = (threadIdx.x & 1) * 8;
    const int bk = threadIdx.x >> 4, bc = (threadIdx.x & 15) * 4;
    const bool arow_ok = w0 + ar < NWAVE;
    const double *arow = dspec + ((size_t)(arow_ok ? w0 + ar : 0) * NPATH + ipath) * E;
    double ra[8], rb[4];
    auto fetch = [&](int e0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = e0 + ak + q;
            ra[q] = (arow_ok && e < E) ? __ldg(arow + e) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = e0 + bk, x = x0 + bc + q;
            rb[q] = (e < E && x < NX) ? __ldg(Mp + (size_t)e * NX + x) : 0.0;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sA[buf][ak + q][ar] = ra[q];
#pragma unroll
        for (int q = 0; q < 4; ++q) sB[buf][bk][bc + q] = rb[q];
    };
    // chunks (optional): the 16-row chunks of M that hold a non-zero for ANY path / column, found once on the host --
    // whole parameters without a state-vector element (a third of the rows at config 4) are then never read
    const int nch = listed ? nchunks : (E + PJ_BK - 1) / PJ_BK;
    if (nch == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int w = w0 + lane * 4 + r;
            if (w >= NWAVE) continue;
            double *o = out + ((size_t)w * NPATH + ipath) * NX + x0 + warp * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (x0 + warp * 8 + c < NX) o[c] = 0.0;
        }
        return;
    }
    fetch(listed ? chunks[0] * PJ_BK : 0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int ci = 0; ci < nch; ++ci) {
        const bool more = ci + 1 < nch;
        if (more) fetch(listed ? chunks[ci + 1] * PJ_BK : (ci + 1) * PJ_BK);   // global loads of the next chunk fly during the FMAs
        // is this warp's 16 x 8 block of M empty?  (lane -> k = lane/2, 4 of the 8 columns)
        const double *bq = &sB[buf][lane >> 1][warp * 8 + (lane & 1) * 4];
        const bool nz = bq[0] != 0.0 || bq[1] != 0.0 || bq[2] != 0.0 || bq[3] != 0.0;
        if (__any_sync(0xffffffffu, nz)) {
#pragma unroll
            for (int kk = 0; kk < PJ_BK; ++kk) {
                const double2 a01 = *reinterpret_cast<const double2 *>(&sA[buf][kk][lane * 4]);
                const double2 a23 = *reinterpret_cast<const double2 *>(&sA[buf][kk][lane * 4 + 2]);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                double bv[8];
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const double2 b2 = *reinterpret_cast<const double2 *>(&sB[buf][kk][warp * 8 + c]);
                    bv[c] = b2.x;
                    bv[c + 1] = b2.y;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
            }
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int w = w0 + lane * 4 + r;
        if (w >= NWAVE) continue;
        double *o = out + ((size_t)w * NPATH + ipath) * NX + x0 + warp * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (x0 + warp * 8 + c < NX) o[c] = acc[r][c];
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
    ans_project_kernel<<<grid, PJ_THREADS, 0, stream>>>(dspec, M, NWAVE, NPAR * NLAYMAX, NPATH, NX, out,
                                                        (size_t)NPAR * NLAYMAX * NX, nullptr, 0, 0);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

// Both products with the list of non-empty 16-row chunks of M (ascending chunk numbers e0 / 16, on the device): rows of
// M outside the listed chunks must be zero for every path and column.  shared != 0: one matrix for all paths.
extern "C" int ansb200_jacobian_project_chunks(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAYMAX,
                                               int NPATH, int NX, int shared, const int32_t *chunks, int nchunks,
                                               double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(dspec && M && out, "jacobian_project: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NPAR > 0 && NLAYMAX > 0 && NPATH > 0 && NX > 0, "jacobian_project: bad shape");
    ANS_REQUIRE(NPATH <= 65535, "jacobian_project: NPATH too large");
    ANS_REQUIRE(nchunks >= 0 && (nchunks == 0 || chunks), "jacobian_project: chunk list missing");
    ANS_REQUIRE(nchunks <= ans_div_up((long long)NPAR * NLAYMAX, PJ_BK), "jacobian_project: more chunks than M has");
    dim3 grid(ans_div_up(NWAVE, PJ_BM), ans_div_up(NX, PJ_BN), NPATH);
    ans_project_kernel<<<grid, PJ_THREADS, 0, stream>>>(dspec, M, NWAVE, NPAR * NLAYMAX, NPATH, NX, out,
                                                        shared ? (size_t)0 : (size_t)NPAR * NLAYMAX * NX,
                                                        chunks, nchunks, 1);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

// The same product with ONE matrix M[NPAR*NLAY, NX] for all paths: layer-space gradients (ANSB200_RAD_LAYER_SPACE), where
// the layer -> profile -> state map does not depend on the path.
extern "C" int ansb200_jacobian_project_shared(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAY,
                                               int NPATH, int NX, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(dspec && M && out, "jacobian_project: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NPAR > 0 && NLAY > 0 && NPATH > 0 && NX > 0, "jacobian_project: bad shape");
    ANS_REQUIRE(NPATH <= 65535, "jacobian_project: NPATH too large");
    dim3 grid(ans_div_up(NWAVE, PJ_BM), ans_div_up(NX, PJ_BN), NPATH);
    ans_project_kernel<<<grid, PJ_THREADS, 0, stream>>>(dspec, M, NWAVE, NPAR * NLAY, NPATH, NX, out, (size_t)0, nullptr, 0, 0);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

// ---- the same product with M as a sparse operator ------------------------------------------------------------------
// M is what map2pro + map2xvec fold into: a state-vector element acts on the layers of ONE parameter, and a temperature
// level on the two or three layers next to it, so M holds a few per cent non-zeros (1.7 % at config 4, 1000 of 60 000)
// and every column is one short run of consecutive rows.  The dense kernel above then spends its time streaming tiles it
// mostly skips.  Here the host hands M over by columns (plan.sparse_projection: for column x of path p the run starts at
// row r0 and holds len values at voff in vals -- zeros inside the run are kept, so there is no index array) and a WARP
// takes one (wavenumber, path) row of dspec: it stages the row in shared memory with coalesced 16-byte loads -- the only
// traffic that matters, each byte of dspec once -- and then every lane sums one column (ascending rows, fused
// multiply-adds, exactly like the dense kernel: bit-identical to it); columns longer than PS_LONG (a gas scaling factor
// touches every layer) are taken together, a few lanes each.  When one operator serves every row (one path, or layer
// space) the CTA keeps it in shared memory as well.
constexpr int PS_WARPS = 8, PS_LONG = 16;
constexpr int PS_PLAN_BYTES = 24 * 1024;          // operator kept in shared memory up to this size

__global__ void __launch_bounds__(PS_WARPS * 32)
ans_project_sparse_kernel(const double *__restrict__ dspec, const int32_t *__restrict__ col_r0,
                          const int32_t *__restrict__ col_len, const int32_t *__restrict__ col_voff,
                          const double *__restrict__ vals, const int32_t *__restrict__ long_cols,
                          const int32_t *__restrict__ long_ptr, long long NROW, int E, int NPATH, int NX, int shared,
                          int nvals, int plan_in_smem, double *__restrict__ out)
{
    extern __shared__ __align__(16) double ps_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ES = (E + 1) & ~1;
    double *a = ps_smem + (size_t)warp * ES;
    // the operator (one for all rows) in shared memory: vals, then r0 / len / voff
    const double *pv = vals;
    const int32_t *pr0 = col_r0, *plen = col_len, *pvo = col_voff;
    if (plan_in_smem) {
        double *sv = ps_smem + (size_t)PS_WARPS * ES;
        int32_t *si = reinterpret_cast<int32_t *>(sv + nvals);
        for (int t = threadIdx.x; t < nvals; t += blockDim.x) sv[t] = vals[t];
        for (int t = threadIdx.x; t < NX; t += blockDim.x) {
            si[t] = col_r0[t];
            si[NX + t] = col_len[t];
            si[2 * NX + t] = col_voff[t];
        }
        __syncthreads();
        pv = sv; pr0 = si; plen = si + NX; pvo = si + 2 * NX;
    }
    const bool wide = (E & 1) == 0;
    for (long long row = (long long)blockIdx.x * PS_WARPS + warp; row < NROW; row += (long long)gridDim.x * PS_WARPS) {
        const int ipath = shared ? 0 : (int)(row % NPATH);
        const double *src = dspec + (size_t)row * E;
        if (wide) {
            const double2 *src2 = reinterpret_cast<const double2 *>(src);
            double2 *a2 = reinterpret_cast<double2 *>(a);
            const int n2 = E >> 1;
            for (int e0 = 0; e0 < n2; e0 += 8 * 32) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + u * 32 + lane;
                    if (e < n2) v[u] = __ldcs(src2 + e);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + u * 32 + lane;
                    if (e < n2) a2[e] = v[u];
                }
            }
        } else {
            for (int e = lane; e < E; e += 32) a[e] = __ldg(src + e);
        }
        __syncwarp();
        const int cb = ipath * NX;
        double *o = out + (size_t)row * NX;
        for (int x0 = 0; x0 < NX; x0 += 32) {
            const int x = x0 + lane;
            int len = 0, r0 = 0, vo = 0;
            if (x < NX) { len = plen[cb + x]; r0 = pr0[cb + x]; vo = pvo[cb + x]; }
            if (len <= PS_LONG) {
                double acc = 0.0;
                for (int i = 0; i < len; ++i) acc = fma(a[r0 + i], pv[vo + i], acc);
                if (x < NX) o[x] = acc;
            }
        }
        // long columns, all at once: S = 32 / (their number, rounded up to a power of two) lanes share a column, each sums
        // a contiguous slice of it in four running sums, and log2(S) shuffle steps add the slices
        const int lq0 = long_ptr[ipath], nlong = long_ptr[ipath + 1] - lq0;
        for (int q0 = 0; q0 < nlong; q0 += 32) {
            const int nq = min(32, nlong - q0);
            int S = 32;
            while (S > 1 && S * nq > 32) S >>= 1;                  // lanes per column
            const int qi = lane / S, sl = lane - qi * S;
            double acc = 0.0;
            int x = -1;
            if (qi < nq) {
                x = long_cols[lq0 + q0 + qi];
                const int len = plen[cb + x], per = (len + S - 1) / S;
                const int ia = min(len, sl * per), ib = min(len, ia + per);
                const double *ap = a + pr0[cb + x], *vp = pv + pvo[cb + x];
                double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
                int i = ia;
                for (; i + 4 <= ib; i += 4) {
                    c0 = fma(ap[i], vp[i], c0);
                    c1 = fma(ap[i + 1], vp[i + 1], c1);
                    c2 = fma(ap[i + 2], vp[i + 2], c2);
                    c3 = fma(ap[i + 3], vp[i + 3], c3);
                }
                for (; i < ib; ++i) c0 = fma(ap[i], vp[i], c0);
                acc = (c0 + c1) + (c2 + c3);
            }
            for (int d = S >> 1; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            if (x >= 0 && sl == 0) o[x] = acc;
        }
        __syncwarp();
    }
}

extern "C" int ansb200_jacobian_project_sparse(const double *dspec, const int32_t *col_r0, const int32_t *col_len,
                                               const int32_t *col_voff, const double *vals, int nvals,
                                               const int32_t *long_cols, const int32_t *long_ptr, int NWAVE, int NPAR,
                                               int NLAYMAX, int NPATH, int NX, int shared, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(dspec && col_r0 && col_len && col_voff && long_ptr && out, "jacobian_project_sparse: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NPAR > 0 && NLAYMAX > 0 && NPATH > 0 && NX > 0 && nvals >= 0, "jacobian_project_sparse: bad shape");
    const int E = NPAR * NLAYMAX, ES = (E + 1) & ~1;
    size_t smem = (size_t)PS_WARPS * ES * 8;
    ANS_REQUIRE(smem <= 200 * 1024, "jacobian_project_sparse: NPAR*NLAYMAX = %d too long for the shared-memory row", E);
    const size_t plan_bytes = (size_t)nvals * 8 + (size_t)3 * NX * 4;
    const int plan_in_smem = (shared || NPATH == 1) && plan_bytes <= PS_PLAN_BYTES && smem + plan_bytes <= 220 * 1024;
    if (plan_in_smem) smem += plan_bytes;
    ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_project_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    const long long NROW = (long long)NWAVE * NPATH;
    long long nblk = (NROW + PS_WARPS - 1) / PS_WARPS;
    const long long cap = 148LL * 8;            // persistent over the rows beyond a few CTAs per SM
    if (nblk > cap) nblk = cap;
    ans_project_sparse_kernel<<<(unsigned)nblk, PS_WARPS * 32, smem, stream>>>(dspec, col_r0, col_len, col_voff, vals,
                                                                              long_cols, long_ptr, NROW, E, NPATH, NX,
                                                                              shared, nvals, plan_in_smem, out);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

// ---- tangent-height interpolation of the path spectra (nemesisSOfmg / nemesisLfmg) ---------------------------------
// Reference: ForwardModel_0.py:1206-1228 / :1464-1486.  The limb / occultation drivers compute one spectrum per path
// (tangent layer) and interpolate the pair of paths that brackets each measured tangent height:
//     SPECMOD[:, i] = SPECOUT[:, lo_i] * wlo_i + SPECOUT[:, hi_i] * whi_i        (hi_i < 0: SPECOUT[:, lo_i] alone)
// and the same for the Jacobian.  out[NWAVE, NGEOM, 1 + NX] = [SPECMOD | dSPECMOD]: the block the line-shape operator
// takes next (ansb200_convolve treats NGEOM * (1 + NX) as columns).  Products and sum are rounded separately, like numpy.
__global__ void __launch_bounds__(256)
ans_path_mix_kernel(const double *__restrict__ spec, const double *__restrict__ dx, const int32_t *__restrict__ lo,
                    const int32_t *__restrict__ hi, const double *__restrict__ wlo, const double *__restrict__ whi,
                    int NPATH, int NX, int NGEOM, double *__restrict__ out)
{
    const int iw = blockIdx.x, ig = blockIdx.y;
    const int pl = lo[ig], ph = hi[ig];
    const double a = wlo[ig], b = whi[ig];
    const double *sl = dx + ((size_t)iw * NPATH + pl) * NX, *sh = dx + ((size_t)iw * NPATH + (ph >= 0 ? ph : pl)) * NX;
    double *o = out + ((size_t)iw * NGEOM + ig) * (NX + 1);
    if (threadIdx.x == 0) {
        const double y = spec[(size_t)iw * NPATH + pl];
        o[0] = ph >= 0 ? __dadd_rn(__dmul_rn(y, a), __dmul_rn(spec[(size_t)iw * NPATH + ph], b)) : y;
    }
    for (int x = threadIdx.x; x < NX; x += blockDim.x)
        o[1 + x] = ph >= 0 ? __dadd_rn(__dmul_rn(sl[x], a), __dmul_rn(sh[x], b)) : sl[x];
}

extern "C" int ansb200_path_mix(const double *spec, const double *dx, const int32_t *lo, const int32_t *hi,
                                const double *wlo, const double *whi, int NWAVE, int NPATH, int NX, int NGEOM,
                                double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(spec && dx && lo && hi && wlo && whi && out, "path_mix: null pointer");
    ANS_REQUIRE(NWAVE > 0 && NPATH > 0 && NX > 0 && NGEOM > 0 && NGEOM <= 65535, "path_mix: bad shape");
    int threads = NX >= 256 ? 256 : ((NX + 31) / 32) * 32;
    ans_path_mix_kernel<<<dim3((unsigned)NWAVE, (unsigned)NGEOM), threads, 0, stream>>>(spec, dx, lo, hi, wlo, whi, NPATH, NX,
                                                                                         NGEOM, out);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
