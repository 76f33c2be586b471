// klbl.cu -- gas opacity from pre-tabulated line-by-line tables (ILBL = LINE_BY_LINE_TABLES).
//
// Reference: Spectroscopy_0.calc_klbl / calc_klblg (archnemesis/Spectroscopy_0.py:1768-1919, :1601-1765)
// followed by the LBL-table branch of calculate_gaseous_line_opacity (ForwardModel_0.py:3795-3815):
//   k_i     = exp( w0 ln K(ip,it1) + w1 ln K(ip+1,it2) + w2 ln K(ip+1,it2+1) + w3 ln K(ip,it1+1) )   (linear in K
//             when all four corners are <= 0, zero when mixed)
//   dk_i/dT = k_i * ( -ln K(ip,it1) (1-v) du1 - ln K(ip+1,it2) v du2 + ln K(ip+1,it2+1) v du2 + ln K(ip,it1+1) (1-v) du1 )
//   TAUGAS  = np.sum_i( k_i * VLOSDENS_i )      dTAUGAS/dAMOUNT_i = k_i (* 1e-4 downstream)
//   dTAUGAS/dT = sum_i dk_i/dT * VLOSDENS_i     (running sum in gas order)
// The table is the NG = 1 resident table of api.cu (plane-major [NP*NT][NWAVE][NGAS], K and ln K); the four
// corner planes of a layer come from the host plan (plan.klbl_plan) as plane numbers, which also carries the
// reference's index wrap for a layer sitting exactly on the first temperature node of calc_klblg.
// One thread per (wavenumber, layer), layers fastest: tau is written fully coalesced, the Jacobian rows go
// through a shared-memory transpose (STAGE below); the table reads of a warp hit at most a few planes at one
// wavenumber (L1 broadcast).  Algorithmic traffic: 8*(NGAS+2) bytes written per (wavenumber, layer) against
// 8*NGAS bytes read per (wavenumber, distinct plane); the FP64 pipe (one exp per gas) is the other bound.
// np.sum over the gas axis is numpy's pairwise kernel: a plain running sum below 8 terms, eight interleaved
// accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder from 8 to 128 terms.
#include "common.cuh"

constexpr int KLBL_THREADS = 128;
constexpr int KLBL_STAGE_MAX = 16;      // NGAS+1 up to which the Jacobian rows are transposed through shared memory

struct KlblTerm { double k, dkdT; };

// One table entry from its four corner logarithms (non-finite = zero / negative table value: the slow path
// looks at K itself and takes the reference's linear or zero branch).
template <bool GRAD>
__device__ __forceinline__ KlblTerm klbl_value(double l00, double l01, double l10, double l11,
                                               const double *__restrict__ K, size_t o00, size_t o01, size_t o10,
                                               size_t o11, double w0, double w1, double w2, double w3, double omv,
                                               double v, double du1, double du2)
{
    KlblTerm r{0.0, 0.0};
    bool logs = isfinite(l00) && isfinite(l01) && isfinite(l10) && isfinite(l11);
    if (!logs) {
        const double k00 = __ldg(K + o00), k01 = __ldg(K + o01), k10 = __ldg(K + o10), k11 = __ldg(K + o11);
        if (k00 > 0.0 && k01 > 0.0 && k10 > 0.0 && k11 > 0.0) {            // +inf entries only
            l00 = log(k00); l01 = log(k01); l10 = log(k10); l11 = log(k11);
            logs = true;
        } else if (k00 <= 0.0 && k01 <= 0.0 && k10 <= 0.0 && k11 <= 0.0) {
            l00 = k00; l01 = k01; l10 = k10; l11 = k11;                      // linear branch: same sums on K itself
        } else {
            return r;
        }
    }
    const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, l00), __dmul_rn(w1, l10)), __dmul_rn(w2, l11)),
                               __dmul_rn(w3, l01));
    r.k = logs ? exp(x) : x;
    if (GRAD) {
        const double t1 = __dmul_rn(__dmul_rn(-l00, omv), du1), t2 = __dmul_rn(__dmul_rn(l10, v), du2);
        const double t3 = __dmul_rn(__dmul_rn(l11, v), du2), t4 = __dmul_rn(__dmul_rn(l01, omv), du1);
        const double s = __dadd_rn(__dadd_rn(__dsub_rn(t1, t2), t3), t4);
        r.dkdT = logs ? __dmul_rn(r.k, s) : s;
    }
    return r;
}

// PAIR: NGAS even and below 8 -- the corners of two gases come in one 16-byte load.  STAGE: the NGAS+1 Jacobian
// entries of the 32 (wavenumber, layer) cells of a warp are one contiguous run of dk; they are transposed through
// shared memory so that every store instruction writes 256 contiguous bytes instead of 32 pieces 8*(NGAS+1) apart.
template <bool GRAD, bool PAIR, bool STAGE>
__global__ void __launch_bounds__(KLBL_THREADS)
ans_klbl_opacity_kernel(const double *__restrict__ lnK, const double *__restrict__ K, const int32_t *__restrict__ corner,
                        const double *__restrict__ w4, const double *__restrict__ omv_, const double *__restrict__ vv_,
                        const double *__restrict__ du1_, const double *__restrict__ du2_,
                        const double *__restrict__ amount, int NWAVE, int NLAY, int NGAS, double *__restrict__ tau,
                        double *__restrict__ dk)
{
    extern __shared__ __align__(16) double klbl_smem[];
    const int lane = threadIdx.x & 31, NP1 = NGAS + 1;
    double *srow = klbl_smem + (size_t)(threadIdx.x - lane) * NP1;     // this warp's [32][NGAS+1] tile
    const size_t plane = (size_t)NWAVE * NGAS;
    const size_t total = (size_t)NWAVE * NLAY;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // the loop bound is warp-uniform (idx - lane) so that the staged stores can use warp collectives
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx - lane < total; idx += stride) {
        const bool live = idx < total;
        double res = 0.0, dT = 0.0;
        if (live) {
            const int w = (int)(idx / NLAY), l = (int)(idx - (size_t)w * NLAY);
            const size_t wb = (size_t)w * NGAS;
            const size_t o00 = (size_t)corner[4 * l] * plane + wb, o01 = (size_t)corner[4 * l + 1] * plane + wb;
            const size_t o10 = (size_t)corner[4 * l + 2] * plane + wb, o11 = (size_t)corner[4 * l + 3] * plane + wb;
            const double w0 = w4[4 * l], w1 = w4[4 * l + 1], w2 = w4[4 * l + 2], w3 = w4[4 * l + 3];
            const double omv = GRAD ? omv_[l] : 0.0, v = GRAD ? vv_[l] : 0.0;
            const double du1 = GRAD ? du1_[l] : 0.0, du2 = GRAD ? du2_[l] : 0.0;
            double *dkrow = GRAD ? (STAGE ? srow + (size_t)lane * NP1 : dk + idx * (size_t)NP1) : nullptr;
            auto finish = [&](int i, const KlblTerm &t) -> double {
                const double a = amount[(size_t)i * NLAY + l];
                if (GRAD) {
                    dkrow[i] = t.k;
                    dT = __dadd_rn(dT, __dmul_rn(t.dkdT, a));
                }
                return __dmul_rn(t.k, a);
            };
            auto term = [&](int i) -> double {
                return finish(i, klbl_value<GRAD>(__ldg(lnK + o00 + i), __ldg(lnK + o01 + i), __ldg(lnK + o10 + i),
                                                  __ldg(lnK + o11 + i), K, o00 + i, o01 + i, o10 + i, o11 + i, w0, w1, w2,
                                                  w3, omv, v, du1, du2));
            };
            if (PAIR) {
                for (int i = 0; i < NGAS; i += 2) {
                    const double2 a00 = __ldg(reinterpret_cast<const double2 *>(lnK + o00 + i));
                    const double2 a01 = __ldg(reinterpret_cast<const double2 *>(lnK + o01 + i));
                    const double2 a10 = __ldg(reinterpret_cast<const double2 *>(lnK + o10 + i));
                    const double2 a11 = __ldg(reinterpret_cast<const double2 *>(lnK + o11 + i));
                    res = __dadd_rn(res, finish(i, klbl_value<GRAD>(a00.x, a01.x, a10.x, a11.x, K, o00 + i, o01 + i, o10 + i,
                                                                    o11 + i, w0, w1, w2, w3, omv, v, du1, du2)));
                    res = __dadd_rn(res, finish(i + 1, klbl_value<GRAD>(a00.y, a01.y, a10.y, a11.y, K, o00 + i + 1,
                                                                        o01 + i + 1, o10 + i + 1, o11 + i + 1, w0, w1, w2,
                                                                        w3, omv, v, du1, du2)));
                }
            } else if (NGAS < 8) {
                for (int i = 0; i < NGAS; ++i) res = __dadd_rn(res, term(i));
            } else {
                const int m8 = NGAS - (NGAS & 7);
                double r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = term(j);
                for (int i0 = 8; i0 < m8; i0 += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], term(i0 + j));
                }
                res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (int i = m8; i < NGAS; ++i) res = __dadd_rn(res, term(i));
            }
            tau[idx] = res;
            if (GRAD) dkrow[NGAS] = dT;
        }
        if (GRAD && STAGE) {
            __syncwarp();
            const size_t base = (idx - lane) * (size_t)NP1;           // first dk element of the warp's cells
            const size_t end = total * (size_t)NP1;
            for (int c = 0; c < NP1; ++c) {
                const size_t o = base + (size_t)c * 32 + lane;
                if (o < end) dk[o] = srow[c * 32 + lane];
            }
            __syncwarp();
        }
    }
}

extern "C" int ansb200_lbl_table_opacity(const ansb200_table *t, int NLAY, const int32_t *corner, const double *w4,
                                         const double *omv, const double *vv, const double *du1dt, const double *du2dt,
                                         const double *amount, int want_grad, double *tau, double *dk, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(t && corner && w4 && amount && tau, "lbl_table_opacity: null pointer");
    ANS_REQUIRE(t->NG == 1, "lbl_table_opacity: the table must have NG = 1 (got %d)", t->NG);
    ANS_REQUIRE(t->storage == ANSB200_TABLE_F64, "lbl_table_opacity: float32 table storage is for k-tables only");
    ANS_REQUIRE(NLAY > 0, "lbl_table_opacity: NLAY must be positive");
    ANS_REQUIRE(t->NGAS <= 128, "lbl_table_opacity: NGAS=%d exceeds 128", t->NGAS);
    ANS_REQUIRE(!want_grad || (omv && vv && du1dt && du2dt && dk), "lbl_table_opacity: gradient requested without omv/vv/du1dt/du2dt/dk");
    const long long total = (long long)t->NWAVE * NLAY;
    int grid = ans_div_up(total, KLBL_THREADS);
    if (grid > 148 * 16) grid = 148 * 16;          // a multiple of the SM count, grid-stride over the rest
    const int NGAS = t->NGAS;
    const bool pair = (NGAS & 1) == 0 && NGAS < 8;   // plane offsets are multiples of NGAS: 16-byte aligned pairs
    const bool stage = want_grad && NGAS + 1 <= KLBL_STAGE_MAX;
    const size_t smem = stage ? (size_t)KLBL_THREADS * (NGAS + 1) * sizeof(double) : 0;
#define KLBL_LAUNCH(G, P, S)                                                                                          \
    ans_klbl_opacity_kernel<G, P, S><<<grid, KLBL_THREADS, smem, stream>>>(t->lnK, t->K, corner, w4, omv, vv, du1dt, du2dt, \
                                                                          amount, t->NWAVE, NLAY, NGAS, tau, dk)
    if (want_grad) {
        if (stage) { if (pair) KLBL_LAUNCH(true, true, true); else KLBL_LAUNCH(true, false, true); }
        else       { if (pair) KLBL_LAUNCH(true, true, false); else KLBL_LAUNCH(true, false, false); }
    } else {
        if (pair) KLBL_LAUNCH(false, true, false); else KLBL_LAUNCH(false, false, false);
    }
#undef KLBL_LAUNCH
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
