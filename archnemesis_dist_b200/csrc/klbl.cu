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
// One thread per (wavenumber, layer), layers fastest: tau is written fully coalesced, the NGAS+1 Jacobian
// entries of a thread are contiguous; the table reads of a warp hit at most a few planes at one wavenumber
// (L1 broadcast).  The kernel is bound by its output stream: 8*(NGAS+2) bytes written per (wavenumber, layer)
// against 32*NGAS bytes read per (wavenumber, distinct plane).
// np.sum over the gas axis is numpy's pairwise kernel: a plain running sum below 8 terms, eight interleaved
// accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder from 8 to 128 terms.
#include "common.cuh"

struct KlblTerm { double k, dkdT; };

template <bool GRAD>
__device__ __forceinline__ KlblTerm klbl_elem(const double *__restrict__ lnK, const double *__restrict__ K, size_t o00,
                                              size_t o01, size_t o10, size_t o11, double w0, double w1, double w2,
                                              double w3, double omv, double v, double du1, double du2)
{
    KlblTerm r{0.0, 0.0};
    double l00 = __ldg(lnK + o00), l01 = __ldg(lnK + o01), l10 = __ldg(lnK + o10), l11 = __ldg(lnK + o11);
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

template <bool GRAD>
__global__ void __launch_bounds__(256)
ans_klbl_opacity_kernel(const double *__restrict__ lnK, const double *__restrict__ K, const int32_t *__restrict__ corner,
                        const double *__restrict__ w4, const double *__restrict__ omv_, const double *__restrict__ vv_,
                        const double *__restrict__ du1_, const double *__restrict__ du2_,
                        const double *__restrict__ amount, int NWAVE, int NLAY, int NGAS, double *__restrict__ tau,
                        double *__restrict__ dk)
{
    const size_t plane = (size_t)NWAVE * NGAS;
    const size_t total = (size_t)NWAVE * NLAY;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int w = (int)(idx / NLAY), l = (int)(idx - (size_t)w * NLAY);
        const size_t wb = (size_t)w * NGAS;
        const size_t o00 = (size_t)corner[4 * l] * plane + wb, o01 = (size_t)corner[4 * l + 1] * plane + wb;
        const size_t o10 = (size_t)corner[4 * l + 2] * plane + wb, o11 = (size_t)corner[4 * l + 3] * plane + wb;
        const double w0 = w4[4 * l], w1 = w4[4 * l + 1], w2 = w4[4 * l + 2], w3 = w4[4 * l + 3];
        const double omv = GRAD ? omv_[l] : 0.0, v = GRAD ? vv_[l] : 0.0;
        const double du1 = GRAD ? du1_[l] : 0.0, du2 = GRAD ? du2_[l] : 0.0;
        double *dkrow = GRAD ? dk + idx * (size_t)(NGAS + 1) : nullptr;
        double dT = 0.0;
        auto term = [&](int i) -> double {
            const KlblTerm t = klbl_elem<GRAD>(lnK, K, o00 + i, o01 + i, o10 + i, o11 + i, w0, w1, w2, w3, omv, v, du1, du2);
            const double a = amount[(size_t)i * NLAY + l];
            if (GRAD) {
                dkrow[i] = t.k;
                dT = __dadd_rn(dT, __dmul_rn(t.dkdT, a));
            }
            return __dmul_rn(t.k, a);
        };
        double res = 0.0;
        if (NGAS < 8) {
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
}

extern "C" int ansb200_lbl_table_opacity(const ansb200_table *t, int NLAY, const int32_t *corner, const double *w4,
                                         const double *omv, const double *vv, const double *du1dt, const double *du2dt,
                                         const double *amount, int want_grad, double *tau, double *dk, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(t && corner && w4 && amount && tau, "lbl_table_opacity: null pointer");
    ANS_REQUIRE(t->NG == 1, "lbl_table_opacity: the table must have NG = 1 (got %d)", t->NG);
    ANS_REQUIRE(NLAY > 0, "lbl_table_opacity: NLAY must be positive");
    ANS_REQUIRE(t->NGAS <= 128, "lbl_table_opacity: NGAS=%d exceeds 128", t->NGAS);
    ANS_REQUIRE(!want_grad || (omv && vv && du1dt && du2dt && dk), "lbl_table_opacity: gradient requested without omv/vv/du1dt/du2dt/dk");
    const long long total = (long long)t->NWAVE * NLAY;
    int grid = ans_div_up(total, 256);
    if (grid > 148 * 8) grid = 148 * 8;            // 8 CTAs of 256 threads per SM, grid-stride over the rest
    if (want_grad)
        ans_klbl_opacity_kernel<true><<<grid, 256, 0, stream>>>(t->lnK, t->K, corner, w4, omv, vv, du1dt, du2dt, amount,
                                                                 t->NWAVE, NLAY, t->NGAS, tau, dk);
    else
        ans_klbl_opacity_kernel<false><<<grid, 256, 0, stream>>>(t->lnK, t->K, corner, w4, omv, vv, du1dt, du2dt, amount,
                                                                  t->NWAVE, NLAY, t->NGAS, tau, dk);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
