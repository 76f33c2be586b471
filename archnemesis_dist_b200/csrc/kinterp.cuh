// kinterp.cuh -- one k-table interpolation, shared by the stand-alone and the fused kernels.
// Mirrors archnemesis/Spectroscopy_0.py:2391-2403 (calc_k) and :2238-2247 (calc_kg): bilinear in
// ln k when all four corners are positive, linear when all four are <= 0, zero when mixed.
// Products and sums are kept un-fused (explicit _rn intrinsics) so the rounding sequence is the
// reference's (numpy evaluates each product and sum separately).
#pragma once
#include "common.cuh"

struct AnsLayerPlan {
    const int32_t *ip_lo, *it_lo;
    const double *w4, *omv, *vv, *dudt;
};

// Device table layout (api.cu): plane-major [NP*NT][NWAVE][NG][NGAS] -- one (p,T) plane is one contiguous
// NWAVE*NG*NGAS run, so the NG*NGAS values a wavenumber needs from a plane are 960 contiguous bytes
// (at 20 x 6) instead of twenty 48-byte pieces 14 KB apart in the reference's layout.
// off00 addresses the (ip_lo, it_lo) corner; `plane` = NWAVE*NG*NGAS elements between consecutive T planes.
__device__ __forceinline__ double ans_tab_lnk(const AnsTab &T, size_t o)
{
    return T.lnK ? __ldg(T.lnK + o) : (double)__ldg(T.lnKf + o);
}
__device__ __forceinline__ double ans_tab_k(const AnsTab &T, size_t o)
{
    return T.K ? __ldg(T.K + o) : (double)__ldg(T.Kf + o);
}

template <bool GRAD>
__device__ __forceinline__ void ans_kinterp_elem(const AnsTab &T,
                                                 size_t off00, int NT, size_t plane, double w0, double w1, double w2,
                                                 double w3, double omv, double v, double dudt, double &kout,
                                                 double &dkout)
{
    const size_t off01 = off00 + plane, off10 = off00 + (size_t)NT * plane, off11 = off10 + plane;
    double l00 = ans_tab_lnk(T, off00), l01 = ans_tab_lnk(T, off01), l10 = ans_tab_lnk(T, off10), l11 = ans_tab_lnk(T, off11);
    bool fast = isfinite(l00) && isfinite(l01) && isfinite(l10) && isfinite(l11);
    if (!fast) {
        double k00 = ans_tab_k(T, off00), k01 = ans_tab_k(T, off01), k10 = ans_tab_k(T, off10), k11 = ans_tab_k(T, off11);
        if (k00 > 0.0 && k01 > 0.0 && k10 > 0.0 && k11 > 0.0) {
            l00 = log(k00); l01 = log(k01); l10 = log(k10); l11 = log(k11);   // only +inf entries land here
            fast = true;
        } else if (k00 <= 0.0 && k01 <= 0.0 && k10 <= 0.0 && k11 <= 0.0) {
            kout = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, k00), __dmul_rn(w1, k10)), __dmul_rn(w2, k11)),
                             __dmul_rn(w3, k01));
            if (GRAD) {
                double s = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(-k00, omv), __dmul_rn(k10, v)), __dmul_rn(k11, v)),
                                     __dmul_rn(k01, omv));
                dkout = __dmul_rn(s, dudt);
            }
            return;
        } else {
            kout = 0.0;
            if (GRAD) dkout = 0.0;
            return;
        }
    }
    double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, l00), __dmul_rn(w1, l10)), __dmul_rn(w2, l11)),
                         __dmul_rn(w3, l01));
    double kv = exp(x);
    kout = kv;
    if (GRAD) {
        double s = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(-l00, omv), __dmul_rn(l10, v)), __dmul_rn(l11, v)),
                             __dmul_rn(l01, omv));
        dkout = __dmul_rn(kv, __dmul_rn(s, dudt));
    }
}
