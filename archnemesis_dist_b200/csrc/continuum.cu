// continuum.cu -- continuum opacities on the device from a host-made plan (archnemesis_dist_b200/continuum.py).
//
// Reference: calc_tau_cia (archnemesis/ForwardModel_0.py:4516-4788), calc_tau_rayleighj / v2 (:5524-5710),
// calc_tau_dust (:4790-4867) and the fold into dTAUCON of calculate_layer_opacity (:3938-3981).
//
// One thread per (wavenumber, layer), layers fastest (the outputs [NWAVE,NLAY] and [NWAVE,NPAR,NLAY] are written
// coalesced).  Per thread, in the reference's order of operations (products and sums un-fused):
//   CIA      for every term t (a pair of gases with a cross-section table, or one of the fixed CO2-CO2 / N2-N2 / N2-H2
//            spectra):  k = table planes of this layer combined with (fhh_t, fhl_t, fhh_f, fhl_f)  [:4688-4695],
//            sum1 += k q1 q2; slot[A] += ca k; slot[B] += cb k; slot[T] += dk/dT q1 q2  [:4737-4750, :4752-4771];
//            TAUCIA = sum1 XFAC, slots *= XFAC  [:4773-4774]
//   Rayleigh TAURAY = sum_r ur vr, dTAURAY = sum_r ur vrd
//   aerosols TAUDUST = sum_i clip(nan_to_num(ud vd), 0, 1e20)  [:4861, :3960-3965]
//   dTAUCON  gas slots: slot / TOTAM + dTAURAY; temperature slot: slot[NVMR]; aerosol slots: ud  [:3941-3981]
// The tables kw[term][plane][wave] are the reference's K_CIA interpolated once onto the calculation wavenumbers
// (host, SciPy interp1d) and stay resident; per evaluation only the per-layer coefficients and a few spectra arrive.
#include "common.cuh"

constexpr int CONT_MAX_SLOTS = 24;

struct ContParams {
    const double *kw;
    const int32_t *nplanes;
    int NTERM, NPL;
    const int32_t *pl;
    const double *wt, *q1, *q2, *ca, *cb;
    const int32_t *slots;
    const double *xfac, *totam, *ur, *vr, *vrd, *ud, *vd;
    int NR, NDUST, NWAVE, NLAY, NVMR, has_cia, want_grad;
    double *taucia, *taudust, *tauray, *dtaucon;
};

__global__ void __launch_bounds__(128) ans_continuum_kernel(ContParams P)
{
    const long long n = (long long)P.NWAVE * P.NLAY;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int w = (int)(idx / P.NLAY), l = (int)(idx - (long long)w * P.NLAY);
    const int NS = P.NVMR + 2, NPAR = P.NVMR + 2 + P.NDUST;
    double d[CONT_MAX_SLOTS];
#pragma unroll
    for (int s = 0; s < CONT_MAX_SLOTS; ++s) d[s] = 0.0;
    double sum1 = 0.0;
    if (P.has_cia && P.NTERM > 0) {
        const double fhh_t = P.wt[5 * l], fhl_t = P.wt[5 * l + 1], fhh_f = P.wt[5 * l + 2], fhl_f = P.wt[5 * l + 3],
                     dfhldT = P.wt[5 * l + 4];
        const int p0 = P.pl[4 * l], p1 = P.pl[4 * l + 1], p2 = P.pl[4 * l + 2], p3 = P.pl[4 * l + 3];
        for (int t = 0; t < P.NTERM; ++t) {
            const double *kt = P.kw + ((size_t)t * P.NPL) * P.NWAVE + w;
            double k, dk = 0.0;
            if (P.nplanes[t] > 1) {
                const double a = kt[(size_t)p0 * P.NWAVE], b = kt[(size_t)p1 * P.NWAVE], c = kt[(size_t)p2 * P.NWAVE],
                             e = kt[(size_t)p3 * P.NWAVE];
                const double ktlo = __dadd_rn(__dmul_rn(a, fhh_t), __dmul_rn(b, fhl_t));
                const double kthi = __dadd_rn(__dmul_rn(c, fhh_t), __dmul_rn(e, fhl_t));
                k = __dadd_rn(__dmul_rn(ktlo, fhh_f), __dmul_rn(kthi, fhl_f));
                dk = __dmul_rn(__dsub_rn(kthi, ktlo), dfhldT);
            } else {
                k = kt[0];
            }
            const double q1 = P.q1[(size_t)t * P.NLAY + l], q2 = P.q2[(size_t)t * P.NLAY + l];
            sum1 = __dadd_rn(sum1, __dmul_rn(__dmul_rn(k, q1), q2));
            const int sa = P.slots[3 * t], sb = P.slots[3 * t + 1], st = P.slots[3 * t + 2];
            if (sa >= 0) d[sa] = __dadd_rn(d[sa], __dmul_rn(P.ca[(size_t)t * P.NLAY + l], k));
            if (sb >= 0) d[sb] = __dadd_rn(d[sb], __dmul_rn(P.cb[(size_t)t * P.NLAY + l], k));
            if (st >= 0 && P.nplanes[t] > 1) d[st] = __dadd_rn(d[st], __dmul_rn(__dmul_rn(dk, q1), q2));
        }
    }
    const double xf = P.xfac[l];
    if (P.taucia) P.taucia[idx] = P.has_cia ? __dmul_rn(sum1, xf) : 0.0;
    double tray = 0.0, dray = 0.0;
    for (int r = 0; r < P.NR; ++r) {
        const double u = P.ur[(size_t)r * P.NWAVE + w];
        tray = __dadd_rn(tray, __dmul_rn(u, P.vr[(size_t)r * P.NLAY + l]));
        dray = __dadd_rn(dray, __dmul_rn(u, P.vrd[(size_t)r * P.NLAY + l]));
    }
    if (P.tauray) P.tauray[idx] = tray;
    double tdust = 0.0;
    for (int i = 0; i < P.NDUST; ++i) {
        double v = __dmul_rn(P.ud[(size_t)i * P.NWAVE + w], P.vd[(size_t)i * P.NLAY + l]);
        v = isnan(v) ? 0.0 : v;                                   // np.nan_to_num, then np.clip(., 0, 1e20)
        v = fmin(fmax(v, 0.0), 1.0e20);
        tdust = __dadd_rn(tdust, v);
    }
    if (P.taudust) P.taudust[idx] = tdust;
    if (P.want_grad && P.dtaucon) {
        double *out = P.dtaucon + (size_t)w * NPAR * P.NLAY + l;
        const double tm = P.totam[l];
        for (int k = 0; k < NPAR; ++k) {
            double v = 0.0;
            if (k < P.NVMR) {
                if (P.has_cia) v = __ddiv_rn(__dmul_rn(d[k < NS ? k : 0], xf), tm);
                if (P.NR > 0) v = __dadd_rn(v, dray);
            } else if (k == P.NVMR) {
                if (P.has_cia) v = __dmul_rn(d[P.NVMR], xf);
            } else if (k <= P.NVMR + P.NDUST) {
                v = P.ud[(size_t)(k - P.NVMR - 1) * P.NWAVE + w];
            }
            out[(size_t)k * P.NLAY] = v;
        }
    }
}

extern "C" int ansb200_continuum(const double *kw, const int32_t *nplanes, int NTERM, int NPL, const int32_t *pl,
                                 const double *wt, const double *q1, const double *q2, const double *ca, const double *cb,
                                 const int32_t *slots, const double *xfac, const double *totam, const double *ur,
                                 const double *vr, const double *vrd, int NR, const double *ud, const double *vd, int NDUST,
                                 int NWAVE, int NLAY, int NVMR, int has_cia, int want_grad, double *taucia,
                                 double *taudust, double *tauray, double *dtaucon, void *stream_)
{
    ANS_REQUIRE(NWAVE > 0 && NLAY > 0 && NVMR >= 0 && NTERM >= 0 && NR >= 0 && NDUST >= 0, "continuum: bad shape");
    ANS_REQUIRE(NVMR + 2 <= CONT_MAX_SLOTS, "continuum: NVMR+2 = %d gradient slots exceed %d", NVMR + 2, CONT_MAX_SLOTS);
    ANS_REQUIRE(xfac && totam, "continuum: null pointer");
    ANS_REQUIRE(!(has_cia && NTERM > 0) || (kw && nplanes && pl && wt && q1 && q2 && ca && cb && slots),
                "continuum: CIA terms without their arrays");
    ANS_REQUIRE(NR == 0 || (ur && vr && vrd), "continuum: Rayleigh terms without their arrays");
    ANS_REQUIRE(NDUST == 0 || (ud && vd), "continuum: aerosol terms without their arrays");
    ANS_REQUIRE(!want_grad || dtaucon, "continuum: gradients requested without dtaucon");
    ContParams P{kw, nplanes, NTERM, NPL, pl, wt, q1, q2, ca, cb, slots, xfac, totam, ur, vr, vrd, ud, vd,
                 NR, NDUST, NWAVE, NLAY, NVMR, has_cia, want_grad, taucia, taudust, tauray, dtaucon};
    const long long n = (long long)NWAVE * NLAY;
    const long long grid = (n + 127) / 128;
    ANS_REQUIRE(grid < 2147483647LL, "continuum: NWAVE*NLAY too large");
    ans_continuum_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream_>>>(P);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
