/*
 * oracle/ansb200_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the archNEMESIS (v1.1.0) correlated-k forward model + Jacobian hot path,
 * written from the behaviour of the reference functions cited next to each routine.  It exists so
 * that the CUDA path can be checked on a machine where the (pure-Python) reference is absent and
 * so that bench.py can time a CPU baseline on the GPU box's host cores.
 *
 * Parity pin: every routine here is compared against the LIVE reference (imported from
 * /root/reference in the build container, see oracle/make_golden.py and tests/test_plan.py::test_oracle_matches_live_reference_functions)
 * and against the golden vectors committed under tests/golden/.  The reference's own test-suite pins
 * none of these functions in isolation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline/reference arms may load this
 * library.  The product package never imports it.
 *
 * Conventions: C-order arrays, float64 unless stated; all functions return void and are
 * re-entrant.  `nthreads` > 1 enables OpenMP over the wavenumber axis (the reference itself is
 * single threaded; the threaded form is the "N independent workers" idiom of
 * archnemesis/ForwardModel_0.py:2312-2337 applied to the embarrassingly parallel axis).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX4(a, b, c, d, B, C, D) ((((size_t)(a) * (B) + (b)) * (C) + (c)) * (D) + (d))

/* ------------------------------------------------------------------------------------------
 * k-table interpolation.  Follows archnemesis/Spectroscopy_0.py:2298-2404 (calc_k) and
 * :2147-2247 (calc_kg).  The per-layer bracket search and the (v,u) weights are made on the host
 * in Python with the reference's dtypes (float32 PRESS/TEMP on the .kta path) and handed in:
 *   ip_lo/it_lo         lower bracket indices (upper is always lower+1 in every branch)
 *   w4[l][0..3]         (1-v)(1-u), v(1-u), v*u, (1-v)u   as float64 values
 *   omv[l], vv[l], dudt[l]   (1-v), v, 1/(thi-tlo)        (gradient only)
 * Table K[NWAVE,NG,NP,NT,NGAS] (gas fastest, Spectroscopy_0.py:213).
 * Output k[NWAVE,NG,NLAY,NGAS] (+ dkdT).
 * ------------------------------------------------------------------------------------------ */
void orc_kinterp(const double *K, int NWAVE, int NG, int NP, int NT, int NGAS, int NLAY,
                 const int32_t *ip_lo, const int32_t *it_lo, const double *w4, const double *omv,
                 const double *vv, const double *dudt, int want_grad, double *k, double *dkdT,
                 int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int iw = 0; iw < NWAVE; ++iw) {
        for (int ig = 0; ig < NG; ++ig) {
            const double *slab = K + ((size_t)iw * NG + ig) * NP * NT * NGAS;
            for (int l = 0; l < NLAY; ++l) {
                const double *c00 = slab + ((size_t)ip_lo[l] * NT + it_lo[l]) * NGAS;       /* klo1 */
                const double *c01 = c00 + NGAS;                                           /* klo2 */
                const double *c10 = c00 + (size_t)NT * NGAS;                              /* khi1 */
                const double *c11 = c10 + NGAS;                                           /* khi2 */
                const double *w = w4 + 4 * l;
                for (int g = 0; g < NGAS; ++g) {
                    double klo1 = c00[g], klo2 = c01[g], khi1 = c10[g], khi2 = c11[g];
                    size_t o = IDX4(iw, ig, l, g, NG, NLAY, NGAS);
                    double kv = 0.0, dv = 0.0;
                    if (klo1 > 0.0 && klo2 > 0.0 && khi1 > 0.0 && khi2 > 0.0) {
                        double l00 = log(klo1), l10 = log(khi1), l11 = log(khi2), l01 = log(klo2);
                        double x = w[0] * l00 + w[1] * l10 + w[2] * l11 + w[3] * l01;
                        kv = exp(x);
                        if (want_grad) {
                            double dxdt = (-l00 * omv[l] - l10 * vv[l] + l11 * vv[l] + l01 * omv[l]) * dudt[l];
                            dv = kv * dxdt;
                        }
                    } else if (klo1 <= 0.0 && klo2 <= 0.0 && khi1 <= 0.0 && khi2 <= 0.0) {
                        kv = w[0] * klo1 + w[1] * khi1 + w[2] * khi2 + w[3] * klo2;
                        if (want_grad)
                            dv = (-klo1 * omv[l] - khi1 * vv[l] + khi2 * vv[l] + klo2 * omv[l]) * dudt[l];
                    }
                    k[o] = kv;
                    if (want_grad) dkdT[o] = dv;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Random-overlap gas mixing.  Follows archnemesis/ForwardModel_0.py:6029-6173 (k_overlap, rank)
 * and :5842-6026 (k_overlapg, rankg).
 *
 * weight[NG*NG]   del_g[i]*del_g[j] evaluated in del_g's own dtype (float32 on the .kta path),
 *                 then widened (ForwardModel_0.py:6087, :5915)
 * g_ord[NG+1]     0, cumsum(del_g) in del_g's dtype, last entry forced to 1 (:6141-6143)
 * Both are produced on the host (archnemesis_dist_b200.plan / oracle.py) so that the float32
 * behaviour of the reference is reproduced bit for bit.
 *
 * The reference sorts with numba's np.argsort, an UNSTABLE quicksort, so the order of equal keys
 * (and with it the split of the gradient rows across bin edges) is a property of numba's
 * implementation.  numba is a third-party dependency absent from /root/reference (setup.py:26
 * requires numba>=0.57; the container has numba 0.65.0): its published algorithm
 * (numba/misc/quicksort.py: median-of-three pivot stashed at the end, Hoare-style partition,
 * explicit stack that always pushes the larger side, insertion sort below 15 elements; float keys
 * compared with `a < b or (isnan(b) and not isnan(a))`, numba/np/arrayobj.py lt_floats) is
 * restated below so that the oracle reproduces the reference's permutation bit for bit, ties
 * included.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double key; int idx; } orc_kv;

static inline int orc_lt(double a, double b) { return a < b || (isnan(b) && !isnan(a)); }

#define ORC_SMALL_QUICKSORT 15
#define ORC_MAX_STACK 100

static void orc_numba_argsort(const double *A, int n, int *R)
{
    for (int i = 0; i < n; ++i) R[i] = i;
    if (n < 2) return;
    int stack_lo[ORC_MAX_STACK], stack_hi[ORC_MAX_STACK];
    int ns = 1;
    stack_lo[0] = 0; stack_hi[0] = n - 1;
#define ORC_SWAP(x, y) do { int t__ = R[x]; R[x] = R[y]; R[y] = t__; } while (0)
    while (ns > 0) {
        ns -= 1;
        int low = stack_lo[ns], high = stack_hi[ns];
        while (high - low >= ORC_SMALL_QUICKSORT) {
            /* partition A[low..high] around the median of {low, mid, high} */
            int mid = (low + high) >> 1;
            if (orc_lt(A[R[mid]], A[R[low]])) ORC_SWAP(low, mid);
            if (orc_lt(A[R[high]], A[R[mid]])) ORC_SWAP(high, mid);
            if (orc_lt(A[R[mid]], A[R[low]])) ORC_SWAP(low, mid);
            double pivot = A[R[mid]];
            ORC_SWAP(high, mid);
            int i = low, j = high - 1;
            for (;;) {
                while (i < high && orc_lt(A[R[i]], pivot)) i += 1;
                while (j >= low && orc_lt(pivot, A[R[j]])) j -= 1;
                if (i >= j) break;
                ORC_SWAP(i, j);
                i += 1;
                j -= 1;
            }
            ORC_SWAP(i, high);
            if (high - i > i - low) {
                if (high > i) { stack_lo[ns] = i + 1; stack_hi[ns] = high; ns += 1; }
                high = i - 1;
            } else {
                if (i > low) { stack_lo[ns] = low; stack_hi[ns] = i - 1; ns += 1; }
                low = i + 1;
            }
        }
        /* insertion sort of A[low..high] */
        for (int i = low + 1; i <= high; ++i) {
            int k = R[i];
            double v = A[k];
            int j = i;
            while (j > low && orc_lt(v, A[R[j - 1]])) { R[j] = R[j - 1]; j -= 1; }
            R[j] = k;
        }
    }
#undef ORC_SWAP
}

/* stable order (key, index): what the CUDA kernels produce for tied keys */
static int orc_kv_cmp(const void *a, const void *b)
{
    const orc_kv *x = (const orc_kv *)a, *y = (const orc_kv *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static int orc_sort_mode = 0;   /* 0: numba quicksort order (the reference), 1: stable (key, index) order */
void orc_set_sort_mode(int mode) { orc_sort_mode = mode; }

/* rank / rankg: sort `cont`, accumulate weights, rebin onto the g_ord bins.
 * grad may be NULL (rank).  npar_tot = row length of grad, n = leading columns that are live. */
static void orc_rank(int ng, const double *weight, const double *cont, const double *g_ord,
                     const double *grad, int npar_tot, int n, orc_kv *scratch, double *gdist,
                     double *k_g, double *dkdq)
{
    int nloop = ng * ng;
    if (orc_sort_mode == 0) {
        int R[1024];
        int *Rp = nloop <= 1024 ? R : (int *)malloc(sizeof(int) * nloop);
        orc_numba_argsort(cont, nloop, Rp);
        for (int i = 0; i < nloop; ++i) { scratch[i].key = cont[Rp[i]]; scratch[i].idx = Rp[i]; }
        if (Rp != R) free(Rp);
    } else {
        for (int i = 0; i < nloop; ++i) { scratch[i].key = cont[i]; scratch[i].idx = i; }
        qsort(scratch, (size_t)nloop, sizeof(orc_kv), orc_kv_cmp);
    }
    double run = 0.0;
    for (int i = 0; i < nloop; ++i) { run += weight[scratch[i].idx]; gdist[i] = run; }
    for (int i = 0; i < ng; ++i) k_g[i] = 0.0;
    if (grad) for (int i = 0; i < ng * npar_tot; ++i) dkdq[i] = 0.0;
    int ig = 0;
    double sum1 = 0.0;
    for (int iloop = 0; iloop < nloop; ++iloop) {
        int src = scratch[iloop].idx;
        double w = weight[src];
        double cw = scratch[iloop].key * w;
        const double *gr = grad ? grad + (size_t)src * npar_tot : NULL;
        if (ig < ng && gdist[iloop] < g_ord[ig + 1]) {
            k_g[ig] += cw;
            if (gr) for (int p = 0; p < n; ++p) dkdq[ig * npar_tot + p] += gr[p] * w;
            sum1 += w;
        } else {
            if (ig >= ng) break;   /* the reference indexes out of bounds here; never reached with sane del_g */
            double prev = gdist[iloop > 0 ? iloop - 1 : nloop - 1];   /* gdist[-1] wraps in the reference */
            double frac = (g_ord[ig + 1] - prev) / (gdist[iloop] - prev);
            k_g[ig] += frac * cw;
            if (gr) for (int p = 0; p < n; ++p) dkdq[ig * npar_tot + p] += frac * (gr[p] * w);
            sum1 += frac * w;
            k_g[ig] /= sum1;
            if (gr) for (int p = 0; p < n; ++p) dkdq[ig * npar_tot + p] /= sum1;
            ig += 1;
            if (ig < ng) {
                sum1 = (1.0 - frac) * w;
                k_g[ig] = (1.0 - frac) * cw;
                if (gr) for (int p = 0; p < n; ++p) dkdq[ig * npar_tot + p] = (1.0 - frac) * (gr[p] * w);
            }
        }
    }
    if (ig == ng - 1) {
        k_g[ig] /= sum1;
        if (grad) for (int p = 0; p < n; ++p) dkdq[ig * npar_tot + p] /= sum1;
    }
}

/* k[NWAVE,NG,NLAY,NGAS], dkdT same (NULL when !want_grad), amount[NGAS,NLAY]
 * -> tau[NWAVE,NG,NLAY], dk[NWAVE,NG,NLAY,NGAS+1] */
void orc_koverlap(const double *k, const double *dkdT, const double *amount, const double *weight,
                  const double *g_ord, int NWAVE, int NG, int NLAY, int NGAS, int want_grad,
                  double *tau, double *dk, int nthreads)
{
    const int NP1 = NGAS + 1, NN = NG * NG;
    if (NGAS == 1) {   /* ForwardModel_0.py:5871-5876, :6056-6058 */
        for (size_t iw = 0; iw < (size_t)NWAVE; ++iw)
            for (int ig = 0; ig < NG; ++ig)
                for (int l = 0; l < NLAY; ++l) {
                    size_t o = (iw * NG + ig) * NLAY + l;
                    tau[o] = k[o] * amount[l];
                    if (want_grad) { dk[o * 2] = k[o]; dk[o * 2 + 1] = dkdT[o] * amount[l]; }
                }
        return;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        double *rw = (double *)malloc(sizeof(double) * NN);   /* random_tau */
        double *rg = (double *)calloc((size_t)NN * NP1, sizeof(double));
        double *gd = (double *)malloc(sizeof(double) * NN);
        orc_kv *sc = (orc_kv *)malloc(sizeof(orc_kv) * NN);
        double *tau_g = (double *)malloc(sizeof(double) * NG);
        double *tmp_g = (double *)malloc(sizeof(double) * NG);
        double *dkp = (double *)malloc(sizeof(double) * NG * NP1);
        double *tmp_dkp = (double *)malloc(sizeof(double) * NG * NP1);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int iw = 0; iw < NWAVE; ++iw) {
            for (int l = 0; l < NLAY; ++l) {
#define KG(ig, gas) k[IDX4(iw, ig, l, gas, NG, NLAY, NGAS)]
#define DT(ig, gas) dkdT[IDX4(iw, ig, l, gas, NG, NLAY, NGAS)]
#define AM(gas) amount[(size_t)(gas) * NLAY + l]
                for (int i = 0; i < NG; ++i) tau_g[i] = 0.0;
                for (int i = 0; i < NG * NP1; ++i) dkp[i] = 0.0;
                memset(rg, 0, sizeof(double) * (size_t)NN * NP1);
                for (int igas = 0; igas < NGAS - 1; ++igas) {
                    int g1 = igas + 1;
                    if (igas == 0) {
                        if (KG(NG - 1, 0) * AM(0) <= 0.0) {
                            for (int i = 0; i < NG; ++i) {
                                tau_g[i] = KG(i, 1) * AM(1);
                                if (want_grad) { dkp[i * NP1 + 1] = KG(i, 1); dkp[i * NP1 + 2] = DT(i, 1) * AM(1); }
                            }
                        } else if (KG(NG - 1, 1) * AM(1) <= 0.0) {
                            for (int i = 0; i < NG; ++i) {
                                tau_g[i] = KG(i, 0) * AM(0);
                                if (want_grad) { dkp[i * NP1 + 0] = KG(i, 0); dkp[i * NP1 + 2] = DT(i, 0) * AM(0); }
                            }
                        } else {
                            int il = 0;
                            for (int ig = 0; ig < NG; ++ig)
                                for (int jg = 0; jg < NG; ++jg, ++il) {
                                    rw[il] = KG(ig, 0) * AM(0) + KG(jg, 1) * AM(1);
                                    if (want_grad) {
                                        rg[(size_t)il * NP1 + 0] = KG(ig, 0);
                                        rg[(size_t)il * NP1 + 1] = KG(jg, 1);
                                        rg[(size_t)il * NP1 + 2] = DT(ig, 0) * AM(0) + DT(jg, 1) * AM(1);
                                    }
                                }
                            orc_rank(NG, weight, rw, g_ord, want_grad ? rg : NULL, NP1, 3, sc, gd, tmp_g, tmp_dkp);
                            memcpy(tau_g, tmp_g, sizeof(double) * NG);
                            /* rankg returns a fresh (ng,nparam) array: columns >= n are zero */
                            if (want_grad) memcpy(dkp, tmp_dkp, sizeof(double) * NG * NP1);
                        }
                    } else {
                        if (KG(NG - 1, g1) * AM(g1) <= 0.0) {
                            if (want_grad)
                                for (int i = 0; i < NG; ++i) {
                                    dkp[i * NP1 + igas + 2] = dkp[i * NP1 + igas + 1];
                                    dkp[i * NP1 + igas + 1] *= 0.0;
                                }
                        } else if (tau_g[NG - 1] <= 0.0) {
                            for (int i = 0; i < NG; ++i) {
                                tau_g[i] = KG(i, g1) * AM(g1);
                                if (want_grad) { dkp[i * NP1 + g1] = KG(i, g1); dkp[i * NP1 + igas + 2] = DT(i, g1) * AM(g1); }
                            }
                        } else {
                            int il = 0;
                            for (int ig = 0; ig < NG; ++ig)
                                for (int jg = 0; jg < NG; ++jg, ++il) {
                                    rw[il] = tau_g[ig] + KG(jg, g1) * AM(g1);
                                    if (want_grad) {
                                        for (int p = 0; p < igas + 1; ++p) rg[(size_t)il * NP1 + p] = dkp[ig * NP1 + p];
                                        rg[(size_t)il * NP1 + igas + 1] = KG(jg, g1);
                                        rg[(size_t)il * NP1 + igas + 2] = dkp[ig * NP1 + igas + 1] + DT(jg, g1) * AM(g1);
                                    }
                                }
                            orc_rank(NG, weight, rw, g_ord, want_grad ? rg : NULL, NP1, igas + 3, sc, gd, tmp_g, tmp_dkp);
                            memcpy(tau_g, tmp_g, sizeof(double) * NG);
                            if (want_grad) memcpy(dkp, tmp_dkp, sizeof(double) * NG * NP1);
                        }
                    }
                }
                for (int ig = 0; ig < NG; ++ig) {
                    size_t o = ((size_t)iw * NG + ig) * NLAY + l;
                    tau[o] = tau_g[ig];
                    if (want_grad) for (int p = 0; p < NP1; ++p) dk[o * NP1 + p] = dkp[ig * NP1 + p];
                }
#undef KG
#undef DT
#undef AM
            }
        }
        free(rw); free(rg); free(gd); free(sc); free(tau_g); free(tmp_g); free(dkp); free(tmp_dkp);
    }
}

/* ------------------------------------------------------------------------------------------
 * Planck function and dB/dT.  archnemesis/ForwardModel_0.py:6183-6283.
 * ------------------------------------------------------------------------------------------ */
static inline double orc_planck(int ispace, double wave, double temp)
{
    const double c1 = 1.1911e-12, c2 = 1.439;
    double y, a;
    if (ispace == 0) { y = wave; a = c1 * pow(y, 3.0); }
    else { y = 1.0e4 / wave; a = c1 * pow(y, 5.0) / 1.0e4; }
    double tmp = c2 * y / temp;
    return a / (exp(tmp) - 1.0);
}

static inline void orc_planckg(int ispace, double wave, double temp, double *bb, double *dbdt)
{
    const double c1 = 1.1911e-12, c2 = 1.439;
    double y, a, ap;
    if (ispace == 0) { y = wave; a = c1 * pow(y, 3.0); ap = c1 * c2 * pow(y, 4.0) / pow(temp, 2.0); }
    else { y = 1.0e4 / wave; a = c1 * pow(y, 5.0) / 1.0e4; ap = c1 * c2 * pow(y, 6.0) / 1.0e4 / pow(temp, 2.0); }
    double tmp = c2 * y / temp;
    double e = exp(tmp);
    double b = e - 1.0;
    *bb = a / b;
    *dbdt = (e * ap) / pow(b, 2.0);
}

/* ------------------------------------------------------------------------------------------
 * Thermal emission along one path, no gradients.  archnemesis/ForwardModel_0.py:6287-6377.
 * tau[NWAVE,NG,NLAYIN] already gathered on the path and scaled.  emitot may be NULL.
 * ------------------------------------------------------------------------------------------ */
void orc_thermal(int ispace, const double *wave, const double *tau, const double *emitot,
                 const double *temp, const double *press, double tsurf, const double *emissivity,
                 const double *solflux, const double *reflectance, double sol_ang, double emiss_ang,
                 int NWAVE, int NG, int NLAYIN, double *spec, int nthreads)
{
    const double p1 = press[NLAYIN / 2 - 1 >= 0 ? NLAYIN / 2 - 1 : NLAYIN - 1];   /* PRESS[-1] wraps when NLAYIN==1 */
    const double p2 = press[NLAYIN - 1];
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (int iw = 0; iw < NWAVE; ++iw)
        for (int ig = 0; ig < NG; ++ig) {
            const double *t = tau + ((size_t)iw * NG + ig) * NLAYIN;
            double taud = 0.0, trold = 1.0, specg = 0.0;
            for (int j = 0; j < NLAYIN; ++j) {
                taud += t[j];
                double tr = exp(-taud);
                double bb = orc_planck(ispace, wave[iw], temp[j]);
                specg += (trold - tr) * bb;
                if (emitot) specg += emitot[(size_t)iw * NLAYIN + j] * tr;
                trold = tr;
            }
            if (p2 > p1) {
                double radground;
                if (tsurf <= 0.0) radground = orc_planck(ispace, wave[iw], temp[NLAYIN - 1]);
                else radground = orc_planck(ispace, wave[iw], tsurf) * emissivity[iw];
                specg += trold * radground;
            }
            if (emiss_ang < 90.0 && sol_ang < 90.0) {
                double mu = cos(emiss_ang / 180.0 * M_PI), mu0 = cos(sol_ang / 180.0 * M_PI);
                specg += trold * exp(-taud * mu / mu0) * solflux[iw] * reflectance[iw];
            }
            spec[(size_t)iw * NG + ig] = specg;
        }
}

/* ------------------------------------------------------------------------------------------
 * Thermal emission with layer gradients: literal O(NLAYIN^2 NPAR) recurrence of
 * archnemesis/ForwardModel_0.py:6380-6504.
 * tau[NWAVE,NG,NLAYIN], dtau[NWAVE,NG,NPAR,NLAYIN]
 * -> spec[NWAVE,NG], dspec[NWAVE,NG,NPAR,NLAYIN], dtsurf[NWAVE,NG]
 * ------------------------------------------------------------------------------------------ */
void orc_thermalg(int ispace, const double *wave, const double *tau, const double *dtau, int NVMR,
                  const double *temp, const double *press, double tsurf, const double *emissivity,
                  int NWAVE, int NG, int NPAR, int NLAYIN, double *spec, double *dspec,
                  double *dtsurf, int nthreads)
{
    const double p1 = press[NLAYIN / 2 - 1 >= 0 ? NLAYIN / 2 - 1 : NLAYIN - 1];
    const double p2 = press[NLAYIN - 1];
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        double *dtold = (double *)malloc(sizeof(double) * NPAR * NLAYIN);
        double *dtr = (double *)malloc(sizeof(double) * NPAR * NLAYIN);
        double *dsg = (double *)malloc(sizeof(double) * NPAR * NLAYIN);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 2)
#endif
        for (int iw = 0; iw < NWAVE; ++iw)
            for (int ig = 0; ig < NG; ++ig) {
                const double *t = tau + ((size_t)iw * NG + ig) * NLAYIN;
                const double *dt = dtau + ((size_t)iw * NG + ig) * NPAR * NLAYIN;
                double trold = 1.0, specg = 0.0;
                for (int i = 0; i < NPAR * NLAYIN; ++i) { dtold[i] = 0.0; dtr[i] = 0.0; dsg[i] = 0.0; }
                for (int j = 0; j < NLAYIN; ++j) {
                    double tlayer = exp(-t[j]);
                    double tr = trold * tlayer;
                    double bb, dbdt;
                    orc_planckg(ispace, wave[iw], temp[j], &bb, &dbdt);
                    specg += (trold - tr) * bb;
                    for (int kk = 0; kk < NPAR; ++kk) {
                        double *o = dtold + (size_t)kk * NLAYIN, *r = dtr + (size_t)kk * NLAYIN,
                               *s = dsg + (size_t)kk * NLAYIN;
                        for (int j1 = 0; j1 < j; ++j1) {
                            r[j1] = o[j1] * tlayer;
                            s[j1] += (o[j1] - r[j1]) * bb;
                        }
                        double tmp = dt[(size_t)kk * NLAYIN + j];
                        r[j] = -tmp * tlayer * trold;
                        s[j] += (o[j] - r[j]) * bb;
                        if (kk == NVMR) s[j] += (trold - tr) * dbdt;
                    }
                    trold = tr;
                    for (int kk = 0; kk < NPAR; ++kk)
                        for (int j1 = 0; j1 <= j; ++j1) dtold[(size_t)kk * NLAYIN + j1] = dtr[(size_t)kk * NLAYIN + j1];
                }
                double tempgtsurf = 0.0;
                if (p2 > p1) {
                    double radground, dradground;
                    if (tsurf <= 0.0) orc_planckg(ispace, wave[iw], temp[NLAYIN - 1], &radground, &dradground);
                    else {
                        orc_planckg(ispace, wave[iw], tsurf, &radground, &dradground);
                        radground *= emissivity[iw];
                        dradground *= emissivity[iw];
                    }
                    specg += trold * radground;
                    tempgtsurf = trold * dradground;
                    for (int i = 0; i < NPAR * NLAYIN; ++i) dsg[i] += radground * dtold[i];
                }
                spec[(size_t)iw * NG + ig] = specg;
                memcpy(dspec + ((size_t)iw * NG + ig) * NPAR * NLAYIN, dsg, sizeof(double) * NPAR * NLAYIN);
                dtsurf[(size_t)iw * NG + ig] = tempgtsurf;
            }
        free(dtold); free(dtr); free(dsg);
    }
}

/* ------------------------------------------------------------------------------------------
 * Line-by-line absorption (Voigt core + 1/dnu^2 wings).  archnemesis/LineData_0.py:123-358.
 * The line profile itself is a callback so that the Python side can hand in
 * scipy.special.cython_special.voigt_profile -- the arithmetic the reference calls
 * (archnemesis/lineshape/_scipy_support.py:13-38, voigt_impl/voigt_scipy.py:52).
 * broadening[3*M, N]: rows (gamma, n, delta) per molecule (self first), mix[M].
 * ------------------------------------------------------------------------------------------ */
typedef double (*orc_lineshape_fn)(double dwn, double alpha_d, double gamma_l);

#define ORC_C2_CGS (2.99792458E10 * 6.62607015E-27 / 1.380649E-16)

void orc_lbl_absorption(const double *wn_grid, int NWAVE, orc_lineshape_fn shape, double t_calc,
                        double t_ref, double p_calc, double p_ref, double q_ratio, double abundance,
                        double mass, const double *mix, int M, const double *broadening,
                        const double *nu, const double *sw, const double *e_lower,
                        const double *stim_ref, int N, double s_floor, double wn_calc_window,
                        double wn_approx_window, double *out)
{
    const double c2 = ORC_C2_CGS;
    const double dconst = (1.0 / 2.99792458E10) * sqrt(2 * log(2.0) * 6.02214129E+23 * 1.380649E-16);
    const double boltz = c2 * (t_calc - t_ref) / (t_calc * t_ref);
    const double t_ratio = t_ref / t_calc, p_ratio = p_calc / p_ref;
    for (int i = 0; i < N; ++i) {
        double strength = sw[i] * ((1 - exp(-c2 * nu[i] / t_calc)) / stim_ref[i]) * exp(boltz * e_lower[i]) * q_ratio;
        if (strength < s_floor) continue;
        double alpha_d = dconst * nu[i] * sqrt(t_calc / mass);
        double gamma_l = 0.0, shift = 0.0;
        for (int j = 0; j < M; ++j) {
            gamma_l += pow(t_ratio, broadening[(size_t)(3 * j + 1) * N + i]) * broadening[(size_t)(3 * j) * N + i] * mix[j] * p_ratio;
            shift += (p_ratio * broadening[(size_t)(3 * j + 2) * N + i]) * mix[j];
        }
        double approx_const = shape(wn_calc_window, alpha_d, gamma_l);
        for (int j = 0; j < NWAVE; ++j) {
            double d = wn_grid[j] - (nu[i] + shift);
            if (d >= wn_approx_window) break;
            if (d < -wn_approx_window) continue;
            if (-wn_calc_window <= d && d < wn_calc_window)
                out[j] += abundance * strength * shape(d, alpha_d, gamma_l);
            else
                out[j] += abundance * strength * approx_const * pow(wn_calc_window, 2.0) / (d * d);
        }
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
