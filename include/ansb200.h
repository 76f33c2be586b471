/*
 * ansb200.h -- C ABI of libansb200.so: the B200 (sm_100a) correlated-k forward model + Jacobian
 * hot path of archNEMESIS (reference: juanaldayparejo/archnemesis-dist v1.1.0).
 *
 * The reference has no FFI layer: its "operator API" is the Python method surface of
 * ForwardModel_0 and the module-level numba functions it calls.  Each entry point below replaces
 * one of those functions; the reference file:line it stands in for is cited beside it.
 *
 * Conventions
 *   - every function returns 0 (ANSB200_OK) or a negative error code; ansb200_last_error() gives
 *     a thread-local message for the last failure;
 *   - every array argument is a DEVICE pointer owned by the caller unless the name ends in
 *     `_host`; the library never frees or retains caller memory except through the opaque table
 *     handle; arrays are C-order float64 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no internal
 *     synchronisation, so calls are asynchronous with respect to the host;
 *   - no global mutable state besides the handles; distinct handles/streams may be used from
 *     distinct threads.
 */
#ifndef ANSB200_H
#define ANSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANSB200_OK 0
#define ANSB200_EINVAL (-1)  /* bad argument / unsupported shape */
#define ANSB200_ECUDA (-2)   /* CUDA runtime error (message holds cudaGetErrorString) */
#define ANSB200_ENOMEM (-3)  /* device allocation failed */

#define ANSB200_MAX_NG 22    /* NG*NG <= 512 keys per register sort */
#define ANSB200_MAX_NGAS 15  /* NGAS+1 gradient columns <= 16 */

/* ansb200_radiance flags */
#define ANSB200_RAD_GRAD 1u          /* produce layer-space gradients */
#define ANSB200_RAD_NAN_TO_NUM 2u    /* np.nan_to_num on g-integrated gradients (ForwardModel_0.py:4507) */
#define ANSB200_RAD_LAYER_SPACE 4u   /* gradients per LAYER, dspec[NWAVE,NPATH,NPAR,NLAY]: the visits of a layer by a limb /
                                        occultation path are added (the projection is linear); where
                                        ansb200_radiance_layer_space() says so */

typedef struct ansb200_table ansb200_table;

const char *ansb200_last_error(void);
int ansb200_version(void);

/* ---- table residency ---------------------------------------------------------------------
 * Replaces the per-call Spectroscopy_0.read_tables (archnemesis/Spectroscopy_0.py:1448-1528):
 * the cropped table K[NWAVE,NG,NP,NT,NGAS] (gas fastest, :213) is copied to the device once and
 * ln K is tabulated beside it (-inf for K==0, NaN for K<0) so k-interp costs one exp per output.
 * `K` may be a host or a device pointer (is_device selects). */
int ansb200_table_create(const double *K, int is_device, int NWAVE, int NG, int NP, int NT, int NGAS,
                         ansb200_table **out, void *stream);
/* Storage variants of the resident copy (k-tables only; ansb200_table_create = F64):
 *   ANSB200_TABLE_F64  K and ln K as float64 (16 bytes per entry)
 *   ANSB200_TABLE_K32  K as float32, ln K as float64 (12 bytes): LOSSLESS for .kta / .lta data, whose values are
 *                      float32 (Spectroscopy_0.py:2849) -- results are bit-identical to F64; refused (EINVAL) if some
 *                      value is not a float32 number
 *   ANSB200_TABLE_F32  K and ln K as float32 (8 bytes): the FP32 k-interp variant BASELINE.json allows, reported
 *                      separately; ln K rounded to float32 costs ~4e-6 relative in k (tolerance 2e-5 in the tests) */
#define ANSB200_TABLE_F64 0
#define ANSB200_TABLE_K32 1
#define ANSB200_TABLE_F32 2
int ansb200_table_create_ex(const double *K, int is_device, int NWAVE, int NG, int NP, int NT, int NGAS, int storage,
                            ansb200_table **out, void *stream);
int ansb200_table_destroy(ansb200_table *t);
int ansb200_table_shape(const ansb200_table *t, int *NWAVE, int *NG, int *NP, int *NT, int *NGAS);
/* device pointers to the resident copies (for tests / diagnostics) */
const double *ansb200_table_k(const ansb200_table *t);
const double *ansb200_table_lnk(const ansb200_table *t);

/* ---- k-table interpolation ----------------------------------------------------------------
 * Replaces Spectroscopy_0.calc_k (Spectroscopy_0.py:2298-2437, WAVECALC=None) and calc_kg
 * (:2147-2295).  The bracket search and the (v,u) weights are evaluated on the host with the
 * reference's dtypes (float32 PRESS/TEMP on the .kta path) and passed per layer:
 *   ip_lo[NLAY], it_lo[NLAY] int32   lower bracket (upper = lower+1 in every branch)
 *   w4[NLAY,4]   (1-v)(1-u), v(1-u), v*u, (1-v)u       omv/vv/dudt[NLAY]: 1-v, v, 1/(thi-tlo)
 * Output k[NWAVE,NG,NLAY,NGAS] and, if want_grad, dkdT of the same shape. */
int ansb200_kinterp(const ansb200_table *t, int NLAY, const int32_t *ip_lo, const int32_t *it_lo,
                    const double *w4, const double *omv, const double *vv, const double *dudt,
                    int want_grad, double *k, double *dkdT, void *stream);

/* ---- random-overlap gas mixing ------------------------------------------------------------
 * Replaces k_overlap + rank (archnemesis/ForwardModel_0.py:6029-6173) and k_overlapg + rankg
 * (:5842-6026).  weight[NG*NG] = del_g[i]*del_g[j] and g_ord[NG+1] = {0, cumsum(del_g)[..], 1}
 * are made on the host in del_g's own dtype (float32 on the .kta path) and widened.
 * del_g[NG] (widened, may be NULL) lets the kernel form the products itself instead of looking them
 * up; it is used only when it reproduces `weight` bit for bit (checked on the device).
 * amount[NGAS,NLAY] in cm-2.  Output tau[NWAVE,NG,NLAY]; if want_grad also
 * dk[NWAVE,NG,NLAY,NGAS+1] (d tau/d amount_gas ..., d tau/dT).
 * Equal sort keys follow the order of numba's (unstable) quicksort, which the reference's np.argsort uses:
 * the split of their gradient rows across a bin edge depends on it ("Ties"). */
int ansb200_koverlap(const double *k, const double *dkdT, const double *amount, const double *weight,
                     const double *g_ord, const double *del_g, int NWAVE, int NG, int NLAY, int NGAS, int want_grad,
                     double *tau, double *dk, void *stream);

/* Fused k-interp + overlap: k_gas / dkgasdT never reach HBM.  Replaces the K_TABLES branch of
 * ForwardModel_0.calculate_gaseous_line_opacity (ForwardModel_0.py:3850-3877). */
int ansb200_gas_opacity(const ansb200_table *t, int NLAY, const int32_t *ip_lo, const int32_t *it_lo,
                        const double *w4, const double *omv, const double *vv, const double *dudt,
                        const double *amount, const double *weight, const double *g_ord, const double *del_g,
                        int want_grad, double *tau, double *dk, void *stream);

/* Diagnostics of the two overlap kernels.  The common case (NG = 20, float32-born quadrature weights, k ascending in
 * g) runs in a fast kernel that hands the cells it declines -- equal sort keys on a bin edge, where the reference's
 * result depends on numba's unstable argsort, non-monotone k, ... -- to the general kernel through a device-side
 * work list.  ansb200_overlap_mode: 0 = that (default), 1 = general kernel only, 2 = as 0 but every call
 * synchronises and records its counts; returns the previous mode (mode < 0 only queries).  ansb200_overlap_stats:
 * out[0] = cells handed over by the last call in mode 2 (-1: all), out[1..5] = reasons (non-monotone input, bin left
 * open, group of equal key bits, exact tie on an edge, non-monotone bins), out[6] / out[7] = folds that used a
 * data-independent / a sorted order, out[8] = static orders rejected for an unseparated straddler. */
int ansb200_overlap_mode(int mode);
void ansb200_overlap_stats(int32_t *out9);

/* ---- path radiance + layer-space Jacobian -------------------------------------------------
 * Replaces, for every path at once, the opacity assembly of calculate_layer_opacity
 * (ForwardModel_0.py:3989-4012: continuum add, gas-gradient scatter :3868-3872, LAYINC gather and
 * SCALE), calc_thermal_emission_spectrum[g] (:6287-6504) with the unit scaling of :4244-4247, or
 * calculate_transmission_spectrum (:4104-4129), and the g-integration of CIRSrad (:4504-4508).
 *   mode            0 = thermal emission (IMOD THERMAL_EMISSION), 1 = transmission
 *   tau[NWAVE,NG,NLAY], dk[NWAVE,NG,NLAY,NGAS+1] (NULL unless GRAD)   gas opacity (koverlap output)
 *   gas_slot[NGAS] int32   parameter index of each active gas (Atmosphere.locate_gas)
 *   taucia/taudust/tauray[NWAVE,NLAY]   continuum opacities, added in the reference's order
 *                          TAUGAS + TAUCIA + TAUDUST + TAURAY (:3989); any may be NULL
 *   dtaucon[NWAVE,NPAR,NLAY] (NULL -> zeros)
 *   layinc[NLAYMAX,NPATH] int32, scale[NLAYMAX,NPATH], nlayin[NPATH] int32, emtemp[NLAYMAX,NPATH]
 *   laypress[NLAY]         layer pressures (for the limb/nadir test of :6354-6357)
 *   wave[NWAVE], delg[NG] (float64 copy of DELG), emissivity[NWAVE], xfac[NWAVE]
 *   solflux/reflectance[NWAVE], sol_ang/emiss_ang[NPATH]: non-gradient solar term (:6368-6373); may be NULL
 * Outputs spec[NWAVE,NPATH]; with GRAD dspec[NWAVE,NPATH,NPAR,NLAYMAX] (path-major, unlike the
 * reference's (NWAVE,NPAR,NLAYIN,NPATH); the Python face transposes) and, thermal only,
 * dtsurf[NWAVE,NPATH].  The (NWAVE,NG,NPAR,NLAYIN,NPATH) tensor of the reference is never formed. */
int ansb200_radiance(int mode, unsigned flags, const double *tau, const double *dk, const int32_t *gas_slot,
                     const double *taucia, const double *taudust, const double *tauray, const double *dtaucon,
                     const int32_t *layinc, const double *scale, const int32_t *nlayin, const double *emtemp,
                     const double *laypress, const double *wave, const double *delg, const double *emissivity,
                     const double *xfac, const double *solflux, const double *reflectance, const double *sol_ang,
                     const double *emiss_ang, int ispace, double tsurf, int NWAVE, int NG, int NLAY, int NGAS,
                     int NVMR, int NPAR, int NLAYMAX, int NPATH, double *spec, double *dspec, double *dtsurf,
                     void *stream);

/* ---- continuum opacities ------------------------------------------------------------------------
 * Replaces the dense host arrays of calc_tau_cia (archnemesis/ForwardModel_0.py:4516-4788), calc_tau_rayleighj / v2
 * (:5524-5710), calc_tau_dust (:4790-4867) and their fold into dTAUCON in calculate_layer_opacity (:3938-3981).  The
 * host (continuum.py) evaluates the reference's per-layer logic and hands over a plan; this makes the arrays
 * ansb200_radiance reads, on the device:
 *   kw[NTERM,NPL,NWAVE]   CIA cross sections on the calculation wavenumbers (resident); nplanes[NTERM] = NPL for a
 *                         table term, 1 for a fixed spectrum (CO2-CO2, N2-N2, N2-H2)
 *   pl[NLAY,4] int32, wt[NLAY,5]   the layer's four (para, T) planes and (fhh_t, fhl_t, fhh_f, fhl_f, dfhl/dT)
 *   q1,q2,ca,cb[NTERM,NLAY], slots[NTERM,3] int32   mixing ratios of the pair, coefficients and gradient slots of
 *                         d/dq1, d/dq2, d/dT (-1: none); xfac[NLAY] = TOTAM^2/XLEN, totam[NLAY]
 *   ur[NR,NWAVE], vr/vrd[NR,NLAY]   Rayleigh: TAURAY = sum ur*vr, dTAURAY = sum ur*vrd
 *   ud[NDUST,NWAVE], vd[NDUST,NLAY] aerosols: kext*1e-4 and column densities
 * Outputs taucia / taudust / tauray [NWAVE,NLAY] (any may be NULL) and, if want_grad, dtaucon[NWAVE,NVMR+2+NDUST,NLAY].
 * NVMR + 2 <= 24. */
int ansb200_continuum(const double *kw, const int32_t *nplanes, int NTERM, int NPL, const int32_t *pl, const double *wt,
                      const double *q1, const double *q2, const double *ca, const double *cb, const int32_t *slots,
                      const double *xfac, const double *totam, const double *ur, const double *vr, const double *vrd,
                      int NR, const double *ud, const double *vd, int NDUST, int NWAVE, int NLAY, int NVMR, int has_cia,
                      int want_grad, double *taucia, double *taudust, double *tauray, double *dtaucon, void *stream);

/* 1 if ansb200_radiance can produce layer-space gradients for this shape: >= 4 paths whose working set fits in shared
 * memory; transmission (mode 1), or thermal emission (mode 0) with dk given and NLAYMAX <= 224 path positions. */
int ansb200_radiance_layer_space(int mode, unsigned flags, int NG, int NLAY, int NGAS, int NPAR, int NPATH, int NLAYMAX,
                                 int has_dk, int has_dtaucon);

/* ---- gas opacity from line-by-line tables ------------------------------------------------------
 * Replaces Spectroscopy_0.calc_klbl / calc_klblg (archnemesis/Spectroscopy_0.py:1768-1919, :1601-1765)
 * and the LBL-table branch of calculate_gaseous_line_opacity (ForwardModel_0.py:3795-3815).  The table
 * is created from K[NWAVE,NP,NT,NGAS] with NG = 1.  Host plan per layer (plan.klbl_plan): corner[NLAY,4]
 * = plane numbers ip*NT+it of (ip,it1), (ip,it1+1), (ip+1,it2), (ip+1,it2+1); w4[NLAY,4] =
 * (1-v)(1-u1), v(1-u2), v u2, (1-v)u1; omv = 1-v, vv = v, du1dt, du2dt (gradient only).
 * amount[NGAS,NLAY] = LayerX.AMOUNT[:,IGAS]*1e-4.  tau[NWAVE,1,NLAY]; dk[NWAVE,1,NLAY,NGAS+1] with
 * entry i = k_i (d tau/d amount_i before the 1e-4) and entry NGAS = d tau/dT, as ansb200_gas_opacity. */
int ansb200_lbl_table_opacity(const ansb200_table *t, int NLAY, const int32_t *corner, const double *w4,
                              const double *omv, const double *vv, const double *du1dt, const double *du2dt,
                              const double *amount, int want_grad, double *tau, double *dk, void *stream);

/* ---- layer -> profile -> state vector ------------------------------------------------------
 * Replaces map2pro + map2xvec (ForwardModel_0.py:5319-5424).  The host folds
 * D[LAYINC[j,path],:] (DAM/DTE/DCO per parameter) with xmap into M[NPATH, NPAR*NLAYMAX, NX];
 * this computes out[NWAVE,NPATH,NX] = sum_{k,j} dspec[NWAVE,path,k,j] * M[path,(k,j),x]. */
int ansb200_jacobian_project(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAYMAX,
                             int NPATH, int NX, double *out, void *stream);
/* The same with layer-space gradients dspec[NWAVE,NPATH,NPAR,NLAY] and ONE matrix M[NPAR*NLAY, NX] for all paths
 * (M[k*NLAY + l, x] = sum_pro D_k[l,pro] xmap[x,k,pro]: D does not depend on the path). */
int ansb200_jacobian_project_shared(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAY, int NPATH,
                                    int NX, double *out, void *stream);
/* Either product with the list of the non-empty 16-row chunks of M (ascending chunk numbers row/16, int32 on the device;
 * every row of M outside the listed chunks must be zero for all paths and columns): whole parameters without a
 * state-vector element are then never read from dspec.  shared != 0: M[1, NPAR*NLAYMAX, NX] for every path. */
int ansb200_jacobian_project_chunks(const double *dspec, const double *M, int NWAVE, int NPAR, int NLAYMAX, int NPATH,
                                    int NX, int shared, const int32_t *chunks, int nchunks, double *out, void *stream);
/* The product with M handed over as a sparse operator (host: plan.sparse_projection): column x of path p (p = 0 for
 * every path when shared != 0) is ONE run of consecutive rows, M[p, r0 : r0+len, x] = vals[voff : voff+len] with
 * r0 / len / voff = col_r0 / col_len / col_voff[p*NX + x] (len = 0: empty column; zeros inside a run are stored);
 * long_cols[long_ptr[p] : long_ptr[p+1]] lists the columns longer than 16.  One warp per (wavenumber, path) row of
 * dspec, NPAR*NLAYMAX*8 bytes of shared memory per warp (NPAR*NLAYMAX <= 3200). */
int ansb200_jacobian_project_sparse(const double *dspec, const int32_t *col_r0, const int32_t *col_len,
                                    const int32_t *col_voff, const double *vals, int nvals, const int32_t *long_cols,
                                    const int32_t *long_ptr, int NWAVE, int NPAR, int NLAYMAX, int NPATH, int NX,
                                    int shared, double *out, void *stream);

/* ---- tangent-height interpolation of the path spectra ------------------------------------------
 * Replaces the SPECMOD / dSPECMOD loop of nemesisSOfmg / nemesisLfmg (ForwardModel_0.py:1206-1228, :1464-1486):
 * out[NWAVE,NGEOM,1+NX] = [spec | dx][:, lo[i]] * wlo[i] + [spec | dx][:, hi[i]] * whi[i]  (hi[i] < 0: path lo[i] alone).
 * spec[NWAVE,NPATH], dx[NWAVE,NPATH,NX] as ansb200_jacobian_project leaves them; lo/hi/wlo/whi[NGEOM] on the device. */
int ansb200_path_mix(const double *spec, const double *dx, const int32_t *lo, const int32_t *hi, const double *wlo,
                     const double *whi, int NWAVE, int NPATH, int NX, int NGEOM, double *out, void *stream);

/* ---- k-distributions from a monochromatic spectrum (k-table generation) ------------------------
 * Replaces the per-bin tail of calc_ktable_chunk (archnemesis/Spectroscopy_0.py:3619-3660): for bin b the absorption
 * coefficients kabs[lo[b]:hi[b]] of the line-by-line grid are sorted, g_i = cumsum(w_i) / sum(w) in sorted order
 * (w = the instrument function at the point times the grid step; NULL: equal weights, g_i = (i+1)/n) and
 * out[b, :] = np.interp(g_ord, g_sorted, k_sorted).  w holds the weights of bin b at w[woff[b] : woff[b] + n_b]
 * (bins may overlap when an instrument function widens them).  max_n = the longest bin; it must not exceed
 * ansb200_kdist_capacity(weighted) (16384 / 8192 points: the sort runs in shared memory). */
int ansb200_kdist_capacity(int weighted);
int ansb200_kdist(const double *kabs, const double *w, const int32_t *lo, const int32_t *hi, const int64_t *woff,
                  int NBIN, int max_n, const double *g_ord, int NG, double *out, void *stream);

/* ---- instrument line shape -------------------------------------------------------------------
 * Replaces Measurement_0.conv / convg for k-tables (archnemesis/Measurement_0.py:2288-2465,
 * :2467-2692): mode 0 = FWHM == 0, scipy interp1d onto the convolution points (rows of two entries
 * (hi, lo) with SciPy's weights; with col0_np_interp column 0 -- the spectrum -- is evaluated in
 * np.interp's slope form from np_lo[NCONV], np_exact[NCONV], xinfo[NCONV,3] = (x_lo, x_hi, x_new);
 * col0_np_interp = 2 evaluates EVERY column that way: Measurement_0.lblconv / lblconvg with FWHM == 0,
 * :2176-2186, :2267-2284);
 * mode 1 = FWHM < 0, filter-weighted mean sum(wval*in)/norm over widx[row_start[c]:row_start[c+1]]
 * (also lblconv / lblconvg with FWHM > 0 or < 0, :3335-4076: plan.lbl_conv_operator).
 * The operator is built on the host (plan.conv_operator).  in[NWAVE, ld] (first NCOL columns used,
 * e.g. [spectrum | Jacobian columns]) -> out[NCONV, NCOL].  Bit-identical to the reference. */
int ansb200_convolve(const double *in, int NWAVE, int NCOL, int ld, int mode, int col0_np_interp,
                     const int32_t *row_start, const int32_t *widx, const double *wval, const double *norm,
                     const int32_t *np_lo, const int32_t *np_exact, const double *xinfo, int NCONV,
                     double *out, void *stream);

/* ---- line-by-line absorption ---------------------------------------------------------------
 * Replaces add_line_set_monochromatic_absorption (archnemesis/LineData_0.py:279-358) with the
 * Voigt profile of lineshape/voigt_impl/voigt_scipy.py:8-52 (SciPy voigt_profile = Re w(z)).
 *   nu, sw, e_lower, stim_ref [N]; broadening[3*M, N] rows (gamma, n, delta) per molecule;
 *   mix[M]; pt[NPT,3] = (t_calc, p_calc, q_ratio) per state point; wn_grid[NWAVE] ascending.
 * out[NPT,NWAVE] is ACCUMULATED into (like the reference). shape_id: 0 Voigt, 1 Lorentz, 2 Gaussian. */
int ansb200_lbl_absorption(const double *wn_grid, int NWAVE, const double *nu, const double *sw,
                           const double *e_lower, const double *stim_ref, const double *broadening,
                           int N, const double *mix, int M, const double *pt, int NPT, double t_ref,
                           double p_ref, double abundance, double mass, double s_floor,
                           double wn_calc_window, double wn_approx_window, int shape_id, double *out,
                           void *stream);

/* Voigt profile exactly as the LBL kernel evaluates it (device arrays of length n; diagnostics). */
int ansb200_voigt(const double *dwn, const double *alpha_d, const double *gamma_l, int n, double *out,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ANSB200_H */
