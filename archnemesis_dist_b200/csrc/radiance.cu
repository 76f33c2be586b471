// radiance.cu -- path radiance and layer-space Jacobian, g-integrated, for every path at once.
//
// Reference: calculate_layer_opacity (archnemesis/ForwardModel_0.py:3989-4012: continuum add, LAYINC
// gather, SCALE), the gas-gradient scatter of calculate_gaseous_line_opacity (:3868-3872),
// calc_thermal_emission_spectrum (:6287-6377), calc_thermal_emission_spectrumg (:6380-6504) with the
// unit scaling of :4244-4247, calculate_transmission_spectrum (:4104-4129) and the g-integration
// of CIRSrad (:4504-4508).
//
// One CTA owns one wavenumber and a subset of the paths (every NPG-th path), which it walks in sequence.
// With several paths per CTA (limb / occultation: NPATH = 16..64) the wavenumber's tau and dk slabs
// (NG*NLAY*(NGAS+2) doubles, 128 KB at 20 x 100 x 8) are staged in shared memory once and reused by every
// path instead of being re-read through L2 (57 GB for 64 paths x 4000 wavenumbers).  Phase 0 tabulates the
// Planck function of the path's
// layers.  Phase 1: one warp per g-ordinate scans the path: lanes own contiguous layer chunks,
// a multiplicative (gradient form, tr = trold*exp(-tau_j), :6446-6448) or additive (no-gradient
// form, tr = exp(-sum tau), :6345-6346) warp scan gives the transmission to each layer, a reverse
// scan gives the suffix sums.  The reference's O(NLAYIN^2 * NPAR) recurrence (:6455-6476) collapses to
//     d spec / d q[k,j] = dtau[k,j] * D_j  +  [k == NVMR] * (T_{j-1} - T_j) * dB_j/dT
//     D_j = T_j B_j - sum_{m>j} (T_{m-1} - T_m) B_m - T_N * radground
// so only D[g][j] and the temperature term are kept (shared memory); the 5-D
// (NWAVE,NG,NPAR,NLAYIN,NPATH) tensor of the reference never exists.  Phase 2: a warp per
// (parameter, 32 path layers) forms dtau on the fly from dk / dtaucon against the pre-weighted
// W[g][j] = D_j(g) SCALE_j DELG_g xfac, integrates over g, then nan_to_num; stores are contiguous in j.
#include <float.h>
#include "common.cuh"

constexpr int RAD_MAX_THREADS = 640;   // phase 1 is one warp per g-ordinate: NG warps (<= 20) when a CTA walks several paths
                                       // with staged slabs (one CTA per SM), half as many for one path per CTA
constexpr unsigned RFULL = 0xffffffffu;

struct RadParams {
    int mode;
    unsigned flags;
    const double *tau, *dk;
    const int32_t *gas_slot;
    const double *taucia, *taudust, *tauray, *dtaucon;
    const int32_t *layinc;
    const double *scale;
    const int32_t *nlayin;
    const double *emtemp, *laypress, *wave, *delg, *emissivity, *xfac, *solflux, *reflectance, *sol_ang, *emiss_ang;
    int ispace;
    double tsurf;
    int NWAVE, NG, NLAY, NGAS, NVMR, NPAR, NLAYMAX, NPATH;
    int NPG;      // path groups: CTA (iw, pg) walks paths pg, pg+NPG, ...
    int stage;    // tau / dk slabs of the wavenumber staged in shared memory (several paths per CTA)
    double *spec, *dspec, *dtsurf;
};

__device__ __forceinline__ double rshfl_up(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(RFULL, lo, d);
    hi = __shfl_up_sync(RFULL, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double rshfl_down(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(RFULL, lo, d);
    hi = __shfl_down_sync(RFULL, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double rshfl_idx(double v, int l)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(RFULL, lo, l);
    hi = __shfl_sync(RFULL, hi, l);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int lo = __double2loint(v), hi = __double2hiint(v);
        lo = __shfl_xor_sync(RFULL, lo, d);
        hi = __shfl_xor_sync(RFULL, hi, d);
        v += __hiloint2double(hi, lo);
    }
    return v;
}

// planck / planckg, archnemesis/ForwardModel_0.py:6183-6283 (c1, c2 as written there)
__device__ __forceinline__ void ans_planckg(int ispace, double wave, double temp, double &bb, double &dbdt)
{
    const double c1 = 1.1911e-12, c2 = 1.439;
    double y, a, ap;
    if (ispace == 0) {
        y = wave;
        a = c1 * (y * y * y);
        ap = c1 * c2 * ((y * y) * (y * y)) / (temp * temp);
    } else {
        y = 1.0e4 / wave;
        const double y2 = y * y;
        a = c1 * (y2 * y2 * y) / 1.0e4;
        ap = c1 * c2 * (y2 * y2 * y2) / 1.0e4 / (temp * temp);
    }
    const double e = exp(c2 * y / temp);
    const double b = e - 1.0;
    bb = a / b;
    dbdt = (e * ap) / (b * b);
}

__device__ __forceinline__ double ans_nan_to_num(double v)
{
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? DBL_MAX : -DBL_MAX;
    return v;
}

__global__ void __launch_bounds__(RAD_MAX_THREADS)
ans_radiance_kernel(RadParams P)
{
    const int RAD_THREADS = blockDim.x, RAD_WARPS = blockDim.x >> 5;
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int iw = blockIdx.x / P.NPG, pg = blockIdx.x - iw * P.NPG;
    const int NG = P.NG, NLAY = P.NLAY, NLM = P.NLAYMAX, NPATH = P.NPATH, NPAR = P.NPAR;
    const bool grad = (P.flags & ANSB200_RAD_GRAD) != 0;
    const bool thermal = (P.mode == 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NP1 = P.NGAS + 1;

    // shared-memory carve-up
    double *sB = reinterpret_cast<double *>(rad_smem);   // [NLM] Planck
    double *sdB = sB + NLM;                               // [NLM] dB/dT
    double *sscale = sdB + NLM;                           // [NLM]
    double *scon = sscale + NLM;                          // [NLM] (unused slot kept for alignment)
    double *sT = scon + NLM;                              // [RAD_WARPS][NLM+1] per-warp transmissions
    double *sD = sT + (size_t)RAD_WARPS * (NLM + 1);      // [NG][NLM]   D_j(g)           (grad)
    double *sTT = sD + (grad ? (size_t)NG * NLM : 0);     // [NG][NLM]   (T_{j-1}-T_j) dB (grad, thermal)
    double *sspec = sTT + ((grad && thermal) ? (size_t)NG * NLM : 0);   // [NG] spec_g
    double *sdts = sspec + NG;                            // [NG] dtsurf_g
    double *sdelg = sdts + NG;                            // [NG] quadrature weights
    double *stau = sdelg + NG;                            // [NG*NLAY]      staged tau slab  (P.stage)
    double *sdk = stau + (P.stage ? (size_t)NG * NLAY : 0);                 // [NG*NLAY*NP1] staged dk slab (P.stage, grad)
    int *slay = reinterpret_cast<int *>(sdk + ((P.stage && grad && P.dk) ? (size_t)NG * NLAY * NP1 : 0));   // [NLM]
    int *scol = slay + NLM;                               // [NPAR] dk column feeding parameter k (-1: none)

    // once per CTA: quadrature weights, the dk column of every parameter (last matching gas wins, like the
    // reference's loop :3868-3872), and the wavenumber's tau / dk slabs when they are reused by several paths
    for (int ig = threadIdx.x; ig < NG; ig += RAD_THREADS) sdelg[ig] = P.delg[ig];
    if (grad) {
        for (int k = threadIdx.x; k < NPAR; k += RAD_THREADS) {
            int col = -1;
            if (P.dk) {   // no active gas (NGAS == 0): dTAUGAS is all zeros in the reference (:3883-3888)
                if (k == P.NVMR) col = P.NGAS;
                else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) col = i;
            }
            scol[k] = col;
        }
    }
    const double *tauw = P.tau + (size_t)iw * NG * NLAY;
    const double *dkw = P.dk ? P.dk + (size_t)iw * NG * NLAY * NP1 : nullptr;
    if (P.stage) {
        for (int t = threadIdx.x; t < NG * NLAY; t += RAD_THREADS) stau[t] = tauw[t];
        tauw = stau;
        if (grad && P.dk) {
            for (int t = threadIdx.x; t < NG * NLAY * NP1; t += RAD_THREADS) sdk[t] = dkw[t];
            dkw = sdk;
        }
    }
    __syncthreads();

    for (int ipath = pg; ipath < NPATH; ipath += P.NPG) {
    const int n = P.nlayin[ipath];
    const double wv = thermal ? P.wave[iw] : 0.0;
    const double xf = P.xfac ? P.xfac[iw] : 1.0;
    for (int j = threadIdx.x; j < n; j += RAD_THREADS) {
        const int l = P.layinc[(size_t)j * NPATH + ipath];
        slay[j] = l;
        sscale[j] = P.scale[(size_t)j * NPATH + ipath];
        if (thermal) {
            double bb, db;
            ans_planckg(P.ispace, wv, P.emtemp[(size_t)j * NPATH + ipath], bb, db);
            sB[j] = bb;
            sdB[j] = db;
        }
    }
    // limb / nadir test and ground term (:6353-6365, :6479-6494)
    double radground = 0.0, dradground = 0.0;
    bool ground = false;
    if (thermal && n > 0) {
        const int jh = n / 2 - 1;
        const double p1 = P.laypress[P.layinc[(size_t)(jh >= 0 ? jh : n - 1) * NPATH + ipath]];
        const double p2 = P.laypress[P.layinc[(size_t)(n - 1) * NPATH + ipath]];
        ground = p2 > p1;
        if (ground) {
            if (P.tsurf <= 0.0) {
                ans_planckg(P.ispace, wv, P.emtemp[(size_t)(n - 1) * NPATH + ipath], radground, dradground);
            } else {
                ans_planckg(P.ispace, wv, P.tsurf, radground, dradground);
                const double em = P.emissivity[iw];
                radground *= em;
                dradground *= em;
            }
        }
    }
    __syncthreads();

    const int CH = (n + 31) / 32;   // layers per lane
    double *myT = sT + (size_t)warp * (NLM + 1);   // myT[j+1] = transmission to the bottom of layer j, myT[0] = 1
    for (int ig = warp; ig < NG; ig += RAD_WARPS) {
        const double *taug = tauw + (size_t)ig * NLAY;
        const int j0 = lane * CH, j1 = min(n, j0 + CH);
        // local scan over this lane's chunk
        const bool prodform = thermal && grad;   // running product (gradient form) or running sum
        double loc = prodform ? 1.0 : 0.0;
        for (int j = j0; j < j1; ++j) {
            const int l = slay[j];
            double t = taug[l];
            if (P.taucia) t += P.taucia[(size_t)iw * NLAY + l];
            if (P.taudust) t += P.taudust[(size_t)iw * NLAY + l];
            if (P.tauray) t += P.tauray[(size_t)iw * NLAY + l];
            t *= sscale[j];
            if (prodform) loc *= exp(-t); else loc += t;
            myT[j + 1] = loc;
        }
        // exclusive warp scan of the chunk totals
        double incl = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = rshfl_up(incl, d);
            if (lane >= d) incl = prodform ? incl * up : incl + up;
        }
        double base = rshfl_up(incl, 1);
        if (lane == 0) base = prodform ? 1.0 : 0.0;
        const double total = rshfl_idx(incl, 31);   // product of all layers / total optical depth
        if (lane == 0) myT[0] = 1.0;
        for (int j = j0; j < j1; ++j) {
            const double v = myT[j + 1];
            myT[j + 1] = prodform ? base * v : (thermal ? exp(-(base + v)) : base + v);
        }
        __syncwarp();

        if (!thermal) {
            // transmission: spec_g = exp(-sum tau) [* xfac]; d/dq[k,j] = -spec_g * dtau[k,j]  (:4110-4126)
            const double sg = exp(-total) * xf;
            if (lane == 0) sspec[ig] = sg;
            // (weighted for phase 2: D_j(g) * SCALE_j * DELG_g; xfac is already inside spec_g)
            if (grad) for (int j = lane; j < n; j += 32) sD[(size_t)ig * NLM + j] = -sg * sscale[j] * sdelg[ig];
            __syncwarp();
            continue;
        }

        const double Tn = prodform ? total : exp(-total);
        // E_j = (T_{j-1} - T_j) B_j ; forward sum for the spectrum, suffix sums for the gradient
        double esum = 0.0;
        for (int j = j0; j < j1; ++j) esum += (myT[j] - myT[j + 1]) * sB[j];
        double specg = warp_sum(esum);
        if (ground) specg += Tn * radground;
        if (!grad && P.emiss_ang && P.sol_ang) {
            const double ea = P.emiss_ang[ipath], sa = P.sol_ang[ipath];
            if (ea < 90.0 && sa < 90.0) {   // :6368-6373
                const double mu = cos(ea / 180.0 * 3.141592653589793), mu0 = cos(sa / 180.0 * 3.141592653589793);
                specg += Tn * exp(-total * mu / mu0) * (P.solflux ? P.solflux[iw] : 0.0) *
                         (P.reflectance ? P.reflectance[iw] : 0.0);
            }
        }
        if (lane == 0) {
            sspec[ig] = specg * xf;
            sdts[ig] = ground ? Tn * dradground * xf : 0.0;
        }
        if (grad) {
            // reverse exclusive scan of chunk sums: S_j = sum_{m>j} E_m + T_N radground
            double rincl = esum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double dn = rshfl_down(rincl, d);
                if (lane + d < 32) rincl += dn;
            }
            double after = rshfl_down(rincl, 1);
            if (lane == 31) after = 0.0;
            double suffix = after + (ground ? Tn * radground : 0.0);
            for (int j = j1 - 1; j >= j0; --j) {
                const double Tj = myT[j + 1], Tjm = myT[j];
                // weighted for phase 2: D_j(g) * SCALE_j * DELG_g * xfac and (T_{j-1}-T_j) dB_j/dT * DELG_g * xfac
                sD[(size_t)ig * NLM + j] = (Tj * sB[j] - suffix) * sscale[j] * (sdelg[ig] * xf);
                sTT[(size_t)ig * NLM + j] = (Tjm - Tj) * sdB[j] * (sdelg[ig] * xf);
                suffix += (Tjm - Tj) * sB[j];
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // g-integration of the spectrum (:4504) and dTSURF (:4508): warp 0, lanes over g
    if (warp == 0) {
        double s = 0.0, d = 0.0;
        for (int ig = lane; ig < NG; ig += 32) { s += sspec[ig] * sdelg[ig]; if (thermal) d += sdts[ig] * sdelg[ig]; }
        s = warp_sum(s);
        d = warp_sum(d);
        if (lane == 0) {
            P.spec[(size_t)iw * NPATH + ipath] = s;
            if (grad && thermal && P.dtsurf) P.dtsurf[(size_t)iw * NPATH + ipath] = d;
        }
    }
    if (grad) {
    // Phase 2: layer-space gradients, g-integrated.  With W[g][j] = D_j(g) SCALE_j DELG_g xfac (phase 1),
    //   d spec / d q[k,j] = unit_k sum_g W[g][j] dk[g, l_j, col_k]  +  dtaucon[k, l_j] sum_g W[g][j]
    //                       (+ sum_g (T_{j-1}-T_j) dB_j/dT DELG_g xfac   for k = NVMR, thermal)
    // i.e. the reference's (dgas + dcon) * SCALE * D * xfac summed over g with DELG (:3993, :4012, :6455-6476,
    // :4244-4247, :4504-4507), regrouped so that the inner loop is two loads and one FMA.
    // A warp takes one parameter k and 32 consecutive path layers j: the stores are contiguous in j and no
    // per-element index division is needed.
    double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLM;
    const int nchunk = (NLM + 31) >> 5;
    if (NG == 1) {
        // One g-ordinate (line-by-line tables): an item is one dk load, one dtaucon load, two FMAs and a store, and a
        // CTA is a single warp, so the items' global loads would run one round trip after another.  Four items are
        // fetched before any of them is finished (same arithmetic as the general loop below with NG = 1).
        const int nitem = NPAR * nchunk;
        for (int item0 = warp * 4; item0 < nitem; item0 += RAD_WARPS * 4) {
            double dkv[4], dcv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int item = item0 + u;
                const int k = item / nchunk, j = ((item - k * nchunk) << 5) + lane;
                dkv[u] = 0.0;
                dcv[u] = 0.0;
                if (item < nitem && j < n) {
                    const int l = slay[j], col = scol[k];
                    if (col >= 0) dkv[u] = dkw[(size_t)l * NP1 + col];
                    if (P.dtaucon) dcv[u] = P.dtaucon[((size_t)iw * NPAR + k) * NLAY + l];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int item = item0 + u;
                const int k = item / nchunk, j = ((item - k * nchunk) << 5) + lane;
                if (item >= nitem || j >= NLM) continue;
                double acc = 0.0;
                if (j < n) {
                    const int col = scol[k];
                    const double wa = sD[j];
                    double wsum = 0.0;
                    if (col >= 0) {
                        const double a0 = fma(wa, dkv[u], 0.0);
                        wsum += wa;
                        wsum += 0.0;
                        acc = (a0 + 0.0) * (col < P.NGAS ? 1.0e-4 : 1.0);
                    } else if (P.dtaucon) {
                        wsum += wa;
                    }
                    if (P.dtaucon) acc = fma(dcv[u], wsum, acc);
                    if (thermal && k == P.NVMR) {
                        double wt = 0.0;
                        wt += sTT[j];
                        acc += wt;
                    }
                    if (P.flags & ANSB200_RAD_NAN_TO_NUM) acc = ans_nan_to_num(acc);
                }
                out[(size_t)k * NLM + j] = acc;
            }
        }
    } else
    for (int item = warp; item < NPAR * nchunk; item += RAD_WARPS) {
        const int k = item / nchunk, j = ((item - k * nchunk) << 5) + lane;
        if (j >= NLM) continue;
        double acc = 0.0;
        if (j < n) {
            const int l = slay[j];
            const int col = scol[k];
            const double *wp = sD + j;
            // issued before the g loop so that its latency overlaps the dk loads
            const double dcon = P.dtaucon ? P.dtaucon[((size_t)iw * NPAR + k) * NLAY + l] : 0.0;
            double wsum = 0.0;
            if (col >= 0) {
                const double *dkp = dkw + (size_t)l * NP1 + col;
                double a0 = 0.0, a1 = 0.0, w1 = 0.0;
                int ig = 0;
                for (; ig + 1 < NG; ig += 2) {
                    const double wa = wp[(size_t)ig * NLM], wb = wp[(size_t)(ig + 1) * NLM];
                    a0 = fma(wa, dkp[(size_t)ig * NLAY * NP1], a0);
                    a1 = fma(wb, dkp[(size_t)(ig + 1) * NLAY * NP1], a1);
                    wsum += wa;
                    w1 += wb;
                }
                if (ig < NG) {
                    const double wa = wp[(size_t)ig * NLM];
                    a0 = fma(wa, dkp[(size_t)ig * NLAY * NP1], a0);
                    wsum += wa;
                }
                wsum += w1;
                acc = (a0 + a1) * (col < P.NGAS ? 1.0e-4 : 1.0);
            } else if (P.dtaucon) {
                for (int ig = 0; ig < NG; ++ig) wsum += wp[(size_t)ig * NLM];
            }
            if (P.dtaucon) acc = fma(dcon, wsum, acc);
            if (thermal && k == P.NVMR) {
                double wt = 0.0;
                for (int ig = 0; ig < NG; ++ig) wt += sTT[(size_t)ig * NLM + j];
                acc += wt;
            }
            if (P.flags & ANSB200_RAD_NAN_TO_NUM) acc = ans_nan_to_num(acc);
        }
        out[(size_t)k * NLM + j] = acc;
    }
    }
    __syncthreads();   // the per-path arrays are reused by the CTA's next path
    }   // paths
}

// ---- transmission, many paths: one warp per path -----------------------------------------------------
// calculate_transmission_spectrum (:4104-4129) is spec_g = exp(-sum_j tau_g[l_j] SCALE_j) (* xfac) and
// d spec_g / d q[k,j] = -spec_g dtau_g[k, l_j] SCALE_j, so after the g-integration of CIRSrad (:4504-4507)
//     d spec / d q[k,j] = -SCALE_j ( unit_k sum_g c_g dk[g, l_j, col_k] + dtaucon[k, l_j] sum_g c_g ),  c_g = spec_g DELG_g:
// a path is described by NG numbers.  A CTA stages the wavenumber's tau / dk slabs and the continuum once and
// each warp then walks its own paths with no CTA barrier: lanes sum the path's opacity per g (phase 1), then
// sweep the (parameter, layer) outputs with consecutive lanes on consecutive layers (contiguous stores).
// 5e3 warp instructions per (wavenumber, path) instead of 2.7e4 in the general kernel.
__device__ __forceinline__ double tl_nan_to_num_fwd(double v)
{
    if ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) v = ans_nan_to_num(v);     // (rare: one integer test otherwise)
    return v;
}

constexpr int RADT_WARPS = 32;
constexpr int RADT_GROUP = 64;      // paths per pass of the layer-space gradient phase (8 tiles of 8 paths)

__global__ void __launch_bounds__(RADT_WARPS * 32)
ans_transmission_paths_kernel(RadParams P)
{
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int iw = blockIdx.x;
    const int NG = P.NG, NLAY = P.NLAY, NLM = P.NLAYMAX, NPATH = P.NPATH, NPAR = P.NPAR, NP1 = P.NGAS + 1;
    const bool grad = (P.flags & ANSB200_RAD_GRAD) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *stau = reinterpret_cast<double *>(rad_smem);          // [NG*NLAY] tau of the wavenumber
    double *scon = stau + (size_t)NG * NLAY;                       // [NLAY] continuum (cia + dust + rayleigh)
    double *sdelg = scon + NLAY;                                   // [NG]
    double *sc = sdelg + NG;                                       // [RADT_WARPS][NG] c_g of the warp's current path
    const bool layer_space = grad && (P.flags & ANSB200_RAD_LAYER_SPACE) != 0;
    double *sS = sc + (size_t)RADT_WARPS * NG;                     // [RADT_GROUP][NLAY] sum of SCALE over a layer's visits
    double *sC = sS + (layer_space ? (size_t)RADT_GROUP * NLAY : 0);      // [RADT_GROUP][NG] c_g of the group's paths
    double *sCs = sC + (layer_space ? (size_t)RADT_GROUP * NG : 0);       // [RADT_GROUP] sum_g c_g
    double *sdcon = sCs + (layer_space ? (size_t)RADT_GROUP : 0);         // [NPAR*NLAY] dtaucon of the wavenumber (grad)
    double *sdk = sdcon + ((grad && P.dtaucon) ? (size_t)NPAR * NLAY : 0);   // [NG*NLAY*NP1] dk of the wavenumber (grad)
    // (layer space: the dk operands of the tensor-core products come straight from global memory -- every row is
    // fetched once per CTA by the warp that owns its layer tile -- so the 112 KB slab is not staged and two CTAs fit)
    int *scol = reinterpret_cast<int *>(sdk + ((grad && P.dk && !layer_space) ? (size_t)NG * NLAY * NP1 : 0));   // [NPAR]
    const int nthr = blockDim.x;
    // NG >= 8: the lanes of a warp are the g-ordinates of its path (phase 1 below); the opacity is then kept
    // layer-major, stau[l][g] = tau + continuum, so that a warp reads one layer's row
    const bool lanes_g = NG >= 8;
    for (int l = threadIdx.x; l < NLAY; l += nthr) {
        double c = 0.0;                                            // TAUCIA + TAUDUST + TAURAY in the reference's order (:3989)
        if (P.taucia) c += P.taucia[(size_t)iw * NLAY + l];
        if (P.taudust) c += P.taudust[(size_t)iw * NLAY + l];
        if (P.tauray) c += P.tauray[(size_t)iw * NLAY + l];
        scon[l] = c;
    }
    if (lanes_g) {
        __syncthreads();
        for (int t = threadIdx.x; t < NG * NLAY; t += nthr) {
            const int g = t / NLAY, l = t - g * NLAY;
            stau[l * NG + g] = P.tau[(size_t)iw * NG * NLAY + t] + scon[l];
        }
    } else {
        for (int t = threadIdx.x; t < NG * NLAY; t += nthr) stau[t] = P.tau[(size_t)iw * NG * NLAY + t];
    }
    for (int g = threadIdx.x; g < NG; g += nthr) sdelg[g] = P.delg[g];
    if (grad) {
        if (P.dtaucon) for (int t = threadIdx.x; t < NPAR * NLAY; t += nthr) sdcon[t] = P.dtaucon[(size_t)iw * NPAR * NLAY + t];
        if (P.dk && !layer_space)
            for (int t = threadIdx.x; t < NG * NLAY * NP1; t += nthr) sdk[t] = P.dk[(size_t)iw * NG * NLAY * NP1 + t];
        for (int k = threadIdx.x; k < NPAR; k += nthr) {
            int col = -1;
            if (P.dk) {
                if (k == P.NVMR) col = P.NGAS;
                else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) col = i;
            }
            scol[k] = col;
        }
    }
    __syncthreads();
    const double xf = P.xfac ? P.xfac[iw] : 1.0;
    double *myc = sc + (size_t)warp * NG;
    // (layer-space gradients: the paths are taken in groups of RADT_GROUP -- phase 1 of every path of the group, then
    // the group's gradients as tensor-core products by the whole CTA; otherwise one pass over all paths)
    for (int gp0 = 0; gp0 < NPATH; gp0 += (layer_space ? RADT_GROUP : NPATH)) {
    const int gp1 = layer_space ? min(NPATH, gp0 + RADT_GROUP) : NPATH;
    if (layer_space) {
        for (int t = threadIdx.x; t < RADT_GROUP * NLAY; t += nthr) sS[t] = 0.0;
        for (int t = threadIdx.x; t < RADT_GROUP * NG; t += nthr) sC[t] = 0.0;
        for (int t = threadIdx.x; t < RADT_GROUP; t += nthr) sCs[t] = 0.0;
        __syncthreads();
    }
    for (int ipath = gp0 + warp; ipath < gp1; ipath += (nthr >> 5)) {
        const int n = P.nlayin[ipath];
        // phase 1: spec_g and c_g.  The lane's path layers (up to 8: NLAYMAX <= 256) stay in registers for all g.
        constexpr int RQ = 8;
        int lq[RQ];
        double sq[RQ], cq[RQ];
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const int j = lane + 32 * q;
            const bool ok = j < n;
            lq[q] = ok ? P.layinc[(size_t)j * NPATH + ipath] : 0;
            sq[q] = ok ? P.scale[(size_t)j * NPATH + ipath] : 0.0;
            cq[q] = scon[lq[q]];
        }
        double spec = 0.0, csum = 0.0;
        if (lanes_g) {
            // every lane walks the path for its own g-ordinate; the (layer, scale) pairs the lanes hold are handed
            // round by shuffles, the NG exponentials are one instruction stream
            for (int g0 = 0; g0 < NG; g0 += 32) {
                const int g = g0 + lane;
                const double *tl = stau + (g < NG ? g : 0);
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < RQ; ++q) {
                    const int cntq = min(32, n - 32 * q);            // (warp-uniform)
#pragma unroll 4
                    for (int jj = 0; jj < cntq; ++jj) {
                        const int l = __shfl_sync(RFULL, lq[q], jj);
                        const double sv = rshfl_idx(sq[q], jj);
                        t = fma(tl[l * NG], sv, t);
                    }
                }
                for (int j = 32 * RQ; j < n; ++j) {                  // (paths longer than 256 layers)
                    const int l = P.layinc[(size_t)j * NPATH + ipath];
                    t = fma(tl[l * NG], P.scale[(size_t)j * NPATH + ipath], t);
                }
                const double cg = g < NG ? exp(-t) * xf * sdelg[g] : 0.0;
                if (g < NG) myc[g] = cg;
                spec += warp_sum(cg);
            }
            csum = spec;
        } else
        for (int g = 0; g < NG; ++g) {
            const double *tg = stau + (size_t)g * NLAY;
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < RQ; ++q) t += (tg[lq[q]] + cq[q]) * sq[q];
            for (int j = lane + 32 * RQ; j < n; j += 32) {          // (paths longer than 256 layers)
                const int l = P.layinc[(size_t)j * NPATH + ipath];
                t += (tg[l] + scon[l]) * P.scale[(size_t)j * NPATH + ipath];
            }
            t = warp_sum(t);
            const double sg = exp(-t) * xf;
            const double cg = sg * sdelg[g];
            spec += cg;
            csum += cg;
            if (lane == 0) myc[g] = cg;
        }
        if (lane == 0) P.spec[(size_t)iw * NPATH + ipath] = spec;
        __syncwarp();
        if (layer_space) {
            // the path's NG numbers c_g, their sum, and the sum of SCALE over the visits of every layer (a limb path
            // meets a layer twice; the projection that follows is linear, so the visits are added)
            const int r = ipath - gp0;
            for (int g = lane; g < NG; g += 32) sC[(size_t)r * NG + g] = myc[g];
            if (lane == 0) sCs[r] = csum;
            for (int j = lane; j < n; j += 32)
                atomicAdd(&sS[(size_t)r * NLAY + P.layinc[(size_t)j * NPATH + ipath]], P.scale[(size_t)j * NPATH + ipath]);
        } else if (grad) {
            double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLM;
            for (int j0 = 0; j0 < NLM; j0 += 32) {
                const int j = j0 + lane;
                const bool live = j < n;
                const int l = live ? P.layinc[(size_t)j * NPATH + ipath] : 0;
                const double scl = live ? P.scale[(size_t)j * NPATH + ipath] : 0.0;
                for (int k = 0; k < NPAR; ++k) {
                    double v = 0.0;
                    if (live) {
                        const int col = scol[k];
                        double a = 0.0;
                        if (col >= 0) {
                            const double *dkp = sdk + (size_t)l * NP1 + col;
                            double a0 = 0.0, a1 = 0.0;
                            int g = 0;
                            for (; g + 1 < NG; g += 2) {
                                a0 = fma(myc[g], dkp[(size_t)g * NLAY * NP1], a0);
                                a1 = fma(myc[g + 1], dkp[(size_t)(g + 1) * NLAY * NP1], a1);
                            }
                            if (g < NG) a0 = fma(myc[g], dkp[(size_t)g * NLAY * NP1], a0);
                            a = (a0 + a1) * (col < P.NGAS ? 1.0e-4 : 1.0);
                        }
                        if (P.dtaucon) a = fma(sdcon[(size_t)k * NLAY + l], csum, a);
                        v = -(a * scl);
                        if (P.flags & ANSB200_RAD_NAN_TO_NUM) v = ans_nan_to_num(v);
                    }
                    if (j < NLM) out[(size_t)k * NLM + j] = v;
                }
            }
        }
        __syncwarp();
    }
    if (layer_space) {
        // Gradients of the group in layer space, dspec[NWAVE,NPATH,NPAR,NLAY]:
        //     d spec / d q[k,l] = -S[path][l] ( unit_k T_k[path][l] + csum[path] dtaucon[k,l] ),
        //     T_k = C[path][g] x dk[g][l][col_k]   -- a [paths x NG] x [NG x NLAY] product per dk column:
        // mma.m8n8k4 tiles of 8 paths x 8 layers (a = C[path][g], b = dk[g][l][col]); a warp takes (tile, parameter) units.
        __syncthreads();
        const int npt = (gp1 - gp0 + 7) >> 3, nlt = (NLAY + 7) >> 3;
        const int kk = lane & 3, mm = lane >> 2;
        // a warp takes (layer tile, parameter) pairs: the pair's B fragments (dk[g][l][col], up to KS k-steps) are
        // loaded once from global memory and stay in registers while the warp walks the path tiles of the group
        const int nwarps = nthr >> 5;
        constexpr int KS = 5;                                      // k-steps held in registers (NG <= 20); beyond that the operands are re-read
        const double *dkw = P.dk ? P.dk + (size_t)iw * NG * NLAY * NP1 : nullptr;
        for (int u = warp; u < nlt * NPAR; u += nwarps) {
            const int lt = u / NPAR, k = u - lt * NPAR;             // (parameters fastest: neighbouring warps share rows)
            const int col = scol[k];
            const int lb = lt * 8 + mm;                             // the layer of this lane's B element
            double bf[KS];
#pragma unroll
            for (int i = 0; i < KS; ++i) {
                const int g = 4 * i + kk;
                bf[i] = (col >= 0 && g < NG && lb < NLAY) ? dkw[((size_t)g * NLAY + lb) * NP1 + col] : 0.0;
            }
            const double unit_k = (col >= 0 && col < P.NGAS) ? 1.0e-4 : 1.0;
            const int l0 = lt * 8 + 2 * kk;                         // this lane's two output layers
            const double dc0 = (P.dtaucon && l0 < NLAY) ? sdcon[(size_t)k * NLAY + l0] : 0.0;
            const double dc1 = (P.dtaucon && l0 + 1 < NLAY) ? sdcon[(size_t)k * NLAY + l0 + 1] : 0.0;
            for (int pt = 0; pt < npt; ++pt) {
                const int r = pt * 8 + mm, path = gp0 + r;
                double d0 = 0.0, d1 = 0.0;
                if (col >= 0) {
                    const double *ar = sC + (size_t)r * NG + kk;
#pragma unroll
                    for (int i = 0; i < KS; ++i) {
                        if (4 * i < NG) {
                            const double av = 4 * i + kk < NG ? ar[4 * i] : 0.0;
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(d0), "+d"(d1) : "d"(av), "d"(bf[i]));
                        }
                    }
                    for (int g0 = 4 * KS; g0 < NG; g0 += 4) {
                        const int g = g0 + kk;
                        const double av = g < NG ? sC[(size_t)r * NG + g] : 0.0;
                        const double bv = (g < NG && lb < NLAY) ? dkw[((size_t)g * NLAY + lb) * NP1 + col] : 0.0;
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(d0), "+d"(d1) : "d"(av), "d"(bv));
                    }
                    d0 *= unit_k;
                    d1 *= unit_k;
                }
                if (path < gp1) {
                    const double cs = sCs[r];
                    const double *srow = sS + (size_t)r * NLAY;
                    double *out = P.dspec + (((size_t)iw * NPATH + path) * NPAR + k) * NLAY;
                    if (l0 < NLAY) {
                        const double scl = srow[l0];
                        double v = scl != 0.0 ? -(fma(dc0, cs, d0) * scl) : 0.0;
                        if (P.flags & ANSB200_RAD_NAN_TO_NUM) v = tl_nan_to_num_fwd(v);
                        out[l0] = v;
                    }
                    if (l0 + 1 < NLAY) {
                        const double scl = srow[l0 + 1];
                        double v = scl != 0.0 ? -(fma(dc1, cs, d1) * scl) : 0.0;
                        if (P.flags & ANSB200_RAD_NAN_TO_NUM) v = tl_nan_to_num_fwd(v);
                        out[l0 + 1] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
    }   // path groups
}

// ---- thermal emission with gradients, many paths: one warp per path ------------------------------------------
// Same arithmetic as phase 1 / phase 2 of ans_radiance_kernel (gradient form tr = trold*exp(-tau_j), closed-form
// D_j), reorganised so that a path never leaves its warp: lanes own contiguous chunks of up to TP_RQ path layers;
// for every g the warp scans the path (chunk products, warp scan, reverse scan of the emission terms) and each
// lane immediately adds W_j(g) dk[g, l_j, :] to the NGAS+1 running sums of its own layers, which stay in registers
// across the g loop (TP_RQ * (TP_NC+1) doubles per lane).  No per-path [NG][NLAYIN] scratch, no CTA barrier after
// the wavenumber's slabs are staged.  Used for NPATH >= 4, NLAYIN <= 32*TP_RQ, NGAS+1 <= TP_NC.
constexpr int TP_RQ = 7;       // path layers per lane (224 per path)
constexpr int TP_NC = 8;       // dk columns (NGAS + 1) kept per layer
constexpr int TP_WARPS = 8;

__global__ void __launch_bounds__(TP_WARPS * 32, 1)
ans_thermal_paths_kernel(RadParams P)
{
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int iw = blockIdx.x;
    const int NG = P.NG, NLAY = P.NLAY, NLM = P.NLAYMAX, NPATH = P.NPATH, NPAR = P.NPAR, NP1 = P.NGAS + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthr = blockDim.x;
    double *stau = reinterpret_cast<double *>(rad_smem);          // [NG*NLAY]
    double *scon = stau + (size_t)NG * NLAY;                       // [NLAY]
    double *sdelg = scon + NLAY;                                   // [NG]
    double *sdcon = sdelg + NG;                                    // [NPAR*NLAY]
    double *sdk = sdcon + (P.dtaucon ? (size_t)NPAR * NLAY : 0);   // [NG*NLAY*TP_NC] (columns padded to TP_NC)
    sdk += (sdk - stau) & 1;                                       // 16-byte aligned: rows are read as double2
    double *spath = sdk + (size_t)NG * NLAY * TP_NC;               // per warp: B[NLM], dB[NLM], scale[NLM]
    int *scol = reinterpret_cast<int *>(spath + (size_t)TP_WARPS * 3 * NLM);   // [NPAR]
    int *slayw = scol + NPAR;                                      // per warp: layer index [NLM]
    for (int t = threadIdx.x; t < NG * NLAY; t += nthr) stau[t] = P.tau[(size_t)iw * NG * NLAY + t];
    for (int l = threadIdx.x; l < NLAY; l += nthr) {
        double c = 0.0;
        if (P.taucia) c += P.taucia[(size_t)iw * NLAY + l];
        if (P.taudust) c += P.taudust[(size_t)iw * NLAY + l];
        if (P.tauray) c += P.tauray[(size_t)iw * NLAY + l];
        scon[l] = c;
    }
    for (int g = threadIdx.x; g < NG; g += nthr) sdelg[g] = P.delg[g];
    if (P.dtaucon) for (int t = threadIdx.x; t < NPAR * NLAY; t += nthr) sdcon[t] = P.dtaucon[(size_t)iw * NPAR * NLAY + t];
    for (int t = threadIdx.x; t < NG * NLAY * TP_NC; t += nthr) {
        const int c = t % TP_NC, gl = t / TP_NC;
        sdk[t] = c < NP1 ? P.dk[((size_t)iw * NG * NLAY + gl) * NP1 + c] : 0.0;
    }
    for (int k = threadIdx.x; k < NPAR; k += nthr) {
        int col = -1;
        if (k == P.NVMR) col = P.NGAS;
        else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) col = i;
        scol[k] = col;
    }
    __syncthreads();
    const double wv = P.wave[iw];
    const double xf = P.xfac ? P.xfac[iw] : 1.0;
    double *sBw = spath + (size_t)warp * 3 * NLM, *sdBw = sBw + NLM, *sscw = sdBw + NLM;
    int *slay = slayw + (size_t)warp * NLM;
    const int tcol = P.NGAS;      // dk column of the temperature parameter
    for (int ipath = warp; ipath < NPATH; ipath += (nthr >> 5)) {
        const int n = P.nlayin[ipath];
        for (int j = lane; j < n; j += 32) {
            slay[j] = P.layinc[(size_t)j * NPATH + ipath];
            sscw[j] = P.scale[(size_t)j * NPATH + ipath];
            double bb, db;
            ans_planckg(P.ispace, wv, P.emtemp[(size_t)j * NPATH + ipath], bb, db);
            sBw[j] = bb;
            sdBw[j] = db;
        }
        // limb / nadir test and ground term (:6353-6365, :6479-6494)
        double radground = 0.0, dradground = 0.0;
        bool ground = false;
        if (n > 0) {
            const int jh = n / 2 - 1;
            const double p1 = P.laypress[P.layinc[(size_t)(jh >= 0 ? jh : n - 1) * NPATH + ipath]];
            const double p2 = P.laypress[P.layinc[(size_t)(n - 1) * NPATH + ipath]];
            ground = p2 > p1;
            if (ground) {
                if (P.tsurf <= 0.0) {
                    ans_planckg(P.ispace, wv, P.emtemp[(size_t)(n - 1) * NPATH + ipath], radground, dradground);
                } else {
                    ans_planckg(P.ispace, wv, P.tsurf, radground, dradground);
                    const double em = P.emissivity[iw];
                    radground *= em;
                    dradground *= em;
                }
            }
        }
        __syncwarp();
        const int CH = (n + 31) / 32;
        const int j0 = lane * CH, cnt = max(0, min(n, j0 + CH) - j0);
        double acc[TP_RQ][TP_NC], wsum[TP_RQ];
#pragma unroll
        for (int q = 0; q < TP_RQ; ++q) {
            wsum[q] = 0.0;
#pragma unroll
            for (int c = 0; c < TP_NC; ++c) acc[q][c] = 0.0;
        }
        double spec = 0.0, dts = 0.0;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
            const double *tg = stau + (size_t)g * NLAY;
            // transmission to the bottom of each of the lane's layers: chunk products, then the warp scan
            double Tq[TP_RQ];
            double loc = 1.0;
#pragma unroll
            for (int q = 0; q < TP_RQ; ++q) {
                if (q < cnt) {
                    const int l = slay[j0 + q];
                    loc *= exp(-((tg[l] + scon[l]) * sscw[j0 + q]));
                }
                Tq[q] = loc;
            }
            double incl = loc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double up = rshfl_up(incl, d);
                if (lane >= d) incl *= up;
            }
            double base = rshfl_up(incl, 1);
            if (lane == 0) base = 1.0;
            const double Tn = rshfl_idx(incl, 31);
            // emission terms E_j = (T_{j-1} - T_j) B_j, their sum and suffix sums
            double esum = 0.0;
#pragma unroll
            for (int q = 0; q < TP_RQ; ++q) {
                if (q < cnt) {
                    Tq[q] *= base;
                    const double Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
                    esum += (Tm - Tq[q]) * sBw[j0 + q];
                }
            }
            double specg = warp_sum(esum);
            if (ground) specg += Tn * radground;
            const double dgx = sdelg[g] * xf;
            spec += specg * xf * sdelg[g];
            if (ground) dts += Tn * dradground * xf * sdelg[g];
            double rincl = esum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double dn = rshfl_down(rincl, d);
                if (lane + d < 32) rincl += dn;
            }
            double after = rshfl_down(rincl, 1);
            if (lane == 31) after = 0.0;
            double suffix = after + (ground ? Tn * radground : 0.0);
#pragma unroll
            for (int qq = 0; qq < TP_RQ; ++qq) {
                constexpr int last = TP_RQ - 1;
                const int q = last - qq;                 // (compile-time after unrolling: the lane's layers, last first)
                if (q < cnt) {
                    const int j = j0 + q;
                    const double Tj = Tq[q], Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
                    const double W = (Tj * sBw[j] - suffix) * sscw[j] * dgx;
                    const double TT = (Tm - Tj) * sdBw[j] * dgx;
                    suffix += (Tm - Tj) * sBw[j];
                    wsum[q] += W;
                    const double2 *dkp = reinterpret_cast<const double2 *>(sdk + ((size_t)g * NLAY + slay[j]) * TP_NC);
#pragma unroll
                    for (int c2 = 0; c2 < TP_NC / 2; ++c2) {
                        const double2 v = dkp[c2];
                        acc[q][2 * c2] = fma(W, v.x, acc[q][2 * c2]);
                        acc[q][2 * c2 + 1] = fma(W, v.y, acc[q][2 * c2 + 1]);
                    }
                    // (T_{j-1}-T_j) dB_j/dT belongs to the temperature parameter, whose dk column is NGAS
#pragma unroll
                    for (int c = 0; c < TP_NC; ++c) acc[q][c] += (c == tcol) ? TT : 0.0;
                }
            }
        }
        if (lane == 0) {
            P.spec[(size_t)iw * NPATH + ipath] = spec;
            if (P.dtsurf) P.dtsurf[(size_t)iw * NPATH + ipath] = dts;
        }
        // d spec / d q[k, j] = unit_k acc[j][col_k] + dtaucon[k, l_j] wsum[j]
        double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLM;
#pragma unroll 1
        for (int k = 0; k < NPAR; ++k) {
            const int col = scol[k];
            const double unit = (col >= 0 && col < P.NGAS) ? 1.0e-4 : 1.0;
#pragma unroll
            for (int q = 0; q < TP_RQ; ++q) {
                if (q < cnt) {
                    const int j = j0 + q;
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < TP_NC; ++c) v = (c == col) ? acc[q][c] * unit : v;
                    if (P.dtaucon) v = fma(sdcon[(size_t)k * NLAY + slay[j]], wsum[q], v);
                    if (P.flags & ANSB200_RAD_NAN_TO_NUM) v = ans_nan_to_num(v);
                    out[(size_t)k * NLM + j] = v;
                }
            }
            for (int j = n + lane; j < NLM; j += 32) out[(size_t)k * NLM + j] = 0.0;    // rows past NLAYIN
        }
        __syncwarp();
    }
}

// ---- thermal emission with layer-space gradients, many paths: tiles of 8 paths, g-sums on the tensor cores -------
// The gradient of a path with respect to the opacity parameters of LAYER l is sum_g sum_{visits j of l} W_j(g) dk[g,l,:]
// (+ the Planck-derivative term for the temperature), so per layer it is a small matrix product
//     G_l[path, c] = sum_g  Wl[path, g] * DK_l[g, c],     Wl[path, g] = sum over the path's visits of l of W_j(g)
// -- 8 paths x (NGAS + 2) columns x NG: one mma.m8n8k4 per 4 g-ordinates.  A CTA owns (wavenumber, 8 paths); warp p
// scans path p like ans_thermal_paths_kernel (lanes own contiguous chunks of visits, whose layer / scale / Planck
// value now live in REGISTERS), but instead of the NGAS+1 fused multiply-adds per visit and g it stores the one
// number W_j(g) into the CTA's W tile [path][g mod 4][layer]; after every 4 g-ordinates all warps fold the tile into
// G (shared memory) with FP64 tensor-core products whose B operand (dk, plus a column of ones that yields sum_g W for
// the continuum term) is read from global memory (prefetched into L1 while the scan runs).  Visits of the same layer
// by the same path (limb: two) are ranked once per path ("legs", last visit first) and stored in as many passes, the
// later ones read-modify-write, so no shared-memory floating-point atomics are needed.  The chunk length per lane is
// forced odd: lanes then hit distinct banks in the tau row and in the W tile.  The scan is branch-free: unused slots
// of a lane point at a zero opacity (exp(-0) = 1, Planck value 0) and exp(-x) is an inlined 13th-degree polynomial
// after the usual 2^k reduction (relative error < 3e-16; the library routine with its range branches costs 2.5 x
// as many instructions).  Output: dspec[NWAVE, NPATH, NPAR, NLAY].
constexpr int TL_PATHS = 8;      // paths per CTA = warps = rows of the product
constexpr int TL_RQ = 7;         // visits per lane: NLAYIN <= 224

__constant__ double TL_EXPC[12] = {      // 1/13!, 1/12!, ..., 1/2!
    1.6059043836821613e-10, 2.0876756987868098e-09, 2.5052108385441720e-08, 2.7557319223985893e-07,
    2.7557319223985888e-06, 2.4801587301587302e-05, 1.9841269841269841e-04, 1.3888888888888889e-03,
    8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 5.0000000000000000e-01};

// exp(y) for |y| <= 708 (the callers clamp): 2^k reduction, 13th-degree polynomial; NaN propagates
__device__ __forceinline__ double tl_exp_core(double y)
{
    const double t = fma(y, 1.4426950408889634, 6755399441055744.0);        // round(y log2 e) in the low word
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, y);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double pl = TL_EXPC[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) pl = fma(pl, r, TL_EXPC[i]);
    pl = fma(pl, r, 1.0);
    pl = fma(pl, r, 1.0);
    return __hiloint2double(__double2hiint(pl) + (k << 20), __double2loint(pl));
}
// exp(-x) for x >= 0 (beyond 708: 2^-1021 instead of a subnormal or 0)
__device__ __forceinline__ double tl_expneg(double x)
{
    return tl_exp_core(x > 708.0 ? -708.0 : -x);
}
// Planck function with the constants of ans_planckg: a = c1 y^3 (or c1 y^5 / 1e4), c2y = c2 y
__device__ __forceinline__ double tl_planck(double a, double c2y, double temp)
{
    const double x = c2y / temp;
    return a / (tl_exp_core(x > 708.0 ? 708.0 : x) - 1.0);
}
__device__ __forceinline__ void tl_planck_consts(int ispace, double wave, double &a, double &c2y)
{
    const double c1 = 1.1911e-12, c2 = 1.439;
    if (ispace == 0) {
        a = c1 * (wave * wave * wave);
        c2y = c2 * wave;
    } else {
        const double y = 1.0e4 / wave, y2 = y * y;
        a = c1 * (y2 * y2 * y) / 1.0e4;
        c2y = c2 * y;
    }
}

__device__ __forceinline__ double tl_nan_to_num(double v)
{
    if ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) v = ans_nan_to_num(v);     // (rare: one integer test otherwise)
    return v;
}

// One g-ordinate of one path, RQ visit slots per lane (RQ = the path's odd chunk length: no slot tests, the seven
// exponentials of a lane are independent instruction streams).  Adds the g-ordinate's share to spec / dts / esq and
// leaves W_j(g) in the path's row of the W tile.
template <int RQ>
__device__ __forceinline__ void tl_scan_g(const char *tg, char *wp, const int (&lay8)[TL_RQ], const double (&sc)[TL_RQ],
                                          const double (&Bq)[TL_RQ], double (&esq)[TL_RQ], unsigned m0, unsigned m1,
                                          unsigned long long legs, int maxleg, bool ground, double radground,
                                          double dradground, double dg, double xf, int lane, double &spec, double &dts)
{
    double Tq[RQ], Wq[RQ];
    double loc = 1.0;
#pragma unroll
    for (int q = 0; q < RQ; ++q) Tq[q] = tl_expneg(*reinterpret_cast<const double *>(tg + lay8[q]) * sc[q]);
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        loc *= Tq[q];
        Tq[q] = loc;
    }
    double incl = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = rshfl_up(incl, d);
        asm("{ .reg .pred p; setp.ge.s32 p, %2, %3; @p mul.f64 %0, %0, %1; }" : "+d"(incl) : "d"(up), "r"(lane), "r"(d));
    }
    double base = rshfl_up(incl, 1);
    if (lane == 0) base = 1.0;
    const double Tn = rshfl_idx(incl, 31);
    double esum = 0.0;
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        Tq[q] *= base;
        const double Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
        esum = fma(Tm - Tq[q], Bq[q], esum);
    }
    // suffix sums of the emission terms; lane 0's inclusive value is the path's sum
    double rincl = esum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double dn = rshfl_down(rincl, d);
        asm("{ .reg .pred p; setp.lt.s32 p, %2, %3; @p add.f64 %0, %0, %1; }" : "+d"(rincl) : "d"(dn), "r"(lane), "r"(32 - d));
    }
    double specg = rshfl_idx(rincl, 0);
    if (ground) specg += Tn * radground;
    const double dgx = dg * xf;
    spec += specg * xf * dg;
    if (ground) dts += Tn * dradground * xf * dg;
    double after = rshfl_down(rincl, 1);
    if (lane == 31) after = 0.0;
    double suffix = after + (ground ? Tn * radground : 0.0);
#pragma unroll
    for (int qq = 0; qq < RQ; ++qq) {
        const int q = RQ - 1 - qq;
        const double Tj = Tq[q], Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
        Wq[q] = (Tj * Bq[q] - suffix) * sc[q] * dgx;
        esq[q] = fma(Tm - Tj, dgx, esq[q]);
        suffix = fma(Tm - Tj, Bq[q], suffix);
    }
    // W_j(g) into the tile: one pass per leg, the later ones add to what the earlier ones stored
#pragma unroll
    for (int q = 0; q < RQ; ++q)
        if ((m0 >> q) & 1u) *reinterpret_cast<double *>(wp + lay8[q]) = Wq[q];
    __syncwarp();
    if (maxleg >= 1) {
#pragma unroll
        for (int q = 0; q < RQ; ++q)
            if ((m1 >> q) & 1u) *reinterpret_cast<double *>(wp + lay8[q]) += Wq[q];
        __syncwarp();
        for (int r = 2; r <= maxleg; ++r) {
#pragma unroll
            for (int q = 0; q < RQ; ++q)
                if ((int)((legs >> (8 * q)) & 0xffull) == r) *reinterpret_cast<double *>(wp + lay8[q]) += Wq[q];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(TL_PATHS * 32, 2)
ans_thermal_layers_kernel(RadParams P)
{
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int iw = blockIdx.x, tile = blockIdx.y;
    const int NG = P.NG, NLAY = P.NLAY, NPATH = P.NPATH, NPAR = P.NPAR, NP1 = P.NGAS + 1;
    const int NT = (NP1 + 1 + 7) / 8, GS = NT * 64 + 2;
    const int TS = NLAY + 1;              // tau row: [NLAY] + one zero for the unused slots of a lane
    const int LS = NLAY | 1;              // W row (odd: the tensor-core A fragment reads 32 rows at one layer)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sG = reinterpret_cast<double *>(rad_smem);            // [NLAY][GS]: G_l[path][column] (+ 2 pad)
    double *stau = sG + (size_t)NLAY * GS;                        // [NG][TS]: gas + continuum opacity
    double *sW = stau + (size_t)NG * TS;                          // [8 paths * 4 g][LS]
    double *sdelg = sW + (size_t)32 * LS;                         // [NG]
    int *sfirst = reinterpret_cast<int *>(sdelg + NG);            // per warp [NLAY]: scratch of the visit ranking
    int *scol = sfirst + TL_PATHS * NLAY;                         // [NPAR]: offset of the parameter's column in a G row (-1: none)
    int *svis = scol + NPAR;                                      // [NLAY]: layer crossed by any of the tile's paths
    for (int l0 = 0; l0 < NLAY; l0 += 32) {
        const int l = l0 + lane;
        double c = 0.0;
        if (l < NLAY) {
            if (P.taucia) c += P.taucia[(size_t)iw * NLAY + l];
            if (P.taudust) c += P.taudust[(size_t)iw * NLAY + l];
            if (P.tauray) c += P.tauray[(size_t)iw * NLAY + l];
            for (int g = warp; g < NG; g += TL_PATHS) stau[g * TS + l] = P.tau[((size_t)iw * NG + g) * NLAY + l] + c;
        }
    }
    for (int g = threadIdx.x; g < NG; g += blockDim.x) {
        stau[g * TS + NLAY] = 0.0;
        sdelg[g] = P.delg[g];
    }
    // (G is not cleared: the first chunk's products overwrite the rows of the layers in play, the others are never read)
    for (int t = threadIdx.x; t < 32 * LS; t += blockDim.x) sW[t] = 0.0;
    for (int t = threadIdx.x; t < TL_PATHS * NLAY; t += blockDim.x) sfirst[t] = -1;
    for (int t = threadIdx.x; t < NLAY; t += blockDim.x) svis[t] = 0;
    for (int k = threadIdx.x; k < NPAR; k += blockDim.x) {
        int col = -1;
        if (k == P.NVMR) col = P.NGAS;
        else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) col = i;
        scol[k] = col < 0 ? -1 : ((col >> 3) * 64 + (col & 7)) | (col < P.NGAS ? 0x40000000 : 0);   // bit 30: unit 1e-4
    }
    __syncthreads();
    // ---- the warp's path: visits into registers, ranking of repeated layers ---------------------------------------
    const int ipath = tile * TL_PATHS + warp;
    const int n = ipath < NPATH ? P.nlayin[ipath] : 0;
    const int CH = ((n + 31) / 32) | 1;
    const int j0 = lane * CH, cnt = max(0, min(n, j0 + CH) - j0);
    const double wv = P.wave[iw];
    const double xf = P.xfac ? P.xfac[iw] : 1.0;
    double pl_a, pl_c2y;
    tl_planck_consts(P.ispace, wv, pl_a, pl_c2y);
    int lay8[TL_RQ];                     // byte offset of the visit's layer in a tau / W row
    double sc[TL_RQ], Bq[TL_RQ], esq[TL_RQ];
#pragma unroll
    for (int q = 0; q < TL_RQ; ++q) {
        lay8[q] = NLAY * 8; sc[q] = 0.0; Bq[q] = 0.0; esq[q] = 0.0;
        if (q < cnt) {
            const size_t at = (size_t)(j0 + q) * NPATH + ipath;
            const int l = P.layinc[at];
            lay8[q] = l * 8;
            sc[q] = P.scale[at];
            Bq[q] = tl_planck(pl_a, pl_c2y, P.emtemp[at]);
            svis[l] = 1;
        }
    }
    unsigned long long legs = 0;         // rank of each visit among the path's visits of its layer, 8 bits per slot
    unsigned m0 = 0u, m1 = 0u;           // slots of rank 0 / rank 1
    int maxleg = -1;
    {
        int *first = sfirst + warp * NLAY;
        unsigned pend = (1u << cnt) - 1u;
        while (__any_sync(RFULL, pend != 0u)) {
            ++maxleg;
#pragma unroll
            for (int q = 0; q < TL_RQ; ++q) if ((pend >> q) & 1u) atomicMax(&first[lay8[q] >> 3], j0 + q);
            __syncwarp();
            unsigned done = 0u;
#pragma unroll
            for (int q = 0; q < TL_RQ; ++q)
                if (((pend >> q) & 1u) && first[lay8[q] >> 3] == j0 + q) {
                    legs |= (unsigned long long)maxleg << (8 * q);
                    done |= 1u << q;
                }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < TL_RQ; ++q) if ((done >> q) & 1u) first[lay8[q] >> 3] = -1;
            __syncwarp();
            if (maxleg == 0) m0 = done;
            if (maxleg == 1) m1 = done;
            pend &= ~done;
        }
        legs |= 0xff00000000000000ull;                      // (unused slots never match a rank)
#pragma unroll
        for (int q = 0; q < TL_RQ; ++q) if (q >= cnt) legs |= 0xffull << (8 * q);
    }
    // limb / nadir test and ground term (:6353-6365, :6479-6494)
    double radground = 0.0, dradground = 0.0;
    bool ground = false;
    if (n > 0) {
        const int jh = n / 2 - 1;
        const double p1 = P.laypress[P.layinc[(size_t)(jh >= 0 ? jh : n - 1) * NPATH + ipath]];
        const double p2 = P.laypress[P.layinc[(size_t)(n - 1) * NPATH + ipath]];
        ground = p2 > p1;
        if (ground) {
            if (P.tsurf <= 0.0) {
                ans_planckg(P.ispace, wv, P.emtemp[(size_t)(n - 1) * NPATH + ipath], radground, dradground);
            } else {
                ans_planckg(P.ispace, wv, P.tsurf, radground, dradground);
                const double em = P.emissivity[iw];
                radground *= em;
                dradground *= em;
            }
        }
    }
    double spec = 0.0, dts = 0.0;
    const int NCH = (NG + 3) / 4;
    // tensor-core products: the warp folds layers warp, warp + 8, ... -- those that any path of the tile crosses
    __syncthreads();
    unsigned vis = 0u;
    for (int i = 0; warp + TL_PATHS * i < NLAY; ++i) if (svis[warp + TL_PATHS * i]) vis |= 1u << i;
    const int t_lo = vis ? __ffs(vis) - 1 : 0, t_hi = vis ? 32 - __clz(vis) : 0;     // the warp's layers warp + 8 [t_lo, t_hi)
    const double *wl = sW + lane * LS + warp + TL_PATHS * t_lo;
#pragma unroll 1
    for (int gc = 0; gc < NCH; ++gc) {
        {   // the chunk's dk rows towards L1 / L2 while the scan runs
            const char *base = reinterpret_cast<const char *>(P.dk + ((size_t)iw * NG + gc * 4) * NLAY * NP1);
            const size_t bytes = (size_t)min(4, NG - gc * 4) * NLAY * NP1 * 8;
            for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(base + o));
        }
#pragma unroll 1
        for (int gi = 0; gi < 4; ++gi) {
            const int g = gc * 4 + gi;
            char *wp = reinterpret_cast<char *>(sW + (warp * 4 + gi) * LS);
            if (g < NG && n > 0) {
                const char *tg = reinterpret_cast<const char *>(stau + g * TS);
                const double dg = sdelg[g];
#define TL_SCAN(RQ) tl_scan_g<RQ>(tg, wp, lay8, sc, Bq, esq, m0, m1, legs, maxleg, ground, radground, dradground, dg, xf, \
                                  lane, spec, dts)
                if (CH == 7) TL_SCAN(7);
                else if (CH == 5) TL_SCAN(5);
                else if (CH == 3) TL_SCAN(3);
                else TL_SCAN(1);
#undef TL_SCAN
            } else if (n > 0) {
                // past the last g-ordinate (NG not a multiple of 4): the tile's column must not keep stale numbers
#pragma unroll
                for (int q = 0; q < TL_RQ; ++q) if (q < cnt) *reinterpret_cast<double *>(wp + lay8[q]) = 0.0;
            }
        }
        __syncthreads();
        // G_l (+)= Wl (8 paths x 4 g) * DK_l (4 g x 8 columns) for the warp's layers in play, four layers' operands in
        // flight (a layer inside the range that no path crosses has a zero W row)
        {
            const int g = gc * 4 + (lane & 3);
            const bool gok = g < NG;
            for (int nt = 0; nt < NT; ++nt) {
                const int c = nt * 8 + (lane >> 2);
                const bool ld = gok && c < NP1;
                const double bconst = (gok && c == NP1) ? 1.0 : 0.0;
                const double *db = P.dk + (((size_t)iw * NG + (gok ? g : 0)) * NLAY + warp + TL_PATHS * t_lo) * NP1 + (ld ? c : 0);
                double *gp = sG + (warp + TL_PATHS * t_lo) * GS + nt * 64 + (lane >> 2) * 8 + 2 * (lane & 3);
                const double *wa = wl;
                const int dstep = TL_PATHS * NP1, gstep = TL_PATHS * GS;
                int i = t_lo;
                for (; i + 4 <= t_hi; i += 4, wa += 4 * TL_PATHS, db += 4 * dstep, gp += 4 * gstep) {
                    double a[4], b[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[u] = wa[u * TL_PATHS];
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = ld ? db[u * dstep] : bconst;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double2 *dp = reinterpret_cast<double2 *>(gp + u * gstep);
                        double2 d = gc ? *dp : make_double2(0.0, 0.0);
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(d.x), "+d"(d.y) : "d"(a[u]), "d"(b[u]));
                        *dp = d;
                    }
                }
                for (; i < t_hi; ++i, wa += TL_PATHS, db += dstep, gp += gstep) {
                    const double a = wa[0], b = ld ? db[0] : bconst;
                    double2 *dp = reinterpret_cast<double2 *>(gp);
                    double2 d = gc ? *dp : make_double2(0.0, 0.0);
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
                    *dp = d;
                }
            }
        }
        __syncthreads();
    }
    if (lane == 0 && ipath < NPATH) {
        P.spec[(size_t)iw * NPATH + ipath] = spec;
        if (P.dtsurf) P.dtsurf[(size_t)iw * NPATH + ipath] = dts;
    }
    // (T_{j-1} - T_j) dB_j/dT, summed over g, belongs to the temperature parameter (dk column NGAS) of the visit's layer.
    // dB/dT from the Planck value already held: B = a / (e - 1)  =>  dB/dT = e ap / (e - 1)^2 = B (1 + B / a) c2 y / T^2
    {
        const int tcol = P.NGAS;
        const double inv_a = 1.0 / pl_a;
        double *gp = sG + (tcol >> 3) * 64 + warp * 8 + (tcol & 7);
        for (int r = 0; r <= maxleg; ++r) {
#pragma unroll
            for (int q = 0; q < TL_RQ; ++q) {
                if ((int)((legs >> (8 * q)) & 0xffull) == r) {
                    const double temp = P.emtemp[(size_t)(j0 + q) * NPATH + ipath];
                    const double db = Bq[q] * fma(Bq[q], inv_a, 1.0) * (pl_c2y / (temp * temp));
                    gp[(lay8[q] >> 3) * GS] += esq[q] * db;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // d spec / d q[k, l] = unit_k G_l[path, col_k] + dtaucon[k, l] G_l[path, ones]: warp p writes path p, a lane
    // keeps its (up to 4) layers' sum_g W and walks the parameters; layers the tile never crosses are zero
    if (ipath < NPATH) {
        const int one_off = (NP1 >> 3) * 64 + (NP1 & 7);
        const bool ntn = (P.flags & ANSB200_RAD_NAN_TO_NUM) != 0;
        for (int l0 = 0; l0 < NLAY; l0 += 128) {
            double ws[4];
            const double *grow[4];
            bool in[4], on[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int l = l0 + 32 * u + lane;
                in[u] = l < NLAY;
                on[u] = in[u] && svis[l] != 0;
                grow[u] = sG + (in[u] ? l : 0) * GS + warp * 8;
                ws[u] = on[u] ? grow[u][one_off] : 0.0;
            }
            double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLAY + l0 + lane;
            const double *dc = P.dtaucon ? P.dtaucon + (size_t)iw * NPAR * NLAY + l0 + lane : nullptr;
#pragma unroll 1
            for (int k = 0; k < NPAR; ++k, out += NLAY) {
                const int cw = scol[k];
                const int off = cw < 0 ? 0 : (cw & 0xffff);
                const double unit = cw < 0 ? 0.0 : ((cw & 0x40000000) ? 1.0e-4 : 1.0);
                double dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) dv[u] = (dc && in[u]) ? dc[k * NLAY + 32 * u] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (in[u]) {
                        double v = 0.0;
                        if (on[u]) v = cw < 0 ? dv[u] * ws[u] : fma(grow[u][off], unit, dv[u] * ws[u]);
                        if (ntn) v = tl_nan_to_num(v);
                        out[32 * u] = v;
                    }
                }
            }
        }
    }
}

// ---- thermal emission with gradients, one to three paths: a warp per (wavenumber, path) --------------------------
// The nadir case of an optimal-estimation iteration (config 2: one path of NLAY visits).  ans_radiance_kernel spends a
// CTA of NG/2 warps, two shared-memory passes and three CTA barriers on every wavenumber (36 000 warp instructions, issue
// slots 34 % busy, half its shared-memory wavefronts bank conflicts); here ONE warp does a wavenumber: lanes own
// contiguous chunks of visits whose layer / scale / Planck value / continuum sit in registers, the g loop scans the path
// (chunk products, warp scan, reverse scan of the emission terms -- the scan of ans_thermal_layers_kernel) and every
// lane adds W_j(g) dk[g, l_j, :] to its visits' NGAS+1 running sums.  A g-ordinate's operands of one wavenumber are two
// contiguous blocks (dk: NLAY*(NGAS+1) doubles, tau: NLAY) that the path uses exactly once: they stream through two
// buffers of the warp's own shared memory with cp.async, the next g-ordinate's while this one is scanned, so DRAM sees
// whole lines, the scan sees no load latency and the address arithmetic is 32-bit.  No CTA barrier.  The chunk
// length is odd (lanes read distinct banks of the staged rows); RQ = 1, 3, 5 or 7 visits per lane.
template <int RQ>
__device__ __forceinline__ void tn_path(const RadParams &P, int iw, int ipath, int lane, double *wbuf)
{
    const int NG = P.NG, NLAY = P.NLAY, NLM = P.NLAYMAX, NPATH = P.NPATH, NPAR = P.NPAR, NP1 = P.NGAS + 1;
    const int n = P.nlayin[ipath];
    const int CH = ((n + 31) / 32) | 1;        // odd: lanes then read distinct banks of the staged rows
    const int j0 = lane * CH, cnt = max(0, min(n, j0 + CH) - j0);
    const double wv = P.wave[iw];
    const double xf = P.xfac ? P.xfac[iw] : 1.0;
    double pl_a, pl_c2y;
    tl_planck_consts(P.ispace, wv, pl_a, pl_c2y);
    int lay[RQ];
    double sc[RQ], Bq[RQ], con[RQ], esq[RQ], wsum[RQ], acc[RQ][TP_NC];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        lay[q] = 0; sc[q] = 0.0; Bq[q] = 0.0; con[q] = 0.0; esq[q] = 0.0; wsum[q] = 0.0;
#pragma unroll
        for (int c = 0; c < TP_NC; ++c) acc[q][c] = 0.0;
        if (q < cnt) {
            const size_t at = (size_t)(j0 + q) * NPATH + ipath;
            const int l = P.layinc[at];
            lay[q] = l;
            sc[q] = P.scale[at];
            Bq[q] = tl_planck(pl_a, pl_c2y, P.emtemp[at]);
            double c = 0.0;                                    // TAUCIA + TAUDUST + TAURAY in the reference's order (:3989)
            if (P.taucia) c += P.taucia[(size_t)iw * NLAY + l];
            if (P.taudust) c += P.taudust[(size_t)iw * NLAY + l];
            if (P.tauray) c += P.tauray[(size_t)iw * NLAY + l];
            con[q] = c;
        }
    }
    // limb / nadir test and ground term (:6353-6365, :6479-6494)
    double radground = 0.0, dradground = 0.0;
    bool ground = false;
    if (n > 0) {
        const int jh = n / 2 - 1;
        const double p1 = P.laypress[P.layinc[(size_t)(jh >= 0 ? jh : n - 1) * NPATH + ipath]];
        const double p2 = P.laypress[P.layinc[(size_t)(n - 1) * NPATH + ipath]];
        ground = p2 > p1;
        if (ground) {
            if (P.tsurf <= 0.0) {
                ans_planckg(P.ispace, wv, P.emtemp[(size_t)(n - 1) * NPATH + ipath], radground, dradground);
            } else {
                ans_planckg(P.ispace, wv, P.tsurf, radground, dradground);
                const double em = P.emissivity[iw];
                radground *= em;
                dradground *= em;
            }
        }
    }
    double spec = 0.0, dts = 0.0;
    // A g-ordinate's operands of this wavenumber -- dk[g, :, :] (NLAY*(NGAS+1) doubles) and tau[g, :] -- are two
    // contiguous blocks that the path uses exactly once: they stream through two shared-memory buffers of the warp
    // (cp.async, the next g-ordinate while this one is scanned), so DRAM sees full lines and the scan sees no latency.
    const int BD = NLAY * NP1, BS = (BD + NLAY + 1) & ~1;       // doubles per buffer (16-byte multiple)
    const double *tw = P.tau + (size_t)iw * NG * NLAY;
    const double *dw = P.dk + (size_t)iw * NG * NLAY * NP1;
    const bool wide = ((BD | NLAY) & 1) == 0;                   // both blocks start 16-byte aligned for every g
    auto issue = [&](int g, double *dst) {
        const double *sd = dw + (size_t)g * BD, *st = tw + (size_t)g * NLAY;
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        if (wide) {
            for (int e = lane; e < (BD >> 1); e += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (unsigned)e * 16u), "l"(sd + 2 * e));
            for (int e = lane; e < (NLAY >> 1); e += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (unsigned)(BD + 2 * e) * 8u), "l"(st + 2 * e));
        } else {
            for (int e = lane; e < BD; e += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (unsigned)e * 8u), "l"(sd + e));
            for (int e = lane; e < NLAY; e += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (unsigned)(BD + e) * 8u), "l"(st + e));
        }
        asm volatile("cp.async.commit_group;");
    };
    int rowoff[RQ];                                             // the visit's dk row in a buffer
#pragma unroll
    for (int q = 0; q < RQ; ++q) rowoff[q] = lay[q] * NP1;
    issue(0, wbuf);
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
        const double *buf = wbuf + (g & 1) * BS;
        if (g + 1 < NG) {
            issue(g + 1, wbuf + ((g + 1) & 1) * BS);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncwarp();
        const double *tg = buf + BD;
        const double *dg_ = buf;
        double Tq[RQ];
        double loc = 1.0;
#pragma unroll
        for (int q = 0; q < RQ; ++q) Tq[q] = tl_expneg((tg[lay[q]] + con[q]) * sc[q]);
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            loc *= Tq[q];
            Tq[q] = loc;
        }
        double incl = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = rshfl_up(incl, d);
            asm("{ .reg .pred p; setp.ge.s32 p, %2, %3; @p mul.f64 %0, %0, %1; }" : "+d"(incl) : "d"(up), "r"(lane), "r"(d));
        }
        double base = rshfl_up(incl, 1);
        if (lane == 0) base = 1.0;
        const double Tn = rshfl_idx(incl, 31);
        double esum = 0.0;
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            Tq[q] *= base;
            const double Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
            esum = fma(Tm - Tq[q], Bq[q], esum);
        }
        double rincl = esum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double dn = rshfl_down(rincl, d);
            asm("{ .reg .pred p; setp.lt.s32 p, %2, %3; @p add.f64 %0, %0, %1; }" : "+d"(rincl) : "d"(dn), "r"(lane), "r"(32 - d));
        }
        double specg = rshfl_idx(rincl, 0);
        if (ground) specg += Tn * radground;
        const double dgv = P.delg[g];
        const double dgx = dgv * xf;
        spec += specg * xf * dgv;
        if (ground) dts += Tn * dradground * xf * dgv;
        double after = rshfl_down(rincl, 1);
        if (lane == 31) after = 0.0;
        double suffix = after + (ground ? Tn * radground : 0.0);
#pragma unroll
        for (int qq = 0; qq < RQ; ++qq) {
            const int q = RQ - 1 - qq;
            const double Tj = Tq[q], Tm = q == 0 ? base : Tq[q > 0 ? q - 1 : 0];
            const double W = (Tj * Bq[q] - suffix) * sc[q] * dgx;        // (unused slots: sc = 0)
            esq[q] = fma(Tm - Tj, dgx, esq[q]);
            suffix = fma(Tm - Tj, Bq[q], suffix);
            wsum[q] += W;
            const double *row = dg_ + rowoff[q];
#pragma unroll
            for (int c = 0; c < TP_NC; ++c)
                if (c < NP1) acc[q][c] = fma(W, row[c], acc[q][c]);
        }
        __syncwarp();          // (this buffer is refilled during the next iteration but one)
    }
    if (lane == 0) {
        P.spec[(size_t)iw * NPATH + ipath] = spec;
        if (P.dtsurf) P.dtsurf[(size_t)iw * NPATH + ipath] = dts;
    }
    // (T_{j-1} - T_j) dB_j/dT summed over g belongs to the temperature parameter (dk column NGAS);
    // dB/dT = B (1 + B / a) c2 y / T^2 from the Planck value already held
    const double inv_a = 1.0 / pl_a;
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        if (q < cnt) {
            const double temp = P.emtemp[(size_t)(j0 + q) * NPATH + ipath];
            const double db = Bq[q] * fma(Bq[q], inv_a, 1.0) * (pl_c2y / (temp * temp));
#pragma unroll
            for (int c = 0; c < TP_NC; ++c) acc[q][c] += (c == P.NGAS) ? esq[q] * db : 0.0;
        }
    }
    // d spec / d q[k, j] = unit_k acc[j][col_k] + dtaucon[k, l_j] wsum[j]
    double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLM;
    const bool ntn = (P.flags & ANSB200_RAD_NAN_TO_NUM) != 0;
#pragma unroll 1
    for (int k = 0; k < NPAR; ++k) {
        int col = -1;
        if (k == P.NVMR) col = P.NGAS;
        else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) col = i;
        const double unit = (col >= 0 && col < P.NGAS) ? 1.0e-4 : 1.0;
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            if (q < cnt) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < TP_NC; ++c) v = (c == col) ? acc[q][c] * unit : v;
                if (P.dtaucon) v = fma(P.dtaucon[((size_t)iw * NPAR + k) * NLAY + lay[q]], wsum[q], v);
                if (ntn) v = tl_nan_to_num(v);
                out[(size_t)k * NLM + j0 + q] = v;
            }
        }
        for (int j = n + lane; j < NLM; j += 32) out[(size_t)k * NLM + j] = 0.0;    // rows past NLAYIN
    }
}

constexpr int TN_WARPS = 4;

template <int RQ>
__global__ void __launch_bounds__(TN_WARPS * 32, RQ <= 3 ? 4 : (RQ <= 5 ? 3 : 2))
ans_thermal_nadir_kernel(RadParams P)          // RQ = (NLAYMAX / 32 rounded up) | 1: chosen by the launcher
{
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * TN_WARPS + warp;
    if (unit >= (long long)P.NWAVE * P.NPATH) return;
    const int iw = (int)(unit / P.NPATH), ipath = (int)(unit - (long long)iw * P.NPATH);
    const int BS = (P.NLAY * (P.NGAS + 1) + P.NLAY + 1) & ~1;
    tn_path<RQ>(P, iw, ipath, lane, reinterpret_cast<double *>(rad_smem) + (size_t)warp * 2 * BS);
}

static size_t thermal_nadir_smem(int NLAY, int NGAS)
{
    return (size_t)TN_WARPS * 2 * ((NLAY * (NGAS + 1) + NLAY + 1) & ~1) * 8;
}

static size_t thermal_layers_smem(int NG, int NLAY, int NGAS, int NPAR)
{
    const int NT = (NGAS + 2 + 7) / 8, GS = NT * 64 + 2;
    return ((size_t)NLAY * GS + (size_t)NG * (NLAY + 1) + (size_t)32 * (NLAY | 1) + NG) * 8 +
           ((size_t)TL_PATHS * NLAY + NPAR + NLAY) * 4 + 16;
}

extern "C" int ansb200_radiance_layer_space(int mode, unsigned flags, int NG, int NLAY, int NGAS, int NPAR, int NPATH,
                                            int NLAYMAX, int has_dk, int has_dtaucon)
{
    if (NPATH < 4 || !(flags & ANSB200_RAD_GRAD)) return 0;
    if (mode == 0)      // thermal emission: ans_thermal_layers_kernel
        return has_dk && NLAYMAX <= 32 * TL_RQ && NLAY <= 32 * TL_PATHS &&      // (a warp's layers fit a 32-bit mask)
                       thermal_layers_smem(NG, NLAY, NGAS, NPAR) <= 227 * 1024 ? 1 : 0;
    if (mode != 1) return 0;
    const size_t nd_t = (size_t)NG * NLAY + NLAY + NG + (size_t)32 * NG + (has_dtaucon ? (size_t)NPAR * NLAY : 0) +
                        (size_t)64 * (NLAY + NG + 1);
    return nd_t * 8 + (size_t)NPAR * 4 + 16 <= 227 * 1024 ? 1 : 0;
}

extern "C" int ansb200_radiance(int mode, unsigned flags, const double *tau, const double *dk, const int32_t *gas_slot,
                                const double *taucia, const double *taudust, const double *tauray,
                                const double *dtaucon, const int32_t *layinc, const double *scale,
                                const int32_t *nlayin, const double *emtemp, const double *laypress,
                                const double *wave, const double *delg, const double *emissivity, const double *xfac,
                                const double *solflux, const double *reflectance, const double *sol_ang,
                                const double *emiss_ang, int ispace, double tsurf, int NWAVE, int NG, int NLAY,
                                int NGAS, int NVMR, int NPAR, int NLAYMAX, int NPATH, double *spec, double *dspec,
                                double *dtsurf, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool grad = (flags & ANSB200_RAD_GRAD) != 0;
    ANS_REQUIRE(mode == 0 || mode == 1, "radiance: mode must be 0 (thermal) or 1 (transmission)");
    ANS_REQUIRE(tau && layinc && scale && nlayin && delg && spec, "radiance: null pointer");
    ANS_REQUIRE(mode != 0 || (emtemp && laypress && wave), "radiance: thermal mode needs emtemp/laypress/wave");
    ANS_REQUIRE(mode != 0 || tsurf <= 0.0 || emissivity, "radiance: TSURF > 0 needs emissivity");
    ANS_REQUIRE(!grad || (dspec && NPAR > 0), "radiance: gradients requested without dspec");
    ANS_REQUIRE(!grad || NGAS == 0 || (dk && gas_slot), "radiance: gradients requested without dk/gas_slot");
    ANS_REQUIRE(NWAVE > 0 && NG > 0 && NLAY > 0 && NLAYMAX > 0 && NPATH > 0, "radiance: bad shape");
    RadParams P{};
    P.mode = mode; P.flags = flags; P.tau = tau; P.dk = dk; P.gas_slot = gas_slot;
    P.taucia = taucia; P.taudust = taudust; P.tauray = tauray; P.dtaucon = dtaucon;
    P.layinc = layinc; P.scale = scale; P.nlayin = nlayin; P.emtemp = emtemp; P.laypress = laypress;
    P.wave = wave; P.delg = delg; P.emissivity = emissivity; P.xfac = xfac; P.solflux = solflux;
    P.reflectance = reflectance; P.sol_ang = sol_ang; P.emiss_ang = emiss_ang; P.ispace = ispace; P.tsurf = tsurf;
    P.NWAVE = NWAVE; P.NG = NG; P.NLAY = NLAY; P.NGAS = NGAS; P.NVMR = NVMR; P.NPAR = NPAR; P.NLAYMAX = NLAYMAX;
    P.NPATH = NPATH; P.spec = spec; P.dspec = dspec; P.dtsurf = dtsurf;
    const bool thermal = mode == 0;
    auto base_bytes = [&](int nthreads) {
        const size_t nd = (size_t)4 * NLAYMAX + (size_t)(nthreads / 32) * (NLAYMAX + 1) + (grad ? (size_t)NG * NLAYMAX : 0) +
                          ((grad && thermal) ? (size_t)NG * NLAYMAX : 0) + 3 * (size_t)NG;
        return nd * 8 + (size_t)(NLAYMAX + NPAR) * 4 + 16;
    };
    const int warps_min = NG == 1 ? 1 : 4;      // line-by-line tables: one g-ordinate, many small CTAs
    const int warps_n = NG < 4 ? warps_min : (NG > RAD_MAX_THREADS / 32 ? RAD_MAX_THREADS / 32 : NG);
    const int warps_1 = (warps_n + 1) / 2 < warps_min ? warps_min : (warps_n + 1) / 2;
    const int RAD_THREADS_1 = 32 * warps_1, RAD_THREADS_N = 32 * warps_n;
    size_t base = base_bytes(RAD_THREADS_1);
    const size_t slab = ((size_t)NG * NLAY + ((grad && dk) ? (size_t)NG * NLAY * (NGAS + 1) : 0)) * 8;
    ANS_REQUIRE(base <= 227 * 1024, "radiance: NG*NLAYMAX too large for shared memory (%zu bytes)", base);
    // Several paths: one CTA per (wavenumber, path group) with the wavenumber's slabs staged once, as many
    // groups as keep every SM busy (>= 4 paths per CTA so that the staging pays).  One path, or slabs that do
    // not fit: one CTA per (wavenumber, path), slabs read through L2 (consecutive CTAs share the wavenumber).
    const bool layer_space = (flags & ANSB200_RAD_LAYER_SPACE) != 0;
    ANS_REQUIRE(!layer_space || ansb200_radiance_layer_space(mode, flags, NG, NLAY, NGAS, NPAR, NPATH, NLAYMAX, dk != nullptr,
                                                             dtaucon != nullptr),
                "radiance: layer-space gradients need >= 4 paths, slabs that fit in shared memory and, for thermal emission, "
                "NLAYIN <= %d", 32 * TL_RQ);
    if (thermal && layer_space) {
        const size_t smem_l = thermal_layers_smem(NG, NLAY, NGAS, NPAR);
        ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_thermal_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(smem_l > 48 * 1024 ? smem_l : 48 * 1024)));
        const dim3 grid((unsigned)NWAVE, (unsigned)((NPATH + TL_PATHS - 1) / TL_PATHS));
        ans_thermal_layers_kernel<<<grid, TL_PATHS * 32, smem_l, stream>>>(P);
        ANS_LAUNCH_CHECK();
        return ANSB200_OK;
    }
    if (!thermal && NPATH >= 4) {
        // transmission with several paths: warp-per-path kernel if the wavenumber's slabs fit in shared memory
        const size_t nd_t = (size_t)NG * NLAY + NLAY + NG + (size_t)RADT_WARPS * NG + ((grad && dtaucon) ? (size_t)NPAR * NLAY : 0) +
                            ((grad && layer_space) ? (size_t)RADT_GROUP * (NLAY + NG + 1) : 0) +
                            ((grad && dk && !layer_space) ? (size_t)NG * NLAY * (NGAS + 1) : 0);
        const size_t smem_t = nd_t * 8 + (size_t)NPAR * 4 + 16;
        if (smem_t <= 227 * 1024) {
            ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_transmission_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_t > 48 * 1024 ? smem_t : 48 * 1024)));
            // layer space: 16 warps per CTA, so that two CTAs (64 registers, ~90 KB each) share an SM and one's
            // staging / barriers overlap the other's arithmetic; path space keeps the dk slab in shared memory (one CTA)
            const int wmax = (grad && layer_space) ? RADT_WARPS / 2 : RADT_WARPS;
            int warps = NPATH < wmax ? NPATH : wmax;
            ans_transmission_paths_kernel<<<(unsigned)NWAVE, warps * 32, smem_t, stream>>>(P);
            ANS_LAUNCH_CHECK();
            return ANSB200_OK;
        }
    }
    if (thermal && grad && dk && NPATH < 4 && NLAYMAX <= 32 * TL_RQ && NGAS + 1 <= TP_NC && NG >= 4 &&
        thermal_nadir_smem(NLAY, NGAS) <= 100 * 1024) {
        // thermal emission with gradients, one to three paths (the nadir case of a retrieval): a warp per (wavenumber,
        // path), a g-ordinate's operands streamed through the warp's own shared-memory buffers
        const long long units = (long long)NWAVE * NPATH;
        const unsigned nblk = (unsigned)((units + TN_WARPS - 1) / TN_WARPS);
        const size_t smem_n = thermal_nadir_smem(NLAY, NGAS);
        const int rq = ((NLAYMAX + 31) / 32) | 1;
#define TN_LAUNCH(RQ)                                                                                                  \
        do {                                                                                                           \
            ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_thermal_nadir_kernel<RQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)(smem_n > 48 * 1024 ? smem_n : 48 * 1024)));                     \
            ans_thermal_nadir_kernel<RQ><<<nblk, TN_WARPS * 32, smem_n, stream>>>(P);                                 \
        } while (0)
        if (rq <= 1) TN_LAUNCH(1);
        else if (rq <= 3) TN_LAUNCH(3);
        else if (rq <= 5) TN_LAUNCH(5);
        else TN_LAUNCH(7);
#undef TN_LAUNCH
        ANS_LAUNCH_CHECK();
        return ANSB200_OK;
    }
    if (thermal && grad && dk && NPATH >= 4 && NLAYMAX <= 32 * TP_RQ && NGAS + 1 <= TP_NC) {
        // thermal emission with gradients over several paths: warp-per-path kernel if the slabs fit
        const size_t nd_p = (size_t)NG * NLAY + NLAY + NG + (dtaucon ? (size_t)NPAR * NLAY : 0) +
                            (size_t)NG * NLAY * TP_NC + (size_t)TP_WARPS * 3 * NLAYMAX + 1;
        const size_t smem_p = nd_p * 8 + ((size_t)NPAR + (size_t)TP_WARPS * NLAYMAX) * 4 + 16;
        if (smem_p <= 227 * 1024) {
            ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_thermal_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(smem_p > 48 * 1024 ? smem_p : 48 * 1024)));
            const int warps = NPATH < TP_WARPS ? NPATH : TP_WARPS;
            ans_thermal_paths_kernel<<<(unsigned)NWAVE, warps * 32, smem_p, stream>>>(P);
            ANS_LAUNCH_CHECK();
            return ANSB200_OK;
        }
    }
    int NPG = NPATH, stage = 0, nthreads = RAD_THREADS_1;
    if (NPATH >= 4 && base_bytes(RAD_THREADS_N) + slab <= 227 * 1024) {
        stage = 1;
        nthreads = RAD_THREADS_N;
        base = base_bytes(RAD_THREADS_N);
        NPG = 1;
        while ((long long)NWAVE * NPG < 2 * 148 && NPATH / (NPG * 2) >= 4) NPG *= 2;
    }
    P.NPG = NPG; P.stage = stage;
    const size_t smem = base + (stage ? slab : 0);
    ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_radiance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    ANS_REQUIRE((long long)NWAVE * NPG < 2147483647LL, "radiance: NWAVE*NPATH too large");
    ans_radiance_kernel<<<(unsigned)(NWAVE * NPG), nthreads, smem, stream>>>(P);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
