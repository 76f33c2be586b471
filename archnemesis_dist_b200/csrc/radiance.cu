// radiance.cu -- path radiance and layer-space Jacobian, g-integrated, for every path at once.
//
// Reference: calculate_layer_opacity (archnemesis/ForwardModel_0.py:3989-4012: continuum add, LAYINC
// gather, SCALE), the gas-gradient scatter of calculate_gaseous_line_opacity (:3868-3872),
// calc_thermal_emission_spectrum (:6287-6377), calc_thermal_emission_spectrumg (:6380-6504) with the
// unit scaling of :4244-4247, calculate_transmission_spectrum (:4104-4129) and the g-integration
// of CIRSrad (:4504-4508).
//
// One CTA owns one (wavenumber, path).  Phase 0 tabulates the Planck function of the path's
// layers.  Phase 1: one warp per g-ordinate scans the path: lanes own contiguous layer chunks,
// a multiplicative (gradient form, tr = trold*exp(-tau_j), :6446-6448) or additive (no-gradient
// form, tr = exp(-sum tau), :6345-6346) warp scan gives the transmission to each layer, a reverse
// scan gives the suffix sums.  The reference's O(NLAYIN^2 * NPAR) recurrence (:6455-6476) collapses to
//     d spec / d q[k,j] = dtau[k,j] * D_j  +  [k == NVMR] * (T_{j-1} - T_j) * dB_j/dT
//     D_j = T_j B_j - sum_{m>j} (T_{m-1} - T_m) B_m - T_N * radground
// so only D[g][j] and the temperature term are kept (shared memory); the 5-D
// (NWAVE,NG,NPAR,NLAYIN,NPATH) tensor of the reference never exists.  Phase 2: one thread per
// (parameter, layer) pair forms dtau on the fly from dk / dtaucon, multiplies, applies xfac and
// integrates over g with DELG, then nan_to_num.
#include <float.h>
#include "common.cuh"

constexpr int RAD_THREADS = 256;
constexpr int RAD_WARPS = RAD_THREADS / 32;
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

__global__ void __launch_bounds__(RAD_THREADS)
ans_radiance_kernel(RadParams P)
{
    extern __shared__ __align__(16) unsigned char rad_smem[];
    const int iw = blockIdx.x, ipath = blockIdx.y;
    const int NG = P.NG, NLAY = P.NLAY, NLM = P.NLAYMAX, NPATH = P.NPATH, NPAR = P.NPAR;
    const bool grad = (P.flags & ANSB200_RAD_GRAD) != 0;
    const bool thermal = (P.mode == 0);
    const int n = P.nlayin[ipath];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

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
    int *slay = reinterpret_cast<int *>(sdts + NG);      // [NLM]

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
        const double *taug = P.tau + ((size_t)iw * NG + ig) * NLAY;
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
            if (grad) for (int j = lane; j < n; j += 32) sD[(size_t)ig * NLM + j] = -sg;
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
                sD[(size_t)ig * NLM + j] = Tj * sB[j] - suffix;
                sTT[(size_t)ig * NLM + j] = (Tjm - Tj) * sdB[j];
                suffix += (Tjm - Tj) * sB[j];
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // g-integration of the spectrum (:4504) and dTSURF (:4508)
    if (threadIdx.x == 0) {
        double s = 0.0, d = 0.0;
        for (int ig = 0; ig < NG; ++ig) { s += sspec[ig] * P.delg[ig]; if (thermal) d += sdts[ig] * P.delg[ig]; }
        P.spec[(size_t)iw * NPATH + ipath] = s;
        if (grad && thermal && P.dtsurf) P.dtsurf[(size_t)iw * NPATH + ipath] = d;
    }
    if (!grad) return;

    // Phase 2: layer-space gradients, g-integrated.  thread -> (j, k) with k fastest so that the
    // NGAS+1 columns of one dk row are read by neighbouring threads.
    const int NP1 = P.NGAS + 1;
    double *out = P.dspec + ((size_t)iw * NPATH + ipath) * NPAR * NLM;
    for (int e = threadIdx.x; e < NPAR * NLM; e += RAD_THREADS) {
        const int j = e / NPAR, k = e - j * NPAR;
        double acc = 0.0;
        if (j < n) {
            const int l = slay[j];
            // which dk column feeds parameter k (last matching gas wins, like the reference's loop :3868-3872)
            int col = -1;
            double unit = 1.0;
            if (P.dk) {   // no active gas (NGAS == 0): dTAUGAS is all zeros in the reference (:3883-3888)
                if (k == P.NVMR) col = P.NGAS;
                else for (int i = 0; i < P.NGAS; ++i) if (P.gas_slot[i] == k) { col = i; unit = 1.0e-4; }
            }
            const double dcon = P.dtaucon ? P.dtaucon[((size_t)iw * NPAR + k) * NLAY + l] : 0.0;
            const double sc = sscale[j];
            for (int ig = 0; ig < NG; ++ig) {
                double dgas = 0.0;
                if (col >= 0) dgas = P.dk[(((size_t)iw * NG + ig) * NLAY + l) * NP1 + col] * unit;
                const double tmp = (dgas + dcon) * sc;
                double v = tmp * sD[(size_t)ig * NLM + j];
                if (thermal && k == P.NVMR) v += sTT[(size_t)ig * NLM + j];
                if (thermal) v *= xf;   // transmission: xfac already inside spec_g
                acc += v * P.delg[ig];
            }
            if (P.flags & ANSB200_RAD_NAN_TO_NUM) acc = ans_nan_to_num(acc);
        }
        out[(size_t)k * NLM + j] = acc;
    }
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
    size_t nd = (size_t)4 * NLAYMAX + (size_t)RAD_WARPS * (NLAYMAX + 1) + (grad ? (size_t)NG * NLAYMAX : 0) +
                ((grad && thermal) ? (size_t)NG * NLAYMAX : 0) + 2 * (size_t)NG;
    size_t smem = nd * 8 + (size_t)NLAYMAX * 4 + 16;
    ANS_REQUIRE(smem <= 227 * 1024, "radiance: NG*NLAYMAX too large for shared memory (%zu bytes)", smem);
    if (smem > 48 * 1024)
        ANS_CUDA_CHECK(cudaFuncSetAttribute(ans_radiance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(NWAVE, NPATH);
    ans_radiance_kernel<<<grid, RAD_THREADS, smem, stream>>>(P);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
