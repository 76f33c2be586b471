// lbl.cu -- line-by-line absorption cross-sections (Voigt core, 1/dnu^2 wings) on sm_100a.
//
// Reference: archnemesis/LineData_0.py:123-226 (line_strength, doppler_width, lorentz_width,
// line_shift), :228-277 (add_line_set_monochromatic_spectrum), :279-358 (..._absorption) and the
// Voigt profile of archnemesis/lineshape/voigt_impl/voigt_scipy.py:8-52, which is
// scipy.special.voigt_profile = Re w(z) / (sigma sqrt(2 pi)), z = (x + i gamma)/(sigma sqrt 2).
// SciPy evaluates w(z) with S. G. Johnson's Faddeeva package (continued fraction for large |z|,
// Zaghloul-Ali "Algorithm 916" sums otherwise); the same published algorithm is restated here
// for the real part so that the profile agrees with SciPy to a few ulp.
//
// Work decomposition: one CTA owns LBL_PTS consecutive wavenumbers of one (p,T) state point; the
// line list streams through shared memory in tiles, each thread deriving the (p,T)-dependent
// parameters of one line per tile (strength, Doppler sigma, Lorentz width, shifted centre, wing
// constant).  Every thread then walks the tile in line order for its own grid points, so the
// accumulation order per grid point is the reference's (line-major).  FP64-pipe bound; no HBM
// traffic to speak of (the line list is read once per CTA from L2).
#include "common.cuh"

constexpr int LBL_THREADS = 256;
constexpr int LBL_GP = 4;                         // grid points per thread
constexpr int LBL_PTS = LBL_THREADS * LBL_GP;     // grid points per CTA
constexpr int LBL_TILE = 256;                     // lines per shared-memory tile

__device__ __forceinline__ double ans_sinc(double x, double sinx)
{
    return fabs(x) < 1e-4 ? 1 - (0.1666666666666666666667) * x * x : sinx / x;
}

__device__ __forceinline__ double ans_sinh_taylor(double x)
{
    return x * (1 + (x * x) * (0.1666666666666666666667 + 0.00833333333333333333333 * (x * x)));
}

// 1/x for normal, finite x (the pair loop: squared distances and |z|^4-sized denominators): the hardware seed
// (>= 20 bits) and one cubic step, <= 1 ulp, without the special-case handling and final rounding fix-up of
// __drcp_rn (a third of its instructions).
__device__ __forceinline__ double ans_rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);            // |e| <= 2^-20 from the seed
    return fma(r, fma(e, e, e), r);              // r (1 + e + e^2): cubic step, relative error e^3 + rounding
}

// Re w(x + i y) for y >= 0 (the Voigt function), following the published Faddeeva-package
// algorithm at relerr = DBL_EPSILON.
__device__ double ans_faddeeva_re(double xin, double y)
{
    const double a = 0.518321480430085929872;    // pi / sqrt(-log(eps*0.5))
    const double c = 0.329973702884629072537;    // (2/pi) * a
    const double a2 = 0.268657157075235951582;   // a^2
    const double relerr = 2.220446049250313e-16;
    const double ispi = 0.56418958354775628694807945156;   // 1/sqrt(pi)
    const double x = fabs(xin);
    const double ya = fabs(y);
    if (xin == 0.0) return erfcx(y);
    if (y == 0.0) return exp(-x * x);

    if (ya > 7 || (x > 6 && (ya > 0.1 || (x > 8 && ya > 1e-10) || x > 28))) {
        // continued fraction; Re w is even in x, evaluate at |x|
        const double xs = x;
        if (x + ya > 4000) {
            if (x + ya > 1e7) {                      // nu == 1
                if (x > ya) {
                    const double yax = ya / xs;
                    const double denom = ispi / (xs + yax * ya);
                    return denom * yax;
                } else if (isinf(ya)) {
                    return isnan(x) ? NAN : 0.0;
                } else {
                    const double xya = xs / ya;
                    return ispi / (xya * xs + ya);
                }
            }
            const double dr = xs * xs - ya * ya - 0.5, di = 2 * xs * ya;   // nu == 2
            const double denom = ispi / (dr * dr + di * di);
            return denom * (xs * di - ya * dr);
        }
        const double c0 = 3.9, c1 = 11.398, c2 = 0.08254, c3 = 0.1421, c4 = 0.2023;
        double nu = floor(c0 + c1 / (c2 * x + c3 * ya + c4));
        double wr = xs, wi = ya;
        // |w|^2 >= x^2 + y^2 > 36 here: the reciprocal helper applies (<= 1 ulp per step of a contracting recurrence)
        for (nu = 0.5 * (nu - 1); nu > 0.4; nu -= 0.5) {
            const double denom = nu * ans_rcp_fast(wr * wr + wi * wi);
            wr = xs - wr * denom;
            wi = ya + wi * denom;
        }
        const double denom = ispi * ans_rcp_fast(wr * wr + wi * wi);
        return denom * wi;
    }

    double sum1 = 0, sum2 = 0, sum3 = 0;
    double ret;
    if (x < 10) {
        double prod2ax = 1, prodm2ax = 1;
        double expx2;
        if (isnan(y)) return y;
        if (x < 5e-4) {
            const double x2 = x * x;
            expx2 = 1 - x2 * (1 - 0.5 * x2);
            const double ax2 = 1.036642960860171859744 * x;   // 2*a*x
            const double exp2ax = 1 + ax2 * (1 + ax2 * (0.5 + 0.166666666666666666667 * ax2));
            const double expm2ax = 1 - ax2 * (1 - ax2 * (0.5 - 0.166666666666666666667 * ax2));
            for (int n = 1; n < 200; ++n) {
                const double coef = exp(-a2 * ((double)n * n)) * expx2 / (a2 * ((double)n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                if (coef * prod2ax < relerr * sum3) break;
            }
        } else {
            expx2 = exp(-x * x);
            const double exp2ax = exp((2 * a) * x), expm2ax = 1 / exp2ax;
            double sum5 = 0;
            for (int n = 1; n < 200; ++n) {
                const double coef = exp(-a2 * ((double)n * n)) * expx2 / (a2 * ((double)n * n) + y * y);
                prod2ax *= exp2ax;
                prodm2ax *= expm2ax;
                sum1 += coef;
                sum2 += coef * prodm2ax;
                sum3 += coef * prod2ax;
                sum5 += (coef * prod2ax) * (a * n);
                if ((coef * prod2ax) * (a * n) < relerr * sum5) break;
            }
        }
        const double expx2erfcxy = y > -6 ? expx2 * erfcx(y) : 2 * exp(y * y - x * x);
        if (y > 5) {
            const double sinxy = sin(x * y);
            ret = (expx2erfcxy - c * y * sum1) * cos(2 * x * y) + (c * x * expx2) * sinxy * ans_sinc(x * y, sinxy);
        } else {
            const double sinxy = sin(x * y);
            const double cos2xy = cos(2 * x * y);
            const double coef1 = expx2erfcxy - c * y * sum1;
            const double coef2 = c * x * expx2;
            ret = coef1 * cos2xy + coef2 * sinxy * ans_sinc(x * y, sinxy);
        }
    } else {
        if (isnan(x)) return x;
        if (isnan(y)) return y;
        ret = exp(-x * x);
        const double n0 = floor(x / a + 0.5);
        const double dx = a * n0 - x;
        sum3 = exp(-dx * dx) / (a2 * (n0 * n0) + y * y);
        double sum5 = a * n0 * sum3;
        const double exp1 = exp(4 * a * dx);
        double exp1dn = 1;
        int dn;
        bool done = false;
        for (dn = 1; n0 - dn > 0; ++dn) {
            const double np = n0 + dn, nm = n0 - dn;
            double tp = exp(-(a * dn + dx) * (a * dn + dx));
            double tm = tp * (exp1dn *= exp1);
            tp /= (a2 * (np * np) + y * y);
            tm /= (a2 * (nm * nm) + y * y);
            sum3 += tp + tm;
            sum5 += a * (np * tp + nm * tm);
            if (a * (np * tp + nm * tm) < relerr * sum5) { done = true; break; }
        }
        while (!done) {
            const double np = n0 + dn++;
            const double tp = exp(-(a * dn + dx) * (a * dn + dx)) / (a2 * (np * np) + y * y);
            sum3 += tp;
            sum5 += a * np * tp;
            if (a * np * tp < relerr * sum5) break;
        }
    }
    return ret + (0.5 * c) * y * (sum2 + sum3);
}

// scipy.special.voigt_profile(x, sigma, gamma)
__device__ __forceinline__ double ans_voigt_profile(double x, double sigma, double gamma)
{
    const double INV_SQRT_2 = 0.707106781186547524401;
    const double SQRT_2PI = 2.5066282746310002416123552393401042;
    if (sigma == 0) {
        if (gamma == 0) {
            if (isnan(x)) return x;
            return x == 0 ? INFINITY : 0.0;
        }
        return gamma / 3.14159265358979323846 / (x * x + gamma * gamma);
    }
    if (gamma == 0) return 1 / SQRT_2PI / sigma * exp(-(x / sigma) * (x / sigma) / 2);
    const double zreal = x / sigma * INV_SQRT_2;
    const double zimag = gamma / sigma * INV_SQRT_2;
    return ans_faddeeva_re(zreal, zimag) / sigma / SQRT_2PI;
}

// lineshape dispatch: 0 Voigt (voigt_scipy.py:52), 1 Lorentz (lorentz.py), 2 Gaussian (gaussian.py)
__device__ __forceinline__ double ans_lineshape(int shape_id, double dwn, double alpha_d, double gamma_l)
{
    const double SQRT_2LOG2 = 1.1774100225154747;   // sqrt(2 ln 2)
    if (shape_id == 0) return ans_voigt_profile(dwn, alpha_d / SQRT_2LOG2, gamma_l);
    if (shape_id == 1) return gamma_l / (3.141592653589793 * (gamma_l * gamma_l + dwn * dwn));
    const double ln2 = 0.6931471805599453;
    return sqrt(ln2 / 3.141592653589793) / alpha_d * exp(-(dwn * dwn * ln2) / (alpha_d * alpha_d));
}

__global__ void ans_voigt_kernel(const double *dwn, const double *alpha_d, const double *gamma_l, int n, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ans_lineshape(0, dwn[i], alpha_d[i], gamma_l[i]);
}

struct LblParams {
    const double *wn;
    int NWAVE;
    const double *nu, *sw, *e_lower, *stim_ref, *broadening;
    int N;
    const double *mix;
    int M;
    const double *pt;
    int NPT;
    double t_ref, p_ref, abundance, mass, s_floor, calc_win, approx_win;
    int shape_id;
    double *out;
};

#define ANS_C2_CGS (2.99792458E10 * 6.62607015E-27 / 1.380649E-16)

// Voigt profile from per-line constants: inv_s2 = 1/(sigma sqrt 2), zimag = gamma/(sigma sqrt 2),
// vnorm = 1/(sigma sqrt(2 pi)).  Within the +-25 cm-1 core window almost every pair has |z| > 4000 (sigma is
// ~1e-3 cm-1), where the Faddeeva package's continued fraction is the closed form of its nu == 2 case:
// one reciprocal instead of the five divisions of voigt_profile + w(z).  Everything else goes through the
// general routine.  Differences from the division-by-division evaluation are a few ulp.
__device__ __forceinline__ double ans_voigt_far(double x, double zimag, double vnorm)
{
    const double ispi = 0.56418958354775628694807945156;
    const double dr = x * x - zimag * zimag - 0.5, di = 2 * x * zimag;
    const double denom = ispi * ans_rcp_fast(dr * dr + di * di);
    return denom * (x * di - zimag * dr) * vnorm;
}

__device__ __forceinline__ double ans_voigt_from_constants(double d, double inv_s2, double zimag, double vnorm)
{
    const double x = fabs(d * inv_s2);
    const double s = x + zimag;
    if (s > 4000.0 && s <= 1.0e7) return ans_voigt_far(x, zimag, vnorm);
    return ans_faddeeva_re(d * inv_s2, zimag) * vnorm;
}

__global__ void __launch_bounds__(LBL_THREADS)
ans_lbl_kernel(LblParams P)
{
    // per-line constants of the tile: shifted centre, abundance*strength, wing constant abundance*S*V(25)*25^2,
    // and either the Voigt constants (fast) or (alpha_D, gamma_L) for the general line-shape routine
    __shared__ double2 s_nw[LBL_TILE], s_a0[LBL_TILE], s_12[LBL_TILE];     // (nus, W), (A, c0), (c1, c2): 16-byte reads
    __shared__ unsigned char s_live[LBL_TILE];
    const int ipt = blockIdx.y;
    const double t_calc = P.pt[3 * ipt], p_calc = P.pt[3 * ipt + 1], q_ratio = P.pt[3 * ipt + 2];
    const int jbase = blockIdx.x * LBL_PTS;
    double wnj[LBL_GP], acc[LBL_GP];
    double *out = P.out + (size_t)ipt * P.NWAVE;
#pragma unroll
    for (int q = 0; q < LBL_GP; ++q) {
        const int j = jbase + threadIdx.x + q * LBL_THREADS;
        wnj[q] = j < P.NWAVE ? P.wn[j] : 0.0;
        acc[q] = j < P.NWAVE ? out[j] : 0.0;
    }
    const double wn_lo = P.wn[jbase];
    const double wn_hi = P.wn[min(jbase + LBL_PTS, P.NWAVE) - 1];

    const double c2 = ANS_C2_CGS;
    const double dconst = (1.0 / 2.99792458E10) * sqrt(2 * log(2.0) * 6.02214129E+23 * 1.380649E-16);
    const double boltz = c2 * (t_calc - P.t_ref) / (t_calc * P.t_ref);
    const double t_ratio = P.t_ref / t_calc, p_ratio = p_calc / P.p_ref;
    const double cw2 = P.calc_win * P.calc_win;   // wn_calc_window_max**2.
    const double SQRT_2LOG2 = 1.1774100225154747, INV_SQRT_2 = 0.707106781186547524401;
    const double SQRT_2PI = 2.5066282746310002416123552393401042;

    for (int i0 = 0; i0 < P.N; i0 += LBL_TILE) {
        const int i = i0 + threadIdx.x;
        // derive the state-dependent parameters of one line per thread (LineData_0.py:123-226)
        unsigned char live = 0;
        if (threadIdx.x < LBL_TILE) {
            double nus = 0, A = 0, W = 0, k0 = 0, k1 = 0, k2 = 0;
            if (i < P.N) {
                const double nui = P.nu[i];
                const double str = P.sw[i] * ((1 - exp(-c2 * nui / t_calc)) / P.stim_ref[i]) * exp(boltz * P.e_lower[i]) * q_ratio;
                double shift = 0, gl = 0;
                for (int m = 0; m < P.M; ++m) {
                    gl += pow(t_ratio, P.broadening[(size_t)(3 * m + 1) * P.N + i]) * P.broadening[(size_t)(3 * m) * P.N + i] *
                          P.mix[m] * p_ratio;
                    shift += (p_ratio * P.broadening[(size_t)(3 * m + 2) * P.N + i]) * P.mix[m];
                }
                const double ad = dconst * nui * sqrt(t_calc / P.mass);
                nus = nui + shift;
                const bool in_reach = !(str < P.s_floor) && (wn_lo - nus < P.approx_win) && !(wn_hi - nus < -P.approx_win);
                if (in_reach) {
                    A = P.abundance * str;
                    W = A * ans_lineshape(P.shape_id, P.calc_win, ad, gl) * cw2;
                    const double sigma = ad / SQRT_2LOG2;
                    if (P.shape_id == 0 && sigma > 0.0 && gl > 0.0) {
                        live = 1;                                  // Voigt from constants
                        k0 = INV_SQRT_2 / sigma;
                        k1 = gl / sigma * INV_SQRT_2;
                        k2 = 1.0 / sigma / SQRT_2PI;
                    } else {
                        live = 2;                                  // general line-shape routine
                        k0 = ad;
                        k1 = gl;
                    }
                    // The CTA spans wn_lo..wn_hi (2 cm-1 at a 0.002 cm-1 grid), so almost every line treats all of
                    // the CTA's points alike.  d = wn - nus is monotone in wn (also after rounding), so the two end
                    // points classify every point in between exactly as the per-pair tests would:
                    //   3  every point in the 1/dnu^2 wing (calc_win <= |d| < approx_win)
                    //   4  every point in the core window with |z| in the Faddeeva closed-form range
                    // Those tiles run branch-free loops below; anything else keeps the per-pair tests.
                    const double d_lo = wn_lo - nus, d_hi = wn_hi - nus;
                    if ((d_lo >= P.calc_win && d_hi < P.approx_win) || (d_hi < -P.calc_win && d_lo >= -P.approx_win)) {
                        live = 3;
                    } else if (live == 1 && -P.calc_win <= d_lo && d_hi < P.calc_win && (d_lo > 0.0 || d_hi < 0.0)) {
                        const double near = d_lo > 0.0 ? d_lo : -d_hi, far = d_lo > 0.0 ? d_hi : -d_lo;
                        if (fabs(near * k0) + k1 > 4000.0 && fabs(far * k0) + k1 <= 1.0e7) live = 4;
                    }
                }
            }
            s_live[threadIdx.x] = live;
            s_nw[threadIdx.x] = make_double2(nus, W);
            s_a0[threadIdx.x] = make_double2(A, k0);
            s_12[threadIdx.x] = make_double2(k1, k2);
        }
        const int any = __syncthreads_or(live ? 1 : 0);
        if (any) {
            const int cnt = min(LBL_TILE, P.N - i0);
            for (int li = 0; li < cnt; ++li) {
                const int kind = s_live[li];
                if (!kind) continue;
                const double2 nw = s_nw[li];
                const double nus = nw.x, W = nw.y;
                if (kind == 3) {
#pragma unroll
                    for (int q = 0; q < LBL_GP; ++q) {
                        const double d = wnj[q] - nus;
                        acc[q] += W * ans_rcp_fast(d * d);
                    }
                    continue;
                }
                const double2 a0 = s_a0[li], c12 = s_12[li];
                const double A = a0.x, k0 = a0.y, k1 = c12.x, k2 = c12.y;
                if (kind == 4) {
#pragma unroll
                    for (int q = 0; q < LBL_GP; ++q) {
                        const double d = wnj[q] - nus;
                        acc[q] += A * ans_voigt_far(fabs(d * k0), k1, k2);
                    }
                    continue;
                }
#pragma unroll
                for (int q = 0; q < LBL_GP; ++q) {
                    const double d = wnj[q] - nus;
                    if (d >= P.approx_win || d < -P.approx_win) continue;
                    if (-P.calc_win <= d && d < P.calc_win)
                        acc[q] += A * (kind != 2 ? ans_voigt_from_constants(d, k0, k1, k2) : ans_lineshape(P.shape_id, d, k0, k1));
                    else
                        acc[q] += W * ans_rcp_fast(d * d);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < LBL_GP; ++q) {
        const int j = jbase + threadIdx.x + q * LBL_THREADS;
        if (j < P.NWAVE) out[j] = acc[q];
    }
}

extern "C" int ansb200_lbl_absorption(const double *wn_grid, int NWAVE, const double *nu, const double *sw,
                                      const double *e_lower, const double *stim_ref, const double *broadening, int N,
                                      const double *mix, int M, const double *pt, int NPT, double t_ref, double p_ref,
                                      double abundance, double mass, double s_floor, double wn_calc_window,
                                      double wn_approx_window, int shape_id, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(wn_grid && nu && sw && e_lower && stim_ref && broadening && mix && pt && out, "lbl_absorption: null pointer");
    ANS_REQUIRE(NWAVE > 0 && N >= 0 && M > 0 && NPT > 0, "lbl_absorption: bad shape");
    ANS_REQUIRE(shape_id >= 0 && shape_id <= 2, "lbl_absorption: unknown line shape %d", shape_id);
    ANS_REQUIRE(NPT <= 65535, "lbl_absorption: at most 65535 state points per call");
    if (N == 0) return ANSB200_OK;
    LblParams P{wn_grid, NWAVE, nu, sw, e_lower, stim_ref, broadening, N, mix, M, pt, NPT, t_ref, p_ref, abundance,
                mass, s_floor, wn_calc_window, wn_approx_window, shape_id, out};
    dim3 grid(ans_div_up(NWAVE, LBL_PTS), NPT);
    ans_lbl_kernel<<<grid, LBL_THREADS, 0, stream>>>(P);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

extern "C" int ansb200_voigt(const double *dwn, const double *alpha_d, const double *gamma_l, int n, double *out,
                             void *stream_)
{
    ANS_REQUIRE(dwn && alpha_d && gamma_l && out && n >= 0, "voigt: bad argument");
    if (n == 0) return ANSB200_OK;
    ans_voigt_kernel<<<ans_div_up(n, 256), 256, 0, (cudaStream_t)stream_>>>(dwn, alpha_d, gamma_l, n, out);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
