// convolve.cu -- instrument line shape applied to the spectrum and its state-vector Jacobian on the device.
//
// Reference: Measurement_0.conv / convg for k-tables (archnemesis/Measurement_0.py:2288-2465, :2467-2692):
//   FWHM == 0   scipy.interpolate.interp1d(Wave, y, axis=0)(VCONV)  (:2632-2640) -- np.interp's slope form
//               for the 1-D spectrum, SciPy's two-weight form for the 2-D gradients;
//   FWHM <  0   filter-weighted mean sum(f1*y)/sum(f1) over the calculation points under the filter,
//               f1 = np.interp(Wave[i], VFIL, AFIL), f1 > 0 only, in ascending order (:2642-2690).
// Both are a sparse operator on the wavenumber axis that the host builds once per geometry
// (plan.conv_operator, with the reference's own arithmetic for the weights); this kernel applies it to
// in[NWAVE, NCOL] = [spectrum | Jacobian columns] so that only [NCONV, NCOL] goes back to the host.
// One thread per (convolution point, column): consecutive threads read consecutive columns; products and
// sums are kept un-fused and in the reference's order, so the result is bit-identical to it.
#include "common.cuh"

constexpr int CV_THREADS = 64;

__global__ void __launch_bounds__(CV_THREADS)
ans_convolve_kernel(const double *__restrict__ in, int NCOL, int ld, int mode, int col0_np,
                    const int32_t *__restrict__ row_start, const int32_t *__restrict__ widx,
                    const double *__restrict__ wval, const double *__restrict__ norm,
                    const int32_t *__restrict__ np_lo, const int32_t *__restrict__ np_exact,
                    const double *__restrict__ xinfo, double *__restrict__ out)
{
    const int c = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= NCOL) return;
    double r;
    if (mode == 0 && (col0_np == 2 || (col0_np && col == 0))) {
        // np.interp: y[j] on a knot, else slope*(x - x[j]) + y[j]
        const int j = np_lo[c];
        const double y0 = in[(size_t)j * ld + col];
        if (np_exact[c]) {
            r = y0;
        } else {
            const double x_lo = xinfo[3 * c], x_hi = xinfo[3 * c + 1], x_new = xinfo[3 * c + 2];
            const double slope = __ddiv_rn(__dsub_rn(in[(size_t)(j + 1) * ld + col], y0), __dsub_rn(x_hi, x_lo));
            r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x_new, x_lo)), y0);
        }
    } else {
        // the sum keeps the reference's order; the loads do not depend on it, so eight rows are fetched ahead of
        // the adds (an analytic line shape spans hundreds of calculation points: a chain of dependent L2 round
        // trips otherwise)
        double acc = 0.0;
        int e = row_start[c];
        const int e1 = row_start[c + 1];
        for (; e + 8 <= e1; e += 8) {
            double w[8], v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                w[q] = __ldg(wval + e + q);
                v[q] = __ldg(in + (size_t)__ldg(widx + e + q) * ld + col);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc = __dadd_rn(acc, __dmul_rn(w[q], v[q]));
        }
        for (; e < e1; ++e) acc = __dadd_rn(acc, __dmul_rn(wval[e], in[(size_t)widx[e] * ld + col]));
        r = mode == 1 ? __ddiv_rn(acc, norm[c]) : acc;
    }
    out[(size_t)c * NCOL + col] = r;
}

extern "C" int ansb200_convolve(const double *in, int NWAVE, int NCOL, int ld, int mode, int col0_np_interp,
                                const int32_t *row_start, const int32_t *widx, const double *wval,
                                const double *norm, const int32_t *np_lo, const int32_t *np_exact,
                                const double *xinfo, int NCONV, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(in && row_start && widx && wval && out, "convolve: null pointer");
    ANS_REQUIRE(mode == 0 || mode == 1, "convolve: mode must be 0 (interpolation) or 1 (filter mean)");
    ANS_REQUIRE(mode != 1 || norm, "convolve: the filter mean needs norm");
    ANS_REQUIRE(!(mode == 0 && col0_np_interp) || (np_lo && np_exact && xinfo), "convolve: np.interp form needs np_lo/np_exact/xinfo");
    ANS_REQUIRE(NWAVE > 0 && NCOL > 0 && ld >= NCOL && NCONV > 0 && NCONV <= 65535, "convolve: bad shape");
    dim3 grid(ans_div_up(NCOL, CV_THREADS), NCONV);
    ans_convolve_kernel<<<grid, CV_THREADS, 0, stream>>>(in, NCOL, ld, mode, col0_np_interp, row_start, widx, wval, norm, np_lo,
                                                  np_exact, xinfo, out);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
