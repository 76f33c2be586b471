// kinterp.cu -- stand-alone k-table interpolation (calc_k / calc_kg), HBM-bound streaming kernel.
//
// Data layout: the resident table is plane-major [NP*NT][NWAVE][NG][NGAS] (kinterp.cuh), the outputs
// k[wave,g,:,:] are one contiguous NLAY*NGAS run per (wave,g) pair.  A thread owns fixed (layer,gas)
// outputs -- its bracket plane offsets and weights sit in registers for the whole kernel -- and the
// CTA walks a contiguous range of (wave,g) pairs, KI_UNROLL pairs at a time so that 4*KI_UNROLL
// independent corner loads are in flight per thread before the first exp.  Stores are fully
// coalesced; consecutive pairs read consecutive 48-byte gas vectors of the same planes, so every
// DRAM burst is used in full.  ln K is pre-tabulated (api.cu): each output costs four loads and one exp.
#include "kinterp.cuh"

constexpr int KI_THREADS = 640;
constexpr int KI_EPT = 1;        // outputs per thread per pair: covers NLAY*NGAS <= 640 in one pass
constexpr int KI_UNROLL = 4;

template <bool GRAD>
__global__ void __launch_bounds__(KI_THREADS, 3)
ans_kinterp_kernel(AnsTab T, AnsLayerPlan plan, int npairs,
                   int NP, int NT, int NGAS, int NLAY, double *__restrict__ kout, double *__restrict__ dkout)
{
    const size_t plane = (size_t)npairs * NGAS;          // elements per (p,T) plane
    const int per_pair = NLAY * NGAS;
    const int ppb = (npairs + gridDim.x - 1) / gridDim.x;   // contiguous pairs per CTA
    const int pbeg = blockIdx.x * ppb, pend = min(npairs, pbeg + ppb);
    for (int e0 = 0; e0 < per_pair; e0 += KI_THREADS * KI_EPT) {
        // this thread's outputs in the pass: e0 + threadIdx.x + q*KI_THREADS
        size_t off[KI_EPT];
        double w0[KI_EPT], w1[KI_EPT], w2[KI_EPT], w3[KI_EPT], omv[KI_EPT], vv[KI_EPT], dudt[KI_EPT];
        bool live[KI_EPT];
#pragma unroll
        for (int q = 0; q < KI_EPT; ++q) {
            const int e = e0 + threadIdx.x + q * KI_THREADS;
            live[q] = e < per_pair;
            const int l = live[q] ? e / NGAS : 0;
            const int gas = live[q] ? e - l * NGAS : 0;
            off[q] = (size_t)(plan.ip_lo[l] * NT + plan.it_lo[l]) * plane + gas;
            w0[q] = plan.w4[4 * l]; w1[q] = plan.w4[4 * l + 1]; w2[q] = plan.w4[4 * l + 2]; w3[q] = plan.w4[4 * l + 3];
            omv[q] = GRAD ? plan.omv[l] : 0.0; vv[q] = GRAD ? plan.vv[l] : 0.0; dudt[q] = GRAD ? plan.dudt[l] : 0.0;
        }
        for (int pair0 = pbeg; pair0 < pend; pair0 += KI_UNROLL) {
#pragma unroll
            for (int u = 0; u < KI_UNROLL; ++u) {
                const int pair = pair0 + u;
                if (pair >= pend) break;
                const size_t tbase = (size_t)pair * NGAS;
                const size_t obase = (size_t)pair * per_pair + e0 + threadIdx.x;
#pragma unroll
                for (int q = 0; q < KI_EPT; ++q) {
                    if (!live[q]) continue;
                    double kv, dv = 0.0;
                    ans_kinterp_elem<GRAD>(T, tbase + off[q], NT, plane, w0[q], w1[q], w2[q], w3[q], omv[q], vv[q],
                                           dudt[q], kv, dv);
                    kout[obase + q * KI_THREADS] = kv;
                    if (GRAD) dkout[obase + q * KI_THREADS] = dv;
                }
            }
        }
    }
}

extern "C" int ansb200_kinterp(const ansb200_table *t, int NLAY, const int32_t *ip_lo, const int32_t *it_lo,
                               const double *w4, const double *omv, const double *vv, const double *dudt,
                               int want_grad, double *k, double *dkdT, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ANS_REQUIRE(t && ip_lo && it_lo && w4 && k, "kinterp: null pointer");
    ANS_REQUIRE(NLAY > 0, "kinterp: NLAY must be positive");
    ANS_REQUIRE(!want_grad || (omv && vv && dudt && dkdT), "kinterp: gradient requested without omv/vv/dudt/dkdT");
    AnsLayerPlan plan{ip_lo, it_lo, w4, omv, vv, dudt};
    const int npairs = t->NWAVE * t->NG;
    int grid = ans_div_up(npairs, KI_UNROLL);
    if (grid > 148 * 12) grid = 148 * 12;      // persistent-style: a multiple of the SM count, pairs strided over CTAs
    if (want_grad)
        ans_kinterp_kernel<true><<<grid, KI_THREADS, 0, stream>>>(ans_tab(t), plan, npairs, t->NP, t->NT, t->NGAS,
                                                                  NLAY, k, dkdT);
    else
        ans_kinterp_kernel<false><<<grid, KI_THREADS, 0, stream>>>(ans_tab(t), plan, npairs, t->NP, t->NT, t->NGAS,
                                                                   NLAY, k, nullptr);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
