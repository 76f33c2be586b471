// kinterp.cu -- stand-alone k-table interpolation (calc_k / calc_kg), HBM-bound streaming kernel.
//
// Data layout: the table stays in the reference's layout K[NWAVE,NG,NP,NT,NGAS] (gas fastest), so
// for one (wave,g) pair every (p,T,gas) entry lives in one contiguous NP*NT*NGAS slab (14.4 KB at
// 20x15x6).  One CTA owns PAIRS_PER_CTA consecutive (wave,g) pairs; its threads sweep the
// (layer,gas) outputs of a pair in order, so the stores to k[wave,g,:,:] are fully coalesced and
// the four corner loads of neighbouring threads fall in the same 48-byte gas vectors (neighbouring
// layers share brackets), i.e. they are served by L1 after the first touch.  ln K is pre-tabulated
// (api.cu) so each output costs four loads and one exp.
#include "kinterp.cuh"

constexpr int KI_THREADS = 256;
constexpr int KI_PAIRS_PER_CTA = 4;

template <bool GRAD>
__global__ void __launch_bounds__(KI_THREADS)
ans_kinterp_kernel(const double *__restrict__ lnK, const double *__restrict__ K, AnsLayerPlan plan, int npairs,
                   int NP, int NT, int NGAS, int NLAY, double *__restrict__ kout, double *__restrict__ dkout)
{
    const size_t slab = (size_t)NP * NT * NGAS;
    const int per_pair = NLAY * NGAS;
    const int pair0 = blockIdx.x * KI_PAIRS_PER_CTA;
    for (int pp = 0; pp < KI_PAIRS_PER_CTA; ++pp) {
        const int pair = pair0 + pp;
        if (pair >= npairs) return;
        const size_t tbase = (size_t)pair * slab;
        const size_t obase = (size_t)pair * per_pair;
        for (int e = threadIdx.x; e < per_pair; e += KI_THREADS) {
            const int l = e / NGAS;
            const int gas = e - l * NGAS;
            const size_t off00 = tbase + ((size_t)__ldg(plan.ip_lo + l) * NT + __ldg(plan.it_lo + l)) * NGAS + gas;
            const double *w = plan.w4 + 4 * l;
            double kv, dv = 0.0;
            ans_kinterp_elem<GRAD>(lnK, K, off00, NT, NGAS, __ldg(w), __ldg(w + 1), __ldg(w + 2), __ldg(w + 3),
                                   GRAD ? __ldg(plan.omv + l) : 0.0, GRAD ? __ldg(plan.vv + l) : 0.0,
                                   GRAD ? __ldg(plan.dudt + l) : 0.0, kv, dv);
            kout[obase + e] = kv;
            if (GRAD) dkout[obase + e] = dv;
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
    const int grid = ans_div_up(npairs, KI_PAIRS_PER_CTA);
    if (want_grad)
        ans_kinterp_kernel<true><<<grid, KI_THREADS, 0, stream>>>(t->lnK, t->K, plan, npairs, t->NP, t->NT, t->NGAS,
                                                                  NLAY, k, dkdT);
    else
        ans_kinterp_kernel<false><<<grid, KI_THREADS, 0, stream>>>(t->lnK, t->K, plan, npairs, t->NP, t->NT, t->NGAS,
                                                                   NLAY, k, nullptr);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}
