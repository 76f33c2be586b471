// koverlap.cu -- C entry points of the random-overlap kernels (implementation: koverlap_impl.cuh).
#include "koverlap_impl.cuh"
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <string.h>

// NGAS == 1: ForwardModel_0.py:5871-5876 / :6056-6058
template <bool GRAD>
__global__ void ans_koverlap_single_kernel(OvParams P)
{
    const size_t n = (size_t)P.NWAVE * P.NG * P.NLAY;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(o % P.NLAY);
        double kv, dv = 0.0;
        if (P.fused) {
            const size_t pair = o / P.NLAY;
            const size_t plane = (size_t)P.NWAVE * P.NG;          // NGAS == 1
            const double *w = P.plan.w4 + 4 * l;
            ans_kinterp_elem<GRAD>(P.tab, ((size_t)P.plan.ip_lo[l] * P.NT + P.plan.it_lo[l]) * plane + pair, P.NT,
                                   plane, w[0], w[1], w[2], w[3], GRAD ? P.plan.omv[l] : 0.0, GRAD ? P.plan.vv[l] : 0.0,
                                   GRAD ? P.plan.dudt[l] : 0.0, kv, dv);
        } else {
            kv = P.k[o];
            if (GRAD) dv = P.dkdT[o];
        }
        const double am = P.amount[l];
        P.tau[o] = __dmul_rn(kv, am);
        if (GRAD) { P.dk[o * 2] = kv; P.dk[o * 2 + 1] = __dmul_rn(dv, am); }
    }
}

bool ov_fast_supported(const OvParams &P, bool grad);
int ov_fast_launch(const OvParams &P, bool grad, int *scratch, int *why, int *work, cudaStream_t stream);

static int ov_general(const OvParams &P, bool grad, cudaStream_t stream)
{
    const int NN = P.NG * P.NG;
    if (NN <= 128) return ov_dispatch_4(P, grad, stream);
    if (NN <= 256) return ov_dispatch_8(P, grad, stream);
    return ov_dispatch_16(P, grad, stream);
}

// Work-list scratch of the fast kernel: device buffers are kept for re-use; a buffer handed back carries an
// event recorded after its last use and the next user makes its stream wait for it, so nothing blocks the host
// and distinct streams / threads never share a live buffer.
namespace {
struct OvScratch { int *p; size_t n; cudaEvent_t ev; int dev; };
std::mutex ov_scratch_mu;
std::vector<OvScratch> ov_scratch_free;

int ov_scratch_get(size_t n, cudaStream_t stream, OvScratch &out)
{
    int dev = 0;
    ANS_CUDA_CHECK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> g(ov_scratch_mu);
        for (size_t i = 0; i < ov_scratch_free.size(); ++i) {
            if (ov_scratch_free[i].dev == dev && ov_scratch_free[i].n >= n) {
                out = ov_scratch_free[i];
                ov_scratch_free.erase(ov_scratch_free.begin() + i);
                ANS_CUDA_CHECK(cudaStreamWaitEvent(stream, out.ev, 0));
                return ANSB200_OK;
            }
        }
    }
    out.n = n < (1u << 20) ? (1u << 20) : n;
    out.dev = dev;
    if (cudaMalloc((void **)&out.p, out.n * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        ansb200_set_error("koverlap: cannot allocate %zu bytes of work-list scratch", out.n * sizeof(int));
        return ANSB200_ENOMEM;
    }
    ANS_CUDA_CHECK(cudaEventCreateWithFlags(&out.ev, cudaEventDisableTiming));
    return ANSB200_OK;
}

void ov_scratch_put(OvScratch &sc, cudaStream_t stream)
{
    cudaEventRecord(sc.ev, stream);
    std::lock_guard<std::mutex> g(ov_scratch_mu);
    ov_scratch_free.push_back(sc);
}
}   // namespace

// Diagnostics.  Mode 0 = fast kernel + work list (default), 1 = general kernel only, 2 = as 0 but every call
// synchronises its stream and records how many cells were handed over and why (ansb200_overlap_stats).  The mode
// comes from ansb200_overlap_mode() or, initially, from the environment (ANSB200_OVERLAP=general|stats).
static std::atomic<int> g_ov_mode{-1};
static std::atomic<int> g_ov_stats[9];

static int ov_mode()
{
    int m = g_ov_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("ANSB200_OVERLAP");
        m = !e ? 0 : (!strcmp(e, "general") ? 1 : (!strcmp(e, "stats") ? 2 : 0));
        g_ov_mode.store(m, std::memory_order_relaxed);
    }
    return m;
}

extern "C" int ansb200_overlap_mode(int mode)
{
    const int old = ov_mode();
    if (mode >= 0 && mode <= 2) g_ov_mode.store(mode, std::memory_order_relaxed);
    return old;
}

extern "C" void ansb200_overlap_stats(int32_t *out9)
{
    for (int i = 0; i < 9; ++i) out9[i] = g_ov_stats[i].load(std::memory_order_relaxed);
}

static int ov_run(OvParams &P, bool grad, cudaStream_t stream)
{
    ANS_REQUIRE(P.NWAVE > 0 && P.NG > 0 && P.NLAY > 0 && P.NGAS > 0, "koverlap: bad shape");
    ANS_REQUIRE(P.NG <= ANSB200_MAX_NG, "koverlap: NG=%d exceeds %d", P.NG, ANSB200_MAX_NG);
    ANS_REQUIRE(P.NGAS <= ANSB200_MAX_NGAS, "koverlap: NGAS=%d exceeds %d", P.NGAS, ANSB200_MAX_NGAS);
    ANS_REQUIRE(P.amount && P.tau && (!grad || P.dk), "koverlap: null pointer");
    if (P.NGAS == 1) {
        const size_t n = (size_t)P.NWAVE * P.NG * P.NLAY;
        int grid = (int)((n + 255) / 256);
        if (grid > 148 * 16) grid = 148 * 16;
        if (grad) ans_koverlap_single_kernel<true><<<grid, 256, 0, stream>>>(P);
        else ans_koverlap_single_kernel<false><<<grid, 256, 0, stream>>>(P);
        ANS_LAUNCH_CHECK();
        return ANSB200_OK;
    }
    ANS_REQUIRE(P.weight && P.g_ord, "koverlap: weight/g_ord tables are required for NGAS > 1");
    if (ov_mode() != 1 && ov_fast_supported(P, grad)) {
        // fast kernel for the common case; the cells it declines (tie order, non-monotone k, ...) are listed for
        // the general kernel, which runs on that list afterwards (usually empty: it exits before its set-up)
        const long long ncell = (long long)P.NWAVE * P.NLAY;
        ANS_REQUIRE(ncell < 0x7fffffffLL, "koverlap: NWAVE*NLAY too large");
        const bool stats = ov_mode() == 2;
        OvScratch sc{};
        {
            const int rc0 = ov_scratch_get((size_t)ncell + 1 + 8 + 1, stream, sc);
            if (rc0 != ANSB200_OK) return rc0;
        }
        int *scratch = sc.p;
        ANS_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(int), stream));
        ANS_CUDA_CHECK(cudaMemsetAsync(scratch + ncell + 1, 0, 9 * sizeof(int), stream));      // statistics + work counter
        int rc = ov_fast_launch(P, grad, scratch, stats ? scratch + ncell + 1 : nullptr, scratch + ncell + 9, stream);
        if (rc == ANSB200_OK) {
            P.cell_count = scratch;
            P.cell_list = scratch + 1;
            rc = ov_general(P, grad, stream);
        }
        if (rc == ANSB200_OK && ov_mode() == 2) {
            int c = 0, why[8];
            cudaMemcpyAsync(&c, scratch, sizeof(int), cudaMemcpyDeviceToHost, stream);
            cudaMemcpyAsync(why, scratch + ncell + 1, sizeof(why), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            g_ov_stats[0].store(c, std::memory_order_relaxed);
            for (int i = 0; i < 8; ++i) g_ov_stats[1 + i].store(why[i], std::memory_order_relaxed);
            if (getenv("ANSB200_OVERLAP_VERBOSE"))
                fprintf(stderr, "[ansb200] overlap: %d of %lld cells left to the general kernel (non-monotone %d, open bin %d, "
                                "group %d, tie %d, walk %d); folds: %d static, %d sorted (%d static orders with an unseparated "
                                "straddler)\n", c, ncell, why[0], why[1], why[2], why[3], why[4], why[5], why[6], why[7]);
        }
        ov_scratch_put(sc, stream);
        return rc;
    }
    return ov_general(P, grad, stream);
}

// The host checks (plan.overlap_tables) that no sorted element can straddle two bin edges; if it
// can, bit 1 of want_grad requests the literal sequential bin-edge scan.
extern "C" int ansb200_koverlap(const double *k, const double *dkdT, const double *amount, const double *weight,
                                const double *g_ord, const double *del_g, int NWAVE, int NG, int NLAY, int NGAS, int want_grad,
                                double *tau, double *dk, void *stream_)
{
    const bool grad = (want_grad & 1) != 0;
    ANS_REQUIRE(k && (!grad || dkdT), "koverlap: null k/dkdT");
    OvParams P{};
    P.k = k; P.dkdT = dkdT; P.amount = amount; P.weight = weight; P.g_ord = g_ord; P.del_g = del_g;
    P.NWAVE = NWAVE; P.NG = NG; P.NLAY = NLAY; P.NGAS = NGAS; P.tau = tau; P.dk = dk;
    P.fused = 0; P.seq_rebin = (want_grad & 2) ? 1 : 0;
    return ov_run(P, grad, (cudaStream_t)stream_);
}

extern "C" int ansb200_gas_opacity(const ansb200_table *t, int NLAY, const int32_t *ip_lo, const int32_t *it_lo,
                                   const double *w4, const double *omv, const double *vv, const double *dudt,
                                   const double *amount, const double *weight, const double *g_ord,
                                   const double *del_g, int want_grad, double *tau, double *dk, void *stream_)
{
    const bool grad = (want_grad & 1) != 0;
    ANS_REQUIRE(t && ip_lo && it_lo && w4, "gas_opacity: null pointer");
    ANS_REQUIRE(!grad || (omv && vv && dudt), "gas_opacity: gradient requested without omv/vv/dudt");
    OvParams P{};
    P.tab = ans_tab(t); P.plan = AnsLayerPlan{ip_lo, it_lo, w4, omv, vv, dudt};
    P.NP = t->NP; P.NT = t->NT;
    P.amount = amount; P.weight = weight; P.g_ord = g_ord; P.del_g = del_g;
    P.NWAVE = t->NWAVE; P.NG = t->NG; P.NLAY = NLAY; P.NGAS = t->NGAS; P.tau = tau; P.dk = dk;
    P.fused = 1; P.seq_rebin = (want_grad & 2) ? 1 : 0;
    return ov_run(P, grad, (cudaStream_t)stream_);
}
