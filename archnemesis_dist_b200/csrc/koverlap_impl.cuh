// koverlap_impl.cuh -- random-overlap gas mixing (k_overlap/rank, k_overlapg/rankg) on sm_100a.
//
// Reference: archnemesis/ForwardModel_0.py:6029-6173 (no gradients), :5842-6026 (gradients).
//
// Work decomposition: persistent CTAs of OV_WARPS warps; one warp owns one (wavenumber, layer) cell at a
// time and folds the NGAS gases in sequence exactly like the reference.  Per fold:
//   0. row-/column-major key orders are data-independent: fixed rebin matrices (ov_rebin_static);
//   1. otherwise the NG*NG keys  tau_i + k_j*amount  are packed as (22 key bits | 10 index bits) and
//      sorted by a 32-bit min/max bitonic network in registers (EPL words per lane, padded to 32*EPL,
//      shuffles for the cross-lane stages); the order is verified / repaired against the exact float64
//      keys only when two neighbours share their key bits, and tied keys follow numba's quicksort
//      (ov_numba_order) when gradients are wanted;
//   2. the cumulative weight is a warp scan; every lane FMA-accumulates cont*w, w and the gradient
//      columns of its EPL consecutive sorted elements, closing a partial sum at every element that
//      straddles a bin edge; lane m assembles bin m (partial sums + the frac / 1-frac parts of its two
//      straddlers) and normalises (ov_rebin_par).
// With identical k inputs ov_rebin_par agrees with the reference to ~1e-15 (another summation order);
// the literal sequential rebin (ov_rebin_seq: lane-per-bin walk in sorted order, requested by the host
// when one element could straddle two bin edges, or by force_seq) is bit-identical to it.
//
// k and dk/dT of the cell come either from the arrays produced by ansb200_kinterp or, in the
// fused entry point, straight from the resident ln K table (k_gas never touches HBM).
#pragma once
#include "kinterp.cuh"

#ifndef OV_NWARPS
#define OV_NWARPS 8
#endif
constexpr int OV_WARPS = OV_NWARPS;
#ifndef OV_MINB
#define OV_MINB 2
#endif
#ifndef OV_MINB_NOGRAD
#define OV_MINB_NOGRAD 3
#endif
constexpr unsigned FULL = 0xffffffffu;
constexpr int OV_CS = 6;   // doubles per column record {b, bT, k, -, -, -}: 3 sixteen-byte units, conflict-free mod 8
constexpr int OV_NONE = 0x7fffffff;
constexpr unsigned OV_KEY_BASE = (227u - 62u) << 23;   // float32 pattern of 2^(100-62): origin of the packed sort keys

struct OvParams {
    const double *k, *dkdT;                 // unfused source [NWAVE,NG,NLAY,NGAS]
    AnsTab tab;                             // fused source (table)
    AnsLayerPlan plan;
    int NP, NT;
    const double *amount, *weight, *g_ord;
    const double *del_g;                    // [NG] quadrature weights (may be NULL): lets the kernel form del_g[i]*del_g[j] itself
    int NWAVE, NG, NLAY, NGAS;
    double *tau, *dk;
    int fused, seq_rebin;
    // work list of the fast kernel (koverlap_fast.cu): cell_count[0] cells, numbered in cell_list; NULL or a
    // negative count = every cell
    const int *cell_count, *cell_list;
};

__device__ __forceinline__ double shfl_xor_d(double v, int m)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(FULL, lo, m);
    hi = __shfl_xor_sync(FULL, hi, m);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double shfl_up_d(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(FULL, lo, d);
    hi = __shfl_up_sync(FULL, hi, d);
    return __hiloint2double(hi, lo);
}

// Sorting network over 32*EPL (key, idx) pairs held blocked across the warp: element e = lane*EPL + r.
//
// Bitonic merge sort in its all-ascending form: the first stage of the merge of width k pairs e with
// e ^ (k-1) (mirror image inside the block), the remaining stages pair e with e ^ j, j = k/4 .. 1,
// and every comparator leaves the smaller element at the smaller index.  Inside a lane all
// comparators are therefore static; across lanes the only run-time flag is "this lane holds the
// smaller index".  EXACT selects the total order (key, idx); otherwise only keys are compared (a valid
// ascending sequence, equal keys in network order) and the caller re-sorts exactly when it sees ties.
// Compare/select is written in PTX (setp/selp) so that ptxas emits predicated selects, not branches.
// Stages with k > EPL keep the k and j loops rolled so the code stays inside the instruction cache.
template <bool EXACT>
__device__ __forceinline__ void ov_ce(double &ka, int &ia, double &kb, int &ib)
{
    // afterwards (ka,ia) sorts before (kb,ib)
    if (EXACT) {
        asm("{\n\t.reg .pred lt, eq, il, sw;\n\t.reg .f64 t;\n\t.reg .s32 u;\n\t"
            "setp.lt.f64 lt, %2, %0;\n\t"
            "setp.eq.f64 eq, %2, %0;\n\t"
            "setp.lt.s32 il, %3, %1;\n\t"
            "and.pred eq, eq, il;\n\t"
            "or.pred sw, lt, eq;\n\t"
            "selp.f64 t, %2, %0, sw;\n\t"
            "selp.f64 %2, %0, %2, sw;\n\t"
            "mov.f64 %0, t;\n\t"
            "selp.s32 u, %3, %1, sw;\n\t"
            "selp.s32 %3, %1, %3, sw;\n\t"
            "mov.s32 %1, u;\n\t}"
            : "+d"(ka), "+r"(ia), "+d"(kb), "+r"(ib));
    } else {
        asm("{\n\t.reg .pred sw;\n\t.reg .f64 t;\n\t.reg .s32 u;\n\t"
            "setp.lt.f64 sw, %2, %0;\n\t"
            "selp.f64 t, %2, %0, sw;\n\t"
            "selp.f64 %2, %0, %2, sw;\n\t"
            "mov.f64 %0, t;\n\t"
            "selp.s32 u, %3, %1, sw;\n\t"
            "selp.s32 %3, %1, %3, sw;\n\t"
            "mov.s32 %1, u;\n\t}"
            : "+d"(ka), "+r"(ia), "+d"(kb), "+r"(ib));
    }
}

// cross-lane comparator: (k,i) is mine, (pk,pi) the partner's; keep the smaller if keep_low else the larger
template <bool EXACT>
__device__ __forceinline__ void ov_ce_x(double &k, int &i, double pk, int pi, int keep_low)
{
    if (EXACT) {
        asm("{\n\t.reg .pred lt, eq, il, kl, tk;\n\t"
            "setp.lt.f64 lt, %2, %0;\n\t"
            "setp.eq.f64 eq, %2, %0;\n\t"
            "setp.lt.s32 il, %3, %1;\n\t"
            "and.pred eq, eq, il;\n\t"
            "or.pred lt, lt, eq;\n\t"          // partner sorts before mine (never equal: idx unique)
            "setp.ne.s32 kl, %4, 0;\n\t"
            "xor.pred tk, lt, kl;\n\t"
            "not.pred tk, tk;\n\t"             // take = (lt == keep_low)
            "selp.f64 %0, %2, %0, tk;\n\t"
            "selp.s32 %1, %3, %1, tk;\n\t}"
            : "+d"(k), "+r"(i) : "d"(pk), "r"(pi), "r"(keep_low));
    } else {
        asm("{\n\t.reg .pred lt, gt, kl, tk;\n\t"
            "setp.lt.f64 lt, %2, %0;\n\t"
            "setp.gt.f64 gt, %2, %0;\n\t"
            "setp.ne.s32 kl, %4, 0;\n\t"
            "and.pred lt, lt, kl;\n\t"
            "not.pred kl, kl;\n\t"
            "and.pred gt, gt, kl;\n\t"
            "or.pred tk, lt, gt;\n\t"          // take = keep_low ? pk < k : pk > k  (equal keys: nobody moves)
            "selp.f64 %0, %2, %0, tk;\n\t"
            "selp.s32 %1, %3, %1, tk;\n\t}"
            : "+d"(k), "+r"(i) : "d"(pk), "r"(pi), "r"(keep_low));
    }
}

template <int EPL, bool EXACT>
__device__ __forceinline__ void ov_intra_halfcleaners(double (&key)[EPL], int (&idx)[EPL], int jstart)
{
#pragma unroll
    for (int j = jstart; j > 0; j >>= 1) {
#pragma unroll
        for (int r = 0; r < EPL; ++r)
            if ((r & j) == 0) ov_ce<EXACT>(key[r], idx[r], key[r | j], idx[r | j]);
    }
}

template <int EPL, bool EXACT>
__device__ __forceinline__ void ov_bitonic_sort(double (&key)[EPL], int (&idx)[EPL], int lane)
{
    // runs inside one lane: k = 2 .. EPL
#pragma unroll
    for (int k = 2; k <= EPL; k <<= 1) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
            const int q = r ^ (k - 1);
            if (r < q) ov_ce<EXACT>(key[r], idx[r], key[q], idx[q]);
        }
        ov_intra_halfcleaners<EPL, EXACT>(key, idx, k >> 2);
    }
    // merges across lanes: k = kl*EPL, kl = 2 .. 32
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {
        {   // mirror stage: partner element = e ^ (k-1): lane ^ (kl-1), register EPL-1-r
            const int keep_low = (lane & (kl >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < EPL / 2; ++r) {
                const int q = EPL - 1 - r;
                const double pkr = shfl_xor_d(key[q], kl - 1), pkq = shfl_xor_d(key[r], kl - 1);
                const int pir = __shfl_xor_sync(FULL, idx[q], kl - 1), piq = __shfl_xor_sync(FULL, idx[r], kl - 1);
                ov_ce_x<EXACT>(key[r], idx[r], pkr, pir, keep_low);
                ov_ce_x<EXACT>(key[q], idx[q], pkq, piq, keep_low);
            }
        }
#pragma unroll 1
        for (int lm = kl >> 2; lm > 0; lm >>= 1) {   // half-cleaners across lanes: partner e ^ (lm*EPL)
            const int keep_low = (lane & lm) == 0;
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                const double pk = shfl_xor_d(key[r], lm);
                const int pi = __shfl_xor_sync(FULL, idx[r], lm);
                ov_ce_x<EXACT>(key[r], idx[r], pk, pi, keep_low);
            }
        }
        ov_intra_halfcleaners<EPL, EXACT>(key, idx, EPL >> 1);
    }
}

// ---- fast path: 32-bit packed (22-bit key | 10-bit index) network --------------------------------
// The keys are scaled by a power of two (kmax -> 2^100) and rounded to float32 (monotone); 6 exponent
// bits relative to 2^(100-62) and 16 mantissa bits of the (positive) float32 pattern are packed above
// the 10-bit element index.  One unsigned min/max then orders (key22, index): a comparator is two instructions on one
// register.  The order is exact except inside groups of keys that agree to 14 mantissa bits
// (~0.5 pairs per fold at NG=20); the caller re-forms the exact float64 keys, repairs such groups with
// odd-even transposition passes on (key, index) and falls back to the float64 network
// (ov_bitonic_sort<EPL, true>) if a few passes do not suffice, so the result is always the exact order.
template <int EPL>
__device__ __forceinline__ void ov_bitonic_sort_u32(unsigned (&v)[EPL], int lane)
{
#define OV_CE32(x, y) do { const unsigned lo__ = min(x, y), hi__ = max(x, y); x = lo__; y = hi__; } while (0)
#pragma unroll
    for (int k = 2; k <= EPL; k <<= 1) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
            const int q = r ^ (k - 1);
            if (r < q) OV_CE32(v[r], v[q]);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < EPL; ++r)
                if ((r & j) == 0) OV_CE32(v[r], v[r | j]);
        }
    }
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {
        {
            const bool keep_low = (lane & (kl >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < EPL / 2; ++r) {
                const int q = EPL - 1 - r;
                const unsigned pr = __shfl_xor_sync(FULL, v[q], kl - 1), pq = __shfl_xor_sync(FULL, v[r], kl - 1);
                v[r] = keep_low ? min(v[r], pr) : max(v[r], pr);
                v[q] = keep_low ? min(v[q], pq) : max(v[q], pq);
            }
        }
#pragma unroll 1
        for (int lm = kl >> 2; lm > 0; lm >>= 1) {
            const bool keep_low = (lane & lm) == 0;
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                const unsigned pv = __shfl_xor_sync(FULL, v[r], lm);
                v[r] = keep_low ? min(v[r], pv) : max(v[r], pv);
            }
        }
#pragma unroll
        for (int j = EPL >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < EPL; ++r)
                if ((r & j) == 0) OV_CE32(v[r], v[r | j]);
        }
    }
#undef OV_CE32
}

// one odd-even transposition pass over the whole sequence on the exact (key, index) order
template <int EPL>
__device__ __forceinline__ void ov_fixup_pass(double (&key)[EPL], int (&idx)[EPL], int lane)
{
#pragma unroll
    for (int r = 0; r + 1 < EPL; r += 2) ov_ce<true>(key[r], idx[r], key[r + 1], idx[r + 1]);
#pragma unroll
    for (int r = 1; r + 1 < EPL; r += 2) ov_ce<true>(key[r], idx[r], key[r + 1], idx[r + 1]);
    // the pair that spans two lanes: my last element against the next lane's first
    double nk = __hiloint2double(__shfl_down_sync(FULL, __double2hiint(key[0]), 1),
                                 __shfl_down_sync(FULL, __double2loint(key[0]), 1));
    int ni = __shfl_down_sync(FULL, idx[0], 1);
    double pk = __hiloint2double(__shfl_up_sync(FULL, __double2hiint(key[EPL - 1]), 1),
                                 __shfl_up_sync(FULL, __double2loint(key[EPL - 1]), 1));
    int pi = __shfl_up_sync(FULL, idx[EPL - 1], 1);
    if (lane == 31) { nk = key[EPL - 1]; ni = idx[EPL - 1]; }     // no neighbour: compare with itself (no-op)
    if (lane == 0) { pk = key[0]; pi = idx[0]; }
    ov_ce_x<true>(key[EPL - 1], idx[EPL - 1], nk, ni, 1);
    ov_ce_x<true>(key[0], idx[0], pk, pi, 0);
}

// order check of the sorted sequence: bit 0 = some neighbour pair is out of order (exact keys),
// bit 1 = some live neighbour pair carries equal keys
template <int EPL>
__device__ __forceinline__ int ov_check_order(const double (&key)[EPL], int lane)
{
    bool bad = false, tie = false;
#pragma unroll
    for (int r = 0; r + 1 < EPL; ++r) {
        bad |= key[r] > key[r + 1];
        tie |= (key[r] == key[r + 1]) & (key[r] != INFINITY);
    }
    const double nxt = __hiloint2double(__shfl_down_sync(FULL, __double2hiint(key[0]), 1),
                                        __shfl_down_sync(FULL, __double2loint(key[0]), 1));
    bad |= (lane < 31) & (key[EPL - 1] > nxt);
    tie |= (lane < 31) & (key[EPL - 1] == nxt) & (nxt != INFINITY);
    return (__any_sync(FULL, bad) ? 1 : 0) | (__any_sync(FULL, tie) ? 2 : 0);
}

// Weight of element (i,j).  The reference's table is del_g[i]*del_g[j] evaluated in del_g's dtype
// (float32 on the .kta path).  A random lookup in the NG*NG table of doubles costs ~7 shared-memory
// wavefronts per warp (bank conflicts) and was 28 % of the kernel's shared-memory traffic; two lookups in
// the NG-entry float table are conflict-free (equal indices broadcast), and the float32 product is the
// table entry bit for bit.  The CTA checks that at start-up (`f32`); otherwise (float64 del_g, or a
// caller-made table) the table is used.
struct OvWeight {
    const double *wtab;
    const float *dgf;
    int NG;
    bool f32;
    __device__ __forceinline__ double operator()(int i, int j) const
    {
        if (f32) return (double)__fmul_rn(dgf[i], dgf[j]);
        return wtab[i * NG + j];
    }
};

// Static-order tables: [order][RA | RB][NG*NG] + [order][NG] (+1 to stay even), see ov_static_setup
__host__ __device__ inline int ov_static_doubles(int NG) { return (4 * NG * NG + 2 * NG + 1) & ~1; }

// Per-warp shared-memory view.
struct OvWarpSmem {
    double *kbuf, *dbuf;     // [NG*NGAS] k and dk/dT of the cell
    double *a, *b, *bT;      // [NG] running tau_g, next gas tau, next gas dk/dT*amount
    double *colB;            // [NG*OV_CS] {b, bT, k of the gas being folded} packed per column for the rebin loop (16-byte aligned)
    double *dkp;             // [NG*DS] rows [tau_i (copy made by the rebin), dT, gas columns] (see ov_ds); 16-byte aligned
    double *frac, *gdn;      // [NG+1] cumulative weight before / after the straddler of every edge
    double *bsum;            // [NG*BS] raw bin sums (see ov_bs)
    double *head;            // [32*BS] per-lane partial sum of the bin a lane starts in (parallel rebin)
    int *strad;              // [NG+1]
    int *spare;              // [NG] (unused; keeps the int block an even count)
    unsigned short *sidx;    // [NG*NG] sorted packed indices
};

// Gradient storage (template NPMAX >= NGAS+1): row i of dkp is [ tau_i, dT, gas 0 .. gas NPMAX-2, - ] with the
// even stride DS = NPMAX+2 (rows 16-byte aligned: the rebin loop reads them as double2; DS/2 is odd for
// NPMAX = 8, so the 8 lanes of a 128-bit wavefront conflict only when their rows are equal mod 8);
// columns of gases not folded yet hold 0.  Raw bin
// sums and head slots are rows of BS = (NPMAX+3)|1 doubles: cont*w, w, dT, gas 0 .. gas NPMAX-2, and the
// column of the gas being folded (kept apart so that every gas column is processed alike).
__host__ __device__ inline int ov_npmax(int NGAS) { return NGAS + 1 <= 4 ? 4 : (NGAS + 1 <= 8 ? 8 : 16); }
__host__ __device__ inline int ov_ds(int npmax) { return npmax + 2; }
__host__ __device__ inline int ov_bs(int npmax) { return (npmax + 3) | 1; }

// head slots of the parallel rebin (32 lanes x BS); the same region is the scratch of the tie-order
// emulation (4 uint16 arrays of the padded sort length = that many doubles)
// (only the gradient kernels replay numba's tie order, so only they need the scratch)
__host__ __device__ inline int ov_head_doubles_np(int NG, int npmax, bool grad)
{
    int nnpad = 128;
    while (nnpad < NG * NG) nnpad <<= 1;
    const int h = 32 * ov_bs(npmax);
    return (!grad || h > nnpad) ? h : nnpad;
}

__host__ __device__ inline size_t ov_per_warp_bytes(int NG, int NGAS, bool grad)
{
    int NN = 128;                       // sorted-index staging is padded to 32*EPL entries
    while (NN < NG * NG) NN <<= 1;
    const int npm = grad ? ov_npmax(NGAS) : 1;      // (the no-gradient kernels are instantiated with NPMAX = 1)
    const int nd = NG * NGAS * (grad ? 2 : 1) + (3 + OV_CS) * NG + (grad ? NG * ov_ds(npm) : 0) + 2 * (NG + 1) + NG * ov_bs(npm) +
                   ov_head_doubles_np(NG, npm, grad);
    return ((size_t)nd * 8 + (size_t)(2 * NG + 2) * 4 + (size_t)NN * 2 + 15) & ~(size_t)15;
}

// Rebin by lane-per-bin walk in sorted order (bit-identical rounding sequence to rank/rankg).  Used when
// the host requests the literal sequential bin-edge scan (an element could straddle two edges).
template <int NPMAX, bool GRAD>
__device__ __noinline__ void ov_rebin_seq(OvWarpSmem s, const double *__restrict__ wtab,
                                          const double *__restrict__ gord, int NG, int NGAS, int igas, int lane)
{
    // s.sidx holds the sorted packed indices (written by the caller)
    const int NN = NG * NG;
    constexpr int DS = NPMAX + 2;
    const int g1 = igas + 1;
    const int seq_rebin = 1;
    __syncwarp();

    double res_tau = 0.0;
    double res_g[GRAD ? NPMAX : 1];
    const int n = igas + 3;   // live gradient columns
    if (GRAD) {
#pragma unroll
        for (int p = 0; p < NPMAX; ++p) res_g[p] = 0.0;
    }

    auto elem = [&](int pos, double &w, double &cw, double (&gw)[GRAD ? NPMAX : 1]) {
        const int pi = s.sidx[pos];
        const int i = pi >> 5, j = pi & 31;
        w = wtab[i * NG + j];
        cw = __dmul_rn(__dadd_rn(s.a[i], s.b[j]), w);
        if (GRAD) {
#pragma unroll
            for (int p = 0; p < NPMAX; ++p) {
                if (p < n) {
                    double g;
                    if (p <= igas) g = s.dkp[i * DS + 2 + p];
                    else if (p == g1) g = s.kbuf[j * NGAS + g1];
                    else g = __dadd_rn(s.dkp[i * DS + 1], s.bT[j]);
                    gw[p] = __dmul_rn(g, w);
                }
            }
        }
    };

    if (seq_rebin) {
        // Literal sequential bin-edge scan of rank/rankg (ForwardModel_0.py:6155-6172) by lane 0,
        // requested by the host when an element could straddle two edges: an element closes at
        // most one bin, exactly like the reference loop.
        if (lane == 0) {
            for (int m = 0; m <= NG; ++m) { s.strad[m] = OV_NONE; s.frac[m] = 0.0; }
            // gdist[iloop-1] wraps to the LAST cumulative weight when the very first element already
            // reaches the first edge (python index -1, ForwardModel_0.py:6160 / :6009): start from the total
            double gprev = 0.0;
            for (int pos = 0; pos < NN; ++pos) {
                const int pi = s.sidx[pos];
                gprev = __dadd_rn(gprev, wtab[(pi >> 5) * NG + (pi & 31)]);
            }
            double run = 0.0;
            int ig = 0;
            for (int pos = 0; pos < NN && ig < NG; ++pos) {
                const int pi = s.sidx[pos];
                run = __dadd_rn(run, wtab[(pi >> 5) * NG + (pi & 31)]);
                if (!(run < gord[ig + 1])) {
                    s.strad[ig + 1] = pos;
                    s.frac[ig + 1] = __ddiv_rn(__dsub_rn(gord[ig + 1], gprev), __dsub_rn(run, gprev));
                    ++ig;
                }
                gprev = run;
            }
        }
        __syncwarp();
    }

    // lane m accumulates bin m in sorted order with the reference's rounding sequence
    const int m = lane;
    if (m < NG) {
        const int e0 = (m == 0) ? -1 : s.strad[m];
        const int e1 = s.strad[m + 1];
        if (e0 != OV_NONE) {
            double acc = 0.0, sum1 = 0.0, w, cw;
            double gw[GRAD ? NPMAX : 1];
            if (m > 0) {
                const double omf = __dsub_rn(1.0, s.frac[m]);
                elem(e0, w, cw, gw);
                acc = __dmul_rn(omf, cw);
                sum1 = __dmul_rn(omf, w);
                if (GRAD) {
#pragma unroll
                    for (int p = 0; p < NPMAX; ++p) if (p < n) res_g[p] = __dmul_rn(omf, gw[p]);
                }
            }
            const int stop = (e1 == OV_NONE) ? NN : e1;
            for (int pos = e0 + 1; pos < stop; ++pos) {
                elem(pos, w, cw, gw);
                acc = __dadd_rn(acc, cw);
                sum1 = __dadd_rn(sum1, w);
                if (GRAD) {
#pragma unroll
                    for (int p = 0; p < NPMAX; ++p) if (p < n) res_g[p] = __dadd_rn(res_g[p], gw[p]);
                }
            }
            bool norm = (m == NG - 1);
            if (e1 != OV_NONE) {
                const double f = s.frac[m + 1];
                elem(e1, w, cw, gw);
                acc = __dadd_rn(acc, __dmul_rn(f, cw));
                sum1 = __dadd_rn(sum1, __dmul_rn(f, w));
                if (GRAD) {
#pragma unroll
                    for (int p = 0; p < NPMAX; ++p) if (p < n) res_g[p] = __dadd_rn(res_g[p], __dmul_rn(f, gw[p]));
                }
                norm = true;
            }
            if (norm) {
                acc = __ddiv_rn(acc, sum1);
                if (GRAD) {
#pragma unroll
                    for (int p = 0; p < NPMAX; ++p) if (p < n) res_g[p] = __ddiv_rn(res_g[p], sum1);
                }
            }
            res_tau = acc;
        }
    }
    __syncwarp();
    if (m < NG) {
        s.a[m] = res_tau;
        if (GRAD) {
#pragma unroll
            for (int p = 0; p < NPMAX; ++p) {
                if (p <= g1) s.dkp[m * DS + 2 + p] = res_g[p];       // gas columns
                else if (p == g1 + 1) s.dkp[m * DS + 1] = res_g[p];      // temperature column
            }
            for (int p = g1 + 1; p < NGAS; ++p) s.dkp[m * DS + 2 + p] = 0.0;
        }
    }
    __syncwarp();
}

// Rebin in parallel over the sorted elements (rank/rankg, ForwardModel_0.py:6155-6172 / :6002-6025).
// Every lane walks its EPL consecutive sorted elements (indices staged in shared memory so the loops
// stay rolled and the hot code fits the instruction cache) with a running cumulative weight and
// accumulates cont*w, w, the T column and the gas columns of the elements that lie WHOLLY inside the
// bin it is in.  The element that straddles a bin edge only closes the lane's partial sum (to the
// lane's head slot if the bin was opened by an earlier lane, else straight to the bin) and leaves
// (element, cumulative weight before / after) in the edge's slot: the straddle branch is divergent
// (some lane takes it in most iterations), so it is kept to a handful of stores.  A bin that spans
// several lanes is the owner lane's tail plus the heads of the following lanes.  Lane m finally adds
// the (1-frac) part of the straddler that opens bin m and the frac part of the one that closes it
// (frac and the products in the reference's rounding order) and normalises.  Same arithmetic as the
// reference in another summation order (agreement ~1e-15); requires that no element straddles two
// edges (host check), else ov_rebin_seq.
template <int EPL, int NPMAX, bool GRAD>
__device__ __forceinline__ void ov_rebin_par(const OvWarpSmem &s, const OvWeight &W,
                                             const double *__restrict__ gord, int NG, int NGAS, int igas, int lane)
{
    // s.sidx[r*32 + lane] = packed index of sorted position lane*EPL + r (staged by the caller)
    constexpr int NQ = GRAD ? NPMAX + 3 : 2;   // cont*w, w, dT, gas 0..NPMAX-2, column of the gas being folded
    constexpr int DS = NPMAX + 2, BS = (NPMAX + 3) | 1;
    const int NN = NG * NG;
    const int g1 = igas + 1;
    for (int t = lane; t < NG * BS; t += 32) s.bsum[t] = 0.0;
    for (int t = lane; t <= NG; t += 32) s.strad[t] = OV_NONE;

    // exclusive prefix of the lane's weight
    double run = 0.0;
#pragma unroll 4
    for (int r = 0; r < EPL; ++r) {
        const int pi = s.sidx[r * 32 + lane];
        run = __dadd_rn(run, (pi >> 5) < NG ? W(pi >> 5, pi & 31) : 0.0);
    }
    double incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = shfl_up_d(incl, d);
        if (lane >= d) incl = __dadd_rn(incl, up);
    }
    // Lane boundaries must be ONE number for both neighbours: the cumulative weight before lane L is lane L-1's
    // inclusive scan value, and lane L-1 uses that same value as the cumulative weight of its last element
    // (below).  With float32-born weights every sum is exact and this changes nothing; with float64 weights
    // (HDF5 tables) a lane's own running sum and the scan differ by an ulp, and an edge lying between the two
    // -- they coincide with the cumulative weight of whole rows in near row-major orders -- would get no
    // straddler at all.
    double prev = shfl_up_d(incl, 1);
    if (lane == 0) prev = 0.0;
    int ig;
    {   // number of edges g_ord[1..NG] that are <= prev
        int lo = 0, hi = NG;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (gord[mid] <= prev) lo = mid; else hi = mid - 1;
        }
        ig = lo;
    }
    __syncwarp();

    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
    double *myhead = s.head + lane * BS;
    if (GRAD) {
        // per-row / per-column operands packed so that the loop below forms two base addresses per element:
        // row i = [dT, gas columns, tau] (the spare slot of the dkp row), column j = [tau, dT part, k]
        for (int t = lane; t < NG; t += 32) {
            s.dkp[t * DS] = s.a[t];
            s.colB[OV_CS * t] = s.b[t];
            s.colB[OV_CS * t + 1] = s.bT[t];
            s.colB[OV_CS * t + 2] = s.kbuf[t * NGAS + g1];
        }
        __syncwarp();
    }
    bool head_pending = lane != 0;   // the bin this lane starts in was opened by an earlier lane
    double edge = gord[ig < NG ? ig + 1 : NG];
    // software pipeline: the index and weight of the next element are fetched while this one is summed
    int pi = s.sidx[lane];
    double w = (pi >> 5) < NG ? W(pi >> 5, pi & 31) : 0.0;
#pragma unroll 1
    for (int r = 0; r < EPL; ++r) {
        const int i = pi >> 5, j = pi & 31;
        const double wc = w;
        const int pn = s.sidx[(r + 1 < EPL ? r + 1 : r) * 32 + lane];
        w = (pn >> 5) < NG ? W(pn >> 5, pn & 31) : 0.0;
        if (i < NG && ig < NG) {
            const double gdn = (r == EPL - 1) ? incl : __dadd_rn(prev, wc);
            if (gdn < edge) {
                acc[1] = __dadd_rn(acc[1], wc);
                if (GRAD) {
                    // row i = [tau_i, dT, gas columns] and column j = [b, bT, k, -] as double2; the pairs that
                    // hold only gases not folded yet (all zero) are not loaded
                    const double2 *row = reinterpret_cast<const double2 *>(s.dkp + i * DS);
                    const double2 *col = reinterpret_cast<const double2 *>(s.colB + OV_CS * j);
                    double2 rv[DS / 2];
#pragma unroll
                    for (int q = 0; q < DS / 2; ++q) rv[q] = (q == 0 || 2 * q - 2 <= igas) ? row[q] : make_double2(0.0, 0.0);
                    const double2 c0 = col[0];
                    const double ck = s.colB[OV_CS * j + 2];
                    acc[0] = __fma_rn(__dadd_rn(rv[0].x, c0.x), wc, acc[0]);
                    acc[2] = __fma_rn(__dadd_rn(rv[0].y, c0.y), wc, acc[2]);
#pragma unroll
                    for (int p = 0; p < NPMAX - 1; ++p)   // unfolded gases are 0
                        acc[3 + p] = __fma_rn((p & 1) ? rv[1 + p / 2].y : rv[1 + p / 2].x, wc, acc[3 + p]);
                    acc[NQ - 1] = __fma_rn(ck, wc, acc[NQ - 1]);
                } else {
                    acc[0] = __fma_rn(__dadd_rn(s.a[i], s.b[j]), wc, acc[0]);
                }
            } else {
                // straddler of edge ig+1: close the partial sum, leave the element for lane ig / ig+1
                double *dst = head_pending ? myhead : s.bsum + ig * BS;
#pragma unroll
                for (int q = 0; q < NQ; ++q) { dst[q] = acc[q]; acc[q] = 0.0; }
                head_pending = false;
                ++ig;
                s.strad[ig] = pi;
                s.frac[ig] = prev;
                s.gdn[ig] = gdn;
                edge = gord[ig < NG ? ig + 1 : NG];
            }
            prev = gdn;
        }
        pi = pn;
    }
    // a lane that never closed the bin it started in passes everything on as its head
    const bool has_tail = !head_pending;
    if (head_pending) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) myhead[q] = acc[q];
    }
    // lanes past the data count as closers so an open last bin stops there
    const unsigned closers = __ballot_sync(FULL, has_tail || lane * EPL >= NN);
    __syncwarp();
    if (has_tail && ig < NG) {
        // tail + heads of the following lanes up to (and including) the first lane that closed a bin
        for (int l2 = lane + 1; l2 < 32 && l2 * EPL < NN; ++l2) {
            const double *h = s.head + l2 * BS;
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = __dadd_rn(acc[q], h[q]);
            if ((closers >> l2) & 1u) break;
        }
        double *dst = s.bsum + ig * BS;
#pragma unroll
        for (int q = 0; q < NQ; ++q) dst[q] = acc[q];
    }
    __syncwarp();
    // lane m: bin m = whole elements + (1-frac) of the straddler of edge m + frac of the straddler of
    // edge m+1; normalised if closed by a straddler, or if it is the last bin and was opened (:6171-6172)
    const int m = lane;
    double res[NQ];
    bool norm = false;
    if (m < NG) {
        const int sp0 = m > 0 ? s.strad[m] : OV_NONE, sp1 = s.strad[m + 1];
        const bool opened = (m == 0) || sp0 != OV_NONE;
        norm = sp1 != OV_NONE || (m == NG - 1 && opened);
        const double *src = s.bsum + m * BS;
#pragma unroll
        for (int q = 0; q < NQ; ++q) res[q] = src[q];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int sp = side == 0 ? sp0 : sp1;
            if (sp != OV_NONE) {
                const int e = m + side;
                const double pv = s.frac[e], gd = s.gdn[e];
                // (gd - pv is the element's weight; exact for float32-born weights)
                const double frac = __ddiv_rn(__dsub_rn(gord[e], pv), __dsub_rn(gd, pv));
                const double f = side == 0 ? __dsub_rn(1.0, frac) : frac;
                const int i = sp >> 5, j = sp & 31;
                const double w = W(i, j);
                res[1] = __dadd_rn(res[1], __dmul_rn(f, w));
                if (GRAD) {
                    const double *row = s.dkp + i * DS;
                    const double *col = s.colB + OV_CS * j;
                    res[0] = __dadd_rn(res[0], __dmul_rn(f, __dmul_rn(__dadd_rn(row[0], col[0]), w)));
                    res[2] = __dadd_rn(res[2], __dmul_rn(f, __dmul_rn(__dadd_rn(row[1], col[1]), w)));
#pragma unroll
                    for (int p = 0; p < NPMAX - 1; ++p)
                        res[3 + p] = __dadd_rn(res[3 + p], __dmul_rn(f, __dmul_rn(row[2 + p], w)));
                    res[NQ - 1] = __dadd_rn(res[NQ - 1], __dmul_rn(f, __dmul_rn(col[2], w)));
                } else {
                    res[0] = __dadd_rn(res[0], __dmul_rn(f, __dmul_rn(__dadd_rn(s.a[i], s.b[j]), w)));
                }
            }
        }
    }
    __syncwarp();
    if (m < NG) {
        const double rs = norm ? __ddiv_rn(1.0, res[1]) : 1.0;
        s.a[m] = __dmul_rn(res[0], rs);
        if (GRAD) {
            double *row = s.dkp + m * DS;
            row[1] = __dmul_rn(res[2], rs);
#pragma unroll
            for (int p = 0; p < NPMAX - 1; ++p) {
                if (p < NGAS) {
                    const double g = (p == g1) ? res[NQ - 1] : res[3 + p];
                    row[2 + p] = p <= g1 ? __dmul_rn(g, rs) : g;
                }
            }
        }
    }
    __syncwarp();
}

// ---- order of EQUAL keys: numba's quicksort --------------------------------------------------------
// The reference sorts with numba's np.argsort, an unstable quicksort (numba/misc/quicksort.py:
// median-of-three pivot stashed at the end, Hoare partition, larger side pushed on an explicit stack,
// insertion sort below 15 elements, `a < b or (isnan(b) and not isnan(a))` as the order).  Which of
// several equal keys comes first is a property of that algorithm, and it decides how the gradient
// rows of tied elements are split across bin edges.  When a fold has tied keys and gradients are
// wanted, the same algorithm is replayed here by the whole warp so the permutation is the reference's:
//   * the Hoare partition is data-parallel: the k-th element from the left that is not < pivot swaps
//     with the k-th element from the right that is not > pivot while the former lies left of the
//     latter; both lists come from one scan over the range, and the pivot's final slot follows;
//   * disjoint ranges may be processed in any order, so small ranges are only recorded and the
//     (stable) insertion sorts are done for all positions at once at the end.
// Output: s.sidx[x] = packed (i,j) of sorted position x (blocked layout).
__device__ __forceinline__ double ov_keyof(const OvWarpSmem &s, int pk)
{
    return __dadd_rn(s.a[pk >> 5], s.b[pk & 31]);
}
__device__ __forceinline__ bool ov_lt(double a, double b) { return a < b || (isnan(b) && !isnan(a)); }

static __device__ __noinline__ void ov_numba_order(OvWarpSmem s, int NG, int lane)
{
    const int NN = NG * NG;
    int nnpad = 128;
    while (nnpad < NN) nnpad <<= 1;
    unsigned short *R = s.sidx;
    unsigned short *Lp = reinterpret_cast<unsigned short *>(s.head);
    unsigned short *Rp = Lp + nnpad;
    unsigned short *rlo = Rp + nnpad;
    unsigned short *rhi = rlo + nnpad;
    int *stk = s.strad;   // (low, high) pairs; depth <= log2(NN) because the larger side is pushed
    for (int x = lane; x < NN; x += 32) {
        R[x] = (unsigned short)(((x / NG) << 5) | (x % NG));
        rlo[x] = (unsigned short)x;
        rhi[x] = (unsigned short)x;
    }
    if (lane == 0) { stk[0] = 0; stk[1] = NN - 1; }
    __syncwarp();
    int ns = 1;
    while (ns > 0) {
        ns -= 1;
        int low = stk[2 * ns], high = stk[2 * ns + 1];
        __syncwarp();
        while (high - low >= 15) {
            const int mid = (low + high) >> 1;
            if (lane == 0) {
                unsigned short rl = R[low], rm = R[mid], rh = R[high], t;
                if (ov_lt(ov_keyof(s, rm), ov_keyof(s, rl))) { t = rl; rl = rm; rm = t; }
                if (ov_lt(ov_keyof(s, rh), ov_keyof(s, rm))) { t = rh; rh = rm; rm = t; }
                if (ov_lt(ov_keyof(s, rm), ov_keyof(s, rl))) { t = rl; rl = rm; rm = t; }
                R[low] = rl; R[mid] = rh; R[high] = rm;     // pivot (median) stashed at the end
            }
            __syncwarp();
            const double pivot = ov_keyof(s, R[high]);
            const int n = high - low;
            const int ch = (n + 31) >> 5;
            const int p0 = low + lane * ch, p1 = min(high, p0 + ch);
            int nl = 0, nr = 0;
            for (int p = p0; p < p1; ++p) {
                const double v = ov_keyof(s, R[p]);
                nl += !ov_lt(v, pivot);
                nr += !ov_lt(pivot, v);
            }
            int il = nl, ir = nr;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(FULL, il, d), v = __shfl_down_sync(FULL, ir, d);
                if (lane >= d) il += u;
                if (lane + d < 32) ir += v;
            }
            const int totL = __shfl_sync(FULL, il, 31), totR = __shfl_sync(FULL, ir, 0);
            int kl = il - nl, kr = ir - nr;
            for (int p = p0; p < p1; ++p) if (!ov_lt(ov_keyof(s, R[p]), pivot)) Lp[kl++] = (unsigned short)p;
            for (int p = p1 - 1; p >= p0; --p) if (!ov_lt(pivot, ov_keyof(s, R[p]))) Rp[kr++] = (unsigned short)p;
            __syncwarp();
            const int kmax = min(totL, totR);
            int cnt = 0;
            for (int k = lane; k < kmax; k += 32) cnt += Lp[k] < Rp[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
            const int K = cnt;
            int ifin;
            if (K >= 1) ifin = (K < totL && Lp[K] < Rp[K - 1]) ? Lp[K] : Rp[K - 1];
            else ifin = totL > 0 ? Lp[0] : high;
            for (int k = lane; k < K; k += 32) {
                const int x = Lp[k], y = Rp[k];
                const unsigned short t = R[x]; R[x] = R[y]; R[y] = t;
            }
            __syncwarp();
            if (lane == 0) { const unsigned short t = R[ifin]; R[ifin] = R[high]; R[high] = t; }
            __syncwarp();
            if (high - ifin > ifin - low) {
                if (high > ifin) { if (lane == 0) { stk[2 * ns] = ifin + 1; stk[2 * ns + 1] = high; } ns += 1; }
                high = ifin - 1;
            } else {
                if (ifin > low) { if (lane == 0) { stk[2 * ns] = low; stk[2 * ns + 1] = ifin - 1; } ns += 1; }
                low = ifin + 1;
            }
            __syncwarp();
        }
        if (high > low)
            for (int x = low + lane; x <= high; x += 32) { rlo[x] = (unsigned short)low; rhi[x] = (unsigned short)high; }
        __syncwarp();
    }
    // stable insertion sort of every small range: rank inside the range
    for (int x0 = 0; x0 < NN; x0 += 32) {
        const int x = x0 + lane;
        if (x < NN) {
            const int lo = rlo[x], hi = rhi[x];
            const double v = ov_keyof(s, R[x]);
            int rank = 0;
            for (int y = lo; y <= hi; ++y) {
                const double u = ov_keyof(s, R[y]);
                rank += ov_lt(u, v) || (y < x && !ov_lt(v, u));
            }
            Lp[lo + rank] = R[x];
        }
    }
    __syncwarp();
    for (int x = lane; x < NN; x += 32) R[x] = Lp[x];
    __syncwarp();
}

// ---- data-independent orders ----------------------------------------------------------------------
// Row-major: the next gas is ascending and too weak to reach the next row, key(i,NG-1) <= key(i+1,0)
// (equal keys already sit in index order).  Column-major: the running opacity is ascending and too weak
// to reach the next column, key(NG-1,j) < key(0,j+1) (strict: an equal pair would be in the wrong index
// order).  Both tests use the exact keys.  Returns 1 (row-major) / 2 (column-major) / 0, plus 4 if WANT_STRICT
// and the whole key sequence is strictly increasing (no ties, so no question about the order of equal keys).
template <bool WANT_STRICT>
__device__ __forceinline__ int ov_trivial_order(const OvWarpSmem &s, int NG, int lane)
{
    bool rowok = true, colok = true;
    if (lane < NG - 1) {
        const double b0 = s.b[0], bl = s.b[NG - 1], a0 = s.a[0], al = s.a[NG - 1];
        rowok = (s.b[lane] <= s.b[lane + 1]) & (__dadd_rn(s.a[lane], bl) <= __dadd_rn(s.a[lane + 1], b0));
        colok = (s.a[lane] <= s.a[lane + 1]) & (__dadd_rn(al, s.b[lane]) < __dadd_rn(a0, s.b[lane + 1]));
    }
    rowok = __all_sync(FULL, rowok);
    colok = !rowok && __all_sync(FULL, colok);
    int ord = rowok ? 1 : (colok ? 2 : 0);
    if (WANT_STRICT && ord) {
        bool strict = true;
        if (lane < NG) {
            // row-major: lane i walks row i; column-major: lane j walks column j
            const double mine = rowok ? s.a[lane] : s.b[lane];
            const double *other = rowok ? s.b : s.a;
            double kp = __dadd_rn(mine, other[0]);
            for (int t = 1; t < NG; ++t) {
                const double kn = __dadd_rn(mine, other[t]);
                strict &= kp < kn;
                kp = kn;
            }
            if (rowok && lane < NG - 1) strict &= kp < __dadd_rn(s.a[lane + 1], s.b[0]);
        }
        if (__all_sync(FULL, strict)) ord |= 4;
    }
    return ord;
}

// When the order of the NG*NG elements is row-major or column-major, the cumulative weights, the element
// that straddles every bin edge and its `frac` depend on the quadrature only.  The rebin
// (rank/rankg, ForwardModel_0.py:6155-6172 / :6002-6025) is then a fixed linear map:
//     bin m = [ sum_i RA[i][m] x_i + sum_j RB[j][m] y_j ] / SW[m]
// with RA[i][m] (RB[j][m]) the weight -- fractions included -- that row i (column j) contributes to bin m,
// x_i the per-row operands (tau_i, dT_i, gas columns) and y_j the per-column ones (b_j, bT_j, k_j).
// ov_static_setup builds the tables once per CTA by replaying the reference's loop over the fixed order
// (thread (order, m) keeps the terms of bin m); `ok` is false if some bin would be left un-normalised
// (degenerate quadrature), in which case the shortcut is not used.
__device__ __forceinline__ bool ov_static_setup(double *stat, const double *wtab, const double *gord, int NG)
{
    const int NN = NG * NG;
    for (int e = threadIdx.x; e < ov_static_doubles(NG); e += blockDim.x) stat[e] = 0.0;
    __syncthreads();
    bool ok = true;
    if ((int)threadIdx.x < 2 * NG) {
        const int o = threadIdx.x / NG, m = threadIdx.x - o * NG;
        double *RA = stat + o * 2 * NN, *RB = RA + NN, *SW = stat + 4 * NN + o * NG;
        double run = 0.0, sw = 0.0;
        int ig = 0;
        bool closed = false, opened = (m == 0);
        for (int q = 0; q < NG && ig < NG; ++q) {
            for (int r = 0; r < NG && ig < NG; ++r) {
                const int i = o == 0 ? q : r, j = o == 0 ? r : q;
                const double w = wtab[i * NG + j];
                const double gdn = __dadd_rn(run, w);
                double f = 0.0;
                if (gdn < gord[ig + 1]) {
                    if (ig == m) f = w;
                } else {
                    const double frac = __ddiv_rn(__dsub_rn(gord[ig + 1], run), __dsub_rn(gdn, run));
                    if (ig == m) { f = __dmul_rn(frac, w); closed = true; }
                    ++ig;
                    if (ig == m) { f = __dmul_rn(__dsub_rn(1.0, frac), w); opened = true; }
                }
                if (f != 0.0) {
                    RA[i * NG + m] = __dadd_rn(RA[i * NG + m], f);
                    RB[j * NG + m] = __dadd_rn(RB[j * NG + m], f);
                    sw = __dadd_rn(sw, f);
                }
                run = gdn;
            }
        }
        SW[m] = sw;
        ok = opened && (closed || m == NG - 1) && sw > 0.0;
    }
    return __syncthreads_and(ok) != 0;
}

template <int NPMAX, bool GRAD>
__device__ __forceinline__ void ov_rebin_static(const OvWarpSmem &s, const double *__restrict__ stat, int ord,
                                                int NG, int NGAS, int igas, int lane)
{
    constexpr int DS = NPMAX + 2;
    const int NN = NG * NG, g1 = igas + 1, o = (ord & 3) - 1, m = lane;
    const double *RA = stat + o * 2 * NN, *RB = RA + NN;
    double ra_a = 0.0, rT = 0.0, rk = 0.0;
    double rg[GRAD ? NPMAX - 1 : 1];
#pragma unroll
    for (int p = 0; p < (GRAD ? NPMAX - 1 : 1); ++p) rg[p] = 0.0;
    if (m < NG) {
#pragma unroll 2
        for (int t = 0; t < NG; ++t) {
            const double ra = RA[t * NG + m], rb = RB[t * NG + m];
            ra_a = __fma_rn(ra, s.a[t], ra_a);
            ra_a = __fma_rn(rb, s.b[t], ra_a);
            if (GRAD) {
                const double2 *row = reinterpret_cast<const double2 *>(s.dkp + t * DS);
                rT = __fma_rn(ra, row[0].y, rT);
                rT = __fma_rn(rb, s.bT[t], rT);
                rk = __fma_rn(rb, s.kbuf[t * NGAS + g1], rk);
#pragma unroll
                for (int q = 1; q < DS / 2; ++q) {
                    if (2 * q - 2 <= igas) {
                        const double2 v = row[q];
                        rg[2 * q - 2] = __fma_rn(ra, v.x, rg[2 * q - 2]);
                        if (2 * q - 1 < NPMAX - 1) rg[2 * q - 1] = __fma_rn(ra, v.y, rg[2 * q - 1]);
                    }
                }
            }
        }
    }
    __syncwarp();
    if (m < NG) {
        const double rs = __ddiv_rn(1.0, stat[4 * NN + o * NG + m]);
        s.a[m] = __dmul_rn(ra_a, rs);
        if (GRAD) {
            double *row = s.dkp + m * DS;
            row[1] = __dmul_rn(rT, rs);
#pragma unroll
            for (int p = 0; p < NPMAX - 1; ++p) {
                if (p < NGAS) row[2 + p] = p <= igas ? __dmul_rn(rg[p], rs) : (p == g1 ? __dmul_rn(rk, rs) : 0.0);
            }
        }
    }
    __syncwarp();
}

// One sort/rebin fold.  a[] holds the running tau_g, b[] the next gas.  Gradient storage: dkp[i][p] is
// d tau_i / d amount_p for p < NGAS and dkp[i][NGAS] is d tau_i / dT at every stage (the reference keeps
// the temperature column at index igas+1 and moves it one to the right per fold, ForwardModel_0.py
// :5931, :5948; the values are the same).  The gradient row of element (i,j) is
// { dkp[i][0..igas], kbuf[j][g1] } for the gases and dkp[i][NGAS] + bT[j] for T (:5946-5949).
template <int EPL, int NPMAX, bool GRAD>
__device__ __forceinline__ void ov_sort_stage(const OvWarpSmem &s, const double *__restrict__ wtab,
                                        const double *__restrict__ gord, int NG, int NGAS, int igas, int lane,
                                        int seq_rebin, int ord)
{
    const int NN = NG * NG;
    const int NP1 = NGAS + 1;
    const int g1 = igas + 1;
    double key[EPL];
    int idx[EPL];
    auto make_keys = [&]() {
        int e = lane * EPL;
        int i = e / NG, j = e - i * NG;
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
            if (e < NN) {
                key[r] = __dadd_rn(s.a[i], s.b[j]);
                idx[r] = (i << 5) | j;
            } else {
                key[r] = INFINITY;
                idx[r] = (32 << 5) + (e - NN);   // distinct, above every live index
            }
            ++e;
            if (++j == NG) { j = 0; ++i; }
        }
    };
    // trivial orders (ov_trivial_order) need no sort
    bool sorted = false, keys_valid = false;
    int chk = -1;
    {
        const bool rowok = (ord & 3) == 1, colok = (ord & 3) == 2;
        if (rowok | colok) {
            int el = lane * EPL;
            int q = el / NG, rm = el - q * NG;      // el = q*NG + rm
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                if (el < NN) idx[r] = rowok ? ((q << 5) | rm) : ((rm << 5) | q);
                else idx[r] = (32 << 5) + (el - NN);
                ++el;
                if (++rm == NG) { rm = 0; ++q; }
            }
            sorted = true;
        }
    }
    // fast path: float32-key packed network, verified against the exact keys
    if (!sorted) {
        const double kmax = __dadd_rn(s.a[NG - 1], s.b[NG - 1]);
        const int e = (__double2hiint(kmax) >> 20) & 0x7ff;
        // (vote: every lane holds the same value; this tells the compiler the branch is warp-uniform)
        if (__all_sync(FULL, kmax > 0.0 && e > 200 && e < 2000)) {
            const double sc = __hiloint2double((2146 - e) << 20, 0);   // 2^(100 - exponent(kmax)): exact scaling
            unsigned v[EPL];
            unsigned orb = 0u, mxb = 0u;
            int el = lane * EPL;
            int i = el / NG, j = el - i * NG;
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                if (el < NN) {
                    // positive keys only: the float32 bit pattern is then monotone (other keys end up misplaced
                    // and are caught by the exact check below)
                    const float kf = __double2float_rn(__dmul_rn(__dadd_rn(s.a[i], s.b[j]), sc));
                    const unsigned kb = __float_as_uint(kf);
                    orb |= kb;
                    mxb = max(mxb, kb);
                    // 22 key bits = 6 exponent bits (kmax sits at 2^100: the 62 binades below it keep their own
                    // exponent, anything smaller collapses to 0 and is then told apart by the exact check) and
                    // 16 mantissa bits
                    const unsigned kt = kb > OV_KEY_BASE ? kb - OV_KEY_BASE : 0u;
                    v[r] = ((kt >> 7) << 10) | (unsigned)((i << 5) | j);
                } else {
                    v[r] = 0xffffffffu;           // padding: above every live element, index field marks it dead
                }
                ++el;
                if (++j == NG) { j = 0; ++i; }
            }
            ov_bitonic_sort_u32<EPL>(v, lane);
            // If every key is a finite positive float and no two neighbours share their 22 key bits, the
            // float32-rounded keys are strictly increasing, hence so are the exact keys (rounding is
            // monotone): the order is exact and free of ties, no need to re-form the float64 keys.
            // (a key beyond the 6-bit exponent range -- only possible when kmax is not the largest key -- or a
            // negative / non-finite one makes the packed order meaningless: verify exactly)
            bool dirty = (orb >> 31) != 0u || mxb >= OV_KEY_BASE + (64u << 23);
#pragma unroll
            for (int r = 0; r + 1 < EPL; ++r) dirty |= ((v[r] | 1023u) >= v[r + 1]) & (v[r + 1] < 0xfffffc00u);
            {
                const unsigned nx = __shfl_down_sync(FULL, v[0], 1);
                dirty |= (lane < 31) & ((v[EPL - 1] | 1023u) >= nx) & (nx < 0xfffffc00u);
            }
#pragma unroll
            for (int r = 0; r < EPL; ++r) idx[r] = (int)(v[r] & 1023u);
            if (!__any_sync(FULL, dirty)) {
                sorted = true;
                chk = 0;
            } else {
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const int ii = idx[r] >> 5, jj = idx[r] & 31;
                    key[r] = ii < NG ? __dadd_rn(s.a[ii], s.b[jj]) : INFINITY;
                }
                chk = ov_check_order<EPL>(key, lane);
                for (int pass = 0; (chk & 1) && pass < 6; ++pass) {
                    ov_fixup_pass<EPL>(key, idx, lane);
                    chk = ov_check_order<EPL>(key, lane);
                }
                sorted = !(chk & 1);
                keys_valid = sorted;
            }
        }
    }
    if (!sorted) {
        // keys outside the scaled float32 range or a large group of near-equal keys: exact network
        make_keys();
        ov_bitonic_sort<EPL, true>(key, idx, lane);
        keys_valid = true;
        chk = -1;
    }

    if (GRAD) {
        // equal keys: take the order numba's quicksort would produce (gradient rows of tied elements
        // are split across bin edges in sort order; tau does not depend on it).  The exact keys and
        // the neighbour check of the fast path are reused when they are still valid.
        if (chk < 0) {
            if (!keys_valid) {
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const int ii = idx[r] >> 5, jj = idx[r] & 31;
                    key[r] = ii < NG ? __dadd_rn(s.a[ii], s.b[jj]) : INFINITY;
                }
            }
            chk = ov_check_order<EPL>(key, lane);
        }
        if (chk & 2) {
            ov_numba_order(s, NG, lane);
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                const int pos = lane * EPL + r;
                idx[r] = pos < NN ? (int)s.sidx[pos] : (32 << 5) + (pos - NN);
            }
            __syncwarp();
        }
    }
    // stage the sorted order for the rebin (blocked layout for the sequential walk, transposed for the
    // parallel one)
    if (seq_rebin) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
            const int pos = lane * EPL + r;
            if (pos < NN) s.sidx[pos] = (unsigned short)idx[r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < EPL; ++r) s.sidx[r * 32 + lane] = (unsigned short)idx[r];
    }
    __syncwarp();
}

template <int EPL, int NPMAX, bool GRAD>
__device__ __forceinline__ void ov_rebin(const OvWarpSmem &s, const OvWeight &W,
                                         const double *__restrict__ gord, int NG, int NGAS, int igas, int lane,
                                         int seq_rebin)
{
    if (seq_rebin) ov_rebin_seq<NPMAX, GRAD>(s, W.wtab, gord, NG, NGAS, igas, lane);
    else ov_rebin_par<EPL, NPMAX, GRAD>(s, W, gord, NG, NGAS, igas, lane);
}

template <int EPL, int NPMAX, bool GRAD>
__global__ void __launch_bounds__(OV_WARPS * 32, GRAD ? OV_MINB : OV_MINB_NOGRAD)
ans_koverlap_kernel(OvParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NG = P.NG, NGAS = P.NGAS, NLAY = P.NLAY, NN = NG * NG, NP1 = NGAS + 1;
    constexpr int DS = NPMAX + 2, BS = (NPMAX + 3) | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long ncell = (long long)P.NWAVE * NLAY;
    const int *cell_list = nullptr;
    if (P.cell_count) {
        const int c = *P.cell_count;
        if (c >= 0) { ncell = c; cell_list = P.cell_list; }
        if ((long long)blockIdx.x * OV_WARPS >= ncell) return;     // (before the set-up: usually nothing is left)
    }

    double *wtab = reinterpret_cast<double *>(smem_raw);
    double *gord = wtab + NN;
    float *dgf = reinterpret_cast<float *>(gord + NG + 1);
    double *stat = wtab + ((NN + NG + 2 + (NG + 1) / 2) & ~1);      // static-order rebin tables (ov_static_doubles)
    double *wbase = stat + ov_static_doubles(NG);                   // keeps the per-warp regions 16-byte aligned
    // per-warp carve-up (doubles first, then ints, then shorts)
    const size_t per_warp_bytes = ov_per_warp_bytes(NG, NGAS, GRAD);
    unsigned char *mine = reinterpret_cast<unsigned char *>(wbase) + per_warp_bytes * warp;
    OvWarpSmem s;
    {
        double *d = reinterpret_cast<double *>(mine);
        // (everything whose offset depends on the run-time NGAS comes last)
        s.colB = d; d += OV_CS * NG;                       // (16-byte aligned: first in the region)
        s.dkp = d; if (GRAD) d += NG * DS;
        s.a = d; d += NG;
        s.b = d; d += NG;
        s.bT = d; d += NG;
        s.frac = d; d += NG + 1;
        s.gdn = d; d += NG + 1;
        s.bsum = d; d += NG * BS;
        s.head = d; d += ov_head_doubles_np(NG, NPMAX, GRAD);
        s.strad = reinterpret_cast<int *>(d);
        s.spare = s.strad + NG + 1;
        int nnpad = 128;
        while (nnpad < NN) nnpad <<= 1;
        s.sidx = reinterpret_cast<unsigned short *>(s.spare + NG + ((2 * NG + 1) & 1));   // (8-byte multiple of ints)
        d = reinterpret_cast<double *>(s.sidx + nnpad);
        s.kbuf = d; d += NG * NGAS;
        s.dbuf = d;
    }
    for (int i = threadIdx.x; i < NN; i += blockDim.x) wtab[i] = P.weight[i];
    for (int i = threadIdx.x; i <= NG; i += blockDim.x) gord[i] = P.g_ord[i];
    OvWeight W{wtab, dgf, NG, false};
    {
        // can del_g[i]*del_g[j] be formed on the fly as a float32 product?  (checked against the table)
        bool ok = P.del_g != nullptr;
        if (ok) {
            for (int i = threadIdx.x; i < NG; i += blockDim.x) dgf[i] = (float)P.del_g[i];
            for (int e = threadIdx.x; e < NN; e += blockDim.x) {
                const int i = e / NG, j = e - i * NG;
                const double di = P.del_g[i], dj = P.del_g[j];
                ok = ok && (double)(float)di == di && (double)__fmul_rn((float)di, (float)dj) == P.weight[e];
            }
        }
        W.f32 = __syncthreads_and(ok) != 0;
    }

    // data-independent rebin of row-/column-major folds (not with the literal sequential scan)
    const bool static_ok = ov_static_setup(stat, wtab, gord, NG) && !P.seq_rebin;

    // persistent CTAs: every warp walks the (wavenumber, layer) cells with the grid's stride
    for (long long cell0 = (long long)blockIdx.x * OV_WARPS; cell0 < ncell; cell0 += (long long)gridDim.x * OV_WARPS) {
    long long cell = cell0 + warp;
    const bool live = cell < ncell;        // idle warps of the last round shadow the last cell (they join the barriers)
    if (!live) cell = ncell - 1;
    if (cell_list) cell = cell_list[cell];
    const int iw = (int)(cell / NLAY);
    const int l = (int)(cell - (long long)iw * NLAY);

    // stage k (and dk/dT) of the cell
    if (P.fused) {
        // plane-major table: the NG*NGAS values of this wavenumber in one (p,T) plane are contiguous
        const size_t plane = (size_t)P.NWAVE * NG * NGAS;
        const size_t toff = ((size_t)__ldg(P.plan.ip_lo + l) * P.NT + __ldg(P.plan.it_lo + l)) * plane +
                            (size_t)iw * NG * NGAS;
        const double *w = P.plan.w4 + 4 * l;
        const double w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        const double omv = GRAD ? __ldg(P.plan.omv + l) : 0.0, v = GRAD ? __ldg(P.plan.vv + l) : 0.0,
                     dudt = GRAD ? __ldg(P.plan.dudt + l) : 0.0;
        for (int e = lane; e < NG * NGAS; e += 32) {
            double kv, dv = 0.0;
            ans_kinterp_elem<GRAD>(P.tab, toff + e, P.NT, plane, w0, w1, w2, w3, omv, v, dudt, kv, dv);
            s.kbuf[e] = kv;
            if (GRAD) s.dbuf[e] = dv;
        }
    } else {
        for (int e = lane; e < NG * NGAS; e += 32) {
            const int g = e / NGAS, gas = e - g * NGAS;
            const size_t o = (((size_t)iw * NG + g) * NLAY + l) * NGAS + gas;
            s.kbuf[e] = __ldg(P.k + o);
            if (GRAD) s.dbuf[e] = __ldg(P.dkdT + o);
        }
    }
    for (int i = lane; i < NG; i += 32) s.a[i] = 0.0;
    if (GRAD) for (int i = lane; i < NG * DS; i += 32) s.dkp[i] = 0.0;
    __syncwarp();

#define KB(g, gas) s.kbuf[(g) * NGAS + (gas)]
#define DB(g, gas) s.dbuf[(g) * NGAS + (gas)]
    for (int igas = 0; igas < NGAS - 1; ++igas) {
        // keep the CTA's warps in the same phase of the code: the hot path is larger than the
        // instruction cache and warps that drift apart evict each other's lines
#ifndef OV_NO_FOLDBAR
        // (measured at config 2: the gradient kernel loses 3 / 10 / 21 % with a barrier every 2nd / 4th /
        // no fold; the smaller no-gradient kernel gains 4 % with one every 4th fold)
        if (GRAD || (igas & 3) == 0) __syncthreads();
#endif
        const int g1 = igas + 1;
        const double am1 = __ldg(P.amount + (size_t)g1 * NLAY + l);
        const bool next_neg = __all_sync(FULL, __dmul_rn(KB(NG - 1, g1), am1) <= 0.0);   // uniform by construction
        bool do_fold = false;
        if (igas == 0) {
            const double am0 = __ldg(P.amount + l);
            const bool first_neg = __all_sync(FULL, __dmul_rn(KB(NG - 1, 0), am0) <= 0.0);
            if (first_neg) {
                for (int i = lane; i < NG; i += 32) {
                    s.a[i] = __dmul_rn(KB(i, 1), am1);
                    if (GRAD) { s.dkp[i * DS + 3] = KB(i, 1); s.dkp[i * DS + 1] = __dmul_rn(DB(i, 1), am1); }
                }
                __syncwarp();
            } else if (next_neg) {
                for (int i = lane; i < NG; i += 32) {
                    s.a[i] = __dmul_rn(KB(i, 0), am0);
                    if (GRAD) { s.dkp[i * DS + 2] = KB(i, 0); s.dkp[i * DS + 1] = __dmul_rn(DB(i, 0), am0); }
                }
                __syncwarp();
            } else {
                for (int i = lane; i < NG; i += 32) {
                    s.a[i] = __dmul_rn(KB(i, 0), am0);
                    s.b[i] = __dmul_rn(KB(i, 1), am1);
                    if (GRAD) {
                        s.dkp[i * DS + 2] = KB(i, 0);
                        s.dkp[i * DS + 1] = __dmul_rn(DB(i, 0), am0);
                        s.bT[i] = __dmul_rn(DB(i, 1), am1);
                    }
                }
                do_fold = true;
            }
        } else {
            if (next_neg) {
                if (GRAD) {
                    for (int i = lane; i < NG; i += 32) {
                        // reference: dk[:,igas+2] = dk[:,igas+1]; dk[:,igas+1] *= 0  (T column moves, gas column = 0*T)
                        s.dkp[i * DS + 2 + g1] = __dmul_rn(s.dkp[i * DS + 1], 0.0);
                    }
                    __syncwarp();
                }
            } else if (__all_sync(FULL, s.a[NG - 1] <= 0.0)) {
                __syncwarp();
                for (int i = lane; i < NG; i += 32) {
                    s.a[i] = __dmul_rn(KB(i, g1), am1);
                    if (GRAD) { s.dkp[i * DS + 2 + g1] = KB(i, g1); s.dkp[i * DS + 1] = __dmul_rn(DB(i, g1), am1); }
                }
                __syncwarp();
            } else {
                for (int i = lane; i < NG; i += 32) {
                    s.b[i] = __dmul_rn(KB(i, g1), am1);
                    if (GRAD) s.bT[i] = __dmul_rn(DB(i, g1), am1);
                }
                do_fold = true;
            }
        }
        if (do_fold) {
            __syncwarp();
            const int ord = ov_trivial_order<GRAD>(s, NG, lane);
            if (static_ok && (ord & 3) && (!GRAD || (ord & 4))) {
                ov_rebin_static<NPMAX, GRAD>(s, stat, ord, NG, NGAS, igas, lane);
            } else {
                ov_sort_stage<EPL, NPMAX, GRAD>(s, wtab, gord, NG, NGAS, igas, lane, P.seq_rebin, ord);
                ov_rebin<EPL, NPMAX, GRAD>(s, W, gord, NG, NGAS, igas, lane, P.seq_rebin);
            }
        }
    }
#undef KB
#undef DB
    for (int g = lane; live && g < NG; g += 32) {
        const size_t o = ((size_t)iw * NG + g) * NLAY + l;
        P.tau[o] = s.a[g];
        if (GRAD) {
            for (int p = 0; p < NGAS; ++p) P.dk[o * NP1 + p] = s.dkp[g * DS + 2 + p];
            P.dk[o * NP1 + NGAS] = s.dkp[g * DS + 1];
        }
    }
    __syncwarp();
    }   // cells
}

template <int EPL, int NPMAX, bool GRAD>
inline int ov_launch(const OvParams &P, cudaStream_t stream)
{
    const int NG = P.NG, NGAS = P.NGAS, NN = NG * NG, NP1 = NGAS + 1;
    const size_t per_warp_bytes = ov_per_warp_bytes(NG, NGAS, GRAD);
    const size_t smem = (size_t)(((NN + NG + 2 + (NG + 1) / 2) & ~1) + ov_static_doubles(NG)) * 8 + per_warp_bytes * OV_WARPS + 16;
    auto kern = ans_koverlap_kernel<EPL, NPMAX, GRAD>;
    if (smem > 48 * 1024) ANS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ncell = (long long)P.NWAVE * P.NLAY;
    // persistent: as many CTAs as are resident at once (the occupancy API accounts for registers and smem)
    int dev = 0, nsm = 148, per_sm = 1;
    ANS_CUDA_CHECK(cudaGetDevice(&dev));
    ANS_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    ANS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, OV_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = ans_div_up(ncell, OV_WARPS);
    if (grid > (long long)nsm * per_sm) grid = (long long)nsm * per_sm;
    kern<<<(unsigned)grid, OV_WARPS * 32, smem, stream>>>(P);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

template <int EPL>
int ov_dispatch_np(const OvParams &P, bool grad, cudaStream_t stream)
{
    if (!grad) return ov_launch<EPL, 1, false>(P, stream);
    const int NP1 = P.NGAS + 1;
    if (NP1 <= 4) return ov_launch<EPL, 4, true>(P, stream);
    if (NP1 <= 8) return ov_launch<EPL, 8, true>(P, stream);
    return ov_launch<EPL, 16, true>(P, stream);
}

int ov_dispatch_4(const OvParams &P, bool grad, cudaStream_t stream);
int ov_dispatch_8(const OvParams &P, bool grad, cudaStream_t stream);
int ov_dispatch_16(const OvParams &P, bool grad, cudaStream_t stream);
