// koverlap_fast.cu -- random-overlap gas mixing, the common case, as "sort + marginal matrices".
//
// Reference: k_overlapg + rankg / k_overlap + rank (archnemesis/ForwardModel_0.py:5842-6026, :6029-6173).
//
// What rankg computes for bin m is  sum_e omega_m(e) x_e / sum_e omega_m(e),  where e runs over the NG*NG
// elements (i,j) of the key matrix  tau_i + k_j*amount,  omega_m(e) is the part of e's weight
// del_g[i]*del_g[j] that falls between the bin edges g_ord[m], g_ord[m+1] on the cumulative-weight axis of
// the key-sorted sequence, and x_e is the key or a gradient entry.  Every x_e is a row term plus a column
// term (key = a_i + b_j; d/d amount of an already folded gas = dk_i; of the gas being folded = k_j;
// d/dT = dT_i + bT_j), so
//     bin m = [ sum_i R[i][m] X[i][:] + sum_j C[j][m] Y[j][:] ] / sum_i R[i][m]
// with the marginals  R[i][m] = sum_j omega_m(i,j),  C[j][m] = sum_i omega_m(i,j); the denominator is the weight
// that falls into bin m, the same in every order of the keys (KfShared::rwid).  The bins, as sets, and
// the element that straddles each edge depend on the ORDER of the keys only where an edge falls, so the
// sort needs to be exact only there:
//   1. keys are packed as (23 key bits | 9-bit element number) -- the exponent bits the fold's keys need (at most 6:
//      62 binades below the largest key) and the 17 to 20 leading mantissa bits that fit beside them, cut straight
//      out of the float64 pattern (truncation is monotone) -- and
//      sorted by the 32-bit min/max bitonic network of koverlap_impl.cuh, in registers;
//   2. cumulative weights: lane-local sums + one warp scan (exact: float32-born weights); every element
//      stores the bin it starts in (one byte), the straddler of every edge its position and the cumulative
//      weight before it;
//   3. lane m checks that the straddler of edge m has key bits of its own (then everything before it is
//      strictly smaller and everything after strictly larger, whatever the order inside groups of equal
//      key bits and whatever order numba's unstable sort would give equal keys); a pair of elements with
//      equal key bits around an edge is put right with the exact float64 keys; anything else (larger
//      groups, exactly equal keys on an edge, non-monotone inputs, ...) sends the CELL to the work list of
//      the general kernel (ans_koverlap_kernel), which replays numba's tie order;
//   4. R and C are built by lane-per-row / lane-per-column walks over the bin bytes (run lengths, no
//      atomics), the straddlers' (1-frac) parts are moved to the next bin, and the two small products
//      above are lane-per-bin FMA loops with broadcast operand rows.
// Row-/column-major key orders (one gas far weaker than the other) have data-independent marginals, made
// once per CTA; they are used when the static straddlers are strictly separated from their neighbours.
//
// One warp per (wavenumber, layer) cell, persistent CTAs, no CTA barrier after the set-up.  Agreement with
// the reference ~1e-15 (another summation order); the literal sequential scan (force_seq) stays with the
// general kernel.
#include "koverlap_impl.cuh"

namespace {

constexpr int KF_NONE = 0xffff;
#ifndef KF_PARTIAL
#define KF_PARTIAL 1        // 0: every sorted fold sorts the whole key matrix
#endif
#ifndef KF_CHECK_WALK
#define KF_CHECK_WALK 0      // 1: assert in the marginal walks that bins never decrease along a row / column
#endif
#ifndef KF_SKIP
#define KF_SKIP 1           // 0: all merge levels even for a small head
#endif

template <int NG>
struct KfShared {
    double gord[NG + 2];            // bin edges, gord[NG+1] = +inf
    double wtabd[NG * NG];          // element weights (float32 products widened)
    float wtabf[NG * NG + 16];      // the same as float32; entry NG*NG = 0 is the padding of the sort
    double stat[2][2][NG * NG];     // [order][rows | columns][m*NG + i]  marginals of the static orders
    unsigned short sstr[2][NG + 1][4];   // [order][edge] -> element before / the straddler / element after
    double scw[NG + 2];             // row-major order: the (1-frac) part of the weight of edge m's straddler
    double sbin[(NG * NG + 16) / 8];   // row-major order: the bin every element starts in (bytes; entry NG*NG: padding)
    double rwid[NG + 4];            // 1 / weight of bin m (the sum over a bin's marginals in any order)
    unsigned char nedge[NG + 4];    // number of edges whose row-major straddler lies in rows < i
    int ok_f32, ok_static;
};

// per-warp shared memory, in doubles (every block a multiple of 16 bytes)
template <int NG, int XS, int KG, bool GRAD>
struct KfWarpLayout {
    static constexpr int NGASMAX = KG;
    static constexpr int RC = 0;                                  // [NG*NG] marginals; aliased by the sorted words
    static constexpr int X = RC + NG * NG;                        // [NG][XS] rows {tau_i, dT_i, gas columns}
    static constexpr int BT = X + NG * XS;                        // [NG] dk/dT * amount of the gas being folded
    static constexpr int AV = BT + NG;                            // [NG] running tau (copy of X[:,0] for the key loop)
    static constexpr int BV = AV + NG;                            // [NG] k * amount of the gas being folded
    static constexpr int KBUF = BV + NG;                          // [NG*NGASMAX]
    static constexpr int DBUF = KBUF + NG * NGASMAX;
    static constexpr int GBS = DBUF + (GRAD ? NG * NGASMAX : 0);   // [NG+2] cumulative weight before the straddler
    static constexpr int SPOS = GBS + NG + 2;                     // [NG+2] ints: sorted position of the straddler
    static constexpr int BIN = SPOS + (NG + 2 + 1) / 2 + 1;       // [NG*NG + 16] bytes (entry NG*NG: padding)
    static constexpr int TOTAL = (BIN + (NG * NG + 16) / 8 + 1) & ~1;
};

// 32-bit min/max bitonic network over 32*16 packed words, blocked layout (element = 16*lane + r), all-ascending form
// (koverlap_impl.cuh).  Every lane arrives with a BITONIC sequence (the key loop lays the tail of one row of the key
// matrix ascending and the head of the next one descending), so the in-lane phases reduce to one 4-stage merge.
// Cross-lane comparators: "take the partner's word if (partner < mine) != keep_high" is one ISETP.LT.XOR + SEL.
__device__ __forceinline__ void kf_sort(unsigned (&v)[16], int lane, int klmax)   // @phase sort
{
#define KF_CE(x, y) do { const unsigned lo__ = min(x, y), hi__ = max(x, y); x = lo__; y = hi__; } while (0)
#pragma unroll
    for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if ((r & j) == 0) KF_CE(v[r], v[r | j]);
    }
#pragma unroll 1
    for (int kl = 2; kl <= klmax; kl <<= 1) {
        {
            const bool keep_high = (lane & (kl >> 1)) != 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int q = 15 - r;
                const unsigned pr = __shfl_xor_sync(FULL, v[q], kl - 1), pq = __shfl_xor_sync(FULL, v[r], kl - 1);
                v[r] = ((pr < v[r]) != keep_high) ? pr : v[r];
                v[q] = ((pq < v[q]) != keep_high) ? pq : v[q];
            }
        }
#pragma unroll 1
        for (int lm = kl >> 2; lm > 0; lm >>= 1) {
            const bool keep_high = (lane & lm) != 0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const unsigned pv = __shfl_xor_sync(FULL, v[r], lm);
                v[r] = ((pv < v[r]) != keep_high) ? pv : v[r];
            }
        }
#pragma unroll
        for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if ((r & j) == 0) KF_CE(v[r], v[r | j]);
        }
    }
#undef KF_CE
}

__device__ __forceinline__ int kf_vaddr(int p) { return ((((p & 15) >> 2) * 32 + (p >> 4)) << 2) + (p & 3); }   // @phase resolve

// Static orders: replay rankg's loop over the row-major (o = 0) / column-major (o = 1) sequence once.
template <int NG>
__device__ void kf_static_setup(KfShared<NG> &S)   // @phase cta_setup
{
    const int o = threadIdx.x;
    if (o < 2) {
        double *RA = S.stat[o][0], *RB = S.stat[o][1];
        double run = 0.0;
        int ig = 0, prev = KF_NONE, pending = -1;
        bool ok = true;
        unsigned char *sbin = reinterpret_cast<unsigned char *>(S.sbin);
        for (int m = 0; m <= NG; ++m) { S.sstr[o][m][0] = S.sstr[o][m][1] = S.sstr[o][m][2] = KF_NONE; }
        if (o == 0) {
            for (int m = 0; m <= NG + 1; ++m) S.scw[m] = 0.0;
            for (int e = NG * NG; e < NG * NG + 16; ++e) sbin[e] = 0;
        }
        for (int q = 0; q < NG; ++q) {
            for (int r = 0; r < NG; ++r) {
                const int i = o == 0 ? q : r, j = o == 0 ? r : q, e = i * NG + j;
                if (pending >= 0) { S.sstr[o][pending][2] = (unsigned short)e; pending = -1; }
                const double w = S.wtabd[e];
                const double gdn = __dadd_rn(run, w);
                if (o == 0) sbin[e] = (unsigned char)(ig < NG ? ig : NG - 1);
                if (ig < NG) {
                    if (gdn < S.gord[ig + 1]) {
                        RA[ig * NG + i] = __dadd_rn(RA[ig * NG + i], w);
                        RB[ig * NG + j] = __dadd_rn(RB[ig * NG + j], w);
                    } else {
                        const double frac = __ddiv_rn(__dsub_rn(S.gord[ig + 1], run), __dsub_rn(gdn, run));
                        const double f = __dmul_rn(frac, w);
                        RA[ig * NG + i] = __dadd_rn(RA[ig * NG + i], f);
                        RB[ig * NG + j] = __dadd_rn(RB[ig * NG + j], f);
                        ++ig;
                        S.sstr[o][ig][0] = (unsigned short)prev;
                        S.sstr[o][ig][1] = (unsigned short)e;
                        pending = ig;
                        if (o == 0)      // (the formula of the sorted folds, step 3 of the kernel)
                            S.scw[ig] = __dsub_rn(__dadd_rn(run, w), S.gord[ig]);
                        if (ig < NG) {
                            const double f2 = __dmul_rn(__dsub_rn(1.0, frac), w);
                            RA[ig * NG + i] = __dadd_rn(RA[ig * NG + i], f2);
                            RB[ig * NG + j] = __dadd_rn(RB[ig * NG + j], f2);
                            if (gdn >= S.gord[ig + 1]) ok = false;      // one element over two edges
                        }
                    }
                }
                prev = e;
                run = gdn;
            }
        }
        if (ig < NG - 1) ok = false;     // some bin never closed
        if (o == 0)
            for (int i = 0; i <= NG; ++i) {
                int c = 0;
                for (int m = 1; m <= NG; ++m) c += S.sstr[0][m][1] != KF_NONE && S.sstr[0][m][1] < i * NG;
                S.nedge[i] = (unsigned char)c;
            }
        if (!ok) atomicAnd(&S.ok_static, 0);
    } else if (o == 32) {
        // in any order: no element may lie over two edges (the host plan checks the same and asks for the
        // sequential scan otherwise)
        double wmax = 0.0, wmin = INFINITY, dmin = INFINITY, total = 0.0;
        for (int e = 0; e < NG * NG; ++e) {
            wmax = fmax(wmax, S.wtabd[e]);
            wmin = fmin(wmin, S.wtabd[e]);
            total = __dadd_rn(total, S.wtabd[e]);
        }
        for (int m = 0; m < NG; ++m) dmin = fmin(dmin, __dsub_rn(S.gord[m + 1], S.gord[m]));
        if (!(wmax < dmin)) atomicAnd(&S.ok_static, 0);
        // and no element may START beyond the last edge (bins are numbered 0 .. NG-1 in the marginal matrices)
        if (!(__dsub_rn(total, wmin) < S.gord[NG])) atomicAnd(&S.ok_static, 0);
    }
}

// The product on the FP64 tensor cores: D[bin][column] += sum_t M[bin][t] * rows[t][column] as mma.m8n8k4 tiles --
// 3 tiles of 8 bins (the last one half empty), XS/8 tiles of 8 columns, NG/4 steps of 4 rows.
// Fragments (PTX ISA, mma.m8n8k4 .f64): a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], d = D[lane>>2][2*(lane&3) + {0,1}].
// With the marginals stored [bin][t] at stride NG = 20 the sixteen 8-byte words of a half-warp's A load fall in
// sixteen different bank pairs.  (The normalisation sum_t M[bin][t] is the weight of the bin: a constant, KfShared::rwid.)
// The second operand is read through one pointer and one stride per lane and column tile (NULL = the column is zero):
// for the row terms that is X[t][column]; the column terms {b_t, bT_t, 0 .., k_t in the column of the gas being
// folded, 0 ..} are read where they lie (kf_yptr), no matrix of them is built.
template <int NG, int XS>
__device__ __forceinline__ void kf_mma(const double *__restrict__ M, const double *const (&bp)[XS / 8],   // @phase mma
                                       const int (&bs)[XS / 8], int lane, double (&d)[3][XS / 8][2])
{
    static_assert(NG % 4 == 0 && NG <= 24, "three 8-bin tiles, whole k-steps");
    const int kk = lane & 3, mm = lane >> 2;
#pragma unroll 1
    for (int k0 = 0; k0 < NG; k0 += 4) {
        double b[XS / 8];
#pragma unroll
        for (int n = 0; n < XS / 8; ++n) b[n] = bp[n] ? bp[n][(k0 + kk) * bs[n]] : 0.0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const double a = M[(8 * t + mm) * NG + k0 + kk];     // (bins >= NG: whatever follows; those rows are dropped)
#pragma unroll
            for (int n = 0; n < XS / 8; ++n)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d[t][n][0]), "+d"(d[t][n][1]) : "d"(a), "d"(b[n]));
        }
    }
}

// Lane-per-row (SA = NG, SB = 1) or lane-per-column (SA = 1, SB = NG) walk over the bin bytes: the weights of a run of
// equal bins are summed in a register and stored to M[bin*NG + lane] (zero-filled before; the set-up has checked that
// no element can start beyond the last edge, so bin < NG).  The run store is two predicated instructions (PTX: ptxas would branch).  Along a row
// or a column of the key matrix the bins cannot decrease: the operands are checked to be non-decreasing, rounding and
// truncation are monotone, equal key bits are ordered by the element number (which grows along rows and columns),
// the pair fix-up of step 3 only touches elements of different rows and columns, and the static bins of a tail follow
// the head's.  KF_CHECK_WALK = 1 asserts it at run time (the cell is handed over if it fails): 2.3 % of the kernel.
template <int NG, int SA, int SB>
__device__ __forceinline__ bool kf_walk(const unsigned char *__restrict__ bin, const double *__restrict__ wtabd,   // @phase walk
                                        double *__restrict__ M, int lane)
{
    int dec = 0;
    if (lane < NG) {
        const unsigned char *bp = bin + lane * SA;
        const double *wp = wtabd + lane;                    // (the weight table is symmetric)
        const unsigned mbase = (unsigned)__cvta_generic_to_shared(M + lane);
        unsigned ma = mbase + (unsigned)bp[0] * (NG * 8);   // shared-memory address of the current run's slot
        double acc = 0.0;
#pragma unroll 2
        for (int t = 0; t < NG; ++t) {
            const unsigned na = mbase + (unsigned)bp[t * SB] * (NG * 8);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, %2;\n\t@p st.shared.f64 [%2], %0;\n\t"
                         "@p mov.f64 %0, 0d0000000000000000;\n\t}" : "+d"(acc) : "r"(na), "r"(ma) : "memory");
            if (KF_CHECK_WALK) dec |= (int)(na - ma);
            ma = na;
            acc = __dadd_rn(acc, wp[t * NG]);
        }
        asm volatile("st.shared.f64 [%0], %1;" :: "r"(ma), "d"(acc) : "memory");
    }
    return dec < 0;
}

// move the (1-frac) part of every straddler from the bin it starts in to the next one; lane = edge, line = the
// straddler's row (or column)
template <int NG>
__device__ __forceinline__ void kf_correct(double *__restrict__ M, int lane, int line, double cw)   // @phase correct
{
    const bool act = line >= 0;
    if (act) M[(lane - 1) * NG + line] = __dsub_rn(M[(lane - 1) * NG + line], cw);
    __syncwarp();
    if (act && lane < NG) M[lane * NG + line] = __dadd_rn(M[lane * NG + line], cw);
    __syncwarp();
}

template <int NG, int XS, int KG, bool GRAD, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1)
ans_koverlap_fast_kernel(OvParams P, int *__restrict__ fb_count, int *__restrict__ fb_list, int *__restrict__ fb_why,
                         int *__restrict__ work)   // @phase cta_setup
{
    using L = KfWarpLayout<NG, XS, KG, GRAD>;
    constexpr int NN = NG * NG;
    constexpr int EPL = 16;
    static_assert(NN <= 512 && NN > 256 && NN % 16 == 0, "16 keys per lane, whole lanes");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KfShared<NG> &S = *reinterpret_cast<KfShared<NG> *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NGAS = P.NGAS, NLAY = P.NLAY, NP1 = NGAS + 1;
    double *wb = reinterpret_cast<double *>(smem_raw + ((sizeof(KfShared<NG>) + 15) & ~(size_t)15)) + (size_t)warp * L::TOTAL;
    double *RC = wb + L::RC, *X = wb + L::X, *bT = wb + L::BT, *av = wb + L::AV, *bv = wb + L::BV;
    double *kbuf = wb + L::KBUF, *dbuf = wb + L::DBUF, *gbs = wb + L::GBS;
    int *spos = reinterpret_cast<int *>(wb + L::SPOS);
    unsigned char *bin = reinterpret_cast<unsigned char *>(wb + L::BIN);
    unsigned *vbuf = reinterpret_cast<unsigned *>(RC);

    // ---- CTA set-up ----------------------------------------------------------------------------
    if (threadIdx.x == 0) { S.ok_f32 = 1; S.ok_static = 1; }
    for (int i = threadIdx.x; i <= NG; i += blockDim.x) S.gord[i] = P.g_ord[i];
    if (threadIdx.x == 0) S.gord[NG + 1] = INFINITY;
    for (int e = threadIdx.x; e < 2 * 2 * NN; e += blockDim.x) (&S.stat[0][0][0])[e] = 0.0;
    __syncthreads();
    {
        bool ok = true;
        for (int e = threadIdx.x; e < NN + 16; e += blockDim.x) {
            float wf = 0.0f;
            if (e < NN) {
                const int i = e / NG, j = e - i * NG;
                const double di = P.del_g[i], dj = P.del_g[j];
                wf = __fmul_rn((float)di, (float)dj);
                ok = ok && (double)(float)di == di && (double)wf == P.weight[e] && wf > 0.0f;
                S.wtabd[e] = (double)wf;
            }
            S.wtabf[e] = wf;
        }
        if (!ok) atomicAnd(&S.ok_f32, 0);
    }
    __syncthreads();
    kf_static_setup<NG>(S);
    __syncthreads();
    if (threadIdx.x < NG) {
        double w = 0.0;
        for (int i = 0; i < NG; ++i) w = __dadd_rn(w, S.stat[0][0][threadIdx.x * NG + i]);
        S.rwid[threadIdx.x] = __ddiv_rn(1.0, w);
    }
    __syncthreads();
    const long long ncell = (long long)P.NWAVE * NLAY;
    if (!S.ok_f32 || !S.ok_static) {
        // weights that are not float32 products, or a quadrature in which one element can lie over two bin
        // edges: everything goes to the general kernel
        if (blockIdx.x == 0 && threadIdx.x == 0) *fb_count = -1;
        return;
    }

    // Cells cost between one and five sorted folds: the first round is dealt out statically, after that a warp takes
    // the next cell from a counter when it is done (the ticket is drawn while the current cell is being worked on).
    long long cell = (long long)blockIdx.x * NWARPS + warp;
    while (cell < ncell) {   // @phase cell_prologue
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(work, 1);
        const int iw = (int)(cell / NLAY);
        const int l = (int)(cell - (long long)iw * NLAY);
        bool fallback = false;
        int why = 0;

        // k (and dk/dT) of the cell
        if (P.fused) {
            const size_t plane = (size_t)P.NWAVE * NG * NGAS;
            const size_t toff = ((size_t)__ldg(P.plan.ip_lo + l) * P.NT + __ldg(P.plan.it_lo + l)) * plane +
                                (size_t)iw * NG * NGAS;
            const double *w = P.plan.w4 + 4 * l;
            const double w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
            const double omv = GRAD ? __ldg(P.plan.omv + l) : 0.0, vv = GRAD ? __ldg(P.plan.vv + l) : 0.0,
                         dudt = GRAD ? __ldg(P.plan.dudt + l) : 0.0;
            for (int e = lane; e < NG * NGAS; e += 32) {
                double kv, dv = 0.0;
                ans_kinterp_elem<GRAD>(P.tab, toff + e, P.NT, plane, w0, w1, w2, w3, omv, vv, dudt, kv, dv);
                kbuf[e] = kv;
                if (GRAD) dbuf[e] = dv;
            }
        } else {
            for (int e = lane; e < NG * NGAS; e += 32) {
                const int g = e / NGAS, gas = e - g * NGAS;
                const size_t o = (((size_t)iw * NG + g) * NLAY + l) * NGAS + gas;
                kbuf[e] = __ldg(P.k + o);
                if (GRAD) dbuf[e] = __ldg(P.dkdT + o);
            }
        }
        for (int i = lane; i < NG * XS; i += 32) X[i] = 0.0;
        for (int i = lane; i < NG; i += 32) av[i] = 0.0;
        __syncwarp();

#define KB(g, gas) kbuf[(g) * NGAS + (gas)]
#define DB(g, gas) dbuf[(g) * NGAS + (gas)]
        for (int igas = 0; igas < NGAS - 1 && !fallback; ++igas) {   // @phase fold_setup
            const int g1 = igas + 1;
            const double am1 = __ldg(P.amount + (size_t)g1 * NLAY + l);
            const bool next_neg = __all_sync(FULL, __dmul_rn(KB(NG - 1, g1), am1) <= 0.0);
            bool do_fold = false;
            // short cuts of the reference (ForwardModel_0.py:5897-5909, :5929-5938)
            if (igas == 0) {
                const double am0 = __ldg(P.amount + l);
                const bool first_neg = __all_sync(FULL, __dmul_rn(KB(NG - 1, 0), am0) <= 0.0);
                if (lane < NG) {
                    const int i = lane;
                    if (first_neg) {
                        av[i] = X[i * XS] = __dmul_rn(KB(i, 1), am1);
                        if (GRAD) { X[i * XS + 3] = KB(i, 1); X[i * XS + 1] = __dmul_rn(DB(i, 1), am1); }
                    } else {
                        av[i] = X[i * XS] = __dmul_rn(KB(i, 0), am0);
                        if (GRAD) { X[i * XS + 2] = KB(i, 0); X[i * XS + 1] = __dmul_rn(DB(i, 0), am0); }
                    }
                }
                do_fold = !first_neg && !next_neg;
            } else {
                if (next_neg) {
                    if (GRAD && lane < NG) X[lane * XS + 2 + g1] = __dmul_rn(X[lane * XS + 1], 0.0);
                } else if (__all_sync(FULL, av[NG - 1] <= 0.0)) {
                    __syncwarp();
                    if (lane < NG) {
                        const int i = lane;
                        av[i] = X[i * XS] = __dmul_rn(KB(i, g1), am1);
                        if (GRAD) { X[i * XS + 2 + g1] = KB(i, g1); X[i * XS + 1] = __dmul_rn(DB(i, g1), am1); }
                    }
                } else {
                    do_fold = true;
                }
            }
            if (do_fold && lane < NG) {
                const double b = __dmul_rn(KB(lane, g1), am1);
                bv[lane] = b;
                if (GRAD) bT[lane] = __dmul_rn(DB(lane, g1), am1);
            }
            __syncwarp();
            if (!do_fold) continue;

            // ---- preconditions and data-independent orders -----------------------------------------   // @phase precond
            const int l0 = lane < NG ? lane : NG - 1, l1 = lane + 1 < NG ? lane + 1 : NG - 1;
            const double a_l = av[l0], a_n = av[l1], b_l = bv[l0], b_n = bv[l1];
            const double a_first = av[0], a_last = av[NG - 1], b_first = bv[0], b_last = bv[NG - 1];
            const double kmax = __dadd_rn(a_last, b_last);
            const int ek = (__double2hiint(kmax) >> 20) & 0x7ff;
            // (lanes >= NG-1 compare an element with itself: false only for NaN)
            const bool mono = (a_l <= a_n) & (b_l <= b_n) & (kmax > 0.0) & (ek > 100) & (ek < 2000);
            if (!__all_sync(FULL, mono)) { fallback = true; why = 1; break; }
            const bool inner = lane < NG - 1;
            const bool rowok = __all_sync(FULL, !inner || __dadd_rn(a_l, b_last) <= __dadd_rn(a_n, b_first));
            const bool colok = !rowok && __all_sync(FULL, !inner || __dadd_rn(a_last, b_l) < __dadd_rn(a_first, b_n));
            int ord = rowok ? 0 : (colok ? 1 : -1);
            // Partly static orders: if no row from some row on interleaves with its neighbours, those rows follow the
            // rest in row-major order -- only the `hl` rows before them (the head) need sorting, and the bins, straddlers
            // and straddler fractions of the tail are those of the row-major order (data-independent).
            int hl = NG;
            if (KF_PARTIAL && ord < 0) {
                // (strictly separated here: no key of the head may equal one of the tail)
                const unsigned rss = __ballot_sync(FULL, !inner || __dadd_rn(a_l, b_last) < __dadd_rn(a_n, b_first));
                hl = 33 - __clz(~rss);      // last interleaving boundary + 2
            }
            if (ord >= 0) {
                // the static straddlers must be strictly separated from their neighbours in the order
                bool sep = true;
                if (lane >= 1 && lane <= NG) {
                    const int pe = S.sstr[ord][lane][0], se = S.sstr[ord][lane][1], ne = S.sstr[ord][lane][2];
                    if (se != KF_NONE) {
                        const double ks = __dadd_rn(av[se / NG], bv[se % NG]);
                        if (pe != KF_NONE) sep &= __dadd_rn(av[pe / NG], bv[pe % NG]) < ks;
                        if (ne != KF_NONE) sep &= ks < __dadd_rn(av[ne / NG], bv[ne % NG]);
                    }
                }
                if (!__all_sync(FULL, sep)) { ord = -1; if (fb_why && lane == 0) atomicAdd(fb_why + 7, 1); }
            }
            if (fb_why && lane == 0) atomicAdd(fb_why + (ord >= 0 ? 5 : 6), 1);

            constexpr int NT8 = XS / 8;   // @phase static_mma
            double dfr[3][NT8][2];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int n = 0; n < NT8; ++n) dfr[t][n][0] = dfr[t][n][1] = 0.0;
            }

            // second operands of the two products, per lane (fragment row kk = lane & 3, column 8n + (lane >> 2))
            const double *xp[NT8], *yp[NT8];
            int xs[NT8], ys[NT8];
#pragma unroll
            for (int n = 0; n < NT8; ++n) {
                const int col = 8 * n + (lane >> 2);
                xp[n] = X + col;
                xs[n] = XS;
                yp[n] = col == 0 ? bv : ((GRAD && col == 1) ? bT : ((GRAD && col == 2 + g1) ? kbuf + g1 : nullptr));
                ys[n] = (GRAD && col == 2 + g1) ? NGAS : 1;
            }

            int se = -1;            // element number of the straddler of edge `lane`
            double cw = 0.0;        // (1-frac) * its weight: goes to the next bin
            bool wbad = false;
            if (ord < 0) {
                const int nhead = hl * NG;
                const int mh = S.nedge[hl];        // edges 1 .. mh fall into the head
                // ---- 1. packed keys, sorted in registers --------------------------------------------   // @phase keys
                unsigned v[EPL];
                {
                    // key bits = the leading 23 bits of (high word of the key - high word of the base), the base being
                    // the binade of the smallest key (at most 62 binades below the largest): the fewer binades the
                    // keys span, the more mantissa bits take part (one binade of head room keeps them below the padding)
                    const int basehi = max(max(__double2hiint(__dadd_rn(a_first, b_first)), 0) & 0x7ff00000, (ek - 62) << 20);
                    const int sh = min(__clz(__double2hiint(kmax) - basehi + (1 << 20)), 11);
                    const int ebase = lane * EPL;
                    const int i0 = ebase / NG, j0 = ebase - i0 * NG;
                    const double a0 = av[i0 < NG ? i0 : NG - 1], a1 = av[i0 + 1 < NG ? i0 + 1 : NG - 1];
                    // slots 0 .. n0-1: the tail of row i0, ascending in j; slots n0 .. 15: the head of row i0+1,
                    // DESCENDING in j -- a bitonic sequence (b is ascending); rows of the tail are padding
                    const int n0 = NG - j0 < EPL ? NG - j0 : EPL;
                    // (j0 and n0 are multiples of 4: a group of four slots lies on one side of n0)
                    static_assert(NG % 4 == 0 && EPL % 4 == 0, "groups of four slots");
#pragma unroll
                    for (int g = 0; g < EPL / 4; ++g) {
                        const bool second = g > 0 && 4 * g >= n0;
                        const double ag = second ? a1 : a0;
                        const int sg = second ? -1 : 1;                               // slot r holds column jg + sg * r,
                        const double *bg = bv + (second ? EPL - 1 : j0);
                        const int eg = second ? ebase + n0 + (EPL - 1) : ebase;       // element eg + sg * r
                        const bool real = (second ? i0 + 1 : i0) < hl;
#pragma unroll
                        for (int r = 4 * g; r < 4 * g + 4; ++r) {
                            const double key = __dadd_rn(ag, bg[sg * r]);
                            const int t = max(__double2hiint(key) - basehi, 0);
                            v[r] = real ? (((unsigned)t << sh) & 0xfffffe00u) | (unsigned)(eg + sg * r) : (0xfffffe00u | (unsigned)NN);
                        }
                    }
                }
                // the head fills lanes 0 .. ceil(nhead / 16) - 1; the padding above it needs no merging
                kf_sort(v, lane, KF_SKIP ? 2 << (31 - __clz(((nhead + EPL - 1) >> 4) - 1)) : 32);   // @phase sort_call
                if (KF_PARTIAL && nhead < NN) {
                    double *bd = reinterpret_cast<double *>(bin);
                    for (int t = lane; t < (NN + 16) / 8; t += 32) bd[t] = S.sbin[t];
                }
                // sorted words for the straddler checks (16-byte units, lane-interleaved: conflict-free)
                {
                    uint4 *vb4 = reinterpret_cast<uint4 *>(vbuf);
#pragma unroll
                    for (int q = 0; q < EPL / 4; ++q) vb4[q * 32 + lane] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
                if (lane <= NG + 1) spos[lane] = KF_NONE;
                // ---- 2. cumulative weights, bin of every element, straddlers ------------------------   // @phase pass1_scan
                float wf[EPL];
                double local = 0.0;
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    wf[r] = S.wtabf[v[r] & 511u];
                    local = __dadd_rn(local, (double)wf[r]);
                }
                double incl = local;
#pragma unroll 1
                for (int d = 1; d < 32; d <<= 1) {
                    const double up = shfl_up_d(incl, d);
                    if (lane >= d) incl = __dadd_rn(incl, up);
                }
                const double base = __dsub_rn(incl, local);      // exact
                int m;
                {   // number of edges g_ord[1..NG] <= base
                    int lo = 0, hi = NG;
#pragma unroll 1
                    for (int it = 0; it < 5; ++it) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (S.gord[mid] <= base) lo = mid; else hi = mid - 1;
                    }
                    m = lo;
                }
                __syncwarp();
                {
                    double rel = 0.0;   // @phase pass2
                    double edge_rel = __dsub_rn(S.gord[m + 1], base);
#pragma unroll
                    for (int r = 0; r < EPL; ++r) {
                        const double relb = rel;
                        rel = __dadd_rn(rel, (double)wf[r]);
                        bin[v[r] & 511u] = (unsigned char)m;
                        if (rel >= edge_rel) {       // straddler of edge m+1 (ForwardModel_0.py:6009-6024)
                            ++m;
                            gbs[m] = __dadd_rn(base, relb);
                            spos[m] = lane * EPL + r;
                            edge_rel = __dsub_rn(S.gord[m + 1], base);
                        }
                    }
                }
                __syncwarp();
                // ---- 3. the straddler of edge `lane` ---------------------------------------------------   // @phase resolve
                bool bad = false;
                if (KF_PARTIAL && lane > mh && lane <= NG) {
                    // an edge of the tail: the straddler of the row-major order, if it is strictly separated from its
                    // neighbours there (as for the fully static orders)
                    const int pe = S.sstr[0][lane][0], s2 = S.sstr[0][lane][1], ne = S.sstr[0][lane][2];
                    if (spos[lane] != KF_NONE) {
                        bad = true;            // found in the head as well: the two counts disagree
                        why = 2;
                    } else if (s2 != KF_NONE) {
                        const double ks = __dadd_rn(av[s2 / NG], bv[s2 % NG]);
                        bool sep = true;
                        if (pe != KF_NONE) sep &= __dadd_rn(av[pe / NG], bv[pe % NG]) < ks;
                        if (ne != KF_NONE) sep &= ks < __dadd_rn(av[ne / NG], bv[ne % NG]);
                        bad = !sep;
                        if (bad) why = 4;
                        se = s2;
                        cw = S.scw[lane];
                    } else {
                        bad = lane < NG;       // only the last edge may stay open (:6025-6027)
                        if (bad) why = 2;
                    }
                } else if (lane >= 1 && lane <= NG) {
                    const int p = spos[lane];
                    if (p == KF_NONE) {
                        bad = lane < NG || nhead < NN;       // only the last edge may stay open (:6025-6027)
                        if (bad) why = 2;
                    } else {
                        const unsigned vp = vbuf[kf_vaddr(p)];
                        const unsigned vl = p > 0 ? vbuf[kf_vaddr(p - 1)] : ~vp, vr = vbuf[kf_vaddr(p + 1)];
                        const bool al = ((vp ^ vl) >> 9) == 0u, ar = ((vp ^ vr) >> 9) == 0u;
                        se = (int)(vp & 511u);
                        double gb = gbs[lane];
                        if (al | ar) {
                            // Equal key bits next to an edge.  The packed order is (key bits, element number); it is
                            // the true one around p if every member of p's group before p has a strictly smaller
                            // exact key and every member after p a strictly larger one (then the elements before p
                            // are exactly those with a smaller key, in any order of equal keys elsewhere).
                            const double kp = __dadd_rn(av[se / NG], bv[se % NG]);
                            bool sep = true, tie = false;
                            int nl = 0, nr = 0;
                            for (int x = p - 1; x >= 0; --x) {
                                const unsigned vx = vbuf[kf_vaddr(x)];
                                if (((vx ^ vp) >> 9) != 0u) break;
                                const int ex = (int)(vx & 511u);
                                const double kx = __dadd_rn(av[ex / NG], bv[ex % NG]);
                                sep &= kx < kp;
                                tie |= kx == kp;
                                ++nl;
                            }
                            for (int x = p + 1; x < 512; ++x) {
                                const unsigned vx = vbuf[kf_vaddr(x)];
                                if (((vx ^ vp) >> 9) != 0u) break;
                                const int ex = (int)(vx & 511u);
                                const double kx = __dadd_rn(av[ex / NG], bv[ex % NG]);
                                sep &= kx > kp;
                                tie |= kx == kp;
                                ++nr;
                            }
                            if (!sep) {
                                const int pm = spos[lane - 1], pn = spos[lane + 1];
                                const int px = nl ? p - 1 : p, py = px + 1;
                                if (tie) {
                                    bad = true;  // a true tie on an edge: numba's order decides (general kernel)
                                    why = 8;
                                } else if (nl + nr != 1 || pm == px || pn == py) {
                                    bad = true;  // a misordered group of three or more, or two edges inside a pair
                                    why = 4;
                                } else {
                                    // a lone pair in the wrong order: x = first, y = second in the packed order, ky < kx
                                    const unsigned vx = nl ? vl : vp, vy = nl ? vp : vr;
                                    const int ex = (int)(vx & 511u), ey = (int)(vy & 511u);
                                    const double wx = S.wtabd[ex], wy = S.wtabd[ey];
                                    const double q = nl ? __dsub_rn(gb, wx) : gb;     // cumulative weight before the pair
                                    const double qy = __dadd_rn(q, wy);
                                    if (qy >= S.gord[lane]) {          // y straddles, x follows it whole
                                        se = ey; gb = q;
                                        bin[ey] = (unsigned char)(lane - 1);
                                        bin[ex] = (unsigned char)lane;
                                    } else {                           // y lies before the edge, x straddles
                                        se = ex; gb = qy;
                                        bin[ey] = (unsigned char)(lane - 1);
                                        bin[ex] = (unsigned char)(lane - 1);
                                    }
                                }
                            }
                        }
                        const double w = S.wtabd[se];
                        // (1 - frac) * w with frac = (g_ord - gb) / w (:6009-6024) is what lies beyond the edge: no division
                        cw = __dsub_rn(__dadd_rn(gb, w), S.gord[lane]);
                    }
                }
                if (__any_sync(FULL, bad)) { fallback = true; break; }
                __syncwarp();       // (the sorted words alias the marginals)
                // ---- 4. marginals and products ----------------------------------------------------------   // @phase marginals
                {
                    double2 *z = reinterpret_cast<double2 *>(RC);
                    for (int t = lane; t < NN / 2; t += 32) z[t] = make_double2(0.0, 0.0);
                }
                __syncwarp();
                wbad = kf_walk<NG, NG, 1>(bin, S.wtabd, RC, lane);
                __syncwarp();
                kf_correct<NG>(RC, lane, se >= 0 ? se / NG : -1, cw);
            }
            // (one copy of each product for both kinds of fold: the hot code must stay inside the instruction cache)
            kf_mma<NG, XS>(ord >= 0 ? S.stat[ord][0] : RC, xp, xs, lane, dfr);
            if (ord < 0) {
                __syncwarp();
                {
                    double2 *z = reinterpret_cast<double2 *>(RC);
                    for (int t = lane; t < NN / 2; t += 32) z[t] = make_double2(0.0, 0.0);
                }
                __syncwarp();
                wbad |= kf_walk<NG, 1, NG>(bin, S.wtabd, RC, lane);
                __syncwarp();
                kf_correct<NG>(RC, lane, se >= 0 ? se % NG : -1, cw);
            }
            kf_mma<NG, XS>(ord >= 0 ? S.stat[ord][1] : RC, yp, ys, lane, dfr);
            if (__any_sync(FULL, wbad)) { fallback = true; why = 16; break; }
            __syncwarp();
            // ---- bin m: normalise (ForwardModel_0.py:6016-6017, :6026-6027) and store -------------------   // @phase normalise
            {
                // the sum of the marginals of a bin is the weight of the bin, whatever the order of the keys: 1 / weight
                // is a table of the CTA set-up
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int mrow = 8 * t + (lane >> 2);
                    if (mrow < NG) {
                        const double rs = S.rwid[mrow];
#pragma unroll
                        for (int n = 0; n < NT8; ++n) {
                            const double x0 = __dmul_rn(dfr[t][n][0], rs), x1 = __dmul_rn(dfr[t][n][1], rs);
                            *reinterpret_cast<double2 *>(X + mrow * XS + 8 * n + 2 * (lane & 3)) = make_double2(x0, x1);
                            if (n == 0 && (lane & 3) == 0) av[mrow] = x0;
                        }
                    }
                }
            }
            __syncwarp();
        }
#undef KB
#undef DB
        if (fallback) {   // @phase cell_epilogue
            if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = (int)cell;
            if (fb_why) {
                const unsigned wm = __reduce_or_sync(FULL, (unsigned)why);
                if (lane == 0) for (int b = 0; b < 5; ++b) if (wm >> b & 1u) atomicAdd(fb_why + b, 1);
            }
        } else {
            for (int g = lane; g < NG; g += 32) {
                const size_t o = ((size_t)iw * NG + g) * NLAY + l;
                P.tau[o] = X[g * XS];
                if (GRAD) {
                    for (int p = 0; p < NGAS; ++p) P.dk[o * NP1 + p] = X[g * XS + 2 + p];
                    P.dk[o * NP1 + NGAS] = X[g * XS + 1];
                }
            }
        }
        __syncwarp();
        cell = (long long)gridDim.x * NWARPS + __shfl_sync(FULL, ticket, 0);
    }
}

// warps per CTA (one CTA per SM): as many as 227 KB of shared memory hold, at most 28 (72 registers per thread)
template <int NG, int XS, int KG, bool GRAD>
constexpr int kf_warps()
{
    constexpr size_t cta = (sizeof(KfShared<NG>) + 15) & ~(size_t)15;
    constexpr size_t per_warp = (size_t)KfWarpLayout<NG, XS, KG, GRAD>::TOTAL * 8;
    constexpr size_t fit = (232448 - cta) / per_warp;
    return fit > 28 ? 28 : (int)fit;
}

template <int NG, int XS, int KG, bool GRAD>
int kf_launch(const OvParams &P, int *scratch, int *why, int *work, cudaStream_t stream)
{
    constexpr int NWARPS = kf_warps<NG, XS, KG, GRAD>();
    static_assert(NWARPS >= 8, "too little shared memory per warp");
    using L = KfWarpLayout<NG, XS, KG, GRAD>;
    const size_t smem = ((sizeof(KfShared<NG>) + 15) & ~(size_t)15) + (size_t)NWARPS * L::TOTAL * 8;
    auto kern = ans_koverlap_fast_kernel<NG, XS, KG, GRAD, NWARPS>;
    ANS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, nsm = 148;
    ANS_CUDA_CHECK(cudaGetDevice(&dev));
    ANS_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const long long ncell = (long long)P.NWAVE * P.NLAY;
    long long grid = ans_div_up(ncell, NWARPS);
    if (grid > nsm) grid = nsm;
    kern<<<(unsigned)grid, NWARPS * 32, smem, stream>>>(P, scratch, scratch + 1, why, work);
    ANS_LAUNCH_CHECK();
    return ANSB200_OK;
}

}   // namespace

// Is there a fast kernel for this shape?  (the general kernel covers everything else)
bool ov_fast_supported(const OvParams &P, bool grad)
{
    (void)grad;
    return P.NG == 20 && P.NGAS >= 2 && P.NGAS <= 14 && !P.seq_rebin && P.del_g != nullptr && P.weight && P.g_ord;
}

// scratch: [0] = number of cells left for the general kernel (-1: all of them), [1..] = their numbers; work: a zeroed
// counter (dynamic cell assignment)
int ov_fast_launch(const OvParams &P, bool grad, int *scratch, int *why, int *work, cudaStream_t stream)
{
    if (!grad) return P.NGAS <= 6 ? kf_launch<20, 8, 6, false>(P, scratch, why, work, stream)
                                  : kf_launch<20, 8, 14, false>(P, scratch, why, work, stream);
    if (P.NGAS <= 6) return kf_launch<20, 8, 6, true>(P, scratch, why, work, stream);
    return kf_launch<20, 16, 14, true>(P, scratch, why, work, stream);
}
