// gik_rtr_cta.cu -- trust-region solve for LARGE, DENSE graphs (32 < N <= 128): one CTA per problem,
// two CTAs per SM.
//
// Same algorithm and arithmetic conventions as the warp kernels (reference trust_region.py:112-599,
// costs.py:79-207, fixed_rank_psd_sym.py:91-137).  With spherical obstacles the reference's graph has
// N = 118 nodes of which 106 are mutually fixed anchors, i.e. 5609 of 6903 node pairs carry an equality
// term (BASELINE configs[2]) and a solve needs ~100 k tCG iterations of 13.8 k directed pair evaluations.
// The pair pass of an inner iteration is FP64-issue bound (15 FP64 instructions per directed pair); the
// rest of the iteration (two CTA-wide reductions, the projection, alpha / beta, ~300 dependent
// instructions) is a latency chain during which the FP64 pipe idles.  Layout:
//
//   * per accepted iterate x the pair quantity c2_ij = 2 act (d_ij - T_ij) is cached in shared memory, the
//     hinge activity as bit masks in registers; both are rebuilt by the cost/gradient pass of the
//     proposal from the static targets in global memory (L2) -- the 2 * n_anchor goal-dependent targets
//     all involve p_n or q_n and live in two per-problem rows Tp, Tq;
//   * c2 is symmetric and stored PACKED: 32 x 32 blocks (a <= b) with row stride 33, element (p in a, q in b)
//     at [q][p].  A warp reading "fixed j, lanes over i" hits a row of the block when block(i) <= block(j)
//     and a column otherwise; the odd stride makes both conflict-free.  74 KB instead of 121 KB for
//     N = 118, so that TWO CTAs (problems) fit on an SM and one problem's latency chain overlaps the
//     other's pair pass (ncu, one CTA of 256 threads per SM: FP64 pipe 51 % active, 40 % of the time in the
//     latency chain);
//   * REGISTER TILE: warp w owns the neighbour slice j in [w JS, (w+1) JS), lane l the four nodes
//     i = l, l+32, l+64, l+96.  Per neighbour j a warp issues 6 shared-memory broadcasts (2 x_j, delta_j)
//     and 4 reads of c2 for 4 x 32 pairs = 60 FP64 instructions per lane in 12 independent accumulator
//     chains;
//   * coordinates are exchanged DOUBLED (2 x): D' = 2 (x_i - x_j) gives 4 <D,w> D = <D',w> D' and
//     2 r D = r D' with exact power-of-two scalings, i.e. bit-identical terms with one multiply less;
//   * the slices' partial node sums go through shared memory and are added in slice order by the node's
//     owner thread (thread t owns node t and its solver state x, g, eta, Heta, r, delta in registers);
//     inner products are warp butterflies + a 4-entry shared-memory stage read by every thread in the
//     same order -> all scalars are block-uniform, every branch of tCG / RTR is taken by the whole CTA,
//     and an inner iteration has 4 CTA barriers.
//
// A pair carrying more than one term kind is not representable densely; gik_launch_rtr_cta then returns 1
// and the caller falls back to k_rtr.
#include <type_traits>

#include "gik_rtr.cuh"

namespace {

constexpr int NPAD = 128;
constexpr int kThreads = NPAD;              // thread t owns node t
constexpr int kWarps = kThreads / 32;       // neighbour slices
constexpr int kOwnWarps = kWarps;
constexpr int kStride = 33;                 // row stride of a packed 32 x 32 block of c2
constexpr int kRedA = 18, kRedB = 18;
constexpr int kMom = 32;                   // 15 moments of the tCG direction, 9 of the cached point (from 16)
static_assert(NPAD / 2 <= 64, "one 64-bit activity mask per tile node: a neighbour slice never exceeds 64 nodes");

// doubles of the packed symmetric pair cache: column block b holds b + 1 blocks of rows_b rows
__host__ __device__ inline int c2_rows(int N, int b) { return min(32, N - 32 * b); }
__host__ __device__ inline int c2_base(int N, int a, int b)   // block (a <= b)
{
    return (16 * b * (b + 1) + a * c2_rows(N, b)) * kStride;
}
__host__ __device__ inline int c2_doubles(int N)
{
    const int nb = (N + 31) / 32;
    return c2_base(N, nb, nb - 1);   // = end of the last column block
}

struct CtaTables {
    const double *target;       // [N][N] squared targets (static part)
    const unsigned char *kind;  // [N][N] GIK_TERM_* or 3 = no term
    const int32_t *goal_i, *goal_j, *goal_slot;
    int n_goal_edges;
    int gp, gq;                 // node indices of p_n, q_n (-1: no goal-dependent targets)
    int hub;                    // node whose pairs carry a second term (-1: none)
    const unsigned char *hub_kind;  // [N] kind of the second term of the pair (hub, i), 3 = none
    const double *hub_target;       // [N]
    // all of the above are indexed by POSITION; perm[position] = node.  Positions >= jF form the equality clique,
    // the 32-node blocks m >= mF lie inside it (mF = NB: no fast path)
    const int32_t *perm;
    int jF, mF;
};

// Contributions of one clique node at the doubled point X to the 9 constants of the point: m1 = sum X, M2 = sum X X^T
// (xx, xy, xz, yy, yz, zz).  Explicit roundings: the same values must come out of every call site (a parked problem
// recomputes what a running one carries).
__device__ __forceinline__ void clique_consts(const double X[3], bool in, double *m)
{
    m[0] = in ? X[0] : 0.0; m[1] = in ? X[1] : 0.0; m[2] = in ? X[2] : 0.0;
    m[3] = in ? __dmul_rn(X[0], X[0]) : 0.0; m[4] = in ? __dmul_rn(X[0], X[1]) : 0.0; m[5] = in ? __dmul_rn(X[0], X[2]) : 0.0;
    m[6] = in ? __dmul_rn(X[1], X[1]) : 0.0; m[7] = in ? __dmul_rn(X[1], X[2]) : 0.0; m[8] = in ? __dmul_rn(X[2], X[2]) : 0.0;
}

// butterfly over a warp, lane 0 stores K partial sums
template <int K>
__device__ __forceinline__ void warp_sum_store(double *v, double *dst, bool store)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(GIK_FULL_MASK, v[k], off, 32);
    }
    if (store) {
#pragma unroll
        for (int k = 0; k < K; ++k) dst[k] = v[k];
    }
}

// every thread adds the owner warps' partial sums in the same order
template <int K, int STRIDE>
__device__ __forceinline__ void block_sum_load(double *v, const double *src)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = src[k];
#pragma unroll
        for (int w = 1; w < kOwnWarps; ++w) s += src[w * STRIDE + k];
        v[k] = s;
    }
}

template <int NB>   // 32-node blocks = nodes per lane in the pair passes
__global__ void __launch_bounds__(kThreads, 2) k_rtr_cta(const RtrArgs a, const CtaTables tb)
{
    extern __shared__ double smem[];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool first_of_own_warp = lane == 0;   // stores its warp's partial inner products
    const bool owner = tid < N;                     // thread tid owns node tid
    // this warp's neighbour slice [jlo, jhi): equal COST per warp -- a neighbour outside the equality clique costs the full
    // pair evaluation for every tile block, one inside it only for the blocks below mF (plus loads either way)
    int jlo = 0, jhi = 0;
    {
        const int NBk = (N + 31) / 32;
        const int mFk = tb.mF < NBk ? tb.mF : NBk;
        const int c_full = 15 * NBk + 8, c_fast = 15 * mFk + 3 * (NBk - mFk) + 8;
        const int total = c_full * min(tb.jF, N) + c_fast * max(0, N - tb.jF);
        int acc = 0;
        bool seen = false;
        for (int j = 0; j < N; ++j) {
            // neighbour j belongs to the warp whose share [w total / kWarps, (w + 1) total / kWarps) holds the middle of its cost
            const int cj = j < tb.jF ? c_full : c_fast;
            const int wj = min(kWarps - 1, (int)(((long long)(2 * acc + cj) * kWarps) / (2LL * total)));
            acc += cj;
            if (wj == warp) {
                if (!seen) { jlo = j; seen = true; }
                jhi = j + 1;
            }
        }
        if (__syncthreads_or(jhi - jlo > 64)) {       // the 64-bit activity masks bound a slice: fall back to equal counts
            const int JS = (N + kWarps - 1) / kWarps;
            jlo = min(N, warp * JS);
            jhi = min(N, jlo + JS);
        }
    }
    double *C2 = smem;                                  // packed symmetric 2 act (d - T) at the cached point
    double *P2 = C2 + c2_doubles(N);                    // [3][NPAD] 2 * coordinates of the cached point
    double *V = P2 + 3 * NPAD;                          // [3][NPAD] direction delta
    double *Zp = V + 3 * NPAD;                          // [kWarps][3][NPAD] partial node sums of the slices
    double *redA = Zp + kWarps * 3 * NPAD;              // [kOwnWarps][kRedA]
    double *redB = redA + kOwnWarps * kRedA;            // [kOwnWarps][kRedB]
    double *redF = redB + kOwnWarps * kRedB;            // [kWarps] cost shares of the slices
    double *Tp = redF + kWarps;                         // [NPAD] targets of the pairs (., p_n) for this problem
    double *Tq = Tp + NPAD;                             // [NPAD] targets of the pairs (., q_n)
    double *Hc2 = Tq + NPAD;                            // [NPAD] 2 act (d - T) of the second terms around the hub node
    double *RS = Hc2 + NPAD;                            // [kWarps][NPAD] slices' partial sums of c2 over the clique pairs
    double *momS = RS + kWarps * NPAD;                  // [kMom] moments (block-uniform values kept out of the registers)
    double *XW = momS + kMom;                           // [NPAD] <X_j, delta_j> of the clique nodes
    double *goal = XW + NPAD;                           // [n_goal]
    int *s_b = reinterpret_cast<int *>(goal + ((a.n_goal + 1) & ~1));                    // [0] work item, [1] park slot
    unsigned long long *s_t = reinterpret_cast<unsigned long long *>(s_b + 2);          // [0] problem start, [1] elapsed
    const int gp = tb.gp, gq = tb.gq;
    const int jF = tb.jF, mF = tb.mF;
    const bool has_fast = mF < NB;
    const int node = owner ? tb.perm[tid] : 0;          // thread tid owns POSITION tid = node perm[tid]
    const bool in_clique = owner && tid >= jF;
    const bool fast_owner = owner && tid >= 32 * mF;     // its pairs with the clique go through the moments
    const double m0 = (double)(N - jF);
    const GikSolveOpts &o = a.o;
    const double eps = 2.220446049250313e-16;

    for (int e = tid; e < c2_doubles(N); e += kThreads) C2[e] = 0.0;
    for (int e = tid; e < 3 * NPAD; e += kThreads) { P2[e] = 0.0; V[e] = 0.0; }   // padding nodes stay at 0
    for (int e = tid; e < NPAD; e += kThreads) Hc2[e] = 0.0;
    __syncthreads();

    double xt[NB][3];           // 2 * coordinates of this lane's tile nodes at the cached point
    unsigned long long amask[NB];   // bit jj: term (tile node m, neighbour jlo + jj) active there
    uint32_t hact = 0u;         // bit m: second term of the pair (tile node m, hub) active there (hub's slice only)
    const int hub = tb.hub;
    const bool hub_warp = hub >= jlo && hub < jhi;   // this warp's slice contains the hub neighbour
    // c2(i = lane + 32 m, j) in the packed cache is a row entry of block (m, jb) when m <= jb and a column
    // entry of block (jb, m) otherwise; padding lanes of a ragged last block re-read its last row
    int col_off[NB];
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        col_off[m] = 16 * m * (m + 1) * kStride + min(lane, c2_rows(N, m) - 1) * kStride;
        amask[m] = 0ull;
    }
    // offsets / strides of the NB reads for the neighbours j0, j0 + 1, ... inside block jb
    auto c2_walk = [&](int jb, int j0, int (&off)[NB], int (&step)[NB]) {
        const int jl = j0 & 31;
        const int rows_jb = c2_rows(N, jb);
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const bool row = m <= jb;
            off[m] = row ? c2_base(N, 0, jb) + (m * rows_jb + jl) * kStride + lane
                         : col_off[m] + jb * c2_rows(N, m) * kStride + jl;
            step[m] = row ? kStride : 1;
        }
    };

    auto load_tile = [&](const double *buf, double (&t)[NB][3]) {
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            t[m][0] = buf[lane + 32 * m];
            t[m][1] = buf[NPAD + lane + 32 * m];
            t[m][2] = buf[2 * NPAD + lane + 32 * m];
        }
    };
    auto store_partials = [&](const double (&z)[NB][3]) {
#pragma unroll
        for (int m = 0; m < NB; ++m) {
#pragma unroll
            for (int q = 0; q < 3; ++q) Zp[(warp * 3 + q) * NPAD + lane + 32 * m] = z[m][q];
        }
    };
    // owner: sum of the slices' partial sums of its node, in slice order
    auto gather = [&](double (&z)[3]) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double acc = Zp[q * NPAD + tid];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += Zp[(w * 3 + q) * NPAD + tid];
            z[q] = acc;
        }
    };
    auto publish = [&](double *buf, const double (&v)[3], double scale) {
        if (owner) { buf[tid] = scale * v[0]; buf[NPAD + tid] = scale * v[1]; buf[2 * NPAD + tid] = scale * v[2]; }
    };

    // direction delta -> V and, for clique nodes, <X_j, delta_j> -> XW (X = the doubled cached point; x in registers)
    auto publish_dir = [&](const double (&v)[3], const double (&xown)[3]) {
        if (owner) {
            V[tid] = v[0]; V[NPAD + tid] = v[1]; V[2 * NPAD + tid] = v[2];
            if (has_fast) XW[tid] = 2.0 * (xown[0] * v[0] + xown[1] * v[1] + xown[2] * v[2]);
        }
    };
    // The 15 moments of the published direction over the clique (positions [jF, N)): e1 = sum delta, E = sum X delta^T,
    // v3 = sum X <X, delta>.  Thread 8 k + p sums every 8th node for moment k, three butterfly steps finish it.  Runs
    // between barriers (1) and (2), next to the pair pass; nothing rides on the latency chain of the iteration.
    auto direction_moments = [&]() {
        {
            const int k = tid < 120 ? tid >> 3 : 14, part = tid & 7;   // the last eight threads repeat moment 14 and drop it
            const double *A = k < 3 ? nullptr : P2 + (k < 12 ? (k - 3) / 3 : k - 12) * NPAD;
            const double *Bv = k < 3 ? V + k * NPAD : (k < 12 ? V + ((k - 3) % 3) * NPAD : XW);
            double s0 = 0.0, s1 = 0.0;
            int jn = jF + part;
            for (; jn + 8 < N; jn += 16) {
                s0 = fma(A ? A[jn] : 1.0, Bv[jn], s0);
                s1 = fma(A ? A[jn + 8] : 1.0, Bv[jn + 8], s1);
            }
            if (jn < N) s0 = fma(A ? A[jn] : 1.0, Bv[jn], s0);
            double sm = s0 + s1;
            sm += __shfl_xor_sync(GIK_FULL_MASK, sm, 1);
            sm += __shfl_xor_sync(GIK_FULL_MASK, sm, 2);
            sm += __shfl_xor_sync(GIK_FULL_MASK, sm, 4);
            if (part == 0 && tid < 120) momS[k] = sm;
        }
    };

    // costs.py:125-169 at the point published (doubled) in P2: rebuilds xt, C2 and the activity mask, leaves the
    // slices' partial half-gradients in Zp and their cost shares in redF.  Caller synchronises before and after.
    auto pair_pass_cost = [&]() {
        load_tile(P2, xt);
        double gpart[NB][3], rsum[NB];
        double fpart = 0.0;
#pragma unroll
        for (int m = 0; m < NB; ++m) { gpart[m][0] = 0.0; gpart[m][1] = 0.0; gpart[m][2] = 0.0; amask[m] = 0ull; rsum[m] = 0.0; }
        for (int jb = jlo >> 5; jb <= (jhi - 1) >> 5; ++jb) {
            const int j0 = max(jlo, 32 * jb), j1 = min(jhi, 32 * jb + 32);
            int off[NB], step[NB];
            c2_walk(jb, j0, off, step);
            for (int j = j0; j < j1; ++j) {
                const double px = P2[j], py = P2[NPAD + j], pz = P2[2 * NPAD + j];
                const unsigned long long bit = 1ull << (j - jlo);
                const bool fj = j >= jF;
#pragma unroll
                for (int m = 0; m < NB; ++m) {
                    const int i = lane + 32 * m;
                    const unsigned kind = i < N ? tb.kind[(size_t)j * N + i] : 3u;
                    const double dx = xt[m][0] - px, dy = xt[m][1] - py, dz = xt[m][2] - pz;   // 2 (x_i - x_j)
                    const double d4 = dx * dx + dy * dy + dz * dz;                               // 4 d_ij
                    const int ic = min(i, N - 1);
                    const double tgt = j == gp ? Tp[ic] : (j == gq ? Tq[ic] :
                                       (i == gp ? Tp[j] : (i == gq ? Tq[j] : tb.target[(size_t)j * N + ic])));
                    double rr = fma(0.25, d4, -tgt);                                             // d_ij - T_ij
                    const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) |
                                     ((kind == GIK_TERM_UP) & (rr > 0.0));
                    rr = act ? rr : 0.0;
                    fpart = fma(rr, rr, fpart);
                    gpart[m][0] = fma(rr, dx, gpart[m][0]);      // 2 r (x_i - x_j)
                    gpart[m][1] = fma(rr, dy, gpart[m][1]);
                    gpart[m][2] = fma(rr, dz, gpart[m][2]);
                    if (m <= jb) C2[off[m]] = 2.0 * rr;          // the mirrored pair computes the same bits
                    off[m] += step[m];
                    if (fj && m >= mF) rsum[m] += 2.0 * rr;      // clique pair: its c2 w term is split (pair_pass_hess)
                    amask[m] |= act ? bit : 0ull;
                }
            }
        }
        double react[3] = {0.0, 0.0, 0.0};
        if (hub_warp) {
            // second terms of the pairs (i, hub), evaluated from the partner's side only; the hub node gets the
            // reaction (the term seen from the hub is the mirror image), the cost counts them twice before the 0.5
            const double px = P2[hub], py = P2[NPAD + hub], pz = P2[2 * NPAD + hub];
            double fh = 0.0;
            hact = 0u;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const int i = lane + 32 * m;
                const int ic = min(i, N - 1);
                const unsigned kind = i < N ? tb.hub_kind[i] : 3u;
                const double dx = xt[m][0] - px, dy = xt[m][1] - py, dz = xt[m][2] - pz;
                const double d4 = dx * dx + dy * dy + dz * dz;
                double rr = fma(0.25, d4, -tb.hub_target[ic]);
                const bool act = (kind == GIK_TERM_EQ) | ((kind == GIK_TERM_LO) & (rr < 0.0)) |
                                 ((kind == GIK_TERM_UP) & (rr > 0.0));
                rr = act ? rr : 0.0;
                fh = fma(rr, rr, fh);
                gpart[m][0] = fma(rr, dx, gpart[m][0]);
                gpart[m][1] = fma(rr, dy, gpart[m][1]);
                gpart[m][2] = fma(rr, dz, gpart[m][2]);
                react[0] = fma(rr, dx, react[0]);
                react[1] = fma(rr, dy, react[1]);
                react[2] = fma(rr, dz, react[2]);
                Hc2[i] = 2.0 * rr;
                hact |= act ? (1u << m) : 0u;
            }
            fpart = fma(2.0, fh, fpart);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int q = 0; q < 3; ++q) react[q] += __shfl_xor_sync(GIK_FULL_MASK, react[q], off, 32);
            }
        }
        store_partials(gpart);
        if (has_fast) {
#pragma unroll
            for (int m = 0; m < NB; ++m) RS[warp * NPAD + lane + 32 * m] = rsum[m];
        }
        if (hub_warp) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) Zp[(warp * 3 + q) * NPAD + hub] -= react[q];
            }
        }
        double f1[1] = {0.5 * fpart};   // every undirected term is seen from both ends
        warp_sum_store<1>(f1, redF + warp, lane == 0);
    };

    // costs.py:171-207 at the cached point along the direction published in V: partial sums into Zp
    auto pair_pass_hess = [&]() {
        double wt[NB][3], z[NB][3];
        load_tile(V, wt);
#pragma unroll
        for (int m = 0; m < NB; ++m) { z[m][0] = 0.0; z[m][1] = 0.0; z[m][2] = 0.0; }
        // neighbours [ja, jb2) of one 32-node block; MF = first tile block whose pairs with these neighbours lie inside
        // the equality clique (compile-time, so that the unrolled tile loop has no branches): for those only
        // -c2_ij delta_j is accumulated here -- (sum_j c2_ij) delta_i and the <D,w> D part (15 moments of delta over the
        // clique) are added by the node's owner
        auto hess_range = [&](auto mf_tag, int ja, int jb2, int (&off)[NB], const int (&step)[NB], unsigned long long &bit) {
            constexpr int MF = decltype(mf_tag)::value;
#pragma unroll 2
            for (int j = ja; j < jb2; ++j) {
                const double vx = V[j], vy = V[NPAD + j], vz = V[2 * NPAD + j];
                double px = 0.0, py = 0.0, pz = 0.0;
                if (MF > 0) { px = P2[j]; py = P2[NPAD + j]; pz = P2[2 * NPAD + j]; }
#pragma unroll
                for (int m = 0; m < NB; ++m) {
                    const double c2 = C2[off[m]];
                    off[m] += step[m];
                    if (m >= MF) {
                        z[m][0] = fma(-c2, vx, z[m][0]);
                        z[m][1] = fma(-c2, vy, z[m][1]);
                        z[m][2] = fma(-c2, vz, z[m][2]);
                    } else {
                        const double dx = xt[m][0] - px, dy = xt[m][1] - py, dz = xt[m][2] - pz;
                        const double wx = wt[m][0] - vx, wy = wt[m][1] - vy, wz = wt[m][2] - vz;
                        double s = dx * wx + dy * wy + dz * wz;                  // 2 <D, w>
                        s = (amask[m] & bit) ? s : 0.0;
                        z[m][0] = fma(s, dx, fma(c2, wx, z[m][0]));              // 4 <D,w> D + 2 r w
                        z[m][1] = fma(s, dy, fma(c2, wy, z[m][1]));
                        z[m][2] = fma(s, dz, fma(c2, wz, z[m][2]));
                    }
                }
                bit <<= 1;
            }
        };
        for (int jb = jlo >> 5; jb <= (jhi - 1) >> 5; ++jb) {
            const int j0 = max(jlo, 32 * jb), j1 = min(jhi, 32 * jb + 32);
            int off[NB], step[NB];
            c2_walk(jb, j0, off, step);
            unsigned long long bit = 1ull << (j0 - jlo);
            const int jm = has_fast ? min(max(j0, jF), j1) : j1;     // [j0, jm): no clique neighbour, [jm, j1): clique
            hess_range(std::integral_constant<int, NB>(), j0, jm, off, step, bit);
            if (jm < j1) {
                switch (mF) {
                    case 0: hess_range(std::integral_constant<int, 0>(), jm, j1, off, step, bit); break;
                    case 1: hess_range(std::integral_constant<int, (1 < NB ? 1 : NB)>(), jm, j1, off, step, bit); break;
                    case 2: hess_range(std::integral_constant<int, (2 < NB ? 2 : NB)>(), jm, j1, off, step, bit); break;
                    default: hess_range(std::integral_constant<int, (3 < NB ? 3 : NB)>(), jm, j1, off, step, bit); break;
                }
            }
        }
        double react[3] = {0.0, 0.0, 0.0};
        if (hub_warp) {
            const double px = P2[hub], py = P2[NPAD + hub], pz = P2[2 * NPAD + hub];
            const double vx = V[hub], vy = V[NPAD + hub], vz = V[2 * NPAD + hub];
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                const double c2 = Hc2[lane + 32 * m];
                const double dx = xt[m][0] - px, dy = xt[m][1] - py, dz = xt[m][2] - pz;
                const double wx = wt[m][0] - vx, wy = wt[m][1] - vy, wz = wt[m][2] - vz;
                double s = dx * wx + dy * wy + dz * wz;
                s = (hact >> m & 1u) ? s : 0.0;
                const double ax = fma(s, dx, c2 * wx), ay = fma(s, dy, c2 * wy), az = fma(s, dz, c2 * wz);
                z[m][0] += ax; z[m][1] += ay; z[m][2] += az;
                react[0] += ax; react[1] += ay; react[2] += az;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int q = 0; q < 3; ++q) react[q] += __shfl_xor_sync(GIK_FULL_MASK, react[q], off, 32);
            }
        }
        store_partials(z);
        if (hub_warp) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) Zp[(warp * 3 + q) * NPAD + hub] -= react[q];
            }
        }
    };

    auto total_cost = [&]() -> double {
        double f = redF[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) f += redF[w];
        return f;
    };

    auto gather_rs = [&]() -> double {
        double acc = RS[tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += RS[w * NPAD + tid];
        return acc;
    };
    // owner of a node whose clique pairs took the fast path: Z_i += sum_j (X_i - X_j)(X_i - X_j)^T (w_i - w_j) over the clique
    // (moments in momS) + (sum_j c2_ij) w_i
    auto add_clique_terms = [&](double (&Z)[3], const double (&xo)[3], const double (&w)[3], double rs) {
        const double X[3] = {2.0 * xo[0], 2.0 * xo[1], 2.0 * xo[2]};
        const double *e1 = momS, *E = momS + 3, *v3 = momS + 12, *m1 = momS + 16, *M2 = momS + 19;
        const double xd = X[0] * w[0] + X[1] * w[1] + X[2] * w[2];
        const double beta = E[0] + E[4] + E[8];
        const double sc = fma(m0, xd, beta) - (X[0] * e1[0] + X[1] * e1[1] + X[2] * e1[2]) -
                          (m1[0] * w[0] + m1[1] * w[1] + m1[2] * w[2]);
        const double Mw[3] = {M2[0] * w[0] + M2[1] * w[1] + M2[2] * w[2], M2[1] * w[0] + M2[3] * w[1] + M2[4] * w[2],
                              M2[2] * w[0] + M2[4] * w[1] + M2[5] * w[2]};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double EX = E[3 * q] * X[0] + E[3 * q + 1] * X[1] + E[3 * q + 2] * X[2];
            Z[q] += fma(X[q], sc, fma(-m1[q], xd, EX + Mw[q] - v3[q])) + rs * w[q];
        }
    };

    // parked problems of the incoming queue are resumed before any new problem starts (gik_rtr.cuh)
    int n_res = 0;
    if (a.carry_in) {
        const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(a.carry_in);
        n_res = min(h->count, h->capacity);
    }

    for (;;) {
        if (tid == 0) s_b[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int w = s_b[0];
        __syncthreads();
        if (w >= n_res + a.B) break;
        const bool resumed = w < n_res;
        const int b = w - n_res;
        const double *ent = resumed ? gik_carry_entry(a.carry_in, w) : nullptr;
        const unsigned long long *entp = reinterpret_cast<const unsigned long long *>(ent);
        const double *goal_row = resumed ? reinterpret_cast<const double *>(entp[CW_GOAL])
                                         : a.goal_d2 + (size_t)b * a.n_goal;

        double x[3] = {0.0, 0.0, 0.0}, g[3] = {0.0, 0.0, 0.0}, eta[3], Heta[3], r[3], dl[3], Hd[3] = {0.0, 0.0, 0.0};
        if (owner) {
            const double *src = resumed ? ent + CW_X + 3 * node : a.Y_init + ((size_t)b * N + node) * 3;
            x[0] = src[0]; x[1] = src[1]; x[2] = src[2];
            if (resumed) {
                src = ent + CW_X + 3 * (N + node);
                g[0] = src[0]; g[1] = src[1]; g[2] = src[2];
            }
        }
        for (int k = tid; k < a.n_goal; k += kThreads) goal[k] = goal_row[k];
        publish(P2, x, 2.0);
        if (gp >= 0) {
            for (int i = tid; i < N; i += kThreads) {
                Tp[i] = tb.target[(size_t)gp * N + i];
                Tq[i] = tb.target[(size_t)gq * N + i];
            }
        }
        __syncthreads();
        if (gp >= 0) {
            for (int e = tid; e < tb.n_goal_edges; e += kThreads) {
                const int i = tb.goal_i[e], j = tb.goal_j[e];
                const double t = goal[tb.goal_slot[e]];
                if (j == gp) Tp[i] = t; else if (i == gp) Tp[j] = t;
                if (j == gq) Tq[i] = t; else if (i == gq) Tq[j] = t;
            }
        }
        __syncthreads();

        double fx, gg, Mi[6], Delta, rs = 0.0;
        int k_outer, inner_total;
        // clique constants of the cached point (momS[16..25)) and this node's sum of c2 over its clique pairs; needs the
        // pair pass at x behind a barrier.  The same per-thread values and the same summation tree as at an accepted step.
        auto refresh_clique = [&]() {
            if (!has_fast) return;
            if (owner) rs = gather_rs();
            const double X[3] = {2.0 * x[0], 2.0 * x[1], 2.0 * x[2]};
            double cc[9];
            clique_consts(X, in_clique, cc);
            warp_sum_store<9>(cc, redB + warp * kRedB, first_of_own_warp);
            __syncthreads();
            if (tid < 9) {          // the sum block_sum_load forms, without indexing a register array by tid
                double sm = redB[tid];
#pragma unroll
                for (int w = 1; w < kOwnWarps; ++w) sm += redB[w * kRedB + tid];
                momS[16 + tid] = sm;
            }
            __syncthreads();
        };
        if (resumed) {
            fx = ent[CW_FX]; gg = ent[CW_GG]; Delta = ent[CW_DELTA];
#pragma unroll
            for (int k = 0; k < 6; ++k) Mi[k] = ent[CW_MI + k];
            const unsigned long long cnt = entp[CW_COUNTS];
            k_outer = (int)(cnt & 0xffffffffu);
            inner_total = (int)(cnt >> 32);
            if (tid == 0) s_t[0] = entp[CW_T0];
            pair_pass_cost();   // tile, pair cache and activity masks at x, as after a rejected step
            __syncthreads();
            refresh_clique();
        } else {
            pair_pass_cost();
            __syncthreads();
            if (owner) gather(g);
            // threads without a node carry zeros in all node state, so every warp can run the butterflies
            double v[7] = {g[0] * g[0] + g[1] * g[1] + g[2] * g[2],
                           x[0] * x[0], x[0] * x[1], x[0] * x[2], x[1] * x[1], x[1] * x[2], x[2] * x[2]};
            warp_sum_store<7>(v, redA + warp * kRedA, first_of_own_warp);
            __syncthreads();
            block_sum_load<7, kRedA>(v, redA);
            fx = total_cost();
            gg = v[0];
            gik_sylvester_inverse(v + 1, Mi);
            Delta = o.Delta0;
            k_outer = 0;
            inner_total = 0;
            if (tid == 0) s_t[0] = a.maxtime_ns ? gik_globaltimer() : 0ull;
            refresh_clique();
        }
        const int inner_entry = inner_total;
        bool may_park = a.carry_out != nullptr;
        int park_slot = -1;
        double norm_grad = sqrt(gg);
        int status = GIK_STATUS_MAXITER;
        if (!(isfinite(fx) && isfinite(gg))) {
            status = GIK_STATUS_NAN;
        } else {
            for (;;) {
                // ================= tCG (trust_region.py:436-599), eta0 = 0, precon = identity
#pragma unroll
                for (int q = 0; q < 3; ++q) { eta[q] = 0.0; Heta[q] = 0.0; r[q] = g[q]; dl[q] = -g[q]; }
                double e_Pe = 0.0, r_r = gg;
                const double norm_r0 = sqrt(r_r);
                double z_r = r_r, d_Pd = r_r, e_Pd = 0.0, model_value = 0.0;
                const double pw = o.theta == 1.0 ? norm_r0 : pow(norm_r0, o.theta);
                const double r_target = norm_r0 * fmin(pw, o.kappa);
                const double r_target2 = r_target * r_target;
                const double Delta2 = Delta * Delta;
                int stop = MAX_INNER_ITER;
                int j = 0;
                __syncthreads();          // every reader of V / redA / redB of the previous phase is done
                publish_dir(dl, x);
                for (j = 0; j < o.maxinner; ++j) {
                    __syncthreads();                                   // (1) delta published
                    if (has_fast) direction_moments();                 // -> momS, read by the owners after barrier (2)
                    pair_pass_hess();
                    __syncthreads();                                   // (2) partial sums complete
                    if (owner) gather(Hd);                             // raw Z; projected below
                    else { Hd[0] = 0.0; Hd[1] = 0.0; Hd[2] = 0.0; }
                    if (fast_owner) add_clique_terms(Hd, x, dl, rs);
                    double v[7];
                    v[0] = dl[0] * Hd[0] + dl[1] * Hd[1] + dl[2] * Hd[2];
                    v[1] = Hd[1] * x[2] - Hd[2] * x[1];                // c = sum Z_i x Y_i
                    v[2] = Hd[2] * x[0] - Hd[0] * x[2];
                    v[3] = Hd[0] * x[1] - Hd[1] * x[0];
                    v[4] = dl[1] * x[2] - dl[2] * x[1];                // u = sum delta_i x Y_i
                    v[5] = dl[2] * x[0] - dl[0] * x[2];
                    v[6] = dl[0] * x[1] - dl[1] * x[0];
                    warp_sum_store<7>(v, redA + warp * kRedA, first_of_own_warp);
                    __syncthreads();                                   // (3)
                    block_sum_load<7, kRedA>(v, redA);
                    double om[3];
                    gik_sym_mul(Mi, v + 1, om);
                    Hd[0] -= x[1] * om[2] - x[2] * om[1];
                    Hd[1] -= x[2] * om[0] - x[0] * om[2];
                    Hd[2] -= x[0] * om[1] - x[1] * om[0];
                    const double d_Hd = v[0] - (om[0] * v[4] + om[1] * v[5] + om[2] * v[6]);
                    ++inner_total;
                    const double alpha = gik_div(z_r, d_Hd, gik_rcp(d_Hd));
                    const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
                    if (!(d_Hd > 0.0) || e_Pe_new >= Delta2) {   // also catches NaN
                        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            eta[q] = fma(tau, dl[q], eta[q]);
                            Heta[q] = fma(tau, Hd[q], Heta[q]);
                        }
                        stop = d_Hd <= 0.0 ? NEGATIVE_CURVATURE : EXCEEDED_TR;
                        break;
                    }
                    e_Pe = e_Pe_new;
                    double ne[3], nh[3], nr[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        ne[q] = fma(alpha, dl[q], eta[q]);
                        nh[q] = fma(alpha, Hd[q], Heta[q]);
                        nr[q] = fma(alpha, Hd[q], r[q]);
                    }
                    {
                        double sdot[3] = {0.0, 0.0, 0.0};
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            sdot[0] = fma(ne[q], g[q], sdot[0]);
                            sdot[1] = fma(ne[q], nh[q], sdot[1]);
                            sdot[2] = fma(nr[q], nr[q], sdot[2]);
                        }
                        warp_sum_store<3>(sdot, redB + warp * kRedB, first_of_own_warp);
                    }
                    __syncthreads();                                   // (4)
                    double sd[3];
                    block_sum_load<3, kRedB>(sd, redB);
                    const double new_model_value = sd[0] + 0.5 * sd[1];
                    if (new_model_value >= model_value) {
                        stop = MODEL_INCREASED;
                        break;
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) { eta[q] = ne[q]; Heta[q] = nh[q]; r[q] = nr[q]; }
                    model_value = new_model_value;
                    r_r = sd[2];
                    if (j >= o.mininner && r_r <= r_target2) {
                        stop = o.kappa < pw ? REACHED_TARGET_LINEAR : REACHED_TARGET_SUPERLINEAR;
                        break;
                    }
                    const double beta = r_r / z_r;
                    z_r = r_r;
#pragma unroll
                    for (int q = 0; q < 3; ++q) dl[q] = fma(beta, dl[q], -r[q]);
                    e_Pd = beta * (e_Pd + alpha * d_Pd);
                    d_Pd = z_r + beta * beta * d_Pd;
                    publish_dir(dl, x);            // the pair pass read V before barrier (2)
                }
                const int numit = j < o.maxinner ? j : o.maxinner - 1;

                // ================= proposal x + eta (trust_region.py:248-251)
#pragma unroll
                for (int q = 0; q < 3; ++q) dl[q] = x[q] + eta[q];
                __syncthreads();
                publish(P2, dl, 2.0);
                __syncthreads();
                pair_pass_cost();
                __syncthreads();
                if (owner) gather(Hd);
                else { Hd[0] = 0.0; Hd[1] = 0.0; Hd[2] = 0.0; }
                const double rs_prop = (has_fast && owner) ? gather_rs() : 0.0;
                double v[18];
                v[0] = g[0] * eta[0] + g[1] * eta[1] + g[2] * eta[2];
                v[1] = eta[0] * Heta[0] + eta[1] * Heta[1] + eta[2] * Heta[2];
                v[2] = Hd[0] * Hd[0] + Hd[1] * Hd[1] + Hd[2] * Hd[2];
                v[3] = dl[0] * dl[0]; v[4] = dl[0] * dl[1]; v[5] = dl[0] * dl[2];
                v[6] = dl[1] * dl[1]; v[7] = dl[1] * dl[2]; v[8] = dl[2] * dl[2];
                if (has_fast) {
                    const double X[3] = {2.0 * dl[0], 2.0 * dl[1], 2.0 * dl[2]};
                    clique_consts(X, in_clique, v + 9);
                    warp_sum_store<18>(v, redA + warp * kRedA, first_of_own_warp);
                    __syncthreads();
                    block_sum_load<18, kRedA>(v, redA);
                } else {
                    warp_sum_store<9>(v, redA + warp * kRedA, first_of_own_warp);
                    __syncthreads();
                    block_sum_load<9, kRedA>(v, redA);
                }
                const double fx_prop = total_cost();
                double rhonum = fx - fx_prop;
                double rhoden = -v[0] - 0.5 * v[1];
                const double rho_reg = fmax(1.0, fabs(fx)) * eps * o.rho_regularization;
                rhonum += rho_reg;
                rhoden += rho_reg;
                const bool model_decreased = rhoden >= 0.0;
                const double rho = rhonum / rhoden;
                const double Delta_used = Delta;
                if (rho < 0.25 || !model_decreased || isnan(rho)) {
                    Delta = Delta / 4.0;
                } else if (rho > 0.75 && (stop == NEGATIVE_CURVATURE || stop == EXCEEDED_TR)) {
                    Delta = fmin(2.0 * Delta, o.Delta_bar);
                }
                const bool accept = model_decreased && rho > o.rho_prime;
                if (accept) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x[q] = dl[q]; g[q] = Hd[q]; }
                    fx = fx_prop;
                    gg = v[2];
                    norm_grad = sqrt(gg);
                    gik_sylvester_inverse(v + 3, Mi);
                    rs = rs_prop;
                    if (has_fast && tid < 9) {   // read after the barriers of the next tCG start
                        double sm = redA[9 + tid];
#pragma unroll
                        for (int w = 1; w < kOwnWarps; ++w) sm += redA[w * kRedA + 9 + tid];
                        momS[16 + tid] = sm;
                    }
                } else {
                    // rejected: bring the exchange buffer, the tile and the pair cache back to x
                    __syncthreads();
                    publish(P2, x, 2.0);
                    __syncthreads();
                    pair_pass_cost();
                }
                if (a.trace && !resumed && k_outer < a.trace_rows && tid == 0) {
                    double *row = a.trace + ((size_t)b * a.trace_rows + k_outer) * 6;
                    row[0] = Delta_used;
                    row[1] = (double)numit;
                    row[2] = (double)stop;
                    row[3] = fx_prop;
                    row[4] = accept ? 1.0 : 0.0;
                    row[5] = accept ? norm_grad : nan("");
                }
                ++k_outer;
                // pymanopt Solver._check_stopping_criterion: maxtime, then maxiter, then mingradnorm
                if (a.maxtime_ns) {
                    __syncthreads();
                    if (tid == 0) s_t[1] = gik_globaltimer() - s_t[0];
                    __syncthreads();
                    if (s_t[1] >= a.maxtime_ns) { status = GIK_STATUS_MAXTIME; break; }
                }
                if (k_outer >= o.maxiter) { status = GIK_STATUS_MAXITER; break; }
                if (norm_grad < o.mingradnorm) { status = GIK_STATUS_CONVERGED; break; }
                if (may_park && inner_total - inner_entry >= a.inner_budget) {
                    __syncthreads();
                    if (tid == 0) s_b[1] = gik_try_park(a, n_res + a.B);
                    __syncthreads();
                    park_slot = s_b[1];
                    if (park_slot >= 0) { status = GIK_STATUS_PENDING; break; }
                    if (park_slot == -1) may_park = false;   // queue full: run this problem to its end
                }
            }
        }
        // ---- final values (or, for a parked problem, its current ones) go where the problem came from
        double *Yrow = resumed ? reinterpret_cast<double *>(entp[CW_Y]) : a.Y_out + (size_t)b * N * 3;
        double *cx = status == GIK_STATUS_PENDING ? gik_carry_slot(a.carry_out, park_slot) : nullptr;
        if (owner) {
            double *dst = Yrow + node * 3;
            dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
            if (cx) {
                dst = cx + CW_X + 3 * node;
                dst[0] = x[0]; dst[1] = x[1]; dst[2] = x[2];
                dst += 3 * N;
                dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2];
            }
        }
        if (tid == 0) {
            const double sg[3] = {0.0, 0.0, 0.0};   // only the warp kernels carry sum g_i x Y_i
            gik_finish_problem(a, resumed, b, entp, goal_row, cx, s_t[0], status, k_outer, inner_total, fx, gg,
                               norm_grad, Delta, Mi, sg, Yrow);
        }
        __syncthreads();
    }
}

size_t cta_smem_bytes(int N, int n_goal)
{
    const int goal_pad = (n_goal + 1) & ~1;
    size_t smem = ((size_t)c2_doubles(N) + (6 + 3 * kWarps + 3 + kWarps + 1) * NPAD + kMom + kOwnWarps * (kRedA + kRedB) + kWarps + goal_pad) *
                      sizeof(double) + 2 * sizeof(int) + 2 * sizeof(unsigned long long);
    return (smem + 15) & ~(size_t)15;
}

}  // namespace

template <int NB>
static int launch_cta(const GikPlan *p, RtrArgs &a, const CtaTables &tb, size_t smem, cudaStream_t st)
{
    auto kern = k_rtr_cta<NB>;
    static size_t cached_smem = ~(size_t)0;
    static int cached_per_sm = 0, cached_dev = -1;
    if (cached_smem != smem || cached_dev != p->device) {
        GIK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GIK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        cached_per_sm = per_sm < 1 ? 1 : per_sm;
        cached_smem = smem;
        cached_dev = p->device;
    }
    int blocks = p->sm_count * cached_per_sm;
    if (!a.carry_in && blocks > a.B) blocks = a.B;
    GIK_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int32_t), st));
    kern<<<blocks, kThreads, smem, st>>>(a, tb);
    return gik_check_cuda(cudaGetLastError(), "k_rtr_cta launch");
}

int gik_launch_rtr_cta(const GikPlan *p, RtrArgs &a, cudaStream_t st)
{
    if (!p->dense_target || p->N > NPAD || p->N <= 32) return 1;
    const int N = p->N;
    const size_t smem = cta_smem_bytes(N, p->n_goal);
    if (smem > 227 * 1024) return 1;
    CtaTables tb;
    tb.target = p->dense_target;
    tb.kind = p->dense_kind;
    tb.goal_i = p->dense_goal_i;
    tb.goal_j = p->dense_goal_j;
    tb.goal_slot = p->dense_goal_slot;
    tb.n_goal_edges = p->n_dense_goal;
    tb.gp = p->n_dense_goal > 0 ? p->dense_goal_p : -1;
    tb.gq = p->n_dense_goal > 0 ? p->dense_goal_q : -1;
    tb.perm = p->dense_perm;
    tb.jF = p->dense_clique_start;
    tb.mF = (p->dense_clique_start + 31) / 32;      // first 32-node block that lies inside the clique
    if (tb.mF > (N + 31) / 32) tb.mF = (N + 31) / 32;
    tb.hub = p->dense_hub;
    tb.hub_kind = p->dense_hub_kind;
    tb.hub_target = p->dense_hub_target;
    switch ((N + 31) / 32) {
        case 2: return launch_cta<2>(p, a, tb, smem, st);
        case 3: return launch_cta<3>(p, a, tb, smem, st);
        default: return launch_cta<4>(p, a, tb, smem, st);
    }
}
