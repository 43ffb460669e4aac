// gik_rtr.cuh -- argument block shared by the trust-region kernels.
#pragma once
#include "gik_common.cuh"

// tCG stop reasons, numbered as in trust_region.py:62-69
enum { NEGATIVE_CURVATURE = 0, EXCEEDED_TR, REACHED_TARGET_LINEAR, REACHED_TARGET_SUPERLINEAR,
       MAX_INNER_ITER, MODEL_INCREASED };

// ---------------------------------------------------------------------------------------------------
// Deferred problems ("carry" queue, gik_rtr_solve_sliced).  A problem whose tCG iterations inside one
// launch reach opts.inner_budget parks at the next outer-iteration boundary: everything the outer loop
// of trust_region.py:179-422 carries from one iteration to the next -- x, the half gradient, f(x),
// <g,g>, Delta, the iteration counters, the inverse of tr(X) I - X and sum g_i x Y_i -- goes to an entry of the
// launch's outgoing queue together with the addresses its final values belong to.  The next launch on
// that queue resumes the parked problems first.  The state is stored bit for bit and everything
// derived from it (slot cache, exchange buffers) is rebuilt by the same code a rejected step uses, so
// a parked problem follows exactly the trajectory it would have followed in one piece.
struct GikCarryHdr {
    int32_t count;      // parked problems (written by the launch that fills the queue)
    int32_t capacity;   // entries the buffer holds
    int32_t stride;     // 8-byte words per entry
    int32_t full;       // problems that wanted to park but found the queue full (they ran on instead)
};
// entry layout in 8-byte words
enum { CW_Y = 0, CW_F, CW_GN, CW_ITERS, CW_STATUS, CW_NINNER, CW_GOAL, CW_PENDING, CW_T0, CW_COUNTS,
       CW_FX, CW_GG, CW_DELTA, CW_MI, CW_SG = CW_MI + 6, CW_X = CW_SG + 3 };
inline __host__ __device__ int gik_carry_stride(int N) { return CW_X + 6 * N; }

struct RtrArgs {
    const uint32_t *slot_info;
    const double *slot_target;
    const int32_t *deg;
    int N, n_goal, maxdeg, tables_in_smem;
    const double *goal_d2;
    const double *Y_init;
    int B;
    GikSolveOpts o;
    double *Y_out, *f, *gradnorm;
    int32_t *iters, *status, *n_inner;
    double *trace;
    int trace_rows;
    int32_t *work_counter;
    // gik_rtr_solve_sliced
    int inner_budget;            // 0: never park
    const char *carry_in;        // GikCarryHdr + entries to resume first (may be null)
    char *carry_out;             // where this launch parks problems (may be null: never park)
    int32_t *pending;            // += 1 per fresh problem parked; an entry carries the address along and the
                                 // launch that finishes the problem does -= 1 (may be null)
    unsigned long long maxtime_ns;   // 0: no time limit
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long gik_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ const double *gik_carry_entry(const char *q, int slot)
{
    const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(q);
    return reinterpret_cast<const double *>(q + sizeof(GikCarryHdr)) + (size_t)slot * h->stride;
}
// Reserve an entry of the outgoing queue; -1 when it is full (the caller keeps iterating instead).
__device__ __forceinline__ int gik_carry_reserve(char *q)
{
    GikCarryHdr *h = reinterpret_cast<GikCarryHdr *>(q);
    const int slot = atomicAdd(&h->count, 1);
    if (slot < h->capacity) return slot;
    atomicSub(&h->count, 1);
    atomicAdd(&h->full, 1);
    return -1;
}
// Park decision of one problem, taken by ONE thread of its group at an outer-iteration boundary once the problem has
// spent its budget in this launch: it parks only when the launch has no new work left to hand out (the work counter
// has passed the last item).  Until then a long problem keeps its warp -- the other warps pull the new problems --
// so it advances without interruption for as long as the launch has a reason to live; once the queue is dry it
// leaves within one outer iteration and the launch ends with its work, not with its slowest problem.
// Returns the reserved entry (>= 0), -1: queue full (never try again), -2: not now.
__device__ __forceinline__ int gik_try_park(const RtrArgs &a, int n_items)
{
    if (*reinterpret_cast<volatile int32_t *>(a.work_counter) < n_items) return -2;
    return gik_carry_reserve(a.carry_out);
}

__device__ __forceinline__ double *gik_carry_slot(char *q, int slot)
{
    const GikCarryHdr *h = reinterpret_cast<const GikCarryHdr *>(q);
    return reinterpret_cast<double *>(q + sizeof(GikCarryHdr)) + (size_t)slot * h->stride;
}

// End of a problem's stay in a launch, executed by ONE thread of the group that owned it: the optlog final_values
// (or, for a problem that parks, its current values with status PENDING) go to the locations the problem came
// with; a parking problem also gets the scalar part of its outgoing entry `cx` (the owner lanes store x and g).
__device__ __forceinline__ void gik_finish_problem(const RtrArgs &a, bool resumed, int b, const unsigned long long *entp,
                                                   const double *goal_row, double *cx, unsigned long long t0,
                                                   int status, int k_outer, int inner_total, double fx, double gg,
                                                   double norm_grad, double Delta, const double (&Mi)[6],
                                                   const double (&sg)[3], double *Yrow)
{
    double *pf = resumed ? reinterpret_cast<double *>(entp[CW_F]) : a.f + b;
    double *pgn = resumed ? reinterpret_cast<double *>(entp[CW_GN]) : a.gradnorm + b;
    int32_t *pit = resumed ? reinterpret_cast<int32_t *>(entp[CW_ITERS]) : a.iters + b;
    int32_t *pst = resumed ? reinterpret_cast<int32_t *>(entp[CW_STATUS]) : a.status + b;
    int32_t *pni = resumed ? reinterpret_cast<int32_t *>(entp[CW_NINNER]) : (a.n_inner ? a.n_inner + b : nullptr);
    int32_t *pend = resumed ? reinterpret_cast<int32_t *>(entp[CW_PENDING]) : a.pending;
    if (cx) {
        unsigned long long *op = reinterpret_cast<unsigned long long *>(cx);
        op[CW_Y] = (unsigned long long)Yrow;
        op[CW_F] = (unsigned long long)pf;
        op[CW_GN] = (unsigned long long)pgn;
        op[CW_ITERS] = (unsigned long long)pit;
        op[CW_STATUS] = (unsigned long long)pst;
        op[CW_NINNER] = (unsigned long long)pni;
        op[CW_GOAL] = (unsigned long long)goal_row;
        op[CW_PENDING] = (unsigned long long)pend;
        op[CW_T0] = t0;
        op[CW_COUNTS] = (unsigned long long)(unsigned)k_outer | ((unsigned long long)(unsigned)inner_total << 32);
        cx[CW_FX] = fx; cx[CW_GG] = gg; cx[CW_DELTA] = Delta;
#pragma unroll
        for (int k = 0; k < 6; ++k) cx[CW_MI + k] = Mi[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) cx[CW_SG + k] = sg[k];
        if (!resumed && pend) atomicAdd(pend, 1);
    }
    *pf = fx;
    *pgn = norm_grad;
    *pit = k_outer;
    *pst = status;
    if (pni) *pni = inner_total;
    if (resumed && !cx && pend) atomicSub(pend, 1);
}
#endif

// gik_rtr_fast.cu: one warp per problem, slot data cached in registers (N <= 32).
// Returns GIK_OK, or 1 if no specialisation covers the plan (caller falls back to k_rtr).
int gik_launch_rtr_fast(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_fast.cu: one warp per problem, two nodes per lane, slot cache in shared memory (32 < N <= 64, sparse).
int gik_launch_rtr_fast2(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_duo.cu: two problems per warp in lock-step (N <= 16); same return convention.
int gik_launch_rtr_duo(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_cta.cu: one CTA per problem with a dense target matrix (32 < N <= 128); same convention.
int gik_launch_rtr_cta(const GikPlan *p, RtrArgs &a, cudaStream_t st);
