// gik_rtr.cuh -- argument block shared by the trust-region kernels.
#pragma once
#include "gik_common.cuh"

// tCG stop reasons, numbered as in trust_region.py:62-69
enum { NEGATIVE_CURVATURE = 0, EXCEEDED_TR, REACHED_TARGET_LINEAR, REACHED_TARGET_SUPERLINEAR,
       MAX_INNER_ITER, MODEL_INCREASED };

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
};

// gik_rtr_fast.cu: one warp per problem, slot data cached in registers (N <= 32).
// Returns GIK_OK, or 1 if no specialisation covers the plan (caller falls back to k_rtr).
int gik_launch_rtr_fast(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_fast.cu: one warp per problem, two nodes per lane, slot cache in shared memory (32 < N <= 64, sparse).
int gik_launch_rtr_fast2(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_duo.cu: two problems per warp in lock-step (N <= 16); same return convention.
int gik_launch_rtr_duo(const GikPlan *p, RtrArgs &a, cudaStream_t st);
// gik_rtr_cta.cu: one CTA per problem with a dense target matrix (32 < N <= 128); same convention.
int gik_launch_rtr_cta(const GikPlan *p, RtrArgs &a, cudaStream_t st);
