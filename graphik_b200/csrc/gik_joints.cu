// gik_joints.cu -- joint-angle recovery and forward kinematics, one thread per problem.
//
//   ProblemGraphRevolute.joint_variables   graphs/graph_revolute.py:251-318
//   RobotRevolute.pose (all joints)        robots/robot_revolute.py:85-103
//   ProblemGraph.realization points        graphs/graph_base.py:112-121
//
// Both are sequential along the kinematic chain and independent across the batch.
#include "gik_common.cuh"

namespace {

struct Frame {  // rigid transform, row-major rotation + translation
    double R[9];
    double t[3];
};

__device__ __forceinline__ void frame_load(const double *T, Frame &f)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        f.R[3 * r] = T[4 * r];
        f.R[3 * r + 1] = T[4 * r + 1];
        f.R[3 * r + 2] = T[4 * r + 2];
        f.t[r] = T[4 * r + 3];
    }
}

// f <- f * g
__device__ __forceinline__ void frame_mul(Frame &f, const Frame &g)
{
    double R[9], t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[3 * r + c] = f.R[3 * r] * g.R[c] + f.R[3 * r + 1] * g.R[3 + c] + f.R[3 * r + 2] * g.R[6 + c];
        t[r] = f.R[3 * r] * g.t[0] + f.R[3 * r + 1] * g.t[1] + f.R[3 * r + 2] * g.t[2] + f.t[r];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) f.R[k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) f.t[k] = t[k];
}

// f <- f * Rz(theta)
__device__ __forceinline__ void frame_rotz(Frame &f, double theta)
{
    double s, c;
    sincos(theta, &s, &c);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double a = f.R[3 * r], b = f.R[3 * r + 1];
        f.R[3 * r] = a * c + b * s;
        f.R[3 * r + 1] = -a * s + b * c;
    }
}

__device__ __forceinline__ void unit3(double v[3])
{
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n != 0.0) { v[0] /= n; v[1] /= n; v[2] /= n; }
}

// graph_revolute.py:251-318.  Node layout: p0=0, x=1, y=2, q0=3, p_i = 2+2i, q_i = 3+2i.
__global__ void k_joints(int N, int n, const double *__restrict__ T0, const double *__restrict__ Trel,
                         const double *__restrict__ qs0, int z_aligned, const double *__restrict__ Y,
                         const double *__restrict__ T_goal, int B, double *__restrict__ q_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *Yb = Y + (size_t)b * N * 3;
    // base frame from the solved points (graph_revolute.py:270-279): R = [x, -y, z], origin p0
    double p0[3] = {Yb[0], Yb[1], Yb[2]};
    double ex[3], ey[3], ez[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ex[k] = Yb[3 + k] - p0[k];
        ey[k] = Yb[6 + k] - p0[k];
        ez[k] = Yb[9 + k] - p0[k];
    }
    unit3(ex); unit3(ey); unit3(ez);
    Frame Tp;
    frame_load(T0, Tp);
    Frame rel;
    for (int i = 1; i <= n; ++i) {
        const double *pc = Yb + 3 * (2 + 2 * i), *qc = Yb + 3 * (3 + 2 * i);
        double d[3] = {qc[0] - pc[0], qc[1] - pc[1], qc[2] - pc[2]};
        unit3(d);
        // q in the base frame: B^{-1} (p + d) = R^T (p + d - p0)
        const double w[3] = {pc[0] + d[0] - p0[0], pc[1] + d[1] - p0[1], pc[2] + d[2] - p0[2]};
        const double qb[3] = {ex[0] * w[0] + ex[1] * w[1] + ex[2] * w[2],
                              -(ey[0] * w[0] + ey[1] * w[1] + ey[2] * w[2]),
                              ez[0] * w[0] + ez[1] * w[1] + ez[2] * w[2]};
        // in the previous joint frame: R_prev^T (q - t_prev)
        const double v[3] = {qb[0] - Tp.t[0], qb[1] - Tp.t[1], qb[2] - Tp.t[2]};
        const double qs[3] = {Tp.R[0] * v[0] + Tp.R[3] * v[1] + Tp.R[6] * v[2],
                              Tp.R[1] * v[0] + Tp.R[4] * v[1] + Tp.R[7] * v[2],
                              Tp.R[2] * v[0] + Tp.R[5] * v[1] + Tp.R[8] * v[2]};
        const double *s0 = qs0 + 3 * (i - 1);
        // theta = atan2(-qs0^T [z]x qs, qs0^T [z]x [z]x^T qs)   (graph_revolute.py:308)
        const double num = s0[0] * qs[1] - s0[1] * qs[0];
        const double den = s0[0] * qs[0] + s0[1] * qs[1];
        const double th = atan2(num, den);
        q_out[(size_t)b * n + (i - 1)] = th;
        frame_rotz(Tp, th);
        frame_load(Trel + 16 * (i - 1), rel);
        frame_mul(Tp, rel);
    }
    // final joint from the goal orientation when its offset is along z (graph_revolute.py:312-316)
    if (T_goal && z_aligned && n > 0) {
        const double *Tg = T_goal + (size_t)b * 16;
        // T_th = Tp^{-1} T_goal; only entries (1,0) and (0,0) of its rotation are needed
        const double t00 = Tp.R[0] * Tg[0] + Tp.R[3] * Tg[4] + Tp.R[6] * Tg[8];
        const double t10 = Tp.R[1] * Tg[0] + Tp.R[4] * Tg[4] + Tp.R[7] * Tg[8];
        const double pi = 3.141592653589793;
        double e = q_out[(size_t)b * n + (n - 1)] + atan2(t10, t00) + pi;
        e = e - floor(e / (2.0 * pi)) * (2.0 * pi);   // np.mod(e + pi, 2 pi) - pi
        q_out[(size_t)b * n + (n - 1)] = e - pi;
    }
}

// robot_revolute.py:85-103: T_k = T0[0] * prod_{i<k} exp(S_i q_{i+1}) * T0[k]; the screw of joint i is
// the z axis of T0[i] through its origin, so exp(S_i q) is a rotation about that line.
__global__ void k_fk(int N, int n, int n_anchor, const int32_t *__restrict__ anchor_node,
                     const double *__restrict__ anchor_pos, double axis_length,
                     const double *__restrict__ T0, const double *__restrict__ q, int B,
                     double *__restrict__ T_ee, double *__restrict__ Y)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Frame acc;
    frame_load(T0, acc);
    double *Yb = Y ? Y + (size_t)b * N * 3 : nullptr;
    if (Yb) {
        for (int a = 0; a < n_anchor; ++a) {
            const int node = anchor_node[a];
#pragma unroll
            for (int k = 0; k < 3; ++k) Yb[3 * node + k] = anchor_pos[3 * a + k];
        }
    }
    Frame cur, f0;
    for (int k = 0; k <= n; ++k) {
        frame_load(T0 + 16 * k, f0);
        cur = acc;
        frame_mul(cur, f0);
        if (Yb) {
            const int pn = k == 0 ? 0 : 2 + 2 * k, qn = k == 0 ? 3 : 3 + 2 * k;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Yb[3 * pn + c] = cur.t[c];
                Yb[3 * qn + c] = cur.t[c] + axis_length * cur.R[3 * c + 2];
            }
        }
        if (k == n) break;
        // acc <- acc * exp(S_k q_{k+1}): rotation by theta about axis w through point c0
        const double th = q[(size_t)b * n + k];
        const double w[3] = {f0.R[2], f0.R[5], f0.R[8]};
        double s, c;
        sincos(th, &s, &c);
        Frame e;
        const double v = 1.0 - c;
        e.R[0] = c + v * w[0] * w[0];        e.R[1] = v * w[0] * w[1] - s * w[2]; e.R[2] = v * w[0] * w[2] + s * w[1];
        e.R[3] = v * w[1] * w[0] + s * w[2]; e.R[4] = c + v * w[1] * w[1];        e.R[5] = v * w[1] * w[2] - s * w[0];
        e.R[6] = v * w[2] * w[0] - s * w[1]; e.R[7] = v * w[2] * w[1] + s * w[0]; e.R[8] = c + v * w[2] * w[2];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            e.t[r] = f0.t[r] - (e.R[3 * r] * f0.t[0] + e.R[3 * r + 1] * f0.t[1] + e.R[3 * r + 2] * f0.t[2]);
        frame_mul(acc, e);
    }
    if (T_ee) {
        double *T = T_ee + (size_t)b * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            T[4 * r] = cur.R[3 * r]; T[4 * r + 1] = cur.R[3 * r + 1]; T[4 * r + 2] = cur.R[3 * r + 2];
            T[4 * r + 3] = cur.t[r];
        }
        T[12] = T[13] = T[14] = 0.0;
        T[15] = 1.0;
    }
}

// graph_base.py:219-260 (intended semantics): one thread per problem over the limit edges.
__global__ void k_check_limits(int N, int n_limits, const int32_t *__restrict__ li, const int32_t *__restrict__ lj,
                               const double *__restrict__ lo, const double *__restrict__ up,
                               const double *__restrict__ Y, double tol, int B, int32_t *__restrict__ n_broken,
                               int32_t *__restrict__ status)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *Yb = Y + (size_t)b * N * 3;
    int count = 0;
    for (int k = 0; k < n_limits; ++k) {
        const int i = li[k], j = lj[k];
        const double dx = Yb[3 * i] - Yb[3 * j], dy = Yb[3 * i + 1] - Yb[3 * j + 1], dz = Yb[3 * i + 2] - Yb[3 * j + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        count += d < lo[k] - tol;
        count += d > up[k] + tol;
    }
    n_broken[b] = count;
    if (status && count > 0 && (status[b] == GIK_STATUS_CONVERGED || status[b] == GIK_STATUS_MAXITER))
        status[b] = GIK_STATUS_LIMITS;
}

}  // namespace

extern "C" int gik_check_limits(const GikPlan *p, const double *Y, double tol, int32_t B, int32_t *n_broken,
                                int32_t *status, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p || !Y || !n_broken || B < 0) { gik_set_error("gik_check_limits: bad argument"); return GIK_EINVAL; }
    k_check_limits<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->N, p->n_limits, p->limit_i, p->limit_j,
                                                                    p->limit_lower, p->limit_upper, Y, tol, B,
                                                                    n_broken, status);
    return gik_check_cuda(cudaGetLastError(), "k_check_limits launch");
}

extern "C" int gik_joints(const GikPlan *p, const double *Y, const double *T_goal, int32_t B, double *q, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p) { gik_set_error("gik_joints: null plan"); return GIK_EINVAL; }
    if (p->n_joints <= 0 || !p->T0) { gik_set_error("gik_joints: plan was created without joint tables"); return GIK_EINVAL; }
    if (!Y || !q || B < 0) { gik_set_error("gik_joints: bad argument"); return GIK_EINVAL; }
    if (p->N < 4 + 2 * p->n_joints) { gik_set_error("gik_joints: node layout does not match n_joints"); return GIK_EINVAL; }
    if (B == 0) return GIK_OK;
    k_joints<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->N, p->n_joints, p->T0, p->Trel, p->qs0,
                                                              p->last_joint_z_aligned, Y, T_goal, B, q);
    return gik_check_cuda(cudaGetLastError(), "k_joints launch");
}

extern "C" int gik_fk(const GikPlan *p, const double *q, int32_t B, double *T_ee, double *Y, void *stream)
{
    if (B == 0) return GIK_OK;
    if (!p) { gik_set_error("gik_fk: null plan"); return GIK_EINVAL; }
    if (p->n_joints <= 0 || !p->T0) { gik_set_error("gik_fk: plan was created without joint tables"); return GIK_EINVAL; }
    if (!q || B < 0) { gik_set_error("gik_fk: bad argument"); return GIK_EINVAL; }
    if (B == 0 || (!T_ee && !Y)) return GIK_OK;
    k_fk<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->N, p->n_joints, p->n_anchor, p->anchor_node,
                                                          p->anchor_pos, p->axis_length, p->T0, q, B, T_ee, Y);
    return gik_check_cuda(cudaGetLastError(), "k_fk launch");
}
