// gik_plan.cu -- plan creation / destruction, error reporting.
//
// A plan is the goal-independent, device-resident description of one robot +
// environment: what the reference rebuilds per goal through networkx
// (graph_base.py:171-180, dgp.py:42-65,124-147, graph_base.py:262-279) compiled
// once into node-centric slot tables for the group kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "gik_common.cuh"

static thread_local char g_err[512] = "";

void gik_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gik_check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return GIK_OK;
    gik_set_error("CUDA error %s (%s) in %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
    return GIK_ECUDA;
}

extern "C" const char *gik_last_error(void) { return g_err; }
extern "C" int gik_version(void) { return 100; }

extern "C" int gik_default_opts(GikSolveOpts *o)
{
    if (!o) { gik_set_error("gik_default_opts: null"); return GIK_EINVAL; }
    o->mingradnorm = 0.5 * 1e-9;     // riemannian_solver.py:45
    o->maxiter = 3000;               // :47
    o->theta = 1.0;                  // :48
    o->kappa = 0.1;                  // :49
    o->rho_prime = 0.1;              // trust_region.py:90
    o->rho_regularization = 1e3;     // :92
    o->mininner = 1;                 // :116
    o->maxinner = 10000;             // :118
    o->Delta_bar = 13.0;             // fixed_rank_psd_sym.py:72 (10 + k)
    o->Delta0 = 13.0 / 8;            // trust_region.py:137-138
    o->kernel = GIK_KERNEL_AUTO;
    o->maxtime = 1000.0;             // pymanopt Solver default (trust_region.py:103 passes none)
    return GIK_OK;
}

template <typename T>
static int upload(T **dst, const T *src, size_t count)
{
    *dst = nullptr;
    if (count == 0) return GIK_OK;
    GIK_CUDA(cudaMalloc((void **)dst, count * sizeof(T)));
    GIK_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return GIK_OK;
}

static void inv4(const double *T, double *out)
{
    // rigid transform inverse
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[r * 4 + c] = T[c * 4 + r];
    for (int r = 0; r < 3; ++r)
        out[r * 4 + 3] = -(out[r * 4 + 0] * T[3] + out[r * 4 + 1] * T[7] + out[r * 4 + 2] * T[11]);
    out[12] = out[13] = out[14] = 0.0;
    out[15] = 1.0;
}

static void mul4(const double *A, const double *B, double *C)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c];
            C[r * 4 + c] = s;
        }
}

extern "C" int gik_plan_create(const GikPlanDesc *d, GikPlan **out)
{
    if (!d || !out) { gik_set_error("gik_plan_create: null argument"); return GIK_EINVAL; }
    *out = nullptr;
    const int N = d->n_nodes;
    if (N < 2 || N > 65535) { gik_set_error("gik_plan_create: n_nodes=%d out of range", N); return GIK_EINVAL; }
    if (d->n_terms < 0 || (d->n_terms && (!d->term_i || !d->term_j || !d->term_kind || !d->term_target))) {
        gik_set_error("gik_plan_create: term arrays missing");
        return GIK_EINVAL;
    }
    if (d->n_goal > 4094) { gik_set_error("gik_plan_create: n_goal=%d exceeds 4094", d->n_goal); return GIK_ELIMIT; }

    GikPlan *p = new GikPlan();
    memset(p, 0, sizeof(*p));
    int rc = GIK_OK;
    cudaGetDevice(&p->device);
    cudaDeviceProp prop;
    if ((rc = gik_check_cuda(cudaGetDeviceProperties(&prop, p->device), "cudaGetDeviceProperties"))) { delete p; return rc; }
    p->sm_count = prop.multiProcessorCount;
    p->N = N;
    p->n_terms = d->n_terms;
    p->n_goal = d->n_goal;
    p->n_anchor = d->n_anchor;
    p->goal_p = d->goal_p;
    p->goal_q = d->goal_q;
    p->axis_length = d->axis_length;
    p->n_joints = d->n_joints;
    p->n_goal_edges = d->n_goal_edges;

    // group geometry
    if (N <= 16) { p->W = 16; p->NPL = 1; }
    else if (N <= 32) { p->W = 32; p->NPL = 1; }
    else if (N <= 64) { p->W = 32; p->NPL = 2; }
    else if (N <= 128) { p->W = 32; p->NPL = 4; }
    else if (N <= 256) { p->W = 32; p->NPL = 8; }    // generic kernels only; part of the per-lane state in local memory
    else if (N <= 480) { p->W = 32; p->NPL = 15; }
    else {
        gik_set_error("gik_plan_create: n_nodes=%d exceeds the compiled limit of 480", N);
        delete p;
        return GIK_ELIMIT;
    }
    const int NP = p->W * p->NPL;

    // node-centric slot tables
    std::vector<int32_t> deg(NP, 0);
    for (int t = 0; t < d->n_terms; ++t) {
        const int i = d->term_i[t], j = d->term_j[t];
        if (i < 0 || j < 0 || i >= N || j >= N || i == j || d->term_kind[t] < 0 || d->term_kind[t] > 2) {
            gik_set_error("gik_plan_create: bad term %d (i=%d j=%d kind=%d)", t, i, j, d->term_kind[t]);
            delete p;
            return GIK_EINVAL;
        }
        const int gs = d->term_goal ? d->term_goal[t] : -1;
        if (gs >= d->n_goal) { gik_set_error("gik_plan_create: term %d goal slot %d >= n_goal", t, gs); delete p; return GIK_EINVAL; }
        deg[i]++;
        deg[j]++;
    }
    int maxdeg = 1;
    for (int i = 0; i < N; ++i) maxdeg = deg[i] > maxdeg ? deg[i] : maxdeg;
    p->maxdeg = maxdeg;
    std::vector<uint32_t> info((size_t)maxdeg * N, 0u);
    std::vector<double> target((size_t)maxdeg * N, 0.0);
    std::vector<int32_t> fill(N, 0);
    for (int t = 0; t < d->n_terms; ++t) {
        const int i = d->term_i[t], j = d->term_j[t];
        const int gs = d->term_goal ? d->term_goal[t] : -1;
        const uint32_t kg = ((uint32_t)d->term_kind[t] << 16) | ((uint32_t)(gs + 1) << 20);
        int k = fill[i]++;
        info[(size_t)k * N + i] = (uint32_t)j | kg;
        target[(size_t)k * N + i] = d->term_target[t];
        k = fill[j]++;
        info[(size_t)k * N + j] = (uint32_t)i | kg;
        target[(size_t)k * N + j] = d->term_target[t];
    }

    // lane-centric tables for k_rtr_fast: lane l serves node l / LPN and owns the slots
    // k = l % LPN, l % LPN + LPN, ... of that node; unused entries are inert (kind 3, self)
    std::vector<uint32_t> finfo;
    std::vector<double> ftarget;
    p->fast_LPN = 0;
    p->fast_SPL = 0;
    if (N <= 32) {
        const int LPN = N <= 16 ? 2 : 1;
        const int SPL = (maxdeg + LPN - 1) / LPN;
        if (SPL <= GIK_FAST_ROWS) {
            p->fast_LPN = LPN;
            p->fast_SPL = SPL;
            finfo.assign((size_t)GIK_FAST_ROWS * 32, 0u);
            ftarget.assign((size_t)GIK_FAST_ROWS * 32, 0.0);
            for (int l = 0; l < 32; ++l) {
                const int node = l / LPN, sub = l % LPN;
                for (int s = 0; s < GIK_FAST_ROWS; ++s) {
                    const int k = sub + s * LPN;
                    if (node < N && k < deg[node]) {
                        finfo[(size_t)s * 32 + l] = info[(size_t)k * N + node];
                        ftarget[(size_t)s * 32 + l] = target[(size_t)k * N + node];
                    } else {
                        finfo[(size_t)s * 32 + l] = (uint32_t)(node < N ? node : 0) | (3u << 16);
                    }
                }
            }
        }
    }

    // k_rtr_fast2 (32 < N <= 64): nodes ordered by degree; the 32 highest-degree nodes are the lanes' first
    // nodes, the others the second node of lanes 0 .. N-33.  (S0, S1) = smallest compiled slot counts that fit.
    std::vector<uint32_t> f2info;
    std::vector<double> f2target;
    std::vector<int32_t> f2node;
    p->fast2_S0 = p->fast2_S1 = 0;
    if (N > 32 && N <= 64) {
        std::vector<int> order(N);
        for (int i = 0; i < N; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return deg[x] > deg[y]; });
        int need0 = 0, need1 = 0;
        for (int k = 0; k < N; ++k) {
            int &need = k < 32 ? need0 : need1;
            need = deg[order[k]] > need ? deg[order[k]] : need;
        }
        static const int combos[][2] = {{6, 5}, {7, 5}, {8, 5}, {8, 8}, {12, 8}, {12, 12}};
        for (const auto &c : combos)
            if (c[0] >= need0 && c[1] >= need1) { p->fast2_S0 = c[0]; p->fast2_S1 = c[1]; break; }
        if (p->fast2_S0) {
            const int S0 = p->fast2_S0, ST = S0 + p->fast2_S1;
            f2info.assign((size_t)ST * 32, 0u);
            f2target.assign((size_t)ST * 32, 0.0);
            f2node.assign(64, -1);
            for (int l = 0; l < 32; ++l) {
                for (int m = 0; m < 2; ++m) {
                    const int k0 = m * 32 + l;
                    const int node = k0 < N ? order[k0] : -1;
                    f2node[m * 32 + l] = node;
                    const int rows = m == 0 ? S0 : p->fast2_S1;
                    for (int s = 0; s < rows; ++s) {
                        const size_t at = (size_t)(m * S0 + s) * 32 + l;
                        if (node >= 0 && s < deg[node]) {
                            f2info[at] = info[(size_t)s * N + node];
                            f2target[at] = target[(size_t)s * N + node];
                        } else {
                            f2info[at] = (uint32_t)(node >= 0 ? node : 63) | (3u << 16);   // inert: self, no term
                        }
                    }
                }
            }
        }
    }

    std::vector<uint32_t> dinfo;
    std::vector<double> dtarget;
    if (N <= 16 && maxdeg <= GIK_FAST_ROWS) {
        dinfo.assign((size_t)GIK_FAST_ROWS * 16, 0u);
        dtarget.assign((size_t)GIK_FAST_ROWS * 16, 0.0);
        for (int node = 0; node < 16; ++node)
            for (int s = 0; s < GIK_FAST_ROWS; ++s) {
                if (node < N && s < deg[node]) {
                    dinfo[(size_t)s * 16 + node] = info[(size_t)s * N + node];
                    dtarget[(size_t)s * 16 + node] = target[(size_t)s * N + node];
                } else {
                    dinfo[(size_t)s * 16 + node] = (uint32_t)(node < N ? node : 0) | (3u << 16);
                }
            }
    }

    // dense pair tables for k_rtr_cta
    std::vector<double> dense_t;
    std::vector<unsigned char> dense_k;
    std::vector<int32_t> dgi, dgj, dgs;
    std::vector<unsigned char> hub_k;
    std::vector<double> hub_t;
    p->dense_hub = -1;
    if (N > 32 && N <= 128) {
        dense_t.assign((size_t)N * N, 0.0);
        dense_k.assign((size_t)N * N, 3);
        // A pair may carry a SECOND term (never goal dependent) when all such pairs share one node, the hub:
        // with obstacle_semantics="intended" the pairs (p_n, obstacle) hold the goal's exact distance and the
        // obstacle hinge.  The second terms live in a row of per-partner tables (hub_kind / hub_target).
        bool representable = true;
        int hub_a = -1, hub_b = -1;   // remaining hub candidates
        hub_k.assign(N, 3);
        hub_t.assign(N, 0.0);
        std::vector<int> second_i, second_j, second_t;
        for (int t = 0; t < d->n_terms && representable; ++t) {
            const int i = d->term_i[t], j = d->term_j[t];
            const int gs = d->term_goal ? d->term_goal[t] : -1;
            if (dense_k[(size_t)i * N + j] != 3) {           // second (or later) term on this pair
                if (gs >= 0) { representable = false; break; }
                if (hub_a < 0) { hub_a = i; hub_b = j; }
                else {
                    const bool a_ok = hub_a == i || hub_a == j, b_ok = hub_b >= 0 && (hub_b == i || hub_b == j);
                    if (!a_ok && !b_ok) { representable = false; break; }
                    if (!a_ok) { hub_a = hub_b; }
                    if (!(a_ok && b_ok)) hub_b = -1;
                }
                second_i.push_back(i); second_j.push_back(j); second_t.push_back(t);
                continue;
            }
            dense_k[(size_t)i * N + j] = dense_k[(size_t)j * N + i] = (unsigned char)d->term_kind[t];
            dense_t[(size_t)i * N + j] = dense_t[(size_t)j * N + i] = d->term_target[t];
            if (gs >= 0) { dgi.push_back(i); dgj.push_back(j); dgs.push_back(gs); }
        }
        p->dense_hub = -1;
        if (representable && !second_t.empty()) {
            p->dense_hub = hub_a;
            for (size_t k = 0; k < second_t.size() && representable; ++k) {
                const int partner = second_i[k] == hub_a ? second_j[k] : second_i[k];
                if ((second_i[k] != hub_a && second_j[k] != hub_a) || hub_k[partner] != 3) { representable = false; break; }  // third term
                hub_k[partner] = (unsigned char)d->term_kind[second_t[k]];
                hub_t[partner] = d->term_target[second_t[k]];
            }
        }
        if (!representable) { dense_t.clear(); dense_k.clear(); dgi.clear(); dgj.clear(); dgs.clear(); p->dense_hub = -1; }
        if (p->dense_hub < 0) { hub_k.clear(); hub_t.clear(); }
    }
    p->n_dense_goal = (int)dgi.size();
    // Equality clique: the largest set of nodes (greedy) whose pairs ALL carry a static equality term -- the anchors
    // (base nodes and, with the reference's obstacle semantics, every obstacle centre): 106 of the 118 nodes of
    // KUKA + table, 5565 of its 6903 pairs.  For those pairs the <D,w> D part of the Hessian-vector product factors
    // through 15 moments of the direction (gik_rtr_cta.cu), so the dense kernel works in an order with the clique last.
    std::vector<int32_t> dperm;          // position -> node
    p->dense_clique_start = N;
    if (!dense_k.empty()) {
        std::vector<char> in(N, 1), goalpair((size_t)N * N, 0);
        for (size_t k = 0; k < dgi.size(); ++k) goalpair[(size_t)dgi[k] * N + dgj[k]] = goalpair[(size_t)dgj[k] * N + dgi[k]] = 1;
        auto bad = [&](int i, int j) {
            return dense_k[(size_t)i * N + j] != GIK_TERM_EQ || goalpair[(size_t)i * N + j] ||
                   (p->dense_hub >= 0 && (i == p->dense_hub || j == p->dense_hub));
        };
        for (;;) {
            int worst = -1, worst_bad = 0;
            for (int i = 0; i < N; ++i) {
                if (!in[i]) continue;
                int nb = 0;
                for (int j = 0; j < N; ++j) nb += (j != i && in[j] && bad(i, j));
                if (nb > worst_bad) { worst_bad = nb; worst = i; }
            }
            if (worst < 0) break;
            in[worst] = 0;
        }
        int nC = 0;
        for (int i = 0; i < N; ++i) nC += in[i];
        if (nC < 32) std::fill(in.begin(), in.end(), 0), nC = 0;    // not worth the extra reduction
        for (int i = 0; i < N; ++i) if (!in[i]) dperm.push_back(i);
        for (int i = 0; i < N; ++i) if (in[i]) dperm.push_back(i);
        p->dense_clique_start = N - nC;
        std::vector<int32_t> pos(N);
        for (int q = 0; q < N; ++q) pos[dperm[q]] = q;
        std::vector<double> t2((size_t)N * N);
        std::vector<unsigned char> k2((size_t)N * N);
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) {
                t2[(size_t)a * N + b] = dense_t[(size_t)dperm[a] * N + dperm[b]];
                k2[(size_t)a * N + b] = dense_k[(size_t)dperm[a] * N + dperm[b]];
            }
        dense_t.swap(t2);
        dense_k.swap(k2);
        for (size_t k = 0; k < dgi.size(); ++k) { dgi[k] = pos[dgi[k]]; dgj[k] = pos[dgj[k]]; }
        if (p->dense_hub >= 0) {
            std::vector<unsigned char> hk(N);
            std::vector<double> ht(N);
            for (int a = 0; a < N; ++a) { hk[a] = hub_k[dperm[a]]; ht[a] = hub_t[dperm[a]]; }
            hub_k.swap(hk);
            hub_t.swap(ht);
            p->dense_hub = pos[p->dense_hub];
        }
        p->dense_goal_p = d->goal_p >= 0 ? pos[d->goal_p] : -1;
        p->dense_goal_q = d->goal_q >= 0 ? pos[d->goal_q] : -1;
    }

    // omega edge list for the initialisation's linear projection
    std::vector<int32_t> oi, oj;
    if (d->omega)
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j)
                if (d->omega[(size_t)i * N + j]) { oi.push_back(i); oj.push_back(j); }
    p->n_omega_edges = (int)oi.size();
    std::vector<int32_t> optr(N + 1, 0), oadj;
    if (d->omega)
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j)
                if (j != i && (d->omega[(size_t)i * N + j] || d->omega[(size_t)j * N + i])) oadj.push_back(j);
            optr[i + 1] = (int32_t)oadj.size();
        }

    // joint-recovery tables (graph_revolute.py:283-310)
    std::vector<double> Trel, qs0;
    if (d->n_joints > 0 && d->T0) {
        const int n = d->n_joints;
        Trel.resize((size_t)n * 16);
        qs0.resize((size_t)n * 3);
        for (int i = 1; i <= n; ++i) {
            double inv[16], Tq[16], rel[16];
            inv4(d->T0 + (size_t)(i - 1) * 16, inv);
            mul4(inv, d->T0 + (size_t)i * 16, &Trel[(size_t)(i - 1) * 16]);
            memcpy(Tq, d->T0 + (size_t)i * 16, sizeof(Tq));
            for (int r = 0; r < 3; ++r) Tq[r * 4 + 3] += d->axis_length * Tq[r * 4 + 2];
            mul4(inv, Tq, rel);
            for (int r = 0; r < 3; ++r) qs0[(size_t)(i - 1) * 3 + r] = rel[r * 4 + 3];
        }
        // final-joint correction applies when the last offset is along z (graph_revolute.py:314)
        const double *tr = &Trel[(size_t)(n - 1) * 16];
        const double cx = tr[7] * 1.0 - 0.0, cy = 0.0 - tr[3] * 1.0;  // cross(t, e_z) = (t_y, -t_x, 0)
        p->last_joint_z_aligned = sqrt(cx * cx + cy * cy) < 1e-10;
    }

    // k_bounds_init works on three N x N fp64 matrices per goal.  It is latency bound, so what counts is the number of
    // goals in flight per SM: the third matrix (written and read once or twice per goal) moves to an L2-resident
    // workspace of the caller's when that admits more CTAs per SM (mode 1), and all three do when not even two fit (mode 2)
    {
        const size_t small = (size_t)gik_bi_small_doubles(N) * sizeof(double);
        const size_t mat = (size_t)N * N * sizeof(double);
        const size_t cap = 227 * 1024, per_cta = 1024;
        const int regs = gik_bi_reg_ctas(N);
        int occ0 = (int)((cap + per_cta) / (small + 3 * mat + per_cta)), occ1 = (int)((cap + per_cta) / (small + 2 * mat + per_cta));
        if (small + 3 * mat > cap) occ0 = 0;
        if (small + 2 * mat > cap) occ1 = 0;
        if (occ0 > regs) occ0 = regs;
        if (occ1 > regs) occ1 = regs;
        p->bi_mode = (occ0 >= 1 && occ0 >= occ1) ? 0 : (occ1 >= 1 ? 1 : 2);
        p->bi_blocks = p->bi_mode == 1 ? p->sm_count * occ1 : p->sm_count * (regs < 2 ? regs : 2);
    }

    bool ok = true;
    ok = ok && !upload(&p->slot_info, info.data(), info.size());
    ok = ok && !upload(&p->slot_target, target.data(), target.size());
    ok = ok && !upload(&p->deg, deg.data(), deg.size());
    ok = ok && !upload(&p->fast_info, finfo.data(), finfo.size());
    ok = ok && !upload(&p->fast_target, ftarget.data(), ftarget.size());
    ok = ok && !upload(&p->fast2_info, f2info.data(), f2info.size());
    ok = ok && !upload(&p->fast2_target, f2target.data(), f2target.size());
    ok = ok && !upload(&p->fast2_node, f2node.data(), f2node.size());
    ok = ok && !upload(&p->duo_info, dinfo.data(), dinfo.size());
    ok = ok && !upload(&p->duo_target, dtarget.data(), dtarget.size());
    ok = ok && !upload(&p->dense_target, dense_t.data(), dense_t.size());
    ok = ok && !upload(&p->dense_kind, dense_k.data(), dense_k.size());
    ok = ok && !upload(&p->dense_hub_kind, hub_k.data(), hub_k.size());
    ok = ok && !upload(&p->dense_hub_target, hub_t.data(), hub_t.size());
    ok = ok && !upload(&p->dense_goal_i, dgi.data(), dgi.size());
    ok = ok && !upload(&p->dense_goal_j, dgj.data(), dgj.size());
    ok = ok && !upload(&p->dense_goal_slot, dgs.data(), dgs.size());
    ok = ok && !upload(&p->dense_perm, dperm.data(), dperm.size());
    ok = ok && !upload(&p->anchor_node, d->anchor_node, (size_t)d->n_anchor);
    ok = ok && !upload(&p->anchor_pos, d->anchor_pos, (size_t)d->n_anchor * 3);
    if (d->bs_lower && d->bs_upper) {
        ok = ok && !upload(&p->bs_lower, d->bs_lower, (size_t)N * N);
        ok = ok && !upload(&p->bs_upper, d->bs_upper, (size_t)N * N);
        // column lists of the positive lower bounds for the first max-plus product of k_bounds_init; the pairs with
        // p_n / q_n are patched per goal and handled from two dense rows there
        const int gp = d->n_goal_edges > 0 ? d->goal_p : -1, gq = d->n_goal_edges > 0 ? d->goal_q : -1;
        std::vector<int32_t> lptr(N + 1, 0), lrow;
        std::vector<double> lval;
        for (int bcol = 0; bcol < N; ++bcol) {
            for (int arow = 0; arow < N; ++arow) {
                const double l = d->bs_lower[(size_t)arow * N + bcol];
                if (l > 0.0 && arow != gp && arow != gq && bcol != gp && bcol != gq) { lrow.push_back(arow); lval.push_back(l); }
            }
            lptr[bcol + 1] = (int32_t)lrow.size();
        }
        if (lrow.empty()) { lrow.push_back(0); lval.push_back(0.0); }   // upload() of an empty array leaves a null pointer
        ok = ok && !upload(&p->low_ptr, lptr.data(), lptr.size());
        ok = ok && !upload(&p->low_row, lrow.data(), lrow.size());
        ok = ok && !upload(&p->low_val, lval.data(), lval.size());
    }
    ok = ok && !upload(&p->goal_edge_i, d->goal_edge_i, (size_t)d->n_goal_edges);
    ok = ok && !upload(&p->goal_edge_j, d->goal_edge_j, (size_t)d->n_goal_edges);
    ok = ok && !upload(&p->goal_edge_slot, d->goal_edge_slot, (size_t)d->n_goal_edges);
    ok = ok && !upload(&p->omega_i, oi.data(), oi.size());
    ok = ok && !upload(&p->omega_j, oj.data(), oj.size());
    ok = ok && !upload(&p->omega_ptr, optr.data(), optr.size());
    ok = ok && !upload(&p->omega_adj, oadj.data(), oadj.size());
    p->n_limits = (d->n_limits > 0 && d->limit_i && d->limit_j && d->limit_lower && d->limit_upper) ? d->n_limits : 0;
    for (int k = 0; k < p->n_limits; ++k)
        if (d->limit_i[k] < 0 || d->limit_i[k] >= N || d->limit_j[k] < 0 || d->limit_j[k] >= N) {
            gik_set_error("gik_plan_create: bad limit edge %d", k);
            gik_plan_destroy(p);
            return GIK_EINVAL;
        }
    ok = ok && !upload(&p->limit_i, d->limit_i, (size_t)p->n_limits);
    ok = ok && !upload(&p->limit_j, d->limit_j, (size_t)p->n_limits);
    ok = ok && !upload(&p->limit_lower, d->limit_lower, (size_t)p->n_limits);
    ok = ok && !upload(&p->limit_upper, d->limit_upper, (size_t)p->n_limits);
    if (d->n_joints > 0 && d->T0) {
        ok = ok && !upload(&p->T0, d->T0, (size_t)(d->n_joints + 1) * 16);
        ok = ok && !upload(&p->Trel, Trel.data(), Trel.size());
        ok = ok && !upload(&p->qs0, qs0.data(), qs0.size());
    }
    if (!ok) { gik_plan_destroy(p); return GIK_ECUDA; }
    *out = p;
    return GIK_OK;
}

extern "C" int gik_plan_destroy(GikPlan *p)
{
    if (!p) return GIK_OK;
    void *ptrs[] = {p->slot_info, p->slot_target, p->deg, p->fast_info, p->fast_target, p->fast2_info, p->fast2_target, p->fast2_node, p->duo_info, p->duo_target, p->dense_target, p->dense_kind, p->dense_hub_kind, p->dense_hub_target, p->dense_goal_i, p->dense_goal_j, p->dense_perm,
                    p->dense_goal_slot, p->anchor_node, p->anchor_pos, p->bs_lower,
                    p->bs_upper, p->low_ptr, p->low_row, p->low_val, p->goal_edge_i, p->goal_edge_j, p->goal_edge_slot, p->omega_ptr, p->omega_adj, p->omega_i,
                    p->omega_j, p->T0, p->Trel, p->qs0, p->limit_i, p->limit_j, p->limit_lower, p->limit_upper};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    delete p;
    return GIK_OK;
}

extern "C" int gik_plan_info(const GikPlan *p, int32_t what[8])
{
    if (!p || !what) { gik_set_error("gik_plan_info: null argument"); return GIK_EINVAL; }
    what[0] = p->N; what[1] = p->n_terms; what[2] = p->n_goal; what[3] = p->maxdeg;
    what[4] = p->n_joints; what[5] = p->W; what[6] = p->NPL; what[7] = p->sm_count;
    return GIK_OK;
}
