/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of the cluster-matching arithmetic and of the
 * 3-NN scale initialisation.  Loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs only.
 *
 * What is restated and how it is pinned:
 *  - oracle_nn_match: `torch.argmin(torch.cdist(a, b), -1)` / `torch.min(cdist, 1)`
 *    (notebooks/10.visualize_and_fit_patch_to_multiple.ipynb cell 34,
 *    notebooks/29.2.Modify_style_clusters.ipynb cell 58).  torch.cdist takes its matmul path
 *    for these sizes: x1_ = [-2a, |a|^2, 1], x2_ = [b, 1, |b|^2], result = x1_ @ x2_^T,
 *    clamp_min(0), sqrt.  On this image's CPU build (torch 2.11, oneMKL 2024.2) that product is
 *    bit-identical to a sequential fused-multiply-add over the 5 terms and |x|^2 is (x0^2+x1^2)+x2^2;
 *    tests/test_match_oracle.py checks this against torch itself, so this part is PINNED.
 *    Only the final sqrt differs: torch's vectorised CPU sqrt is not correctly rounded
 *    (about 0.7% of values differ by 1 ulp from IEEE sqrt), which can only move assignments
 *    between candidates whose squared distances are within a few ulp; the test asserts exactly that.
 *  - oracle_w2_cost / oracle_w2_match: closed-form squared 2-Wasserstein distance between
 *    Gaussians.  The reference contains NO implementation of it, so
 *    PARITY IS UNPINNED for this term: the oracle is the textbook formula
 *        W2^2 = |m1-m2|^2 + tr S1 + tr S2 - 2 tr((S1^1/2 S2 S1^1/2)^1/2)
 *    evaluated in fp32 in the fixed operation order below, cross-checked in the tests against a
 *    float64 scipy.linalg.sqrtm evaluation.  The CUDA kernel executes the same operations
 *    (explicit fma / IEEE div / IEEE sqrt), so assignments and costs are compared bit for bit.
 *  - oracle_knn: brute-force statement of SimpleKNN::knn's result
 *    (submodules/simple-knn/simple_knn.cu:131-145 updateKBest, :147-183 boxMeanDist): mean of the
 *    three smallest d^2 to other points, d^2 = fma(dz,dz,fma(dx,dx,dy*dy)) (the contraction nvcc
 *    picks for the reference; verified bit for bit against its sm_100 build), self excluded by index, FLT_MAX for missing neighbours.  Pinned on the GPU box against
 *    the real simple-knn (oracle/_ref) by tests/test_knn_gpu.py.
 *  - oracle_cluster_stats: per-cluster mean and covariance of member xyz (the statistics the
 *    notebooks derive from K-Means memberships); float64 accumulation, population covariance.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- M3: torch.cdist order -------------------------------------------------------------- */
static float norm2_torch(const float* x) { return (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]; }

float oracle_cdist_sq(const float* a, const float* b) {
    const float na = norm2_torch(a), nb = norm2_torch(b);
    float acc = 0.f;
    acc = fmaf(-2.f * a[0], b[0], acc);
    acc = fmaf(-2.f * a[1], b[1], acc);
    acc = fmaf(-2.f * a[2], b[2], acc);
    acc = fmaf(na, 1.f, acc);
    acc = fmaf(1.f, nb, acc);
    return fmaxf(acc, 0.f);
}

void oracle_cdist(int Na, int Nb, const float* a, const float* b, float* out, int take_sqrt) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Na; ++i)
        for (int j = 0; j < Nb; ++j) {
            float v = oracle_cdist_sq(a + 3 * i, b + 3 * j);
            out[(size_t)i * Nb + j] = take_sqrt ? sqrtf(v) : v;
        }
}

/* argmin over j of sqrt(cdist_sq) with ties to the lowest j; out_dist = the distance */
void oracle_nn_match(int Na, int Nb, const float* a, const float* b, int32_t* out_idx, float* out_dist) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Na; ++i) {
        float best = INFINITY;
        int bi = 0;
        for (int j = 0; j < Nb; ++j) {
            const float d = sqrtf(oracle_cdist_sq(a + 3 * i, b + 3 * j));
            if (d < best) { best = d; bi = j; }
        }
        out_idx[i] = Nb > 0 ? bi : -1;
        out_dist[i] = best;
    }
}

/* ---- M2: k smallest per row of cdist, ordered by (distance, index) == stable sort of the row
 * (torch.sort / torch.topk call sites: aux_optimize_cluster_D_W_distance.py:79-82,
 * aux_optimize_cluster_D_W_distance2.py:269-273, notebooks/25.4 cell 73).  PINNED against
 * torch.sort(torch.cdist(...), stable=True) on the squared-distance values in tests/test_match_oracle.py
 * (same 1-ulp sqrt caveat as oracle_nn_match). */
void oracle_cdist_topk(int Na, int Nb, const float* a, const float* b, int k, float* out_dist, int32_t* out_idx) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < Na; ++i) {
        float* bd = out_dist + (size_t)i * k;
        int32_t* bi = out_idx + (size_t)i * k;
        int filled = 0;
        for (int j = 0; j < Nb; ++j) {
            const float d = sqrtf(oracle_cdist_sq(a + 3 * i, b + 3 * j));
            if (filled == k && !(d < bd[k - 1])) continue; /* later index never displaces an equal distance */
            int pos = filled < k ? filled : k - 1;
            while (pos > 0 && d < bd[pos - 1]) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = j;
            if (filled < k) ++filled;
        }
    }
}

/* ---- M4: ot.emd2 with uniform weights (aux_optimize_cluster_D_W_distance.py:260-270) ------------
 * POT (`import ot`) is an un-vendored, un-pinned dependency of the reference and is not installed
 * here, so PARITY IS UNPINNED against POT itself.  Restated from its published definition:
 * emd2(a, b, M) = min_{G >= 0, G 1 = a, G^T 1 = b} <G, M>, M = ot.dist(xa, xb) = squared Euclidean
 * (a2 + b2 - 2 a.b^T clamped at 0, ot/utils.py euclidean_distances).  For a = b = 1/n the vertices of
 * the feasible polytope are permutation matrices / n, so the optimum is a linear assignment:
 * solved by the Jonker-Volgenant shortest augmenting path method in double precision.  The optimal
 * VALUE is unique; tests cross-check it against scipy.optimize.linear_sum_assignment. */
static void emd_cost_matrix(int n, const float* xa, const float* xb, float* M) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const float* a = xa + 3 * i;
            const float* b = xb + 3 * j;
            const float a2 = (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2];
            const float b2 = (b[0] * b[0] + b[1] * b[1]) + b[2] * b[2];
            const float dot = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
            float c = -2.f * dot;
            c = c + a2;
            c = c + b2;
            M[(size_t)i * n + j] = fmaxf(c, 0.f);
        }
}

float oracle_emd2_uniform(int n, const float* xa, const float* xb, int32_t* perm, float* M_out) {
    float* M = (float*)malloc(sizeof(float) * (size_t)n * n);
    emd_cost_matrix(n, xa, xb, M);
    double* u = (double*)calloc((size_t)n + 1, sizeof(double));
    double* v = (double*)calloc((size_t)n + 1, sizeof(double));
    double* minv = (double*)malloc(sizeof(double) * ((size_t)n + 1));
    int* p = (int*)calloc((size_t)n + 1, sizeof(int));
    int* way = (int*)calloc((size_t)n + 1, sizeof(int));
    char* used = (char*)malloc((size_t)n + 1);
    for (int i = 1; i <= n; ++i) {
        p[0] = i;
        int j0 = 0;
        for (int j = 0; j <= n; ++j) { minv[j] = DBL_MAX; used[j] = 0; }
        do {
            used[j0] = 1;
            const int i0 = p[j0];
            double delta = DBL_MAX;
            int j1 = 0;
            for (int j = 1; j <= n; ++j)
                if (!used[j]) {
                    const double cur = (double)M[(size_t)(i0 - 1) * n + (j - 1)] - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < delta) { delta = minv[j]; j1 = j; }
                }
            for (int j = 0; j <= n; ++j)
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            j0 = j1;
        } while (p[j0] != 0);
        do {
            const int j1 = way[j0];
            p[j0] = p[j1];
            j0 = j1;
        } while (j0);
    }
    double tot = 0.0;
    for (int j = 1; j <= n; ++j) {
        tot += (double)M[(size_t)(p[j] - 1) * n + (j - 1)];
        if (perm) perm[p[j] - 1] = j - 1;
    }
    if (M_out) memcpy(M_out, M, sizeof(float) * (size_t)n * n);
    free(M); free(u); free(v); free(minv); free(p); free(way); free(used);
    return (float)(tot / (double)n);
}

/* ---- M5: closed-form Gaussian W2^2, fixed fp32 operation order --------------------------
 * cov6 = (xx, xy, xz, yy, yz, zz).  Per cluster "descriptor" (derived once):
 *   tr = (xx + yy) + zz
 *   adj = adjugate(S) (6 entries), det = xx*adj_xx + xy*adj_xy + xz*adj_xz (clamped at 0)
 * Per pair, with <A,B> = Axx Bxx + Ayy Byy + Azz Bzz + 2 (Axy Bxy + Axz Bxz + Ayz Byz):
 *   c2 = <S1,S2>           = tr(S1 S2)            = sum mu_k^2      (mu_k^2 = eig(S1 S2) >= 0)
 *   c1 = <adj S1, adj S2>  = tr(adj(S1 S2))       = sum_{k<l} mu_k^2 mu_l^2
 *   e3 = sqrt(det1 det2)                           = mu_1 mu_2 mu_3
 *   s  = mu_1 + mu_2 + mu_3 solves s = sqrt(c2 + 2 sqrt(c1 + 2 e3 s)); the map is a contraction
 *        with factor <= 2/9, iterated W2_ITERS times from s0 = sqrt(c2) (exact when det = 0).
 *   W2^2 = max(0, (|m1-m2|^2 + (tr1 + tr2)) - 2 s)
 */
#define W2_ITERS 10

static float dotsym(const float* A, const float* B) {
    float d = A[0] * B[0];
    d = fmaf(A[3], B[3], d);
    d = fmaf(A[5], B[5], d);
    float o = A[1] * B[1];
    o = fmaf(A[2], B[2], o);
    o = fmaf(A[4], B[4], o);
    return fmaf(2.f, o, d);
}

/* desc[16]: 0-2 mean, 3-8 cov, 9-14 adj, 15 det (>=0); returns trace */
void oracle_w2_descriptor(const float* mean, const float* c, float* desc) {
    memcpy(desc, mean, 3 * sizeof(float));
    memcpy(desc + 3, c, 6 * sizeof(float));
    float* adj = desc + 9;
    adj[0] = fmaf(c[3], c[5], -(c[4] * c[4])); /* yy zz - yz^2 */
    adj[1] = fmaf(c[2], c[4], -(c[1] * c[5])); /* xz yz - xy zz */
    adj[2] = fmaf(c[1], c[4], -(c[2] * c[3])); /* xy yz - xz yy */
    adj[3] = fmaf(c[0], c[5], -(c[2] * c[2])); /* xx zz - xz^2 */
    adj[4] = fmaf(c[1], c[2], -(c[0] * c[4])); /* xy xz - xx yz */
    adj[5] = fmaf(c[0], c[3], -(c[1] * c[1])); /* xx yy - xy^2 */
    float det = c[0] * adj[0];
    det = fmaf(c[1], adj[1], det);
    det = fmaf(c[2], adj[2], det);
    desc[15] = fmaxf(det, 0.f);
}

float oracle_w2_cost_desc(const float* d1, const float* d2) {
    const float dx = d1[0] - d2[0], dy = d1[1] - d2[1], dz = d1[2] - d2[2];
    const float dist2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float tr1 = (d1[3] + d1[6]) + d1[8], tr2 = (d2[3] + d2[6]) + d2[8];
    const float c2 = fmaxf(dotsym(d1 + 3, d2 + 3), 0.f);
    const float c1 = fmaxf(dotsym(d1 + 9, d2 + 9), 0.f);
    const float e3 = sqrtf(d1[15] * d2[15]);
    float s = sqrtf(c2);
    for (int it = 0; it < W2_ITERS; ++it) {
        const float e2 = sqrtf(fmaf(2.f * e3, s, c1));
        s = sqrtf(fmaf(2.f, e2, c2));
    }
    const float w = fmaf(-2.f, s, dist2 + (tr1 + tr2));
    return fmaxf(w, 0.f);
}

float oracle_w2_cost(const float* m1, const float* c1, const float* m2, const float* c2) {
    float d1[16], d2[16];
    oracle_w2_descriptor(m1, c1, d1);
    oracle_w2_descriptor(m2, c2, d2);
    return oracle_w2_cost_desc(d1, d2);
}

/* argmin_j W2^2(i, j), ties to the lowest j.  cost_matrix (optional) receives all Kc*Ks costs. */
void oracle_w2_match(int Kc, int Ks, const float* mean_c, const float* cov_c, const float* mean_s,
                     const float* cov_s, int32_t* out_idx, float* out_cost, float* cost_matrix) {
    float* dc = (float*)malloc(sizeof(float) * 16 * (size_t)(Kc > 0 ? Kc : 1));
    float* ds = (float*)malloc(sizeof(float) * 16 * (size_t)(Ks > 0 ? Ks : 1));
    for (int i = 0; i < Kc; ++i) oracle_w2_descriptor(mean_c + 3 * i, cov_c + 6 * i, dc + 16 * i);
    for (int j = 0; j < Ks; ++j) oracle_w2_descriptor(mean_s + 3 * j, cov_s + 6 * j, ds + 16 * j);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Kc; ++i) {
        float best = INFINITY;
        int bi = -1;
        for (int j = 0; j < Ks; ++j) {
            const float w = oracle_w2_cost_desc(dc + 16 * i, ds + 16 * j);
            if (cost_matrix) cost_matrix[(size_t)i * Ks + j] = w;
            if (w < best) { best = w; bi = j; }
        }
        out_idx[i] = bi;
        out_cost[i] = best;
    }
    free(dc);
    free(ds);
}

/* per-cluster mean / population covariance (float64 accumulation, two passes) */
void oracle_cluster_stats(int n, int K, const float* pts, const int32_t* labels, float* mean,
                          float* cov6, int32_t* count) {
    double* sum = (double*)calloc((size_t)K * 3, sizeof(double));
    double* acc = (double*)calloc((size_t)K * 6, sizeof(double));
    memset(count, 0, sizeof(int32_t) * (size_t)K);
    for (int i = 0; i < n; ++i) {
        const int k = labels[i];
        if (k < 0 || k >= K) continue;
        count[k]++;
        for (int c = 0; c < 3; ++c) sum[3 * k + c] += pts[3 * i + c];
    }
    for (int k = 0; k < K; ++k)
        for (int c = 0; c < 3; ++c) sum[3 * k + c] = count[k] ? sum[3 * k + c] / count[k] : 0.0;
    for (int i = 0; i < n; ++i) {
        const int k = labels[i];
        if (k < 0 || k >= K) continue;
        const double dx = pts[3 * i] - sum[3 * k], dy = pts[3 * i + 1] - sum[3 * k + 1],
                     dz = pts[3 * i + 2] - sum[3 * k + 2];
        double* a = acc + 6 * k;
        a[0] += dx * dx; a[1] += dx * dy; a[2] += dx * dz; a[3] += dy * dy; a[4] += dy * dz; a[5] += dz * dz;
    }
    for (int k = 0; k < K; ++k) {
        for (int c = 0; c < 3; ++c) mean[3 * k + c] = (float)sum[3 * k + c];
        for (int c = 0; c < 6; ++c) cov6[6 * k + c] = count[k] ? (float)(acc[6 * k + c] / count[k]) : 0.f;
    }
    free(sum);
    free(acc);
}

/* ---- N0: brute-force 3-NN ---------------------------------------------------------------- */
void oracle_knn(int P, const float* pts, float* mean_dist2, int32_t* nn_idx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float bd[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        int32_t bi[3] = {-1, -1, -1};
        const float* q = pts + 3 * i;
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            const float dx = pts[3 * j] - q[0], dy = pts[3 * j + 1] - q[1], dz = pts[3 * j + 2] - q[2];
            float dist = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            int32_t id = j;
            for (int k = 0; k < 3; ++k) { /* (distance, index) lexicographic: ties keep the lowest index */
                if (dist < bd[k] || (dist == bd[k] && (uint32_t)id < (uint32_t)bi[k])) {
                    float td = bd[k]; int32_t ti = bi[k];
                    bd[k] = dist; bi[k] = id; dist = td; id = ti;
                }
            }
        }
        mean_dist2[i] = (bd[0] + bd[1] + bd[2]) / 3.0f;
        if (nn_idx) { nn_idx[3 * i] = bi[0]; nn_idx[3 * i + 1] = bi[1]; nn_idx[3 * i + 2] = bi[2]; }
    }
}
