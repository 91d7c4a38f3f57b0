/*
 * oracle.c -- CPU restatement of the pointnet-autoencoder reconstruction-loss ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / baseline.
 *
 * Every function restates one reference routine in scalar C.  Where the
 * reference GPU kernel and the reference CPU loop disagree
 * the GPU kernel is the one restated, because that is the parity target:
 * float accumulators, the FMA contractions nvcc emits for the reference source
 * (checked in the sm_100a SASS: d = fma(dz,dz,fma(dx,dx,dy*dy)); sweep sums
 * accumulate with fma), 10 levels j=7..-2 and the (b,m,n) match layout.
 *
 * The one thing that cannot be restated bit-for-bit on a CPU is MUFU.EX2 /
 * MUFU.RSQ (hardware approximations, <=2 ulp).  They are replaced by correctly
 * rounded exp2 / 1/sqrt, so EMD results agree with the reference GPU kernels
 * to ~1e-6 relative, not bit-exactly.  Chamfer (no transcendental) is bit-exact.
 *
 * Parity pinning: the reference ships no golden vectors.
 * This restatement is pinned (tests/test_oracle_cpu.py) against the reference's
 * own CPU loops compiled from /root/reference into oracle/_ref/libref_cpu.so, and
 * (tests/test_gpu_parity.py::TestAgainstReferenceKernels, on the GPU box) against the reference's own CUDA
 * kernels compiled unmodified into oracle/_ref/libref_gpu.so, plus the golden
 * fixtures under tests/golden/ that were generated from those kernels.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared oracle.c -o liboracle.so -lm
 * (-ffp-contract=off: every fused operation below is an explicit fmaf()).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Chamfer forward.                                                           */
/* ------------------------------------------------------------------------- */

/* One direction: for every point of `a` the nearest point of `c`.
 * Follows tf_nndistance_g.cu:5-127 (GPU) and tf_nndistance.cpp:21-43 (CPU):
 * squared distance, strict '<' so the lowest index wins ties.
 * contract=1: d = fma(dz,dz,fma(dx,dx,dy*dy))  -- what nvcc emits for
 *             `x2*x2+y2*y2+z2*z2` (tf_nndistance_g.cu:28), the GPU parity target.
 * contract=0: d = (dx*dx+dy*dy)+dz*dz, each op rounded -- the reference CPU loop
 *             (tf_nndistance.cpp:33; the `double d` there is assigned from a
 *             float expression, so the sum is formed in float).
 */
static void nn_one_direction(int b, int n, int m, const float *a, const float *c,
                             float *dist, int *idx, int contract)
{
    for (int i = 0; i < b; i++) {
        const float *pa = a + (size_t)i * n * 3;
        const float *pc = c + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            float x1 = pa[j * 3 + 0], y1 = pa[j * 3 + 1], z1 = pa[j * 3 + 2];
            float best = 0.0f;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                float dx = pc[k * 3 + 0] - x1;
                float dy = pc[k * 3 + 1] - y1;
                float dz = pc[k * 3 + 2] - z1;
                float d;
                if (contract)
                    d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                else {
                    float t = dx * dx + dy * dy;
                    d = t + dz * dz;
                }
                if (k == 0 || d < best) {
                    best = d;
                    besti = k;
                }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = besti;
        }
    }
}

/* NmDistanceKernelLauncher (tf_nndistance_g.cu:128-131): both directions. */
ORACLE_API void oracle_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2,
                                   float *dist1, int *idx1, float *dist2, int *idx2, int contract)
{
    nn_one_direction(b, n, m, xyz1, xyz2, dist1, idx1, contract);
    nn_one_direction(b, m, n, xyz2, xyz1, dist2, idx2, contract);
}

/* ------------------------------------------------------------------------- */
/* Chamfer gradient.                                                          */
/* ------------------------------------------------------------------------- */

/* NmDistanceGradKernel x2 (tf_nndistance_g.cu:132-157) == the CPU loops at
 * tf_nndistance.cpp:126-163.  The GPU scatters with float atomicAdd in an
 * unspecified order; this restatement uses the CPU loop's index order, which
 * is one of the orders the GPU may take.  Outputs are zeroed inside the op.
 */
ORACLE_API void oracle_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                        const float *grad_dist1, const int *idx1,
                                        const float *grad_dist2, const int *idx2,
                                        float *grad_xyz1, float *grad_xyz2)
{
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        float *g1 = grad_xyz1 + (size_t)i * n * 3;
        float *g2 = grad_xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            int j2 = idx1[(size_t)i * n + j];
            float g = grad_dist1[(size_t)i * n + j] * 2;
            for (int c = 0; c < 3; c++) {
                float t = g * (p1[j * 3 + c] - p2[j2 * 3 + c]);
                g1[j * 3 + c] += t;
                g2[j2 * 3 + c] -= t;
            }
        }
        for (int j = 0; j < m; j++) {
            int j2 = idx2[(size_t)i * m + j];
            float g = grad_dist2[(size_t)i * m + j] * 2;
            for (int c = 0; c < 3; c++) {
                float t = g * (p2[j * 3 + c] - p1[j2 * 3 + c]);
                g2[j * 3 + c] += t;
                g1[j2 * 3 + c] -= t;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* approx_match, GPU schedule.                                                */
/* ------------------------------------------------------------------------- */

static inline float sqdist_fma(float ax, float ay, float az, float bx, float by, float bz)
{
    /* (bx-ax)^2+(by-ay)^2+(bz-az)^2 as contracted in the reference SASS:
     * FMUL on the y term first, then FFMA x, then FFMA z. */
    float dx = bx - ax, dy = by - ay, dz = bz - az;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* __expf(level*d) as compiled for the reference (no -ftz, no fast-math):
 * t = (d*level)*1.4426950216f ; 2^t with denormal results preserved
 * (tf_approxmatch_g.cu:51-52).  MUFU.EX2 itself is replaced by a correctly
 * rounded exp2. */
static inline float expf_dev(float d, float level)
{
    float t = (d * level) * 1.4426950216293334961f;
    return (float)exp2((double)t);
}

/* levels of tf_approxmatch_g.cu:21-25: j=jstart..-2, level=-4^j, 0 at j=-2.
 * The GPU kernel uses jstart=7 (10 levels); the reference CPU function uses
 * jstart=8 (tf_approxmatch.cpp:31) and is a different schedule. */
ORACLE_API int oracle_num_levels(int jstart) { return jstart + 3; }

/*
 * approxmatch (tf_approxmatch_g.cu:1-179).
 *   xyz1 (b,n,3) "dataset", xyz2 (b,m,3) "query"
 *   match   (b,m,n) or NULL   -- dense soft assignment, match[i][l][k]
 *   factors (b,nlev,n+m) or NULL -- per level: ratioL[0..n) then ratioR[0..m)
 *           (the only per-level state that enters `match`: SURVEY.md 0.4)
 */
ORACLE_API void oracle_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2,
                                   float *match, float *factors, int jstart)
{
    int nlev = jstart + 3;
    float multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (float)(n / m); }
    else        { multiL = (float)(m / n); multiR = 1; }
    float *remainL = (float *)malloc(sizeof(float) * (size_t)(n + m) * 2);
    float *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        float *mt = match ? match + (size_t)i * n * m : NULL;
        if (mt) memset(mt, 0, sizeof(float) * (size_t)n * m);
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        int lev = 0;
        for (int j = jstart; j >= -2; j--, lev++) {
            float level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            /* sweep A (:26-59): ratioL[k] = remainL[k] / (1e-9 + sum_l E_kl remainR[l]) */
            for (int k = 0; k < n; k++) {
                float suml = 1e-9f;
                for (int l = 0; l < m; l++) {
                    float d = sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]);
                    suml = fmaf(expf_dev(d, level), remainR[l], suml);
                }
                ratioL[k] = remainL[k] / suml;
            }
            /* sweep B (:75-108) */
            for (int l = 0; l < m; l++) {
                float sumr = 0;
                for (int k = 0; k < n; k++) {
                    float d = sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]);
                    sumr = fmaf(expf_dev(d, level), ratioL[k], sumr);
                }
                sumr *= remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
            /* sweep C (:127-160): match[l][k] += E rl rr ; remainL[k] -= sum_l w */
            for (int k = 0; k < n; k++) {
                float suml = 0;
                float rl = ratioL[k];
                for (int l = 0; l < m; l++) {
                    float d = sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]);
                    float t = expf_dev(d, level) * rl;
                    if (mt) mt[(size_t)l * n + k] = fmaf(t, ratioR[l], mt[(size_t)l * n + k]);
                    suml = fmaf(t, ratioR[l], suml);
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            }
            if (factors) {
                float *f = factors + ((size_t)i * nlev + lev) * (n + m);
                memcpy(f, ratioL, sizeof(float) * n);
                memcpy(f + n, ratioR, sizeof(float) * m);
            }
        }
    }
    free(remainL);
}

/*
 * The same schedule with every per-point sum formed as partial sums over blocks of `chunk` streamed points, added
 * in block order afterwards (chunk <= 0: one sequential sum, i.e. oracle_approxmatch).  Not a restatement of
 * anything: it exists to show how far a change of SUMMATION ORDER ALONE moves the result on a given cloud
 * (tools/emd_order_sensitivity.py), which is the yardstick for comparing two fp32 implementations.
 */
/* running sum in blocks: `part` collects the current block, `total` the finished blocks (block order) */
typedef struct { float total, part; int left, chunk; } blocksum;
static inline void bs_init(blocksum *s, float init, int chunk, int count) { s->total = init; s->part = 0; s->chunk = chunk > 0 ? chunk : count; s->left = s->chunk; }
static inline void bs_add(blocksum *s, float term)
{
    s->part += term;
    if (--s->left == 0) { s->total += s->part; s->part = 0; s->left = s->chunk; }
}
static inline float bs_result(const blocksum *s) { return s->left == s->chunk ? s->total : s->total + s->part; }

ORACLE_API void oracle_approxmatch_order(int b, int n, int m, const float *xyz1, const float *xyz2,
                                         float *factors, int jstart, int chunk)
{
    int nlev = jstart + 3;
    float multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (float)(n / m); }
    else        { multiL = (float)(m / n); multiR = 1; }
    float *remainL = (float *)malloc(sizeof(float) * (size_t)(n + m) * 2);
    float *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        int lev = 0;
        for (int j = jstart; j >= -2; j--, lev++) {
            float level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
#pragma omp parallel for schedule(static)
            for (int k = 0; k < n; k++) {
                blocksum s; bs_init(&s, 1e-9f, chunk, m);
                for (int l = 0; l < m; l++)
                    bs_add(&s, expf_dev(sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]), level) * remainR[l]);
                ratioL[k] = remainL[k] / bs_result(&s);
            }
#pragma omp parallel for schedule(static)
            for (int l = 0; l < m; l++) {
                blocksum s; bs_init(&s, 0.0f, chunk, n);
                for (int k = 0; k < n; k++)
                    bs_add(&s, expf_dev(sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]), level) * ratioL[k]);
                float sumr = bs_result(&s) * remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
#pragma omp parallel for schedule(static)
            for (int k = 0; k < n; k++) {
                float rl = ratioL[k];
                blocksum s; bs_init(&s, 0.0f, chunk, m);
                for (int l = 0; l < m; l++)
                    bs_add(&s, (expf_dev(sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]), level) * rl) * ratioR[l]);
                remainL[k] = fmaxf(0.0f, remainL[k] - bs_result(&s));
            }
            float *f = factors + ((size_t)i * nlev + lev) * (n + m);
            memcpy(f, ratioL, sizeof(float) * n);
            memcpy(f + n, ratioR, sizeof(float) * m);
        }
    }
    free(remainL);
}

/* Dense match from the per-level factors, accumulated in level order exactly
 * like the `match[...]+=w` of tf_approxmatch_g.cu:152. */
ORACLE_API void oracle_match_from_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                                          const float *factors, int jstart, float *match)
{
    int nlev = jstart + 3;
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        float *mt = match + (size_t)i * n * m;
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) {
                float d = sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]);
                float acc = 0;
                int lev = 0;
                for (int j = jstart; j >= -2; j--, lev++) {
                    float level = (j == -2) ? 0.0f : -powf(4.0f, (float)j);
                    const float *f = factors + ((size_t)i * nlev + lev) * (n + m);
                    acc = fmaf(expf_dev(d, level) * f[k], f[n + l], acc);
                }
                mt[(size_t)l * n + k] = acc;
            }
    }
}

/* ------------------------------------------------------------------------- */
/* match_cost and its gradient (dense match, (b,m,n) layout).                 */
/* ------------------------------------------------------------------------- */

/* matchcost (tf_approxmatch_g.cu:183-225): cost[i] = sum_kl sqrtf(d_kl) match[i][l][k].
 * IEEE sqrtf.  The GPU sums per thread then over a 512-wide tree; the order
 * is not restated -- the sum is taken in double and rounded once, so it is the
 * value any fp32 summation order approximates. */
ORACLE_API void oracle_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 const float *match, float *cost)
{
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        double s = 0;
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) {
                float d = sqdist_fma(p1[k*3], p1[k*3+1], p1[k*3+2], p2[l*3], p2[l*3+1], p2[l*3+2]);
                s += (double)(sqrtf(d) * mt[(size_t)l * n + k]);
            }
        cost[i] = (float)s;
    }
}

/* matchcostgrad1/2 (tf_approxmatch_g.cu:229-291):
 *   grad1[k] = sum_l match[l][k] (x1_k - x2_l) rsqrtf(max(d,1e-20))
 *   grad2[l] = sum_k match[l][k] (x2_l - x1_k) rsqrtf(max(d,1e-20))
 * float terms as on the GPU (d' = match*rsqrt ; acc += diff*d'), double
 * accumulation (order-free reference value).  rsqrtf (MUFU.RSQ, 2 ulp) is
 * replaced by a correctly rounded 1/sqrt.
 */
ORACLE_API void oracle_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2,
                                     const float *match, float *grad1, float *grad2)
{
    double *acc = (double *)malloc(sizeof(double) * 3 * (size_t)(n + m));
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        double *a1 = acc, *a2 = acc + 3 * (size_t)n;
        memset(acc, 0, sizeof(double) * 3 * (size_t)(n + m));
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) {
                float dx = p1[k*3] - p2[l*3], dy = p1[k*3+1] - p2[l*3+1], dz = p1[k*3+2] - p2[l*3+2];
                float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float r = (float)(1.0 / sqrt((double)fmaxf(d, 1e-20f)));
                float w = mt[(size_t)l * n + k] * r;
                a1[k*3+0] += (double)(dx * w); a1[k*3+1] += (double)(dy * w); a1[k*3+2] += (double)(dz * w);
                a2[l*3+0] -= (double)(dx * w); a2[l*3+1] -= (double)(dy * w); a2[l*3+2] -= (double)(dz * w);
            }
        for (int k = 0; k < n * 3; k++) grad1[(size_t)i * n * 3 + k] = (float)a1[k];
        for (int l = 0; l < m * 3; l++) grad2[(size_t)i * m * 3 + l] = (float)a2[l];
    }
    free(acc);
}

/* match_cost and both gradients straight from the factors (what the product's
 * fused kernel computes) -- used to check the factor path without a dense
 * (b,m,n) tensor at sizes where that tensor would not fit a test. */
ORACLE_API void oracle_matchcost_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                                         const float *factors, int jstart,
                                         float *cost, float *grad1, float *grad2)
{
    int nlev = jstart + 3;
    double *acc = (double *)malloc(sizeof(double) * 3 * (size_t)(n + m));
    float *lv = (float *)malloc(sizeof(float) * nlev);
    for (int j = jstart, t = 0; j >= -2; j--, t++) lv[t] = (j == -2) ? 0.0f : -powf(4.0f, (float)j);
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        double *a1 = acc, *a2 = acc + 3 * (size_t)n;
        memset(acc, 0, sizeof(double) * 3 * (size_t)(n + m));
        double s = 0;
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) {
                float dx = p1[k*3] - p2[l*3], dy = p1[k*3+1] - p2[l*3+1], dz = p1[k*3+2] - p2[l*3+2];
                float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float mv = 0;
                for (int t = 0; t < nlev; t++) {
                    const float *f = factors + ((size_t)i * nlev + t) * (n + m);
                    mv = fmaf(expf_dev(d, lv[t]) * f[k], f[n + l], mv);
                }
                s += (double)(sqrtf(d) * mv);
                float r = (float)(1.0 / sqrt((double)fmaxf(d, 1e-20f)));
                float w = mv * r;
                a1[k*3+0] += (double)(dx * w); a1[k*3+1] += (double)(dy * w); a1[k*3+2] += (double)(dz * w);
                a2[l*3+0] -= (double)(dx * w); a2[l*3+1] -= (double)(dy * w); a2[l*3+2] -= (double)(dz * w);
            }
        cost[i] = (float)s;
        if (grad1) for (int k = 0; k < n * 3; k++) grad1[(size_t)i * n * 3 + k] = (float)a1[k];
        if (grad2) for (int l = 0; l < m * 3; l++) grad2[(size_t)i * m * 3 + l] = (float)a2[l];
    }
    free(acc);
    free(lv);
}

/* ------------------------------------------------------------------------- */
/* fp64 ground truth of the whole EMD pipeline.                               */
/* ------------------------------------------------------------------------- */

/* approxmatch -> matchcost -> matchcostgrad (tf_approxmatch_g.cu:1-295, GPU schedule, jstart levels) evaluated
 * end to end in double: distances, exponentials, every sum and division.  The float constants of the reference
 * (1e-9f, 1e-20f, multiL/multiR, the levels) keep their float values.  This is the value every fp32 evaluation
 * order of the algorithm approximates; tests use it to rank the product and the reference kernels by their
 * distance to it (the algorithm is ill-conditioned in fp32 on some shapes, so "within 1e-4 of another fp32
 * evaluation" is not always attainable -- "no farther from the truth than the reference" is).
 *   cost (b,) , grad1 (b,n,3), grad2 (b,m,3) as doubles; factors64 (b,nlev,n+m) doubles or NULL. */
ORACLE_API void oracle_emd_fp64(int b, int n, int m, const float *xyz1, const float *xyz2, int jstart,
                                double *cost, double *grad1, double *grad2, double *factors64)
{
    int nlev = jstart + 3;
    double multiL, multiR;
    if (n >= m) { multiL = 1; multiR = (double)(n / m); }
    else        { multiL = (double)(m / n); multiR = 1; }
    const double eps9 = (double)1e-9f, eps20 = (double)1e-20f;
    double *remainL = (double *)malloc(sizeof(double) * (size_t)(n + m) * 2);
    double *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;
    double *fac = (double *)malloc(sizeof(double) * (size_t)nlev * (n + m));
    double *lv = (double *)malloc(sizeof(double) * nlev);
    for (int j = jstart, t = 0; j >= -2; j--, t++) lv[t] = (j == -2) ? 0.0 : -pow(4.0, (double)j);
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        for (int lev = 0; lev < nlev; lev++) {
            const double level = lv[lev];
#pragma omp parallel for schedule(static)
            for (int k = 0; k < n; k++) {
                double suml = eps9;
                for (int l = 0; l < m; l++) {
                    double dx = (double)p2[l*3] - p1[k*3], dy = (double)p2[l*3+1] - p1[k*3+1], dz = (double)p2[l*3+2] - p1[k*3+2];
                    suml += exp(level * (dx*dx + dy*dy + dz*dz)) * remainR[l];
                }
                ratioL[k] = remainL[k] / suml;
            }
#pragma omp parallel for schedule(static)
            for (int l = 0; l < m; l++) {
                double sumr = 0;
                for (int k = 0; k < n; k++) {
                    double dx = (double)p2[l*3] - p1[k*3], dy = (double)p2[l*3+1] - p1[k*3+1], dz = (double)p2[l*3+2] - p1[k*3+2];
                    sumr += exp(level * (dx*dx + dy*dy + dz*dz)) * ratioL[k];
                }
                sumr *= remainR[l];
                double consumption = fmin(remainR[l] / (sumr + eps9), 1.0);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmax(0.0, remainR[l] - sumr);
            }
#pragma omp parallel for schedule(static)
            for (int k = 0; k < n; k++) {
                double suml = 0;
                for (int l = 0; l < m; l++) {
                    double dx = (double)p2[l*3] - p1[k*3], dy = (double)p2[l*3+1] - p1[k*3+1], dz = (double)p2[l*3+2] - p1[k*3+2];
                    suml += exp(level * (dx*dx + dy*dy + dz*dz)) * ratioL[k] * ratioR[l];
                }
                remainL[k] = fmax(0.0, remainL[k] - suml);
            }
            memcpy(fac + (size_t)lev * (n + m), ratioL, sizeof(double) * n);
            memcpy(fac + (size_t)lev * (n + m) + n, ratioR, sizeof(double) * m);
        }
        if (factors64) memcpy(factors64 + (size_t)i * nlev * (n + m), fac, sizeof(double) * (size_t)nlev * (n + m));
        double s = 0;
        double *g1 = grad1 + (size_t)i * n * 3, *g2 = grad2 + (size_t)i * m * 3;
        /* two passes so that every output has one owner: rows own cost terms and grad1, columns own grad2
         * (each thread sums its own point in index order: the result does not depend on the thread count) */
        double *rowcost = (double *)malloc(sizeof(double) * n);
#pragma omp parallel for schedule(static)
        for (int k = 0; k < n; k++) {
            double c = 0, ax = 0, ay = 0, az = 0;
            for (int l = 0; l < m; l++) {
                double dx = (double)p1[k*3] - p2[l*3], dy = (double)p1[k*3+1] - p2[l*3+1], dz = (double)p1[k*3+2] - p2[l*3+2];
                double d = dx*dx + dy*dy + dz*dz;
                double mv = 0;
                for (int t = 0; t < nlev; t++)
                    mv += exp(lv[t] * d) * fac[(size_t)t * (n + m) + k] * fac[(size_t)t * (n + m) + n + l];
                c += sqrt(d) * mv;
                double w = mv / sqrt(fmax(d, eps20));
                ax += dx * w; ay += dy * w; az += dz * w;
            }
            rowcost[k] = c; g1[k*3] = ax; g1[k*3+1] = ay; g1[k*3+2] = az;
        }
        for (int k = 0; k < n; k++) s += rowcost[k];
        free(rowcost);
#pragma omp parallel for schedule(static)
        for (int l = 0; l < m; l++) {
            double ax = 0, ay = 0, az = 0;
            for (int k = 0; k < n; k++) {
                double dx = (double)p1[k*3] - p2[l*3], dy = (double)p1[k*3+1] - p2[l*3+1], dz = (double)p1[k*3+2] - p2[l*3+2];
                double d = dx*dx + dy*dy + dz*dz;
                double mv = 0;
                for (int t = 0; t < nlev; t++)
                    mv += exp(lv[t] * d) * fac[(size_t)t * (n + m) + k] * fac[(size_t)t * (n + m) + n + l];
                double w = mv / sqrt(fmax(d, eps20));
                ax -= dx * w; ay -= dy * w; az -= dz * w;
            }
            g2[l*3] = ax; g2[l*3+1] = ay; g2[l*3+2] = az;
        }
        cost[i] = s;
    }
    free(remainL);
    free(fac);
    free(lv);
}
