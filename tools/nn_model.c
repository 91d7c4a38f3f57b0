/*
 * nn_model.c -- CPU model of the Chamfer forward's "approximate sweep + exact finalize" scheme
 * (csrc/nn_distance.cu), used to check the candidate logic without a GPU:
 *   sweep:    s = A_j + B_k - 2 a'_j.b'_k (4 fp32 ops, centred coordinates, inflated norms => s >= 0),
 *             per (row, span) the best chunk, one more in-band chunk, or an overflow mark;
 *             per (column, row block) the minimum and the ballot of in-band lanes;
 *   finalize: exact re-evaluation (reference rounding) of every in-band candidate.
 * The result must equal the oracle bit for bit on any input.  Same arithmetic as the kernel (fmaf).
 *
 *   gcc -O2 -ffp-contract=off tools/nn_model.c -o /tmp/nn_model -lm && /tmp/nn_model
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KR 8
#define ROWS 256
#define CHUNK 32
static const float K_INFL = 1.0f + 1.0f / 524288.0f;   /* 1 + 2^-19 */
static const float K_BAND = 1.0f / 16384.0f;            /* 2^-14 */

typedef unsigned long long u64;

static float sqd(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, dy * dy)); }
static float thr_of(float s, float nrm) { return fmaf(nrm, K_BAND, fmaf(s, K_BAND, s)); }
static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void centre(const float *p1, int n, const float *p2, int m, float *o)
{
    for (int c = 0; c < 3; c++)
        o[c] = ((p1[c] + p1[(n / 2) * 3 + c]) + (p2[c] + p2[(m / 2) * 3 + c])) * 0.25f;
}
static float norm_infl(float x, float y, float z) { return sqd(x, y, z) * K_INFL; }

static long stat_rows_slow, stat_rows_rescan, stat_cols_slow, stat_neg;

/* one element; warps/units emulate the stream-K split (spans of `span` chunks inside a row block) */
static void model(int n, const float *p1, int m, const float *p2, int span, float *d1, int *i1, float *d2, int *i2)
{
    const int nrb = (n + ROWS - 1) / ROWS, nch = (m + CHUNK - 1) / CHUNK;
    int tsh = 0;
    while (((nch - 1) >> tsh) + 1 > 65534) tsh++;
    float o[3];
    centre(p1, n, p2, m, o);
    const int nslot = (nch + span - 1) / span, maxspan = span;
    u64 *rowkeys = malloc(sizeof(u64) * (size_t)nrb * nslot * ROWS);
    u64 *colkeys = malloc(sizeof(u64) * (size_t)nrb * m);
    float *A = malloc(sizeof(float) * nrb * ROWS), *ax = malloc(sizeof(float) * nrb * ROWS * 3);
    float *B = malloc(sizeof(float) * nch * CHUNK), *bx = malloc(sizeof(float) * nch * CHUNK * 3);
    for (int j = 0; j < nrb * ROWS; j++) {
        int jj = j < n ? j : n - 1;
        float x = p1[jj * 3] - o[0], y = p1[jj * 3 + 1] - o[1], z = p1[jj * 3 + 2] - o[2];
        ax[j * 3] = -2.0f * x; ax[j * 3 + 1] = -2.0f * y; ax[j * 3 + 2] = -2.0f * z;
        A[j] = j < n ? norm_infl(x, y, z) : INFINITY;
    }
    for (int k = 0; k < nch * CHUNK; k++) {
        int kk = k < m ? k : m - 1;
        float x = p2[kk * 3] - o[0], y = p2[kk * 3 + 1] - o[1], z = p2[kk * 3 + 2] - o[2];
        bx[k * 3] = x; bx[k * 3 + 1] = y; bx[k * 3 + 2] = z;
        B[k] = k < m ? norm_infl(x, y, z) : INFINITY;
    }
    /* ---- sweep */
    for (int rb = 0; rb < nrb; rb++)
        for (int sl = 0; sl < nslot; sl++) {
            float snap[ROWS]; uint32_t tw[ROWS];
            for (int r = 0; r < ROWS; r++) { snap[r] = INFINITY; tw[r] = 0; }
            int ch_end = (sl + 1) * span < nch ? (sl + 1) * span : nch;
            for (int ch = sl * span; ch < ch_end; ch++) {
                float cm[ROWS];
                for (int r = 0; r < ROWS; r++) cm[r] = INFINITY;
                for (int c = 0; c < CHUNK; c++) {
                    int k = ch * CHUNK + c;
                    float lanemin[32];
                    for (int l = 0; l < 32; l++) lanemin[l] = INFINITY;
                    for (int r = 0; r < ROWS; r++) {
                        int j = rb * ROWS + r;
                        float s = fmaf(ax[j * 3 + 2], bx[k * 3 + 2], fmaf(ax[j * 3 + 1], bx[k * 3 + 1], fmaf(ax[j * 3], bx[k * 3], A[j] + B[k])));
                        if (s < 0) stat_neg++;
                        cm[r] = fminf(cm[r], s);
                        lanemin[r / KR] = fminf(lanemin[r / KR], s);
                    }
                    float mn = INFINITY;
                    for (int l = 0; l < 32; l++) mn = fminf(mn, lanemin[l]);
                    float th = thr_of(mn, B[k]);
                    uint32_t who = 0;
                    for (int l = 0; l < 32; l++) if (lanemin[l] <= th) who |= 1u << l;
                    if (k < m) colkeys[(size_t)rb * m + k] = ((u64)fbits(mn) << 32) | who;
                }
                const uint32_t bit = 1u << ((ch - sl * span) & 15);
                for (int r = 0; r < ROWS; r++) {
                    int j = rb * ROWS + r;
                    float lo = fminf(cm[r], snap[r]), hi = fmaxf(cm[r], snap[r]);
                    int inb = hi <= thr_of(lo, A[j]);
                    int imp = cm[r] < snap[r];
                    tw[r] = ((imp && !inb) ? 0u : tw[r]) | ((imp || inb) ? bit : 0u);
                    snap[r] = lo;
                }
            }
            for (int r = 0; r < ROWS; r++)
                rowkeys[((size_t)rb * nslot + sl) * ROWS + r] = ((u64)fbits(snap[r]) << 32) | ((uint32_t)((sl * span) >> tsh) << 16) | tw[r];
        }
    /* ---- finalize: rows */
    for (int j = 0; j < n; j++) {
        const int rb = j / ROWS, r = j % ROWS;
        const float x = p1[j * 3], y = p1[j * 3 + 1], z = p1[j * 3 + 2];
        const float Aj = norm_infl(x - o[0], y - o[1], z - o[2]);
        u64 g = ~0ull;
        for (int sl = 0; sl < nslot; sl++) { u64 k = rowkeys[((size_t)rb * nslot + sl) * ROWS + r]; if (k < g) g = k; }
        const float th = thr_of(bitsf((uint32_t)(g >> 32)), Aj);
        u64 best = ~0ull;
        int ncand = 0;
        for (int sl = 0; sl < nslot; sl++) {
            u64 k = rowkeys[((size_t)rb * nslot + sl) * ROWS + r];
            if (!(bitsf((uint32_t)(k >> 32)) <= th)) continue;
            const int c0 = (int)(((uint32_t)k >> 16) & 0xffff) << tsh;
            const int cend = c0 + maxspan + ((1 << tsh) - 1) < nch ? c0 + maxspan + ((1 << tsh) - 1) : nch;
            for (int beta = 0; beta < 16; beta++) if (((uint32_t)k >> beta) & 1)
                for (int ch = c0 + beta; ch < cend; ch += 16) {
                    ncand++;
                    for (int c = ch * CHUNK; c < (ch + 1) * CHUNK && c < m; c++) {
                        float d = sqd(p2[c * 3] - x, p2[c * 3 + 1] - y, p2[c * 3 + 2] - z);
                        u64 kk = ((u64)fbits(d) << 32) | (uint32_t)c;
                        if (kk < best) best = kk;
                    }
                }
        }
        if (ncand > 1) stat_rows_slow++;
        if (ncand > 4) stat_rows_rescan++;
        d1[j] = bitsf((uint32_t)(best >> 32)); i1[j] = (int)(uint32_t)best;
    }
    /* ---- finalize: columns */
    for (int k = 0; k < m; k++) {
        const float x = p2[k * 3], y = p2[k * 3 + 1], z = p2[k * 3 + 2];
        const float Bk = norm_infl(x - o[0], y - o[1], z - o[2]);
        uint32_t gm = 0xffffffffu;
        for (int rb = 0; rb < nrb; rb++) { uint32_t v = (uint32_t)(colkeys[(size_t)rb * m + k] >> 32); if (v < gm) gm = v; }
        const float th = thr_of(bitsf(gm), Bk);
        u64 best = ~0ull;
        int ncand = 0;
        for (int rb = 0; rb < nrb; rb++) {
            u64 v = colkeys[(size_t)rb * m + k];
            if (!(bitsf((uint32_t)(v >> 32)) <= th)) continue;
            for (int l = 0; l < 32; l++) if (((uint32_t)v >> l) & 1) {
                ncand++;
                for (int r = 0; r < KR; r++) {
                    int j = rb * ROWS + l * KR + r;
                    if (j >= n) j = n - 1;
                    float d = sqd(x - p1[j * 3], y - p1[j * 3 + 1], z - p1[j * 3 + 2]);
                    u64 kk = ((u64)fbits(d) << 32) | (uint32_t)j;
                    if (kk < best) best = kk;
                }
            }
        }
        if (ncand > 1) stat_cols_slow++;
        d2[k] = bitsf((uint32_t)(best >> 32)); i2[k] = (int)(uint32_t)best;
    }
    free(rowkeys); free(colkeys); free(A); free(ax); free(B); free(bx);
}

static void oracle_dir(int n, const float *a, int m, const float *c, float *dist, int *idx, int flip)
{
    for (int j = 0; j < n; j++) {
        float best = 0; int bi = 0;
        for (int k = 0; k < m; k++) {
            float d = flip ? sqd(a[j * 3] - c[k * 3], a[j * 3 + 1] - c[k * 3 + 1], a[j * 3 + 2] - c[k * 3 + 2])
                           : sqd(c[k * 3] - a[j * 3], c[k * 3 + 1] - a[j * 3 + 1], c[k * 3 + 2] - a[j * 3 + 2]);
            if (k == 0 || d < best) { best = d; bi = k; }
        }
        dist[j] = best; idx[j] = bi;
    }
}

static uint64_t rng = 88172645463325252ull;
static double urand(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (rng >> 11) * (1.0 / 9007199254740992.0); }
static double nrand(void) { double u = urand() + 1e-300, v = urand(); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); }

static int run_case(const char *name, int n, int m, int span, int kind, float offset, float scale)
{
    float *p1 = malloc(sizeof(float) * n * 3), *p2 = malloc(sizeof(float) * m * 3);
    for (int i = 0; i < n * 3; i++) p1[i] = 0;
    if (kind == 0) {            /* gaussian clouds */
        for (int i = 0; i < n * 3; i++) p1[i] = (float)(nrand() * scale + offset);
        for (int i = 0; i < m * 3; i++) p2[i] = (float)(nrand() * scale + offset);
    } else if (kind == 1) {     /* pred = label permuted + small noise, with duplicated label points */
        for (int i = 0; i < m; i++) {
            int src = (i > 0 && urand() < 0.3) ? (int)(urand() * i) : -1;
            for (int c = 0; c < 3; c++) p2[i * 3 + c] = src >= 0 ? p2[src * 3 + c] : (float)((urand() * 2 - 1) * scale + offset);
        }
        for (int i = 0; i < n; i++) {
            int src = (int)(urand() * m);
            for (int c = 0; c < 3; c++) p1[i * 3 + c] = p2[src * 3 + c] + (float)(nrand() * 0.02 * scale);
        }
    } else if (kind == 2) {     /* integer lattice: masses of exact ties */
        for (int i = 0; i < n * 3; i++) p1[i] = (float)((int)(urand() * 6) * scale + offset);
        for (int i = 0; i < m * 3; i++) p2[i] = (float)((int)(urand() * 6) * scale + offset);
    } else {                    /* a few distinct points repeated many times */
        for (int i = 0; i < m; i++) for (int c = 0; c < 3; c++) p2[i * 3 + c] = (float)(((i % 7) * 0.37 + c * 0.11) * scale + offset);
        for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) p1[i * 3 + c] = (float)(((i % 5) * 0.37 + c * 0.11) * scale + offset);
    }
    float *d1 = malloc(4 * n), *d2 = malloc(4 * m), *e1 = malloc(4 * n), *e2 = malloc(4 * m);
    int *i1 = malloc(4 * n), *i2 = malloc(4 * m), *j1 = malloc(4 * n), *j2 = malloc(4 * m);
    stat_rows_slow = stat_rows_rescan = stat_cols_slow = stat_neg = 0;
    model(n, p1, m, p2, span, d1, i1, d2, i2);
    oracle_dir(n, p1, m, p2, e1, j1, 0);
    oracle_dir(m, p2, n, p1, e2, j2, 1);
    int bad = 0;
    for (int j = 0; j < n; j++) if (fbits(d1[j]) != fbits(e1[j]) || i1[j] != j1[j]) bad++;
    for (int k = 0; k < m; k++) if (fbits(d2[k]) != fbits(e2[k]) || i2[k] != j2[k]) bad++;
    printf("%-28s n=%5d m=%5d span=%2d  mismatches %d | rows: >1 chunk %ld, >4 chunks %ld of %d | cols slow %ld of %d | negative s %ld\n",
           name, n, m, span, bad, stat_rows_slow, stat_rows_rescan, n, stat_cols_slow, m, stat_neg);
    free(p1); free(p2); free(d1); free(d2); free(e1); free(e2); free(i1); free(i2); free(j1); free(j2);
    return bad;
}

int main(void)
{
    int bad = 0;
    bad += run_case("randn", 2048, 2048, 7, 0, 0.f, 1.f);
    bad += run_case("randn span 64", 2048, 2048, 64, 0, 0.f, 1.f);
    bad += run_case("randn ragged", 777, 1030, 3, 0, 0.f, 1.f);
    bad += run_case("randn tiny", 5, 6, 1, 0, 0.f, 1.f);
    bad += run_case("randn far from origin", 1024, 1024, 7, 0, 1000.f, 1.f);
    bad += run_case("randn huge scale", 1024, 1024, 7, 0, 0.f, 1e6f);
    bad += run_case("randn tiny scale", 1024, 1024, 7, 0, 0.f, 1e-6f);
    bad += run_case("pred~label, duplicates", 2048, 2048, 7, 1, 0.f, 1.f);
    bad += run_case("pred~label, dup, span 64", 2048, 2048, 64, 1, 0.f, 1.f);
    bad += run_case("pred~label, dup, offset", 2048, 2048, 7, 1, 50.f, 1.f);
    bad += run_case("lattice ties", 1500, 1300, 5, 2, 0.f, 1.f);
    bad += run_case("lattice ties offset", 1500, 1300, 5, 2, 3.f, 0.1f);
    bad += run_case("7 points repeated", 600, 700, 4, 3, 0.f, 1.f);
    bad += run_case("n=1", 1, 700, 7, 0, 0.f, 1.f);
    bad += run_case("m=1", 300, 1, 7, 0, 0.f, 1.f);
    bad += run_case("16384 x 1024", 16384, 1024, 32, 0, 0.f, 1.f);
    printf(bad ? "FAILED\n" : "all cases bit-exact\n");
    return bad != 0;
}
