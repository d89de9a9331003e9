/*
 * synth_lp.c -- deterministic synthetic LP generator for the BASELINE.json configs (test/bench
 * infrastructure; the reference ships no generator -- SURVEY.md 8d defines this one).
 *
 *   min c'x  s.t.  AL <= A x <= AU,  l <= x <= u          CSR, int32 indices, fp64
 *
 * Counter-based RNG (splitmix64 of (seed, stream, index)): generation is independent of thread
 * count and order, so any row block can be regenerated on its own (needed for sharded generation).
 *   kind 0 "uniform"  : every row has exactly nnz/m distinct columns drawn uniformly, sorted
 *   kind 1 "powerlaw" : row length clamp(floor(s * u^(-1/1.5)), 1, 200000), s chosen so that the
 *                       total is exactly nnz (heavy tail; load-balance stress)
 *   values U(-1,1) with |v| >= 1e-3.
 * Feasible and bounded by construction: a primal-dual pair (x*, y*, z*) is drawn first
 *   x*_j = 0 for half of the columns (then z*_j >= 0), else U(0,1) interior (z*_j = 0);
 *   l = 0, u = +inf except 10 % of columns with u = 1;
 *   rows: 1/3 equality (AL = AU = (Ax*)_i, y*_i free), 1/3 "<=" (AL = -inf; half active with
 *   y*_i <= 0, half slack with y*_i = 0), 1/3 ranged (half active at the lower side with y*_i >= 0,
 *   half strictly inside with y*_i = 0);
 *   c = A'y* + z*   =>  (x*, y*, z*) satisfies the KKT conditions, c'x* is the known optimal value.
 * vec_seed selects the (x*, y*, z*) draw independently of the matrix seed: batch member k of the
 * shared-A config uses vec_seed = seed + k.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t rng(uint64_t seed, uint64_t stream, uint64_t idx) {
    return splitmix64(splitmix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + idx * 0x9E3779B97F4A7C15ULL);
}
static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

static long long powerlaw_len(double scale, double u, int maxlen) {
    double v = floor(scale * pow(u, -1.0 / 1.5));
    if (v < 1.0) v = 1.0;
    if (v > (double)maxlen) v = (double)maxlen;
    return (long long)v;
}

/* Fills rowPtr (m+1). Returns nnz. */
long long synth_lp_rowptr(int kind, int m, int n, long long nnz_target, uint64_t seed, int *rowPtr) {
    long long total = 0;
    if (kind == 0) {
        long long k = nnz_target / m;
        if (k < 1) k = 1;
        if (k > n) k = n;
        rowPtr[0] = 0;
        for (int i = 0; i < m; ++i) rowPtr[i + 1] = (int)((long long)(i + 1) * k);
        return (long long)m * k;
    }
    const int maxlen = n < 200000 ? n : 200000;
    double lo = 1e-3, hi = (double)nnz_target / m * 4.0 + 8.0;
    for (int it = 0; it < 80; ++it) {
        double mid = 0.5 * (lo + hi);
        long long s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int i = 0; i < m; ++i) s += powerlaw_len(mid, 1.0 - u01(rng(seed, 1, (uint64_t)i)), maxlen);
        if (s > nnz_target) hi = mid; else lo = mid;
    }
    long long *len = (long long *)malloc(sizeof(long long) * (size_t)m);
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (int i = 0; i < m; ++i) {
        len[i] = powerlaw_len(lo, 1.0 - u01(rng(seed, 1, (uint64_t)i)), maxlen);
        total += len[i];
    }
    /* exact total: hand the remaining nonzeros out one per row, cyclically */
    long long rem = nnz_target - total;
    for (int i = 0; rem > 0; i = (i + 1) % m)
        if (len[i] < maxlen) { len[i]++; rem--; }
    rowPtr[0] = 0;
    for (int i = 0; i < m; ++i) rowPtr[i + 1] = rowPtr[i] + (int)len[i];
    total = rowPtr[m];
    free(len);
    return total;
}

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* Fills colIndex/values for rows [r0, r1) given rowPtr. */
void synth_lp_matrix_rows(int n, uint64_t seed, const int *rowPtr, int r0, int r1, int *col, double *val) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = r0; i < r1; ++i) {
        const int p0 = rowPtr[i], len = rowPtr[i + 1] - rowPtr[i];
        int *c = col + p0;
        uint64_t ctr = 0;
        for (int k = 0; k < len; ++k) c[k] = (int)(rng(seed, 2 + 4 * (uint64_t)i, ctr++) % (uint64_t)n);
        for (;;) {   /* make the columns distinct: re-draw duplicates */
            qsort(c, (size_t)len, sizeof(int), cmp_int);
            int dup = 0;
            for (int k = 1; k < len; ++k)
                if (c[k] == c[k - 1]) { c[k - 1] = (int)(rng(seed, 2 + 4 * (uint64_t)i, ctr++) % (uint64_t)n); dup = 1; }
            if (!dup) break;
        }
        uint64_t vctr = 0;
        for (int k = 0; k < len; ++k) {
            double v;
            do { v = 2.0 * u01(rng(seed, 3 + 4 * (uint64_t)i, vctr++)) - 1.0; } while (fabs(v) < 1e-3);
            val[p0 + k] = v;
        }
    }
}

/* "banded" twin (structured matrix of the same size): the columns of row i are drawn from a `window`-wide slice around
 * i*n/m, so consecutive rows gather neighbouring entries (and so do the rows of the transpose). */
void synth_lp_matrix_rows_banded(int m, int n, int window, uint64_t seed, const int *rowPtr, int r0, int r1, int *col, double *val) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = r0; i < r1; ++i) {
        const int p0 = rowPtr[i], len = rowPtr[i + 1] - rowPtr[i];
        int *c = col + p0;
        int w = window < len ? len : window;
        if (w > n) w = n;
        long long lo = (long long)i * n / m - w / 2;
        if (lo < 0) lo = 0;
        if (lo + w > n) lo = n - w;
        uint64_t ctr = 0;
        for (int k = 0; k < len; ++k) c[k] = (int)(lo + (long long)(rng(seed, 2 + 4 * (uint64_t)i, ctr++) % (uint64_t)w));
        for (;;) {
            qsort(c, (size_t)len, sizeof(int), cmp_int);
            int dup = 0;
            for (int k = 1; k < len; ++k)
                if (c[k] == c[k - 1]) { c[k - 1] = (int)(lo + (long long)(rng(seed, 2 + 4 * (uint64_t)i, ctr++) % (uint64_t)w)); dup = 1; }
            if (!dup) break;
        }
        uint64_t vctr = 0;
        for (int k = 0; k < len; ++k) {
            double v;
            do { v = 2.0 * u01(rng(seed, 3 + 4 * (uint64_t)i, vctr++)) - 1.0; } while (fabs(v) < 1e-3);
            val[p0 + k] = v;
        }
    }
}

/* "blocked" twin: dense run x run blocks.  The `run` rows of a row group share their columns, which come as runs of
 * `run` consecutive columns (aligned to `run`) drawn from a `window`-wide slice around the group's diagonal position, so a
 * warp's gathers coalesce into whole 64-byte segments in the pass over A and in the pass over the transpose. */
void synth_lp_matrix_rows_blocked(int m, int n, int window, int run, uint64_t seed, const int *rowPtr, int r0, int r1, int *col, double *val) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = r0; i < r1; ++i) {
        const int p0 = rowPtr[i], len = rowPtr[i + 1] - rowPtr[i];
        int *c = col + p0;
        const int g = i / run;
        const int nb = (len + run - 1) / run;
        int w = window < nb * run ? nb * run : window;
        if (w > n) w = n;
        long long lo = (long long)g * run * n / m - w / 2;
        if (lo < 0) lo = 0;
        if (lo + w > n) lo = n - w;
        lo = lo / run * run;
        const int wb = w / run;               /* block columns in the window */
        int blk[64];
        if (nb > 64 || nb > wb) { fprintf(stderr, "synth_lp blocked: row too long\n"); abort(); }
        uint64_t ctr = 0;
        for (int k = 0; k < nb; ++k) blk[k] = (int)(rng(seed, 2 + 4 * (uint64_t)g, ctr++) % (uint64_t)wb);
        for (;;) {
            qsort(blk, (size_t)nb, sizeof(int), cmp_int);
            int dup = 0;
            for (int k = 1; k < nb; ++k)
                if (blk[k] == blk[k - 1]) { blk[k - 1] = (int)(rng(seed, 2 + 4 * (uint64_t)g, ctr++) % (uint64_t)wb); dup = 1; }
            if (!dup) break;
        }
        for (int k = 0; k < len; ++k) c[k] = (int)(lo + (long long)blk[k / run] * run + k % run);
        uint64_t vctr = 0;
        for (int k = 0; k < len; ++k) {
            double v;
            do { v = 2.0 * u01(rng(seed, 3 + 4 * (uint64_t)i, vctr++)) - 1.0; } while (fabs(v) < 1e-3);
            val[p0 + k] = v;
        }
    }
}

/* Draws (x*, y*, z*) with vec_seed and fills AL, AU, l, u, c; returns c'x*. */
double synth_lp_vectors(int m, int n, const int *rowPtr, const int *col, const double *val, uint64_t seed,
                        uint64_t vec_seed, double *AL, double *AU, double *l, double *u, double *c,
                        double *xs, double *ys, double *zs) {
    const uint64_t S = vec_seed;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j) {
        const int at_zero = (rng(S, 11, (uint64_t)j) & 1ULL) != 0;
        l[j] = 0.0;
        u[j] = (rng(seed, 12, (uint64_t)j) % 10ULL == 0) ? 1.0 : INFINITY;   /* bound pattern tied to the matrix seed */
        if (at_zero) { xs[j] = 0.0; zs[j] = u01(rng(S, 13, (uint64_t)j)); }
        else { xs[j] = 0.05 + 0.9 * u01(rng(S, 14, (uint64_t)j)); zs[j] = 0.0; }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        double ax = 0.0;
        for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) ax += val[k] * xs[col[k]];
        const uint64_t type = rng(seed, 15, (uint64_t)i) % 3ULL;        /* row type tied to the matrix seed */
        const int active = (rng(seed, 16, (uint64_t)i) & 1ULL) != 0;
        const double r1 = u01(rng(S, 17, (uint64_t)i)), r2 = u01(rng(S, 18, (uint64_t)i));
        if (type == 0) { AL[i] = ax; AU[i] = ax; ys[i] = 2.0 * r1 - 1.0; }
        else if (type == 1) {
            AL[i] = -INFINITY;
            if (active) { AU[i] = ax; ys[i] = -r1; } else { AU[i] = ax + 0.1 + r2; ys[i] = 0.0; }
        } else {
            if (active) { AL[i] = ax; AU[i] = ax + 0.5 + r2; ys[i] = r1; }
            else { AL[i] = ax - 0.1 - r1; AU[i] = ax + 0.1 + r2; ys[i] = 0.0; }
        }
    }
    for (int j = 0; j < n; ++j) c[j] = zs[j];
    for (int i = 0; i < m; ++i) {
        const double yi = ys[i];
        if (yi == 0.0) continue;
        for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) c[col[k]] += val[k] * yi;
    }
    double obj = 0.0;
    for (int j = 0; j < n; ++j) obj += c[j] * xs[j];
    return obj;
}
