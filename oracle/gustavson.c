/*
 * oracle/gustavson.c -- TEST INFRASTRUCTURE ONLY (not the product, never shipped).
 *
 * CPU restatement of the MH-SpGEMM hot path (C = A*B, CSR in / CSR out, int32
 * indices, fp64 or fp32 values).  The reference has no CPU implementation of this
 * path; the functions below restate the *semantics* of its GPU pipeline so that
 * the CUDA product can be checked bit-for-bit on structure and to tolerance on
 * values.  Each function cites the reference file:line whose result it restates
 * (paths relative to the reference tree root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path never links it.
 *
 * Parity pinning: the reference ships no golden vectors.
 * This oracle is pinned against (i) fixtures recorded from the reference's own
 * kernels rebuilt for sm_100 and run on a B200 (tests/golden/ref_*.json, written
 * by tests/golden/make_golden.py through oracle/_ref), and (ii) scipy's
 * independent CSR product (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE_SHIFT 5  /* inc/common.h:74  BLOCK_SIZE_BIT 5 */
#define TILE_MASK 31  /* inc/common.h:75  BLOCK_SIZE 32    */

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Total intermediate products: sum over nz (i,k) of A of rownnz_B(k).
 * Restates src/main.cu:102-107 (the GFLOPS numerator). */
int64_t orc_intprod(int M, const int *Ap, const int *Ac, const int *Bp)
{
    int64_t total = 0;
    int64_t nnzA = Ap[M];
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (int64_t j = 0; j < nnzA; ++j)
    {
        int k = Ac[j];
        total += (int64_t)(Bp[k + 1] - Bp[k]);
    }
    return total;
}

/* Per-row intermediate products (the quantity k_calculate_flop_tmp leaves in
 * C.d_tileptr, inc/Form_mask_matrix_B.cuh:56-95). */
void orc_row_intprod(int M, const int *Ap, const int *Ac, const int *Bp, int64_t *out)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < M; ++i)
    {
        int64_t s = 0;
        for (int j = Ap[i]; j < Ap[i + 1]; ++j)
        {
            int k = Ac[j];
            s += (int64_t)(Bp[k + 1] - Bp[k]);
        }
        out[i] = s;
    }
}

/* B mask matrix, pass 1: number of distinct 32-column tiles per row of B,
 * returned as exclusive offsets tileptr[0..K] (tileptr[K] = total tiles).
 * Restates Calculate_B_tilePtr + the in-place scan (inc/MH_spgemm.cuh:45-99,269;
 * kernels inc/Form_mask_matrix_B.cuh:97-388).  Does not assume sorted rows. */
int64_t orc_mask_count(int K, int N, const int *Bp, const int *Bc, int *tileptr)
{
    int nt = (N + TILE_MASK) >> TILE_SHIFT;
    int *cnt = (int *)calloc((size_t)K + 1, sizeof(int));
#pragma omp parallel
    {
        /* stamp[t] == k+1 marks tile t as already seen for row k */
        int *stamp = (int *)calloc((size_t)nt + 1, sizeof(int));
#pragma omp for schedule(dynamic, 1024)
        for (int k = 0; k < K; ++k)
        {
            int c = 0;
            for (int j = Bp[k]; j < Bp[k + 1]; ++j)
            {
                int t = Bc[j] >> TILE_SHIFT;
                if (stamp[t] != k + 1)
                {
                    stamp[t] = k + 1;
                    ++c;
                }
            }
            cnt[k] = c;
        }
        free(stamp);
    }
    int64_t run = 0;
    for (int k = 0; k < K; ++k)
    {
        tileptr[k] = (int)run;
        run += cnt[k];
    }
    tileptr[K] = (int)run;
    free(cnt);
    return run;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* B mask matrix, pass 2: (tilecol, 32-bit occupancy mask) per tile; tile = col>>5,
 * bit = col&31 (inc/Form_mask_matrix_B.cuh:409-415).  The reference emits the tiles
 * of a row in hash-slot order (unordered); the oracle emits them in
 * ascending tilecol order, which is also what the product emits -- comparisons
 * against the reference must be per-row set comparisons. */
void orc_mask_fill(int K, int N, const int *Bp, const int *Bc, const int *tileptr,
                   int *tilecol, uint32_t *tilemask)
{
    int nt = (N + TILE_MASK) >> TILE_SHIFT;
#pragma omp parallel
    {
        uint32_t *acc = (uint32_t *)calloc((size_t)nt + 1, sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1024)
        for (int k = 0; k < K; ++k)
        {
            int base = tileptr[k], c = 0;
            for (int j = Bp[k]; j < Bp[k + 1]; ++j)
            {
                int t = Bc[j] >> TILE_SHIFT;
                if (acc[t] == 0)
                    tilecol[base + c++] = t;
                acc[t] |= (uint32_t)1u << (Bc[j] & TILE_MASK);
            }
            qsort(tilecol + base, (size_t)c, sizeof(int), cmp_int);
            for (int q = 0; q < c; ++q)
            {
                int t = tilecol[base + q];
                tilemask[base + q] = acc[t];
                acc[t] = 0;
            }
        }
        free(acc);
    }
}

/* Tile-flop per A row: sum over A(i,:) of tilecount_B[k].
 * Restates k_calculate_flop (inc/Form_mask_matrix_B.cuh:14-54). */
void orc_row_tileflop(int M, const int *Ap, const int *Ac, const int *tileptr, int64_t *out)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < M; ++i)
    {
        int64_t s = 0;
        for (int j = Ap[i]; j < Ap[i + 1]; ++j)
        {
            int k = Ac[j];
            s += (int64_t)(tileptr[k + 1] - tileptr[k]);
        }
        out[i] = s;
    }
}

/* Symbolic phase, column formulation: nnz of every C row, then exclusive scan.
 * Result equals what Calculate_C_nnz + the scan at src/main.cu:55-57 leave in
 * C.d_ptr (row offsets, Cp[M] = nnz(C)).  Structural product: numerical
 * cancellation never removes an entry (inc/numeric.cuh:237-241 accumulates
 * without testing for zero).  Cp is int64 so overflow of the int32 contract is
 * detectable by the caller. */
int orc_symbolic(int M, int K, int N, const int *Ap, const int *Ac,
                 const int *Bp, const int *Bc, int64_t *Cp)
{
    (void)K;
    int64_t *cnt = (int64_t *)calloc((size_t)M + 1, sizeof(int64_t));
    if (!cnt)
        return -1;
#pragma omp parallel
    {
        int *stamp = (int *)calloc((size_t)N + 1, sizeof(int));
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < M; ++i)
        {
            int64_t c = 0;
            for (int j = Ap[i]; j < Ap[i + 1]; ++j)
            {
                int k = Ac[j];
                for (int q = Bp[k]; q < Bp[k + 1]; ++q)
                {
                    int col = Bc[q];
                    if (stamp[col] != i + 1)
                    {
                        stamp[col] = i + 1;
                        ++c;
                    }
                }
            }
            cnt[i] = c;
        }
        free(stamp);
    }
    int64_t run = 0;
    for (int i = 0; i < M; ++i)
    {
        Cp[i] = run;
        run += cnt[i];
    }
    Cp[M] = run;
    free(cnt);
    return 0;
}

/* Symbolic phase, mask formulation: per C row OR the B tile masks per distinct
 * tile column and popcount.  Restates Calculate_C_tilePtr +
 * Calculate_C_nnz_by_OR_CtileMask (inc/MH_spgemm.cuh:149-240; kernels
 * inc/Calculate_C_nnz.cuh:88-835).  Also returns the distinct C tiles per row
 * (what binning<3> consumes) when ctiles != NULL. */
int orc_symbolic_mask(int M, int N, const int *Ap, const int *Ac, const int *tileptr,
                      const int *tilecol, const uint32_t *tilemask, int64_t *Cp,
                      int64_t *ctiles)
{
    int nt = (N + TILE_MASK) >> TILE_SHIFT;
    int64_t *cnt = (int64_t *)calloc((size_t)M + 1, sizeof(int64_t));
    if (!cnt)
        return -1;
#pragma omp parallel
    {
        uint32_t *acc = (uint32_t *)calloc((size_t)nt + 1, sizeof(uint32_t));
        int *touched = (int *)malloc(((size_t)nt + 1) * sizeof(int));
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < M; ++i)
        {
            int nt_i = 0;
            for (int j = Ap[i]; j < Ap[i + 1]; ++j)
            {
                int k = Ac[j];
                for (int q = tileptr[k]; q < tileptr[k + 1]; ++q)
                {
                    int t = tilecol[q];
                    if (acc[t] == 0)
                        touched[nt_i++] = t;
                    acc[t] |= tilemask[q];
                }
            }
            int64_t c = 0;
            for (int q = 0; q < nt_i; ++q)
            {
                c += __builtin_popcount(acc[touched[q]]);
                acc[touched[q]] = 0;
            }
            cnt[i] = c;
            if (ctiles)
                ctiles[i] = nt_i;
        }
        free(acc);
        free(touched);
    }
    int64_t run = 0;
    for (int i = 0; i < M; ++i)
    {
        Cp[i] = run;
        run += cnt[i];
    }
    Cp[M] = run;
    free(cnt);
    return 0;
}

/* Numeric phase: column indices ascending per row (the rank / bitonic sort of
 * inc/numeric.cuh:287-297,424-456) and accumulated values
 * (inc/numeric.cuh:215-258).  Products are accumulated in A-row order, then
 * B-row order, in the value type -- the GPU accumulates in hash/atomic order, so
 * values agree only to rounding (1e-12 rel fp64, 1e-5 rel fp32). */
#define ORC_NUMERIC(NAME, T)                                                           \
    int NAME(int M, int K, int N, const int *Ap, const int *Ac, const T *Av,           \
             const int *Bp, const int *Bc, const T *Bv, const int64_t *Cp, int *Cc,    \
             T *Cv)                                                                    \
    {                                                                                  \
        (void)K;                                                                       \
        int bad = 0;                                                                   \
        _Pragma("omp parallel")                                                        \
        {                                                                              \
            T *acc = (T *)calloc((size_t)N + 1, sizeof(T));                            \
            int *stamp = (int *)calloc((size_t)N + 1, sizeof(int));                    \
            _Pragma("omp for schedule(dynamic, 256)")                                  \
            for (int i = 0; i < M; ++i)                                                \
            {                                                                          \
                int64_t base = Cp[i];                                                  \
                int64_t n = 0, cap = Cp[i + 1] - Cp[i];                                \
                for (int j = Ap[i]; j < Ap[i + 1]; ++j)                                \
                {                                                                      \
                    int k = Ac[j];                                                     \
                    T a = Av[j];                                                       \
                    for (int q = Bp[k]; q < Bp[k + 1]; ++q)                            \
                    {                                                                  \
                        int col = Bc[q];                                               \
                        if (stamp[col] != i + 1)                                       \
                        {                                                              \
                            stamp[col] = i + 1;                                        \
                            acc[col] = a * Bv[q];                                      \
                            if (n < cap)                                               \
                                Cc[base + n] = col;                                    \
                            ++n;                                                       \
                        }                                                              \
                        else                                                           \
                            acc[col] += a * Bv[q];                                     \
                    }                                                                  \
                }                                                                      \
                if (n != cap)                                                          \
                {                                                                      \
                    _Pragma("omp atomic write") bad = 1;                               \
                    continue;                                                          \
                }                                                                      \
                qsort(Cc + base, (size_t)n, sizeof(int), cmp_int);                     \
                for (int64_t q = 0; q < n; ++q)                                        \
                    Cv[base + q] = acc[Cc[base + q]];                                  \
            }                                                                          \
            free(acc);                                                                 \
            free(stamp);                                                               \
        }                                                                              \
        return bad ? -2 : 0;                                                           \
    }

ORC_NUMERIC(orc_numeric_f64, double)
ORC_NUMERIC(orc_numeric_f32, float)

/* Two-phase host Gustavson SpGEMM, the "host Gustavson reference" of
 * BASELINE.json configs[0] and the cpu_baseline of bench.py: symbolic, then the
 * caller allocates Cc/Cv from Cp[M], then numeric -- the same symbolic-then-numeric
 * contract as src/main.cu:33-66.  Exposed as two calls above; this helper reports
 * nnz(C) only. */
int64_t orc_nnzC(int M, int K, int N, const int *Ap, const int *Ac, const int *Bp,
                 const int *Bc)
{
    int64_t *Cp = (int64_t *)malloc(((size_t)M + 1) * sizeof(int64_t));
    if (!Cp)
        return -1;
    orc_symbolic(M, K, N, Ap, Ac, Bp, Bc, Cp);
    int64_t r = Cp[M];
    free(Cp);
    return r;
}

/* Tolerance comparison of two CSR results, restating CSR::operator==
 * (src/CSR.cu:48-96) with a caller-chosen relative tolerance: ptr and col must be
 * identical, values must satisfy |x-y| <= rtol*max(|x|,|y|) (or both be NaN).
 * Returns the number of mismatches (0 == equal); first_bad gets the first index. */
#include <math.h>
#define ORC_COMPARE(NAME, T)                                                           \
    int64_t NAME(int M, const int64_t *Cp1, const int *Cc1, const T *Cv1,              \
                 const int64_t *Cp2, const int *Cc2, const T *Cv2, double rtol,        \
                 int64_t *first_bad)                                                   \
    {                                                                                  \
        int64_t bad = 0;                                                               \
        *first_bad = -1;                                                               \
        for (int i = 0; i <= M; ++i)                                                   \
            if (Cp1[i] != Cp2[i])                                                      \
            {                                                                          \
                if (*first_bad < 0)                                                    \
                    *first_bad = i;                                                    \
                ++bad;                                                                 \
            }                                                                          \
        if (bad)                                                                       \
            return bad;                                                                \
        int64_t nnz = Cp1[M];                                                          \
        for (int64_t j = 0; j < nnz; ++j)                                              \
        {                                                                              \
            int ok = Cc1[j] == Cc2[j];                                                 \
            double x = (double)Cv1[j], y = (double)Cv2[j];                             \
            double d = fabs(x - y), m = fmax(fabs(x), fabs(y));                        \
            if (!(d <= rtol * m) && !(isnan(x) && isnan(y)))                           \
                ok = 0;                                                                \
            if (!ok)                                                                   \
            {                                                                          \
                if (*first_bad < 0)                                                    \
                    *first_bad = j;                                                    \
                ++bad;                                                                 \
            }                                                                          \
        }                                                                              \
        return bad;                                                                    \
    }

ORC_COMPARE(orc_compare_f64, double)
ORC_COMPARE(orc_compare_f32, float)

/* CSR transpose T = A^T (N x M): count per column, exclusive scan, then append every
 * nonzero to its column's list while walking A in (row, column) order, so that each row of
 * T holds A's row indices ascending.  Restates matrix_transposition (src/utils.cpp:20-46),
 * the host step that forms B for the reference's AAT mode (src/main.cu:98-101).
 * Tp has N+1 entries, Tc / Tv nnz entries.  `vsize` = bytes per value (8 or 4). */
void orc_transpose(int M, int N, const int *Ap, const int *Ac, const void *Av, int vsize, int *Tp, int *Tc,
                   void *Tv)
{
    int64_t nnz = Ap[M];
    memset(Tp, 0, sizeof(int) * ((size_t)N + 1));
    for (int64_t j = 0; j < nnz; ++j)
        Tp[Ac[j] + 1]++;
    for (int c = 0; c < N; ++c)
        Tp[c + 1] += Tp[c];
    int *cursor = (int *)malloc(sizeof(int) * ((size_t)N + 1));
    memcpy(cursor, Tp, sizeof(int) * ((size_t)N + 1));
    for (int r = 0; r < M; ++r)
        for (int j = Ap[r]; j < Ap[r + 1]; ++j)
        {
            int pos = cursor[Ac[j]]++;
            Tc[pos] = r;
            memcpy((char *)Tv + (size_t)pos * vsize, (const char *)Av + (size_t)j * vsize, (size_t)vsize);
        }
    free(cursor);
}
