/*
 * mhb_spgemm.h -- C ABI of the B200-native CSR SpGEMM (C = A*B) that replaces the hot
 * path of yyssys/MH-SpGEMM.  Plain C types only; every entry point returns 0 on success
 * and a non-zero status otherwise (never throws); mhb_last_error() gives the message.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference tree).  Layout contract (inc/CSR.h:4-44): 0-based int32 row_ptr / col_idx,
 * rows sorted ascending and duplicate-free, values double (VALUE_TYPE, inc/common.h:8)
 * or float.  All `d*` pointers are DEVICE pointers, all `h*` pointers are HOST pointers.
 *
 * One handle per (host thread, device).  Calls on one handle are stream-ordered on the
 * handle's stream and return after their result is complete unless noted.
 */
#ifndef MHB_SPGEMM_H
#define MHB_SPGEMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mhb_context *mhb_handle_t;

/* status codes */
enum {
    MHB_OK = 0,
    MHB_ERR_CUDA = 1,      /* a CUDA runtime call or kernel failed */
    MHB_ERR_ARG = 2,       /* invalid argument / call order */
    MHB_ERR_OVERFLOW = 3,  /* nnz(C) does not fit the int32 CSR contract: shard the rows */
    MHB_ERR_NOMEM = 4      /* device or pinned allocation failed */
};

/* Per-stage device times of the last call, in ms, named after the reference's Timing
 * fields (inc/Timing.h:6-12) so reports line up side by side. Measured with CUDA events. */
typedef struct mhb_timing {
    double mem_alloc;          /* workspace growth (0 in steady state)              */
    double form_mask_matrix_B; /* family 1: B mask matrix                           */
    double symbolic_binning;   /* family 2: row metrics + symbolic binning          */
    double calculate_C_nnz;    /* family 3: symbolic nnz(C) + scan                  */
    double malloc_C_col_val;   /* nnz hand-off (D2H of nnz, bin sizes)              */
    double numeric_binning;    /* family 2: numeric binning                         */
    double numeric;            /* family 4: numeric                                 */
    double total;              /* first kernel to last kernel, mask build included  */
} mhb_timing;

/* Per-call statistics (replaces the printf side channel of src/main.cu:58,115-116). */
typedef struct mhb_stats {
    long long intprod;       /* sum over nz (i,k) of A of rownnz_B(k)  (src/main.cu:102-107) */
    long long tileflop;      /* same sum over B tile counts (inc/Form_mask_matrix_B.cuh:14-54) */
    long long ntiles_B;      /* tiles in B's mask matrix                                */
    long long nnzC;
    int sym_bin_size[16];    /* rows per symbolic bin  */
    int num_bin_size[16];    /* rows per numeric bin   */
    int gpu_launches;        /* kernels launched by the last call */
    /* option "count_probes": probes of the hash kernels that found their slot taken by another
     * key -- the reference's HASH_CONFLICT counter (inc/common.h:18, inc/numeric.cuh:116-118,
     * inc/Calculate_C_nnz.cuh:153); 0 when the option is off */
    long long hash_probes;     /* numeric phase (column hash)  */
    long long sym_hash_probes; /* symbolic phase (tile hash)   */
} mhb_stats;

/* ---- lifetime (replaces Tool::allocate / Tool::release, src/Tool.cu:4-69) ---- */
int mhb_create(mhb_handle_t *out, int device);
int mhb_destroy(mhb_handle_t h);
const char *mhb_last_error(mhb_handle_t h);
/* cudaStream_t to run on; NULL selects the handle-owned non-blocking stream. */
int mhb_set_stream(mhb_handle_t h, void *cuda_stream);
/* Tuning / test knobs: "force_sym_path", "force_num_path" (0 auto, 1 window/bitmap only
 * where it fits, 2 hash only), "serial_bins" (1: per-bin kernels on one stream),
 * "nnz_limit" (report MHB_ERR_OVERFLOW above this nnz(C); default INT_MAX), "verbose";
 * kernel selection: "compact_rows" (1), "claim_list" (1), "row_twins" (0), "sym_twins" (1), "pdl" (1:
 * programmatic dependent launch of the main-stream kernel chain; env MHB_PDL overrides at create);
 * "count_probes" (0): count failed hash probes into mhb_stats.hash_probes / sym_hash_probes. */
int mhb_set_option(mhb_handle_t h, const char *key, long long value);

/* ---- the symbolic-then-numeric contract (MH_spgemm, src/main.cu:12-72) ---- */

/* Symbolic phase: steps 2-7 of src/main.cu:21-57.  Writes the exclusive-scanned row
 * offsets of C to dC_ptr[0..M] and nnz(C) to *nnzC; the caller then allocates
 * dC_col / dC_val (the hand-off of src/main.cu:55-60).  A and B may alias. */
int mhb_symbolic(mhb_handle_t h, int M, int K, int N,
                 int nnzA, const int *dA_ptr, const int *dA_col,
                 int nnzB, const int *dB_ptr, const int *dB_col,
                 int *dC_ptr, long long *nnzC);

/* Numeric phase: h_numeric (inc/MH_spgemm.cuh:364-430).  Uses the pattern of the last
 * mhb_symbolic on this handle (the A/B/C index arrays must still be valid); may be
 * called repeatedly with new values.  dC_col ascending per row. */
int mhb_numeric_f64(mhb_handle_t h, const double *dA_val, const double *dB_val,
                    int *dC_col, double *dC_val);
int mhb_numeric_f32(mhb_handle_t h, const float *dA_val, const float *dB_val,
                    int *dC_col, float *dC_val);

/* One-shot MH_spgemm: symbolic, cudaMalloc of C, numeric.  *dC_ptr/*dC_col/*dC_val are
 * caller-owned afterwards (cudaFree / mhb_device_free), as CSR::d_release_csr expects
 * (src/CSR.cu:14-22). */
int mhb_spgemm_f64(mhb_handle_t h, int M, int K, int N,
                   int nnzA, const int *dA_ptr, const int *dA_col, const double *dA_val,
                   int nnzB, const int *dB_ptr, const int *dB_col, const double *dB_val,
                   int **dC_ptr, int **dC_col, double **dC_val, long long *nnzC);
int mhb_spgemm_f32(mhb_handle_t h, int M, int K, int N,
                   int nnzA, const int *dA_ptr, const int *dA_col, const float *dA_val,
                   int nnzB, const int *dB_ptr, const int *dB_col, const float *dB_val,
                   int **dC_ptr, int **dC_col, float **dC_val, long long *nnzC);
int mhb_device_free(void *dptr);
/* Raw device buffers and copies for callers without their own CUDA runtime binding: the
 * pieces of CSR::H2D / CSR::D2H (src/CSR.cu:97-120).  Synchronous. */
int mhb_device_alloc(void **dptr, size_t bytes);
int mhb_memcpy_h2d(void *dptr, const void *hptr, size_t bytes);
int mhb_memcpy_d2h(void *hptr, const void *dptr, size_t bytes);

/* ---- host-buffer entry point (CSR::H2D + MH_spgemm + CSR::D2H, src/main.cu:110-124,
 *      src/CSR.cu:97-120).  Inputs are host arrays (pinned memory recommended); if the
 *      B pointers equal the A pointers (C = A*A) B is uploaded once.  The result arrays
 *      are pinned, handle-owned and valid until the next host call or mhb_destroy. ---- */
int mhb_spgemm_host_f64(mhb_handle_t h, int M, int K, int N,
                        const int *hA_ptr, const int *hA_col, const double *hA_val,
                        const int *hB_ptr, const int *hB_col, const double *hB_val,
                        const int **hC_ptr, const int **hC_col, const double **hC_val,
                        long long *nnzC);
int mhb_spgemm_host_f32(mhb_handle_t h, int M, int K, int N,
                        const int *hA_ptr, const int *hA_col, const float *hA_val,
                        const int *hB_ptr, const int *hB_col, const float *hB_val,
                        const int **hC_ptr, const int **hC_col, const float **hC_val,
                        long long *nnzC);
/* ---- device-side CSR transpose: T = A^T (N x M), canonical CSR (rows of T hold A's row
 *      indices ascending).  Replaces matrix_transposition (src/utils.cpp:20-46), which the
 *      reference runs on the host to form B for its AAT mode (src/main.cu:98-101, flag
 *      inc/common.h:37).  dT_ptr has N+1 entries, dT_col / dT_val nnz entries, all caller-owned
 *      device arrays.  Deterministic (stable radix sort by column, no atomic cursors). ---- */
int mhb_transpose_f64(mhb_handle_t h, int M, int N, int nnz, const int *dA_ptr, const int *dA_col,
                      const double *dA_val, int *dT_ptr, int *dT_col, double *dT_val);
int mhb_transpose_f32(mhb_handle_t h, int M, int N, int nnz, const int *dA_ptr, const int *dA_col,
                      const float *dA_val, int *dT_ptr, int *dT_col, float *dT_val);

/* pinned host memory for callers that want zero-staging uploads */
int mhb_host_alloc(void **hptr, size_t bytes);
int mhb_host_free(void *hptr);

/* ---- stage-level entry points (the four kernel families, for parity tests) ---- */

/* Family 1 -- Form_mask_matrix_B (inc/MH_spgemm.cuh:242-295): B's mask matrix as
 * exclusive tile offsets d_tileptr[0..K], tile column (col>>5) and 32-bit occupancy mask
 * (bit col&31) per tile, tiles ascending inside a row.  Arrays are handle-owned. */
int mhb_form_mask_matrix_B(mhb_handle_t h, int K, int N, int nnzB, const int *dB_ptr, const int *dB_col,
                           const int **d_tileptr, const int **d_tilecol,
                           const unsigned **d_tilemask, long long *ntiles);

/* Family 2 -- per-row counts and bins of the last mhb_symbolic (k_calculate_flop,
 * k_calculate_flop_tmp, k_binning1/2: inc/Form_mask_matrix_B.cuh:14-95, inc/binning.cuh).
 * d_row_info[i] = {intprod (saturating), tileflop, min col, max col} of C row i.
 * which = 0: symbolic bins, 1: numeric bins.  d_bins holds row ids grouped by bin in
 * ascending row order inside a bin; h_bin_offset has nbins+1 entries. */
int mhb_get_row_info(mhb_handle_t h, const int **d_row_info /* int4 per row */);
int mhb_get_bins(mhb_handle_t h, int which, int *nbins, const int **d_bins,
                 int *h_bin_offset /* 17 ints */);

int mhb_get_timing(mhb_handle_t h, mhb_timing *out);
int mhb_get_stats(mhb_handle_t h, mhb_stats *out);

/* version / build info */
const char *mhb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MHB_SPGEMM_H */
