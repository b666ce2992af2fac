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
    MHB_ERR_NOMEM = 4,     /* device or pinned allocation failed */
    MHB_ERR_CAPACITY = 5   /* mhb_spgemm_into_*: nnz(C) exceeds the caller's C.col / C.val capacity */
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
    int sym_bin_size[24];    /* rows per symbolic bin  */
    int num_bin_size[24];    /* rows per numeric bin   */
    int gpu_launches;        /* kernels launched by the last call */
    /* option "count_probes": probes of the hash kernels that found their slot taken by another
     * key -- the reference's HASH_CONFLICT counter (inc/common.h:18, inc/numeric.cuh:116-118,
     * inc/Calculate_C_nnz.cuh:153); 0 when the option is off */
    long long hash_probes;     /* numeric phase (column hash)  */
    long long sym_hash_probes; /* symbolic phase (tile hash)   */
    /* symbolic phases on this handle that were launched from the previous call's bin sizes
     * without the mid-pipeline host read (option "speculate"), and how many of them had to be
     * re-run the ordinary way because the guess did not cover the input */
    int speculative_launches;
    int speculative_misses;
    /* mhb_spgemm_into_* calls on this handle that ran with a single host synchronisation */
    int fused_calls;
} mhb_stats;

/* ---- lifetime (replaces Tool::allocate / Tool::release, src/Tool.cu:4-69) ---- */
int mhb_create(mhb_handle_t *out, int device);
int mhb_destroy(mhb_handle_t h);
const char *mhb_last_error(mhb_handle_t h);
/* cudaStream_t to run on; NULL selects the handle-owned non-blocking stream (to run on the
 * default stream pass cudaStreamLegacy or cudaStreamPerThread, not 0). */
int mhb_set_stream(mhb_handle_t h, void *cuda_stream);
/* Tuning / test knobs: "force_sym_path", "force_num_path" (0 auto, 1 window/bitmap only
 * where it fits, 2 hash only), "serial_bins" (1: per-bin kernels on one stream),
 * "nnz_limit" (report MHB_ERR_OVERFLOW above this nnz(C); default INT_MAX), "verbose";
 * kernel selection: "compact_rows" (1), "claim_list" (1), "row_twins" (0), "sym_twins" (1), "pdl" (1:
 * programmatic dependent launch of the main-stream kernel chain; env MHB_PDL overrides at create);
 * "count_probes" (0): count failed hash probes into mhb_stats.hash_probes / sym_hash_probes;
 * "speculate" (1): a call with the shape of the previous one launches its symbolic kernels from
 * that call's bin sizes (one host read per SpGEMM instead of two; verified, re-run on a miss);
 * "mask_onepass" (2): builder of B's mask matrix -- 2: two passes over B.col around a scan of the per-chunk tile
 * counts, 1: one pass with a chained (look-back) scan, 0: the round-1 five-kernel chain;
 * "row_chunks" (2, at most 8) / "row_chunk_bytes" (32 MiB): mhb_spgemm_host_* runs the numeric phase in
 * that many row chunks (sizes 1 : 2 : 4 ... by nnz) when C.col + C.val are at least that large, and
 * downloads a finished chunk while the next one computes (1: off). */
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

/* One-shot MH_spgemm: symbolic, cudaMalloc of C, numeric.  The arrays behind dC_ptr, dC_col and dC_val are
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
/* MH_spgemm into CALLER-OWNED C arrays: dC_ptr[M+1], dC_col / dC_val with room for `capacity`
 * entries.  For the caller that multiplies repeatedly (the timed loop of src/main.cu:118-125
 * re-allocates C every iteration, src/main.cu:59-60; a caller that keeps its buffers has nothing
 * to allocate at the hand-off of src/main.cu:55-60).  When the call has the shape of the previous
 * one on the handle, both phases are launched without any host read in between (bin sizes and
 * nnz(C) stay on the device, the guess is verified by a device-side gate and re-run the ordinary
 * way on a miss): ONE host synchronisation per SpGEMM, at the end.  If nnz(C) > capacity nothing
 * is written to dC_col / dC_val, dC_ptr and *nnzC are valid and the call returns
 * MHB_ERR_CAPACITY: grow the arrays and call again.  mhb_numeric_* may follow (pattern reuse). */
int mhb_spgemm_into_f64(mhb_handle_t h, int M, int K, int N,
                        int nnzA, const int *dA_ptr, const int *dA_col, const double *dA_val,
                        int nnzB, const int *dB_ptr, const int *dB_col, const double *dB_val,
                        int *dC_ptr, int *dC_col, double *dC_val, long long capacity, long long *nnzC);
int mhb_spgemm_into_f32(mhb_handle_t h, int M, int K, int N,
                        int nnzA, const int *dA_ptr, const int *dA_col, const float *dA_val,
                        int nnzB, const int *dB_ptr, const int *dB_col, const float *dB_val,
                        int *dC_ptr, int *dC_col, float *dC_val, long long capacity, long long *nnzC);
/* The same call in two halves, for callers that overlap host work with the SpGEMM or time it with
 * their own events: begin queues the work on the handle's stream and returns without waiting (in
 * steady state; the first call of a shape runs its symbolic phase to completion inside begin),
 * end synchronises, verifies, re-runs on a miss and reports nnz(C) / MHB_ERR_CAPACITY.  One
 * outstanding begin per handle; the arrays must stay valid until end returns. */
int mhb_spgemm_into_begin_f64(mhb_handle_t h, int M, int K, int N,
                              int nnzA, const int *dA_ptr, const int *dA_col, const double *dA_val,
                              int nnzB, const int *dB_ptr, const int *dB_col, const double *dB_val,
                              int *dC_ptr, int *dC_col, double *dC_val, long long capacity);
int mhb_spgemm_into_begin_f32(mhb_handle_t h, int M, int K, int N,
                              int nnzA, const int *dA_ptr, const int *dA_col, const float *dA_val,
                              int nnzB, const int *dB_ptr, const int *dB_col, const float *dB_val,
                              int *dC_ptr, int *dC_col, float *dC_val, long long capacity);
int mhb_spgemm_into_end(mhb_handle_t h, long long *nnzC);
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
                 int *h_bin_offset /* 25 ints */);

/* cudaStream_t the handle currently launches on (for callers that order their own work). */
int mhb_get_stream(mhb_handle_t h, void **cuda_stream);

/* Device-side results of the call last queued on the handle, readable in stream order by the
 * caller's own kernels without a host round trip: nnz(C), and the gate word of a fused call
 * (non-zero: the numeric phase stood down and mhb_spgemm_into_end will redo or report). */
int mhb_get_device_scalars(mhb_handle_t h, const long long **d_nnzC, const int **d_gate);
int mhb_get_timing(mhb_handle_t h, mhb_timing *out);
int mhb_get_stats(mhb_handle_t h, mhb_stats *out);

/* =========================================================================================
 * Row-sharded SpGEMM across the GPUs of one box -- one process (or thread) per GPU, one
 * mhb_shard_t per rank on top of that rank's mhb_handle_t.
 *
 * The reference is single-GPU (no NCCL / MPI / peer access anywhere under src/ or inc/); this
 * is the multi-GPU form of the same call, MH_spgemm (src/main.cu:12-72), for a C++ driver
 * shaped like src/main.cu:74-217 that owns one row block per GPU:
 *   A  rows [a0, a1) on this rank: local CSR (row_ptr rebased to 0), GLOBAL column ids in [0, K)
 *   B  row-sharded by `bounds` (bounds[r] .. bounds[r+1] owned by rank r): local row_ptr
 *      rebased to 0, GLOBAL column ids in [0, N)
 *   C  every rank keeps its own CSR slice (local int32 row_ptr); the global matrix is the
 *      concatenation, row_ptr shifted by the slice offsets (int64) of mhb_shard_offsets().
 * Gustavson rows are independent, so the only exchange is B: a rank needs the rows of B that
 * the columns of its A block reference.  They are read ONE-SIDEDLY over NVLink from windows
 * the owners export with CUDA IPC (peer-mapped memory; no send/recv pairing, no host-side
 * collective in a step): mhb_shard_exchange() = one flag store per peer ("my shard of B is
 * final for this step") + one kernel that waits for its owners' flags and copies the missing
 * pieces into this rank's contiguous image of B.  Slice sizes travel the same way.
 *
 * The library does not own a transport for the set-up: the caller moves fixed-size blobs
 * between the ranks with whatever it has (MPI, torch.distributed, files, shared memory).
 * Set-up sequence (every rank; the all-gathers are the caller's):
 *   mhb_shard_create -> mhb_shard_set_A -> mhb_shard_export(1) -> [all-gather blobs]
 *   -> mhb_shard_import(1) -> mhb_shard_export(2) -> [all-gather] -> mhb_shard_import(2)
 *   -> mhb_shard_own_B: where this rank keeps its shard of B's col / val (inside its image).
 * Per step: [write B's shard] -> mhb_shard_exchange -> mhb_shard_symbolic -> (allocate C)
 *   -> mhb_shard_numeric_* -> mhb_shard_post_size; mhb_shard_offsets when C is assembled.
 * A rank whose block holds more than 2^31-1 products cuts it into row slices
 * [r_lo, r_hi) and calls symbolic / numeric once per slice after ONE exchange.
 * ========================================================================================= */
typedef struct mhb_shard *mhb_shard_t;
#define MHB_SHARD_BLOB_BYTES 128

int mhb_shard_create(mhb_shard_t *out, mhb_handle_t h, int rank, int world, int K, int N,
                     int value_bytes /* 8 or 4 */, const long long *bounds /* world+1 row bounds of B */);
int mhb_shard_destroy(mhb_shard_t s);
const char *mhb_shard_last_error(mhb_shard_t s);
/* This rank's block of A (device arrays, kept by reference) and its shard of B's row_ptr
 * (device, bounds[rank+1]-bounds[rank]+1 entries, rebased to 0; copied). */
int mhb_shard_set_A(mhb_shard_t s, int M_local, int nnzA, const int *dA_ptr, const int *dA_col,
                    const int *dBown_ptr);
int mhb_shard_export(mhb_shard_t s, int phase /* 1 or 2 */, void *blob /* MHB_SHARD_BLOB_BYTES */);
int mhb_shard_import(mhb_shard_t s, int phase, const void *blobs /* world * MHB_SHARD_BLOB_BYTES */);
/* Where this rank's shard of B lives (col: GLOBAL column ids; val: value_bytes each), and the
 * image the SpGEMM reads: rows [*k0, *k1) of B, *nnz_image entries. */
int mhb_shard_own_B(mhb_shard_t s, int **dB_col_own, void **dB_val_own, long long *nnz_own);
int mhb_shard_image(mhb_shard_t s, int *k0, int *k1, long long *nnz_image, long long *halo_bytes_per_step);
/* The exchange step (stream-ordered on the handle's stream, returns without synchronising):
 * mhb_shard_exchange = mhb_shard_publish ("my shard of B is final for this step": one flag store
 * per peer) + mhb_shard_pull (wait for the owners' flags, copy their pieces into the image).
 * The pull kernel SPINS until the owners have published, so the ranks' kernels must be able to
 * run at the same time: one GPU per rank.  Ranks that share one GPU (tests on a single-GPU box)
 * call the two halves separately with a HOST barrier in between, so that no kernel ever waits
 * for a kernel of another process on the same device; likewise they replace mhb_shard_barrier by
 * a host barrier and synchronise before mhb_shard_offsets. */
int mhb_shard_exchange(mhb_shard_t s);
int mhb_shard_publish(mhb_shard_t s);
int mhb_shard_pull(mhb_shard_t s);
/* Device-side barrier over all ranks (peer flags, stream-ordered): call between the end of a
 * step and the next modification of B's shard, and to align ranks before timing. */
int mhb_shard_barrier(mhb_shard_t s);
/* mhb_symbolic / mhb_numeric_* for rows [r_lo, r_hi) of this rank's block against the image. */
int mhb_shard_symbolic(mhb_shard_t s, int r_lo, int r_hi, int *dC_ptr, long long *nnzC);
int mhb_shard_numeric_f64(mhb_shard_t s, const double *dA_val, int *dC_col, double *dC_val);
int mhb_shard_numeric_f32(mhb_shard_t s, const float *dA_val, int *dC_col, float *dC_val);
/* mhb_spgemm_into_* for rows [r_lo, r_hi) of this rank's block against the image: one host
 * synchronisation per step (see mhb_spgemm_into_f64). */
int mhb_shard_spgemm_into_f64(mhb_shard_t s, int r_lo, int r_hi, const double *dA_val, int *dC_ptr, int *dC_col,
                              double *dC_val, long long capacity, long long *nnzC);
int mhb_shard_spgemm_into_f32(mhb_shard_t s, int r_lo, int r_hi, const float *dA_val, int *dC_ptr, int *dC_col,
                              float *dC_val, long long capacity, long long *nnzC);
/* ... and in two halves (mhb_spgemm_into_begin_* / mhb_spgemm_into_end). */
int mhb_shard_spgemm_into_begin_f64(mhb_shard_t s, int r_lo, int r_hi, const double *dA_val, int *dC_ptr, int *dC_col,
                                    double *dC_val, long long capacity);
int mhb_shard_spgemm_into_begin_f32(mhb_shard_t s, int r_lo, int r_hi, const float *dA_val, int *dC_ptr, int *dC_col,
                                    float *dC_val, long long capacity);
int mhb_shard_spgemm_into_end(mhb_shard_t s, long long *nnzC);
/* Publish this rank's nnz(C slice) of the step to every rank (one-sided, stream-ordered);
 * mhb_shard_offsets waits for all of them: offset of this rank's slice and the total. */
int mhb_shard_post_size(mhb_shard_t s, long long nnzC_local);
/* nnzC_local = -1 posts the device-side nnz(C) of the SpGEMM just queued with
 * mhb_shard_spgemm_into_begin_* (no host read in between).  If that call turns out to need a redo
 * (speculation miss) the posted value is a "pending" marker that mhb_shard_offsets waits on;
 * mhb_shard_repost_size(s, nnz) after mhb_shard_spgemm_into_end then supplies the value for the
 * SAME step (idempotent when nothing was pending). */
int mhb_shard_repost_size(mhb_shard_t s, long long nnzC_local);
int mhb_shard_offsets(mhb_shard_t s, long long *slice_offset, long long *nnzC_total,
                      long long *all_sizes /* world entries, may be NULL */);

/* NCCL from C++ (optional; libnccl.so.2 is loaded at run time): the north-star exchange
 * "B is NCCL-broadcast over NVLink" for inputs where every rank needs all of B.
 * mhb_nccl_unique_id on one rank -> [caller distributes the 128 bytes] -> mhb_shard_init_nccl
 * on every rank -> mhb_shard_broadcast(buffer, bytes, root) each step, on the handle's stream. */
int mhb_nccl_unique_id(void *id128);
int mhb_shard_init_nccl(mhb_shard_t s, const void *id128);
int mhb_shard_broadcast(mhb_shard_t s, void *dbuf, size_t bytes, int root);

/* version / build info */
const char *mhb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MHB_SPGEMM_H */
