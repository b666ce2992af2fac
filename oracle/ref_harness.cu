// oracle/ref_harness.cu -- TEST / BASELINE INFRASTRUCTURE ONLY (not the product).
//
// Thin C-ABI wrapper around the UNMODIFIED reference (yyssys/MH-SpGEMM), whose
// sources are compiled in place from /root/reference by oracle/Makefile (`make ref`)
// into oracle/_ref/libmhref.so.  Nothing of the reference is copied into this
// repository; this file only *calls* its public entry points:
//
//   MH_spgemm(const CSR&, CSR&, CSR&, Timing&, Tool&)      src/main.cu:12
//   CSR::alloc / H2D / D2H / d_release_csr / d_release_tile src/CSR.cu
//   Tool::release                                           src/Tool.cu:47
//   cusparse_spgemm(CSR*, CSR*, CSR*, double*)              inc/cusparse_spgemm.cuh:96
//
// Used (a) as the GPU oracle: row_ptr / col_idx of the product must be bit-exact
// with what these kernels produce on the same input, (b) as the "reference arm" of
// bench.py, timed externally with std::chrono because the reference's own
// rdtsc/cpuid(0x16) timer (inc/common.h:115-133) returns inf/NaN on CPUs whose
// leaf 0x16 is empty.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.h"
#include "CSR.h"
#include "Timing.h"
#include "Tool.h"

void MH_spgemm(const CSR &A, CSR &B, CSR &C, Timing &Timing, Tool &tools);
void cusparse_spgemm(CSR *a, CSR *b, CSR *c, double *time);
void warm_gpu();
void matrix_transposition(CSR const &A, CSR &B); // src/utils.cpp:20 (host code, no GPU needed)

namespace
{
using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0)
{
    return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}
double median(std::vector<double> v)
{
    if (v.empty())
        return 0.0;
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}
void fill_csr(CSR &X, int M, int N, const int *p, const int *c, const double *v)
{
    int nnz = p[M];
    X.alloc(M, N, nnz);
    std::memcpy(X.ptr, p, sizeof(int) * (size_t)(M + 1));
    std::memcpy(X.col, c, sizeof(int) * (size_t)nnz);
    std::memcpy(X.val, v, sizeof(double) * (size_t)nnz);
}
void drop_device(CSR &X)
{
    if (X.d_ptr || X.d_col || X.d_val)
        X.d_release_csr();
}
} // namespace

extern "C"
{

    void mhref_free(void *p) { std::free(p); }

    // The reference's own host transpose (src/utils.cpp:20-46), the B operand of its AAT mode.
    // Runs without a GPU.  Tp[N+1], Tc[nnz], Tv[nnz] are caller arrays.
    int mhref_transpose(int M, int N, const int *Ap, const int *Ac, const double *Av, int *Tp, int *Tc, double *Tv)
    {
        // ~CSR calls cudaFree (src/CSR.cu:14-22), which throws on a box without a GPU: the two
        // object shells are therefore heap-allocated and only their host arrays are released
        CSR *A = new CSR, *T = new CSR;
        fill_csr(*A, M, N, Ap, Ac, Av);
        matrix_transposition(*A, *T);
        std::memcpy(Tp, T->ptr, sizeof(int) * (size_t)(N + 1));
        std::memcpy(Tc, T->col, sizeof(int) * (size_t)A->nnz);
        std::memcpy(Tv, T->val, sizeof(double) * (size_t)A->nnz);
        A->h_release_csr();
        T->h_release_csr();
        return 0;
    }

    // One C = A*B through the reference.  Inputs are HOST CSR arrays.
    //   reps/warmup : timed / untimed repetitions of the device-resident call
    //   Cp_out[M+1], *Cc_out, *Cv_out (malloc'd, free with mhref_free), *nnzC_out
    //   ms_device   : median ms of MH_spgemm alone (A,B resident; includes the
    //                 reference's per-call allocations and mask build -- end to end)
    //   ms_e2e      : median ms of CSR::H2D(A,B) + MH_spgemm + CSR::D2H(C)
    //   stage_ms[7] : the reference's own Timing fields of the last call, in the order
    //                 mem_alloc, Form_mask_matrix_B, symbolic_binning, Calculate_C_nnz,
    //                 Malloc_C_col_val, numeric_binning, Numeric (may be NaN, see above)
    //   tile outputs (optional, may be NULL): B's mask matrix as left attached to B.
    int mhref_spgemm(int M, int K, int N, const int *Ap, const int *Ac, const double *Av,
                     const int *Bp, const int *Bc, const double *Bv, int reps, int warmup,
                     int e2e_reps, int *Cp_out, int **Cc_out, double **Cv_out, int *nnzC_out,
                     double *ms_device, double *ms_e2e, double *stage_ms,
                     int *tileptr_out, int **tilecol_out, unsigned **tilemask_out, double *ms_device_min)
    {
        try
        {
            CSR A, B;
            fill_csr(A, M, K, Ap, Ac, Av);
            fill_csr(B, K, N, Bp, Bc, Bv);
            warm_gpu();
            A.H2D();
            B.H2D();
            std::vector<double> t_dev, t_e2e;
            Timing timing;
            int total = warmup + (reps < 1 ? 1 : reps);
            for (int it = 0; it < total; ++it)
            {
                CSR C;
                Tool tools;
                CHECK_ERROR(cudaDeviceSynchronize());
                auto t0 = clk::now();
                MH_spgemm(A, B, C, timing, tools);
                CHECK_ERROR(cudaDeviceSynchronize());
                double ms = ms_since(t0);
                if (it >= warmup)
                    t_dev.push_back(ms);
                bool last = (it == total - 1);
                if (last)
                {
                    C.D2H();
                    *nnzC_out = C.nnz;
                    std::memcpy(Cp_out, C.ptr, sizeof(int) * (size_t)(M + 1));
                    *Cc_out = (int *)std::malloc(sizeof(int) * (size_t)std::max(C.nnz, 1));
                    *Cv_out = (double *)std::malloc(sizeof(double) * (size_t)std::max(C.nnz, 1));
                    std::memcpy(*Cc_out, C.col, sizeof(int) * (size_t)C.nnz);
                    std::memcpy(*Cv_out, C.val, sizeof(double) * (size_t)C.nnz);
                    if (tileptr_out && tilecol_out && tilemask_out)
                    {
                        CHECK_ERROR(cudaMemcpy(tileptr_out, B.d_tileptr, sizeof(int) * (size_t)(K + 1),
                                               cudaMemcpyDeviceToHost));
                        int nt = tileptr_out[K];
                        *tilecol_out = (int *)std::malloc(sizeof(int) * (size_t)std::max(nt, 1));
                        *tilemask_out = (unsigned *)std::malloc(sizeof(unsigned) * (size_t)std::max(nt, 1));
                        CHECK_ERROR(cudaMemcpy(*tilecol_out, B.d_tilecol, sizeof(int) * (size_t)nt,
                                               cudaMemcpyDeviceToHost));
                        CHECK_ERROR(cudaMemcpy(*tilemask_out, B.d_tilemask, sizeof(unsigned) * (size_t)nt,
                                               cudaMemcpyDeviceToHost));
                    }
                }
                // what main() does between iterations (src/main.cu:126-131) plus the
                // C.d_tileptr the reference leaks
                CHECK_ERROR(cudaFree(C.d_tileptr));
                C.d_tileptr = nullptr;
                tools.release();
                B.d_release_tile();
                // ~CSR releases C's host+device arrays
            }
            if (stage_ms)
            {
                stage_ms[0] = timing.mem_alloc;
                stage_ms[1] = timing.Form_mask_matrix_B;
                stage_ms[2] = timing.symbolic_binning;
                stage_ms[3] = timing.Calculate_C_nnz;
                stage_ms[4] = timing.Malloc_C_col_val;
                stage_ms[5] = timing.numeric_binning;
                stage_ms[6] = timing.Numeric;
            }
            drop_device(A);
            drop_device(B);
            // end-to-end: host buffers in, host buffers out, through the reference's
            // own CSR::H2D / MH_spgemm / CSR::D2H
            for (int it = 0; it < e2e_reps; ++it)
            {
                CSR C;
                Tool tools;
                Timing tm;
                CHECK_ERROR(cudaDeviceSynchronize());
                auto t0 = clk::now();
                A.H2D();
                B.H2D();
                MH_spgemm(A, B, C, tm, tools);
                C.D2H();
                CHECK_ERROR(cudaDeviceSynchronize());
                t_e2e.push_back(ms_since(t0));
                CHECK_ERROR(cudaFree(C.d_tileptr));
                C.d_tileptr = nullptr;
                tools.release();
                B.d_release_tile();
                drop_device(A);
                drop_device(B);
            }
            *ms_device = median(t_dev);
            *ms_e2e = median(t_e2e);
            if (ms_device_min) // best repetition: the reference's host path (mallocs, syncs) is noisy
                *ms_device_min = *std::min_element(t_dev.begin(), t_dev.end());
            return 0;
        }
        catch (const std::exception &e)
        {
            std::fprintf(stderr, "mhref_spgemm: reference threw: %s\n", e.what());
            return -1;
        }
    }

    // cuSPARSE SpGEMM exactly as the reference drives it (inc/cusparse_spgemm.cuh:6-94:
    // CUSPARSE_SPGEMM_DEFAULT, fp64, 32-bit indices, timed including its mallocs).
    // Timed externally for the same reason as above.
    int mhref_cusparse(int M, int K, int N, const int *Ap, const int *Ac, const double *Av,
                       const int *Bp, const int *Bc, const double *Bv, int reps, int warmup,
                       int *Cp_out, int **Cc_out, double **Cv_out, int *nnzC_out, double *ms_device,
                       double *ms_device_min)
    {
        try
        {
            CSR A, B;
            fill_csr(A, M, K, Ap, Ac, Av);
            fill_csr(B, K, N, Bp, Bc, Bv);
            A.H2D();
            B.H2D();
            std::vector<double> t_dev;
            int total = warmup + (reps < 1 ? 1 : reps);
            for (int it = 0; it < total; ++it)
            {
                CSR C;
                double t_unused = 0;
                CHECK_ERROR(cudaDeviceSynchronize());
                auto t0 = clk::now();
                cusparse_spgemm(&A, &B, &C, &t_unused);
                CHECK_ERROR(cudaDeviceSynchronize());
                double ms = ms_since(t0);
                if (it >= warmup)
                    t_dev.push_back(ms);
                if (it == total - 1)
                {
                    C.D2H();
                    *nnzC_out = C.nnz;
                    std::memcpy(Cp_out, C.ptr, sizeof(int) * (size_t)(M + 1));
                    *Cc_out = (int *)std::malloc(sizeof(int) * (size_t)std::max(C.nnz, 1));
                    *Cv_out = (double *)std::malloc(sizeof(double) * (size_t)std::max(C.nnz, 1));
                    std::memcpy(*Cc_out, C.col, sizeof(int) * (size_t)C.nnz);
                    std::memcpy(*Cv_out, C.val, sizeof(double) * (size_t)C.nnz);
                }
            }
            *ms_device = median(t_dev);
            if (ms_device_min)
                *ms_device_min = *std::min_element(t_dev.begin(), t_dev.end());
            return 0;
        }
        catch (const std::exception &e)
        {
            std::fprintf(stderr, "mhref_cusparse: threw: %s\n", e.what());
            return -1;
        }
    }

} // extern "C"
