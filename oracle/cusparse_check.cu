// oracle/cusparse_check.cu -- TEST INFRASTRUCTURE ONLY (not the product).
//
// Independent third oracle: cusparseSpGEMM through the generic API with a
// selectable algorithm.  The reference drives cuSPARSE with CUSPARSE_SPGEMM_DEFAULT only
// (inc/cusparse_spgemm.cuh:46-77), which runs out of memory on the largest suite shapes
// (wb-edu); ALG2 / ALG3 bound the work buffers through cusparseSpGEMM_estimateMemory with a
// chunk fraction.  Host CSR in, host CSR out (column order inside a row as cuSPARSE leaves it:
// the caller sorts before comparing), plus the device time of the call.
//
//   alg: 0 = DEFAULT, 1 = ALG1, 2 = ALG2, 3 = ALG3;  chunk_fraction only for ALG2 / ALG3
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cusparse.h>

#define CK(call)                                                                                     \
    do                                                                                               \
    {                                                                                                \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
        {                                                                                            \
            std::fprintf(stderr, "cusparse_check: %s: %s\n", #call, cudaGetErrorString(e__));        \
            rc = -1;                                                                                 \
            goto done;                                                                               \
        }                                                                                            \
    } while (0)
#define CS(call)                                                                                     \
    do                                                                                               \
    {                                                                                                \
        cusparseStatus_t s__ = (call);                                                               \
        if (s__ != CUSPARSE_STATUS_SUCCESS)                                                          \
        {                                                                                            \
            std::fprintf(stderr, "cusparse_check: %s: status %d\n", #call, (int)s__);                \
            rc = (s__ == CUSPARSE_STATUS_INSUFFICIENT_RESOURCES || s__ == CUSPARSE_STATUS_ALLOC_FAILED) ? -2 : -1; \
            goto done;                                                                               \
        }                                                                                            \
    } while (0)

extern "C"
{
    void csk_free(void *p) { std::free(p); }

    // Returns 0, -1 (error) or -2 (out of memory / insufficient resources).
    int csk_spgemm(int M, int K, int N, const int *Ap, const int *Ac, const double *Av, const int *Bp,
                   const int *Bc, const double *Bv, int alg_sel, float chunk_fraction, int *Cp_out,
                   int **Cc_out, double **Cv_out, long long *nnzC_out, double *ms_out)
    {
        int rc = 0;
        const int nnzA = Ap[M], nnzB = Bp[K];
        const cusparseSpGEMMAlg_t alg = alg_sel == 1   ? CUSPARSE_SPGEMM_ALG1
                                        : alg_sel == 2 ? CUSPARSE_SPGEMM_ALG2
                                        : alg_sel == 3 ? CUSPARSE_SPGEMM_ALG3
                                                       : CUSPARSE_SPGEMM_DEFAULT;
        int *dAp = nullptr, *dAc = nullptr, *dBp = nullptr, *dBc = nullptr, *dCp = nullptr, *dCc = nullptr;
        double *dAv = nullptr, *dBv = nullptr, *dCv = nullptr;
        void *buf1 = nullptr, *buf2 = nullptr, *buf3 = nullptr;
        size_t sz1 = 0, sz2 = 0, sz3 = 0;
        cusparseHandle_t handle = nullptr;
        cusparseSpMatDescr_t matA = nullptr, matB = nullptr, matC = nullptr;
        cusparseSpGEMMDescr_t desc = nullptr;
        const double alpha = 1.0, beta = 0.0;
        const cusparseOperation_t op = CUSPARSE_OPERATION_NON_TRANSPOSE;
        int64_t rows = 0, cols = 0, nnzC = 0;
        auto t0 = std::chrono::steady_clock::now();

        CK(cudaMalloc(&dAp, sizeof(int) * (size_t)(M + 1)));
        CK(cudaMalloc(&dAc, sizeof(int) * (size_t)(nnzA > 0 ? nnzA : 1)));
        CK(cudaMalloc(&dAv, sizeof(double) * (size_t)(nnzA > 0 ? nnzA : 1)));
        CK(cudaMalloc(&dBp, sizeof(int) * (size_t)(K + 1)));
        CK(cudaMalloc(&dBc, sizeof(int) * (size_t)(nnzB > 0 ? nnzB : 1)));
        CK(cudaMalloc(&dBv, sizeof(double) * (size_t)(nnzB > 0 ? nnzB : 1)));
        CK(cudaMalloc(&dCp, sizeof(int) * (size_t)(M + 1)));
        CK(cudaMemcpy(dAp, Ap, sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAc, Ac, sizeof(int) * (size_t)nnzA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAv, Av, sizeof(double) * (size_t)nnzA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBp, Bp, sizeof(int) * (size_t)(K + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBc, Bc, sizeof(int) * (size_t)nnzB, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBv, Bv, sizeof(double) * (size_t)nnzB, cudaMemcpyHostToDevice));
        CS(cusparseCreate(&handle));
        CS(cusparseCreateCsr(&matA, M, K, nnzA, dAp, dAc, dAv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                             CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
        CS(cusparseCreateCsr(&matB, K, N, nnzB, dBp, dBc, dBv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                             CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
        CS(cusparseCreateCsr(&matC, M, N, 0, dCp, nullptr, nullptr, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                             CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
        CS(cusparseSpGEMM_createDescr(&desc));
        CK(cudaDeviceSynchronize());
        t0 = std::chrono::steady_clock::now();
        CS(cusparseSpGEMM_workEstimation(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc, &sz1,
                                         nullptr));
        CK(cudaMalloc(&buf1, sz1 > 0 ? sz1 : 1));
        CS(cusparseSpGEMM_workEstimation(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc, &sz1,
                                         buf1));
        if (alg == CUSPARSE_SPGEMM_ALG2 || alg == CUSPARSE_SPGEMM_ALG3)
        {
            CS(cusparseSpGEMM_estimateMemory(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc,
                                             chunk_fraction, &sz3, nullptr, nullptr));
            CK(cudaMalloc(&buf3, sz3 > 0 ? sz3 : 1));
            CS(cusparseSpGEMM_estimateMemory(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc,
                                             chunk_fraction, &sz3, buf3, &sz2));
            CK(cudaFree(buf3));
            buf3 = nullptr;
        }
        else
            CS(cusparseSpGEMM_compute(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc, &sz2,
                                      nullptr));
        CK(cudaMalloc(&buf2, sz2 > 0 ? sz2 : 1));
        CS(cusparseSpGEMM_compute(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc, &sz2, buf2));
        CS(cusparseSpMatGetSize(matC, &rows, &cols, &nnzC));
        CK(cudaMalloc(&dCc, sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1)));
        CK(cudaMalloc(&dCv, sizeof(double) * (size_t)(nnzC > 0 ? nnzC : 1)));
        CS(cusparseCsrSetPointers(matC, dCp, dCc, dCv));
        CS(cusparseSpGEMM_copy(handle, op, op, &alpha, matA, matB, &beta, matC, CUDA_R_64F, alg, desc));
        CK(cudaDeviceSynchronize());
        *ms_out = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *nnzC_out = (long long)nnzC;
        *Cc_out = (int *)std::malloc(sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1));
        *Cv_out = (double *)std::malloc(sizeof(double) * (size_t)(nnzC > 0 ? nnzC : 1));
        CK(cudaMemcpy(Cp_out, dCp, sizeof(int) * (size_t)(M + 1), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(*Cc_out, dCc, sizeof(int) * (size_t)nnzC, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(*Cv_out, dCv, sizeof(double) * (size_t)nnzC, cudaMemcpyDeviceToHost));
    done:
        if (desc)
            cusparseSpGEMM_destroyDescr(desc);
        if (matA)
            cusparseDestroySpMat(matA);
        if (matB)
            cusparseDestroySpMat(matB);
        if (matC)
            cusparseDestroySpMat(matC);
        if (handle)
            cusparseDestroy(handle);
        cudaFree(buf1), cudaFree(buf2), cudaFree(buf3);
        cudaFree(dAp), cudaFree(dAc), cudaFree(dAv), cudaFree(dBp), cudaFree(dBc), cudaFree(dBv);
        cudaFree(dCp), cudaFree(dCc), cudaFree(dCv);
        if (rc != 0)
            cudaGetLastError();
        return rc;
    }
}
