// mhb_compat.cu -- implementation of include/mhb_compat.hpp (the reference's CSR / Tool /
// Timing / MH_spgemm surface) on top of the C ABI.  Behaviour follows src/CSR.cu,
// src/Tool.cu, src/Timing.cpp and src/main.cu:12-72 of the reference; no code is shared.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "../../include/mhb_compat.hpp"

namespace
{
struct ShimError : std::exception
{
    const char *what() const noexcept override { return "mhb_compat: CUDA / library call failed"; }
};
void cuda_ok(cudaError_t e, const char *what, int line)
{
    if (e == cudaSuccess)
        return;
    std::printf("%s in %s at line %d\n", cudaGetErrorString(e), what, line);
    throw ShimError();
}
#define SHIM_CUDA(x) cuda_ok((x), #x, __LINE__)
} // namespace

// ---- CSR --------------------------------------------------------------------------------
CSR::~CSR() { release(); }

void CSR::alloc(int rows, int cols, int nonzeros)
{
    M = rows, N = cols, nnz = nonzeros;
    ptr = new int[(size_t)rows + 1]();
    col = new int[(size_t)std::max(nonzeros, 1)];
    val = new VALUE_TYPE[(size_t)std::max(nonzeros, 1)];
}

CSR &CSR::operator=(const CSR &o)
{
    if (this == &o)
        return *this;
    h_release_csr();
    alloc(o.M, o.N, o.nnz);
    isSymmetric = o.isSymmetric;
    std::memcpy(ptr, o.ptr, sizeof(int) * ((size_t)M + 1));
    std::memcpy(col, o.col, sizeof(int) * (size_t)nnz);
    std::memcpy(val, o.val, sizeof(VALUE_TYPE) * (size_t)nnz);
    return *this;
}

bool CSR::operator==(const CSR &o)
{
    if (nnz != o.nnz)
    {
        std::printf("nnz not equal %d %d\n", nnz, o.nnz);
        throw std::runtime_error("nnz not equal");
    }
    if (M != o.M || N != o.N)
        throw std::runtime_error("dimension not same");
    const double eps = 1e-9;
    int errors = 0;
    for (int i = 0; i <= M && errors <= 10; ++i)
        if (ptr[i] != o.ptr[i])
        {
            std::printf("ptr not equal at %d rows, %d != %d\n", i, ptr[i], o.ptr[i]);
            ++errors;
        }
    for (int j = 0; j < nnz && errors <= 10; ++j)
    {
        if (col[j] != o.col[j])
        {
            std::printf("col not equal, index %d != %d\n", col[j], o.col[j]);
            ++errors;
        }
        const double d = std::fabs((double)val[j] - (double)o.val[j]);
        if (!(d < eps || d < eps * std::fabs((double)val[j])))
        {
            std::printf("val not equal, value %.18le != %.18le\n", (double)val[j], (double)o.val[j]);
            ++errors;
        }
    }
    if (errors > 10)
        throw std::runtime_error("matrix compare: error num exceed threshold");
    return errors == 0;
}

void CSR::H2D()
{
    SHIM_CUDA(cudaMalloc((void **)&d_ptr, sizeof(int) * ((size_t)M + 1)));
    SHIM_CUDA(cudaMalloc((void **)&d_col, sizeof(int) * (size_t)std::max(nnz, 1)));
    SHIM_CUDA(cudaMalloc((void **)&d_val, sizeof(VALUE_TYPE) * (size_t)std::max(nnz, 1)));
    SHIM_CUDA(cudaMemcpy(d_ptr, ptr, sizeof(int) * ((size_t)M + 1), cudaMemcpyHostToDevice));
    SHIM_CUDA(cudaMemcpy(d_col, col, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
    SHIM_CUDA(cudaMemcpy(d_val, val, sizeof(VALUE_TYPE) * (size_t)nnz, cudaMemcpyHostToDevice));
}

void CSR::D2H()
{
    h_release_csr();
    ptr = new int[(size_t)M + 1];
    col = new int[(size_t)std::max(nnz, 1)];
    val = new VALUE_TYPE[(size_t)std::max(nnz, 1)];
    SHIM_CUDA(cudaMemcpy(ptr, d_ptr, sizeof(int) * ((size_t)M + 1), cudaMemcpyDeviceToHost));
    SHIM_CUDA(cudaMemcpy(col, d_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
    SHIM_CUDA(cudaMemcpy(val, d_val, sizeof(VALUE_TYPE) * (size_t)nnz, cudaMemcpyDeviceToHost));
}

void CSR::h_release_csr()
{
    delete[] ptr;
    delete[] col;
    delete[] val;
    ptr = col = nullptr;
    val = nullptr;
}

void CSR::d_release_csr()
{
    SHIM_CUDA(cudaFree(d_ptr));
    SHIM_CUDA(cudaFree(d_col));
    SHIM_CUDA(cudaFree(d_val));
    d_ptr = d_col = nullptr;
    d_val = nullptr;
}

void CSR::d_release_tile()
{
    SHIM_CUDA(cudaFree(d_tileptr));
    SHIM_CUDA(cudaFree(d_tilecol));
    SHIM_CUDA(cudaFree(d_tilemask));
    d_tileptr = d_tilecol = nullptr;
    d_tilemask = nullptr;
}

void CSR::release()
{
    h_release_csr();
    d_release_csr();
}

// ---- Timing -----------------------------------------------------------------------------
void Timing::operator+=(const Timing &t)
{
    mem_alloc += t.mem_alloc, Form_mask_matrix_B += t.Form_mask_matrix_B, Calculate_C_nnz += t.Calculate_C_nnz;
    Malloc_C_col_val += t.Malloc_C_col_val, Numeric += t.Numeric, symbolic_binning += t.symbolic_binning;
    numeric_binning += t.numeric_binning;
}
void Timing::operator/=(const double x)
{
    for (double *f : {&mem_alloc, &Form_mask_matrix_B, &Calculate_C_nnz, &Malloc_C_col_val, &Numeric,
                      &symbolic_binning, &numeric_binning})
        *f /= x;
}
void Timing::print_step_time()
{
    std::printf("  -------------time-------------\n");
    std::printf("    mem_alloc: \t\t%.3lfms\n    form_mask_matrix_B: %.3lfms\n    symbolic_binning: \t%.3lfms\n"
                "    calculate_C_nnz: \t%.3lfms\n    malloc_C_col_val: \t%.3lfms\n    numeric_binning: \t%.3lfms\n"
                "    numeric: \t\t%.3lfms\n",
                mem_alloc, Form_mask_matrix_B, symbolic_binning, Calculate_C_nnz, Malloc_C_col_val,
                numeric_binning, Numeric);
    std::printf("  ------------------------------\n");
}
double Timing::getTotal()
{
    return mem_alloc + symbolic_binning + Calculate_C_nnz + Malloc_C_col_val + numeric_binning + Numeric;
}

// ---- Tool -------------------------------------------------------------------------------
Tool::~Tool() {}
void Tool::allocate(const CSR &, const CSR &)
{
    if (!handle && mhb_create(&handle, 0) != MHB_OK)
    {
        std::printf("mhb_create failed: no usable CUDA device\n");
        throw ShimError();
    }
}
void Tool::release()
{
    if (handle)
        mhb_destroy(handle);
    handle = nullptr;
}

// ---- the entry point --------------------------------------------------------------------
void MH_spgemm(const CSR &A, CSR &B, CSR &C, Timing &tm, Tool &tools)
{
    C.M = A.M;
    C.N = B.N;
    tools.allocate(B, C);
    mhb_handle_t h = tools.handle;
    mhb_set_option(h, "verbose", tools.verbose);
    auto fail = [&]() {
        std::printf("%s\n", mhb_last_error(h));
        throw ShimError();
    };
    SHIM_CUDA(cudaMalloc((void **)&C.d_ptr, sizeof(int) * ((size_t)C.M + 1)));
    long long nnzC = 0;
    if (mhb_symbolic(h, A.M, A.N, B.N, A.nnz, A.d_ptr, A.d_col, B.nnz, B.d_ptr, B.d_col, C.d_ptr, &nnzC) != MHB_OK)
        fail();
    C.nnz = (int)nnzC;
    // the hand-off of src/main.cu:55-60: the caller-visible allocation of C.col / C.val
    SHIM_CUDA(cudaMalloc((void **)&C.d_col, sizeof(int) * (size_t)std::max(C.nnz, 1)));
    SHIM_CUDA(cudaMalloc((void **)&C.d_val, sizeof(VALUE_TYPE) * (size_t)std::max(C.nnz, 1)));
    int rc;
    if (sizeof(VALUE_TYPE) == 8)
        rc = mhb_numeric_f64(h, (const double *)A.d_val, (const double *)B.d_val, C.d_col, (double *)C.d_val);
    else
        rc = mhb_numeric_f32(h, (const float *)A.d_val, (const float *)B.d_val, C.d_col, (float *)C.d_val);
    if (rc != MHB_OK)
        fail();
    mhb_timing t;
    mhb_get_timing(h, &t);
    tm.mem_alloc = t.mem_alloc;
    tm.Form_mask_matrix_B = t.form_mask_matrix_B;
    tm.symbolic_binning = t.symbolic_binning;
    tm.Calculate_C_nnz = t.calculate_C_nnz;
    tm.Malloc_C_col_val = t.malloc_C_col_val;
    tm.numeric_binning = t.numeric_binning;
    tm.Numeric = t.numeric;
}
