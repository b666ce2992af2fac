// mhb_compat.hpp -- C++ source-compatibility shim: the reference's host API for the hot
// path, implemented on top of the C ABI in mhb_spgemm.h.
//
// A driver written against the reference (its main(), src/main.cu:74-217) keeps compiling
// when it includes this header instead of inc/CSR.h, inc/Tool.h and inc/Timing.h and links
// libmhb_spgemm.so:  the class names, public members and the entry point
//
//     void MH_spgemm(const CSR &A, CSR &B, CSR &C, Timing &Timing, Tool &tools);   // src/main.cu:12
//
// are the reference's.  Differences a maintainer should know:
//   * Tool wraps one mhb handle (workspace + stream) that is reused across calls;
//     Tool::release() destroys it (src/Tool.cu:47-69).  Tool::allocate() is a no-op kept for
//     source compatibility -- the workspace grows inside MH_spgemm.
//   * B's mask matrix is workspace-owned: B.d_tileptr / d_tilecol / d_tilemask and
//     C.d_tileptr stay null, so the reference's clean-up calls (B.d_release_tile(),
//     cudaFree(C.d_tileptr), src/main.cu:128-131) remain valid no-ops.
//   * C.d_ptr / d_col / d_val are cudaMalloc'd here and owned by the caller, exactly as
//     before (CSR::d_release_csr frees them, src/CSR.cu:14-22).
//   * Failures throw std::exception after printing the CUDA / library message, like
//     CHECK_ERROR (inc/common.h:85-95).  The "C.nnz = ..." print (src/main.cu:58) is off
//     unless Tool::verbose is set.
#pragma once
#include <exception>

#include "mhb_spgemm.h"

#ifndef VALUE_TYPE
#define VALUE_TYPE double // inc/common.h:8
#endif
#ifndef MASK_TYPE
#define MASK_TYPE unsigned int // inc/common.h:10
#endif

class CSR
{
  public:
    // dimensions and host arrays (inc/CSR.h:7-13)
    int M = 0, N = 0, nnz = 0;
    int *ptr = nullptr, *col = nullptr;
    VALUE_TYPE *val = nullptr;
    // device twins (inc/CSR.h:15-17)
    int *d_ptr = nullptr, *d_col = nullptr;
    VALUE_TYPE *d_val = nullptr;
    int isSymmetric = 0;
    // mask-matrix fields of the reference (inc/CSR.h:21-27); unused by this implementation
    int *tileptr = nullptr, *tilecol = nullptr;
    MASK_TYPE *tilemask = nullptr;
    int *d_tileptr = nullptr, *d_tilecol = nullptr;
    MASK_TYPE *d_tilemask = nullptr;

    CSR() = default;
    ~CSR();
    void alloc(int rows, int cols, int nonzeros); // host arrays, ptr zero-filled
    CSR &operator=(const CSR &other);            // deep copy of the host side
    bool operator==(const CSR &other);           // ptr/col exact, val 1e-9 abs-or-rel (src/CSR.cu:48-96)
    void H2D();                                  // cudaMalloc + copy of ptr/col/val
    void D2H();                                  // new[] + copy back
    void h_release_csr();
    void d_release_csr();
    void d_release_tile();
    void release();
};

class Timing
{
  public:
    // per-stage milliseconds, the reference's field names (inc/Timing.h:6-12)
    double mem_alloc = 0, Form_mask_matrix_B = 0, Calculate_C_nnz = 0, Malloc_C_col_val = 0, Numeric = 0,
           symbolic_binning = 0, numeric_binning = 0;
    void operator+=(const Timing &t);
    void operator/=(const double x);
    void print_step_time();
    double getTotal(); // the reference's convention: everything but the mask build (src/Timing.cpp:39-42)
};

class Tool
{
  public:
    mhb_handle_t handle = nullptr;
    int verbose = 0;
    Tool() = default;
    ~Tool(); // like the reference's, does not release (src/Tool.cu:70-73)
    void allocate(const CSR &B, const CSR &C);
    void release();
};

void MH_spgemm(const CSR &A, CSR &B, CSR &C, Timing &Timing, Tool &tools);
