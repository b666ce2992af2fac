// A driver in the style of the reference's main() (src/main.cu:98-133): build A, B = A,
// H2D, MH_spgemm(A, B, C, timing, tools), D2H, compare -- compiled against
// include/mhb_compat.hpp and linked with libmhb_spgemm.so.  The comparison target is a
// tiny host Gustavson inside this file, handed through CSR::operator==.
#include <cstdio>
#include <map>
#include <vector>

#include "mhb_compat.hpp"

static void poisson(CSR &A, int n)
{
    std::vector<std::map<int, double>> rows(n * n);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
        {
            int i = y * n + x;
            rows[i][i] = 4.0;
            if (x > 0) rows[i][i - 1] = -1.0;
            if (x < n - 1) rows[i][i + 1] = -1.0;
            if (y > 0) rows[i][i - n] = -1.0;
            if (y < n - 1) rows[i][i + n] = -1.0;
        }
    int nnz = 0;
    for (auto &r : rows) nnz += (int)r.size();
    A.alloc(n * n, n * n, nnz);
    int p = 0;
    for (int i = 0; i < n * n; ++i)
    {
        A.ptr[i] = p;
        for (auto &kv : rows[i]) A.col[p] = kv.first, A.val[p] = kv.second, ++p;
    }
    A.ptr[n * n] = p;
}

static void host_product(const CSR &A, const CSR &B, CSR &C)
{
    std::vector<std::map<int, double>> rows(A.M);
    int nnz = 0;
    for (int i = 0; i < A.M; ++i)
    {
        for (int j = A.ptr[i]; j < A.ptr[i + 1]; ++j)
            for (int q = B.ptr[A.col[j]]; q < B.ptr[A.col[j] + 1]; ++q)
                rows[i][B.col[q]] += A.val[j] * B.val[q];
        nnz += (int)rows[i].size();
    }
    C.alloc(A.M, B.N, nnz);
    int p = 0;
    for (int i = 0; i < A.M; ++i)
    {
        C.ptr[i] = p;
        for (auto &kv : rows[i]) C.col[p] = kv.first, C.val[p] = kv.second, ++p;
    }
    C.ptr[A.M] = p;
}

int main()
{
    CSR A, B, C, want;
    poisson(A, 48);
    B = A;
    A.H2D();
    B.H2D();
    Timing timing, bench_timing;
    Tool tools;
    try
    {
        for (int it = 0; it < 3; ++it)
        {
            MH_spgemm(A, B, C, timing, tools);
            bench_timing += timing;
            if (it < 2)
                C.d_release_csr(), B.d_release_tile();
        }
        bench_timing /= 3;
        C.D2H();
        host_product(A, B, want);
        bool same = (C == want);
        std::printf("C.nnz = %d total %.3f ms\n", C.nnz, bench_timing.getTotal());
        std::printf(same ? "pass\n" : "error\n");
        tools.release();
        B.d_release_tile();
        return same ? 0 : 1;
    }
    catch (const std::exception &e)
    {
        std::printf("MH-SpGEMM failed!!! %s\n", e.what());
        return 2;
    }
}
