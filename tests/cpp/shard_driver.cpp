// A C++ caller of the row-sharded SpGEMM through the C ABI alone (include/mhb_spgemm.h; no CUDA
// runtime, no Python, no MPI): two ranks = two processes forked BEFORE any CUDA call, set-up blobs
// and host barriers through an anonymous shared mapping.  Each rank owns half of the rows of a 2-D
// Poisson matrix (A and B = A sharded alike), runs three steps of
//   exchange -> mhb_shard_spgemm_into_begin_f64 -> mhb_shard_post_size(-1) -> ..._end -> repost
// with B's values changing every step, and checks ITS slice of C = A*B and the slice offsets against a
// host Gustavson in this file.  Ranks may share one GPU: publish / pull are then separated by a host
// barrier, as the header prescribes.  Prints "SHARD-CPP-OK rank r" per rank; exit code 0 when both do.
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "mhb_spgemm.h"

struct Shared
{
    std::atomic<int> arrive[64];
    unsigned char blob[2][2][MHB_SHARD_BLOB_BYTES]; // [phase][rank]
};
static Shared *sh;
static int bar_no = 0;
static void host_barrier()
{
    const int k = bar_no++;
    sh->arrive[k].fetch_add(1);
    while (sh->arrive[k].load() < 2)
        usleep(50);
}
#define OK(call)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        int rc__ = (call);                                                                         \
        if (rc__ != MHB_OK)                                                                        \
        {                                                                                          \
            std::fprintf(stderr, "rank %d: %s -> %d (%s)\n", rank, #call, rc__,                    \
                         s ? mhb_shard_last_error(s) : (h ? mhb_last_error(h) : ""));              \
            std::_Exit(2);                                                                         \
        }                                                                                          \
    } while (0)

template <class T>
static T *to_device(const std::vector<T> &v)
{
    void *d = nullptr;
    if (mhb_device_alloc(&d, v.size() * sizeof(T) + 16) || (v.size() && mhb_memcpy_h2d(d, v.data(), v.size() * sizeof(T))))
        std::_Exit(3);
    return static_cast<T *>(d);
}

static int run_rank(int rank)
{
    mhb_handle_t h = nullptr;
    mhb_shard_t s = nullptr;
    // the matrix: 5-point Poisson on an n x n grid (every rank builds the same one)
    const int n = 40, M = n * n;
    std::vector<int> ptr(M + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < M; ++i)
    {
        const int x = i % n, y = i / n;
        const int nb[5] = {y > 0 ? i - n : -1, x > 0 ? i - 1 : -1, i, x < n - 1 ? i + 1 : -1, y < n - 1 ? i + n : -1};
        for (int c : nb)
            if (c >= 0)
                col.push_back(c), val.push_back(c == i ? 4.0 : -1.0 - 0.001 * (c % 7));
        ptr[i + 1] = (int)col.size();
    }
    const long long bounds[3] = {0, M / 2 + 7, M}; // uneven on purpose
    const int r0 = (int)bounds[rank], r1 = (int)bounds[rank + 1], Ml = r1 - r0;
    std::vector<int> lptr(Ml + 1), lcol(col.begin() + ptr[r0], col.begin() + ptr[r1]);
    std::vector<double> lval(val.begin() + ptr[r0], val.begin() + ptr[r1]);
    for (int i = 0; i <= Ml; ++i)
        lptr[i] = ptr[r0 + i] - ptr[r0];
    int ndev = 0;
    OK(mhb_create(&h, 0));
    (void)ndev;
    int *dAp = to_device(lptr), *dAc = to_device(lcol), *dBp = to_device(lptr);
    double *dAv = to_device(lval);
    OK(mhb_shard_create(&s, h, rank, 2, M, M, 8, bounds));
    OK(mhb_shard_set_A(s, Ml, (int)lcol.size(), dAp, dAc, dBp));
    for (int phase = 1; phase <= 2; ++phase)
    {
        OK(mhb_shard_export(s, phase, sh->blob[phase - 1][rank]));
        host_barrier();
        OK(mhb_shard_import(s, phase, sh->blob[phase - 1]));
        host_barrier();
    }
    int *own_col = nullptr;
    void *own_val = nullptr;
    long long n_own = 0;
    OK(mhb_shard_own_B(s, &own_col, &own_val, &n_own));
    if (n_own != (long long)lcol.size())
        return 4;
    if (mhb_memcpy_h2d(own_col, lcol.data(), lcol.size() * 4))
        return 5;
    // caller-owned C arrays: upper bound 25 entries per row
    const long long cap = 25LL * Ml;
    void *dCp = nullptr, *dCc = nullptr, *dCv = nullptr;
    if (mhb_device_alloc(&dCp, (Ml + 1) * 4) || mhb_device_alloc(&dCc, cap * 4) || mhb_device_alloc(&dCv, cap * 8))
        return 6;
    for (int step = 0; step < 3; ++step)
    {
        const double scale = 1.0 + step;
        std::vector<double> bv(lval);
        for (double &x : bv)
            x *= scale;
        if (mhb_memcpy_h2d(own_val, bv.data(), bv.size() * 8))
            return 7;
        OK(mhb_shard_publish(s)); // ranks may share a GPU: no kernel ever waits for another process's kernel
        host_barrier();
        OK(mhb_shard_pull(s));
        OK(mhb_shard_spgemm_into_begin_f64(s, 0, Ml, dAv, (int *)dCp, (int *)dCc, (double *)dCv, cap));
        OK(mhb_shard_post_size(s, -1));
        long long nnz = 0;
        OK(mhb_shard_spgemm_into_end(s, &nnz));
        OK(mhb_shard_repost_size(s, nnz));
        host_barrier(); // both sizes are posted (the calls above have synchronised their streams)
        long long off = 0, tot = 0, sizes[2] = {0, 0};
        OK(mhb_shard_offsets(s, &off, &tot, sizes));
        // host Gustavson of this rank's rows
        std::vector<int> cp(Ml + 1), cc(nnz);
        std::vector<double> cv(nnz);
        if (mhb_memcpy_d2h(cp.data(), dCp, (Ml + 1) * 4) || mhb_memcpy_d2h(cc.data(), dCc, nnz * 4) ||
            mhb_memcpy_d2h(cv.data(), dCv, nnz * 8))
            return 8;
        long long want_before = 0, want_total = 0, bad = 0;
        int p = 0;
        for (int i = 0; i < M; ++i)
        {
            std::map<int, double> row;
            for (int j = ptr[i]; j < ptr[i + 1]; ++j)
                for (int q = ptr[col[j]]; q < ptr[col[j] + 1]; ++q)
                    row[col[q]] += val[j] * (val[q] * scale);
            want_total += (long long)row.size();
            if (i < r0)
                want_before += (long long)row.size();
            if (i < r0 || i >= r1)
                continue;
            if (cp[i - r0] != p)
                ++bad;
            for (auto &kv : row)
            {
                if (p >= nnz || cc[p] != kv.first || std::fabs(cv[p] - kv.second) > 1e-12 * std::fabs(kv.second))
                    ++bad;
                ++p;
            }
        }
        if (bad || p != nnz || cp[Ml] != nnz || off != want_before || tot != want_total || sizes[rank] != nnz)
        {
            std::fprintf(stderr, "rank %d step %d: bad %lld, nnz %lld vs %d, off %lld vs %lld, total %lld vs %lld\n", rank, step,
                         bad, nnz, p, off, want_before, tot, want_total);
            return 9;
        }
        host_barrier(); // nobody is still pulling when the next step rewrites the shard
    }
    mhb_stats st;
    mhb_get_stats(h, &st);
    std::printf("SHARD-CPP-OK rank %d rows [%d,%d) fused_calls %d\n", rank, r0, r1, st.fused_calls);
    std::fflush(stdout);
    mhb_shard_destroy(s);
    mhb_destroy(h);
    return 0;
}

int main()
{
    sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (sh == MAP_FAILED)
        return 1;
    new (sh) Shared();
    pid_t pid[2];
    for (int r = 0; r < 2; ++r)
    {
        pid[r] = fork(); // before any CUDA call in this process tree
        if (pid[r] == 0)
            std::_Exit(run_rank(r));
    }
    int bad = 0;
    for (int r = 0; r < 2; ++r)
    {
        int status = 0;
        waitpid(pid[r], &status, 0);
        bad |= !(WIFEXITED(status) && WEXITSTATUS(status) == 0);
    }
    return bad;
}
