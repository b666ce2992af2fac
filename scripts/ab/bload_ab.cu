// bload_ab.cu -- A/B/C of the three ways to bring the rows of B to the lanes of the numeric kernel
// (north_star: "128-bit B loads", "SMEM or TMA staging" -- VERDICT r1 item 6), on the B-side access
// stream of C = A*A for the cant-like FEM matrix (27-point stencil, 3 dof, 8 x 8 x nz nodes):
// every warp walks rows of A and, for every nonzero (i, k), streams row k of B -- columns (int32)
// and values (fp64) -- exactly as the numeric kernels do, and folds them into a checksum (the
// accumulator traffic is left out on purpose: this measures the LOAD path alone).
//   A  per-lane 32-bit column / 64-bit value loads, lane l takes entry q + l + 32 t   (what ships)
//   B  per-lane 64-bit column-pair / 128-bit value-pair loads with an alignment peel
//   C  1-D bulk async copies (cp.async.bulk, the non-tensor TMA path) of the row's columns and values
//      into a two-stage shared-memory ring guarded by mbarriers, then LDS by the lanes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bload_ab bload_ab.cu
// Run:   ./bload_ab [nz]      prints one line per variant: ms, GB/s of B bytes requested, checksum
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do                                                                                         \
    {                                                                                          \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess)                                                                  \
        {                                                                                      \
            std::fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);   \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

constexpr int kThreads = 128, kWarps = kThreads / 32;

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- A: what ships ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_lane32(int M, const int *__restrict__ Ap, const int *__restrict__ Ac,
                                                     const int *__restrict__ Bp, const int *__restrict__ Bc,
                                                     const double *__restrict__ Bv, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int row = blockIdx.x * kWarps + (threadIdx.x >> 5); row < M; row += gridDim.x * kWarps)
        for (int j = Ap[row]; j < Ap[row + 1]; ++j)
        {
            const int k = __ldg(&Ac[j]);
            const int qs = __ldg(&Bp[k]), qe = __ldg(&Bp[k + 1]);
            for (int q = qs + lane; q < qe; q += 32)
                acc += (double)__ldg(&Bc[q]) * __ldg(&Bv[q]);
        }
    acc = warp_sum(acc);
    if (lane == 0)
        atomicAdd(out, acc);
}

// ---- B: 64-/128-bit per-lane loads, peel to an even entry -------------------------------------
__global__ void __launch_bounds__(kThreads) k_lane128(int M, const int *__restrict__ Ap, const int *__restrict__ Ac,
                                                      const int *__restrict__ Bp, const int *__restrict__ Bc,
                                                      const double *__restrict__ Bv, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int row = blockIdx.x * kWarps + (threadIdx.x >> 5); row < M; row += gridDim.x * kWarps)
        for (int j = Ap[row]; j < Ap[row + 1]; ++j)
        {
            const int k = __ldg(&Ac[j]);
            int qs = __ldg(&Bp[k]);
            const int qe = __ldg(&Bp[k + 1]);
            if ((qs & 1) && qs < qe) // peel: entry qs is not 8-/16-byte aligned
            {
                if (lane == 0)
                    acc += (double)__ldg(&Bc[qs]) * __ldg(&Bv[qs]);
                ++qs;
            }
            const int pairs = (qe - qs) >> 1;
            const int2 *c2 = reinterpret_cast<const int2 *>(Bc + qs);
            const double2 *v2 = reinterpret_cast<const double2 *>(Bv + qs);
            for (int p = lane; p < pairs; p += 32)
            {
                const int2 c = __ldg(&c2[p]);
                const double2 v = __ldg(&v2[p]);
                acc += (double)c.x * v.x + (double)c.y * v.y;
            }
            if (((qe - qs) & 1) && lane == 31)
                acc += (double)__ldg(&Bc[qe - 1]) * __ldg(&Bv[qe - 1]);
        }
    acc = warp_sum(acc);
    if (lane == 0)
        atomicAdd(out, acc);
}

// ---- C: bulk async copies into a shared-memory ring -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase)
{
    uint32_t done;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(smem_u32(b)), "r"(phase)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

constexpr int kRowCap = 128; // entries of a B row staged per slot (cant-like rows hold <= 81; +3 of alignment slack)
struct __align__(16) Slot
{
    double v[kRowCap + 8];
    int c[kRowCap + 8];
};

__global__ void __launch_bounds__(kThreads) k_bulk(int M, const int *__restrict__ Ap, const int *__restrict__ Ac,
                                                   const int *__restrict__ Bp, const int *__restrict__ Bc,
                                                   const double *__restrict__ Bv, double *__restrict__ out)
{
    __shared__ Slot ring[kWarps][2];
    __shared__ __align__(8) uint64_t bar[kWarps][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
    {
        mbar_init(&bar[w][0], 1);
        mbar_init(&bar[w][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    double acc = 0.0;
    uint32_t phase[2] = {0, 0};
    // issue the copies of B row k into slot s (16-byte aligned source ranges: over-fetch the ends)
    auto issue = [&](int k, int s, int &coff, int &voff, int &len) {
        const int qs = __ldg(&Bp[k]), qe = __ldg(&Bp[k + 1]);
        len = min(qe - qs, kRowCap);
        const int c0 = qs & ~3, c1 = (qs + len + 3) & ~3; // int32: 4 per 16 B
        const int v0 = qs & ~1, v1 = (qs + len + 1) & ~1; // fp64: 2 per 16 B
        coff = qs - c0, voff = qs - v0;
        if (lane == 0 && len > 0)
        {
            const uint32_t cb = (uint32_t)(c1 - c0) * 4u, vb = (uint32_t)(v1 - v0) * 8u;
            mbar_expect(&bar[w][s], cb + vb);
            bulk_g2s(ring[w][s].c, Bc + c0, cb, &bar[w][s]);
            bulk_g2s(ring[w][s].v, Bv + v0, vb, &bar[w][s]);
        }
    };
    for (int row = blockIdx.x * kWarps + w; row < M; row += gridDim.x * kWarps)
    {
        const int s0 = Ap[row], e0 = Ap[row + 1];
        if (s0 == e0)
            continue;
        int coff[2], voff[2], len[2];
        issue(__ldg(&Ac[s0]), 0, coff[0], voff[0], len[0]);
        for (int j = s0; j < e0; ++j)
        {
            const int s = (j - s0) & 1;
            if (j + 1 < e0)
                issue(__ldg(&Ac[j + 1]), s ^ 1, coff[s ^ 1], voff[s ^ 1], len[s ^ 1]);
            if (len[s] > 0)
            {
                mbar_wait(&bar[w][s], phase[s]);
                phase[s] ^= 1;
                for (int t = lane; t < len[s]; t += 32)
                    acc += (double)ring[w][s].c[coff[s] + t] * ring[w][s].v[voff[s] + t];
            }
            __syncwarp(); // slot s is free for the copy issued two steps from now
        }
    }
    acc = warp_sum(acc);
    if (lane == 0)
        atomicAdd(out, acc);
}

int main(int argc, char **argv)
{
    const int nx = 8, ny = 8, nz = argc > 1 ? std::atoi(argv[1]) : 325, dof = 3;
    const int nn = nx * ny * nz, M = nn * dof;
    std::vector<int> ptr(M + 1, 0), col;
    std::vector<double> val;
    for (int node = 0; node < nn; ++node)
    {
        const int x = node % nx, y = (node / nx) % ny, z = node / (nx * ny);
        std::vector<int> nb;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if (x + dx >= 0 && x + dx < nx && y + dy >= 0 && y + dy < ny && z + dz >= 0 && z + dz < nz)
                        nb.push_back(node + dx + dy * nx + dz * nx * ny);
        for (int a = 0; a < dof; ++a)
        {
            for (int n2 : nb)
                for (int b = 0; b < dof; ++b)
                {
                    col.push_back(n2 * dof + b);
                    val.push_back(0.5 + ((n2 * 7 + a * 3 + b) % 97) / 97.0);
                }
            ptr[node * dof + a + 1] = (int)col.size();
        }
    }
    const long long nnz = (long long)col.size();
    long long products = 0;
    for (long long j = 0; j < nnz; ++j)
        products += ptr[col[j] + 1] - ptr[col[j]];
    int *dp, *dc;
    double *dv, *dout;
    CK(cudaMalloc(&dp, (M + 1) * 4));
    CK(cudaMalloc(&dc, nnz * 4 + 64)); // the bulk copies over-fetch to 16-byte boundaries
    CK(cudaMalloc(&dv, nnz * 8 + 64));
    CK(cudaMalloc(&dout, 8));
    CK(cudaMemcpy(dp, ptr.data(), (M + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dc, col.data(), nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv, val.data(), nnz * 8, cudaMemcpyHostToDevice));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 8;
    std::printf("cant-like FEM 8x8x%d x3dof: %d rows, %lld nnz, %lld products, %.1f MB of B requested per pass\n", nz, M, nnz,
                products, products * 12.0 / 1e6);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](const char *name, void (*kern)(int, const int *, const int *, const int *, const int *, const double *, double *)) {
        float best = 1e30f;
        double sum = 0.0;
        for (int it = 0; it < 12; ++it)
        {
            CK(cudaMemset(dout, 0, 8));
            CK(cudaEventRecord(e0));
            kern<<<grid, kThreads>>>(M, dp, dc, dp, dc, dv, dout);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 2 && ms < best)
                best = ms;
            CK(cudaMemcpy(&sum, dout, 8, cudaMemcpyDeviceToHost));
        }
        std::printf("%-28s %8.3f ms  %8.1f GB/s of B entries  checksum %.6e\n", name, best, products * 12.0 / best / 1e6, sum);
    };
    run("A lane 32/64-bit loads", k_lane32);
    run("B lane 64/128-bit loads", k_lane128);
    run("C cp.async.bulk + mbarrier", k_bulk);
    return 0;
}
