// mhb_shard.cu -- row-sharded SpGEMM across the GPUs of one box (declared in
// include/mhb_spgemm.h, "Row-sharded SpGEMM").  A layer on top of the single-GPU C ABI: one
// mhb_shard per rank drives that rank's mhb_handle_t.
//
// The reference has no multi-GPU path.  Gustavson rows are independent:
// C(i,:) needs A(i,:) and the rows of B that A(i,:) selects, so the ONLY exchange is B.
// B is row-sharded like A; every rank keeps one contiguous IMAGE of the B rows [i0, i1) it
// needs (its own shard lives inside the image) in a window exported with CUDA IPC, and reads
// the pieces it does not own straight out of the owners' windows over NVLink / NVSwitch:
//
//   exchange  = k_shard_publish  one 8-byte store per peer: "my shard is final for step e"
//             + k_shard_pull     waits for its owners' flags, copies their pieces (peer loads)
//   sizes     = k_shard_post     nnz(C slice) + epoch stored into every rank's mailbox
//   barrier   = k_shard_barrier  arrive-stores to all ranks, then wait on the own mailbox
//
// No send/recv pairing, no rendezvous, no host-side collective inside a step: the fixed
// ~0.4 ms that grouped NCCL send/recv + an all-gather cost a 0.5 ms step in round 1 becomes two
// kernels of a few microseconds.  NCCL (loaded at run time) remains for the broadcast layout.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mhb_spgemm.h"

namespace
{

// ---- mailbox (8-byte words, one per source rank and kind) ------------------------------------
enum MailKind
{
    MAIL_PUB = 0,   // epoch of the source's last publish
    MAIL_BAR,       // epoch of the source's last barrier arrival
    MAIL_SIZE_VAL,  // source's nnz(C slice)
    MAIL_SIZE_EP,   // epoch that value belongs to
    MAIL_KINDS
};

constexpr unsigned long long kWaitNs = 20ull * 1000 * 1000 * 1000; // a peer that never shows up: give up after 20 s

__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= want; false on time-out
__device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long want)
{
    const unsigned long long t0 = now_ns();
    while (ld_sys(p) < want)
    {
        if (now_ns() - t0 > kWaitNs)
            return false;
        __nanosleep(100);
    }
    return true;
}

struct PeerMail
{
    unsigned long long *mail[8]; // mailbox base of every rank (own one included), world <= 8
};

// "everything I wrote to my shard of B before this point is final for step `epoch`"
__global__ void k_shard_publish(PeerMail pm, int world, int rank, unsigned long long epoch)
{
    const int p = threadIdx.x;
    if (p < world && p != rank)
    {
        __threadfence_system();
        st_sys(pm.mail[p] + MAIL_PUB * world + rank, epoch);
    }
}

struct PullPiece
{
    const int *src_col;           // in the owner's window (peer-mapped)
    const unsigned char *src_val;
    long long dst;                // element offset in this rank's image
    long long count;
    int owner;
    int pad;
};
struct PullPlan
{
    PullPiece piece[8];
    int npieces;
};

template <typename V>
__global__ void __launch_bounds__(256) k_shard_pull(PullPlan plan, const unsigned long long *mail, int world,
                                                    unsigned long long epoch, int *__restrict__ img_col,
                                                    V *__restrict__ img_val, int *__restrict__ err)
{
    __shared__ int ok;
    if (threadIdx.x == 0)
        ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < plan.npieces)
        if (!wait_ge(mail + MAIL_PUB * world + plan.piece[threadIdx.x].owner, epoch))
            ok = 0;
    __syncthreads();
    if (!ok)
    {
        if (threadIdx.x == 0)
            atomicExch(err, 1);
        return;
    }
    __threadfence_system();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int p = 0; p < plan.npieces; ++p)
    {
        const PullPiece pc = plan.piece[p];
        const V *sv = reinterpret_cast<const V *>(pc.src_val);
        // four independent loads in flight per thread: NVLink latency, not width, is what a
        // one-element-per-iteration loop would pay for
        long long i = t0;
        for (; i + 3 * stride < pc.count; i += 4 * stride)
        {
            const int c0 = __ldcg(pc.src_col + i), c1 = __ldcg(pc.src_col + i + stride),
                      c2 = __ldcg(pc.src_col + i + 2 * stride), c3 = __ldcg(pc.src_col + i + 3 * stride);
            const V v0 = __ldcg(sv + i), v1 = __ldcg(sv + i + stride), v2 = __ldcg(sv + i + 2 * stride),
                    v3 = __ldcg(sv + i + 3 * stride);
            img_col[pc.dst + i] = c0, img_col[pc.dst + i + stride] = c1;
            img_col[pc.dst + i + 2 * stride] = c2, img_col[pc.dst + i + 3 * stride] = c3;
            img_val[pc.dst + i] = v0, img_val[pc.dst + i + stride] = v1;
            img_val[pc.dst + i + 2 * stride] = v2, img_val[pc.dst + i + 3 * stride] = v3;
        }
        for (; i < pc.count; i += stride)
        {
            img_col[pc.dst + i] = __ldcg(pc.src_col + i);
            img_val[pc.dst + i] = __ldcg(sv + i);
        }
    }
}

__global__ void k_shard_post(PeerMail pm, int world, int rank, unsigned long long value, unsigned long long epoch)
{
    const int p = threadIdx.x;
    if (p < world)
    {
        st_sys(pm.mail[p] + MAIL_SIZE_VAL * world + rank, value);
        __threadfence_system();
        st_sys(pm.mail[p] + MAIL_SIZE_EP * world + rank, epoch);
    }
}

// The same post with the value taken from the device: nnz(C) of the SpGEMM the handle has just
// queued (mhb_spgemm_into_begin_*), without the host having seen it.  When the call's speculation
// gate says the numeric phase stood down (the host will redo it in mhb_spgemm_into_end), the value
// is not final: kSizePending is posted instead and mhb_shard_repost_size follows the redo.
constexpr unsigned long long kSizePending = ~0ull;
__global__ void k_shard_post_dev(PeerMail pm, int world, int rank, const long long *__restrict__ d_nnz,
                                 const int *__restrict__ d_gate, unsigned long long epoch)
{
    const int p = threadIdx.x;
    if (p < world)
    {
        const unsigned long long value = (*d_gate != 0) ? kSizePending : (unsigned long long)*d_nnz;
        st_sys(pm.mail[p] + MAIL_SIZE_VAL * world + rank, value);
        __threadfence_system();
        st_sys(pm.mail[p] + MAIL_SIZE_EP * world + rank, epoch);
    }
}

__global__ void k_shard_wait_sizes(const unsigned long long *mail, int world, unsigned long long epoch,
                                   unsigned long long *out, int *err)
{
    const int o = threadIdx.x;
    if (o < world)
    {
        if (!wait_ge(mail + MAIL_SIZE_EP * world + o, epoch))
            atomicExch(err, 1);
        __threadfence_system();
        unsigned long long v = ld_sys(mail + MAIL_SIZE_VAL * world + o);
        const unsigned long long t0 = now_ns();
        while (v == kSizePending) // the owner is redoing its SpGEMM and will repost
        {
            if (now_ns() - t0 > kWaitNs)
            {
                atomicExch(err, 1);
                break;
            }
            __nanosleep(200);
            v = ld_sys(mail + MAIL_SIZE_VAL * world + o);
        }
        out[o] = v;
    }
}

__global__ void k_shard_barrier(PeerMail pm, int world, int rank, unsigned long long epoch, int *err)
{
    const int p = threadIdx.x;
    if (p < world)
    {
        __threadfence_system();
        st_sys(pm.mail[p] + MAIL_BAR * world + rank, epoch);
        if (!wait_ge(pm.mail[rank] + MAIL_BAR * world + p, epoch))
            atomicExch(err, 1);
    }
}

// min / max column of A's block: the rows of B it references
__global__ void __launch_bounds__(256) k_col_range(const int *__restrict__ col, long long n, int *__restrict__ mm)
{
    int lo = INT_MAX, hi = -1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const int c = col[i];
        lo = min(lo, c);
        hi = max(hi, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && hi >= 0)
    {
        atomicMin(mm, lo);
        atomicMax(mm + 1, hi);
    }
}

__global__ void __launch_bounds__(256) k_shift_cols(const int *__restrict__ in, long long n, int shift,
                                                    int *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = in[i] - shift;
}

// ---- blobs the caller moves between the ranks --------------------------------------------------
struct Blob1
{
    cudaIpcMemHandle_t handle; // window 1: mailbox | own shard of B.ptr
    unsigned long long raw;    // the same window as a plain pointer (ranks that share a process)
    int pid, device;
    int i0, i1;                // image rows this rank will hold
    long long nnz_own;
};
struct Blob2
{
    cudaIpcMemHandle_t handle; // window 2: image col | image val
    unsigned long long raw;
    int pid, device;
    long long own_off;         // element offset of the own shard inside the image
    long long val_byte_off;    // byte offset of the value section inside window 2
};
static_assert(sizeof(Blob1) <= MHB_SHARD_BLOB_BYTES && sizeof(Blob2) <= MHB_SHARD_BLOB_BYTES, "blob size");

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- NCCL, loaded at run time (the single-GPU library must not depend on it) -------------------
struct Id128 // ncclUniqueId (nccl.h: 128 opaque bytes, passed by value)
{
    char internal[128];
};
struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string &err)
{
    if (g_nccl.lib)
        return true;
    const char *names[] = {"libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names)
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
            break;
    if (!lib)
    {
        err = std::string("libnccl.so.2 not found: ") + dlerror();
        return false;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(dlsym(lib, "ncclBroadcast"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Broadcast || !g_nccl.CommDestroy)
    {
        err = "libnccl.so.2 lacks an expected symbol";
        return false;
    }
    g_nccl.lib = lib;
    return true;
}

} // namespace

struct mhb_shard
{
    mhb_handle_t h = nullptr;
    int rank = 0, world = 1, device = 0;
    int K = 0, N = 0, vbytes = 8;
    std::vector<long long> bounds;
    std::string err;
    // A block
    int M = 0, nnzA = 0;
    const int *Ap = nullptr, *Ac = nullptr;
    int *Ac_local = nullptr; // A.col - i0
    int k0 = 0, k1 = 0;      // rows of B the block references
    int i0 = 0, i1 = 0;      // image rows: union of [k0,k1) and the own shard
    long long nnz_own = 0, nnz_img = 0, own_off = 0, halo_bytes = 0;
    // windows
    unsigned char *w1 = nullptr, *w2 = nullptr;
    size_t w1_bytes = 0, w2_bytes = 0, ptr_off = 0, val_byte_off = 0;
    unsigned char *peer_w1[8] = {nullptr}, *peer_w2[8] = {nullptr};
    bool opened1[8] = {false}, opened2[8] = {false};
    std::vector<std::vector<int>> peer_ptr_piece; // B.ptr of the rows each owner contributes (host)
    std::vector<int> piece_ra, piece_rb;
    int *img_ptr = nullptr;
    PeerMail pm{};
    PullPlan plan{};
    int *dev_err = nullptr;
    unsigned long long *dev_sizes = nullptr, *host_sizes = nullptr;
    unsigned long long pub_epoch = 0, bar_epoch = 0, size_epoch = 0;
    int phase = 0;
    void *nccl_comm = nullptr;
};

namespace
{

int sfail(mhb_shard *s, int code, const std::string &msg)
{
    if (s)
        s->err = msg;
    return code;
}
#define SCU(call)                                                                                        \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return sfail(s, e__ == cudaErrorMemoryAllocation ? MHB_ERR_NOMEM : MHB_ERR_CUDA,             \
                         std::string(#call) + ": " + cudaGetErrorString(e__) + " (mhb_shard.cu:" +       \
                             std::to_string(__LINE__) + ")");                                            \
    } while (0)

cudaStream_t stream_of(mhb_shard *s)
{
    void *st = nullptr;
    mhb_get_stream(s->h, &st);
    return (cudaStream_t)st;
}

int check_dev_err(mhb_shard *s, const char *what)
{
    int e = 0;
    SCU(cudaMemcpy(&e, s->dev_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e)
    {
        cudaMemset(s->dev_err, 0, sizeof(int));
        return sfail(s, MHB_ERR_CUDA, std::string(what) + ": a peer did not arrive within 20 s");
    }
    return MHB_OK;
}

int open_peer(mhb_shard *s, const cudaIpcMemHandle_t &hdl, unsigned long long raw, int pid, int device,
              unsigned char **out, bool *opened)
{
    if (pid == (int)getpid())
    {
        // ranks that share a process (one thread per GPU): plain pointers + peer access
        if (device != s->device)
        {
            cudaError_t e = cudaDeviceEnablePeerAccess(device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return sfail(s, MHB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        *out = reinterpret_cast<unsigned char *>(raw);
        *opened = false;
        return MHB_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
        return sfail(s, MHB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) +
                                          " (peer-mapped B windows need CUDA IPC between the ranks' processes)");
    *out = reinterpret_cast<unsigned char *>(p);
    *opened = true;
    return MHB_OK;
}

} // namespace

extern "C"
{

    int mhb_shard_create(mhb_shard_t *out, mhb_handle_t h, int rank, int world, int K, int N, int value_bytes,
                         const long long *bounds)
    {
        if (!out || !h || !bounds || world < 1 || world > 8 || rank < 0 || rank >= world || K < 0 || N < 0 ||
            (value_bytes != 8 && value_bytes != 4))
            return MHB_ERR_ARG;
        for (int r = 0; r < world; ++r)
            if (bounds[r] > bounds[r + 1] || bounds[0] != 0 || bounds[world] != K)
                return MHB_ERR_ARG;
        mhb_shard *s = new mhb_shard();
        s->h = h, s->rank = rank, s->world = world, s->K = K, s->N = N, s->vbytes = value_bytes;
        s->bounds.assign(bounds, bounds + world + 1);
        cudaGetDevice(&s->device);
        if (cudaMalloc((void **)&s->dev_err, sizeof(int)) != cudaSuccess ||
            cudaMemset(s->dev_err, 0, sizeof(int)) != cudaSuccess ||
            cudaMalloc((void **)&s->dev_sizes, sizeof(unsigned long long) * 8) != cudaSuccess ||
            cudaMallocHost((void **)&s->host_sizes, sizeof(unsigned long long) * 8) != cudaSuccess)
        {
            delete s;
            return MHB_ERR_NOMEM;
        }
        *out = s;
        return MHB_OK;
    }

    int mhb_shard_destroy(mhb_shard_t s)
    {
        if (!s)
            return MHB_OK;
        cudaStreamSynchronize(stream_of(s));
        for (int p = 0; p < s->world; ++p)
        {
            if (s->opened1[p])
                cudaIpcCloseMemHandle(s->peer_w1[p]);
            if (s->opened2[p])
                cudaIpcCloseMemHandle(s->peer_w2[p]);
        }
        if (s->nccl_comm && g_nccl.CommDestroy)
            g_nccl.CommDestroy(s->nccl_comm);
        cudaFree(s->w1), cudaFree(s->w2), cudaFree(s->Ac_local), cudaFree(s->img_ptr), cudaFree(s->dev_err);
        cudaFree(s->dev_sizes);
        cudaFreeHost(s->host_sizes);
        delete s;
        return MHB_OK;
    }

    const char *mhb_shard_last_error(mhb_shard_t s) { return s ? s->err.c_str() : "null shard"; }

    int mhb_shard_set_A(mhb_shard_t s, int M_local, int nnzA, const int *dA_ptr, const int *dA_col,
                        const int *dBown_ptr)
    {
        if (!s || M_local < 0 || nnzA < 0 || !dA_ptr || (nnzA > 0 && !dA_col) || !dBown_ptr)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_set_A: bad argument");
        if (s->phase != 0)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_set_A: the plan of this shard is already being built");
        s->M = M_local, s->nnzA = nnzA, s->Ap = dA_ptr, s->Ac = dA_col;
        cudaStream_t st = stream_of(s);
        // rows of B the block references
        int mm[2] = {INT_MAX, -1};
        int *dmm = reinterpret_cast<int *>(s->dev_sizes); // scratch
        SCU(cudaMemcpyAsync(dmm, mm, sizeof(mm), cudaMemcpyHostToDevice, st));
        if (nnzA > 0)
            k_col_range<<<(int)std::min<long long>(((long long)nnzA + 255) / 256, 148 * 8), 256, 0, st>>>(dA_col, (long long)nnzA, dmm);
        SCU(cudaMemcpyAsync(mm, dmm, sizeof(mm), cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        if (mm[1] >= s->K)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_set_A: a column index of A is >= K");
        const int b0 = (int)s->bounds[s->rank], b1 = (int)s->bounds[s->rank + 1];
        s->k0 = mm[1] >= 0 ? mm[0] : b0;
        s->k1 = mm[1] >= 0 ? mm[1] + 1 : b0;
        s->i0 = b1 > b0 ? std::min(s->k0, b0) : s->k0;
        s->i1 = b1 > b0 ? std::max(s->k1, b1) : s->k1;
        // window 1: mailbox | own shard of B.ptr
        const int kown = b1 - b0;
        int last = 0;
        SCU(cudaMemcpy(&last, dBown_ptr + kown, sizeof(int), cudaMemcpyDeviceToHost));
        s->nnz_own = last;
        s->ptr_off = align256(sizeof(unsigned long long) * MAIL_KINDS * s->world);
        s->w1_bytes = s->ptr_off + sizeof(int) * ((size_t)kown + 1);
        SCU(cudaMalloc((void **)&s->w1, s->w1_bytes));
        SCU(cudaMemset(s->w1, 0, s->ptr_off));
        SCU(cudaMemcpy(s->w1 + s->ptr_off, dBown_ptr, sizeof(int) * ((size_t)kown + 1), cudaMemcpyDeviceToDevice));
        SCU(cudaDeviceSynchronize());
        s->phase = 1;
        return MHB_OK;
    }

    int mhb_shard_export(mhb_shard_t s, int phase, void *blob)
    {
        if (!s || !blob || (phase != 1 && phase != 2))
            return sfail(s, MHB_ERR_ARG, "mhb_shard_export: bad argument");
        std::memset(blob, 0, MHB_SHARD_BLOB_BYTES);
        if (phase == 1)
        {
            if (s->phase < 1)
                return sfail(s, MHB_ERR_ARG, "mhb_shard_export(1) before mhb_shard_set_A");
            Blob1 b{};
            SCU(cudaIpcGetMemHandle(&b.handle, s->w1));
            b.raw = (unsigned long long)(uintptr_t)s->w1;
            b.pid = (int)getpid(), b.device = s->device, b.i0 = s->i0, b.i1 = s->i1, b.nnz_own = s->nnz_own;
            std::memcpy(blob, &b, sizeof(b));
        }
        else
        {
            if (s->phase < 2)
                return sfail(s, MHB_ERR_ARG, "mhb_shard_export(2) before mhb_shard_import(1)");
            Blob2 b{};
            SCU(cudaIpcGetMemHandle(&b.handle, s->w2));
            b.raw = (unsigned long long)(uintptr_t)s->w2;
            b.pid = (int)getpid(), b.device = s->device, b.own_off = s->own_off, b.val_byte_off = (long long)s->val_byte_off;
            std::memcpy(blob, &b, sizeof(b));
        }
        return MHB_OK;
    }

    int mhb_shard_import(mhb_shard_t s, int phase, const void *blobs)
    {
        if (!s || !blobs || (phase != 1 && phase != 2))
            return sfail(s, MHB_ERR_ARG, "mhb_shard_import: bad argument");
        const unsigned char *bb = static_cast<const unsigned char *>(blobs);
        const int W = s->world, me = s->rank;
        if (phase == 1)
        {
            if (s->phase != 1)
                return sfail(s, MHB_ERR_ARG, "mhb_shard_import(1) out of order");
            // open every rank's mailbox window
            for (int p = 0; p < W; ++p)
            {
                Blob1 b;
                std::memcpy(&b, bb + (size_t)p * MHB_SHARD_BLOB_BYTES, sizeof(b));
                if (p == me)
                    s->peer_w1[p] = s->w1;
                else if (int rc = open_peer(s, b.handle, b.raw, b.pid, b.device, &s->peer_w1[p], &s->opened1[p]))
                    return rc;
                s->pm.mail[p] = reinterpret_cast<unsigned long long *>(s->peer_w1[p]);
            }
            // row offsets of the image: the B.ptr pieces of the owners that overlap [i0, i1)
            const int rows = s->i1 - s->i0;
            std::vector<int> img((size_t)rows + 1, 0);
            s->peer_ptr_piece.assign(W, {});
            s->piece_ra.assign(W, 0), s->piece_rb.assign(W, 0);
            long long run = 0;
            for (int o = 0; o < W; ++o)
            {
                const int ra = std::max<long long>(s->i0, s->bounds[o]), rb = std::min<long long>(s->i1, s->bounds[o + 1]);
                if (ra >= rb)
                    continue;
                std::vector<int> piece((size_t)(rb - ra) + 1);
                const int *src = reinterpret_cast<const int *>(s->peer_w1[o] + s->ptr_off) + (ra - s->bounds[o]);
                SCU(cudaMemcpy(piece.data(), src, sizeof(int) * piece.size(), cudaMemcpyDefault));
                for (int r = ra; r < rb; ++r)
                    img[r - s->i0] = (int)(run + piece[r - ra] - piece[0]);
                if (o == me)
                    s->own_off = run; // ra == bounds[me]: the whole own shard is part of the image
                run += piece.back() - piece[0];
                if (run > INT_MAX)
                    return sfail(s, MHB_ERR_OVERFLOW, "the image of B needed by this rank exceeds 2^31-1 entries");
                s->piece_ra[o] = ra, s->piece_rb[o] = rb;
                s->peer_ptr_piece[o] = std::move(piece);
            }
            img[rows] = (int)run;
            s->nnz_img = run;
            SCU(cudaMalloc((void **)&s->img_ptr, sizeof(int) * ((size_t)rows + 1)));
            SCU(cudaMemcpy(s->img_ptr, img.data(), sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice));
            // A's columns relative to the image
            SCU(cudaMalloc((void **)&s->Ac_local, sizeof(int) * (size_t)std::max(s->nnzA, 1)));
            if (s->nnzA > 0)
                k_shift_cols<<<(int)std::min<long long>(((long long)s->nnzA + 255) / 256, 148 * 8), 256>>>(s->Ac, (long long)s->nnzA, s->i0, s->Ac_local);
            // window 2: the image (col | val), own shard inside it
            s->val_byte_off = align256(sizeof(int) * (size_t)std::max<long long>(s->nnz_img, 1));
            s->w2_bytes = s->val_byte_off + (size_t)s->vbytes * (size_t)std::max<long long>(s->nnz_img, 1);
            SCU(cudaMalloc((void **)&s->w2, s->w2_bytes));
            SCU(cudaDeviceSynchronize());
            s->phase = 2;
            return MHB_OK;
        }
        if (s->phase != 2)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_import(2) out of order");
        s->plan.npieces = 0;
        s->halo_bytes = 0;
        for (int o = 0; o < W; ++o)
        {
            Blob2 b;
            std::memcpy(&b, bb + (size_t)o * MHB_SHARD_BLOB_BYTES, sizeof(b));
            if (o == me)
            {
                s->peer_w2[o] = s->w2;
                continue;
            }
            if (s->piece_rb[o] <= s->piece_ra[o])
                continue; // nothing needed from this owner
            const std::vector<int> &piece = s->peer_ptr_piece[o];
            const long long count = piece.back() - piece[0];
            if (count == 0)
                continue;
            if (int rc = open_peer(s, b.handle, b.raw, b.pid, b.device, &s->peer_w2[o], &s->opened2[o]))
                return rc;
            // piece[0] is the owner's local offset of row ra (its B.ptr is rebased to its first row)
            const long long src_el = b.own_off + piece[0];
            int dst_first = 0;
            SCU(cudaMemcpy(&dst_first, s->img_ptr + (s->piece_ra[o] - s->i0), sizeof(int), cudaMemcpyDeviceToHost));
            PullPiece &pc = s->plan.piece[s->plan.npieces++];
            pc.src_col = reinterpret_cast<const int *>(s->peer_w2[o]) + src_el;
            pc.src_val = s->peer_w2[o] + b.val_byte_off + (size_t)src_el * s->vbytes;
            pc.dst = dst_first, pc.count = count, pc.owner = o, pc.pad = 0;
            s->halo_bytes += count * (4 + s->vbytes);
        }
        s->phase = 3;
        return MHB_OK;
    }

    int mhb_shard_own_B(mhb_shard_t s, int **dB_col_own, void **dB_val_own, long long *nnz_own)
    {
        if (!s || s->phase < 2)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_own_B before mhb_shard_import(1)");
        if (dB_col_own)
            *dB_col_own = reinterpret_cast<int *>(s->w2) + s->own_off;
        if (dB_val_own)
            *dB_val_own = s->w2 + s->val_byte_off + (size_t)s->own_off * s->vbytes;
        if (nnz_own)
            *nnz_own = s->nnz_own;
        return MHB_OK;
    }

    int mhb_shard_image(mhb_shard_t s, int *k0, int *k1, long long *nnz_image, long long *halo_bytes_per_step)
    {
        if (!s || s->phase < 2)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_image before mhb_shard_import(1)");
        if (k0)
            *k0 = s->i0;
        if (k1)
            *k1 = s->i1;
        if (nnz_image)
            *nnz_image = s->nnz_img;
        if (halo_bytes_per_step)
            *halo_bytes_per_step = s->halo_bytes;
        return MHB_OK;
    }

    int mhb_shard_publish(mhb_shard_t s)
    {
        if (!s || s->phase != 3)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_publish before the plan is complete");
        if (s->world == 1)
            return MHB_OK;
        ++s->pub_epoch;
        k_shard_publish<<<1, 32, 0, stream_of(s)>>>(s->pm, s->world, s->rank, s->pub_epoch);
        SCU(cudaGetLastError());
        return MHB_OK;
    }

    int mhb_shard_exchange(mhb_shard_t s)
    {
        if (int rc = mhb_shard_publish(s))
            return rc;
        return mhb_shard_pull(s);
    }

    int mhb_shard_pull(mhb_shard_t s)
    {
        if (!s || s->phase != 3)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_pull before the plan is complete");
        if (s->world == 1)
            return MHB_OK;
        cudaStream_t st = stream_of(s);
        if (s->plan.npieces > 0)
        {
            long long total = 0;
            for (int p = 0; p < s->plan.npieces; ++p)
                total += s->plan.piece[p].count;
            const int grid = (int)std::max<long long>(1, std::min<long long>((total + 1023) / 1024, 148 * 4));
            const unsigned long long *mail = s->pm.mail[s->rank];
            int *ic = reinterpret_cast<int *>(s->w2);
            if (s->vbytes == 8)
                k_shard_pull<double><<<grid, 256, 0, st>>>(s->plan, mail, s->world, s->pub_epoch, ic,
                                                           reinterpret_cast<double *>(s->w2 + s->val_byte_off), s->dev_err);
            else
                k_shard_pull<float><<<grid, 256, 0, st>>>(s->plan, mail, s->world, s->pub_epoch, ic,
                                                          reinterpret_cast<float *>(s->w2 + s->val_byte_off), s->dev_err);
        }
        SCU(cudaGetLastError());
        return MHB_OK;
    }

    int mhb_shard_barrier(mhb_shard_t s)
    {
        if (!s || s->phase < 2)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_barrier before mhb_shard_import(1)");
        if (s->world == 1)
            return MHB_OK;
        ++s->bar_epoch;
        k_shard_barrier<<<1, 32, 0, stream_of(s)>>>(s->pm, s->world, s->rank, s->bar_epoch, s->dev_err);
        SCU(cudaGetLastError());
        return MHB_OK;
    }

    int mhb_shard_symbolic(mhb_shard_t s, int r_lo, int r_hi, int *dC_ptr, long long *nnzC)
    {
        if (!s || s->phase != 3 || r_lo < 0 || r_hi < r_lo || r_hi > s->M)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_symbolic: bad row range or incomplete plan");
        int rc = mhb_symbolic(s->h, r_hi - r_lo, s->i1 - s->i0, s->N, s->nnzA, s->Ap + r_lo, s->Ac_local,
                              (int)s->nnz_img, s->img_ptr, reinterpret_cast<const int *>(s->w2), dC_ptr, nnzC);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }

    int mhb_shard_numeric_f64(mhb_shard_t s, const double *dA_val, int *dC_col, double *dC_val)
    {
        if (!s || s->phase != 3 || s->vbytes != 8)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_numeric_f64: incomplete plan or value type mismatch");
        int rc = mhb_numeric_f64(s->h, dA_val, reinterpret_cast<const double *>(s->w2 + s->val_byte_off), dC_col, dC_val);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }
    int mhb_shard_numeric_f32(mhb_shard_t s, const float *dA_val, int *dC_col, float *dC_val)
    {
        if (!s || s->phase != 3 || s->vbytes != 4)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_numeric_f32: incomplete plan or value type mismatch");
        int rc = mhb_numeric_f32(s->h, dA_val, reinterpret_cast<const float *>(s->w2 + s->val_byte_off), dC_col, dC_val);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }

    int mhb_shard_spgemm_into_f64(mhb_shard_t s, int r_lo, int r_hi, const double *dA_val, int *dC_ptr, int *dC_col,
                                  double *dC_val, long long capacity, long long *nnzC)
    {
        if (!s || s->phase != 3 || s->vbytes != 8 || r_lo < 0 || r_hi < r_lo || r_hi > s->M)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_spgemm_into_f64: bad row range, incomplete plan or value type mismatch");
        int rc = mhb_spgemm_into_f64(s->h, r_hi - r_lo, s->i1 - s->i0, s->N, s->nnzA, s->Ap + r_lo, s->Ac_local, dA_val,
                                     (int)s->nnz_img, s->img_ptr, reinterpret_cast<const int *>(s->w2),
                                     reinterpret_cast<const double *>(s->w2 + s->val_byte_off), dC_ptr, dC_col, dC_val,
                                     capacity, nnzC);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }
    int mhb_shard_spgemm_into_f32(mhb_shard_t s, int r_lo, int r_hi, const float *dA_val, int *dC_ptr, int *dC_col,
                                  float *dC_val, long long capacity, long long *nnzC)
    {
        if (!s || s->phase != 3 || s->vbytes != 4 || r_lo < 0 || r_hi < r_lo || r_hi > s->M)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_spgemm_into_f32: bad row range, incomplete plan or value type mismatch");
        int rc = mhb_spgemm_into_f32(s->h, r_hi - r_lo, s->i1 - s->i0, s->N, s->nnzA, s->Ap + r_lo, s->Ac_local, dA_val,
                                     (int)s->nnz_img, s->img_ptr, reinterpret_cast<const int *>(s->w2),
                                     reinterpret_cast<const float *>(s->w2 + s->val_byte_off), dC_ptr, dC_col, dC_val,
                                     capacity, nnzC);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }

    int mhb_shard_spgemm_into_begin_f64(mhb_shard_t s, int r_lo, int r_hi, const double *dA_val, int *dC_ptr,
                                        int *dC_col, double *dC_val, long long capacity)
    {
        if (!s || s->phase != 3 || s->vbytes != 8 || r_lo < 0 || r_hi < r_lo || r_hi > s->M)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_spgemm_into_begin_f64: bad row range, incomplete plan or value type mismatch");
        int rc = mhb_spgemm_into_begin_f64(s->h, r_hi - r_lo, s->i1 - s->i0, s->N, s->nnzA, s->Ap + r_lo, s->Ac_local,
                                           dA_val, (int)s->nnz_img, s->img_ptr, reinterpret_cast<const int *>(s->w2),
                                           reinterpret_cast<const double *>(s->w2 + s->val_byte_off), dC_ptr, dC_col,
                                           dC_val, capacity);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }
    int mhb_shard_spgemm_into_begin_f32(mhb_shard_t s, int r_lo, int r_hi, const float *dA_val, int *dC_ptr,
                                        int *dC_col, float *dC_val, long long capacity)
    {
        if (!s || s->phase != 3 || s->vbytes != 4 || r_lo < 0 || r_hi < r_lo || r_hi > s->M)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_spgemm_into_begin_f32: bad row range, incomplete plan or value type mismatch");
        int rc = mhb_spgemm_into_begin_f32(s->h, r_hi - r_lo, s->i1 - s->i0, s->N, s->nnzA, s->Ap + r_lo, s->Ac_local,
                                           dA_val, (int)s->nnz_img, s->img_ptr, reinterpret_cast<const int *>(s->w2),
                                           reinterpret_cast<const float *>(s->w2 + s->val_byte_off), dC_ptr, dC_col,
                                           dC_val, capacity);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }
    int mhb_shard_spgemm_into_end(mhb_shard_t s, long long *nnzC)
    {
        if (!s)
            return MHB_ERR_ARG;
        int rc = mhb_spgemm_into_end(s->h, nnzC);
        if (rc)
            s->err = mhb_last_error(s->h);
        return rc;
    }

    int mhb_shard_post_size(mhb_shard_t s, long long nnzC_local)
    {
        if (!s || s->phase < 2 || nnzC_local < -1)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_post_size: bad argument");
        ++s->size_epoch;
        if (nnzC_local == -1)
        {
            // from the device: nnz(C) of the call queued by mhb_shard_spgemm_into_begin_*
            const long long *d_nnz = nullptr;
            const int *d_gate = nullptr;
            if (mhb_get_device_scalars(s->h, &d_nnz, &d_gate) != MHB_OK || !d_nnz)
                return sfail(s, MHB_ERR_ARG, "mhb_shard_post_size(-1): no SpGEMM queued on the handle");
            k_shard_post_dev<<<1, 32, 0, stream_of(s)>>>(s->pm, s->world, s->rank, d_nnz, d_gate, s->size_epoch);
        }
        else
            k_shard_post<<<1, 32, 0, stream_of(s)>>>(s->pm, s->world, s->rank, (unsigned long long)nnzC_local, s->size_epoch);
        SCU(cudaGetLastError());
        return MHB_OK;
    }
    int mhb_shard_repost_size(mhb_shard_t s, long long nnzC_local)
    {
        if (!s || s->phase < 2 || nnzC_local < 0 || s->size_epoch == 0)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_repost_size: bad argument or nothing posted yet");
        k_shard_post<<<1, 32, 0, stream_of(s)>>>(s->pm, s->world, s->rank, (unsigned long long)nnzC_local, s->size_epoch);
        SCU(cudaGetLastError());
        return MHB_OK;
    }

    int mhb_shard_offsets(mhb_shard_t s, long long *slice_offset, long long *nnzC_total, long long *all_sizes)
    {
        if (!s || s->phase < 2 || s->size_epoch == 0)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_offsets before mhb_shard_post_size");
        cudaStream_t st = stream_of(s);
        k_shard_wait_sizes<<<1, 32, 0, st>>>(s->pm.mail[s->rank], s->world, s->size_epoch, s->dev_sizes, s->dev_err);
        SCU(cudaMemcpyAsync(s->host_sizes, s->dev_sizes, sizeof(unsigned long long) * s->world, cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        if (int rc = check_dev_err(s, "mhb_shard_offsets"))
            return rc;
        long long off = 0, tot = 0;
        for (int r = 0; r < s->world; ++r)
        {
            if (r < s->rank)
                off += (long long)s->host_sizes[r];
            tot += (long long)s->host_sizes[r];
            if (all_sizes)
                all_sizes[r] = (long long)s->host_sizes[r];
        }
        if (slice_offset)
            *slice_offset = off;
        if (nnzC_total)
            *nnzC_total = tot;
        return MHB_OK;
    }

    // ---- NCCL from C++ -------------------------------------------------------------------------
    int mhb_nccl_unique_id(void *id128)
    {
        std::string err;
        if (!id128 || !load_nccl(err))
        {
            if (!err.empty())
                std::fprintf(stderr, "mhb_nccl_unique_id: %s\n", err.c_str());
            return MHB_ERR_ARG;
        }
        return g_nccl.GetUniqueId(id128) == 0 ? MHB_OK : MHB_ERR_CUDA;
    }

    int mhb_shard_init_nccl(mhb_shard_t s, const void *id128)
    {
        if (!s || !id128)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_init_nccl: bad argument");
        std::string err;
        if (!load_nccl(err))
            return sfail(s, MHB_ERR_CUDA, err);
        Id128 id;
        std::memcpy(&id, id128, sizeof(id));
        const int rc = g_nccl.CommInitRank(&s->nccl_comm, s->world, id, s->rank);
        if (rc != 0)
            return sfail(s, MHB_ERR_CUDA, std::string("ncclCommInitRank: ") +
                                              (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
        return MHB_OK;
    }

    int mhb_shard_broadcast(mhb_shard_t s, void *dbuf, size_t bytes, int root)
    {
        if (!s || !dbuf || root < 0 || root >= s->world)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_broadcast: bad argument");
        if (s->world == 1)
            return MHB_OK;
        if (!s->nccl_comm)
            return sfail(s, MHB_ERR_ARG, "mhb_shard_broadcast before mhb_shard_init_nccl");
        const int rc = g_nccl.Broadcast(dbuf, dbuf, bytes, /* ncclUint8 */ 1, root, s->nccl_comm, stream_of(s));
        if (rc != 0)
            return sfail(s, MHB_ERR_CUDA, std::string("ncclBroadcast: ") +
                                              (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
        return MHB_OK;
    }

} // extern "C"
