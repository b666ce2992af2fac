// mhb_capi.cu -- host orchestration of the four kernel families and the C ABI declared in
// include/mhb_spgemm.h.  Replaces MH_spgemm (src/main.cu:12-72), the launchers of
// inc/MH_spgemm.cuh and the Tool workspace (src/Tool.cu).
//
// Host-path design (vs the reference's 9 cudaMalloc, 12 stream creations, 7 blocking D2H
// copies and 7 cudaDeviceSynchronize per call): one handle-owned, grow-only
// workspace (no allocation in steady state), one stream, bin offsets computed on the
// device.  Host reads per SpGEMM: two small ones in the symbolic phase on the first call of a
// shape (symbolic bin sizes; nnz(C) + numeric bin sizes, the hand-off the contract requires);
// ONE when the call has the shape of the previous one (the symbolic kernels are launched
// speculatively from that call's bin sizes, RowList); and for mhb_spgemm_into_* (caller-owned
// C arrays) a single synchronisation at the very end -- both phases launched speculatively,
// verified by k_fused_gate on the device, redone the ordinary way on a miss.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/mhb_spgemm.h"
#include "mhb_binning.cuh"
#include "mhb_common.cuh"
#include "mhb_mask.cuh"
#include "mhb_numeric.cuh"
#include "mhb_symbolic.cuh"
#include "mhb_transpose.cuh"

using namespace mhb;

namespace
{

struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;
    bool grew = false;
    cudaError_t ensure(size_t bytes)
    {
        grew = false;
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256; // a little slack so near-equal sizes do not regrow
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess)
            return e;
        cap = want;
        grew = true;
        return cudaSuccess;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf
{
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess)
            return e;
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

enum Ev
{
    EV_START = 0,
    EV_ALLOC,
    EV_MASK,
    EV_SYMBIN,
    EV_SYM,
    EV_NUMBIN,
    EV_HANDOFF,
    EV_NUM0,
    EV_NUM1,
    EV_COUNT
};

} // namespace

namespace
{
// a mhb_spgemm_into_begin_* whose mhb_spgemm_into_end is outstanding
struct IntoCall
{
    bool active = false, pending = false, over = false;
    int M = 0, K = 0, N = 0, nnzA = 0, nnzB = 0, vbytes = 0;
    const int *Ap = nullptr, *Ac = nullptr, *Bp = nullptr, *Bc = nullptr;
    const void *Av = nullptr, *Bv = nullptr;
    int *Cp = nullptr, *Cc = nullptr;
    void *Cv = nullptr;
    long long capacity = 0, nnzC = 0;
};
} // namespace

struct mhb_context
{
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_vals = nullptr, ev_ready = nullptr;
    cudaEvent_t ev_chunk[kMaxRowChunks] = {nullptr}; // host-buffer path: numeric row chunk c is complete
    int row_chunks = 2;                              // option "row_chunks": chunks of the host path (1: off)
    long long row_chunk_bytes = 32LL << 20;          // option "row_chunk_bytes": ... used when C.col + C.val are at least this large
    static constexpr int kAux = 5; // per-bin kernels of one phase run concurrently (the reference uses 12 streams)
    cudaStream_t aux[kAux] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {nullptr};
    int aux_used = 0;
    bool serial = false; // option "serial_bins": one stream, for per-kernel timing
    bool phase_serial = false; // this phase runs on the main stream only (serial, or one populated bin)
    DevBuf bsame, asame_buf, bm_store, bm_slot;
    bool have_bm_store = false;          // symbolic kept the bitmaps of its SB_BM_G8 rows
    int compact_rows = 1;                // option "compact_rows": use NB_WIN_COMPACT
    int claim_list = 1;                  // option "claim_list": k_num_hash_list for the 1 024 / 4 096-slot bins
    const unsigned char *asame = nullptr; // twin flags of A's rows (== bsame when A aliases B)
    int sym_twins = 1;                   // option "sym_twins": symbolic computes one row per run of twin rows of A
    int pdl = 1;                         // option "pdl": programmatic dependent launch of the main-stream chain
    int speculate = 1;                   // option "speculate": launch the symbolic bins from the previous call's bin sizes (no host read #1)
    bool spec_ok = false;                // the previous symbolic on this handle finished and left its bin sizes behind
    int spec_key[5] = {0, 0, 0, 0, 0};   // M, K, N, nnzA, nnzB of that call
    int spec_calls = 0, spec_misses = 0; // statistics (mhb_stats)
    // fused call (mhb_spgemm_into_*): do_symbolic left its host read to the caller, who launches the
    // numeric kernels first; the symbolic bins it launched and the pool size it planned for
    bool fused_pending = false;
    int fused_launched[MHB_MAX_BINS + 1] = {0};
    int fused_planned_tileflop = 0;
    int fused_calls = 0;
    IntoCall into;
    int mask_onepass = 2;                // option "mask_onepass": 2 = two passes around a scan of the chunk counts, 1 = one pass
                                         // with a chained scan, 0 = the round-1 five-kernel chain
    int count_probes = 0;                // option "count_probes": hash kernels count failed probes (HASH_CONFLICT)
    bool asame_early = false;            // A's twin flags were computed beside the mask build (A is not B)
    int row_twins = 0;                   // option "row_twins": dense-window A-row twin fusion (slower: 8 warps/SM)
    std::string err;
    // options
    int force_sym = 0, force_num = 0, verbose = 0;
    long long nnz_limit = INT_MAX; // option "nnz_limit": lets tests exercise the int32-overflow path
    // problem of the last symbolic call
    bool have_pattern = false;
    int M = 0, K = 0, N = 0, nnzA = 0, nnzB = 0;
    const int *Ap = nullptr, *Ac = nullptr, *Bp = nullptr, *Bc = nullptr;
    int *Cp = nullptr;
    long long nnzC = 0;
    int sym_off[MHB_MAX_BINS + 1] = {0}, num_off[MHB_MAX_BINS + 1] = {0};
    int max_tileflop = 0, max_rownnz = 0;
    // workspace
    DevBuf flags, wordprefix, tileptr, tilecol, tilemask, binfo, arow, binid, bins_sym, bins_num, blockhist,
        scan_tmp, scal, pool;
    HostBuf h_scal;
    // host-API staging
    DevBuf sA_ptr, sA_col, sA_val, sB_ptr, sB_col, sB_val, sC_ptr, sC_col, sC_val;
    DevBuf tr_key[2], tr_idx[2], tr_hist; // radix-sort scratch of mhb_transpose_*
    DevBuf chunk_off;                     // [bins][chunks + 1] positions in bins_num (host-buffer path)
    HostBuf hC_ptr, hC_col, hC_val;
    cudaEvent_t ev[EV_COUNT] = {nullptr};
    bool ev_sym_valid = false, ev_num_valid = false;
    mhb_timing timing{};
    mhb_stats stats{};
    int launches = 0;
};

namespace
{

int fail(mhb_context *h, int code, const std::string &msg)
{
    if (h)
        h->err = msg;
    return code;
}

#define CU(call)                                                                                         \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(h, e__ == cudaErrorMemoryAllocation ? MHB_ERR_NOMEM : MHB_ERR_CUDA,              \
                        std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +    \
                            std::to_string(__LINE__) + ")");                                             \
    } while (0)

// Kernels of the main-stream chain (mask build, scans, metrics, binning): every one of them
// starts with pdl_prologue(), so they are launched with programmatic stream serialization --
// the launch latency of kernel i+1 hides behind kernel i (a dozen 3-25 us kernels per call).
#define LAUNCH(h, kern, grid, block, smem, ...)                                                          \
    do                                                                                                   \
    {                                                                                                    \
        cudaLaunchConfig_t cfg__ = {};                                                                   \
        cfg__.gridDim = dim3((unsigned)(grid));                                                          \
        cfg__.blockDim = dim3((unsigned)(block));                                                        \
        cfg__.dynamicSmemBytes = (size_t)(smem);                                                         \
        cfg__.stream = (h)->stream;                                                                      \
        cudaLaunchAttribute at__[1];                                                                     \
        at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                 \
        at__[0].val.programmaticStreamSerializationAllowed = (h)->pdl ? 1 : 0;                           \
        cfg__.attrs = at__;                                                                              \
        cfg__.numAttrs = 1;                                                                              \
        ++(h)->launches;                                                                                 \
        CU(cudaLaunchKernelEx(&cfg__, kern, __VA_ARGS__));                                               \
    } while (0)

#define LAUNCH_ON(h, st, kern, grid, block, smem, ...)                                                  \
    do                                                                                                   \
    {                                                                                                    \
        kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                            \
        ++(h)->launches;                                                                                 \
        CU(cudaGetLastError());                                                                          \
    } while (0)

// Fork / join of the per-bin kernels of one phase: bin kernels are independent (disjoint rows,
// disjoint outputs), so they are spread over the main stream and kAux helper streams.
int fork_bins(mhb_context *h, int populated /* kernels of the phase */)
{
    h->aux_used = 0;
    // a phase with a single kernel (FEM-like inputs) stays on the main stream: no event round
    // trip to a helper stream and back
    h->phase_serial = h->serial || populated <= 1;
    if (h->phase_serial)
        return MHB_OK;
    CU(cudaEventRecord(h->ev_fork, h->stream));
    return MHB_OK;
}
int next_bin_stream(mhb_context *h, cudaStream_t *out)
{
    if (h->phase_serial)
    {
        *out = h->stream;
        return MHB_OK;
    }
    // slots 0..kAux-1 -> helper streams (the first, big-row bins land on the high-priority
    // ones), slot kAux -> main stream, then round robin
    int slot = h->aux_used++;
    int a = slot % (mhb_context::kAux + 1);
    if (a == mhb_context::kAux)
    {
        *out = h->stream;
        return MHB_OK;
    }
    if (slot < mhb_context::kAux)
        CU(cudaStreamWaitEvent(h->aux[a], h->ev_fork, 0));
    *out = h->aux[a];
    return MHB_OK;
}
int join_bins(mhb_context *h)
{
    if (h->phase_serial)
        return MHB_OK;
    int n = std::min(h->aux_used, (int)mhb_context::kAux);
    for (int a = 0; a < n; ++a)
    {
        CU(cudaEventRecord(h->ev_join[a], h->aux[a]));
        CU(cudaStreamWaitEvent(h->stream, h->ev_join[a], 0));
    }
    return MHB_OK;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

template <class K>
cudaError_t allow_smem(K kern, int bytes)
{
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

int log2_ceil(long long v)
{
    int l = 0;
    while ((1LL << l) < v)
        ++l;
    return l;
}

// ---- exclusive scan helper: out[i] (i<n) exclusive prefix, out[n] total if write_total ----
template <class Load>
int run_scan(mhb_context *h, Load in, long long n, int *out, int write_total, long long *total64_dev)
{
    if (n <= 0)
    {
        if (total64_dev)
            CU(cudaMemsetAsync(total64_dev, 0, sizeof(long long), h->stream));
        if (write_total)
            CU(cudaMemsetAsync(out, 0, sizeof(int), h->stream));
        return MHB_OK;
    }
    int nb = cdiv(n, kScanTile);
    long long *bs = h->scan_tmp.as<long long>();
    if (nb == 1) // one tile: the scan kernel alone
    {
        LAUNCH(h, k_scan_apply<Load>, 1, kScanThreads, 0, in, n, (const long long *)nullptr, 1, out, write_total, total64_dev);
        return MHB_OK;
    }
    LAUNCH(h, k_scan_blocksums<Load>, nb, kScanThreads, 0, in, n, bs);
    LAUNCH(h, k_scan_apply<Load>, nb, kScanThreads, 0, in, n, (const long long *)bs, nb, out, write_total, total64_dev);
    return MHB_OK;
}

// ---- stable binning of M rows by binid into `bins`; sizes/offsets to scal[size_at/off_at] ----
int run_binning(mhb_context *h, int M, int nbins, int *bins, int size_at, int off_at)
{
    int *scal = h->scal.as<int>();
    if (M <= 0)
    {
        CU(cudaMemsetAsync(scal + size_at, 0, sizeof(int) * MHB_MAX_BINS, h->stream));
        CU(cudaMemsetAsync(scal + off_at, 0, sizeof(int) * (MHB_MAX_BINS + 1), h->stream));
        return MHB_OK;
    }
    int nb = cdiv(M, kBinThreads);
    const unsigned char *binid = h->binid.as<unsigned char>();
    int *bh = h->blockhist.as<int>();
    LAUNCH(h, k_bin_count, nb, kBinThreads, 0, M, binid, bh, nb);
    LAUNCH(h, k_bin_offsets, 1, 32 * MHB_MAX_BINS, 0, bh, nb, nbins, scal + size_at, scal + off_at);
    LAUNCH(h, k_bin_scatter, nb, kBinThreads, 0, M, binid, bh, nb, bins);
    return MHB_OK;
}

// The scalar block is followed, in the same allocation, by everything else that must be zero
// when a call starts -- the chunk ticket and chunk status words of the chained scan and the
// row-start flag words of the mask builder -- so that ONE memset per call clears them all.
constexpr size_t kCtrlOffset = (size_t)SC_COUNT * 4;    // unsigned ticket (+ padding to 64 bytes)
constexpr size_t kStatusOffset = kCtrlOffset + 64;      // unsigned long long status[nchunks]
inline long long mask_words(long long nnz) { return (nnz + 31) / 32; }
inline int mask_chunks(long long nnz) { return (int)((mask_words(nnz) + kMaskChunkWords - 1) / kMaskChunkWords); }
inline size_t flags_offset(long long nnz) { return kStatusOffset + (size_t)(mask_chunks(nnz) + 1) * 8; }
inline size_t scal_bytes(long long nnz) { return (flags_offset(nnz) + (size_t)(mask_words(nnz) + 2) * 4 + 15) / 16 * 16; }

// zero the first `bytes` (a multiple of 16) of the scalar block on the handle's stream
int zero_scal(mhb_context *h, size_t bytes);

int ensure_workspace(mhb_context *h, int M, int K, int nnzB, bool *grew)
{
    *grew = false;
    long long nW = ((long long)nnzB + 31) / 32;
    struct Req
    {
        DevBuf *b;
        size_t bytes;
    } reqs[] = {
        {&h->flags, (size_t)(nW + 1) * 4},
        {&h->wordprefix, (size_t)(nW + 2) * 4},
        {&h->tileptr, ((size_t)K + 2) * 4},
        {&h->tilecol, ((size_t)nnzB + 1) * 4},
        {&h->tilemask, ((size_t)nnzB + 1) * 4},
        {&h->binfo, ((size_t)K + 1) * 16},
        {&h->bsame, ((size_t)K + 1)},
        {&h->arow, ((size_t)M + 1) * 16},
        {&h->binid, ((size_t)M + 1)},
        {&h->bins_sym, ((size_t)M + 1) * 4},
        {&h->bins_num, ((size_t)M + 1) * 4},
        {&h->blockhist, (size_t)MHB_MAX_BINS * (size_t)(cdiv(std::max(M, 1), kBinThreads) + 1) * 4},
        {&h->scan_tmp, (size_t)(cdiv(std::max<long long>(std::max<long long>(nW, (long long)M + 1), 1), kScanTile) + 2) * 8},
        {&h->scal, scal_bytes(nnzB)},
    };
    for (auto &r : reqs)
    {
        CU(r.b->ensure(r.bytes));
        *grew |= r.b->grew;
    }
    CU(h->h_scal.ensure(SC_COUNT * 4));
    return MHB_OK;
}

int zero_scal(mhb_context *h, size_t bytes)
{
    const long long n16 = (long long)(bytes / 16);
    LAUNCH(h, k_zero16, (int)std::min<long long>(std::max<long long>(cdiv(n16, 256), 1), h->num_sms * 8), 256, 0,
           h->scal.as<uint4>(), n16);
    return MHB_OK;
}

// ---- family 1 -----------------------------------------------------------------------------
int build_mask_matrix(mhb_context *h, int K, int nnzB, const int *Bp, const int *Bc)
{
    const long long nnz = nnzB;
    const long long nW = (nnz + 31) / 32;
    int *wp = h->wordprefix.as<int>();
    int *scal = h->scal.as<int>();
    long long *ntiles_dev = reinterpret_cast<long long *>(scal + SC_NTILES_LO);
    if (h->mask_onepass)
    {
        // the caller has zeroed scal_bytes(nnzB) bytes at h->scal: scalars, ticket, status, flags
        unsigned char *base = h->scal.as<unsigned char>();
        unsigned *oflags = reinterpret_cast<unsigned *>(base + flags_offset(nnz));
        if (nnz > 0)
        {
            LAUNCH(h, k_mask_rowstarts, cdiv(K, 256), 256, 0, K, Bp, oflags);
            unsigned *ctrl = reinterpret_cast<unsigned *>(base + kCtrlOffset);
            unsigned long long *status = reinterpret_cast<unsigned long long *>(base + kStatusOffset);
            const int nch = mask_chunks(nnz);
            if (h->mask_onepass == 2)
            {
                // two passes around a scan of the per-chunk tile counts (they live in the status words)
                int *ct = reinterpret_cast<int *>(status), *cpfx = ct + nch + 1;
                LAUNCH(h, k_mask_build<1>, nch, kMaskThreads, 0, Bc, nnz, nW, oflags, wp, h->tilecol.as<int>(),
                       h->tilemask.as<unsigned>(), ctrl, status, nch, ntiles_dev, ct, (const int *)cpfx);
                int rc = run_scan(h, LoadInt{ct}, nch, cpfx, 0, ntiles_dev);
                if (rc)
                    return rc;
                LAUNCH(h, k_mask_build<2>, nch, kMaskThreads, 0, Bc, nnz, nW, oflags, wp, h->tilecol.as<int>(),
                       h->tilemask.as<unsigned>(), ctrl, status, nch, ntiles_dev, ct, (const int *)cpfx);
            }
            else
                LAUNCH(h, k_mask_build<0>, nch, kMaskThreads, 0, Bc, nnz, nW, oflags, wp, h->tilecol.as<int>(),
                       h->tilemask.as<unsigned>(), ctrl, status, nch, ntiles_dev, (int *)nullptr, (const int *)nullptr);
        }
        LAUNCH(h, k_mask_rows, cdiv((long long)K + 1, 256), 256, 0, K, nnz, Bp, Bc, (const unsigned *)oflags,
               (const int *)wp, (const long long *)ntiles_dev, (const int *)h->tilecol.as<int>(),
               (const unsigned *)h->tilemask.as<unsigned>(), h->tileptr.as<int>(), h->binfo.as<int4>(),
               h->bsame.as<unsigned char>());
        return MHB_OK;
    }
    unsigned *flags = h->flags.as<unsigned>();
    if (nnz > 0)
    {
        CU(cudaMemsetAsync(h->tilemask.p, 0, (size_t)nnz * 4, h->stream));
        int grid = std::min(cdiv(nW * 32, 256), h->num_sms * 32);
        LAUNCH(h, k_mask_flags, grid, 256, 0, Bc, nnz, nW, flags);
        LAUNCH(h, k_mask_rowstarts, cdiv(K, 256), 256, 0, K, Bp, flags);
    }
    int rc = run_scan(h, LoadPopc{flags}, nW, wp, 0, ntiles_dev);
    if (rc)
        return rc;
    LAUNCH(h, k_mask_tileptr, cdiv(K + 1, 256), 256, 0, K, nnz, Bp, Bc, flags, wp, ntiles_dev,
           h->tileptr.as<int>(), h->binfo.as<int4>());
    if (nnz > 0)
    {
        int grid = std::min(cdiv(nW * 32, 256), h->num_sms * 32);
        LAUNCH(h, k_mask_fill, grid, 256, 0, Bc, nnz, nW, flags, wp, h->tilecol.as<int>(),
               h->tilemask.as<unsigned>());
    }
    if (K > 0)
        LAUNCH(h, k_mask_same, cdiv(K, 256), 256, 0, K, h->binfo.as<int4>(), h->tileptr.as<int>(),
               h->tilecol.as<int>(), h->tilemask.as<unsigned>(), h->bsame.as<unsigned char>());
    return MHB_OK;
}

int check_dev_error(mhb_context *h, const int *hs)
{
    if (hs[SC_ERROR] != DEVERR_NONE)
        return fail(h, MHB_ERR_CUDA, "internal: a hash table filled up (device error flag " +
                                         std::to_string(hs[SC_ERROR]) + ")");
    return MHB_OK;
}

// Bins that get a kernel when a phase is launched from the offsets `off`: the populated ones; the
// three cost classes of the thread-per-row kernel (adjacent bins t0 .. t0+2) share ONE launch, so
// all of them are covered as soon as one is populated.
unsigned launched_mask(const int *off, int nbins, int t0)
{
    unsigned m = 0;
    for (int b = 1; b < nbins; ++b) // bin 0 (rows without products) has no kernel
        m |= (off[b + 1] > off[b]) ? (1u << b) : 0u;
    const unsigned tiny = 7u << t0;
    if (m & tiny)
        m |= tiny;
    return m;
}
// kernels behind a launched mask (the tiny family counts once)
int launched_kernels(unsigned m, int t0)
{
    const unsigned tiny = 7u << t0;
    return __builtin_popcount(m & ~tiny) + ((m & tiny) ? 1 : 0);
}

// ---- family 3 launches ------------------------------------------------------------------
// spec: the host has NOT read this call's bin sizes; h->sym_off / h->max_tileflop are those of the
// previous call on the handle and only size the grids and scratch -- the kernels take their
// row ranges from the device-side offsets (RowList), and do_symbolic verifies the guess later.
int launch_symbolic_bins(mhb_context *h, bool spec)
{
    const int *off = h->sym_off;
    auto n_of = [&](int b) { return off[b + 1] - off[b]; };
    const int *bins = h->bins_sym.as<int>();
    auto list = [&](int b) {
        return spec ? RowList{bins, h->scal.as<int>() + SC_SYM_OFF + b, -1} : RowList{bins + off[b], nullptr, n_of(b)};
    };
    const int *tp = h->tileptr.as<int>();
    const int *tc = h->tilecol.as<int>();
    const unsigned *tm = h->tilemask.as<unsigned>();
    const int4 *arow = h->arow.as<int4>();
    int *scal = h->scal.as<int>();
    int *counts = h->Cp;
    const int cap_blocks = h->num_sms * 16;
    unsigned long long *probes =
        h->count_probes ? reinterpret_cast<unsigned long long *>(scal + SC_SYM_PROBES_LO) : nullptr;
    int n;
    cudaStream_t st;
    // twin flags of A's rows are known at this point only when A is B (family 1 computed them)
    const unsigned char *a_twins = nullptr;
    if (h->sym_twins)
        a_twins = (h->Ap == h->Bp && h->Ac == h->Bc) ? h->bsame.as<unsigned char>()
                                                     : (h->asame_early ? h->asame_buf.as<unsigned char>() : nullptr);
    h->have_bm_store = false;
    int frc = fork_bins(h, launched_kernels(launched_mask(off, SB_COUNT, SB_TINY), SB_TINY));
    if (frc)
        return frc;
    // big-row bins first (see launch_numeric_bins)
    if ((n = n_of(SB_H_GLOBAL)) > 0)
    {
        long long nt = ((long long)h->N + 31) / 32;
        long long ub = std::min<long long>(h->max_tileflop, nt);
        long long slots = 1LL << std::max(10, log2_ceil(ub + (ub >> 1) + 1));
        size_t slice = (size_t)slots * 2 * 4;
        int nblk = (int)std::min<long long>(std::min(n, h->num_sms * 2), std::max<long long>(1, (1LL << 30) / slice));
        CU(h->pool.ensure(slice * nblk));
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_block, nblk, kSymThreads, 0, list(SB_H_GLOBAL), h->Ap, h->Ac, tp, tc, tm,
               arow, counts, 0, h->pool.as<int>(), slots, scal, probes);
    }
    if ((n = n_of(SB_H_BLOCK_L)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_block, std::min(n, cap_blocks), kSymThreads, 2 * SB_H_BLOCK_L_SLOTS * 4,
               list(SB_H_BLOCK_L), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               log2_ceil(SB_H_BLOCK_L_SLOTS), (int *)nullptr, 0LL, scal, probes);
    }
    if ((n = n_of(SB_BM_BLOCK)) > 0)
    {
        int words = std::min<long long>(SB_BM_BLOCK_WORDS, ((long long)h->N + 31) / 32 + 1);
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_bitmap_block, std::min(n, cap_blocks), kSymThreads, words * 4, list(SB_BM_BLOCK),
               h->Ap, h->Ac, tp, tc, tm, arow, counts);
    }
    if ((n = n_of(SB_H_BLOCK_S)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_block, std::min(n, cap_blocks), kSymThreads, 2 * SB_H_BLOCK_S_SLOTS * 4,
               list(SB_H_BLOCK_S), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               log2_ceil(SB_H_BLOCK_S_SLOTS), (int *)nullptr, 0LL, scal, probes);
    }
    if ((n = n_of(SB_H_WARP)) > 0)
    {
        constexpr int G = 32, GPB = kSymThreads / G;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_group<G>, std::min(cdiv(n, GPB), cap_blocks), kSymThreads,
               GPB * 2 * SB_H_WARP_SLOTS * 4, list(SB_H_WARP), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               log2_ceil(SB_H_WARP_SLOTS), scal, probes);
    }
    if ((n = n_of(SB_H_G16)) > 0)
    {
        constexpr int G = 16, GPB = kSymThreads / G;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_group<G>, std::min(cdiv(n, GPB), cap_blocks), kSymThreads,
               GPB * 2 * SB_H_G16_SLOTS * 4, list(SB_H_G16), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               log2_ceil(SB_H_G16_SLOTS), scal, probes);
    }
    if ((n = n_of(SB_BM_WARP)) > 0)
    {
        constexpr int G = 32, GPB = kSymThreads / G;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_bitmap_group<G>, std::min(cdiv(n, 3 * GPB), cap_blocks), kSymThreads,
               GPB * SB_BM_WARP_WORDS * 4, list(SB_BM_WARP), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               SB_BM_WARP_WORDS, h->bsame.as<unsigned char>(), (unsigned *)nullptr, (int *)nullptr, a_twins, 0, scal);
    }
    if ((n = n_of(SB_H_G8)) > 0)
    {
        constexpr int G = 8, GPB = kSymThreads / G;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_hash_group<G>, std::min(cdiv(n, GPB), cap_blocks), kSymThreads,
               GPB * 2 * SB_H_G8_SLOTS * 4, list(SB_H_G8), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               log2_ceil(SB_H_G8_SLOTS), scal, probes);
    }
    if ((n = n_of(SB_BM_G8)) > 0)
    {
        constexpr int G = 8, GPB = kSymThreads / G;
        // keep the bitmaps of these rows for the numeric pass (NB_WIN_COMPACT) when that is cheap
        const size_t store_bytes = (size_t)n * SB_BM_STORE_WORDS * 4;
        h->have_bm_store = h->compact_rows && store_bytes <= ((size_t)1 << 31);
        if (h->have_bm_store)
        {
            CU(h->bm_store.ensure(store_bytes));
            CU(h->bm_slot.ensure(((size_t)h->M + 1) * 4));
        }
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_sym_bitmap_group<G>, std::min(cdiv(n, 3 * GPB), cap_blocks), kSymThreads,
               GPB * SB_BM_G8_WORDS * 4, list(SB_BM_G8), h->Ap, h->Ac, tp, tc, tm, arow, counts,
               SB_BM_G8_WORDS, h->bsame.as<unsigned char>(), h->have_bm_store ? h->bm_store.as<unsigned>() : nullptr,
               h->have_bm_store ? h->bm_slot.as<int>() : nullptr, a_twins, n, scal);
    }
    // the three cost classes of the thread-per-row kernel: adjacent bins, one launch over rows sorted by class
    static_assert(SB_TINY_S == SB_TINY + 1 && SB_TINY_M == SB_TINY + 2, "tiny classes must be adjacent bins");
    if ((n = off[SB_TINY + 3] - off[SB_TINY]) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        RowList tl = spec ? RowList{bins, scal + SC_SYM_OFF + SB_TINY, -1, nullptr, 3}
                          : RowList{bins + off[SB_TINY], nullptr, n, nullptr, 1};
        LAUNCH_ON(h, st, k_sym_tiny, std::min(cdiv(n, kTinyThreads), cap_blocks), kTinyThreads, 0, tl,
                  h->Ap, h->Ac, tp, tc, tm, counts);
    }
    return join_bins(h);
}

// ---- family 4 launches ------------------------------------------------------------------
// shared memory of k_num_hash_list for a table of S slots (see the layout in the kernel)
template <typename T>
size_t hash_list_smem(int S)
{
    return ((size_t)S * (sizeof(T) + 4) + (size_t)(S / 4 + 4) * 4 + (size_t)(S / 8) * 5 * (4 + 2 + 2) + 15) / 16 * 16;
}

template <typename T>
int launch_numeric_bins(mhb_context *h, const T *Av, const T *Bv, int *Cc, T *Cv, bool spec = false,
                        const int *chunk_off = nullptr, int chunk = 0, int nchunks = 1)
{
    // chunk_off (host-buffer path): launch only the rows of row chunk `chunk`; the piece of every
    // bin's list comes from the device-side table k_chunk_bounds filled
    // spec: the fused call (do_spgemm_into) has not read this call's bin sizes; h->num_off and
    // h->max_rownnz are the previous call's and only size grids and scratch, the kernels take
    // their ranges from the device-side offsets and stand down when the capacity gate is set
    const int *off = h->num_off;
    auto n_of = [&](int b) { return off[b + 1] - off[b]; };
    const int *bins = h->bins_num.as<int>();
    auto list = [&](int b) {
        if (chunk_off)
            return RowList{bins, chunk_off + b * (nchunks + 1) + chunk, -1, nullptr};
        return spec ? RowList{bins, h->scal.as<int>() + SC_NUM_OFF + b, -1, h->scal.as<int>() + SC_GATE}
                    : RowList{bins + off[b], nullptr, n_of(b), nullptr};
    };
    const int4 *arow = h->arow.as<int4>();
    int *scal = h->scal.as<int>();
    const int cap_blocks = h->num_sms * 16;
    unsigned long long *probes = h->count_probes ? reinterpret_cast<unsigned long long *>(scal + SC_PROBES_LO) : nullptr;
    const int *Ap = h->Ap, *Ac = h->Ac, *Bp = h->Bp, *Bc = h->Bc, *Cp = h->Cp;
    int n;
    cudaStream_t st;
    int frc = fork_bins(h, launched_kernels(launched_mask(off, NB_COUNT, NB_TINY), NB_TINY));
    if (frc)
        return frc;
    // k_num_hash_list over one bin.  threads = 0: four rows (warps) per 128-thread block, each warp
    // with a table of its own; otherwise one block of `threads` threads per row.
    auto launch_hash_list = [&](int bin, int slots, int threads, int grid_cap) -> int {
        const int rows_in_bin = off[bin + 1] - off[bin];
        const int table = (int)hash_list_smem<T>(slots);
        if (threads == 0)
        {
            auto kern = k_num_hash_list<T, true>;
            LAUNCH_ON(h, st, kern, std::min(cdiv(rows_in_bin, 4), grid_cap), 128, 4 * table, list(bin),
                      Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, log2_ceil(slots), scal, table, probes);
        }
        else
        {
            auto kern = k_num_hash_list<T, false>;
            LAUNCH_ON(h, st, kern, std::min(rows_in_bin, grid_cap), threads, table, list(bin), Ap, Ac,
                      Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, log2_ceil(slots), scal, 0, probes);
        }
        return MHB_OK;
    };
    // launch order: bins with the fewest, largest rows first, so that their long-running
    // blocks start at once and overlap the bulk bins instead of forming a tail
    if ((n = n_of(NB_H_GLOBAL)) > 0)
    {
        long long slots = 1LL << std::max(10, log2_ceil(2LL * h->max_rownnz));
        size_t slice = (size_t)slots * (sizeof(T) + 4);
        int nblk = (int)std::min<long long>(std::min(n, h->num_sms * 2), std::max<long long>(1, (1LL << 31) / slice));
        CU(h->pool.ensure(slice * nblk));
        if (int e_ = next_bin_stream(h, &st)) return e_;
        // shared memory only for the sort of the compacted keys (6 B per entry + buckets)
        const int sort_smem = (int)std::min<long long>(MHB_SMEM_MAX - 1024, 7LL * h->max_rownnz + 1024);
        LAUNCH_ON(h, st, k_num_hash_block<T>, nblk, 1024, sort_smem, list(NB_H_GLOBAL), Ap, Ac, Av, Bp, Bc, Bv, arow,
                  Cp, Cc, Cv, 0, h->pool.as<unsigned char>(), slots, scal, sort_smem, probes);
    }
    if ((n = n_of(NB_H_BLOCK_L)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_num_hash_block<T>, std::min(n, cap_blocks), 1024, NB_H_BLOCK_L_SLOTS * (sizeof(T) + 4),
               list(NB_H_BLOCK_L), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, log2_ceil(NB_H_BLOCK_L_SLOTS),
               (unsigned char *)nullptr, 0LL, scal, 0, probes);
    }
    if ((n = n_of(NB_WIN_BLOCK_L)) > 0)
    {
        int wcap = (int)std::min<long long>(NB_WIN_BLOCK_L_COLS, (((long long)h->N + 31) / 32) * 32);
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_num_win_block<T>, std::min(n, cap_blocks), 1024, wcap * sizeof(T) + (wcap / 32) * 8,
               list(NB_WIN_BLOCK_L), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, wcap);
    }
    if ((n = n_of(NB_H_BLOCK_M)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_BLOCK_M, NB_H_BLOCK_M_SLOTS, 512, cap_blocks)) return e_;
        }
        else
            LAUNCH_ON(h, st, k_num_hash_block<T>, std::min(n, cap_blocks), 1024, NB_H_BLOCK_L_SLOTS * (sizeof(T) + 4),
                      list(NB_H_BLOCK_M), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv,
                      log2_ceil(NB_H_BLOCK_L_SLOTS), (unsigned char *)nullptr, 0LL, scal, 0, probes);
    }
    if ((n = n_of(NB_H_BLOCK_S)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_BLOCK_S, NB_H_BLOCK_S_SLOTS, 256, cap_blocks)) return e_;
        }
        else
            LAUNCH_ON(h, st, k_num_hash_block<T>, std::min(n, cap_blocks), 256, NB_H_BLOCK_S_SLOTS * (sizeof(T) + 4),
                      list(NB_H_BLOCK_S), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv,
                      log2_ceil(NB_H_BLOCK_S_SLOTS), (unsigned char *)nullptr, 0LL, scal, 0, probes);
    }
    if ((n = n_of(NB_H_BLOCK_XS)) > 0)
    {
        // 2 048-slot tables: 36 KB per row instead of the 4 096-slot bin's 74 KB, i.e. twice the rows in flight
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_BLOCK_XS, NB_H_BLOCK_XS_SLOTS, 128, cap_blocks * 2)) return e_;
        }
        else
            LAUNCH_ON(h, st, k_num_hash_block<T>, std::min(n, cap_blocks), 256, NB_H_BLOCK_S_SLOTS * (sizeof(T) + 4),
                      list(NB_H_BLOCK_XS), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv,
                      log2_ceil(NB_H_BLOCK_S_SLOTS), (unsigned char *)nullptr, 0LL, scal, 0, probes);
    }
    if ((n = n_of(NB_WIN_BLOCK_S)) > 0)
    {
        int wcap = NB_WIN_BLOCK_S_COLS;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_num_win_block<T>, std::min(n, cap_blocks), 256, wcap * sizeof(T) + (wcap / 32) * 8,
               list(NB_WIN_BLOCK_S), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, wcap);
    }
    if ((n = n_of(NB_H_WARP_L)) > 0)
    {
        // 1 024-slot tables: two warps share one table (64 threads) so that ~28-36 warps stay resident
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_WARP_L, NB_H_WARP_L_SLOTS, 64, cap_blocks * 4)) return e_;
        }
        else
            LAUNCH_ON(h, st, k_num_hash_block<T>, std::min(n, cap_blocks * 4), 64, NB_H_WARP_L_SLOTS * (sizeof(T) + 4),
                      list(NB_H_WARP_L), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv,
                      log2_ceil(NB_H_WARP_L_SLOTS), (unsigned char *)nullptr, 0LL, scal, 0, probes);
    }
    if ((n = n_of(NB_WIN_COMPACT)) > 0)
    {
        constexpr int WPB = kRowTwinThreads / 32;
        const size_t smem = (size_t)WPB * 3 * (NB_WIN_COMPACT_MAXN + 34) * sizeof(T) +
                            (size_t)WPB * (SB_BM_STORE_WORDS + 2 + 34) * 8;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, k_num_compact_rowtwins<T>, std::min(cdiv(cdiv(n, 3), WPB), cap_blocks), kRowTwinThreads, smem,
                  list(NB_WIN_COMPACT), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, NB_WIN_COMPACT_MAXN,
                  h->bsame.as<unsigned char>(), h->asame, h->bm_store.as<unsigned>(), h->bm_slot.as<int>());
    }
    if ((n = n_of(NB_WIN_WARP)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->row_twins)
        {
            constexpr int WPB = kRowTwinThreads / 32;
            const size_t smem = (size_t)WPB * 3 * (NB_WIN_WARP_COLS + 34) * sizeof(T) + (size_t)WPB * 34 * 8;
            LAUNCH_ON(h, st, k_num_win_rowtwins<T>, std::min(cdiv(cdiv(n, 3), WPB), cap_blocks), kRowTwinThreads, smem,
                      list(NB_WIN_WARP), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv, NB_WIN_WARP_COLS,
                      h->bsame.as<unsigned char>(), h->asame);
        }
        else
        {
            constexpr int G = 32, GPB = kNumGroupThreads / G;
            auto kern = k_num_win_group<G, T>;
            LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
                      GPB * (NB_WIN_WARP_COLS * sizeof(T) + (G + 2) * 16), list(NB_WIN_WARP), Ap, Ac, Av, Bp, Bc,
                      Bv, arow, Cp, Cc, Cv, NB_WIN_WARP_COLS, h->bsame.as<unsigned char>());
        }
    }
    if ((n = n_of(NB_H_WARP_M)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_WARP_M, NB_H_WARP_M_SLOTS, 32, cap_blocks * 8)) return e_;
        }
        else
        {
            constexpr int G = 32, GPB = kNumGroupThreads / G;
            auto kern = k_num_hash_group<G, T>;
            LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
                      GPB * NB_H_WARP_M_SLOTS * (sizeof(T) + 4), list(NB_H_WARP_M), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp,
                      Cc, Cv, log2_ceil(NB_H_WARP_M_SLOTS), scal, probes);
        }
    }
    if ((n = n_of(NB_H_WARP_S)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_WARP_S, NB_H_WARP_S_SLOTS, 0, cap_blocks * 2)) return e_;
        }
        else
        {
            constexpr int G = 32, GPB = kNumGroupThreads / G;
            auto kern = k_num_hash_group<G, T>;
            LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
                      GPB * NB_H_WARP_S_SLOTS * (sizeof(T) + 4), list(NB_H_WARP_S), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp,
                      Cc, Cv, log2_ceil(NB_H_WARP_S_SLOTS), scal, probes);
        }
    }
    if ((n = n_of(NB_H_WARP_XS)) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (h->claim_list)
        {
            if (int e_ = launch_hash_list(NB_H_WARP_XS, NB_H_WARP_XS_SLOTS, 0, cap_blocks * 2)) return e_;
        }
        else
        {
            constexpr int G = 32, GPB = kNumGroupThreads / G;
            auto kern = k_num_hash_group<G, T>;
            LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
                      GPB * NB_H_WARP_XS_SLOTS * (sizeof(T) + 4), list(NB_H_WARP_XS), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp,
                      Cc, Cv, log2_ceil(NB_H_WARP_XS_SLOTS), scal, probes);
        }
    }
    if ((n = n_of(NB_WIN_G8)) > 0)
    {
        constexpr int G = 8, GPB = kNumGroupThreads / G;
        auto kern = k_num_win_group<G, T>;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
               GPB * (NB_WIN_G8_COLS * sizeof(T) + (G + 2) * 16), list(NB_WIN_G8), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc,
               Cv, NB_WIN_G8_COLS, h->bsame.as<unsigned char>());
    }
    if ((n = n_of(NB_H_G8)) > 0)
    {
        constexpr int G = 8, GPB = kNumGroupThreads / G;
        auto kern = k_num_hash_group<G, T>;
        if (int e_ = next_bin_stream(h, &st)) return e_;
        LAUNCH_ON(h, st, kern, std::min(cdiv(n, GPB), cap_blocks), kNumGroupThreads,
               GPB * NB_H_G8_SLOTS * (sizeof(T) + 4), list(NB_H_G8), Ap, Ac, Av, Bp, Bc, Bv, arow, Cp, Cc, Cv,
               log2_ceil(NB_H_G8_SLOTS), scal, probes);
    }
    // the three cost classes of the thread-per-row kernel: adjacent bins, one launch over rows sorted by class
    static_assert(NB_TINY_S == NB_TINY + 1 && NB_TINY_M == NB_TINY + 2, "tiny classes must be adjacent bins");
    if ((n = off[NB_TINY + 3] - off[NB_TINY]) > 0)
    {
        if (int e_ = next_bin_stream(h, &st)) return e_;
        if (chunk_off) // a row chunk is a separate piece of each class's list: one launch per class
        {
            for (int tb = NB_TINY; tb < NB_TINY + 3; ++tb)
                if (n_of(tb) > 0)
                    LAUNCH_ON(h, st, k_num_tiny<T>, std::min(cdiv(n_of(tb), kTinyRowThreads), cap_blocks), kTinyRowThreads,
                              NB_TINY_MAX * kTinyRowThreads * (sizeof(T) + 4), list(tb), Ap, Ac, Av, Bp, Bc, Bv, Cp, Cc, Cv);
        }
        else
        {
            RowList tl = spec ? RowList{bins, scal + SC_NUM_OFF + NB_TINY, -1, scal + SC_GATE, 3}
                              : RowList{bins + off[NB_TINY], nullptr, n, nullptr, 1};
            LAUNCH_ON(h, st, k_num_tiny<T>, std::min(cdiv(n, kTinyRowThreads), cap_blocks), kTinyRowThreads,
                      NB_TINY_MAX * kTinyRowThreads * (sizeof(T) + 4), tl, Ap, Ac, Av, Bp, Bc, Bv, Cp, Cc, Cv);
        }
    }
    return join_bins(h);
}

int set_kernel_attributes(mhb_context *h)
{
    CU(allow_smem(k_sym_bitmap_group<32>, 8 * SB_BM_WARP_WORDS * 4));
    CU(allow_smem(k_sym_bitmap_block, MHB_SMEM_MAX - 256));
    CU(allow_smem(k_sym_hash_block, 2 * SB_H_BLOCK_L_SLOTS * 4));
    CU(allow_smem(k_num_win_group<8, double>, 32 * (NB_WIN_G8_COLS * 8 + 160)));
    CU(allow_smem(k_num_win_group<32, double>, 8 * (NB_WIN_WARP_COLS * 8 + 544)));
    CU(allow_smem(k_num_win_group<8, float>, 32 * (NB_WIN_G8_COLS * 4 + 160)));
    CU(allow_smem(k_num_win_group<32, float>, 8 * (NB_WIN_WARP_COLS * 4 + 544)));
    CU(allow_smem(k_num_hash_list<double, false>, (int)hash_list_smem<double>(NB_H_BLOCK_M_SLOTS)));
    CU(allow_smem(k_num_hash_list<float, false>, (int)hash_list_smem<float>(NB_H_BLOCK_M_SLOTS)));
    CU(allow_smem(k_num_tiny<double>, NB_TINY_MAX * kTinyRowThreads * 12));
    CU(allow_smem(k_num_tiny<float>, NB_TINY_MAX * kTinyRowThreads * 8));
    CU(allow_smem(k_num_compact_rowtwins<double>, 4 * 3 * (NB_WIN_COMPACT_MAXN + 34) * 8 + 4 * 100 * 8));
    CU(allow_smem(k_num_compact_rowtwins<float>, 4 * 3 * (NB_WIN_COMPACT_MAXN + 34) * 4 + 4 * 100 * 8));
    CU(allow_smem(k_num_win_rowtwins<double>, 4 * 3 * (NB_WIN_WARP_COLS + 34) * 8 + 4 * 34 * 8));
    CU(allow_smem(k_num_win_rowtwins<float>, 4 * 3 * (NB_WIN_WARP_COLS + 34) * 4 + 4 * 34 * 8));
    CU(allow_smem(k_num_win_block<double>, MHB_SMEM_MAX - 256));
    CU(allow_smem(k_num_win_block<float>, MHB_SMEM_MAX - 256));
    CU(allow_smem(k_num_hash_group<32, double>, 8 * NB_H_WARP_L_SLOTS * 12));
    CU(allow_smem(k_num_hash_group<32, float>, 8 * NB_H_WARP_L_SLOTS * 8));
    CU(allow_smem(k_num_hash_block<double>, MHB_SMEM_MAX - 1024));
    CU(allow_smem(k_num_hash_block<float>, MHB_SMEM_MAX - 1024));
    return MHB_OK;
}

float ev_ms(mhb_context *h, int a, int b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess)
    {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

int finish_symbolic(mhb_context *h, const int *hs, long long *nnzC_out);
int prepare_a_twins(mhb_context *h);

int do_symbolic(mhb_context *h, int M, int K, int N, int nnzA, const int *Ap, const int *Ac, int nnzB,
                const int *Bp, const int *Bc, int *Cp, long long *nnzC_out, bool allow_spec = true,
                long long fused_capacity = -1)
{
    h->fused_pending = false;
    if (h->into.active && h->into.pending)
        return fail(h, MHB_ERR_ARG, "a mhb_spgemm_into_begin_* is outstanding on this handle: call mhb_spgemm_into_end first");
    if (M < 0 || K < 0 || N < 0 || nnzA < 0 || nnzB < 0)
        return fail(h, MHB_ERR_ARG, "negative dimension");
    if (!Ap || !Bp || !Cp || (nnzA > 0 && !Ac) || (nnzB > 0 && !Bc))
        return fail(h, MHB_ERR_ARG, "null CSR pointer");
    CU(cudaSetDevice(h->device));
    h->have_pattern = false;
    h->launches = 0;
    h->ev_sym_valid = h->ev_num_valid = false;
    std::memset(&h->timing, 0, sizeof(h->timing));
    std::memset(&h->stats, 0, sizeof(h->stats));
    h->M = M, h->K = K, h->N = N, h->nnzA = nnzA, h->nnzB = nnzB;
    h->Ap = Ap, h->Ac = Ac, h->Bp = Bp, h->Bc = Bc, h->Cp = Cp;
    CU(cudaEventRecord(h->ev[EV_START], h->stream));
    bool grew = false;
    int rc = ensure_workspace(h, M, K, nnzB, &grew);
    if (rc)
        return rc;
    int *scal = h->scal.as<int>();
    int *hs = h->h_scal.as<int>();
    rc = zero_scal(h, h->mask_onepass ? scal_bytes(nnzB) : (size_t)SC_COUNT * 4);
    if (rc)
        return rc;
    CU(cudaEventRecord(h->ev[EV_ALLOC], h->stream));
    // twin rows of A (same column list as the previous row): B's flags when A is B; otherwise
    // compared now on a helper stream, hidden behind the mask build that only needs B
    h->asame_early = false;
    if (!(Ap == Bp && Ac == Bc) && h->sym_twins && M > 0)
    {
        CU(h->asame_buf.ensure((size_t)M + 1));
        cudaStream_t side = h->serial ? h->stream : h->aux[0];
        if (!h->serial)
        {
            CU(cudaEventRecord(h->ev_fork, h->stream));
            CU(cudaStreamWaitEvent(side, h->ev_fork, 0));
        }
        LAUNCH_ON(h, side, k_rows_same_cols, cdiv((long long)M * 8, 256), 256, 0, M, Ap, Ac,
                  h->asame_buf.as<unsigned char>());
        if (!h->serial)
            CU(cudaEventRecord(h->ev_join[0], side));
        h->asame_early = true;
    }

    // family 1: B mask matrix
    rc = build_mask_matrix(h, K, nnzB, Bp, Bc);
    if (rc)
        return rc;
    CU(cudaEventRecord(h->ev[EV_MASK], h->stream));

    // family 2: row metrics + symbolic bins
    if (M > 0)
    {
        double avg = (double)nnzA / M;
        // twin flags of A's rows: B's flags (family 1) if A is B, else the comparison that ran on the
        // helper stream beside the mask build (long finished: joined here instead of before the symbolic bins)
        const unsigned char *tw = nullptr;
        if (h->sym_twins && Ap == Bp && Ac == Bc)
            tw = h->bsame.as<unsigned char>();
        else if (h->sym_twins && h->asame_early)
        {
            if (!h->serial)
                CU(cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
            tw = h->asame_buf.as<unsigned char>();
        }
        if (avg > 12.0)
            LAUNCH(h, k_arow_metrics<32>, std::min(cdiv((long long)M * 32, 256), h->num_sms * 16), 256, 0, M, Ap, Ac, h->binfo.as<int4>(),
                   h->arow.as<int4>(), h->binid.as<unsigned char>(), Cp, scal, h->force_sym, tw);
        else
            LAUNCH(h, k_arow_metrics<4>, std::min(cdiv((long long)M * 4, 256), h->num_sms * 16), 256, 0, M, Ap, Ac, h->binfo.as<int4>(),
                   h->arow.as<int4>(), h->binid.as<unsigned char>(), Cp, scal, h->force_sym, tw);
    }
    else
        CU(cudaMemsetAsync(Cp, 0, sizeof(int), h->stream));
    rc = run_binning(h, M, SB_COUNT, h->bins_sym.as<int>(), SC_SYM_SIZE, SC_SYM_OFF);
    if (rc)
        return rc;
    // Speculative launch: when the previous call on this handle had the same shape, its bin sizes
    // size the grids and scratch of this call's symbolic kernels, the kernels read their row
    // ranges from the device-side offsets, and host read #1 (a blocking round trip in the middle
    // of the pipeline) is dropped; the guess is verified at the hand-off read below, and a miss
    // (a bin that was not launched turned out populated, or a scratch capacity was exceeded)
    // re-runs the phase the ordinary way.
    const int key[5] = {M, K, N, nnzA, nnzB};
    const bool spec = allow_spec && h->speculate && h->spec_ok && std::memcmp(key, h->spec_key, sizeof(key)) == 0;
    h->spec_ok = false;
    if (!spec)
    {
        CU(cudaMemcpyAsync(hs, scal, SC_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->ev[EV_SYMBIN], h->stream));
        CU(cudaStreamSynchronize(h->stream)); // read #1: symbolic bin sizes
        std::memcpy(h->sym_off, hs + SC_SYM_OFF, sizeof(h->sym_off));
        h->max_tileflop = hs[SC_MAX_TILEFLOP];
    }
    else
    {
        ++h->spec_calls;
        CU(cudaEventRecord(h->ev[EV_SYMBIN], h->stream));
    }
    int launched[MHB_MAX_BINS + 1];
    std::memcpy(launched, h->sym_off, sizeof(launched));
    const int planned_tileflop = h->max_tileflop;

    // family 3: nnz per C row
    if (h->asame_early && !h->serial)
        CU(cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
    rc = launch_symbolic_bins(h, spec);
    if (rc)
        return rc;
    CU(cudaEventRecord(h->ev[EV_SYM], h->stream));

    // family 2 again: numeric bins from the exact row sizes, then the row-offset scan
    if (M > 0)
        LAUNCH(h, k_classify_num, cdiv(M, 256), 256, 0, M, Ap, Cp, h->arow.as<int4>(), h->binid.as<unsigned char>(),
               scal, h->force_num, h->force_sym, h->have_bm_store ? 1 : 0);
    rc = run_binning(h, M, NB_COUNT, h->bins_num.as<int>(), SC_NUM_SIZE, SC_NUM_OFF);
    if (rc)
        return rc;
    CU(cudaEventRecord(h->ev[EV_NUMBIN], h->stream));
    rc = run_scan(h, LoadInt{Cp}, M, Cp, 1, reinterpret_cast<long long *>(scal + SC_NNZC_LO));
    if (rc)
        return rc;
    if (spec && fused_capacity >= 0)
    {
        // fused call: no host read here.  The verdict on this call's speculation is formed on the
        // device (k_fused_gate) and the caller launches the numeric kernels behind it; the host reads
        // the scalars once, after the numeric phase (finish_symbolic).
        const unsigned sym_mask = launched_mask(launched, SB_COUNT, SB_TINY);
        const unsigned num_mask = launched_mask(h->num_off, NB_COUNT, NB_TINY);
        LAUNCH(h, k_fused_gate, 1, 32, 0, scal, sym_mask, planned_tileflop, num_mask, h->max_rownnz,
               std::min<long long>(fused_capacity, h->nnz_limit), (int)SB_H_GLOBAL, (int)NB_H_GLOBAL, (int)SB_COUNT,
               (int)NB_COUNT);
        CU(cudaEventRecord(h->ev[EV_HANDOFF], h->stream));
        std::memcpy(h->fused_launched, launched, sizeof(launched));
        h->fused_planned_tileflop = planned_tileflop;
        h->fused_pending = true;
        return MHB_OK;
    }
    CU(cudaMemcpyAsync(hs, scal, SC_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev[EV_HANDOFF], h->stream));
    CU(cudaStreamSynchronize(h->stream)); // read #2: nnz(C) + numeric bin sizes (the hand-off)
    rc = check_dev_error(h, hs);
    if (rc)
        return rc;
    if (spec)
    {
        bool miss = hs[SC_SPEC_MISS] != 0;
        const unsigned sym_mask = launched_mask(launched, SB_COUNT, SB_TINY);
        for (int b = 1; b < SB_COUNT; ++b) // bin 0 (rows without products) has no kernel
            miss |= hs[SC_SYM_SIZE + b] > 0 && !((sym_mask >> b) & 1u);
        // the global tile-hash pool was sized from the previous call's largest row (checked per row in the kernel too)
        miss |= hs[SC_SYM_SIZE + SB_H_GLOBAL] > 0 && hs[SC_MAX_TILEFLOP] > planned_tileflop;
        if (miss)
        {
            ++h->spec_misses;
            return do_symbolic(h, M, K, N, nnzA, Ap, Ac, nnzB, Bp, Bc, Cp, nnzC_out, false);
        }
    }
    rc = finish_symbolic(h, hs, nnzC_out);
    if (rc)
        return rc;
    rc = prepare_a_twins(h);
    if (rc)
        return rc;
    h->have_pattern = true;
    if (h->verbose)
        std::printf("C.nnz = %lld\n", h->nnzC); // the reference's print (src/main.cu:58), opt-in
    return MHB_OK;
}

// The host has this call's scalar block in hs: bin sizes, nnz(C), statistics, stage times.
int finish_symbolic(mhb_context *h, const int *hs, long long *nnzC_out)
{
    const int key[5] = {h->M, h->K, h->N, h->nnzA, h->nnzB};
    std::memcpy(h->sym_off, hs + SC_SYM_OFF, sizeof(h->sym_off));
    h->max_tileflop = hs[SC_MAX_TILEFLOP];
    std::memcpy(&h->stats.intprod, hs + SC_INTPROD_LO, 8);
    std::memcpy(&h->stats.tileflop, hs + SC_TILEFLOP_LO, 8);
    std::memcpy(&h->stats.ntiles_B, hs + SC_NTILES_LO, 8);
    std::memcpy(h->stats.sym_bin_size, hs + SC_SYM_SIZE, sizeof(int) * MHB_MAX_BINS);
    h->stats.speculative_launches = h->spec_calls;
    h->stats.speculative_misses = h->spec_misses;
    h->stats.fused_calls = h->fused_calls;
    std::memcpy(h->spec_key, key, sizeof(key));
    h->spec_ok = true;
    std::memcpy(h->num_off, hs + SC_NUM_OFF, sizeof(h->num_off));
    h->max_rownnz = hs[SC_MAX_ROWNNZ];
    std::memcpy(&h->nnzC, hs + SC_NNZC_LO, 8);
    std::memcpy(h->stats.num_bin_size, hs + SC_NUM_SIZE, sizeof(int) * MHB_MAX_BINS);
    h->stats.nnzC = h->nnzC;
    h->stats.gpu_launches = h->launches;
    std::memcpy(&h->stats.sym_hash_probes, hs + SC_SYM_PROBES_LO, 8);
    *nnzC_out = h->nnzC;
    h->ev_sym_valid = true;
    h->timing.mem_alloc = ev_ms(h, EV_START, EV_ALLOC);
    h->timing.form_mask_matrix_B = ev_ms(h, EV_ALLOC, EV_MASK);
    h->timing.symbolic_binning = ev_ms(h, EV_MASK, EV_SYMBIN);
    h->timing.calculate_C_nnz = ev_ms(h, EV_SYMBIN, EV_SYM);
    h->timing.numeric_binning = ev_ms(h, EV_SYM, EV_NUMBIN);
    h->timing.malloc_C_col_val = ev_ms(h, EV_NUMBIN, EV_HANDOFF);
    h->timing.total = ev_ms(h, EV_START, EV_HANDOFF);
    if (h->nnzC > h->nnz_limit)
        return fail(h, MHB_ERR_OVERFLOW,
                    "nnz(C) = " + std::to_string(h->nnzC) + " exceeds the int32 CSR contract; shard the rows of A");
    return MHB_OK;
}

// Twin rows of A (same column list as the previous row), needed only by the row-twin numeric
// kernels: B's flags when A is B, else compared now -- off the symbolic critical path.  Reads
// h->num_off (in a fused call: the previous call's, like the numeric launch that follows).
int prepare_a_twins(mhb_context *h)
{
    const int M = h->M;
    if (h->Ap == h->Bp && h->Ac == h->Bc)
        h->asame = h->bsame.as<unsigned char>();
    else
    {
        CU(h->asame_buf.ensure((size_t)M + 1));
        const bool needed = (h->num_off[NB_WIN_COMPACT + 1] > h->num_off[NB_WIN_COMPACT]) ||
                            (h->row_twins && h->num_off[NB_WIN_WARP + 1] > h->num_off[NB_WIN_WARP]);
        if (M > 0 && needed && !h->asame_early)
            LAUNCH(h, k_rows_same_cols, cdiv((long long)M * 8, 256), 256, 0, M, h->Ap, h->Ac,
                   h->asame_buf.as<unsigned char>());
        h->asame = h->asame_buf.as<unsigned char>();
    }
    return MHB_OK;
}

template <typename T>
int do_numeric(mhb_context *h, const T *Av, const T *Bv, int *Cc, T *Cv, bool sync)
{
    if (h->into.active && h->into.pending)
        return fail(h, MHB_ERR_ARG, "a mhb_spgemm_into_begin_* is outstanding on this handle: call mhb_spgemm_into_end first");
    if (!h->have_pattern)
        return fail(h, MHB_ERR_ARG, "mhb_numeric called without a successful mhb_symbolic");
    if (h->nnzC > 0 && (!Av || !Bv || !Cc || !Cv))
        return fail(h, MHB_ERR_ARG, "null value / output pointer");
    CU(cudaSetDevice(h->device));
    int before = h->launches;
    if (h->count_probes)
        CU(cudaMemsetAsync(h->scal.as<int>() + SC_PROBES_LO, 0, 8, h->stream));
    CU(cudaEventRecord(h->ev[EV_NUM0], h->stream));
    int rc = launch_numeric_bins<T>(h, Av, Bv, Cc, Cv);
    if (rc)
        return rc;
    CU(cudaEventRecord(h->ev[EV_NUM1], h->stream));
    h->stats.gpu_launches = h->launches;
    (void)before;
    if (sync)
    {
        int *hs = h->h_scal.as<int>();
        CU(cudaMemcpyAsync(hs + SC_ERROR, h->scal.as<int>() + SC_ERROR, 6 * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        rc = check_dev_error(h, hs);
        if (rc)
            return rc;
        std::memcpy(&h->stats.hash_probes, hs + SC_PROBES_LO, 8);
        h->timing.numeric = ev_ms(h, EV_NUM0, EV_NUM1);
        if (h->ev_sym_valid)
            h->timing.total = ev_ms(h, EV_START, EV_HANDOFF) + h->timing.numeric;
    }
    return MHB_OK;
}

// begin: validates, launches (in steady state: everything, without a host read), returns without
// waiting.  end: the one synchronisation, the verdict of the gate, the redo on a miss.
template <typename T>
int into_begin(mhb_context *h, int M, int K, int N, int nnzA, const int *Ap, const int *Ac, const T *Av, int nnzB,
               const int *Bp, const int *Bc, const T *Bv, int *Cp, int *Cc, T *Cv, long long capacity)
{
    IntoCall &ic = h->into;
    if (ic.active && ic.pending)
        return fail(h, MHB_ERR_ARG, "a mhb_spgemm_into_begin_* is outstanding on this handle: call mhb_spgemm_into_end first");
    ic = IntoCall{};
    if (capacity < 0 || (capacity > 0 && (!Cc || !Cv)))
        return fail(h, MHB_ERR_ARG, "null output pointer / negative capacity");
    if ((nnzA > 0 && !Av) || (nnzB > 0 && !Bv))
        return fail(h, MHB_ERR_ARG, "null value pointer");
    ic.M = M, ic.K = K, ic.N = N, ic.nnzA = nnzA, ic.nnzB = nnzB;
    ic.Ap = Ap, ic.Ac = Ac, ic.Av = Av, ic.Bp = Bp, ic.Bc = Bc, ic.Bv = Bv, ic.Cp = Cp, ic.Cc = Cc, ic.Cv = Cv;
    ic.capacity = capacity, ic.vbytes = (int)sizeof(T);
    int rc = do_symbolic(h, M, K, N, nnzA, Ap, Ac, nnzB, Bp, Bc, Cp, &ic.nnzC, true, capacity);
    if (rc)
        return rc;
    ic.active = true;
    if (h->fused_pending)
    {
        h->fused_pending = false;
        ++h->fused_calls;
        ic.pending = true;
        rc = prepare_a_twins(h);
        if (rc)
            return rc;
        CU(cudaEventRecord(h->ev[EV_NUM0], h->stream));
        rc = launch_numeric_bins<T>(h, Av, Bv, Cc, Cv, true);
        if (rc)
            return rc;
        CU(cudaEventRecord(h->ev[EV_NUM1], h->stream));
        CU(cudaMemcpyAsync(h->h_scal.as<int>(), h->scal.as<int>(), SC_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
        return MHB_OK;
    }
    // ordinary path (first call of a shape): the symbolic phase has synchronised and nnz(C) is known
    if (ic.nnzC > capacity)
    {
        ic.over = true; // reported by end
        return MHB_OK;
    }
    return do_numeric<T>(h, Av, Bv, Cc, Cv, false);
}

int numeric_finish(mhb_context *h)
{
    int *hs = h->h_scal.as<int>();
    CU(cudaMemcpyAsync(hs + SC_ERROR, h->scal.as<int>() + SC_ERROR, 6 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int rc = check_dev_error(h, hs);
    if (rc)
        return rc;
    std::memcpy(&h->stats.hash_probes, hs + SC_PROBES_LO, 8);
    h->timing.numeric = ev_ms(h, EV_NUM0, EV_NUM1);
    if (h->ev_sym_valid)
        h->timing.total = ev_ms(h, EV_START, EV_HANDOFF) + h->timing.numeric;
    return MHB_OK;
}

int capacity_error(mhb_context *h, long long nnz, long long capacity)
{
    return fail(h, MHB_ERR_CAPACITY, "nnz(C) = " + std::to_string(nnz) + " exceeds the capacity of the caller's C arrays (" +
                                         std::to_string(capacity) + "); row_ptr is valid, grow C.col / C.val and call again");
}

template <typename T>
int into_end(mhb_context *h, long long *nnzC)
{
    IntoCall ic = h->into;
    h->into.active = false;
    if (!nnzC)
        return fail(h, MHB_ERR_ARG, "null nnzC");
    *nnzC = 0;
    if (!ic.active || ic.vbytes != (int)sizeof(T))
        return fail(h, MHB_ERR_ARG, "mhb_spgemm_into_end without a matching mhb_spgemm_into_begin_*");
    const T *Av = static_cast<const T *>(ic.Av), *Bv = static_cast<const T *>(ic.Bv);
    T *Cv = static_cast<T *>(ic.Cv);
    if (!ic.pending)
    {
        *nnzC = ic.nnzC;
        if (ic.over)
            return capacity_error(h, ic.nnzC, ic.capacity);
        return numeric_finish(h);
    }
    int *hs = h->h_scal.as<int>();
    CU(cudaStreamSynchronize(h->stream)); // the one host read of the call
    const int gate = hs[SC_GATE];
    // SC_SPEC_MISS raised after the gate: a pool row of the numeric phase outgrew the planned table
    const bool late_miss = gate == 0 && hs[SC_SPEC_MISS] != 0;
    if ((gate & (GATE_SYM_MISS | GATE_NUM_MISS)) || late_miss)
    {
        ++h->spec_misses; // redo the ordinary way
        int rc = do_symbolic(h, ic.M, ic.K, ic.N, ic.nnzA, ic.Ap, ic.Ac, ic.nnzB, ic.Bp, ic.Bc, ic.Cp, nnzC, false);
        if (rc)
            return rc;
        if (*nnzC > ic.capacity)
            return capacity_error(h, *nnzC, ic.capacity);
        return do_numeric<T>(h, Av, Bv, ic.Cc, Cv, true);
    }
    int rc = check_dev_error(h, hs);
    if (rc)
        return rc;
    rc = finish_symbolic(h, hs, nnzC);
    if (rc)
        return rc;
    if (gate & GATE_CAPACITY)
        return capacity_error(h, h->nnzC, ic.capacity);
    h->have_pattern = true; // (A's twin flags were formed in begin, for the same set of bins)
    h->stats.gpu_launches = h->launches;
    std::memcpy(&h->stats.hash_probes, hs + SC_PROBES_LO, 8);
    h->timing.numeric = ev_ms(h, EV_NUM0, EV_NUM1);
    h->timing.total = ev_ms(h, EV_START, EV_HANDOFF) + h->timing.numeric;
    return MHB_OK;
}

// The fused call: C = A*B into caller-owned C arrays of `capacity` entries.  A caller that
// multiplies the same shapes repeatedly (the loop of src/main.cu:118-125, an AMG setup, a
// time-stepping code) keeps its C buffers, so the hand-off of src/main.cu:55-60 -- read nnz(C),
// allocate, continue -- has nothing left to allocate, and with it goes the last reason for the host
// to look at the device in the middle of the pipeline.  In steady state (same shape as the previous
// call on the handle) every kernel of both phases is launched from the previous call's bin sizes,
// the kernels take their row ranges from the device-side offsets, a one-thread kernel checks the
// guess on the device (k_fused_gate) and the numeric kernels stand down if it fails; the host
// synchronises ONCE, at the end, and re-runs the ordinary two-read path on a miss.
template <typename T>
int do_spgemm_into(mhb_context *h, int M, int K, int N, int nnzA, const int *Ap, const int *Ac, const T *Av, int nnzB,
                   const int *Bp, const int *Bc, const T *Bv, int *Cp, int *Cc, T *Cv, long long capacity,
                   long long *nnzC)
{
    if (!nnzC)
        return fail(h, MHB_ERR_ARG, "null nnzC");
    *nnzC = 0;
    const bool outstanding = h->into.active && h->into.pending;
    int rc = into_begin<T>(h, M, K, N, nnzA, Ap, Ac, Av, nnzB, Bp, Bc, Bv, Cp, Cc, Cv, capacity);
    if (rc)
    {
        if (!outstanding)
            h->into.active = false;
        return rc;
    }
    return into_end<T>(h, nnzC);
}

template <typename T>
int do_spgemm(mhb_context *h, int M, int K, int N, int nnzA, const int *Ap, const int *Ac, const T *Av, int nnzB,
              const int *Bp, const int *Bc, const T *Bv, int **Cp, int **Cc, T **Cv, long long *nnzC)
{
    if (!Cp || !Cc || !Cv || !nnzC)
        return fail(h, MHB_ERR_ARG, "null output pointer");
    *Cp = nullptr, *Cc = nullptr, *Cv = nullptr;
    CU(cudaMalloc((void **)Cp, sizeof(int) * ((size_t)M + 1)));
    int rc = do_symbolic(h, M, K, N, nnzA, Ap, Ac, nnzB, Bp, Bc, *Cp, nnzC);
    if (rc == MHB_OK)
    {
        size_t n = (size_t)std::max<long long>(*nnzC, 1);
        cudaError_t e1 = cudaMalloc((void **)Cc, sizeof(int) * n);
        cudaError_t e2 = cudaMalloc((void **)Cv, sizeof(T) * n);
        if (e1 != cudaSuccess || e2 != cudaSuccess)
            rc = fail(h, MHB_ERR_NOMEM, "cudaMalloc of C.col / C.val failed");
        else
            rc = do_numeric<T>(h, Av, Bv, *Cc, *Cv, true);
    }
    if (rc != MHB_OK)
    {
        cudaFree(*Cp), cudaFree(*Cc), cudaFree(*Cv);
        *Cp = nullptr, *Cc = nullptr, *Cv = nullptr;
    }
    return rc;
}

template <typename T>
int do_spgemm_host(mhb_context *h, int M, int K, int N, const int *hAp, const int *hAc, const T *hAv,
                   const int *hBp, const int *hBc, const T *hBv, const int **hCp, const int **hCc,
                   const T **hCv, long long *nnzC)
{
    if (!hAp || !hBp || !hCp || !hCc || !hCv || !nnzC)
        return fail(h, MHB_ERR_ARG, "null pointer");
    if (M < 0 || K < 0 || N < 0)
        return fail(h, MHB_ERR_ARG, "negative dimension");
    CU(cudaSetDevice(h->device));
    const int nnzA = hAp[M], nnzB = hBp[K];
    if (nnzA < 0 || nnzB < 0 || (nnzA > 0 && (!hAc || !hAv)) || (nnzB > 0 && (!hBc || !hBv)))
        return fail(h, MHB_ERR_ARG, "bad host CSR: negative nnz or null col / val array");
    const bool alias = (hAp == hBp && hAc == hBc && (const void *)hAv == (const void *)hBv && M == K);
    cudaStream_t st = h->stream;
    CU(h->sA_ptr.ensure(((size_t)M + 1) * 4));
    CU(h->sA_col.ensure(((size_t)nnzA + 1) * 4));
    CU(h->sA_val.ensure(((size_t)nnzA + 1) * sizeof(T)));
    CU(h->sC_ptr.ensure(((size_t)M + 2) * 4));
    CU(h->hC_ptr.ensure(((size_t)M + 2) * 4));
    // index arrays first: the symbolic phase needs no values, so their upload overlaps it
    CU(cudaMemcpyAsync(h->sA_ptr.p, hAp, ((size_t)M + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->sA_col.p, hAc, (size_t)nnzA * 4, cudaMemcpyHostToDevice, st));
    const int *dBp = h->sA_ptr.as<int>(), *dBc = h->sA_col.as<int>();
    const T *dBv = h->sA_val.as<T>();
    if (!alias)
    {
        CU(h->sB_ptr.ensure(((size_t)K + 1) * 4));
        CU(h->sB_col.ensure(((size_t)nnzB + 1) * 4));
        CU(h->sB_val.ensure(((size_t)nnzB + 1) * sizeof(T)));
        CU(cudaMemcpyAsync(h->sB_ptr.p, hBp, ((size_t)K + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->sB_col.p, hBc, (size_t)nnzB * 4, cudaMemcpyHostToDevice, st));
        dBp = h->sB_ptr.as<int>(), dBc = h->sB_col.as<int>(), dBv = h->sB_val.as<T>();
    }
    // values go up on a second stream: the symbolic phase reads no values, so this copy
    // overlaps the mask build / binning / symbolic kernels and numeric waits on ev_vals
    CU(cudaEventRecord(h->ev_ready, st)); // staging buffers are free once earlier work on st is done
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
    CU(cudaMemcpyAsync(h->sA_val.p, hAv, (size_t)nnzA * sizeof(T), cudaMemcpyHostToDevice, h->copy_stream));
    if (!alias)
        CU(cudaMemcpyAsync(h->sB_val.p, hBv, (size_t)nnzB * sizeof(T), cudaMemcpyHostToDevice, h->copy_stream));
    CU(cudaEventRecord(h->ev_vals, h->copy_stream));
    // every error exit below drains copy_stream first: the value upload reads the CALLER's
    // host arrays, which the caller may reuse the moment this function returns
    auto drain = [&](int code) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(st);
        return code;
    };
#define CUD(call)                                                                                        \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return drain(fail(h, e__ == cudaErrorMemoryAllocation ? MHB_ERR_NOMEM : MHB_ERR_CUDA,        \
                              std::string(#call) + ": " + cudaGetErrorString(e__)));                     \
    } while (0)
    int rc = do_symbolic(h, M, K, N, nnzA, h->sA_ptr.as<int>(), h->sA_col.as<int>(), nnzB, dBp, dBc,
                         h->sC_ptr.as<int>(), nnzC);
    if (rc)
        return drain(rc);
    size_t n = (size_t)std::max<long long>(*nnzC, 1);
    CUD(h->sC_col.ensure(n * 4));
    CUD(h->sC_val.ensure(n * sizeof(T)));
    CUD(h->hC_col.ensure(n * 4));
    CUD(h->hC_val.ensure(n * sizeof(T)));
    CUD(cudaMemcpyAsync(h->hC_ptr.p, h->sC_ptr.p, ((size_t)M + 1) * 4, cudaMemcpyDeviceToHost, st));
    // The download of C is most of this call (PCIe: 12 bytes per entry).  With enough of it, the
    // numeric phase runs in row chunks balanced by nnz and the copy of a finished chunk (on
    // copy_stream) overlaps the computation of the next: the tail of the call is then the download
    // alone instead of numeric + download.
    const int nc = (h->row_chunks > 1 && *nnzC * (long long)(4 + sizeof(T)) >= h->row_chunk_bytes && M >= h->row_chunks)
                       ? std::min(h->row_chunks, (int)kMaxRowChunks)
                       : 1;
    if (nc > 1)
    {
        CUD(cudaStreamSynchronize(st)); // C.ptr is on the host: chunk bounds and byte ranges come from it
        const int *hp = h->hC_ptr.as<int>();
        ChunkRows cr;
        cr.r[0] = 0;
        for (int c = 1; c < nc; ++c)
        {
            // geometric sizes (1 : 2 : 4 ...): a small first chunk starts the download early, and few,
            // large copies keep PCIe efficient (r2t: 8 copies of 17-34 MB cost 0.28 ms more than 2 of 67-135 MB,
            // as much as four equal chunks gained)
            const long long want = *nnzC * ((1LL << c) - 1) / ((1LL << nc) - 1);
            int r = (int)(std::lower_bound(hp, hp + M + 1, (int)std::min<long long>(want, INT_MAX)) - hp);
            cr.r[c] = std::max(cr.r[c - 1], std::min(r, M));
        }
        for (int c = nc; c <= (int)kMaxRowChunks; ++c)
            cr.r[c] = M;
        CUD(h->chunk_off.ensure((size_t)MHB_MAX_BINS * (kMaxRowChunks + 1) * 4));
        LAUNCH(h, k_chunk_bounds, cdiv((long long)NB_COUNT * (nc + 1), 128), 128, 0, (const int *)h->bins_num.as<int>(),
               (const int *)(h->scal.as<int>() + SC_NUM_OFF), (int)NB_COUNT, nc, cr, h->chunk_off.as<int>(), h->asame,
               (int)NB_WIN_COMPACT, h->row_twins ? (int)NB_WIN_WARP : -1);
        CUD(cudaStreamWaitEvent(st, h->ev_vals, 0));
        CUD(cudaEventRecord(h->ev[EV_NUM0], st));
        // MHB_TRACE_HOST=1: device timeline of the chunks (numeric done / download done), printed after the call
        static const bool trace = std::getenv("MHB_TRACE_HOST") != nullptr;
        cudaEvent_t tn[kMaxRowChunks] = {nullptr}, td[kMaxRowChunks] = {nullptr}, tv = nullptr;
        if (trace)
        {
            for (int c = 0; c < nc; ++c)
                cudaEventCreate(&tn[c]), cudaEventCreate(&td[c]);
            cudaEventCreate(&tv);
            cudaEventRecord(tv, st); // = value upload complete (st has just waited for it)
        }
        for (int c = 0; c < nc; ++c)
        {
            rc = launch_numeric_bins<T>(h, h->sA_val.as<T>(), dBv, h->sC_col.as<int>(), h->sC_val.as<T>(), false,
                                        h->chunk_off.as<int>(), c, nc);
            if (rc)
                return drain(rc);
            CUD(cudaEventRecord(h->ev_chunk[c], st));
            if (trace)
                cudaEventRecord(tn[c], st);
            CUD(cudaStreamWaitEvent(h->copy_stream, h->ev_chunk[c], 0));
            const size_t e0 = (size_t)hp[cr.r[c]], e1 = (size_t)hp[cr.r[c + 1]];
            if (e1 > e0)
            {
                CUD(cudaMemcpyAsync(h->hC_col.as<int>() + e0, h->sC_col.as<int>() + e0, (e1 - e0) * 4, cudaMemcpyDeviceToHost,
                                    h->copy_stream));
                CUD(cudaMemcpyAsync(h->hC_val.as<T>() + e0, h->sC_val.as<T>() + e0, (e1 - e0) * sizeof(T),
                                    cudaMemcpyDeviceToHost, h->copy_stream));
            }
            if (trace)
                cudaEventRecord(td[c], h->copy_stream);
        }
        CUD(cudaEventRecord(h->ev[EV_NUM1], st));
        if (trace)
        {
            cudaStreamSynchronize(st);
            cudaStreamSynchronize(h->copy_stream);
            float a = 0.f;
            cudaEventElapsedTime(&a, h->ev[EV_START], tv);
            std::fprintf(stderr, "[mhb host trace] values up at %.3f ms after the first kernel;", a);
            for (int c = 0; c < nc; ++c)
            {
                float n_ = 0.f, d_ = 0.f;
                cudaEventElapsedTime(&n_, h->ev[EV_START], tn[c]);
                cudaEventElapsedTime(&d_, h->ev[EV_START], td[c]);
                std::fprintf(stderr, " chunk %d: numeric %.3f download %.3f (%.1f MB);", c, n_, d_,
                             (hp[cr.r[c + 1]] - hp[cr.r[c]]) * (4.0 + sizeof(T)) / 1e6);
                cudaEventDestroy(tn[c]), cudaEventDestroy(td[c]);
            }
            cudaEventDestroy(tv);
            std::fprintf(stderr, "\n");
        }
        h->stats.gpu_launches = h->launches;
    }
    else
    {
        CUD(cudaStreamWaitEvent(st, h->ev_vals, 0));
        rc = do_numeric<T>(h, h->sA_val.as<T>(), dBv, h->sC_col.as<int>(), h->sC_val.as<T>(), false);
        if (rc)
            return drain(rc);
        CUD(cudaMemcpyAsync(h->hC_col.p, h->sC_col.p, (size_t)*nnzC * 4, cudaMemcpyDeviceToHost, st));
        CUD(cudaMemcpyAsync(h->hC_val.p, h->sC_val.p, (size_t)*nnzC * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    int *hs = h->h_scal.as<int>();
    CUD(cudaMemcpyAsync(hs + SC_ERROR, h->scal.as<int>() + SC_ERROR, 6 * 4, cudaMemcpyDeviceToHost, st));
    CUD(cudaStreamSynchronize(st));
    if (nc > 1)
        CUD(cudaStreamSynchronize(h->copy_stream));
#undef CUD
    rc = check_dev_error(h, hs);
    if (rc)
        return rc;
    std::memcpy(&h->stats.hash_probes, hs + SC_PROBES_LO, 8);
    h->timing.numeric = ev_ms(h, EV_NUM0, EV_NUM1);
    h->timing.total = ev_ms(h, EV_START, EV_HANDOFF) + h->timing.numeric;
    *hCp = h->hC_ptr.as<int>();
    *hCc = h->hC_col.as<int>();
    *hCv = h->hC_val.as<T>();
    return MHB_OK;
}

// ---- device-side CSR transpose (the AAT mode's B operand; src/utils.cpp:20-46) -------------
template <typename T>
int do_transpose(mhb_context *h, int M, int N, int nnz, const int *Ap, const int *Ac, const T *Av, int *Tp, int *Tc,
                 T *Tv)
{
    if (M < 0 || N < 0 || nnz < 0)
        return fail(h, MHB_ERR_ARG, "negative dimension");
    if (!Ap || !Tp || (nnz > 0 && (!Ac || !Av || !Tc || !Tv)))
        return fail(h, MHB_ERR_ARG, "null CSR pointer");
    CU(cudaSetDevice(h->device));
    h->launches = 0;
    h->have_pattern = false; // scan_tmp / scal are shared with the SpGEMM phases
    const int nblocks = cdiv(std::max(nnz, 1), kRadixTile);
    const long long nhist = (long long)kRadixBins * nblocks;
    CU(h->scan_tmp.ensure((size_t)(cdiv(std::max<long long>(std::max<long long>(nhist, (long long)N + 1), 1), kScanTile) + 2) * 8));
    CU(h->scal.ensure((size_t)SC_COUNT * 4));
    CU(cudaMemsetAsync(Tp, 0, ((size_t)N + 1) * 4, h->stream));
    if (nnz == 0)
    {
        CU(cudaStreamSynchronize(h->stream));
        return MHB_OK;
    }
    // T.ptr: column histogram of A, scanned in place
    LAUNCH(h, k_tr_count_cols, std::min(cdiv(nnz, 256), h->num_sms * 16), 256, 0, Ac, (long long)nnz, Tp);
    int rc = run_scan(h, LoadInt{Tp}, N, Tp, 1, nullptr);
    if (rc)
        return rc;
    // stable LSD radix sort of the nonzeros by column, 8 bits per pass
    int bits = 1;
    while (bits < 31 && (1LL << bits) < (long long)N)
        ++bits;
    const int passes = (bits + kRadixBits - 1) / kRadixBits;
    CU(h->tr_hist.ensure((size_t)nhist * 4));
    if (passes > 1)
        for (int b = 0; b < 2; ++b)
        {
            CU(h->tr_key[b].ensure((size_t)nnz * 4));
            CU(h->tr_idx[b].ensure((size_t)nnz * 4));
        }
    const int *kin = Ac, *iin = nullptr;
    int *hist = h->tr_hist.as<int>();
    for (int p = 0; p < passes; ++p)
    {
        const int shift = p * kRadixBits;
        LAUNCH(h, k_radix_count, nblocks, kRadixThreads, 0, kin, (long long)nnz, shift, hist, nblocks);
        rc = run_scan(h, LoadInt{hist}, nhist, hist, 0, nullptr);
        if (rc)
            return rc;
        if (p == passes - 1)
        {
            auto kern = k_radix_scatter<T, true>;
            LAUNCH(h, kern, nblocks, kRadixThreads, 0, kin, iin, (long long)nnz, shift, (const int *)hist, nblocks,
                   (int *)nullptr, (int *)nullptr, Ap, M, Av, Tc, Tv);
        }
        else
        {
            int *kout = h->tr_key[p & 1].as<int>(), *iout = h->tr_idx[p & 1].as<int>();
            auto kern = k_radix_scatter<T, false>;
            LAUNCH(h, kern, nblocks, kRadixThreads, 0, kin, iin, (long long)nnz, shift, (const int *)hist, nblocks, kout,
                   iout, Ap, M, Av, (int *)nullptr, (T *)nullptr);
            kin = kout, iin = iout;
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    h->stats.gpu_launches = h->launches;
    return MHB_OK;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C"
{

    const char *mhb_version(void) { return "mhb-spgemm 0.2 (sm_100a)"; }

    int mhb_create(mhb_handle_t *out, int device)
    {
        if (!out)
            return MHB_ERR_ARG;
        *out = nullptr;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
            return MHB_ERR_CUDA; // no CPU fallback: without a CUDA device the library refuses to run
        mhb_context *h = new mhb_context();
        h->device = device;
        if (const char *e = std::getenv("MHB_PDL")) // A/B switch for measurements
            h->pdl = std::atoi(e);
        auto bail = [&](int code) {
            delete h;
            return code;
        };
        if (cudaSetDevice(device) != cudaSuccess)
            return bail(MHB_ERR_CUDA);
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
            return bail(MHB_ERR_CUDA);
        h->num_sms = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess)
            return bail(MHB_ERR_CUDA);
        h->stream = h->own_stream;
        if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_vals, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming) != cudaSuccess)
            return bail(MHB_ERR_CUDA);
        for (auto &e : h->ev)
            if (cudaEventCreate(&e) != cudaSuccess)
                return bail(MHB_ERR_CUDA);
        for (auto &e : h->ev_chunk)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
                return bail(MHB_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess)
            return bail(MHB_ERR_CUDA);
        int prio_lo = 0, prio_hi = 0; // helper streams 0-1 (big-row bins) get the high priority
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        for (int a = 0; a < mhb_context::kAux; ++a)
            if (cudaStreamCreateWithPriority(&h->aux[a], cudaStreamNonBlocking, a < 2 ? prio_hi : prio_lo) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->ev_join[a], cudaEventDisableTiming) != cudaSuccess)
                return bail(MHB_ERR_CUDA);
        if (set_kernel_attributes(h) != MHB_OK)
        {
            std::fprintf(stderr, "mhb_create: %s\n", h->err.c_str());
            return bail(MHB_ERR_CUDA);
        }
        *out = h;
        return MHB_OK;
    }

    int mhb_destroy(mhb_handle_t h)
    {
        if (!h)
            return MHB_OK;
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        for (DevBuf *b : {&h->flags, &h->wordprefix, &h->tileptr, &h->tilecol, &h->tilemask, &h->binfo, &h->arow,
                          &h->binid, &h->bsame, &h->asame_buf, &h->bm_store, &h->bm_slot, &h->bins_sym, &h->bins_num, &h->blockhist, &h->scan_tmp, &h->scal, &h->pool,
                          &h->sA_ptr, &h->sA_col, &h->sA_val, &h->sB_ptr, &h->sB_col, &h->sB_val, &h->sC_ptr,
                          &h->sC_col, &h->sC_val, &h->tr_key[0], &h->tr_key[1], &h->tr_idx[0], &h->tr_idx[1], &h->tr_hist, &h->chunk_off})
            b->release();
        for (HostBuf *b : {&h->h_scal, &h->hC_ptr, &h->hC_col, &h->hC_val})
            b->release();
        for (auto &e : h->ev)
            if (e)
                cudaEventDestroy(e);
        for (auto &e : h->ev_chunk)
            if (e)
                cudaEventDestroy(e);
        for (int a = 0; a < mhb_context::kAux; ++a)
        {
            if (h->aux[a])
                cudaStreamDestroy(h->aux[a]);
            if (h->ev_join[a])
                cudaEventDestroy(h->ev_join[a]);
        }
        if (h->ev_fork)
            cudaEventDestroy(h->ev_fork);
        if (h->ev_vals)
            cudaEventDestroy(h->ev_vals);
        if (h->ev_ready)
            cudaEventDestroy(h->ev_ready);
        if (h->copy_stream)
            cudaStreamDestroy(h->copy_stream);
        if (h->own_stream)
            cudaStreamDestroy(h->own_stream);
        delete h;
        return MHB_OK;
    }

    const char *mhb_last_error(mhb_handle_t h) { return h ? h->err.c_str() : "null handle"; }

    int mhb_set_stream(mhb_handle_t h, void *cuda_stream)
    {
        if (!h)
            return MHB_ERR_ARG;
        h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
        return MHB_OK;
    }

    int mhb_get_stream(mhb_handle_t h, void **cuda_stream)
    {
        if (!h || !cuda_stream)
            return MHB_ERR_ARG;
        *cuda_stream = (void *)h->stream;
        return MHB_OK;
    }

    int mhb_set_option(mhb_handle_t h, const char *key, long long value)
    {
        if (!h || !key)
            return MHB_ERR_ARG;
        std::string k(key);
        if (k == "force_sym_path")
            h->force_sym = (int)value;
        else if (k == "force_num_path")
            h->force_num = (int)value;
        else if (k == "claim_list")
            h->claim_list = (int)value;
        else if (k == "compact_rows")
            h->compact_rows = (int)value;
        else if (k == "row_twins")
            h->row_twins = (int)value;
        else if (k == "sym_twins")
            h->sym_twins = (int)value;
        else if (k == "pdl")
            h->pdl = (int)value;
        else if (k == "count_probes")
            h->count_probes = (int)value;
        else if (k == "mask_onepass")
            h->mask_onepass = (int)value;
        else if (k == "speculate")
            h->speculate = (int)value;
        else if (k == "row_chunk_bytes")
            h->row_chunk_bytes = std::max<long long>(1, value);
        else if (k == "row_chunks")
            h->row_chunks = (int)std::max<long long>(1, std::min<long long>(value, kMaxRowChunks));
        else if (k == "nnz_limit")
            h->nnz_limit = std::min<long long>(value > 0 ? value : INT_MAX, INT_MAX);
        else if (k == "serial_bins")
            h->serial = value != 0;
        else if (k == "verbose")
            h->verbose = (int)value;
        else
            return fail(h, MHB_ERR_ARG, "unknown option " + k);
        h->spec_ok = false; // bin sizes remembered from the previous call were formed under the old options
        return MHB_OK;
    }

    int mhb_symbolic(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col, int nnzB,
                     const int *dB_ptr, const int *dB_col, int *dC_ptr, long long *nnzC)
    {
        if (!h || !nnzC)
            return MHB_ERR_ARG;
        return do_symbolic(h, M, K, N, nnzA, dA_ptr, dA_col, nnzB, dB_ptr, dB_col, dC_ptr, nnzC);
    }

    int mhb_numeric_f64(mhb_handle_t h, const double *dA_val, const double *dB_val, int *dC_col, double *dC_val)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_numeric<double>(h, dA_val, dB_val, dC_col, dC_val, true);
    }
    int mhb_numeric_f32(mhb_handle_t h, const float *dA_val, const float *dB_val, int *dC_col, float *dC_val)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_numeric<float>(h, dA_val, dB_val, dC_col, dC_val, true);
    }

    int mhb_spgemm_f64(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                       const double *dA_val, int nnzB, const int *dB_ptr, const int *dB_col, const double *dB_val,
                       int **dC_ptr, int **dC_col, double **dC_val, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm<double>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr,
                                 dC_col, dC_val, nnzC);
    }
    int mhb_spgemm_f32(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                       const float *dA_val, int nnzB, const int *dB_ptr, const int *dB_col, const float *dB_val,
                       int **dC_ptr, int **dC_col, float **dC_val, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm<float>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr,
                                dC_col, dC_val, nnzC);
    }
    int mhb_spgemm_into_f64(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                            const double *dA_val, int nnzB, const int *dB_ptr, const int *dB_col, const double *dB_val,
                            int *dC_ptr, int *dC_col, double *dC_val, long long capacity, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm_into<double>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr,
                                      dC_col, dC_val, capacity, nnzC);
    }
    int mhb_spgemm_into_f32(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                            const float *dA_val, int nnzB, const int *dB_ptr, const int *dB_col, const float *dB_val,
                            int *dC_ptr, int *dC_col, float *dC_val, long long capacity, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm_into<float>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr,
                                     dC_col, dC_val, capacity, nnzC);
    }
    int mhb_spgemm_into_begin_f64(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                                  const double *dA_val, int nnzB, const int *dB_ptr, const int *dB_col,
                                  const double *dB_val, int *dC_ptr, int *dC_col, double *dC_val, long long capacity)
    {
        if (!h)
            return MHB_ERR_ARG;
        return into_begin<double>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr, dC_col,
                                  dC_val, capacity);
    }
    int mhb_spgemm_into_begin_f32(mhb_handle_t h, int M, int K, int N, int nnzA, const int *dA_ptr, const int *dA_col,
                                  const float *dA_val, int nnzB, const int *dB_ptr, const int *dB_col,
                                  const float *dB_val, int *dC_ptr, int *dC_col, float *dC_val, long long capacity)
    {
        if (!h)
            return MHB_ERR_ARG;
        return into_begin<float>(h, M, K, N, nnzA, dA_ptr, dA_col, dA_val, nnzB, dB_ptr, dB_col, dB_val, dC_ptr, dC_col,
                                 dC_val, capacity);
    }
    int mhb_spgemm_into_end(mhb_handle_t h, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return h->into.vbytes == 4 ? into_end<float>(h, nnzC) : into_end<double>(h, nnzC);
    }
    int mhb_device_free(void *dptr) { return cudaFree(dptr) == cudaSuccess ? MHB_OK : MHB_ERR_CUDA; }
    int mhb_device_alloc(void **dptr, size_t bytes)
    {
        if (!dptr)
            return MHB_ERR_ARG;
        return cudaMalloc(dptr, bytes ? bytes : 1) == cudaSuccess ? MHB_OK : MHB_ERR_NOMEM;
    }
    int mhb_memcpy_h2d(void *dptr, const void *hptr, size_t bytes)
    {
        return cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? MHB_OK : MHB_ERR_CUDA;
    }
    int mhb_memcpy_d2h(void *hptr, const void *dptr, size_t bytes)
    {
        return cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? MHB_OK : MHB_ERR_CUDA;
    }

    int mhb_spgemm_host_f64(mhb_handle_t h, int M, int K, int N, const int *hA_ptr, const int *hA_col,
                            const double *hA_val, const int *hB_ptr, const int *hB_col, const double *hB_val,
                            const int **hC_ptr, const int **hC_col, const double **hC_val, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm_host<double>(h, M, K, N, hA_ptr, hA_col, hA_val, hB_ptr, hB_col, hB_val, hC_ptr, hC_col,
                                      hC_val, nnzC);
    }
    int mhb_spgemm_host_f32(mhb_handle_t h, int M, int K, int N, const int *hA_ptr, const int *hA_col,
                            const float *hA_val, const int *hB_ptr, const int *hB_col, const float *hB_val,
                            const int **hC_ptr, const int **hC_col, const float **hC_val, long long *nnzC)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_spgemm_host<float>(h, M, K, N, hA_ptr, hA_col, hA_val, hB_ptr, hB_col, hB_val, hC_ptr, hC_col,
                                     hC_val, nnzC);
    }

    int mhb_transpose_f64(mhb_handle_t h, int M, int N, int nnz, const int *dA_ptr, const int *dA_col,
                          const double *dA_val, int *dT_ptr, int *dT_col, double *dT_val)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_transpose<double>(h, M, N, nnz, dA_ptr, dA_col, dA_val, dT_ptr, dT_col, dT_val);
    }
    int mhb_transpose_f32(mhb_handle_t h, int M, int N, int nnz, const int *dA_ptr, const int *dA_col,
                          const float *dA_val, int *dT_ptr, int *dT_col, float *dT_val)
    {
        if (!h)
            return MHB_ERR_ARG;
        return do_transpose<float>(h, M, N, nnz, dA_ptr, dA_col, dA_val, dT_ptr, dT_col, dT_val);
    }

    int mhb_host_alloc(void **hptr, size_t bytes)
    {
        if (!hptr)
            return MHB_ERR_ARG;
        return cudaMallocHost(hptr, bytes ? bytes : 1) == cudaSuccess ? MHB_OK : MHB_ERR_NOMEM;
    }
    int mhb_host_free(void *hptr) { return cudaFreeHost(hptr) == cudaSuccess ? MHB_OK : MHB_ERR_CUDA; }

    int mhb_form_mask_matrix_B(mhb_handle_t h, int K, int N, int nnzB, const int *dB_ptr, const int *dB_col,
                               const int **d_tileptr, const int **d_tilecol, const unsigned **d_tilemask,
                               long long *ntiles)
    {
        if (!h || !d_tileptr || !d_tilecol || !d_tilemask || !ntiles)
            return MHB_ERR_ARG;
        if (K < 0 || N < 0 || nnzB < 0 || !dB_ptr || (nnzB > 0 && !dB_col))
            return fail(h, MHB_ERR_ARG, "bad B");
        CU(cudaSetDevice(h->device));
        h->have_pattern = false;
        h->launches = 0;
        bool grew;
        int rc = ensure_workspace(h, 0, K, nnzB, &grew);
        if (rc)
            return rc;
        rc = zero_scal(h, h->mask_onepass ? scal_bytes(nnzB) : (size_t)SC_COUNT * 4);
        if (rc)
            return rc;
        rc = build_mask_matrix(h, K, nnzB, dB_ptr, dB_col);
        if (rc)
            return rc;
        int *hs = h->h_scal.as<int>();
        CU(cudaMemcpyAsync(hs, h->scal.p, SC_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        std::memcpy(ntiles, hs + SC_NTILES_LO, 8);
        *d_tileptr = h->tileptr.as<int>();
        *d_tilecol = h->tilecol.as<int>();
        *d_tilemask = h->tilemask.as<unsigned>();
        h->stats.gpu_launches = h->launches;
        return MHB_OK;
    }

    int mhb_get_row_info(mhb_handle_t h, const int **d_row_info)
    {
        if (!h || !d_row_info)
            return MHB_ERR_ARG;
        if (!h->have_pattern)
            return fail(h, MHB_ERR_ARG, "no symbolic result on this handle");
        *d_row_info = h->arow.as<int>();
        return MHB_OK;
    }

    int mhb_get_bins(mhb_handle_t h, int which, int *nbins, const int **d_bins, int *h_bin_offset)
    {
        if (!h || !nbins || !d_bins || !h_bin_offset)
            return MHB_ERR_ARG;
        if (!h->have_pattern)
            return fail(h, MHB_ERR_ARG, "no symbolic result on this handle");
        *nbins = which == 0 ? (int)SB_COUNT : (int)NB_COUNT;
        *d_bins = which == 0 ? h->bins_sym.as<int>() : h->bins_num.as<int>();
        std::memcpy(h_bin_offset, which == 0 ? h->sym_off : h->num_off, sizeof(int) * (MHB_MAX_BINS + 1));
        return MHB_OK;
    }

    int mhb_get_device_scalars(mhb_handle_t h, const long long **d_nnzC, const int **d_gate)
    {
        if (!h || !d_nnzC || !d_gate)
            return MHB_ERR_ARG;
        if (!h->scal.p)
            return fail(h, MHB_ERR_ARG, "no call on this handle yet");
        *d_nnzC = reinterpret_cast<const long long *>(h->scal.as<int>() + SC_NNZC_LO);
        *d_gate = h->scal.as<int>() + SC_GATE;
        return MHB_OK;
    }

    int mhb_get_timing(mhb_handle_t h, mhb_timing *out)
    {
        if (!h || !out)
            return MHB_ERR_ARG;
        *out = h->timing;
        return MHB_OK;
    }
    int mhb_get_stats(mhb_handle_t h, mhb_stats *out)
    {
        if (!h || !out)
            return MHB_ERR_ARG;
        *out = h->stats;
        return MHB_OK;
    }

} // extern "C"
