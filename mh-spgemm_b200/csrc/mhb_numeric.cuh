// mhb_numeric.cuh -- kernel family 4: the numeric accumulate / compact / sort pass.
//
// Replaces h_numeric (inc/MH_spgemm.cuh:364-430) and its kernels (inc/numeric.cuh):
// column indices (ascending) and values of every row of C = A*B, given C's row offsets.
//
// The reference does, per intermediate product, a 64-bit modulo hash, an ATOMS.CAS on the
// key and an LDS.64 + DADD + ATOMS.CAST.SPIN.64 loop on the value (shared-memory fp64 add
// has no native atomic on sm_100), then compacts through an atomic cursor and rank-sorts
// in O(n^2).  Two accumulators replace that here:
//
//  * DENSE WINDOW: a C row whose column span W fits shared memory (227 KB holds 27 648
//    fp64 columns) accumulates into acc[col - first_col].  No key, no hash, no probing,
//    and the compaction sweep emits the row already column-sorted -- no sort at all.
//  * HASH (multiplicative hash, power-of-two table, linear probing) for rows whose span is
//    large relative to their nnz; compaction by ballot/prefix and a bitonic (or, for tiny
//    rows, rank) sort; a global-memory table for rows above 12 288 nnz.
//
// Group kernels (8 lanes or one warp per row) walk A's row one B row at a time.  The
// columns of one B row are distinct, so in one step no two lanes update the same
// accumulator: the value update is a plain LDS / FMA / STS ordered by __syncwarp, with no
// shared-memory atomic at all.  Block kernels (one block per row, warps on different B
// rows) fall back to atomicAdd.  Loads of B are coalesced along the B row.
#pragma once
#include "mhb_common.cuh"
#include "mhb_stream.cuh"

namespace mhb
{

constexpr int kNumGroupThreads = 256;
constexpr int kTinyRowThreads = 256;
constexpr int kPre = 3;      // chunks of the next B row prefetched by the dense-window kernel

// =========================================================================================
// Dense window, G lanes per row.
// =========================================================================================
template <int G, typename T>
__global__ void __launch_bounds__(kNumGroupThreads, 3) // 3 blocks/SM is what the 64 KB windows allow
    k_num_win_group(RowList list, const int *__restrict__ Ap,
                    const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                    const int *__restrict__ Bc, const T *__restrict__ Bv, const int4 *__restrict__ arow,
                    const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int wcap,
                    const unsigned char *__restrict__ same)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int GPB = kNumGroupThreads / G;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    T *acc = reinterpret_cast<T *>(sm_raw) + (size_t)g * wcap;
    // per-group staging of the current chunk of A's row (G + 2 int4), after all the windows
    int4 *stage = reinterpret_cast<int4 *>(reinterpret_cast<T *>(sm_raw) + (size_t)GPB * wcap) + (size_t)g * (G + 2);
    for (int r = blockIdx.x * GPB + g; r < nrows; r += gridDim.x * GPB)
    {
        const int row = rows[r];
        const int4 info = arow[row];
        const int cmin = info.z;
        const int W = info.w - cmin + 1;
        for (int i = l; i < W; i += G)
            acc[i] = Unset<T>::value();
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // the columns of one B row are distinct: plain read-modify-write, no atomics;
        // twin B rows (same pattern, adjacent in A's row) are folded into one update
        walk_sequential_twins<G, kPre, T>(gm, l, s, e, Ac, Av, Bp, Bc, Bv, same, stage, [&](int c, T v, bool active) {
            const int idx = active ? c - cmin : 0; // idle lanes read slot 0 (a broadcast) and store nothing
            const T o = acc[idx];
            const T nv = Unset<T>::is(o) ? v : o + v;
            if (active)
                acc[idx] = nv;
        });
        // ordered compaction: the window is already sorted by column
        int out = __ldg(&Cp[row]);
        for (int i0 = 0; i0 < W; i0 += G)
        {
            const int i = i0 + l;
            const T v = (i < W) ? acc[i] : Unset<T>::value();
            const bool p = !Unset<T>::is(v);
            const unsigned bal = __ballot_sync(gm, p);
            if (p)
            {
                const int pos = out + __popc(bal & lanemask_lt());
                Cc[pos] = cmin + i;
                Cv[pos] = v;
            }
            out += __popc(bal);
        }
        __syncwarp(gm);
    }
}

// =========================================================================================
// Dense window, one warp per group of up to three TWIN ROWS OF A (rows of A with one column
// pattern, e.g. the dofs of one FEM node).  Twin rows of A select the same rows of B, so the
// B data (columns once, values of up to three twin B rows) is loaded once per step and
// feeds three accumulator windows: v_r = sum_j a[r][j] * b_j.  Per three C rows this
// issues ~1.9x fewer instructions and ~1.5x fewer LSU wavefronts than three passes of
// k_num_win_group (the kernel is bound by exactly those, profiles/r1e_numwin_final.md).
// Rows that are not twins take the same code with one window.
// smem per warp: 3 windows of wcap | stage_be[G+2] (int2) | stage_a[3][G+2]
// =========================================================================================
constexpr int kRowTwinThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kRowTwinThreads)
    k_num_win_rowtwins(RowList list, const int *__restrict__ Ap,
                       const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                       const int *__restrict__ Bc, const T *__restrict__ Bv, const int4 *__restrict__ arow,
                       const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int wcap,
                       const unsigned char *__restrict__ bsame, const unsigned char *__restrict__ asame)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int G = 32, WPB = kRowTwinThreads / 32, SG = G + 2;
    const int warp = threadIdx.x >> 5, l = lane_id();
    T *acc0 = reinterpret_cast<T *>(sm_raw) + (size_t)warp * 3 * wcap;
    T *stage_a = reinterpret_cast<T *>(sm_raw) + (size_t)WPB * 3 * wcap + (size_t)warp * 3 * SG;
    int2 *stage_be = reinterpret_cast<int2 *>(reinterpret_cast<T *>(sm_raw) + (size_t)WPB * 3 * (wcap + SG)) +
                     (size_t)warp * SG;
    for (int chunk = blockIdx.x * WPB + warp; chunk * 3 < nrows; chunk += gridDim.x * WPB)
    {
        const int base = chunk * 3, cnt3 = min(3, nrows - base);
        int j = 0;
        while (j < cnt3)
        {
            const int row = __ldg(&rows[base + j]);
            int R = 1;
            if (j + 1 < cnt3 && __ldg(&rows[base + j + 1]) == row + 1 && __ldg(&asame[row + 1]))
            {
                R = 2;
                if (j + 2 < cnt3 && __ldg(&rows[base + j + 2]) == row + 2 && __ldg(&asame[row + 2]))
                    R = 3;
            }
            j += R;
            const int4 info = __ldg(&arow[row]);
            const int cmin = info.z, W = info.w - cmin + 1;
            for (int r = 0; r < R; ++r)
                for (int i = l; i < W; i += G)
                    acc0[r * wcap + i] = Unset<T>::value();
            if (l < 2)
            {
                stage_be[G + l] = make_int2(0, 0);
                stage_a[0 * SG + G + l] = stage_a[1 * SG + G + l] = stage_a[2 * SG + G + l] = T(0);
            }
            __syncwarp();
            const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
            const int d1 = (R > 1) ? __ldg(&Ap[row + 1]) - s : 0, d2 = (R > 2) ? __ldg(&Ap[row + 2]) - s : 0;
            int bs, be, kk, nbs, nbe, nkk;
            T av0, av1 = T(0), av2 = T(0), nav0, nav1 = T(0), nav2 = T(0);
            auto meta = [&](int jj, int &ms, int &me, int &mk, T &m0, T &m1, T &m2) {
                ms = 0, me = 0, mk = -2, m0 = T(0), m1 = T(0), m2 = T(0);
                if (jj < e)
                {
                    mk = __ldg(&Ac[jj]);
                    m0 = __ldg(&Av[jj]);
                    if (R > 1)
                        m1 = __ldg(&Av[jj + d1]);
                    if (R > 2)
                        m2 = __ldg(&Av[jj + d2]);
                    ms = __ldg(&Bp[mk]);
                    me = __ldg(&Bp[mk + 1]);
                    if (__ldg(&bsame[mk]))
                        mk |= kTwinTag;
                }
            };
            meta(s + l, bs, be, kk, av0, av1, av2);
            for (int j0 = s; j0 < e; j0 += G)
            {
                meta(j0 + G + l, nbs, nbe, nkk, nav0, nav1, nav2);
                const int cnt = min(G, e - j0);
                stage_be[l] = make_int2(bs, be);
                stage_a[0 * SG + l] = av0;
                stage_a[1 * SG + l] = av1;
                stage_a[2 * SG + l] = av2;
                const int kprev = __shfl_up_sync(kFull, kk & kTwinMask, 1);
                const bool fol = l > 0 && l < cnt && (kk & kTwinTag) && (kk & kTwinMask) == kprev + 1;
                const unsigned fmask = __ballot_sync(kFull, fol);
                __syncwarp();
                int pc[kPre], nq = 0, nqe = 0, nsz = 1, nb1 = 0, nb2 = 0, ni = 0;
                T pv0[kPre], pv1[kPre], pv2[kPre];
                auto issue = [&](int i) {
                    ni = i;
                    nsz = 1 + ((fmask >> (i + 1)) & 1u);
                    if (nsz == 2)
                        nsz += (fmask >> (i + 2)) & 1u;
                    if (i + nsz > cnt)
                        nsz = cnt - i;
                    const int2 m0 = stage_be[i], m1 = stage_be[i + 1], m2 = stage_be[i + 2];
                    nq = m0.x;
                    nqe = m0.y;
                    nb1 = m1.x - nq;
                    nb2 = m2.x - nq;
#pragma unroll
                    for (int t = 0; t < kPre; ++t)
                    {
                        const int p = nq + t * G + l;
                        pc[t] = -1;
                        if (p < nqe)
                        {
                            pc[t] = __ldg(&Bc[p]);
                            pv0[t] = __ldg(&Bv[p]);
                            if (nsz > 1)
                                pv1[t] = __ldg(&Bv[p + nb1]);
                            if (nsz > 2)
                                pv2[t] = __ldg(&Bv[p + nb2]);
                        }
                    }
                };
                issue(0);
                for (int i = 0; i < cnt;)
                {
                    int cc[kPre];
                    T b0[kPre], b1[kPre], b2[kPre];
                    const int q = nq, qe = nqe, sz = nsz, o1 = nb1, o2 = nb2, ci = ni;
#pragma unroll
                    for (int t = 0; t < kPre; ++t)
                        cc[t] = pc[t], b0[t] = pv0[t], b1[t] = pv1[t], b2[t] = pv2[t];
                    i += sz;
                    if (i < cnt)
                        issue(i);
                    // a[r][jj]: value of twin row r of A at nonzero ci + jj (broadcast loads)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                    {
                        if (r >= R)
                            break;
                        const T a0 = stage_a[r * SG + ci];
                        const T a1 = (sz > 1) ? stage_a[r * SG + ci + 1] : T(0);
                        const T a2 = (sz > 2) ? stage_a[r * SG + ci + 2] : T(0);
                        T *acc = acc0 + r * wcap;
#pragma unroll
                        for (int t = 0; t < kPre; ++t)
                        {
                            T v = a0 * b0[t];
                            if (sz > 1)
                                v = fma(a1, b1[t], v);
                            if (sz > 2)
                                v = fma(a2, b2[t], v);
                            const bool active = cc[t] >= 0;
                            const int idx = active ? cc[t] - cmin : 0;
                            const T o = acc[idx];
                            const T nv = Unset<T>::is(o) ? v : o + v;
                            if (active)
                                acc[idx] = nv;
                        }
                        for (int p = q + kPre * G + l; p < qe; p += G) // B rows longer than kPre*G
                        {
                            T v = a0 * __ldg(&Bv[p]);
                            if (sz > 1)
                                v = fma(a1, __ldg(&Bv[p + o1]), v);
                            if (sz > 2)
                                v = fma(a2, __ldg(&Bv[p + o2]), v);
                            const int idx = __ldg(&Bc[p]) - cmin;
                            const T o = acc[idx];
                            acc[idx] = Unset<T>::is(o) ? v : o + v;
                        }
                    }
                    __syncwarp();
                }
                bs = nbs, be = nbe, kk = nkk, av0 = nav0, av1 = nav1, av2 = nav2;
            }
            // ordered compaction; twin rows share one structure, so one ballot serves all of them
            int out0 = __ldg(&Cp[row]);
            const int o1 = (R > 1) ? __ldg(&Cp[row + 1]) - out0 : 0, o2 = (R > 2) ? __ldg(&Cp[row + 2]) - out0 : 0;
            for (int i0 = 0; i0 < W; i0 += G)
            {
                const int i = i0 + l;
                const T v = (i < W) ? acc0[i] : Unset<T>::value();
                const bool p = !Unset<T>::is(v);
                const unsigned bal = __ballot_sync(kFull, p);
                if (p)
                {
                    const int pos = out0 + __popc(bal & lanemask_lt());
                    Cc[pos] = cmin + i;
                    Cv[pos] = v;
                    if (R > 1)
                    {
                        Cc[pos + o1] = cmin + i;
                        Cv[pos + o1] = acc0[wcap + i];
                    }
                    if (R > 2)
                    {
                        Cc[pos + o2] = cmin + i;
                        Cv[pos + o2] = acc0[2 * wcap + i];
                    }
                }
                out0 += __popc(bal);
            }
            __syncwarp();
        }
    }
}

// One step of the rank-mapped kernel: up to kPre*32 entries of a B row (and of its one or two
// folded twins) held in registers.
template <typename T>
struct CompactStep
{
    int c[kPre];                     // column per chunk, -1 = lane idle
    T v0[kPre], v1[kPre], v2[kPre];  // values of the B row and of its folded twins
    int q, qe, sz, o1, o2, ci;       // entry range, folded rows (1..3), twin offsets, index into the A stage
};

// acc_r[rank] += sum_j a[r][j] * b_j for R twin rows of A and SZ folded rows of B.  The ranks of
// one step are distinct (distinct columns of one B row), so all loads may precede all stores.
template <typename T, int R, int SZ>
__device__ __forceinline__ void compact_accumulate(T *acc0, int ncap, const T *sa, int SG, const int (&rk)[kPre],
                                                   const bool (&act)[kPre], const CompactStep<T> &P)
{
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
        const T a0 = sa[r * SG];
        const T a1 = (SZ > 1) ? sa[r * SG + 1] : T(0);
        const T a2 = (SZ > 2) ? sa[r * SG + 2] : T(0);
        T *acc = acc0 + r * ncap;
        T v[kPre], o[kPre];
#pragma unroll
        for (int t = 0; t < kPre; ++t)
        {
            v[t] = a0 * P.v0[t];
            if (SZ > 1)
                v[t] = fma(a1, P.v1[t], v[t]);
            if (SZ > 2)
                v[t] = fma(a2, P.v2[t], v[t]);
            o[t] = acc[rk[t]];
        }
#pragma unroll
        for (int t = 0; t < kPre; ++t)
            if (act[t])
                acc[rk[t]] = o[t] + v[t];
    }
}

// =========================================================================================
// Rank-mapped ("compact") window, one warp per group of up to three twin rows of A.
// The symbolic pass kept the occupancy bitmap of these rows (bm_store, <= 64 words).  A warp
// scan of the word popcounts gives, for every 32-column word, the number of C entries in
// front of it; a product with column c then accumulates into
//     acc[ prefix[word(c)] + popc(mask[word(c)] & bits_below(c)) ]        (one LDS.64 lookup)
// i.e. directly into its final CSR position.  Against the dense window this needs no
// "unset" marker (init 0, always add), no compaction sweep, and 8 bytes per ENTRY instead of
// per COLUMN of the span -- which is what makes three accumulator sets per warp affordable
// (the dense-window row-twin kernel above drops to 8 warps/SM and loses).  The lookup is
// shared by the three twin rows: v_r = sum_j a[r][j] * b_j lands in acc_r[rank].
// smem per warp: lut[66] (mask, prefix) | 3 x acc[ncap] | stage_be[G+2] | stage_a[3][G+2]
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(kRowTwinThreads)
    k_num_compact_rowtwins(RowList list, const int *__restrict__ Ap,
                           const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                           const int *__restrict__ Bc, const T *__restrict__ Bv, const int4 *__restrict__ arow,
                           const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int ncap,
                           const unsigned char *__restrict__ bsame, const unsigned char *__restrict__ asame,
                           const unsigned *__restrict__ bm_store, const int *__restrict__ bm_slot)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int G = 32, WPB = kRowTwinThreads / 32, SG = G + 2, LUT = SB_BM_STORE_WORDS + 2;
    const int warp = threadIdx.x >> 5, l = lane_id();
    // layout (per block): acc[WPB][3][ncap] T | stage_a[WPB][3][SG] T | lut[WPB][LUT] uint2 | stage_be[WPB][SG] int2
    T *acc0 = reinterpret_cast<T *>(sm_raw) + (size_t)warp * 3 * ncap;
    T *stage_a = reinterpret_cast<T *>(sm_raw) + (size_t)WPB * 3 * ncap + (size_t)warp * 3 * SG;
    uint2 *lut = reinterpret_cast<uint2 *>(reinterpret_cast<T *>(sm_raw) + (size_t)WPB * 3 * (ncap + SG)) +
                 (size_t)warp * LUT;
    int2 *stage_be = reinterpret_cast<int2 *>(reinterpret_cast<uint2 *>(
                         reinterpret_cast<T *>(sm_raw) + (size_t)WPB * 3 * (ncap + SG)) + (size_t)WPB * LUT) +
                     (size_t)warp * SG;
    for (int chunk = blockIdx.x * WPB + warp; chunk * 3 < nrows; chunk += gridDim.x * WPB)
    {
        const int base = chunk * 3, cnt3 = min(3, nrows - base);
        int j = 0;
        while (j < cnt3)
        {
            const int row = __ldg(&rows[base + j]);
            int R = 1;
            if (j + 1 < cnt3 && __ldg(&rows[base + j + 1]) == row + 1 && __ldg(&asame[row + 1]))
            {
                R = 2;
                if (j + 2 < cnt3 && __ldg(&rows[base + j + 2]) == row + 2 && __ldg(&asame[row + 2]))
                    R = 3;
            }
            j += R;
            const int4 info = __ldg(&arow[row]);
            const int tbase = info.z >> MHB_TILE_SHIFT;
            const int wt = (info.w >> MHB_TILE_SHIFT) - tbase + 1; // <= 64 (SB_BM_G8 rows)
            // lookup table: (mask, entries in front of the word), two words per lane
            const unsigned *bmrow = bm_store + (size_t)__ldg(&bm_slot[row]) * SB_BM_STORE_WORDS;
            const unsigned m0 = (l < wt) ? __ldg(&bmrow[l]) : 0u;
            const unsigned m1 = (l + 32 < wt) ? __ldg(&bmrow[l + 32]) : 0u;
            int c0 = __popc(m0), c1 = __popc(m1);
            int i0 = c0, i1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t0 = __shfl_up_sync(kFull, i0, o), t1 = __shfl_up_sync(kFull, i1, o);
                if (l >= o)
                    i0 += t0, i1 += t1;
            }
            const int n0 = __shfl_sync(kFull, i0, 31);
            const int n = n0 + __shfl_sync(kFull, i1, 31); // nnz of the row(s)
            lut[l] = make_uint2(m0, (unsigned)(i0 - c0));
            lut[l + 32] = make_uint2(m1, (unsigned)(n0 + i1 - c1));
            for (int r = 0; r < R; ++r)
                for (int i = l; i < n; i += G)
                    acc0[r * ncap + i] = T(0);
            if (l < 2)
            {
                stage_be[G + l] = make_int2(0, 0);
                stage_a[0 * SG + G + l] = stage_a[1 * SG + G + l] = stage_a[2 * SG + G + l] = T(0);
            }
            __syncwarp();
            const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
            const int d1 = (R > 1) ? __ldg(&Ap[row + 1]) - s : 0, d2 = (R > 2) ? __ldg(&Ap[row + 2]) - s : 0;
            int bs, be, kk, nbs, nbe, nkk;
            T av0, av1 = T(0), av2 = T(0), nav0, nav1 = T(0), nav2 = T(0);
            auto meta = [&](int jj, int &ms, int &me, int &mk, T &a0, T &a1, T &a2) {
                ms = 0, me = 0, mk = -2, a0 = T(0), a1 = T(0), a2 = T(0);
                if (jj < e)
                {
                    mk = __ldg(&Ac[jj]);
                    a0 = __ldg(&Av[jj]);
                    if (R > 1)
                        a1 = __ldg(&Av[jj + d1]);
                    if (R > 2)
                        a2 = __ldg(&Av[jj + d2]);
                    ms = __ldg(&Bp[mk]);
                    me = __ldg(&Bp[mk + 1]);
                    if (__ldg(&bsame[mk]))
                        mk |= kTwinTag;
                }
            };
            meta(s + l, bs, be, kk, av0, av1, av2);
            for (int j0 = s; j0 < e; j0 += G)
            {
                meta(j0 + G + l, nbs, nbe, nkk, nav0, nav1, nav2);
                const int cnt = min(G, e - j0);
                stage_be[l] = make_int2(bs, be);
                stage_a[0 * SG + l] = av0;
                stage_a[1 * SG + l] = av1;
                stage_a[2 * SG + l] = av2;
                const int kprev = __shfl_up_sync(kFull, kk & kTwinMask, 1);
                const bool fol = l > 0 && l < cnt && (kk & kTwinTag) && (kk & kTwinMask) == kprev + 1;
                const unsigned fmask = __ballot_sync(kFull, fol);
                __syncwarp();
                // two register sets: while one B-row step is accumulated the loads of the next are in flight
                CompactStep<T> PA, PB;
                auto issue = [&](CompactStep<T> &P, int i) {
                    P.ci = i;
                    int sz = 1 + ((fmask >> (i + 1)) & 1u);
                    if (sz == 2)
                        sz += (fmask >> (i + 2)) & 1u;
                    if (i + sz > cnt)
                        sz = cnt - i;
                    P.sz = sz;
                    const int2 e0 = stage_be[i], e1 = stage_be[i + 1], e2 = stage_be[i + 2];
                    P.q = e0.x;
                    P.qe = e0.y;
                    P.o1 = e1.x - e0.x;
                    P.o2 = e2.x - e0.x;
                    const int *pcol = Bc + e0.x + l;
                    const T *p0 = Bv + e0.x + l, *p1 = p0 + P.o1, *p2 = p0 + P.o2;
                    const int left = e0.y - e0.x - l; // entries of the B row at or behind this lane
#pragma unroll
                    for (int t = 0; t < kPre; ++t)
                    {
                        P.c[t] = -1;
                        if (t * G < left)
                        {
                            P.c[t] = __ldg(pcol + t * G);
                            P.v0[t] = __ldg(p0 + t * G);
                            if (sz > 1)
                                P.v1[t] = __ldg(p1 + t * G);
                            if (sz > 2)
                                P.v2[t] = __ldg(p2 + t * G);
                        }
                    }
                };
                auto rank_of = [&](int c) {
                    const uint2 w = lut[(c >> MHB_TILE_SHIFT) - tbase];
                    return (int)w.y + __popc(w.x & ((1u << (c & 31)) - 1u));
                };
                auto consume = [&](const CompactStep<T> &P) {
                    int rk[kPre];
                    bool act[kPre];
#pragma unroll
                    for (int t = 0; t < kPre; ++t)
                    {
                        act[t] = P.c[t] >= 0;
                        rk[t] = act[t] ? rank_of(P.c[t]) : 0; // idle lanes point at entry 0 and store nothing
                    }
                    const T *sa = stage_a + P.ci;
                    // straight-line code for every (twin rows of A) x (folded rows of B) shape
                    if (R == 3)
                    {
                        if (P.sz == 3)
                            compact_accumulate<T, 3, 3>(acc0, ncap, sa, SG, rk, act, P);
                        else if (P.sz == 2)
                            compact_accumulate<T, 3, 2>(acc0, ncap, sa, SG, rk, act, P);
                        else
                            compact_accumulate<T, 3, 1>(acc0, ncap, sa, SG, rk, act, P);
                    }
                    else if (R == 2)
                    {
                        if (P.sz == 3)
                            compact_accumulate<T, 2, 3>(acc0, ncap, sa, SG, rk, act, P);
                        else if (P.sz == 2)
                            compact_accumulate<T, 2, 2>(acc0, ncap, sa, SG, rk, act, P);
                        else
                            compact_accumulate<T, 2, 1>(acc0, ncap, sa, SG, rk, act, P);
                    }
                    else
                    {
                        if (P.sz == 3)
                            compact_accumulate<T, 1, 3>(acc0, ncap, sa, SG, rk, act, P);
                        else if (P.sz == 2)
                            compact_accumulate<T, 1, 2>(acc0, ncap, sa, SG, rk, act, P);
                        else
                            compact_accumulate<T, 1, 1>(acc0, ncap, sa, SG, rk, act, P);
                    }
                    if (P.qe - P.q > kPre * G) // B rows longer than kPre*G (warp-uniform, rare)
                    {
                        for (int r = 0; r < R; ++r)
                        {
                            const T a0 = sa[r * SG];
                            const T a1 = (P.sz > 1) ? sa[r * SG + 1] : T(0);
                            const T a2 = (P.sz > 2) ? sa[r * SG + 2] : T(0);
                            T *acc = acc0 + r * ncap;
                            for (int p = P.q + kPre * G + l; p < P.qe; p += G)
                            {
                                T v = a0 * __ldg(&Bv[p]);
                                if (P.sz > 1)
                                    v = fma(a1, __ldg(&Bv[p + P.o1]), v);
                                if (P.sz > 2)
                                    v = fma(a2, __ldg(&Bv[p + P.o2]), v);
                                const int k = rank_of(__ldg(&Bc[p]));
                                acc[k] += v;
                            }
                        }
                    }
                    __syncwarp();
                };
                issue(PA, 0);
                for (int i = 0;;)
                {
                    i += PA.sz;
                    if (i < cnt)
                        issue(PB, i);
                    consume(PA);
                    if (i >= cnt)
                        break;
                    i += PB.sz;
                    if (i < cnt)
                        issue(PA, i);
                    consume(PB);
                    if (i >= cnt)
                        break;
                }
                bs = nbs, be = nbe, kk = nkk, av0 = nav0, av1 = nav1, av2 = nav2;
            }
            // values are already in CSR order: coalesced copies; columns come from the bitmap
            const int out0 = __ldg(&Cp[row]);
            const int q1 = (R > 1) ? __ldg(&Cp[row + 1]) : 0, q2 = (R > 2) ? __ldg(&Cp[row + 2]) : 0;
            for (int i = l; i < n; i += G)
            {
                Cv[out0 + i] = acc0[i];
                if (R > 1)
                    Cv[q1 + i] = acc0[ncap + i];
                if (R > 2)
                    Cv[q2 + i] = acc0[2 * ncap + i];
            }
            for (int w = l; w < wt; w += G)
            {
                const uint2 e = lut[w];
                unsigned m = e.x;
                int k = (int)e.y;
                const int cbase = ((tbase + w) << MHB_TILE_SHIFT);
                while (m)
                {
                    const int c = cbase + __ffs(m) - 1;
                    m &= m - 1;
                    Cc[out0 + k] = c;
                    if (R > 1)
                        Cc[q1 + k] = c;
                    if (R > 2)
                        Cc[q2 + k] = c;
                    ++k;
                }
            }
            __syncwarp();
        }
    }
}

// =========================================================================================
// Block-wide helpers
// =========================================================================================
// exclusive scan of one int per thread across the block; returns exclusive prefix, *total
__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot /*32*/, int *total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(kFull, incl, o);
        if (lane_id() >= o)
            incl += t;
    }
    __syncthreads();
    if (lane_id() == 31)
        warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    // every warp scans the (at most 32) warp totals itself: ten shuffles instead of a serial
    // loop over the warps in every thread
    const int nw = blockDim.x >> 5;
    const int wt = lane_id() < nw ? warp_tot[lane_id()] : 0;
    int wi = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(kFull, wi, o);
        if (lane_id() >= o)
            wi += t;
    }
    const int pre = __shfl_sync(kFull, wi - wt, threadIdx.x >> 5);
    *total = __shfl_sync(kFull, wi, 31);
    return pre + incl - v;
}

// =========================================================================================
// Dense window, one block per row (W up to 27 648 columns).  Warps work on different B rows
// at once, so values use atomicAdd and presence is a bitmap set with atomicOr.
// smem: acc[wcap] | flags[wcap/32] | wpre[wcap/32]
// =========================================================================================
template <typename T>
__global__ void k_num_win_block(RowList list, const int *__restrict__ Ap,
                                const int *__restrict__ Ac, const T *__restrict__ Av,
                                const int *__restrict__ Bp, const int *__restrict__ Bc,
                                const T *__restrict__ Bv, const int4 *__restrict__ arow,
                                const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int wcap)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int warp_tot[32];
    T *acc = reinterpret_cast<T *>(sm_raw);
    unsigned *flags = reinterpret_cast<unsigned *>(acc + wcap);
    int *wpre = reinterpret_cast<int *>(flags + wcap / 32);
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        const int4 info = arow[row];
        const int cmin = info.z;
        const int W = info.w - cmin + 1;
        const int nwords = (W + 31) >> 5;
        for (int i = threadIdx.x; i < W; i += blockDim.x)
            acc[i] = T(0);
        for (int i = threadIdx.x; i < nwords; i += blockDim.x)
            flags[i] = 0u;
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        walk_flat<32, T, T>(kFull, lane, s, e, warp, nwarp, Ac, Av, Bp, Bc, Bv, [&](int c, T v, T a) {
            const int idx = c - cmin;
            atomicAdd(&acc[idx], a * v);
            const unsigned bit = 1u << (idx & 31);
            if (!(flags[idx >> 5] & bit))
                atomicOr(&flags[idx >> 5], bit);
        });
        __syncthreads();
        // prefix of the per-word popcounts -> output position of every present column
        int carry = 0;
        for (int w0 = 0; w0 < nwords; w0 += blockDim.x)
        {
            const int w = w0 + threadIdx.x;
            const int c = (w < nwords) ? __popc(flags[w]) : 0;
            int tot;
            const int ex = block_excl_scan(c, warp_tot, &tot);
            if (w < nwords)
                wpre[w] = carry + ex;
            carry += tot;
        }
        __syncthreads();
        const int out = __ldg(&Cp[row]);
        for (int i = threadIdx.x; i < W; i += blockDim.x)
        {
            const unsigned f = flags[i >> 5];
            if ((f >> (i & 31)) & 1u)
            {
                const int pos = out + wpre[i >> 5] + __popc(f & ((1u << (i & 31)) - 1u));
                Cc[pos] = cmin + i;
                Cv[pos] = acc[i];
            }
        }
        __syncthreads();
    }
}

// =========================================================================================
// Hash accumulation
// =========================================================================================
// Find-or-claim the slot of `key`; returns the slot, or -1 if the table is full.
__device__ __forceinline__ int key_slot(int *keys, int logS, int key, int &np)
{
    const unsigned S1 = (1u << logS) - 1u;
    unsigned h = hash_slot((unsigned)key, logS);
    for (unsigned it = 0; it <= S1; ++it)
    {
        int old = keys[h];
        if (old == key)
            return (int)h;
        if (old == -1)
        {
            old = atomicCAS(&keys[h], -1, key);
            if (old == -1 || old == key)
                return (int)h;
        }
        ++np; // slot taken by another key (the reference's HASH_CONFLICT event)
        h = (h + 1) & S1;
    }
    return -1;
}

// In-place bitonic sort of (keys[0..P), vals[0..P)) by key, P a power of two, executed by
// `nthr` cooperating threads (tid in [0, nthr)); `sync` orders the stages.
template <typename T, class Sync>
__device__ __forceinline__ void bitonic_sort_kv(int *keys, T *vals, int P, int tid, int nthr, Sync sync)
{
    for (int k = 2; k <= P; k <<= 1)
    {
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int t = tid; t < (P >> 1); t += nthr)
            {
                // t-th compare-exchange pair of this stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int x = i | j;
                const int a = keys[i], b = keys[x];
                const bool asc = (i & k) == 0;
                if ((a > b) == asc)
                {
                    keys[i] = b;
                    keys[x] = a;
                    const T va = vals[i];
                    vals[i] = vals[x];
                    vals[x] = va;
                }
            }
            sync();
        }
    }
}

// -----------------------------------------------------------------------------------------
// Bucket-rank sort of the compacted (keys[0..n), vals[0..n)) of one C row, emitted straight
// into C.  The first profile of the hash kernels (profiles/r1b_*) showed the bitonic sort
// taking ~2/3 of their instructions.  Columns of a C row lie in [cmin, cmin + W): bucket =
// (key - cmin) >> sh splits that span into NB = S/8 equal ranges (a counting sort on the
// top bits), and inside a bucket (a handful of keys) the rank is found by direct
// comparison.  Scratch lives in the unused tail of the table (fill <= 5/8):
//   keys[n ..]            : start[NB+1], cursor[NB]      (NB = S/8  ->  S/4+1 ints <= 3S/8)
//   vals[n ..] as ushort  : idx[n] = indices grouped by bucket (2n bytes <= (S-n)*sizeof(T))
// ~40 instructions per element instead of ~220 (n = 640).  Returns false without emitting
// when one bucket holds more than kBucketMax keys (heavily clustered columns): the caller
// then falls back to the bitonic sort.
// -----------------------------------------------------------------------------------------
constexpr int kBucketMax = 96;

__device__ __forceinline__ int ceil_log2_dev(int v) { return v <= 1 ? 0 : 32 - __clz(v - 1); }

template <typename T>
__device__ __forceinline__ bool bucket_sort_emit_warp(int *keys, T *vals, int n, int logS, int cmin, int W,
                                                      int l, int *__restrict__ Cc, T *__restrict__ Cv)
{
    const int logNB = logS - 3, NB = 1 << logNB;
    const int sh = max(0, ceil_log2_dev(W) - logNB);
    int *start = keys + n;            // NB + 1
    int *cursor = start + NB + 1;     // NB
    unsigned short *idx = reinterpret_cast<unsigned short *>(vals + n);
    for (int b = l; b < NB; b += 32)
    {
        start[b] = 0;
        cursor[b] = 0;
    }
    __syncwarp();
    for (int i = l; i < n; i += 32)
        atomicAdd(&start[(keys[i] - cmin) >> sh], 1);
    __syncwarp();
    // exclusive scan of the NB counts: lane l owns NB/32 consecutive buckets (NB >= 32)
    // lane l owns `per` consecutive buckets (NB = 16 .. 128: per = 1, 1, 2, 4; lanes >= NB idle)
    const int per = max(NB >> 5, 1);
    int cnt[4], tot = 0, mx = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
        const int b = l * per + t;
        cnt[t] = (t < per && b < NB) ? start[b] : 0;
        tot += cnt[t];
        mx = max(mx, cnt[t]);
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (l >= o)
            incl += t;
    }
    mx = group_max<32>(mx, kFull);
    if (mx > kBucketMax)
        return false;
    int run = incl - tot;
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
        const int b = l * per + t;
        if (t < per && b < NB)
        {
            start[b] = run;
            run += cnt[t];
        }
    }
    if (l == 31)
        start[NB] = n;
    __syncwarp();
    for (int i = l; i < n; i += 32)
    {
        const int b = (keys[i] - cmin) >> sh;
        idx[start[b] + atomicAdd(&cursor[b], 1)] = (unsigned short)i;
    }
    __syncwarp();
    for (int p = l; p < n; p += 32)
    {
        const int i = idx[p];
        const int k = keys[i];
        const int b = (k - cmin) >> sh;
        const int lo = start[b], hi = start[b + 1];
        int rank = lo;
        for (int q = lo; q < hi; ++q)
            rank += keys[idx[q]] < k;
        Cc[rank] = k;
        Cv[rank] = vals[i];
    }
    return true;
}

// Block-wide variant (one block per row, S up to 16 384): same layout, scan via block_excl_scan.
template <typename T>
__device__ __forceinline__ bool bucket_sort_emit_block(int *keys, T *vals, int n, int logS, int cmin, int W,
                                                       int *warp_tot, int *__restrict__ Cc,
                                                       T *__restrict__ Cv)
{
    const int logNB = logS - 3, NB = 1 << logNB;
    const int sh = max(0, ceil_log2_dev(W) - logNB);
    const int tid = threadIdx.x, nthr = blockDim.x;
    int *start = keys + n;
    int *cursor = start + NB + 1;
    unsigned short *idx = reinterpret_cast<unsigned short *>(vals + n);
    for (int b = tid; b < NB; b += nthr)
    {
        start[b] = 0;
        cursor[b] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr)
        atomicAdd(&start[(keys[i] - cmin) >> sh], 1);
    __syncthreads();
    int carry = 0, mx = 0;
    for (int b0 = 0; b0 < NB; b0 += nthr)
    {
        const int b = b0 + tid;
        const int c = (b < NB) ? start[b] : 0;
        mx = max(mx, c);
        int tot;
        const int ex = block_excl_scan(c, warp_tot, &tot);
        if (b < NB)
            start[b] = carry + ex;
        carry += tot;
    }
    if (tid == 0)
        start[NB] = n;
    const bool bad = __syncthreads_or(mx > kBucketMax);
    if (bad)
        return false;
    for (int i = tid; i < n; i += nthr)
    {
        const int b = (keys[i] - cmin) >> sh;
        idx[start[b] + atomicAdd(&cursor[b], 1)] = (unsigned short)i;
    }
    __syncthreads();
    for (int p = tid; p < n; p += nthr)
    {
        const int i = idx[p];
        const int k = keys[i];
        const int b = (k - cmin) >> sh;
        const int lo = start[b], hi = start[b + 1];
        int rank = lo;
        for (int q = lo; q < hi; ++q)
            rank += keys[idx[q]] < k;
        Cc[rank] = k;
        Cv[rank] = vals[i];
    }
    return true;
}

// ---- hash, G lanes per row, table of 2^logS (key, value) slots per group -----------------
template <int G, typename T>
__global__ void __launch_bounds__(kNumGroupThreads)
    k_num_hash_group(RowList list, const int *__restrict__ Ap,
                     const int *__restrict__ Ac, const T *__restrict__ Av, const int *__restrict__ Bp,
                     const int *__restrict__ Bc, const T *__restrict__ Bv, const int4 *__restrict__ arow,
                     const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int logS,
                     int *__restrict__ scal, unsigned long long *__restrict__ probes)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int GPB = kNumGroupThreads / G;
    int np = 0;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int S = 1 << logS;
    // layout: all value tables first (8-byte aligned), then all key tables
    T *vals = reinterpret_cast<T *>(sm_raw) + (size_t)g * S;
    int *keys = reinterpret_cast<int *>(reinterpret_cast<T *>(sm_raw) + (size_t)GPB * S) + (size_t)g * S;
    for (int r = blockIdx.x * GPB + g; r < nrows; r += gridDim.x * GPB)
    {
        const int row = rows[r];
        for (int i = l; i < S; i += G)
        {
            keys[i] = -1;
            vals[i] = T(0);
        }
        __syncwarp(gm);
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // B rows long enough to fill the group: sequential walk, plain accumulate.  Short B
        // rows (power-law graphs): flat expansion, all lanes busy, atomic accumulate.
        const int products = __ldg(&arow[row]).x;
        if ((long long)products * 2 >= (long long)(e - s) * G)
            walk_sequential<G, 2, T, T>(gm, l, s, e, Ac, Av, Bp, Bc, Bv, [&](int c, T v, T a) {
                const int h = key_slot(keys, logS, c, np);
                if (h >= 0)
                    vals[h] = fma(a, v, vals[h]);
                else
                    atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
            });
        else
            walk_flat<G, T, T>(gm, l, s, e, 0, 1, Ac, Av, Bp, Bc, Bv, [&](int c, T v, T a) {
                const int h = key_slot(keys, logS, c, np);
                if (h >= 0)
                    atomicAdd(&vals[h], a * v);
                else
                    atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
            });
        __syncwarp(gm);
        // in-place compaction to the front of the table, G slots per step
        int n = 0;
        for (int i0 = 0; i0 < S; i0 += G)
        {
            const int k = keys[i0 + l];
            const T v = vals[i0 + l];
            __syncwarp(gm);
            const bool p = k != -1;
            const unsigned bal = __ballot_sync(gm, p);
            if (p)
            {
                const int pos = n + __popc(bal & lanemask_lt());
                keys[pos] = k;
                vals[pos] = v;
            }
            n += __popc(bal);
            __syncwarp(gm);
        }
        const int out = __ldg(&Cp[row]);
        if (G < 32)
        {
            // tiny rows: rank sort straight into C
            for (int i = l; i < n; i += G)
            {
                const int k = keys[i];
                int rank = 0;
                for (int j = 0; j < n; ++j)
                    rank += keys[j] < k;
                Cc[out + rank] = k;
                Cv[out + rank] = vals[i];
            }
        }
        else if (n <= 32)
        {
            // one key per lane: rank by direct comparison
            const int k = (l < n) ? keys[l] : INT_MAX;
            int rank = 0;
            for (int j = 0; j < n; ++j)
                rank += keys[j] < k;
            if (l < n)
            {
                Cc[out + rank] = k;
                Cv[out + rank] = vals[l];
            }
        }
        else
        {
            const int4 info = __ldg(&arow[row]);
            if (!bucket_sort_emit_warp<T>(keys, vals, n, logS, info.z, info.w - info.z + 1, l, Cc + out, Cv + out))
            {
                int P = 2;
                while (P < n)
                    P <<= 1;
                for (int i = n + l; i < P; i += G)
                    keys[i] = INT_MAX;
                __syncwarp(gm);
                bitonic_sort_kv(keys, vals, P, l, G, [&]() { __syncwarp(gm); });
                for (int i = l; i < n; i += G)
                {
                    Cc[out + i] = keys[i];
                    Cv[out + i] = vals[i];
                }
            }
        }
        __syncwarp(gm);
    }
    flush_probes(probes, np);
}

// Bucket-rank sort for a row whose table lives in the GLOBAL pool: the compacted keys are
// copied into shared memory (4 B each; values stay in the pool), bucketed on the top bits of
// the column into ~n/2 buckets, ranked inside the bucket and emitted; the value is gathered
// from the pool at the end.  A 21 920-entry hub row of the webbase-like input took 1.2 ms in
// the bitonic network over global memory (120 stages x L2 latency); this needs 6n + 4NB bytes
// of shared memory, i.e. rows up to ~32 K entries.  Returns false (nothing emitted) otherwise.
template <typename T>
__device__ __forceinline__ bool pool_bucket_sort_emit(const int *gkeys, const T *gvals, int n, int cmin, int W,
                                                      unsigned char *sm, int sm_bytes, int *warp_tot,
                                                      int *__restrict__ Cc, T *__restrict__ Cv)
{
    if (n > 65535)
        return false;
    int logNB = max(5, 31 - __clz(max(n >> 1, 1)));
    while (logNB > 5 && (size_t)n * 6 + ((size_t)(1 << logNB) + 1) * 4 + 16 > (size_t)sm_bytes)
        --logNB;
    if ((size_t)n * 6 + ((size_t)(1 << logNB) + 1) * 4 + 16 > (size_t)sm_bytes)
        return false;
    const int NB = 1 << logNB;
    const int sh = max(0, ceil_log2_dev(W) - logNB);
    const int tid = threadIdx.x, nthr = blockDim.x;
    int *skey = reinterpret_cast<int *>(sm);
    int *start = skey + n; // [NB + 1] counts -> begins -> ends
    unsigned short *idx = reinterpret_cast<unsigned short *>(start + NB + 1);
    for (int b = tid; b <= NB; b += nthr)
        start[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nthr)
    {
        const int k = __ldcg(&gkeys[i]);
        skey[i] = k;
        atomicAdd(&start[(k - cmin) >> sh], 1);
    }
    __syncthreads();
    int carry = 0;
    for (int b0 = 0; b0 < NB; b0 += nthr)
    {
        const int b = b0 + tid;
        const int c = (b < NB) ? start[b] : 0;
        int tot;
        const int ex = block_excl_scan(c, warp_tot, &tot);
        if (b < NB)
            start[b] = carry + ex;
        carry += tot;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr)
        idx[atomicAdd(&start[(skey[i] - cmin) >> sh], 1)] = (unsigned short)i; // start[b] ends as the END of bucket b
    __syncthreads();
    for (int p = tid; p < n; p += nthr)
    {
        const int i = idx[p];
        const int k = skey[i];
        const int b = (k - cmin) >> sh;
        const int lo = b ? start[b - 1] : 0, hi = start[b];
        int rank = lo;
        for (int q = lo; q < hi; ++q)
            rank += skey[idx[q]] < k;
        Cc[rank] = k;
        Cv[rank] = __ldcg(&gvals[i]);
    }
    return true;
}

// ---- hash, one block per row; table in shared memory or in the global pool ---------------
template <typename T>
__global__ void k_num_hash_block(RowList list, const int *__restrict__ Ap,
                                 const int *__restrict__ Ac, const T *__restrict__ Av,
                                 const int *__restrict__ Bp, const int *__restrict__ Bc,
                                 const T *__restrict__ Bv, const int4 *__restrict__ arow,
                                 const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv,
                                 int logS_fixed, unsigned char *__restrict__ pool,
                                 long long pool_slots, int *__restrict__ scal, int sort_smem,
                                 unsigned long long *__restrict__ probes)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int warp_tot[32];
    int np = 0;
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x)
    {
        const int row = rows[r];
        const int out = __ldg(&Cp[row]);
        const int n_row = __ldg(&Cp[row + 1]) - out;
        int logS = logS_fixed;
        T *vals;
        int *keys;
        if (pool)
        {
            logS = 10;
            while ((1LL << logS) < 2LL * n_row)
                ++logS;
            if ((1LL << logS) > pool_slots) // a speculative launch sized the pool from the previous call's largest row: redo
            {
                if (threadIdx.x == 0)
                    atomicMax(scal + SC_SPEC_MISS, 1);
                continue;
            }
            unsigned char *base = pool + (size_t)blockIdx.x * (size_t)pool_slots * (sizeof(T) + sizeof(int));
            vals = reinterpret_cast<T *>(base);
            keys = reinterpret_cast<int *>(vals + pool_slots);
        }
        else
        {
            // the table of this row: smallest power of two >= 2 n (>= 1 024), at most the bin's;
            // initialisation, compaction sweep and sort scratch then scale with the row
            int lrow = 10;
            while ((1 << lrow) < 2 * n_row)
                ++lrow;
            logS = min(logS_fixed, lrow);
            vals = reinterpret_cast<T *>(sm_raw);
            keys = reinterpret_cast<int *>(vals + ((size_t)1 << logS));
        }
        const int S = 1 << logS;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
        {
            keys[i] = -1;
            vals[i] = T(0);
        }
        __syncthreads();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        walk_flat<32, T, T>(kFull, lane, s, e, warp, nwarp, Ac, Av, Bp, Bc, Bv, [&](int c, T v, T a) {
            const int h = key_slot(keys, logS, c, np);
            if (h >= 0)
                atomicAdd(&vals[h], a * v);
            else
                atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
        });
        __syncthreads();
        // in-place compaction, blockDim slots per step
        int n = 0;
        for (int i0 = 0; i0 < S; i0 += blockDim.x)
        {
            const int i = i0 + threadIdx.x;
            int k = -1;
            T v = T(0);
            if (i < S)
            {
                k = pool ? __ldcg(&keys[i]) : keys[i];
                v = pool ? __ldcg(&vals[i]) : vals[i];
            }
            const bool p = k != -1;
            int tot;
            const int ex = block_excl_scan(p ? 1 : 0, warp_tot, &tot); // has __syncthreads inside
            if (p)
            {
                keys[n + ex] = k;
                vals[n + ex] = v;
            }
            n += tot;
            __syncthreads();
        }
        bool done = false;
        if (!pool && n > 0)
        {
            const int4 info = __ldg(&arow[row]);
            done = bucket_sort_emit_block<T>(keys, vals, n, logS, info.z, info.w - info.z + 1, warp_tot, Cc + out,
                                             Cv + out);
        }
        else if (pool && n > 0 && sort_smem > 0)
        {
            const int4 info = __ldg(&arow[row]);
            done = pool_bucket_sort_emit<T>(keys, vals, n, info.z, info.w - info.z + 1, sm_raw, sort_smem, warp_tot,
                                            Cc + out, Cv + out);
        }
        if (!done)
        {
            int P = 2;
            while (P < n)
                P <<= 1;
            for (int i = n + threadIdx.x; i < P; i += blockDim.x)
                keys[i] = INT_MAX;
            __syncthreads();
            bitonic_sort_kv(keys, vals, P, (int)threadIdx.x, (int)blockDim.x, [&]() { __syncthreads(); });
            for (int i = threadIdx.x; i < n; i += blockDim.x)
            {
                Cc[out + i] = keys[i];
                Cv[out + i] = vals[i];
            }
        }
        __syncthreads();
    }
    flush_probes(probes, np);
}

// ---- hash with a claim list: one block per row, table in shared memory ------------------
// The table is cleared once per block.  A lane that claims an empty slot appends the slot
// index to a list (one warp-aggregated shared-memory atomicAdd per step) and counts the key
// into its column bucket, so after the walk the row's entries and the bucket histogram are
// known without sweeping the table: no per-row initialisation, no compaction pass, no
// histogram pass.  The probe loop only finds the slot; the value is added once, after the
// lanes have reconverged (one LDS/DADD/CAS sequence per step instead of one per divergent
// path).  fp64 atomicAdd on shared memory is a CAS loop that costs ~2 LSU wavefronts per LANE
// (ncu: a third of the kernel's shared-memory traffic), so a single-warp block avoids it where
// it can: the lane that claimed a slot owns its first value and writes it with a plain store
// (which also makes zeroing the values unnecessary); only lanes that hit an existing key add
// atomically, after a __syncwarp, and rows with little compression have almost none of those.  The bucket-rank sort (S/4 buckets: ~2 keys per bucket) scatters (slot, key) pairs,
// ranks each key inside its bucket, emits into C and resets the slot on the way out.
// WROWS: every WARP of the block owns a row and a table of its own (table_bytes apart) and
// synchronises with __syncwarp only -- for the small tables this lifts the 32-blocks-per-SM
// limit of one-warp blocks (32 -> 48..64 resident warps), which is what a latency-bound
// kernel needs.
// smem: vals[S] | keys[S] | start[NB+1] | misc[3] | bkey[5S/8] | list[5S/8] u16 | idx[5S/8] u16
template <typename T, bool WROWS>
__global__ void __launch_bounds__(WROWS ? 128 : 1024, WROWS ? 10 : 1) k_num_hash_list(RowList rowlist, const int *__restrict__ Ap,
                                const int *__restrict__ Ac, const T *__restrict__ Av,
                                const int *__restrict__ Bp, const int *__restrict__ Bc,
                                const T *__restrict__ Bv, const int4 *__restrict__ arow,
                                const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv, int logS,
                                int *__restrict__ scal, int table_bytes, unsigned long long *__restrict__ probes)
{
    const int *__restrict__ rows = rowlist.begin();
    const int nrows = rowlist.size();
    int np = 0;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ int warp_tot[32];
    const int S = 1 << logS, NBmax = S >> 2, nmax = (S >> 3) * 5;
    T *vals = reinterpret_cast<T *>(sm_raw + (WROWS ? (size_t)(threadIdx.x >> 5) * table_bytes : 0));
    int *keys = reinterpret_cast<int *>(vals + S);
    int *start = keys + S;         // [NB + 1] bucket counts -> bucket begins -> bucket ends
    int *misc = start + NBmax + 1; // [0] entries claimed so far
    int *bkey = misc + 3;       // keys in bucket order
    unsigned short *list = reinterpret_cast<unsigned short *>(bkey + nmax);
    unsigned short *idx = list + nmax; // slots in bucket order
    const int lane = lane_id();
    const int tid = WROWS ? lane : (int)threadIdx.x, nthr = WROWS ? 32 : (int)blockDim.x;
    const int warp = WROWS ? 0 : tid >> 5, nwarp = nthr >> 5;
    const bool solo = nthr == 32; // one warp per row: claimed slots are written, not added to
    const int row0 = WROWS ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
    const int rstep = WROWS ? gridDim.x * (blockDim.x >> 5) : gridDim.x;
    auto bar = [&]() {
        if (WROWS)
            __syncwarp();
        else
            __syncthreads();
    };
    for (int i = tid; i < S; i += nthr)
    {
        keys[i] = -1;
        vals[i] = T(0);
    }
    for (int r = row0; r < nrows; r += rstep)
    {
        const int row = rows[r];
        const int out = __ldg(&Cp[row]);
        const int n = __ldg(&Cp[row + 1]) - out;
        const int4 info = __ldg(&arow[row]);
        const int cmin = info.z, W = info.w - info.z + 1;
        // table of THIS row: the smallest power of two with fill <= 5/8 (the shared-memory layout
        // is that of the bin's largest table; a smaller row uses a prefix of it, so its bucket
        // arrays, scans and probe cycles are sized by the row, not by the bin)
        const int lS = min(logS, max(5, ceil_log2_dev((n * 8 + 4) / 5)));
        const int NB = 1 << (lS - 2);
        const int sh = max(0, ceil_log2_dev(W) - (lS - 2));
        for (int b = tid; b <= NB; b += nthr)
            start[b] = 0;
        if (tid == 0)
            misc[0] = 0;
        bar();
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // A row whose nnz equals its product count has no two products on one column: nothing to
        // find, nothing to add up.  Its products go straight into the front of the table arrays in
        // expansion order (keys[pos], vals[pos]) -- no hashing, no CAS, no claim list -- and only the
        // sort below is left.  On power-law graphs that is most rows (webbase-like input: 94 % of the
        // rows of these bins, 57 % of all products).
        const bool uniq = info.x == n;
        // ... and a row with FEW repeated columns (products <= the entries the bin's arrays hold) goes the
        // same way -- expand, sort, compress: the sort below then ranks equal columns by position, marks
        // the first of each as its head, and the others are added onto their head's entry of C.  Banded
        // random / power-law rows have compression 1.00-1.05: nearly all of them qualify.
        const bool esc = !uniq && info.x > n && info.x <= nmax;
        const bool direct = uniq || esc;
        const int ne = direct ? info.x : n; // entries to sort
        if (direct)
            walk_flat_indexed<32, T, T>(kFull, lane, s, e, warp, nwarp, Ac, Av, Bp, Bc, Bv, [&](int pos, int c, T v, T a) {
                keys[pos] = c;
                vals[pos] = a * v;
                atomicAdd(&start[(c - cmin) >> sh], 1);
            });
        else
        walk_flat_post<32, T, T>(
            kFull, lane, s, e, warp, nwarp, Ac, Av, Bp, Bc, Bv,
            [&](int c, T v, T a) {
                // find-or-claim, all lanes in lockstep (c < 0: nothing to insert): the loop leaves
                // when every lane has its slot, so the value update below runs once per step
                const unsigned S1 = (1u << lS) - 1u;
                unsigned h = hash_slot((unsigned)c, lS);
                int claimed = -1;
                bool need = c >= 0;
                for (unsigned it = 0; it <= S1 && __any_sync(kFull, need); ++it)
                {
                    int old = -2;
                    if (need)
                        old = atomicCAS(&keys[h], -1, c);
                    if (old == -1)
                        claimed = (int)h;
                    need = need && old != -1 && old != c;
                    if (need)
                    {
                        ++np;
                        h = (h + 1) & S1;
                    }
                }
                if (need) // S probes without a home: the row has more entries than symbolic promised
                    atomicMax(scal + SC_ERROR, (int)DEVERR_TABLE_FULL);
                const bool ok = c >= 0 && !need;
                if (!solo)
                {
                    if (ok)
                        atomicAdd(&vals[h], a * v);
                }
                else
                {
                    if (claimed >= 0)
                        vals[h] = a * v;
                    __syncwarp();
                    const bool hit = ok && claimed < 0;
                    if (__any_sync(kFull, hit))
                        if (hit)
                            atomicAdd(&vals[h], a * v);
                }
                if (claimed >= 0)
                    atomicAdd(&start[(c - cmin) >> sh], 1);
                return claimed;
            },
            [&](int claimed) {
                const unsigned cm = __ballot_sync(kFull, claimed >= 0);
                if (cm)
                {
                    const int leader = __ffs(cm) - 1;
                    int base = 0;
                    if (lane == leader)
                        base = atomicAdd(&misc[0], __popc(cm));
                    base = __shfl_sync(kFull, base, leader);
                    if (claimed >= 0)
                    {
                        const int pos = base + __popc(cm & lanemask_lt());
                        if (pos < nmax)
                            list[pos] = (unsigned short)claimed;
                    }
                }
            });
        bar();
        // bucket counts -> bucket begins (see bucket_sort_emit_warp for the idea of the sort)
        int carry = 0, mx = 0;
        for (int b0 = 0; b0 < NB; b0 += nthr)
        {
            const int b = b0 + tid;
            const int c = (b < NB) ? start[b] : 0;
            mx = max(mx, c);
            int tot, ex;
            if (nthr == 32)
            {
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const int t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o)
                        incl += t;
                }
                ex = incl - c;
                tot = __shfl_sync(kFull, incl, 31);
            }
            else
                ex = block_excl_scan(c, warp_tot, &tot);
            if (b < NB)
                start[b] = carry + ex;
            carry += tot;
        }
        bool clustered;
        if (WROWS)
        {
            __syncwarp();
            clustered = __any_sync(kFull, mx > kBucketMax);
        }
        else
            clustered = __syncthreads_or(mx > kBucketMax);
        if (!clustered || esc)
        {
            for (int i = tid; i < ne; i += nthr)
            {
                const unsigned short sl = direct ? (unsigned short)i : list[i];
                const int k = keys[sl];
                const int pos = atomicAdd(&start[(k - cmin) >> sh], 1); // start[b] ends as the END of bucket b
                idx[pos] = sl;
                bkey[pos] = k;
            }
            bar();
            if (esc)
            {
                // stable rank (column, then position) of every product; heads = first of their column
                unsigned *hb = reinterpret_cast<unsigned *>(vals + nmax); // head bits | heads in front of each word
                const int nW = (ne + 31) >> 5;
                int *hpre = reinterpret_cast<int *>(hb + nW);
                for (int w = tid; w < nW; w += nthr)
                    hb[w] = 0u;
                bar();
                for (int p = tid; p < ne; p += nthr)
                {
                    const int k = bkey[p];
                    const int b = (k - cmin) >> sh;
                    const int lo = b ? start[b - 1] : 0, hi = start[b];
                    int less = 0, eqb = 0;
                    for (int q = lo; q < hi; ++q)
                    {
                        const int kk = bkey[q];
                        less += kk < k;
                        eqb += (kk == k) & (q < p);
                    }
                    const int j = lo + less + eqb;
                    list[p] = (unsigned short)j;
                    if (eqb == 0)
                        atomicOr(&hb[j >> 5], 1u << (j & 31));
                }
                bar();
                int hcarry = 0;
                for (int w0 = 0; w0 < nW; w0 += nthr)
                {
                    const int w = w0 + tid;
                    const int c = (w < nW) ? __popc(hb[w]) : 0;
                    int tot, ex;
                    if (nthr == 32)
                    {
                        int incl = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1)
                        {
                            const int t = __shfl_up_sync(kFull, incl, o);
                            if (lane >= o)
                                incl += t;
                        }
                        ex = incl - c;
                        tot = __shfl_sync(kFull, incl, 31);
                    }
                    else
                        ex = block_excl_scan(c, warp_tot, &tot);
                    if (w < nW)
                        hpre[w] = hcarry + ex;
                    hcarry += tot;
                }
                bar();
                // heads write their entry of C, then the repeats are added onto it
                for (int pass = 0; pass < 2; ++pass)
                {
                    for (int p = tid; p < ne; p += nthr)
                    {
                        const int j = list[p];
                        const unsigned word = hb[j >> 5];
                        const bool head = (word >> (j & 31)) & 1u;
                        if (head != (pass == 0))
                            continue;
                        const int dest = out + hpre[j >> 5] + __popc(word & ((2u << (j & 31)) - 1u)) - 1;
                        const int sl = idx[p];
                        if (head)
                        {
                            Cc[dest] = bkey[p];
                            Cv[dest] = vals[sl];
                        }
                        else
                            atomicAdd(&Cv[dest], vals[sl]);
                        keys[sl] = -1; // leave the table clean for the next row
                        if (!solo)
                            vals[sl] = T(0);
                    }
                    bar();
                }
                for (int w = tid; w < 2 * nW; w += nthr) // the scratch lies in the table's value array: back to zero
                    hb[w] = 0u;
            }
            else
            for (int p = tid; p < n; p += nthr)
            {
                const int sl = idx[p];
                const int k = bkey[p];
                const int b = (k - cmin) >> sh;
                const int lo = b ? start[b - 1] : 0, hi = start[b];
                int rank = lo;
                for (int q = lo; q < hi; ++q)
                    rank += bkey[q] < k;
                Cc[out + rank] = k;
                Cv[out + rank] = vals[sl];
                keys[sl] = -1; // leave the table clean for the next row
                if (!solo)
                    vals[sl] = T(0);
            }
        }
        else
        {
            // heavily clustered columns: pull the entries into registers, rebuild them as dense
            // arrays at the front of the table and bitonic-sort there (at most 10 per thread)
            int rk[10];
            T rv[10];
#pragma unroll
            for (int j = 0; j < 10; ++j)
            {
                const int i = tid + j * nthr;
                rk[j] = INT_MAX;
                rv[j] = T(0);
                if (i < n)
                {
                    const int sl = uniq ? i : (int)list[i];
                    rk[j] = keys[sl];
                    rv[j] = vals[sl];
                }
            }
            bar();
            int P = 2;
            while (P < n)
                P <<= 1;
#pragma unroll
            for (int j = 0; j < 10; ++j)
            {
                const int i = tid + j * nthr;
                if (i < n)
                {
                    keys[i] = rk[j];
                    vals[i] = rv[j];
                }
            }
            for (int i = n + tid; i < P; i += nthr)
                keys[i] = INT_MAX;
            bar();
            bitonic_sort_kv(keys, vals, P, tid, nthr, bar);
            for (int i = tid; i < n; i += nthr)
            {
                Cc[out + i] = keys[i];
                Cv[out + i] = vals[i];
            }
            bar();
            for (int i = tid; i < (1 << lS); i += nthr) // full clear of the row's table: slots outside [0, P) may still be set
            {
                keys[i] = -1;
                vals[i] = T(0);
            }
        }
        bar();
    }
    flush_probes(probes, np);
}

// ---- tiny rows: one thread per row ------------------------------------------------------
// n <= NB_TINY_MAX (24) entries and <= NB_TINY_PRODUCTS products.  The thread appends (column,
// value) pairs to its own shared-memory column (bank = thread, conflict-free), merging
// repeats by a linear scan, insertion-sorts the <= 24 pairs and writes them out.  ~10x fewer
// warp instructions per row than the 8-lane hash kernel (profiles/r1c_tiny_rows.md).
template <typename T>
__global__ void __launch_bounds__(kTinyRowThreads)
    k_num_tiny(RowList list, const int *__restrict__ Ap, const int *__restrict__ Ac,
               const T *__restrict__ Av, const int *__restrict__ Bp, const int *__restrict__ Bc,
               const T *__restrict__ Bv, const int *__restrict__ Cp, int *__restrict__ Cc, T *__restrict__ Cv)
{
    const int *__restrict__ rows = list.begin();
    const int nrows = list.size();
    extern __shared__ __align__(16) unsigned char sm_raw[]; // vals[NB_TINY_MAX][threads] | keys[NB_TINY_MAX][threads]
    T *vals = reinterpret_cast<T *>(sm_raw);
    int *keys = reinterpret_cast<int *>(vals + NB_TINY_MAX * kTinyRowThreads);
    const int t = threadIdx.x;
    for (int r = blockIdx.x * kTinyRowThreads + t; r < nrows; r += gridDim.x * kTinyRowThreads)
    {
        const int row = rows[r];
        const int out = __ldg(&Cp[row]);
        const int cap = __ldg(&Cp[row + 1]) - out; // <= NB_TINY_MAX by the bin's definition
        int n = 0;
        const int s = __ldg(&Ap[row]), e = __ldg(&Ap[row + 1]);
        // ONE loop over the products of the row (the nested form -- for every nonzero of A, for every
        // entry of its B row -- reconverges the warp at the end of every B row, so a warp paid the
        // longest B row of every round: 5.6 of 32 lanes active on the power-law input, r2n)
        int j = s, q = 0, qe = 0;
        T a = T(0);
        while (true)
        {
            while (q == qe && j < e) // next nonzero of A (skipping empty B rows)
            {
                const int k = __ldg(&Ac[j]);
                a = __ldg(&Av[j]);
                q = __ldg(&Bp[k]), qe = __ldg(&Bp[k + 1]);
                ++j;
            }
            if (q == qe)
                break;
            const int c = __ldg(&Bc[q]);
            const T x = a * __ldg(&Bv[q]);
            ++q;
            int p = 0;
            while (p < n && keys[p * kTinyRowThreads + t] != c)
                ++p;
            if (p < n)
                vals[p * kTinyRowThreads + t] += x;
            else if (n < cap)
            {
                keys[n * kTinyRowThreads + t] = c;
                vals[n * kTinyRowThreads + t] = x;
                ++n;
            }
        }
        for (int i = 1; i < n; ++i) // insertion sort by column
        {
            const int kk = keys[i * kTinyRowThreads + t];
            const T vv = vals[i * kTinyRowThreads + t];
            int p = i;
            while (p > 0 && keys[(p - 1) * kTinyRowThreads + t] > kk)
            {
                keys[p * kTinyRowThreads + t] = keys[(p - 1) * kTinyRowThreads + t];
                vals[p * kTinyRowThreads + t] = vals[(p - 1) * kTinyRowThreads + t];
                --p;
            }
            keys[p * kTinyRowThreads + t] = kk;
            vals[p * kTinyRowThreads + t] = vv;
        }
        for (int i = 0; i < n; ++i)
        {
            Cc[out + i] = keys[i * kTinyRowThreads + t];
            Cv[out + i] = vals[i * kTinyRowThreads + t];
        }
    }
}

} // namespace mhb
