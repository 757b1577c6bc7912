// Sharded exact sequential sum, normaliser and estimate: every rank works on ITS slice of the particle order and the
// ranks exchange only what the sequential dependency needs -- over NVLink, by peer stores into one CUDA-IPC-mapped
// "exchange block" per rank, ordered by flag barriers in the same block (no collective call on the data path).
//
// The reference sums weights left to right in double (particle_filter.cpp:93-99, :126-133); mcl_kernels.cuh (K1) explains
// how that sum is reproduced exactly from composable "add D[parity]" maps.  Sharded, per stage (resample / normalise):
//   S1, S2  rank r: plain chunk sums of its slice, scan of its tiles, slice total  -> totals[r] on every rank   (8 B)
//   -- barrier A --
//   S3      rank r: binade guess from (sum of the lower ranks' totals + its own scan), level-1 maps of its chunks
//                   -> q0/q1/ebias[k] on every rank (20 B per 128 particles); the raw weights of the few chunks whose
//                   running sum may cross a binade edge -> side buffer on every rank (1 KB per such chunk)
//   S4      rank r: level-2 maps of its groups -> g0/g1/gebias[j] on every rank (20 B per 8192 particles)
//   -- barrier B --
//   S5      every rank: the walk over all groups (redundant, and local: everything it can need was pushed)
//           -> exact total, exact entry sum of every group
//   resample only: S6/S7 materialise the exact running sum for the groups THIS rank's children draw from (their
//           weights are read from the owning rank over NVLink), then the search over its own children.
// Slices start at multiples of kL1 * kL2 = 8192 particles, so chunks, groups and the estimate's 4096-particle partials
// never straddle ranks; results do not depend on the partition (the maps compose associatively; the walk verifies
// every assumption against exact values).
#pragma once
#include "mcl_kernels.cuh"

namespace mcl {

constexpr int kSliceAlign = kL1 * kL2;          // 8192
constexpr int kFbSlots = 256;                   // raw-chunk side buffer: slots per source rank
constexpr int kXFlagSlots = 16;

// Byte offsets inside a rank's exchange block (identical on every rank).
struct XLayout {
    size_t w[2];            // weights, double-buffered: [N] doubles each (only the owner's slice is meaningful)
    size_t q0, q1, eb;      // level-1 maps: [n1] int64, [n1] int64, [n1] int32
    size_t g0, g1, ge;      // level-2 maps: [n2]
    size_t tot;             // [kMaxPeers] doubles: approximate slice totals
    size_t est;             // [est_count] double4 + [est_count] double (sum w^2)
    size_t est_w2;
    size_t fbraw;           // [kMaxPeers][kFbSlots][kL1] doubles
    size_t flags;           // [kXFlagSlots][kMaxPeers] ints
    size_t err;             // int: barrier timeout
    size_t bytes;
};

struct XPeers {
    int world, rank;
    unsigned char* base[kMaxPeers];     // exchange block of every rank (own entry = local block)
    long long lo[kMaxPeers + 1];        // slice boundaries: rank r owns [lo[r], lo[r+1])
};

__host__ __device__ inline int xowner(const XPeers& xp, long long elem)
{
    int r = 0;
    while (r + 1 < xp.world && elem >= xp.lo[r + 1]) ++r;
    return r;
}

// ---- flag barrier -------------------------------------------------------------------------------------------------
// signal: after everything this rank enqueued before it (kernel boundary), tell every rank "rank `rank` reached `epoch`
// on slot `slot`".  wait: spin until every rank has.  Epochs only grow, so a fast rank can never be mistaken.
__global__ void xsignal_kernel(const XPeers xp, size_t flags_off, int slot, int epoch)
{
    const int r = threadIdx.x;
    if (r < xp.world) {
        __threadfence_system();
        volatile int* f = reinterpret_cast<volatile int*>(xp.base[r] + flags_off) + slot * kMaxPeers + xp.rank;
        *f = epoch;
        __threadfence_system();
    }
}

// signal + wait in one launch (lane r tells rank r, then waits for rank r)
__global__ void xbarrier_kernel(const XPeers xp, size_t flags_off, size_t err_off, int slot, int epoch)
{
    const int r = threadIdx.x;
    if (r < xp.world) {
        __threadfence_system();
        volatile int* out = reinterpret_cast<volatile int*>(xp.base[r] + flags_off) + slot * kMaxPeers + xp.rank;
        *out = epoch;
        __threadfence_system();
        unsigned char* base = xp.base[xp.rank];
        volatile int* f = reinterpret_cast<volatile int*>(base + flags_off) + slot * kMaxPeers + r;
        const long long t0 = clock64();
        while (*f < epoch) {
            if (clock64() - t0 > 20000000000ll) {       // ~10 s: a peer died; do not hang the GPU
                *reinterpret_cast<volatile int*>(base + err_off) = 1;
                break;
            }
        }
        __threadfence_system();
    }
}

__global__ void xwait_kernel(unsigned char* base, size_t flags_off, size_t err_off, int slot, int epoch, int world)
{
    const int r = threadIdx.x;
    if (r < world) {
        volatile int* f = reinterpret_cast<volatile int*>(base + flags_off) + slot * kMaxPeers + r;
        const long long t0 = clock64();
        while (*f < epoch) {
            if (clock64() - t0 > 20000000000ll) {       // ~10 s: a peer died; do not hang the GPU
                *reinterpret_cast<volatile int*>(base + err_off) = 1;
                break;
            }
        }
        __threadfence_system();
    }
}

// ---- S1: plain per-chunk sums and per-tile totals of the slice's tiles [tile_lo, tile_lo + gridDim.x) -----------------
__global__ void __launch_bounds__(128) xseq_chunk_sums_kernel(const double* w, long long n, long long n1, long long tile_lo,
                                                              double* sums, double* tile_sums)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    const long long tidx = tile_lo + blockIdx.x;
    const long long chunk0 = tidx * kSeqTileChunks;
    seq_stage(w, n, chunk0 * kL1, tile);
    __syncthreads();
    if (threadIdx.x < kSeqTileChunks) {
        double s = 0.0;
        if (chunk0 + threadIdx.x < n1) {
            const double* row = tile + threadIdx.x * kSeqRowPitch;
#pragma unroll 8
            for (int i = 0; i < kL1; ++i) s += row[i];
            sums[chunk0 + threadIdx.x] = s;
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (threadIdx.x == 0) tile_sums[tidx] = s;
    }
}

// ---- S2: exclusive scan of the slice's tile totals (slice-relative) + the slice total to every rank ----------------------
// (also re-arms fb_count, the side-buffer slot counter S3 is about to use)
__global__ void __launch_bounds__(1024) xseq_tile_scan_kernel(const double* tile_sums, long long tile_lo, long long tile_hi,
                                                              double* tile_excl, const XPeers xp, size_t tot_off, int* fb_count)
{
    __shared__ double warp_tot[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) { carry_s = 0.0; *fb_count = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = tile_lo; base < tile_hi; base += 1024) {
        const long long k = base + threadIdx.x;
        const double v = k < tile_hi ? tile_sums[k] : 0.0;
        double inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            double t = warp_tot[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, t, off);
                if (lane >= off) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const double incl = carry_s + (wid ? warp_tot[wid - 1] : 0.0) + inc;
        if (k < tile_hi) tile_excl[k] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if ((int)threadIdx.x < xp.world)
        reinterpret_cast<double*>(xp.base[threadIdx.x] + tot_off)[xp.rank] = carry_s;
}

// ---- S3: level-1 maps of the slice's chunks, pushed to every rank ----------------------------------------------------------
struct XSeqOut { size_t q0, q1, eb, fbraw; };
__global__ void __launch_bounds__(128) xseq_chunk_maps_kernel(const double* w, long long n, long long n1, long long tile_lo,
                                                              const double* sums, const double* tile_excl,
                                                              const double* totals, int* fb_count, const XPeers xp,
                                                              const XSeqOut o)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    __shared__ int s_slot[kSeqTileChunks];
    const long long tidx = tile_lo + blockIdx.x;
    const long long chunk0 = tidx * kSeqTileChunks;
    seq_stage(w, n, chunk0 * kL1, tile);
    __syncthreads();
    const long long k = chunk0 + threadIdx.x;
    if (threadIdx.x < kSeqTileChunks) {                 // warp 0, all 32 lanes (shuffles below)
        double below = 0.0;                             // approximate sum of the lower ranks' slices (fixed order)
        for (int r = 0; r < xp.rank; ++r) below += totals[r];
        const double v = k < n1 ? sums[k] : 0.0;
        double inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)threadIdx.x >= off) inc += t;
        }
        const double incl = below + tile_excl[tidx] + inc, excl = incl - v;
        const double lo = excl * (1.0 - 1e-6), hi = incl * (1.0 + 1e-6);
        const int elo = dbl_exp(lo), ehi = dbl_exp(hi);
        int e = (lo > 0.0 && elo == ehi && elo > 60 && elo < 1900) ? elo : 0;
        long long f0 = 0, f1 = 0;
        if (k < n1 && e != 0) {
            const double base0 = __longlong_as_double((long long)e << 52);
            const double ulp = __longlong_as_double((long long)(e - 52) << 52);
            const double base1 = base0 + ulp;
            const double top = base0 + base0;
            const double* row = tile + threadIdx.x * kSeqRowPitch;
            double c0 = base0, c1 = base1;
#pragma unroll 8
            for (int i = 0; i < kL1; ++i) {
                const double vv = row[i];
                c0 = __dadd_rn(c0, vv);
                c1 = __dadd_rn(c1, vv);
            }
            if (c0 < top && c1 < top && c0 >= base0 && c1 >= base1) {
                f0 = __double_as_longlong(c0) - __double_as_longlong(base0);   // same binade: bit patterns count ulps
                f1 = __double_as_longlong(c1) - __double_as_longlong(base1);
            } else {
                e = 0;
            }
        }
        // chunks that will be added element by element: their raw weights go to every rank's side buffer
        int slot = -1;
        if (k < n1 && e == 0 && xp.world > 1) {
            slot = atomicAdd(fb_count, 1);
            if (slot >= kFbSlots) slot = -1;
        }
        s_slot[threadIdx.x] = slot;
        if (k < n1) {
            if (e == 0) f0 = slot;
            for (int r = 0; r < xp.world; ++r) {
                reinterpret_cast<long long*>(xp.base[r] + o.q0)[k] = f0;
                reinterpret_cast<long long*>(xp.base[r] + o.q1)[k] = f1;
                reinterpret_cast<int*>(xp.base[r] + o.eb)[k] = e;
            }
        }
        __syncwarp();
        if (xp.world > 1) {
            for (int c = 0; c < kSeqTileChunks; ++c) {
                const int sl = s_slot[c];
                if (sl < 0) continue;
                const double* row = tile + c * kSeqRowPitch;
                for (int r = 0; r < xp.world; ++r) {
                    double* dst = reinterpret_cast<double*>(xp.base[r] + o.fbraw) + ((size_t)xp.rank * kFbSlots + sl) * kL1;
                    for (int i = threadIdx.x; i < kL1; i += 32) dst[i] = row[i];
                }
            }
        }
    }
}

// ---- S4: level-2 maps of the slice's groups [group_lo, group_hi), pushed to every rank ---------------------------------------
// One warp per group: every lane composes two consecutive chunk maps, a shuffle tree composes the 32 results in order
// (composition is associative, not commutative).
__global__ void __launch_bounds__(128) xseq_group_maps_kernel(const int* ebias, const long long* q0, const long long* q1,
                                                              long long n1, long long group_lo, long long group_hi,
                                                              const XPeers xp, size_t g0_off, size_t g1_off, size_t ge_off)
{
    const int lane = threadIdx.x & 31;
    const long long j = group_lo + blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= group_hi) return;
    const long long first = j * kL2;
    const long long last = first + kL2 < n1 ? first + kL2 : n1;
    const long long ka = first + 2 * lane, kb = ka + 1;
    const int e = ebias[first];
    long long a0 = 0, a1 = 0;                               // identity map for the chunks past the end
    bool same = true;
    if (ka < last) { same = ebias[ka] == e; a0 = q0[ka]; a1 = q1[ka]; }
    if (kb < last) {
        same = same && ebias[kb] == e;
        long long r0, r1;
        seq_compose(a0, a1, q0[kb], q1[kb], r0, r1);
        a0 = r0; a1 = r1;
    }
    const bool ok = e != 0 && __all_sync(0xffffffffu, same);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long b0 = __shfl_down_sync(0xffffffffu, a0, off), b1 = __shfl_down_sync(0xffffffffu, a1, off);
        if ((lane & (2 * off - 1)) == 0) {
            long long r0, r1;
            seq_compose(a0, a1, b0, b1, r0, r1);
            a0 = r0; a1 = r1;
        }
    }
    a0 = __shfl_sync(0xffffffffu, a0, 0); a1 = __shfl_sync(0xffffffffu, a1, 0);        // lane 0 holds the group's map
    if (lane < xp.world) {
        reinterpret_cast<long long*>(xp.base[lane] + g0_off)[j] = a0;
        reinterpret_cast<long long*>(xp.base[lane] + g1_off)[j] = a1;
        reinterpret_cast<int*>(xp.base[lane] + ge_off)[j] = ok ? e : 0;
    }
}

// Advances the exact running sum c over the longest valid prefix of up to cnt consecutive maps (lane l holds map l:
// exponent e_l and (m0_l, m1_l)).  A map is valid while it assumes c's binade and the sum stays inside that binade.
// Returns L in [0, cnt]: maps 0 .. L-1 were applied, c = the sum after them, and for every lane l <= L cin_l = the exact
// sum at the entry of map l (so lane L, the first map that did not apply, knows where it starts).
__device__ __forceinline__ int seq_apply_prefix(double& c, int cnt, int e_l, long long m0_l, long long m1_l, double& cin_l)
{
    const int lane = threadIdx.x & 31;
    const int e = dbl_exp(c);
    const bool mine_ok = lane < cnt && e_l == e && e_l != 0;
    const unsigned okmask = __ballot_sync(0xffffffffu, mine_ok);
    const int l0 = okmask == 0xffffffffu ? 32 : __ffs(~okmask) - 1;          // leading maps in c's binade
    const long long bits = __double_as_longlong(c);
    cin_l = c;
    if (l0 == 0) return 0;
    long long p0 = lane < l0 ? m0_l : 0, p1 = lane < l0 ? m1_l : 0;          // inclusive prefix composition
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long q0 = __shfl_up_sync(0xffffffffu, p0, off), q1 = __shfl_up_sync(0xffffffffu, p1, off);
        if (lane >= off) {
            long long r0, r1;
            seq_compose(q0, q1, p0, p1, r0, r1);
            p0 = r0; p1 = r1;
        }
    }
    const long long incl = bits + ((bits & 1) ? p1 : p0);
    // weights are non-negative, so the sums grow with the lane: the maps that leave the binade form a suffix
    const bool stay = lane >= l0 || ((((incl >> 52) & 0x7ff) == e) && incl >= bits);
    const unsigned smask = __ballot_sync(0xffffffffu, stay);
    const int l1 = smask == 0xffffffffu ? 32 : __ffs(~smask) - 1;
    const int L = l1 < l0 ? l1 : l0;
    long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = bits;
    cin_l = __longlong_as_double(excl);
    if (L > 0) c = __longlong_as_double(__shfl_sync(0xffffffffu, incl, L - 1));
    return L;
}

// ---- S5: the walk (one warp; the CTA stages the group maps).  Runs of maps that stay inside one binade advance up to
// 32 at a time (seq_apply_prefix); the map that does not apply is opened into the next level down -- a group into its
// chunks, a chunk into its raw weights (real sequential double adds) -- and the walk resumes right behind it.  The raw
// weights of such a chunk come from the rank's own slice, from the side buffer its owner filled, or from the owner's
// weights over NVLink.  Produces the exact entry sum of every group (cin2), of every chunk of the groups that had to
// be opened (cin1, flagged in opened[j]), the exact total and the number of chunks added element by element.
constexpr int kWalkOpen = 16;        // groups whose chunk maps are staged ahead of the walk
constexpr int kWalkFb = 16;          // chunks whose raw weights are staged ahead of the walk
__host__ __device__ inline size_t xseq_walk_smem(long long n2)
{
    return (size_t)n2 * 20 + (size_t)kWalkOpen * kL2 * 20 + (size_t)kWalkFb * kL1 * 8;
}

// Where the raw weights of chunk k live for this rank: its own slice, the side buffer its owner filled (S3 knew the
// chunk would be added element by element), or the owner's weights over NVLink.  bounded: not zero padded past n.
__device__ __forceinline__ const double* xseq_raw_chunk(const XPeers& xp, size_t w_off, size_t fbraw_off, long long k, int e,
                                                        long long f0, bool& bounded)
{
    const long long efirst = k * kL1;
    const int owner = xowner(xp, efirst);
    bounded = true;
    if (owner == xp.rank) return reinterpret_cast<const double*>(xp.base[xp.rank] + w_off) + efirst;
    if (e == 0 && f0 >= 0) {
        bounded = false;
        return reinterpret_cast<const double*>(xp.base[xp.rank] + fbraw_off) + ((size_t)owner * kFbSlots + (size_t)f0) * kL1;
    }
    return reinterpret_cast<const double*>(xp.base[owner] + w_off) + efirst;
}

__global__ void __launch_bounds__(1024) xseq_walk_kernel(long long n, long long n1, long long n2, const int* ebias,
                                                       const long long* q0, const long long* q1, const int* gebias,
                                                       const long long* g0, const long long* g1, double* cin2, double* cin1,
                                                       int* opened, double* total, long long* fallback_chunks, int staged,
                                                       const XPeers xp, size_t w_off, size_t fbraw_off, double* zero_after)
{
    __shared__ double fb_chunk[kL1];
    __shared__ long long sc0[kL2], sc1[kL2];      // level-1 maps of the group being opened
    __shared__ int sce[kL2];
    __shared__ int s_nopen, s_nfb;
    __shared__ long long s_open_j[kWalkOpen], s_fb_k[kWalkFb];
    __shared__ int s_fb_e[kWalkFb];
    __shared__ long long s_fb_f0[kWalkFb];
    extern __shared__ __align__(16) unsigned char walk_smem[];
    long long* sg0 = reinterpret_cast<long long*>(walk_smem);
    long long* sg1 = sg0 + (staged ? n2 : 0);
    long long* so0 = sg1 + (staged ? n2 : 0);                 // [kWalkOpen][kL2]
    long long* so1 = so0 + (staged ? kWalkOpen * kL2 : 0);
    double* sfb = reinterpret_cast<double*>(so1 + (staged ? kWalkOpen * kL2 : 0));      // [kWalkFb][kL1]
    int* sge = reinterpret_cast<int*>(sfb + (staged ? kWalkFb * kL1 : 0));
    int* soe = sge + (staged ? n2 : 0);                       // [kWalkOpen][kL2]
    int nopen = 0, nfb = 0;
    if (staged) {
        // Everything the walk will read, staged by the whole CTA so that its single warp never waits for global memory:
        // the group maps; the chunk maps of the groups that do not apply as a whole (S4 marked them); the raw weights
        // of their chunks that are added element by element (S3 marked those).
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        if (threadIdx.x == 0) { s_nopen = 0; s_nfb = 0; }
        for (long long j = threadIdx.x; j < n2; j += blockDim.x) { sg0[j] = g0[j]; sg1[j] = g1[j]; sge[j] = gebias[j]; }
        __syncthreads();
        for (long long j = wrp; j < n2; j += nwarps) {
            if (sge[j] != 0) continue;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&s_nopen, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot >= kWalkOpen) continue;
            if (lane == 0) s_open_j[slot] = j;
            const long long kfirst = j * kL2;
            for (int i = lane; i < kL2; i += 32) {
                const bool in = kfirst + i < n1;
                so0[slot * kL2 + i] = in ? q0[kfirst + i] : 0;
                so1[slot * kL2 + i] = in ? q1[kfirst + i] : 0;
                soe[slot * kL2 + i] = in ? ebias[kfirst + i] : -1;
            }
        }
        __syncthreads();
        nopen = min(s_nopen, kWalkOpen);
        for (int t = threadIdx.x; t < nopen * kL2; t += blockDim.x) {
            if (soe[t] == 0) {
                const int fs = atomicAdd(&s_nfb, 1);
                if (fs < kWalkFb) {
                    s_fb_k[fs] = s_open_j[t / kL2] * kL2 + (t % kL2);
                    s_fb_e[fs] = 0; s_fb_f0[fs] = so0[t];
                }
            }
        }
        __syncthreads();
        nfb = min(s_nfb, kWalkFb);
        for (int fs = wrp; fs < nfb; fs += nwarps) {
            bool bounded;
            const long long k = s_fb_k[fs];
            const double* src = xseq_raw_chunk(xp, w_off, fbraw_off, k, s_fb_e[fs], s_fb_f0[fs], bounded);
            for (int i = lane; i < kL1; i += 32) sfb[fs * kL1 + i] = (!bounded || k * kL1 + i < n) ? src[i] : 0.0;
        }
        __syncthreads();
    }
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double c = 0.0;
    long long fallbacks = 0;
    long long j = 0;
    while (j < n2) {
        const long long jl = j + lane;
        const int cnt = (int)(n2 - j < 32 ? n2 - j : 32);
        const int ge_l = jl < n2 ? (staged ? sge[jl] : gebias[jl]) : 0;
        const long long g0_l = jl < n2 ? (staged ? sg0[jl] : g0[jl]) : 0, g1_l = jl < n2 ? (staged ? sg1[jl] : g1[jl]) : 0;
        double cin_l;
        const int L = seq_apply_prefix(c, cnt, ge_l, g0_l, g1_l, cin_l);
        if (lane < L) { cin2[jl] = cin_l; opened[jl] = 0; }
        j += L;
        if (L == cnt) continue;
        // group j does not apply as a whole: open it
        if (lane == 0) { cin2[j] = c; opened[j] = 1; }
        const long long kfirst = j * kL2;
        const int kcnt = (int)((kfirst + kL2 < n1 ? kfirst + kL2 : n1) - kfirst);
        const unsigned hit = __ballot_sync(0xffffffffu, lane < nopen && s_open_j[lane < nopen ? lane : 0] == j);
        if (hit) {
            const int slot = __ffs(hit) - 1;
            for (int i = lane; i < kL2; i += 32) {
                const bool in = i < kcnt;
                sce[i] = in ? soe[slot * kL2 + i] : 0; sc0[i] = in ? so0[slot * kL2 + i] : 0; sc1[i] = in ? so1[slot * kL2 + i] : 0;
            }
        } else {
            for (int i = lane; i < kL2; i += 32) {
                const bool in = i < kcnt;
                sce[i] = in ? ebias[kfirst + i] : 0; sc0[i] = in ? q0[kfirst + i] : 0; sc1[i] = in ? q1[kfirst + i] : 0;
            }
        }
        __syncwarp();
        int kd = 0;
        while (kd < kcnt) {
            const int kb = kcnt - kd < 32 ? kcnt - kd : 32;
            const int ki = kd + lane;
            const int e_l = lane < kb ? sce[ki] : 0;
            const long long f0_l = lane < kb ? sc0[ki] : 0, f1_l = lane < kb ? sc1[ki] : 0;
            const int K = seq_apply_prefix(c, kb, e_l, f0_l, f1_l, cin_l);
            if (lane < K) cin1[kfirst + ki] = cin_l;
            kd += K;
            if (K == kb) continue;
            // chunk kfirst + kd is added element by element
            const long long k = kfirst + kd;
            if (lane == 0) cin1[k] = c;
            ++fallbacks;
            const unsigned fhit = __ballot_sync(0xffffffffu, lane < nfb && s_fb_k[lane < nfb ? lane : 0] == k);
            if (fhit) {
                const double* src = sfb + (__ffs(fhit) - 1) * kL1;
#pragma unroll
                for (int q = 0; q < kL1 / 32; ++q) fb_chunk[q * 32 + lane] = src[q * 32 + lane];
            } else {
                bool bounded;
                const double* src = xseq_raw_chunk(xp, w_off, fbraw_off, k, sce[kd], sc0[kd], bounded);
                const long long efirst = k * kL1;
                double v_l[kL1 / 32];
#pragma unroll
                for (int q = 0; q < kL1 / 32; ++q) {                      // all loads in flight before the adds
                    const long long gi = efirst + q * 32 + lane;
                    v_l[q] = (!bounded || gi < n) ? src[q * 32 + lane] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < kL1 / 32; ++q) fb_chunk[q * 32 + lane] = v_l[q];
            }
            __syncwarp();
#pragma unroll 16
            for (int i = 0; i < kL1; ++i) c = __dadd_rn(c, fb_chunk[i]);
            __syncwarp();
            kd += 1;
        }
        __syncwarp();
        j += 1;
    }
    if (lane == 0) {
        *total = c;
        *fallback_chunks = fallbacks;
        if (zero_after) *zero_after = 0.0;          // (the normaliser's sum of squares, accumulated by the divide that follows)
    }
}

// ---- resampling: which groups do this rank's children draw from? ---------------------------------------------------------
// Child m compares U_m = r + m/N with the running sum; its parent lies in the last group whose ENTRY sum is below U_m.
// range[0], range[1] = first and last such group over the rank's children [lo, hi).
// (also re-arms the search's overrun counter)
__global__ void xresample_range_kernel(const double* cin2, long long n2, long long children, double r, long long lo,
                                       long long hi, long long* range, unsigned long long* overruns)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *overruns = 0ull;
    const double m_inv = __ddiv_rn(1.0, (double)children);
    auto last_below = [&](double u) {
        long long a = 0, b = n2;                 // first j with !(cin2[j] < u)
        while (a < b) {
            const long long mid = (a + b) >> 1;
            if (cin2[mid] < u) a = mid + 1; else b = mid;
        }
        return a > 0 ? a - 1 : 0;
    };
    if (hi <= lo) { range[0] = 0; range[1] = -1; return; }
    range[0] = last_below(__dadd_rn(r, __dmul_rn((double)lo, m_inv)));
    range[1] = last_below(__dadd_rn(r, __dmul_rn((double)(hi - 1), m_inv)));
}

// S6 for the groups in range: entry sums of their chunks (opened groups already have theirs from the walk).
__global__ void xseq_group_expand_kernel(const long long* q0, const long long* q1, long long n1, const double* cin2,
                                         const int* opened, double* cin1, const long long* range)
{
    const long long j = range[0] + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j > range[1] || opened[j]) return;
    const long long first = j * kL2;
    const long long last = first + kL2 < n1 ? first + kL2 : n1;
    long long bits = __double_as_longlong(cin2[j]);
    for (long long k = first; k < last; ++k) {
        cin1[k] = __longlong_as_double(bits);
        bits += (bits & 1) ? q1[k] : q0[k];
    }
}

// S7 for the chunks of the groups in range: exact running sum per element; weights come from the owning rank.
__global__ void __launch_bounds__(128) xseq_materialize_kernel(long long n, long long n1, const double* cin1, double* cum,
                                                               const long long* range, const XPeers xp, size_t w_off)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    const long long tile_first = range[0] * kL2 / kSeqTileChunks, tile_last = ((range[1] + 1) * kL2 - 1) / kSeqTileChunks;
    for (long long tidx = tile_first + blockIdx.x; tidx <= tile_last && range[1] >= range[0]; tidx += gridDim.x) {
        const long long chunk0 = tidx * kSeqTileChunks;
        const long long efirst = chunk0 * kL1;
        if (efirst >= n) break;
        const double* w = reinterpret_cast<const double*>(xp.base[xowner(xp, efirst)] + w_off);     // a tile never straddles ranks
        __syncthreads();
        seq_stage(w, n, efirst, tile);
        __syncthreads();
        const long long k = chunk0 + threadIdx.x;
        if (threadIdx.x < kSeqTileChunks && k < n1) {
            double* row = tile + threadIdx.x * kSeqRowPitch;
            double c = cin1[k];
#pragma unroll 8
            for (int i = 0; i < kL1; ++i) {
                c = __dadd_rn(c, row[i]);
                row[i] = c;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kSeqTileChunks * kL1; i += blockDim.x) {
            const long long g = efirst + i;
            if (g < n) cum[g] = tile[(i / kL1) * kSeqRowPitch + (i % kL1)];
        }
    }
}

// Systematic search over the materialised range (same indices as resample_search_kernel over the whole array).
// children: the number of draws the offsets U_m = r + m / children are spread over (N for the filter's resampling).
__global__ void xresample_search_kernel(const double* cum, long long n, long long children, double r, long long lo,
                                        long long hi, const long long* range, int32_t* idx, unsigned long long* overruns)
{
    const double m_inv = __ddiv_rn(1.0, (double)children);
    const long long first = range[0] * (long long)kSliceAlign;
    const long long lim = (range[1] + 1) * (long long)kSliceAlign < n ? (range[1] + 1) * (long long)kSliceAlign : n;
    const long long runs = (hi - lo + kSearchRun - 1) / kSearchRun;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < runs; t += (long long)gridDim.x * blockDim.x) {
        const long long m0 = lo + t * kSearchRun;
        const long long m1 = m0 + kSearchRun < hi ? m0 + kSearchRun : hi;
        long long a = first;
        for (long long m = m0; m < m1; ++m) {
            const double u = __dadd_rn(r, __dmul_rn((double)m, m_inv));
            int steps = 0;
            if (m != m0)
                while (a < lim && u > cum[a] && steps < 32) { ++a; ++steps; }
            if (m == m0 || steps == 32) {
                long long b = lim;
                while (a < b) {
                    const long long mid = (a + b) >> 1;
                    if (u > cum[mid]) a = mid + 1; else b = mid;
                }
            }
            long long out = a;
            if (out >= n) { out = n - 1; atomicAdd(overruns, 1ull); }
            idx[m] = (int32_t)out;
        }
    }
}

// ---- normaliser and estimate over the slice -----------------------------------------------------------------------------------
// particle_filter.cpp:136-138: w /= wSum (IEEE double division, correctly rounded on both sides).  ess_acc collects the
// slice's sum of squares (a statistic; the estimate's final reduction replaces it with the deterministic global one).
__global__ void xdivide_kernel(double* w, long long lo, long long hi, const double* wsum, double* ess_acc)
{
    const double s = *wsum;
    double sq = 0.0;
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const double q = __ddiv_rn(w[i], s);
        w[i] = q;
        sq += q * q;
    }
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(ess_acc, sq);
}

// Partials of (sum w x, sum w y, sum w sinf, sum w cosf) and of sum w^2 over the slice's 4096-particle chunks
// [chunk_lo, chunk_lo + gridDim.x), pushed to every rank.
__global__ void __launch_bounds__(kEstBlock) xestimate_partial_kernel(const float* x, const float* y, const float* th,
                                                                      const double* w, long long n, long long chunk_lo,
                                                                      const XPeers xp, size_t est_off, size_t w2_off)
{
    __shared__ double4 red[kEstBlock];
    __shared__ double red2[kEstBlock];
    const long long cidx = chunk_lo + blockIdx.x;
    const long long base = cidx * kEstChunk;
    double4 acc = make_double4(0, 0, 0, 0);
    double acc2 = 0.0;
    for (int k = threadIdx.x; k < kEstChunk; k += kEstBlock) {
        const long long i = base + k;
        if (i < n) {
            const double wi = w[i];
            float s, c;
            glibc_sincosf(th[i], &s, &c);
            acc.x = __dadd_rn(acc.x, __dmul_rn(wi, (double)x[i]));
            acc.y = __dadd_rn(acc.y, __dmul_rn(wi, (double)y[i]));
            acc.z = __dadd_rn(acc.z, __dmul_rn(wi, (double)s));
            acc.w = __dadd_rn(acc.w, __dmul_rn(wi, (double)c));
            acc2 = __dadd_rn(acc2, __dmul_rn(wi, wi));
        }
    }
    red[threadIdx.x] = acc;
    red2[threadIdx.x] = acc2;
    __syncthreads();
    for (int off = kEstBlock / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            double4 o = red[threadIdx.x + off], m = red[threadIdx.x];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[threadIdx.x] = m;
            red2[threadIdx.x] += red2[threadIdx.x + off];
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < xp.world) {
        reinterpret_cast<double4*>(xp.base[threadIdx.x] + est_off)[cidx] = red[0];
        reinterpret_cast<double*>(xp.base[threadIdx.x] + w2_off)[cidx] = red2[0];
    }
}

// out4 = (x, y, theta, unused) as floats, ess_acc = sum w^2; single block, fixed order.
__global__ void __launch_bounds__(kEstBlock) xestimate_final_kernel(const double4* partials, const double* w2, int count,
                                                                    float* out4, double* ess_acc)
{
    __shared__ double4 red[kEstBlock];
    __shared__ double red2[kEstBlock];
    double4 acc = make_double4(0, 0, 0, 0);
    double acc2 = 0.0;
    for (int k = threadIdx.x; k < count; k += kEstBlock) {
        const double4 p = partials[k];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        acc2 += w2[k];
    }
    red[threadIdx.x] = acc;
    red2[threadIdx.x] = acc2;
    __syncthreads();
    for (int off = kEstBlock / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            double4 o = red[threadIdx.x + off], m = red[threadIdx.x];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[threadIdx.x] = m;
            red2[threadIdx.x] += red2[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out4[0] = (float)red[0].x;
        out4[1] = (float)red[0].y;
        out4[2] = (float)atan2(red[0].z, red[0].w);
        out4[3] = 0.0f;
        *ess_acc = red2[0];
    }
}

// Order-independent, position-sensitive digest of the rank's slice: sums (mod 2^64) of mixed (global index, bit pattern)
// pairs of the resample indices, the half-unit scores, the weights and the poses.  Slices partition the cloud, so the
// sum of the ranks' digests does not depend on the GPU count iff the results do not.
__device__ __forceinline__ unsigned long long xmix(unsigned long long i, unsigned long long v)
{
    unsigned long long z = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull ^ v;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void xdigest_kernel(const int32_t* idx, const int32_t* score2, const double* w, const float* x, const float* y,
                               const float* th, long long lo, long long hi, unsigned long long* out4)
{
    unsigned long long d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        d0 += xmix((unsigned long long)i, (unsigned long long)(unsigned)idx[i]);
        d1 += xmix((unsigned long long)i, (unsigned long long)(unsigned)score2[i]);
        d2 += xmix((unsigned long long)i, (unsigned long long)__double_as_longlong(w[i]));
        d3 += xmix((unsigned long long)i, ((unsigned long long)__float_as_uint(x[i]) << 32) ^ ((unsigned long long)__float_as_uint(y[i]) << 16) ^
                                              (unsigned long long)__float_as_uint(th[i]));
    }
    for (int off = 16; off > 0; off >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, off); d1 += __shfl_xor_sync(0xffffffffu, d1, off);
        d2 += __shfl_xor_sync(0xffffffffu, d2, off); d3 += __shfl_xor_sync(0xffffffffu, d3, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out4 + 0, d0); atomicAdd(out4 + 1, d1); atomicAdd(out4 + 2, d2); atomicAdd(out4 + 3, d3);
    }
}

}  // namespace mcl
