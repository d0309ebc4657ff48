// binning.cu -- depth ordering, instance emission, stable tile partition, tile ranges (sm_100a).
//
// Replaces the external rasterizer's InclusiveSum + blocking D2H + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs<uint64,uint32> + identifyTileRanges (SURVEY.md section 2c,
// section 8a rows a10-a13; semantics Appendix A.6-A.7) with a different, cheaper decomposition
// that yields the *same* per-tile depth-ordered lists (minus pairs that provably contribute nothing):
//
//   reference:  emit R (tile<<32 | depth) keys  ->  one stable 64-bit LSD sort over R items
//               (6 eight-bit passes at 1080p, ~152 B per instance of HBM traffic)
//   here:       an LSD sort processes the least-significant field first, and the depth field is
//               a property of the *Gaussian*, not of the instance.  So:
//                 1. stable 32-bit radix sort of the P Gaussians by depth bits   (P items)
//                 2. emit instances in that order                                (R items, 8 B each)
//                 3. stable partition by tile id: ceil(bits(Tn)/8) passes        (R items)
//               Stability makes equal-depth ties resolve by Gaussian index, exactly as A.6.
//               At 1080p this is 2 passes over R instead of 6, on 8-byte instead of 12-byte pairs,
//               and R itself is ~40 % smaller because emission keeps only the tiles in which the
//               Gaussian can reach alpha >= 1/255 (exact closed-form test, common.cuh).
//
// Every radix pass is ONE kernel ("onesweep"): the global digit histograms are produced up front
// (by the depth-key kernel for the depth keys, by the emission kernel for the tile ids), each CTA
// takes a dynamic tile ticket, ranks its 4096 items stably with warp match + per-warp counters,
// obtains its global digit offsets by decoupled look-back over the preceding CTAs, reorders the
// items through shared memory and writes them out in coalesced runs.
//
// Nothing here needs R on the host: every kernel reads the instance count from device memory
// and clamps its work to the binning capacity (overflow is flagged, never written past).
// All passes are HBM/L2-bound integer work -- no tensor cores.
#include <cstdlib>

#include "common.cuh"

namespace scgr {

namespace {

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide inclusive scan (blockDim.x a multiple of 32, <= 1024); total of the block in *total
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0u;
        const uint32_t xi = warp_inclusive_scan(x, lane);
        s_warp[lane] = xi;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t base = w > 0 ? s_warp[w - 1] : 0u;
    if (total) *total = s_warp[(blockDim.x >> 5) - 1];
    return inc + base;
}

// Single-pass chained scan (decoupled look-back): offsets[s] = inclusive prefix of
// tiles_touched[order[s]] in depth order; the grand total R goes to status[0].
// state[0] = ticket, state[1 + tile] = flag (2 msb) | value (62 lsb); all zero before the launch.
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;
constexpr unsigned long long SFLAG_PARTIAL = 1ull << 62, SFLAG_INCLUSIVE = 2ull << 62, SFLAG_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_offsets_kernel(const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ order, int P,
                    uint32_t* __restrict__ offsets, int64_t* __restrict__ status, unsigned long long* __restrict__ state,
                    volatile int64_t* status_mapped) {
    pdl_trigger();                  // lets the next kernel of the chain become resident early (common.cuh); it waits for this grid to finish
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(state, 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int ntiles = (P + SCAN_TILE - 1) / SCAN_TILE;
    const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t mine = 0u;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < P ? tiles_touched[order[base + i]] : 0u;
        mine += v[i];
    }
    uint32_t total;
    const uint32_t inc = block_inclusive_scan(mine, s_warp, &total);
    unsigned long long* words = state + 1;
    if (threadIdx.x == 0)
        st_volatile64(words + tile, (tile == 0 ? SFLAG_INCLUSIVE : SFLAG_PARTIAL) | (unsigned long long)total);
    if (threadIdx.x < 32) {
        // warp 0 looks back 32 predecessors at a time
        unsigned long long excl = 0ull;
        int p = (int)tile - 1;
        while (p >= 0) {
            const int q = p - (int)threadIdx.x;
            unsigned long long w = SFLAG_INCLUSIVE;           // virtual tile -1: inclusive prefix 0
            if (q >= 0) w = ld_volatile64(words + q);
            const uint32_t unpublished = __ballot_sync(0xffffffffu, (w & SFLAG_MASK) == 0ull);
            const uint32_t inclusive = __ballot_sync(0xffffffffu, (w & SFLAG_MASK) == SFLAG_INCLUSIVE);
            // usable lanes: those before the first unpublished one, up to and including the first inclusive one
            const int first_unpub = unpublished ? __ffs(unpublished) - 1 : 32;
            const int first_incl = inclusive ? __ffs(inclusive) - 1 : 32;
            const int take = min(first_unpub, first_incl + 1);      // number of lanes consumed
            unsigned long long val = (int)threadIdx.x < take ? (w & ~SFLAG_MASK) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            excl += val;
            if (first_incl < first_unpub) break;              // reached an inclusive word
            p -= take;                                         // retry from the first unpublished word
        }
        if (threadIdx.x == 0) {
            if (tile > 0) st_volatile64(words + tile, SFLAG_INCLUSIVE | (excl + total));
            s_prefix = excl;
            if ((int)tile == ntiles - 1) {
                status[0] = (int64_t)(excl + total);   // R = num_rendered
                status[1] = 0;                         // overflow flag, raised later by the emission kernel
                if (status_mapped) {
                    // zero-copy store into the caller's pinned host word: scgr_forward() spins on it
                    // instead of paying a D2H copy + stream synchronisation
                    status_mapped[0] = (int64_t)(excl + total);
                    __threadfence_system();
                }
            }
        }
    }
    __syncthreads();
    uint32_t run = (uint32_t)s_prefix + inc - mine;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        run += v[i];
        if (base + i < P) offsets[base + i] = run;
    }
}

// ------------------------------------------------------------------------------------------
// onesweep radix pass
// ------------------------------------------------------------------------------------------
constexpr uint32_t FLAG_PARTIAL = 1u << 30;
constexpr uint32_t FLAG_INCLUSIVE = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

__device__ __forceinline__ uint32_t load_count(const int64_t* n_dev, int64_t n_host, int64_t cap) {
    int64_t n = n_dev ? *n_dev : n_host;
    if (n > cap) n = 0;   // overflow: caller re-runs with a larger buffer; do nothing now
    return (uint32_t)n;
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// peers of this lane = lanes holding the same (<= 8-bit) digit.  8 independent ballots instead of
// one MATCH.ANY: VOTE has a short fixed latency and the ballots of all items pipeline.
__device__ __forceinline__ uint32_t peers_by_ballot(const uint32_t d, const bool valid, const int nbits) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        if (b >= nbits) break;      // warp-uniform: the pass's digit has only nbits bits
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// One stable pass.  sweep = this pass's state (layout in common.cuh), zero before the launch except
// for the histogram, which must be complete.  THREADS * ITEMS == RADIX_TILE.
// RANGES (last tile-partition pass): instead of writing the sorted keys, derive the per-tile
// [start, end) ranges (A.7) on the fly -- inside a CTA the reordered items are sorted by full key,
// so every CTA-local run boundary contributes an atomicMin(start) / atomicMax(end).
// MATCH: how a lane finds its peers (lanes of the warp holding the same digit) -- 0: MATCH.ANY, 1: one ballot per digit
// bit, 2: atomic OR of the lane's bit into a per-warp table of 256 masks in shared memory (one ATOMS + one LDS per item
// whatever the digit width; the pass is issue-bound, and the ballots were 40 % of its instructions).
// MINB 0: registers left to the compiler's default for the block size (64 at 512 threads: two CTAs per SM; an explicit
// minimum of 1 lets it take 128 and halves the occupancy: +30 % time).
template <int THREADS, int ITEMS, int MATCH, bool RANGES, int MINB = 0>
__global__ void __launch_bounds__(THREADS, MINB == 0 ? (2048 / THREADS >= 4 ? 2 : 2048 / THREADS / 2 > 0 ? 2048 / THREADS / 2 : 1) : MINB)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const int64_t* __restrict__ n_dev, int64_t n_host, int64_t cap, int shift,
                     uint32_t mask, uint32_t* __restrict__ sweep, uint2* __restrict__ ranges) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    static_assert(THREADS * ITEMS == RADIX_TILE, "tile size is part of the scratch layout");
    constexpr int WARPS = THREADS / 32;
    constexpr int PER_WARP = RADIX_TILE / WARPS;
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t* const s_keys = s_dyn;                          // [RADIX_TILE]
    uint32_t* const s_vals = s_keys + RADIX_TILE;            // [RADIX_TILE]
    uint32_t* const s_start = s_vals + RADIX_TILE;           // [256] CTA-local sorted position of each digit's run
    uint32_t* const s_gbase = s_start + RADIX_BINS;          // [256] global position of the run's element 0, minus s_start
    uint32_t* const s_warp = s_gbase + RADIX_BINS;           // [32]
    uint32_t& s_tile = s_warp[32];                           // [1] (+3 pad)
    uint16_t (*s_cnt)[RADIX_BINS] = reinterpret_cast<uint16_t (*)[RADIX_BINS]>(s_warp + 36);   // [WARPS][256] counts -> offsets
    // MATCH == 2: [WARPS][256] peer masks
    uint32_t (*s_match)[RADIX_BINS] = reinterpret_cast<uint32_t (*)[RADIX_BINS]>(&s_cnt[WARPS][0]);

    const uint32_t* ghist = sweep;
    uint32_t* ticket = sweep + 256;
    uint32_t* lookback = sweep + 260;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * RADIX_BINS; i += THREADS) (&s_cnt[0][0])[i] = 0;
    if (MATCH == 2)
        for (int i = threadIdx.x; i < WARPS * RADIX_BINS; i += THREADS) (&s_match[0][0])[i] = 0u;
    if (threadIdx.x < RADIX_BINS) s_start[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n = load_count(n_dev, n_host, cap);
    const uint32_t base = tile * RADIX_TILE;
    if (base >= n) return;      // tiles past the end are never waited on
    const int nbits = __popc(mask);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t in_tile = min((uint32_t)RADIX_TILE, n - base);

    // ---- stable ranks: (warp, iteration, lane) order == input order ----
    uint32_t key[ITEMS], val[ITEMS], rank[ITEMS], peers[ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const uint32_t idx = base + w * PER_WARP + it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? keys_in[idx] : 0u;
        val[it] = valid ? vals_in[idx] : 0u;
    }
    // Publish this tile's digit counts EARLY (a cheap unordered shared-memory histogram), long
    // before they are needed: the ranking below gives the preceding tiles time to resolve their
    // own prefixes, so the look-back further down usually finds an inclusive word one hop away.
#pragma unroll
    for (int it = 0; it < ITEMS; it++)
        if (base + w * PER_WARP + it * 32 + lane < n) atomicAdd(&s_start[(key[it] >> shift) & mask], 1u);
    __syncthreads();
    if (threadIdx.x < RADIX_BINS)
        st_volatile(lookback + (size_t)tile * RADIX_BINS + threadIdx.x,
                    (tile == 0 ? FLAG_INCLUSIVE : FLAG_PARTIAL) | s_start[threadIdx.x]);
    if (MATCH != 2) {
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const bool valid = base + w * PER_WARP + it * 32 + lane < n;
            const uint32_t d = (key[it] >> shift) & mask;
            // invalid lanes become singletons that match nobody
            peers[it] = MATCH == 1 ? peers_by_ballot(d, valid, nbits)
                                   : __match_any_sync(0xffffffffu, valid ? d : (0x10000u | (uint32_t)lane));
        }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const bool valid = base + w * PER_WARP + it * 32 + lane < n;
        const uint32_t d = (key[it] >> shift) & mask;
        if (MATCH == 2) {
            uint32_t* const mm = s_match[w];
            if (valid) atomicOr(&mm[d], 1u << lane);
            __syncwarp();
            peers[it] = valid ? mm[d] : (1u << lane);
        }
        const int leader = __ffs(peers[it]) - 1;
        uint32_t c = 0u;
        if (MATCH == 2) __syncwarp();          // every peer has read the mask before its leader clears it
        if (valid && lane == leader) {
            c = s_cnt[w][d];
            s_cnt[w][d] = (uint16_t)(c + __popc(peers[it]));
            if (MATCH == 2) s_match[w][d] = 0u;      // (the warp barrier that ends the iteration orders this before the next OR)
        }
        c = __shfl_sync(0xffffffffu, c, leader);
        rank[it] = c + __popc(peers[it] & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread d < 256): CTA total, publish, look back, global base ----
    {
        const uint32_t d = threadIdx.x;
        uint32_t total = 0u;
        if (d < RADIX_BINS) {
#pragma unroll
            for (int k = 0; k < WARPS; k++) {
                const uint32_t c = s_cnt[k][d];
                s_cnt[k][d] = (uint16_t)total;      // offset of warp k inside the digit's run
                total += c;
            }
        }
        // CTA-local start of the digit's run, and global start of the digit (all tiles)
        const uint32_t inc_local = block_inclusive_scan(total, s_warp, nullptr);
        __syncthreads();
        const uint32_t gh = d < RADIX_BINS ? ghist[d] : 0u;
        const uint32_t inc_global = block_inclusive_scan(gh, s_warp, nullptr);
        if (d < RADIX_BINS) {
            uint32_t excl = 0u;          // items with this digit in preceding tiles
            if (tile > 0) {
                // Decoupled look-back, LOOK predecessors per round trip: the loads of a window are
                // independent, so a chain of k unresolved predecessors costs k / LOOK L2 latencies.
                constexpr int LOOK = 8;
                int p = (int)tile - 1;
                bool resolved = false;
                while (!resolved) {
                    uint32_t v[LOOK];
#pragma unroll
                    for (int i = 0; i < LOOK; i++)
                        v[i] = p - i >= 0 ? ld_volatile(lookback + (size_t)(p - i) * RADIX_BINS + d) : FLAG_INCLUSIVE;
#pragma unroll
                    for (int i = 0; i < LOOK; i++) {
                        if (resolved) break;
                        const uint32_t f = v[i] & FLAG_MASK;
                        if (f == 0u) break;                 // not published yet: re-read from here
                        excl += v[i] & VALUE_MASK;
                        p--;
                        if (f == FLAG_INCLUSIVE) resolved = true;
                    }
                }
                st_volatile(lookback + (size_t)tile * RADIX_BINS + d, FLAG_INCLUSIVE | (excl + total));
            }
            const uint32_t start = inc_local - total;
            s_start[d] = start;
            s_gbase[d] = (inc_global - gh) + excl - start;
        }
    }
    __syncthreads();

    // ---- reorder through shared memory, then coalesced runs to global ----
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        if (base + w * PER_WARP + it * 32 + lane < n) {
            const uint32_t d = (key[it] >> shift) & mask;
            const uint32_t lp = s_start[d] + s_cnt[w][d] + rank[it];
            s_keys[lp] = key[it];
            s_vals[lp] = val[it];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < in_tile; i += THREADS) {
        const uint32_t k = s_keys[i];
        const uint32_t pos = s_gbase[(k >> shift) & mask] + i;
        vals_out[pos] = s_vals[i];
        if (RANGES) {
            if (i == 0 || s_keys[i - 1] != k) atomicMin(&ranges[k].x, pos);
            if (i == in_tile - 1 || s_keys[i + 1] != k) atomicMax(&ranges[k].y, pos + 1u);
        } else {
            keys_out[pos] = k;
        }
    }
}

constexpr size_t onesweep_smem_bytes(int threads, int match) {
    return (size_t)(2 * RADIX_TILE + 2 * RADIX_BINS + 36) * 4 + (size_t)(threads / 32) * RADIX_BINS * 2 +
           (match == 2 ? (size_t)(threads / 32) * RADIX_BINS * 4 : 0);
}

__global__ void init_ranges_kernel(uint2* __restrict__ ranges, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ranges[i] = make_uint2(0xFFFFFFFFu, 0u);     // empty: start > end
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int THREADS, int ITEMS, int MATCH, bool RANGES, int MINB = 0>
void launch_onesweep_variant(uint32_t nb, const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout,
                             const int64_t* n_dev, int64_t n_host, int64_t cap, int shift, uint32_t mask,
                             uint32_t* sweep, uint2* ranges, const Launch& L) {
    auto kern = onesweep_pass_kernel<THREADS, ITEMS, MATCH, RANGES, MINB>;
    constexpr size_t smem = onesweep_smem_bytes(THREADS, MATCH);
    static bool configured = false;     // per-process, idempotent
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    chain(kern, dim3(nb), dim3(THREADS), smem, L)(kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges);
}

// variant 1: 256 threads x 16 items; 2: 512 x 8; 4: 1024 x 4 (all ballot ranking); 0: 256 x 16, 3: 512 x 8 with MATCH.ANY;
// 5: 512 x 8, 6: 1024 x 4, 7: 256 x 16 with the shared-memory atomic-OR match
template <bool RANGES>
void launch_onesweep(const char* name, const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout,
                     const int64_t* n_dev, int64_t n_host, int64_t cap, int shift, uint32_t mask, uint32_t* sweep,
                     uint2* ranges, const Launch& L) {
    static const int variant = env_int("SCGR_SORT_VARIANT", 5);
    const uint32_t nb = radix_blocks(cap > 0 ? cap : 1);
    begin_kernel(name, L);
    switch (variant) {
        default: launch_onesweep_variant<512, 8, 2, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 8: launch_onesweep_variant<512, 8, 2, RANGES, 3>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 9: launch_onesweep_variant<256, 16, 2, RANGES, 4>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 6: launch_onesweep_variant<1024, 4, 2, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 7: launch_onesweep_variant<256, 16, 2, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 0: launch_onesweep_variant<256, 16, 0, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 1: launch_onesweep_variant<256, 16, 1, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 3: launch_onesweep_variant<512, 8, 0, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 4: launch_onesweep_variant<1024, 4, 1, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
        case 2: launch_onesweep_variant<512, 8, 1, RANGES>(nb, kin, vin, kout, vout, n_dev, n_host, cap, shift, mask, sweep, ranges, L); break;
    }
    check_launch(name, L);
}

// ------------------------------------------------------------------------------------------
// instance emission in depth order (A.6 without the depth half of the key).  Also accumulates the
// global digit histograms of the tile ids for the partition passes.
// ------------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;

struct TilePasses {
    int passes;
    int shift[MAX_TILE_PASSES];
    uint32_t mask[MAX_TILE_PASSES];
};

// position of the k-th (0-based) set bit of w; k < popc(w)
__device__ __forceinline__ uint32_t select_bit32(uint32_t w, uint32_t k) {
    uint32_t pos = 0u;
    uint32_t c = __popc(w & 0xFFFFu);
    if (k >= c) { k -= c; pos = 16u; w >>= 16; }
    c = __popc(w & 0xFFu);
    if (k >= c) { k -= c; pos += 8u; w >>= 8; }
    c = __popc(w & 0xFu);
    if (k >= c) { k -= c; pos += 4u; w >>= 4; }
    c = __popc(w & 0x3u);
    if (k >= c) { k -= c; pos += 2u; w >>= 2; }
    if (k >= (w & 1u)) pos += 1u;
    return pos;
}

// A warp owns 32 depth-ordered Gaussians, whose instance slices [offsets[s-1], offsets[s]) are one
// contiguous run of the output.  The run is produced 32 instances at a time, ONE INSTANCE PER LANE:
// the lane finds the Gaussian that owns its output position (5-step search over the 32 slice ends),
// fetches that Gaussian's survivor mask by shuffle and selects the k-th set bit -- so the work is
// balanced however uneven the rects are, and the (tile id, id) stores are fully coalesced.
// Rects of more than 64 tiles (no mask) are handled afterwards by the whole warp: the 32 lanes
// re-run the exact test on 32 tiles at a time and compact the survivors with a ballot.
template <int PASSES>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const uint2* __restrict__ rect, const unsigned long long* __restrict__ tile_mask,
                      const Record* __restrict__ rec, int P, int grid_x, int64_t capacity,
                      int64_t* __restrict__ status, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                      uint32_t* __restrict__ sweep, size_t pass_words, const TilePasses tp) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    __shared__ uint32_t s_hist[PASSES][RADIX_BINS];
    const int64_t R = status[0];
    if (R > capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 1;
        return;
    }
    // the flag describes THIS emission: a re-run with a large enough buffer (SCGR_NEED_CAPACITY / optimistic
    // recovery, which skip stage 1 and its reset) clears what the refused attempt raised
    if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 0;
#pragma unroll
    for (int p = 0; p < PASSES; p++) s_hist[p][threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int s = blockIdx.x * EMIT_THREADS + threadIdx.x;
    uint32_t gid = 0u, end = 0xFFFFFFFFu, begin = 0u, x0 = 0u, y0 = 0u, rw = 0u, rh = 0u;
    if (s < P) {
        gid = order[s];
        end = offsets[s];
        begin = s > 0 ? offsets[s - 1] : 0u;
        if (end != begin) {
            const uint2 rc = rect[gid];
            x0 = rc.x & 0xffffu; y0 = rc.x >> 16;
            rw = (rc.y & 0xffffu) - x0; rh = (rc.y >> 16) - y0;
        }
    }
    const bool has = s < P && end != begin;
    const bool big = rw * rh > 64u;
    // digit extraction of the partition passes, in registers
    int sh[PASSES];
    uint32_t mk[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; p++) { sh[p] = tp.shift[p]; mk[p] = tp.mask[p]; }
    {
        unsigned long long m = 0ull;
        if (has && !big) m = tile_mask[gid];
        const uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32);
        const uint32_t rowbase = y0 * (uint32_t)grid_x + x0;                        // tile id of the rect's origin
        // rw (7 bits) | ceil(2^16 / rw) (17 bits): ty = (bit * inv) >> 16 is exact for bit < 64
        const uint32_t geom = rw | ((rw ? (65536u + rw - 1u) / rw : 0u) << 7) | (big ? 0x80000000u : 0u);
        const uint32_t run_begin = __shfl_sync(0xffffffffu, begin, 0);
        const uint32_t run_end = __reduce_max_sync(0xffffffffu, s < P ? end : 0u);
        for (uint32_t base = run_begin; base < run_end; base += 32u) {
            const uint32_t pos = base + (uint32_t)lane;
            // owner = number of slices that end at or before pos (ends are non-decreasing; lanes past P hold +inf)
            int owner = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t e = __shfl_sync(0xffffffffu, end, owner + step - 1);
                if (e <= pos) owner += step;
            }
            owner &= 31;                                                               // (lanes past run_end)
            const uint32_t ob = __shfl_sync(0xffffffffu, begin, owner);
            const uint32_t olo = __shfl_sync(0xffffffffu, mlo, owner), ohi = __shfl_sync(0xffffffffu, mhi, owner);
            const uint32_t obase = __shfl_sync(0xffffffffu, rowbase, owner);
            const uint32_t ogeom = __shfl_sync(0xffffffffu, geom, owner);
            const uint32_t og = __shfl_sync(0xffffffffu, gid, owner);
            if (pos < run_end && !(ogeom >> 31)) {
                uint32_t k = pos - ob;
                const uint32_t clo = __popc(olo);
                const bool upper = k >= clo;
                const uint32_t bit = select_bit32(upper ? ohi : olo, upper ? k - clo : k) + (upper ? 32u : 0u);
                const uint32_t orw = ogeom & 127u, inv = (ogeom >> 7) & 0x1FFFFu;
                const uint32_t ty = (bit * inv) >> 16, tx = bit - ty * orw;
                const uint32_t tile = obase + ty * (uint32_t)grid_x + tx;
                keys[pos] = tile;
                vals[pos] = og;
#pragma unroll
                for (int p = 0; p < PASSES; p++) atomicAdd(&s_hist[p][(tile >> sh[p]) & mk[p]], 1u);
            }
        }
    }
    // large rects, warp-cooperatively
    uint32_t todo = __ballot_sync(0xffffffffu, has && big);
    if (todo) {
        float4 q0 = make_float4(0.f, 0.f, -1.f, 0.f), q1 = make_float4(-1.f, 0.f, 0.f, 0.f);
        if (has && big) {
            const float4* r = reinterpret_cast<const float4*>(rec + gid);
            q0 = __ldg(r);
            q1 = __ldg(r + 1);
        }
        const CullParams mine = make_cull(q0, q1);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t b = __shfl_sync(0xffffffffu, begin, j);
            const uint32_t e = __shfl_sync(0xffffffffu, end, j);
            const uint32_t g = __shfl_sync(0xffffffffu, gid, j);
            const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, j), by0 = __shfl_sync(0xffffffffu, y0, j);
            const uint32_t brw = __shfl_sync(0xffffffffu, rw, j), brh = __shfl_sync(0xffffffffu, rh, j);
            CullParams c;
            c.cA = __shfl_sync(0xffffffffu, mine.cA, j); c.cB = __shfl_sync(0xffffffffu, mine.cB, j);
            c.cC = __shfl_sync(0xffffffffu, mine.cC, j); c.kx = __shfl_sync(0xffffffffu, mine.kx, j);
            c.ky = __shfl_sync(0xffffffffu, mine.ky, j); c.mx = __shfl_sync(0xffffffffu, mine.mx, j);
            c.my = __shfl_sync(0xffffffffu, mine.my, j); c.thr = __shfl_sync(0xffffffffu, mine.thr, j);
            const uint32_t ntiles = brw * brh;
            uint32_t out = b;
            const SpanParams sp = make_span(c);      // same decision path as the counting pass (preprocess)
            if (sp.robust) {
                // a lane per tile row: closed-form column span, prefix sum over the rows, each lane writes its run
                for (uint32_t r0 = 0; r0 < brh; r0 += 32) {
                    const uint32_t r = r0 + lane;
                    int first = 0, n = 0;
                    if (r < brh) n = row_span(sp, (int)(by0 + r), (int)bx0, (int)(bx0 + brw), &first);
                    const uint32_t inc = warp_inclusive_scan((uint32_t)n, lane);
                    uint32_t pos = out + inc - (uint32_t)n;
                    for (int t = 0; t < n; t++, pos++) {
                        if (pos < e) {      // always true: same bit-exact spans as the counting pass
                            const uint32_t tile = (by0 + r) * (uint32_t)grid_x + (uint32_t)(first + t);
                            keys[pos] = tile;
                            vals[pos] = g;
#pragma unroll
                            for (int p = 0; p < PASSES; p++) atomicAdd(&s_hist[p][(tile >> sh[p]) & mk[p]], 1u);
                        }
                    }
                    out += __shfl_sync(0xffffffffu, inc, 31);
                }
                continue;
            }
            for (uint32_t k0 = 0; k0 < ntiles; k0 += 32) {
                const uint32_t k = k0 + lane;
                bool pass = false;
                uint32_t tile = 0u;
                if (k < ntiles) {
                    const uint32_t ty = k / brw, tx = k - ty * brw;
                    pass = tile_may_contribute(c, (int)(bx0 + tx), (int)(by0 + ty));
                    tile = (by0 + ty) * (uint32_t)grid_x + bx0 + tx;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const uint32_t pos = out + __popc(bal & lt_mask);
                    if (pos < e) {      // always true: same bit-exact test as the counting pass
                        keys[pos] = tile;
                        vals[pos] = g;
#pragma unroll
                        for (int p = 0; p < PASSES; p++) atomicAdd(&s_hist[p][(tile >> sh[p]) & mk[p]], 1u);
                    }
                }
                out += __popc(bal);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PASSES; p++) {
        const uint32_t c = s_hist[p][threadIdx.x];
        if (c) atomicAdd(sweep + p * pass_words + threadIdx.x, c);
    }
}

int bits_for(uint32_t n_values) {   // bits needed to represent 0 .. n_values-1
    int b = 0;
    while (b < 32 && (1ull << b) < n_values) b++;
    return b > 0 ? b : 1;
}

TilePasses plan_tile_passes(uint32_t n_tiles) {
    TilePasses tp{};
    const int bits = bits_for(n_tiles);
    tp.passes = (bits + 7) / 8;
    int shift = 0;
    for (int p = 0; p < tp.passes; p++) {
        // spread the bits evenly over the passes (e.g. 13 bits -> 7 + 6)
        const int pb = (bits - shift + (tp.passes - p) - 1) / (tp.passes - p);
        tp.shift[p] = shift;
        tp.mask[p] = (1u << pb) - 1u;
        shift += pb;
    }
    return tp;
}

}  // namespace

// Number of ping-pong flips the tile partition performs for a given tile count (needed by the
// backward to find the final point list without any saved host state).
int tile_partition_final_buffer(uint32_t n_tiles) { return plan_tile_passes(n_tiles).passes & 1; }

// Depth order of the Gaussians (ascending depth bits, ties by index; near-plane-culled ones last).
// `L` is the auxiliary stream: key generation + the four radix passes run beside the preprocess.
void launch_depth_sort(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G, const Launch& L) {
    const int32_t P = g.P;
    if (P <= 0) return;
    const size_t pw = sweep_pass_words(P);
    // one memset clears the onesweep state of the 4 passes and the scan state behind it
    cudaMemsetAsync(G.sweep, 0, (size_t)((char*)G.scan_state - (char*)G.sweep) + scan_state_bytes(P), L.stream);
    launch_depth_keys(v, g, G, L);      // writes sort_keys[0] (= depth_key), sort_vals[0] (= 0..P-1), the 4 histograms
    int cur = 0;
    for (int p = 0; p < 4; p++) {
        launch_onesweep<false>("depth_sort_pass", G.sort_keys[cur], G.sort_vals[cur], G.sort_keys[cur ^ 1],
                               G.sort_vals[cur ^ 1], nullptr, P, P, 8 * p, 255u, G.sweep + p * pw, nullptr, L);
        cur ^= 1;
    }
    // 4 passes -> result is back in buffer 0
}

// Inclusive prefix sum of tiles touched in depth order, and R.  Needs both the depth order
// (auxiliary stream) and tiles_touched (preprocess).
void launch_scan_offsets(const GeometryLayout& G, int32_t P, int64_t* status_mapped, const Launch& L) {
    if (P <= 0) return;
    const uint32_t* order = G.sort_vals[0];
    const int nblk = (P + SCAN_TILE - 1) / SCAN_TILE;
    begin_kernel("scan_offsets", L);
    scan_offsets_kernel<<<nblk, SCAN_BLOCK, 0, L.stream>>>(G.tiles_touched, order, P, G.offsets, G.status, G.scan_state,
                                                           status_mapped);
    check_launch("scan_offsets", L);
}

// State of the tile partition that does not depend on R: empty ranges, zeroed onesweep words.
// Split off so that scgr_forward() can enqueue it before it waits for R.
void launch_binning_prologue(const ScgrView& v, const BinningLayout& B, int32_t P, int64_t capacity, const Launch& L) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const uint32_t n_tiles = (uint32_t)gx * gy;
    begin_kernel("init_ranges", L);
    init_ranges_kernel<<<(n_tiles + 255) / 256, 256, 0, L.stream>>>(B.ranges, n_tiles);
    check_launch("init_ranges", L);
    if (P <= 0) return;
    const TilePasses tp = plan_tile_passes(n_tiles);
    cudaMemsetAsync(B.sweep, 0, sweep_pass_words(capacity) * tp.passes * 4, L.stream);
}

void launch_emit_and_partition(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                               int32_t P, int64_t capacity, int* final_buffer, const Launch& L) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const uint32_t n_tiles = (uint32_t)gx * gy;
    if (P <= 0) { if (final_buffer) *final_buffer = 0; return; }
    const TilePasses tp = plan_tile_passes(n_tiles);
    const size_t pw = sweep_pass_words(capacity);
    const uint32_t* order = G.sort_vals[0];   // 32-bit sort = 4 passes = even number of flips
    begin_kernel("emit_instances", L);
#define SCGR_EMIT(N_) chain(emit_instances_kernel<N_>, dim3((P + EMIT_THREADS - 1) / EMIT_THREADS), dim3(EMIT_THREADS), 0, L)( \
        order, G.offsets, G.rect, G.tile_mask, G.rec, P, gx, capacity, G.status, B.keys[0], B.vals[0], B.sweep, pw, tp)
    switch (tp.passes) {
        case 1: SCGR_EMIT(1); break;
        case 2: SCGR_EMIT(2); break;
        case 3: SCGR_EMIT(3); break;
        default: SCGR_EMIT(4); break;
    }
#undef SCGR_EMIT
    check_launch("emit_instances", L);
    int cur = 0;
    for (int p = 0; p < tp.passes; p++) {
        if (p == tp.passes - 1)
            launch_onesweep<true>("tile_partition_pass", B.keys[cur], B.vals[cur], B.keys[cur ^ 1], B.vals[cur ^ 1], G.status,
                                  0, capacity, tp.shift[p], tp.mask[p], B.sweep + p * pw, B.ranges, L);
        else
            launch_onesweep<false>("tile_partition_pass", B.keys[cur], B.vals[cur], B.keys[cur ^ 1], B.vals[cur ^ 1], G.status,
                                   0, capacity, tp.shift[p], tp.mask[p], B.sweep + p * pw, nullptr, L);
        cur ^= 1;
    }
    if (final_buffer) *final_buffer = cur;
}

}  // namespace scgr
