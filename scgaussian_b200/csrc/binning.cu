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
// (by a histogram kernel for the depth keys, by the emission kernel for the tile ids), each CTA
// takes a dynamic tile ticket, ranks its 4096 items stably with warp match + per-warp counters,
// obtains its global digit offsets by decoupled look-back over the preceding CTAs, reorders the
// items through shared memory and writes them out in coalesced runs.
//
// Nothing here needs R on the host: every kernel reads the instance count from device memory
// and clamps its work to the binning capacity (overflow is flagged, never written past).
// All passes are HBM/L2-bound integer work -- no tensor cores.
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

// pass 1: per-block sums of tiles_touched[order[s]]
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_reduce_kernel(const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ order, int P,
                   uint32_t* __restrict__ partials) {
    __shared__ uint32_t s_warp[32];
    const int s = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = s < P ? tiles_touched[order[s]] : 0u;
    uint32_t total;
    block_inclusive_scan(v, s_warp, &total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// pass 2 (single block): exclusive scan of the partials in place; grand total -> status[0]
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_partials_kernel(uint32_t* __restrict__ partials, int n, int64_t* __restrict__ status) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? partials[i] : 0u;
        uint32_t total;
        const uint32_t inc = block_inclusive_scan(v, s_warp, &total);
        const uint32_t carry = s_carry;
        if (i < n) partials[i] = carry + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        status[0] = (int64_t)s_carry;   // R = num_rendered
        status[1] = 0;                  // overflow flag, raised later by the emission kernel
    }
}

// pass 3: offsets[s] = inclusive prefix of tiles_touched in depth order
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_finish_kernel(const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ order, int P,
                   const uint32_t* __restrict__ partials, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t s_warp[32];
    const int s = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = s < P ? tiles_touched[order[s]] : 0u;
    const uint32_t inc = block_inclusive_scan(v, s_warp, nullptr);
    if (s < P) offsets[s] = partials[blockIdx.x] + inc;
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

// global digit histograms of the 4 byte-digits of the depth keys -> sweep[pass][0..255]
constexpr int HIST_ITEMS = 32;
__global__ void __launch_bounds__(256)
depth_hist_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ sweep, size_t pass_words) {
    __shared__ uint32_t s_hist[4][RADIX_BINS];
#pragma unroll
    for (int p = 0; p < 4; p++) s_hist[p][threadIdx.x] = 0u;
    __syncthreads();
    const int base = blockIdx.x * 256 * HIST_ITEMS;
#pragma unroll 4
    for (int it = 0; it < HIST_ITEMS; it++) {
        const int idx = base + it * 256 + threadIdx.x;
        if (idx < n) {
            const uint32_t k = keys[idx];
            atomicAdd(&s_hist[0][k & 255u], 1u);
            atomicAdd(&s_hist[1][(k >> 8) & 255u], 1u);
            atomicAdd(&s_hist[2][(k >> 16) & 255u], 1u);
            atomicAdd(&s_hist[3][k >> 24], 1u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const uint32_t c = s_hist[p][threadIdx.x];
        if (c) atomicAdd(sweep + p * pass_words + threadIdx.x, c);
    }
}

// One stable pass.  sweep = this pass's state (layout in common.cuh), zero before the launch except
// for the histogram, which must be complete.
template <bool WRITE_KEYS>
__global__ void __launch_bounds__(RADIX_THREADS)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const int64_t* __restrict__ n_dev, int64_t n_host, int64_t cap, int shift,
                     uint32_t mask, uint32_t* __restrict__ sweep) {
    constexpr int WARPS = RADIX_THREADS / 32;
    constexpr int PER_WARP = RADIX_TILE / WARPS;   // 512
    __shared__ uint32_t s_cnt[WARPS][RADIX_BINS];  // per-warp digit counts -> offsets
    __shared__ uint32_t s_start[RADIX_BINS];       // CTA-local sorted position of each digit's run
    __shared__ uint32_t s_gbase[RADIX_BINS];       // global position of element 0 of each digit's run, minus s_start
    __shared__ uint32_t s_keys[RADIX_TILE];
    __shared__ uint32_t s_vals[RADIX_TILE];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_tile;

    const uint32_t* ghist = sweep;
    uint32_t* ticket = sweep + 256;
    uint32_t* lookback = sweep + 260;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int k = 0; k < WARPS; k++) s_cnt[k][threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n = load_count(n_dev, n_host, cap);
    const uint32_t base = tile * RADIX_TILE;
    if (base >= n) return;      // tiles past the end are never waited on
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t in_tile = min((uint32_t)RADIX_TILE, n - base);

    // ---- stable ranks: (warp, iteration, lane) order == input order ----
    uint32_t key[RADIX_ITEMS], val[RADIX_ITEMS], rank[RADIX_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        const uint32_t idx = base + w * PER_WARP + it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? keys_in[idx] : 0u;
        val[it] = valid ? vals_in[idx] : 0u;
    }
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        const bool valid = base + w * PER_WARP + it * 32 + lane < n;
        const uint32_t d = (key[it] >> shift) & mask;
        // invalid lanes become singletons that match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : (0x10000u | (uint32_t)lane));
        const int leader = __ffs(peers) - 1;
        uint32_t c = 0u;
        if (valid && lane == leader) {
            c = s_cnt[w][d];
            s_cnt[w][d] = c + __popc(peers);
        }
        c = __shfl_sync(0xffffffffu, c, leader);
        rank[it] = c + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread d): CTA total, publish, look back, global base ----
    {
        const uint32_t d = threadIdx.x;
        uint32_t total = 0u;
#pragma unroll
        for (int k = 0; k < WARPS; k++) {
            const uint32_t c = s_cnt[k][d];
            s_cnt[k][d] = total;      // offset of warp k inside the digit's run
            total += c;
        }
        st_volatile(lookback + (size_t)tile * RADIX_BINS + d, (tile == 0 ? FLAG_INCLUSIVE : FLAG_PARTIAL) | total);
        // CTA-local start of the digit's run, and global start of the digit (all tiles)
        const uint32_t inc_local = block_inclusive_scan(total, s_warp, nullptr);
        __syncthreads();
        const uint32_t gh = ghist[d];
        const uint32_t inc_global = block_inclusive_scan(gh, s_warp, nullptr);
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
    __syncthreads();

    // ---- reorder through shared memory, then coalesced runs to global ----
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        if (base + w * PER_WARP + it * 32 + lane < n) {
            const uint32_t d = (key[it] >> shift) & mask;
            const uint32_t lp = s_start[d] + s_cnt[w][d] + rank[it];
            s_keys[lp] = key[it];
            s_vals[lp] = val[it];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < in_tile; i += RADIX_THREADS) {
        const uint32_t k = s_keys[i];
        const uint32_t pos = s_gbase[(k >> shift) & mask] + i;
        if (WRITE_KEYS) keys_out[pos] = k;
        vals_out[pos] = s_vals[i];
    }
}

// ------------------------------------------------------------------------------------------
// instance emission in depth order (A.6 without the depth half of the key)
// One warp per 32 consecutive depth-ordered Gaussians; for each of them the 32 lanes test the
// tiles of its rect in parallel (exact culling), compact the survivors with a ballot and write
// them out -> coalesced stores, no per-thread rect loops.  Also accumulates the global digit
// histograms of the tile ids for the partition passes.
// ------------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_GROUPS_PER_WARP = 4;   // 8 warps x 4 x 32 = 1024 Gaussians per CTA

struct TilePasses {
    int passes;
    int shift[MAX_TILE_PASSES];
    uint32_t mask[MAX_TILE_PASSES];
};

__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const uint2* __restrict__ rect, const Record* __restrict__ rec, int P, int grid_x,
                      int64_t capacity, int64_t* __restrict__ status, uint32_t* __restrict__ keys,
                      uint32_t* __restrict__ vals, uint32_t* __restrict__ sweep, size_t pass_words,
                      const TilePasses tp) {
    __shared__ uint32_t s_hist[MAX_TILE_PASSES][RADIX_BINS];
    const int64_t R = status[0];
    if (R > capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 1;
        return;
    }
    for (int p = 0; p < tp.passes; p++) s_hist[p][threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int gi = 0; gi < EMIT_GROUPS_PER_WARP; gi++) {
        const int s = ((blockIdx.x * (EMIT_THREADS / 32) + w) * EMIT_GROUPS_PER_WARP + gi) * 32 + lane;
        uint32_t gid = 0u, end = 0u, begin = 0u;
        uint2 rc = make_uint2(0u, 0u);
        float4 q0 = make_float4(0.f, 0.f, -1.f, 0.f), q1 = make_float4(-1.f, 0.f, 0.f, 0.f);
        if (s < P) {
            gid = order[s];
            end = offsets[s];
            begin = s > 0 ? offsets[s - 1] : 0u;
            if (end != begin) {
                rc = rect[gid];
                const float4* r = reinterpret_cast<const float4*>(rec + gid);
                q0 = __ldg(r);
                q1 = __ldg(r + 1);
            }
        }
        const CullParams mine = make_cull(q0, q1);
        uint32_t todo = __ballot_sync(0xffffffffu, end != begin);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t b = __shfl_sync(0xffffffffu, begin, j);
            const uint32_t e = __shfl_sync(0xffffffffu, end, j);
            const uint32_t g = __shfl_sync(0xffffffffu, gid, j);
            const uint32_t rmin = __shfl_sync(0xffffffffu, rc.x, j);
            const uint32_t rmax = __shfl_sync(0xffffffffu, rc.y, j);
            CullParams c;
            c.cA = __shfl_sync(0xffffffffu, mine.cA, j); c.cB = __shfl_sync(0xffffffffu, mine.cB, j);
            c.cC = __shfl_sync(0xffffffffu, mine.cC, j); c.kx = __shfl_sync(0xffffffffu, mine.kx, j);
            c.ky = __shfl_sync(0xffffffffu, mine.ky, j); c.mx = __shfl_sync(0xffffffffu, mine.mx, j);
            c.my = __shfl_sync(0xffffffffu, mine.my, j); c.thr = __shfl_sync(0xffffffffu, mine.thr, j);
            const uint32_t x0 = rmin & 0xffffu, y0 = rmin >> 16;
            const uint32_t rw = (rmax & 0xffffu) - x0, rh = (rmax >> 16) - y0;
            const uint32_t ntiles = rw * rh;
            uint32_t out = b;
            for (uint32_t k0 = 0; k0 < ntiles; k0 += 32) {
                const uint32_t k = k0 + lane;
                bool pass = false;
                uint32_t tile = 0u;
                if (k < ntiles) {
                    const uint32_t ty = k / rw, tx = k - ty * rw;
                    pass = tile_may_contribute(c, (int)(x0 + tx), (int)(y0 + ty));
                    tile = (y0 + ty) * (uint32_t)grid_x + x0 + tx;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const uint32_t pos = out + __popc(bal & lt_mask);
                    if (pos < e) {      // always true: same bit-exact test as the counting pass
                        keys[pos] = tile;
                        vals[pos] = g;
                        for (int p = 0; p < tp.passes; p++) atomicAdd(&s_hist[p][(tile >> tp.shift[p]) & tp.mask[p]], 1u);
                    }
                }
                out += __popc(bal);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < tp.passes; p++) {
        const uint32_t c = s_hist[p][threadIdx.x];
        if (c) atomicAdd(sweep + p * pass_words + threadIdx.x, c);
    }
}

// A.7
__global__ void identify_ranges_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ status,
                                       int64_t capacity, uint2* __restrict__ ranges) {
    const int64_t R = status[0];
    if (R > capacity) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t t = keys[i];
    if (i == 0 || keys[i - 1] != t) ranges[t].x = (uint32_t)i;
    if (i == R - 1 || keys[i + 1] != t) ranges[t].y = (uint32_t)(i + 1);
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

// depth order of the Gaussians (ascending depth bits, ties by index; culled ones last), then the
// inclusive prefix sum of tiles touched in that order and R.
void launch_depth_order(const GeometryLayout& G, int32_t P, const Launch& L) {
    if (P <= 0) return;
    // preprocess already wrote sort_keys[0] (= depth_key) and sort_vals[0] (= 0..P-1)
    const size_t pw = sweep_pass_words(P);
    cudaMemsetAsync(G.sweep, 0, sweep_words(P, 4) * 4, L.stream);
    begin_kernel("depth_hist", L);
    depth_hist_kernel<<<(P + 256 * HIST_ITEMS - 1) / (256 * HIST_ITEMS), 256, 0, L.stream>>>(G.sort_keys[0], P, G.sweep, pw);
    check_launch("depth_hist", L);
    const uint32_t nb = radix_blocks(P);
    int cur = 0;
    for (int p = 0; p < 4; p++) {
        begin_kernel("depth_sort_pass", L);
        onesweep_pass_kernel<true><<<nb, RADIX_THREADS, 0, L.stream>>>(
            G.sort_keys[cur], G.sort_vals[cur], G.sort_keys[cur ^ 1], G.sort_vals[cur ^ 1], nullptr, P, P, 8 * p, 255u,
            G.sweep + p * pw);
        check_launch("depth_sort_pass", L);
        cur ^= 1;
    }
    // 4 passes -> result is back in buffer 0
    const uint32_t* order = G.sort_vals[0];
    const int nblk = (P + SCAN_BLOCK - 1) / SCAN_BLOCK;
    begin_kernel("scan_reduce", L);
    scan_reduce_kernel<<<nblk, SCAN_BLOCK, 0, L.stream>>>(G.tiles_touched, order, P, G.scan_partials);
    check_launch("scan_reduce", L);
    begin_kernel("scan_partials", L);
    scan_partials_kernel<<<1, SCAN_BLOCK, 0, L.stream>>>(G.scan_partials, nblk, G.status);
    check_launch("scan_partials", L);
    begin_kernel("scan_finish", L);
    scan_finish_kernel<<<nblk, SCAN_BLOCK, 0, L.stream>>>(G.tiles_touched, order, P, G.scan_partials, G.offsets);
    check_launch("scan_finish", L);
}

void launch_emit_and_partition(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                               int32_t P, int64_t capacity, int* final_buffer, const Launch& L) {
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    const uint32_t n_tiles = (uint32_t)gx * gy;
    cudaMemsetAsync(B.ranges, 0, (size_t)n_tiles * sizeof(uint2), L.stream);
    if (P <= 0) { if (final_buffer) *final_buffer = 0; return; }
    const TilePasses tp = plan_tile_passes(n_tiles);
    const size_t pw = sweep_pass_words(capacity);
    cudaMemsetAsync(B.sweep, 0, pw * tp.passes * 4, L.stream);
    const uint32_t* order = G.sort_vals[0];   // 32-bit sort = 4 passes = even number of flips
    const int per_cta = (EMIT_THREADS / 32) * EMIT_GROUPS_PER_WARP * 32;
    begin_kernel("emit_instances", L);
    emit_instances_kernel<<<(P + per_cta - 1) / per_cta, EMIT_THREADS, 0, L.stream>>>(
        order, G.offsets, G.rect, G.rec, P, gx, capacity, G.status, B.keys[0], B.vals[0], B.sweep, pw, tp);
    check_launch("emit_instances", L);
    const uint32_t nb = radix_blocks(capacity > 0 ? capacity : 1);
    int cur = 0;
    for (int p = 0; p < tp.passes; p++) {
        begin_kernel("tile_partition_pass", L);
        onesweep_pass_kernel<true><<<nb, RADIX_THREADS, 0, L.stream>>>(
            B.keys[cur], B.vals[cur], B.keys[cur ^ 1], B.vals[cur ^ 1], G.status, 0, capacity, tp.shift[p], tp.mask[p],
            B.sweep + p * pw);
        check_launch("tile_partition_pass", L);
        cur ^= 1;
    }
    if (final_buffer) *final_buffer = cur;
    const int64_t blocks = (capacity + 255) / 256;
    if (blocks > 0) {
        begin_kernel("identify_ranges", L);
        identify_ranges_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(B.keys[cur], G.status, capacity, B.ranges);
        check_launch("identify_ranges", L);
    }
}

}  // namespace scgr
