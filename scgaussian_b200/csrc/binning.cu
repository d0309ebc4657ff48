// binning.cu -- depth ordering, instance emission, stable tile partition, tile ranges (sm_100a).
//
// Replaces the external rasterizer's InclusiveSum + blocking D2H + duplicateWithKeys +
// cub::DeviceRadixSort::SortPairs<uint64,uint32> + identifyTileRanges (SURVEY.md section 2c,
// section 8a rows a10-a13; semantics Appendix A.6-A.7) with a different, cheaper decomposition
// that yields the *same* per-tile lists:
//
//   reference:  emit R (tile<<32 | depth) keys  ->  one stable 64-bit LSD sort over R items
//               (6 eight-bit passes at 1080p, ~152 B per instance of HBM traffic)
//   here:       an LSD sort processes the least-significant field first, and the depth field is
//               a property of the *Gaussian*, not of the instance.  So:
//                 1. stable 32-bit radix sort of the P Gaussians by depth bits   (P items)
//                 2. emit instances in that order                                (R items, 8 B each)
//                 3. stable partition by tile id: ceil(bits(Tn)/8) passes        (R items)
//               Stability makes equal-depth ties resolve by Gaussian index, exactly as A.6.
//               At 1080p this is 2 passes over R instead of 6, on 8-byte instead of 12-byte pairs.
//
// Nothing here needs R on the host: every kernel reads the instance count from device memory
// and clamps its work to the binning capacity (overflow is flagged, never written past).
//
// All passes are HBM/L2-bound integer work: coalesced 128-bit-friendly loads, shared-memory
// histograms, warp match/ballot ranking -- no tensor cores, no atomics on global memory.
#include "common.cuh"

namespace scgr {

namespace {

// ------------------------------------------------------------------------------------------
// exclusive/inclusive scan of tiles_touched gathered in depth order
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide inclusive scan for SCAN_BLOCK (1024) threads; returns inclusive value, total in *total
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
// stable LSD radix pass:  histogram -> per-digit row scan -> ranked scatter
// hist layout: hist[digit * nb_max + block]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_count(const int64_t* n_dev, int64_t n_host, int64_t cap) {
    int64_t n = n_dev ? *n_dev : n_host;
    if (n > cap) n = 0;   // overflow: caller re-runs with a larger buffer; do nothing now
    return (uint32_t)n;
}

__global__ void __launch_bounds__(RADIX_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ n_dev, int64_t n_host,
                  int64_t cap, int shift, uint32_t mask, uint32_t* __restrict__ hist, uint32_t nb_max) {
    __shared__ uint32_t s_hist[RADIX_BINS];
    const uint32_t n = load_count(n_dev, n_host, cap);
    const uint32_t base = blockIdx.x * RADIX_TILE;
    if (base >= n) return;
    s_hist[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        const uint32_t idx = base + it * RADIX_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&s_hist[(keys[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nb_max + blockIdx.x] = s_hist[threadIdx.x];
}

// one CTA per digit: exclusive scan along the active blocks of its row, row total -> totals[digit]
__global__ void __launch_bounds__(RADIX_THREADS)
radix_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ totals, const int64_t* __restrict__ n_dev,
                  int64_t n_host, int64_t cap, uint32_t nb_max) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t n = load_count(n_dev, n_host, cap);
    const uint32_t nb = (n + RADIX_TILE - 1) / RADIX_TILE;
    uint32_t* row = hist + (size_t)blockIdx.x * nb_max;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += RADIX_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? row[i] : 0u;
        uint32_t total;
        const uint32_t inc = block_inclusive_scan(v, s_warp, &total);
        const uint32_t carry = s_carry;
        if (i < nb) row[i] = carry + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

// Ranked scatter.  Each warp owns a contiguous run of 512 items of the CTA's 4096-item tile and
// walks it 32 items at a time, so that (warp, iteration, lane) order == input order: ranks
// computed with __match_any_sync + a per-warp digit counter are stable by construction.
template <bool WRITE_KEYS>
__global__ void __launch_bounds__(RADIX_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const int64_t* __restrict__ n_dev, int64_t n_host, int64_t cap, int shift,
                     uint32_t mask, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals,
                     uint32_t nb_max) {
    constexpr int WARPS = RADIX_THREADS / 32;
    constexpr int PER_WARP = RADIX_TILE / WARPS;   // 512
    __shared__ uint32_t s_cnt[WARPS][RADIX_BINS];
    __shared__ uint32_t s_warp[32];
    const uint32_t n = load_count(n_dev, n_host, cap);
    const uint32_t base = blockIdx.x * RADIX_TILE;
    if (base >= n) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < WARPS; k++) s_cnt[k][threadIdx.x] = 0u;
    __syncthreads();

    uint32_t key[RADIX_ITEMS], val[RADIX_ITEMS], rank[RADIX_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        const uint32_t idx = base + w * PER_WARP + it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? keys_in[idx] : 0u;
        val[it] = valid ? vals_in[idx] : 0u;
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
    // digit base = (#items with smaller digit, all blocks) + (#items with this digit in earlier
    // blocks) ; then running offsets across the 8 warps of this CTA
    {
        const uint32_t d = threadIdx.x;
        const uint32_t tot = totals[d];
        const uint32_t inc = block_inclusive_scan(tot, s_warp, nullptr);
        uint32_t run = inc - tot + hist[d * nb_max + blockIdx.x];
#pragma unroll
        for (int k = 0; k < WARPS; k++) {
            const uint32_t c = s_cnt[k][d];
            s_cnt[k][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RADIX_ITEMS; it++) {
        const uint32_t idx = base + w * PER_WARP + it * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[it] >> shift) & mask;
            const uint32_t pos = s_cnt[w][d] + rank[it];
            if (WRITE_KEYS) keys_out[pos] = key[it];
            vals_out[pos] = val[it];
        }
    }
}

// ------------------------------------------------------------------------------------------
// instance emission in depth order (A.6 without the depth half of the key)
// one warp per 32 consecutive depth-ordered Gaussians; for each of them all 32 lanes write its
// tile ids cooperatively -> coalesced stores, no per-thread rect loops (the reference's
// duplicateWithKeys is one thread per Gaussian looping over its whole rect)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emit_instances_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const uint2* __restrict__ rect, int P, int grid_x, int64_t capacity,
                      int64_t* __restrict__ status, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t R = status[0];
    if (R > capacity) {
        if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 1;
        return;
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int s = warp_global * 32 + lane;
    uint32_t gid = 0u, end = 0u, begin = 0u;
    uint2 rc = make_uint2(0u, 0u);
    if (s < P) {
        gid = order[s];
        end = offsets[s];
        begin = s > 0 ? offsets[s - 1] : 0u;
        rc = rect[gid];
    }
    for (int j = 0; j < 32; j++) {
        const uint32_t b = __shfl_sync(0xffffffffu, begin, j);
        const uint32_t e = __shfl_sync(0xffffffffu, end, j);
        if (e == b) continue;   // culled or out of range (warp-uniform)
        const uint32_t g = __shfl_sync(0xffffffffu, gid, j);
        const uint32_t rmin = __shfl_sync(0xffffffffu, rc.x, j);
        const uint32_t rmax = __shfl_sync(0xffffffffu, rc.y, j);
        const uint32_t x0 = rmin & 0xffffu, y0 = rmin >> 16;
        const uint32_t w = (rmax & 0xffffu) - x0;
        for (uint32_t k = lane; k < e - b; k += 32) {
            const uint32_t ty = k / w, tx = k - ty * w;
            keys[b + k] = (y0 + ty) * (uint32_t)grid_x + x0 + tx;
            vals[b + k] = g;
        }
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

}  // namespace

// Sorts (keys, vals) pairs on key bits [begin_bit, end_bit) with ceil(bits/8) stable passes.
// Result lands in keys[*final_buffer], vals[*final_buffer].
void radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], const int64_t* n_dev, int64_t n_host,
                      int64_t cap, int begin_bit, int end_bit, uint32_t* hist, uint32_t* totals,
                      int* final_buffer, const Launch& L) {
    const int bits = end_bit - begin_bit;
    const int passes = (bits + 7) / 8;
    const uint32_t nb_max = radix_blocks(cap > 0 ? cap : 1);
    int cur = 0;
    int shift = begin_bit;
    for (int p = 0; p < passes; p++) {
        // spread the bits evenly over the passes (e.g. 13 bits -> 7 + 6)
        const int pb = (bits - (shift - begin_bit) + (passes - p) - 1) / (passes - p);
        const uint32_t mask = (1u << pb) - 1u;
        begin_kernel("radix_hist", L);
        radix_hist_kernel<<<nb_max, RADIX_THREADS, 0, L.stream>>>(keys[cur], n_dev, n_host, cap, shift, mask, hist, nb_max);
        check_launch("radix_hist", L);
        begin_kernel("radix_scan", L);
        radix_scan_kernel<<<RADIX_BINS, RADIX_THREADS, 0, L.stream>>>(hist, totals, n_dev, n_host, cap, nb_max);
        check_launch("radix_scan", L);
        begin_kernel("radix_scatter", L);
        radix_scatter_kernel<true><<<nb_max, RADIX_THREADS, 0, L.stream>>>(
            keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_dev, n_host, cap, shift, mask, hist, totals, nb_max);
        check_launch("radix_scatter", L);
        cur ^= 1;
        shift += pb;
    }
    if (final_buffer) *final_buffer = cur;
}

// Number of ping-pong flips the tile partition performs for a given tile count (needed by the
// backward to find the final point list without any saved host state).
int tile_partition_final_buffer(uint32_t n_tiles) {
    const int passes = (bits_for(n_tiles) + 7) / 8;
    return passes & 1;
}

// depth order of the Gaussians (ascending depth bits, ties by index; culled ones last), then the
// inclusive prefix sum of tiles touched in that order and R.
void launch_depth_order(const GeometryLayout& G, int32_t P, const Launch& L) {
    if (P <= 0) return;
    // preprocess already wrote sort_keys[0] (= depth_key) and sort_vals[0] (= 0..P-1)
    uint32_t* keys[2] = {G.sort_keys[0], G.sort_keys[1]};
    uint32_t* vals[2] = {G.sort_vals[0], G.sort_vals[1]};
    int fin = 0;
    radix_sort_pairs(keys, vals, nullptr, P, P, 0, 32, G.radix_hist, G.radix_totals, &fin, L);
    // 4 passes -> result is back in buffer 0
    const uint32_t* order = vals[fin];
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
    const uint32_t* order = G.sort_vals[0];   // 32-bit sort = 4 passes = even number of flips
    begin_kernel("emit_instances", L);
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, L.stream>>>(order, G.offsets, G.rect, P, gx, capacity,
                                                                  G.status, B.keys[0], B.vals[0]);
    check_launch("emit_instances", L);
    uint32_t* keys[2] = {B.keys[0], B.keys[1]};
    uint32_t* vals[2] = {B.vals[0], B.vals[1]};
    int fin = 0;
    radix_sort_pairs(keys, vals, G.status, 0, capacity, 0, bits_for(n_tiles), B.radix_hist, B.radix_totals, &fin, L);
    if (final_buffer) *final_buffer = fin;
    const int64_t blocks = (capacity + 255) / 256;
    if (blocks > 0) {
        begin_kernel("identify_ranges", L);
        identify_ranges_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(keys[fin], G.status, capacity, B.ranges);
        check_launch("identify_ranges", L);
    }
}

}  // namespace scgr
