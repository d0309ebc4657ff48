// collective.cu -- the path's single collective as a hand-written NVSwitch (NVLS) kernel (sm_100a).
//
// the gradients of the N views rendered by the N ranks are summed with ONE
// all-reduce over one flat fp32 buffer per step.  When that buffer lives in symmetric memory with a
// multicast mapping (torch.distributed._symmetric_memory; scgaussian_b200/parallel.py), the sum is
// done here in two shots through the switch instead of by NCCL:
//   shot 1  rank r pulls ITS 1/N of the buffer with multimem.ld_reduce: the switch reads the N
//           replicas and returns their sum (in-network reduction, fp32 accumulate);
//   shot 2  rank r pushes the reduced values with multimem.st: the switch writes them into all N
//           replicas.
// Every GPU sends and receives the buffer once; every replica ends up with bit-identical sums
// (one reduction per element, broadcast).  The caller brackets the kernel with cross-rank barriers
// (all replicas written before shot 1, all shards broadcast before anyone reads).
// HBM / NVLink-bound streaming of fp32: no tensor cores.
#include <cstdlib>

#include "common.cuh"

namespace scgr {

namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int NVLS_THREADS, int NVLS_UNROLL>
__device__ __forceinline__ void dense_shot(float4* __restrict__ mc, const size_t first, const size_t count) {
    const size_t stride = (size_t)gridDim.x * NVLS_THREADS;
    size_t i = (size_t)blockIdx.x * NVLS_THREADS + threadIdx.x;
    // NVLS_UNROLL independent reductions in flight per thread
    for (; i + (NVLS_UNROLL - 1) * stride < count; i += NVLS_UNROLL * stride) {
        float4 v[NVLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++) v[u] = multimem_ld_reduce_add(mc + first + i + u * stride);
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++) multimem_st(mc + first + i + u * stride, v[u]);
    }
    for (; i < count; i += stride) multimem_st(mc + first + i, multimem_ld_reduce_add(mc + first + i));
    // the broadcast must have LANDED in every replica before the cross-rank barrier that follows lets a peer read it:
    // the barrier's signal travels as a unicast store and may overtake multicast data still in the switch
    __threadfence_system();
}

template <int NVLS_THREADS, int NVLS_UNROLL>
__global__ void __launch_bounds__(NVLS_THREADS)
nvls_allreduce_kernel(float4* __restrict__ mc, const size_t first, const size_t count) {
    dense_shot<NVLS_THREADS, NVLS_UNROLL>(mc, first, count);
}

// ---- row-sparse variant: only the rows some rank actually wrote ----
// In one view most Gaussians receive no gradient at all (hidden behind saturated pixels, or outside the view): their
// rows of the gradient arrays are exact zeros on every rank.  The backward counts, per Gaussian, the views that gave
// it a gradient (ScgrGrads.live_count); once that small array has been summed over the ranks (dense shot, above) every
// rank knows the same set of rows worth reducing, and the 192-byte dL/dSH rows -- 79 % of the payload -- go through
// the switch only for those.  Rank r owns rows [r n / world, (r + 1) n / world).  A warp takes 128 consecutive rows,
// compacts the live ones with four ballots, and spreads the (row, float4) items of the live rows over its lanes,
// UNROLL items per lane in flight.
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t w, uint32_t k) {      // position of the k-th (0-based) set bit
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

template <int THREADS, int UNROLL>
__device__ __forceinline__ void rows_shot(float4* __restrict__ mc, const float* __restrict__ live, const long long row_first,
                                          const long long row_end, const int row_f4) {
    constexpr int CHUNK = 128;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * THREADS) >> 5;
    for (long long base = row_first + warp * CHUNK; base < row_end; base += n_warps * CHUNK) {
        uint32_t bal[4], pre[5];
        pre[0] = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const long long i = base + 32 * q + lane;
            bal[q] = __ballot_sync(0xffffffffu, i < row_end && live[i] != 0.f);
            pre[q + 1] = pre[q] + __popc(bal[q]);
        }
        const uint32_t items = pre[4] * (uint32_t)row_f4;
        for (uint32_t it0 = 0; it0 < items; it0 += 32u * UNROLL) {
            float4 v[UNROLL];
            float4* addr[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const uint32_t it = it0 + (uint32_t)u * 32u + (uint32_t)lane;
                addr[u] = nullptr;
                if (it < items) {
                    const uint32_t k = it / (uint32_t)row_f4, c = it - k * (uint32_t)row_f4;
                    const int q = (k >= pre[1]) + (k >= pre[2]) + (k >= pre[3]);
                    const uint32_t w = q == 0 ? bal[0] : (q == 1 ? bal[1] : (q == 2 ? bal[2] : bal[3]));
                    const uint32_t p0 = q == 0 ? pre[0] : (q == 1 ? pre[1] : (q == 2 ? pre[2] : pre[3]));
                    const long long row = base + 32 * q + (long long)nth_set_bit(w, k - p0);
                    addr[u] = mc + row * row_f4 + c;
                    v[u] = multimem_ld_reduce_add(addr[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (addr[u]) multimem_st(addr[u], v[u]);
        }
    }
    __threadfence_system();      // (as above)
}

template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
nvls_allreduce_rows_kernel(float4* __restrict__ mc, const float* __restrict__ live, const long long row_first,
                           const long long row_end, const int row_f4) {
    rows_shot<THREADS, UNROLL>(mc, live, row_first, row_end, row_f4);
}

// ---- the whole collective of a step in ONE launch: the cross-rank barriers that bracket the two shots are taken inside the
// kernel, on flag words in symmetric memory, instead of three host-issued barrier kernels between two launches ----
// Barrier s (a monotonically increasing epoch): rank r release-stores s into word r of every peer's flag array, then waits
// until every word of its own array has reached s.  Inside the grid CTA 0 does that once all CTAs have arrived (atomic
// counter) and releases the others through a `go` word.  Every wait is bounded (~10 s): on a timeout the kernel raises
// sync[2] and carries on (wrong sums, reported by the host, rather than a hung GPU).
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- the same two shots over plain peer-to-peer loads and stores (no multicast) ----
// With few ranks the switch's reduction engine is the slow way: at 2 ranks multimem.ld_reduce + multimem.st move the 60 MB
// of small blocks at a third of the link rate (0.19 ms; ncclAllReduce's direct path: 0.13 ms).  Here the owner of an
// element loads it from every replica through the peer mappings, adds the values in rank order (one reduction per
// element, in a fixed order: replicas stay bit-identical) and stores the sum into every replica.
template <int WORLD>
__device__ __forceinline__ float4 p2p_sum(float4* const* __restrict__ peer, const size_t i) {
    float4 v[WORLD];
#pragma unroll
    for (int q = 0; q < WORLD; q++) v[q] = __ldcg(peer[q] + i);      // (L2 of the owning GPU: never a stale L1 line)
    float4 a = v[0];
#pragma unroll
    for (int q = 1; q < WORLD; q++) { a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w; }
    return a;
}
template <int WORLD>
__device__ __forceinline__ void p2p_store(float4* const* __restrict__ peer, const size_t i, const float4 a) {
#pragma unroll
    for (int q = 0; q < WORLD; q++) __stcg(peer[q] + i, a);
}

template <int THREADS, int WORLD>
__device__ __forceinline__ void dense_shot_p2p(float4* const* __restrict__ peer, const size_t first, const size_t count) {
    const size_t stride = (size_t)gridDim.x * THREADS;
    constexpr int U = WORLD <= 2 ? 4 : 2;
    size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < count; i += U * stride) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; u++) a[u] = p2p_sum<WORLD>(peer, first + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) p2p_store<WORLD>(peer, first + i + u * stride, a[u]);
    }
    for (; i < count; i += stride) p2p_store<WORLD>(peer, first + i, p2p_sum<WORLD>(peer, first + i));
    __threadfence_system();
}

template <int THREADS, int WORLD>
__device__ __forceinline__ void rows_shot_p2p(float4* const* __restrict__ peer, const size_t rows_f4, const float* __restrict__ live,
                                              const long long row_first, const long long row_end, const int row_f4) {
    constexpr int CHUNK = 128;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * THREADS) >> 5;
    for (long long base = row_first + warp * CHUNK; base < row_end; base += n_warps * CHUNK) {
        uint32_t bal[4], pre[5];
        pre[0] = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const long long i = base + 32 * q + lane;
            bal[q] = __ballot_sync(0xffffffffu, i < row_end && live[i] != 0.f);
            pre[q + 1] = pre[q] + __popc(bal[q]);
        }
        const uint32_t items = pre[4] * (uint32_t)row_f4;
        for (uint32_t it = lane; it < items; it += 32u) {
            const uint32_t k = it / (uint32_t)row_f4, c = it - k * (uint32_t)row_f4;
            const int q = (k >= pre[1]) + (k >= pre[2]) + (k >= pre[3]);
            const uint32_t w = q == 0 ? bal[0] : (q == 1 ? bal[1] : (q == 2 ? bal[2] : bal[3]));
            const uint32_t p0 = q == 0 ? pre[0] : (q == 1 ? pre[1] : (q == 2 ? pre[2] : pre[3]));
            const long long row = base + 32 * q + (long long)nth_set_bit(w, k - p0);
            const size_t i = rows_f4 + (size_t)row * row_f4 + c;
            p2p_store<WORLD>(peer, i, p2p_sum<WORLD>(peer, i));
        }
    }
    __threadfence_system();
}

struct FusedArgs {
    float4* mc;
    size_t dense_first, dense_count;      // this rank's shard of the dense part, float4 units
    float4* mc_rows;
    const float* live;
    long long row_first, row_end;
    int row_f4;
    int rank, world;
    uint32_t* flags[SCGR_NVLS_MAX_WORLD];   // flags[q]: the flag array inside rank q's replica (world words)
    uint32_t* sync;                          // local: [0] arrivals, [1] go, [2] timeout raised
    uint32_t epoch;                          // first of the three barrier values of this call
    float4* peer[SCGR_NVLS_MAX_WORLD];       // peer-to-peer shots: rank q's mapping of the flat buffer (NULL: multicast shots)
    size_t rows_f4;                          // offset of the rows block inside the flat buffer, float4 units
    int p2p_dense, p2p_rows;                 // which shot goes peer to peer (measured: both at 2 ranks, the dense one at 4)
};

constexpr long long SPIN_LIMIT = 20000000000ll;     // cycles (~10 s): far beyond any legitimate skew between ranks

template <typename F>
__device__ __forceinline__ void bounded_spin(uint32_t* sync, F&& ready) {
    const long long t0 = clock64();
    while (!ready()) {
        if (clock64() - t0 > SPIN_LIMIT) { sync[2] = 1u; break; }
    }
}

// every thread of the grid calls this; `arrive`: the CTAs have work of the previous shot to finish first
__device__ __forceinline__ void fused_barrier(const FusedArgs& a, const uint32_t s, const bool arrive, const uint32_t arrivals) {
    __syncthreads();
    if (arrive && threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.sync, 1u);
    }
    if (blockIdx.x == 0) {
        if (arrive && threadIdx.x == 0) bounded_spin(a.sync, [&] { return ld_acquire_gpu(a.sync) >= arrivals; });
        __syncthreads();
        if ((int)threadIdx.x < a.world) {
            st_release_sys(a.flags[threadIdx.x] + a.rank, s);
            const uint32_t* mine = a.flags[a.rank] + threadIdx.x;
            bounded_spin(a.sync, [&] { return (int)(ld_acquire_sys(mine) - s) >= 0; });
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(a.sync + 1, s);
    } else {
        if (threadIdx.x == 0) bounded_spin(a.sync, [&] { return (int)(ld_acquire_gpu(a.sync + 1) - s) >= 0; });
        __syncthreads();
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
nvls_allreduce_fused_kernel(const FusedArgs a) {
    fused_barrier(a, a.epoch, false, 0u);                    // every replica has been written by its rank's backward
    if (a.dense_count) {
        if (!a.p2p_dense) dense_shot<THREADS, 4>(a.mc, a.dense_first, a.dense_count);
        else if (a.world == 2) dense_shot_p2p<THREADS, 2>(a.peer, a.dense_first, a.dense_count);
        else if (a.world == 4) dense_shot_p2p<THREADS, 4>(a.peer, a.dense_first, a.dense_count);
        else dense_shot_p2p<THREADS, 8>(a.peer, a.dense_first, a.dense_count);
    }
    fused_barrier(a, a.epoch + 1u, true, gridDim.x);         // the summed live counts are in place on every rank
    if (a.row_end > a.row_first) {
        if (!a.p2p_rows) rows_shot<THREADS, 4>(a.mc_rows, a.live, a.row_first, a.row_end, a.row_f4);
        else if (a.world == 2) rows_shot_p2p<THREADS, 2>(a.peer, a.rows_f4, a.live, a.row_first, a.row_end, a.row_f4);
        else if (a.world == 4) rows_shot_p2p<THREADS, 4>(a.peer, a.rows_f4, a.live, a.row_first, a.row_end, a.row_f4);
        else rows_shot_p2p<THREADS, 8>(a.peer, a.rows_f4, a.live, a.row_first, a.row_end, a.row_f4);
    }
    // every shard has been broadcast; the last arrival leaves the counter clean for the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.sync, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            bounded_spin(a.sync, [&] { return ld_acquire_gpu(a.sync) >= 2u * gridDim.x; });
            a.sync[0] = 0u;
        }
        __syncthreads();
        if ((int)threadIdx.x < a.world) {
            st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch + 2u);
            const uint32_t* mine = a.flags[a.rank] + threadIdx.x;
            bounded_spin(a.sync, [&] { return (int)(ld_acquire_sys(mine) - (a.epoch + 2u)) >= 0; });
        }
    }
}

}  // namespace

void launch_nvls_allreduce_rows(void* multicast_rows, const float* live_count, long long n_rows, int row_floats, int rank,
                                int world, const Launch& L) {
    if (n_rows <= 0) return;
    const long long per = (n_rows + world - 1) / world;
    const long long first = per * rank, end = first + per < n_rows ? first + per : n_rows;
    if (first >= end) return;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int ctas_per_sm = getenv("SCGR_NVLS_CTAS") ? atoi(getenv("SCGR_NVLS_CTAS")) : 2;
    constexpr int THREADS = 512;
    const long long chunks = (end - first + 127) / 128;
    const long long want = (chunks + THREADS / 32 - 1) / (THREADS / 32);
    const long long cap = (long long)ctas_per_sm * sms;
    begin_kernel("nvls_allreduce_rows", L);
    nvls_allreduce_rows_kernel<THREADS, 4><<<(unsigned)(want < cap ? want : cap), THREADS, 0, L.stream>>>(
        reinterpret_cast<float4*>(multicast_rows), live_count, first, end, row_floats / 4);
    check_launch("nvls_allreduce_rows", L);
}

void launch_nvls_allreduce(void* multicast_ptr, size_t n_floats, int rank, int world, const Launch& L) {
    const size_t n4 = n_floats / 4;                       // caller guarantees n_floats % (4 * world) == 0
    const size_t per = n4 / (size_t)world;
    if (per == 0) return;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int ctas_per_sm = getenv("SCGR_NVLS_CTAS") ? atoi(getenv("SCGR_NVLS_CTAS")) : 2;
    static const int variant = getenv("SCGR_NVLS_VARIANT") ? atoi(getenv("SCGR_NVLS_VARIANT")) : 0;
    float4* const mc = reinterpret_cast<float4*>(multicast_ptr);
    const size_t first = per * (size_t)rank;
    auto grid_for = [&](int threads, int unroll) {
        const size_t want = (per + (size_t)threads * unroll - 1) / ((size_t)threads * unroll);
        const size_t cap = (size_t)ctas_per_sm * sms;
        return (unsigned)(want < cap ? want : cap);
    };
    begin_kernel("nvls_allreduce", L);
    switch (variant) {
        case 1: nvls_allreduce_kernel<512, 8><<<grid_for(512, 8), 512, 0, L.stream>>>(mc, first, per); break;
        case 2: nvls_allreduce_kernel<1024, 4><<<grid_for(1024, 4), 1024, 0, L.stream>>>(mc, first, per); break;
        case 3: nvls_allreduce_kernel<256, 2><<<grid_for(256, 2), 256, 0, L.stream>>>(mc, first, per); break;
        default: nvls_allreduce_kernel<512, 4><<<grid_for(512, 4), 512, 0, L.stream>>>(mc, first, per); break;
    }
    check_launch("nvls_allreduce", L);
}

void launch_nvls_allreduce_fused(const ScgrNvlsFused& f, const Launch& L) {
    FusedArgs a{};
    const size_t n4 = f.dense_floats / 4;                  // caller guarantees dense_floats % (4 * world) == 0
    const size_t per = n4 / (size_t)f.world;
    a.mc = reinterpret_cast<float4*>(f.multicast_ptr);
    a.dense_first = per * (size_t)f.rank;
    a.dense_count = per;
    a.mc_rows = reinterpret_cast<float4*>(f.multicast_rows);
    a.live = f.live_count;
    if (f.multicast_rows && f.n_rows > 0) {
        const long long rper = (f.n_rows + f.world - 1) / f.world;
        a.row_first = rper * f.rank;
        a.row_end = a.row_first + rper < f.n_rows ? a.row_first + rper : f.n_rows;
        if (a.row_first > a.row_end) a.row_first = a.row_end;
    }
    a.row_f4 = f.row_floats / 4;
    a.rank = f.rank;
    a.world = f.world;
    for (int q = 0; q < f.world; q++) a.flags[q] = f.flags[q];
    // peer-to-peer shots when the caller passed the peer mappings and the world is one of the compiled sizes
    if (f.peer_ptrs[0] && (f.world == 2 || f.world == 4 || f.world == 8)) {
        for (int q = 0; q < f.world; q++) a.peer[q] = reinterpret_cast<float4*>(f.peer_ptrs[q]);
        a.rows_f4 = f.multicast_rows ? (size_t)((char*)f.multicast_rows - (char*)f.multicast_ptr) / 16 : 0;
        // tools/nvls_time.py on 8 x B200: peer-to-peer wins both shots at 2 ranks, the dense one only at 4, neither at 8
        static const int force = getenv("SCGR_NVLS_P2P_SHOTS") ? atoi(getenv("SCGR_NVLS_P2P_SHOTS")) : -1;      // bit 0 dense, bit 1 rows
        a.p2p_dense = force >= 0 ? (force & 1) : (f.world <= 4);
        a.p2p_rows = force >= 0 ? ((force >> 1) & 1) : (f.world <= 2);
    }
    a.sync = f.sync_local;
    a.epoch = f.epoch;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int ctas_per_sm = getenv("SCGR_NVLS_CTAS") ? atoi(getenv("SCGR_NVLS_CTAS")) : 2;
    constexpr int THREADS = 512;
    // the grid must be co-resident (CTAs wait for one another): never more CTAs than the device holds at once
    int max_per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, nvls_allreduce_fused_kernel<THREADS>, THREADS, 0);
    const int per_sm = max_per_sm < ctas_per_sm ? (max_per_sm > 0 ? max_per_sm : 1) : ctas_per_sm;
    begin_kernel("nvls_allreduce_fused", L);
    nvls_allreduce_fused_kernel<THREADS><<<(unsigned)(per_sm * sms), THREADS, 0, L.stream>>>(a);
    check_launch("nvls_allreduce_fused", L);
}

}  // namespace scgr
