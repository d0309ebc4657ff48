// render.cu -- per-tile alpha compositing, forward and backward (sm_100a).
//
// Replaces the external rasterizer's FORWARD::renderCUDA / BACKWARD::renderCUDA (SURVEY.md
// section 2c, section 8a rows a14/a15; semantics Appendix A.8 / A.9) behind
// reference gaussian_renderer/__init__.py:100-108.
//
// Design (differs from the reference's one-thread-per-pixel, 256-thread CTA):
//   * ONE WARP per 16x16 tile.  The tile is cut into 8 "slots" of 8x4 pixels; lane l owns pixel
//     (l % 8, l / 8) of every slot, i.e. 8 pixels per thread held in registers.
//   * Work items are dispatched longest first.  Forward: most tiles whole (8 slots per warp), then a
//     share of the tiles as two 16x8 halves, the last ones as four 16x4 quarters -- the tail of the grid
//     is made of short items, so no SM idles while a few warps finish whole tiles.  Backward: the
//     forward records how deep every tile's walk is, and the tiles are launched in descending order.
//     Images with few tiles are cut in halves / quarters throughout (parallelism over overhead).
//   * The tile's depth-ordered list is consumed 32 Gaussians at a time: lane l fetches the packed
//     48-byte record of the l-th one and computes, for its Gaussian, an exact
//     closed-form bound of the maximum of the Gaussian's exponent over each slot rectangle.
//     Three staging schemes, all double-buffered, chosen per kernel by measurement: three 16-byte `cp.async` per
//     lane straight into shared memory, completion per thread (wait_group) + a warp barrier -- no register holds
//     data in flight, which lets 18 warps per SM fit in the backward (the default of both kernels since round 2);
//     TMA, one `cp.async.bulk` per lane + an mbarrier (round 1's backward: a bulk copy takes its addresses from
//     uniform registers, so per-lane copies are issued one lane at a time, 256 instructions per batch); and
//     prefetching the next batch into registers with 3 x 128-bit loads (round 1's forward).
//     Slots whose bound says alpha < 1/255 everywhere are skipped without touching a pixel
//     (conservative: the per-pixel test that follows is the reference's, so no output changes).
//   * Backward: every lane accumulates its 10 per-Gaussian gradient terms over its 8 pixels in
//     registers; one 12-shuffle transposing butterfly per (tile, Gaussian) leaves each term summed
//     over all 256 pixels in one lane, and those 10 lanes issue one coalesced `red.global.add.f32`
//     -- instead of the reference's 10 global atomics per contributing (pixel, Gaussian) pair.
//   * exp() is one MUFU.EX2: the conic is stored pre-multiplied by -0.5*log2(e).
// No block-level barriers at all (a warp never waits for another); no tensor cores (blend is not a
// contraction).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace scgr {

namespace {


__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (pos < lc && power <= 0 && og_raw >= 1/255) ? og_raw : 0  as three chained SETP and one SELP
// (left to itself the compiler materialises each test with its own select)
__device__ __forceinline__ float gate_pair(const float og_raw, const float power, const int pos, const int lc) {
    float og;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.lt.s32 p, %2, %3;\n\t"
        "setp.le.and.f32 p, %4, 0f00000000, p;\n\t"
        "setp.ge.and.f32 p, %1, 0f3B808081, p;\n\t"      // 1/255
        "selp.f32 %0, %1, 0f00000000, p;\n\t"
        "}"
        : "=f"(og)
        : "f"(og_raw), "r"(pos), "r"(lc), "f"(power));
    return og;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32 (sm_100a FFMA2 / FMUL2 / FADD2): ONE instruction works on the two floats of a 64-bit register pair.
// A lane owns the same pixel position in the two 8-pixel-wide slot columns of a tile row, so everything the blend
// does per pixel is done for the pair at once; per-lane IEEE semantics are those of fmaf / * / + (round to nearest).
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// (TMA / mbarrier plumbing: common.cuh)
struct Rec {
    float4 q0, q1, q2;
};

// ---- optional work counters (tools/render_stats.py builds a separate library with -DSCGR_STATS;
// the product build compiles none of this) ----
#ifdef SCGR_STATS
__device__ unsigned long long d_render_stats[32];
#define SCGR_STAT_DECL(n_) uint32_t st_##n_ = 0u
#define SCGR_STAT_ADD(n_, v_) st_##n_ += (uint32_t)(v_)
#define SCGR_STAT_FLUSH_WARP(slot_, n_)                                                              \
    do {                                                                                             \
        if ((threadIdx.x & 31) == 0) atomicAdd(&d_render_stats[slot_], (unsigned long long)st_##n_); \
    } while (0)
#define SCGR_STAT_FLUSH_LANES(slot_, n_)                                                         \
    do {                                                                                         \
        const uint32_t t_ = __reduce_add_sync(0xffffffffu, st_##n_);                             \
        if ((threadIdx.x & 31) == 0) atomicAdd(&d_render_stats[slot_], (unsigned long long)t_); \
    } while (0)
#else
#define SCGR_STAT_DECL(n_)
#define SCGR_STAT_ADD(n_, v_)
#define SCGR_STAT_FLUSH_WARP(slot_, n_)
#define SCGR_STAT_FLUSH_LANES(slot_, n_)
#endif

__device__ __forceinline__ Rec load_rec(const Record* __restrict__ rec, uint32_t id) {
    const float4* p = reinterpret_cast<const float4*>(rec + id);
    Rec r;
    r.q0 = __ldg(p);
    r.q1 = __ldg(p + 1);
    r.q2 = __ldg(p + 2);
    return r;
}

// slot mask of one Gaussian over a region of SPW slots whose first pixel is (X0, Yr):
// bit i set <=> local slot i (column i & 1, row i >> 1) may receive a contribution
template <int SPW>
__device__ __forceinline__ uint32_t slot_mask(const Rec& r, const float X0, const float Yr, const uint32_t live_slots) {
    const CullParams c = make_cull(r.q0, r.q1);
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < SPW; i++) {
        const float x0 = X0 + (float)((i & 1) << 3), y0 = Yr + (float)((i >> 1) << 2);
        const float mp = max_power_over_rect(c.cA, c.cB, c.cC, c.kx, c.ky, c.mx, c.my, x0, x0 + 7.f, y0, y0 + 3.f);
        if (mp >= c.thr) m |= 1u << i;
    }
    return m & live_slots;
}

// Per-Gaussian values of a staged record, laid out for the packed blend: every scalar the two slot columns share
// appears twice, so that one 64-bit half of an LDS.128 is a ready-made {s, s} operand of FFMA2 / FMUL2.
//   d0 = {x, x, cA, cA}   d1 = {cB, cB, cC, cC}   d2 = {opacity, opacity, depth, depth}   d3 = {r, r, g, g}   d4 = {b, b, y, y}
constexpr int DUP_F4 = 5;
__device__ __forceinline__ void stage_dup(float4* s_dup, const int lane, const Rec& c) {
    float4* d = s_dup + lane * DUP_F4;      // 80-byte stride: conflict-free STS.128
    d[0] = make_float4(c.q0.x, c.q0.x, c.q0.z, c.q0.z);
    d[1] = make_float4(c.q0.w, c.q0.w, c.q1.x, c.q1.x);
    d[2] = make_float4(c.q1.y, c.q1.y, c.q1.z, c.q1.z);
    d[3] = make_float4(c.q2.x, c.q2.x, c.q2.y, c.q2.y);
    d[4] = make_float4(c.q2.z, c.q2.z, c.q0.y, c.q0.y);
}

// Work items of a render launch, in dispatch order: tiles [0, n8) whole, tiles [n8, n8 + n4) as two
// halves each, tiles [n8 + n4, n8 + n4 + n2) as four quarters each.
struct WorkSplit {
    int n8, n4, n2;
    int deep_first;     // backward only: the n2 / n4 tiles cut in quarters / halves are the DEEPEST of the longest-first order
    int items() const { return n8 + 2 * n4 + 4 * n2; }
};

// ------------------------------------------------------------------------------------------
// forward (A.8)
// ------------------------------------------------------------------------------------------
// One warp renders a region of SPW slots (SPW = 8: the whole tile, 4: a 16x8 half, 2: a 16x4 quarter)
// whose first slot is slot k0 of the tile.  Warps that share a tile never synchronise with each
// other (disjoint pixels, private staging buffers).
// TMA: 0 register-prefetched gathers, 1 one bulk copy per lane + mbarrier, 2 three cp.async per lane (see backward_region)
template <int SPW, int TMA>
__device__ __forceinline__ void
forward_region(const int tile, const int tiles_x, const int k0, float4 (*s_rec)[96], unsigned long long* bars,
               const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
               const Record* __restrict__ rec, const int W, const int H, const float* __restrict__ bg,
               float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
               uint32_t* __restrict__ n_contrib, float* __restrict__ final_T, uint4* __restrict__ tile_todo) {
    const int lane = threadIdx.x & 31;
    if (TMA == 1) {
        if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
        mbar_fence_init();
        __syncwarp();
    }
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = (tile % tiles_x) * TILE, Yr = (tile / tiles_x) * TILE + ((k0 >> 1) << 2);
    const uint2 range = ranges[tile];
    const int total = range.y > range.x ? (int)(range.y - range.x) : 0;   // empty tiles hold (0xffffffff, 0)

    // Tt = transmittance of the pixel while it is open; a finished pixel (saturated, or outside the
    // image) keeps its final transmittance NEGATED: test_T of a finished pixel is then negative, so
    // it can never blend again and no separate "done" flag is tested or updated in the inner loop.
    float Tt[SPW], Cr[SPW], Cg[SPW], Cb[SPW], Dd[SPW];
    uint32_t last[SPW];
#pragma unroll
    for (int i = 0; i < SPW; i++) {
        Cr[i] = 0.f; Cg[i] = 0.f; Cb[i] = 0.f; Dd[i] = 0.f; last[i] = 0u;
        const int px = X0 + ((i & 1) << 3) + lx, py = Yr + ((i >> 1) << 2) + ly;
        Tt[i] = (px >= W || py >= H) ? -1.f : 1.f;
    }
    const float pxf = (float)(X0 + lx), pyf = (float)(Yr + ly);

    // stage b of the list = entries [32 b, 32 b + 32); buffer b & 1
    auto issue = [&](const int b) {
        const int c = min(32, total - 32 * b);
        if (TMA == 1) {
            if (lane == 0) mbar_expect_tx(&bars[b & 1], 48u * (uint32_t)c);
            __syncwarp();
            if (lane < c)
                bulk_g2s(&s_rec[b & 1][lane * 3], rec + point_list[range.x + 32 * b + lane], 48u, &bars[b & 1]);
        } else {
            if (lane < c) {
                const float4* src = reinterpret_cast<const float4*>(rec + point_list[range.x + 32 * b + lane]);
                float4* dst = &s_rec[b & 1][lane * 3];
                cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
            }
            cp_async_commit();
        }
    };
    Rec nxt;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nbatch = (total + 31) >> 5;
    if (TMA) { if (nbatch > 0) issue(0); }
    else if (lane < total) nxt = load_rec(rec, point_list[range.x + lane]);

    SCGR_STAT_DECL(batches); SCGR_STAT_DECL(scanned); SCGR_STAT_DECL(hit); SCGR_STAT_DECL(slots);
    SCGR_STAT_DECL(cand); SCGR_STAT_DECL(go);
    int b = 0;
    for (; b < nbatch; b++) {
        const int base = 32 * b;
        // slots in which some pixel is still open
        uint32_t live = 0u;
#pragma unroll
        for (int i = 0; i < SPW; i++)
            if (__any_sync(0xffffffffu, Tt[i] > 0.f)) live |= 1u << i;
        if (live == 0u) break;
        const int cnt = min(32, total - base);
        const float4* const srec = s_rec[TMA ? (b & 1) : 0];
        Rec cur;
        if (TMA == 2) {
            if (b + 1 < nbatch) { issue(b + 1); cp_async_wait_all_but_one(); } else cp_async_wait_all();
            __syncwarp();
            cur.q0 = srec[lane * 3]; cur.q1 = srec[lane * 3 + 1]; cur.q2 = srec[lane * 3 + 2];
        } else if (TMA) {
            mbar_wait(&bars[b & 1], (uint32_t)(b >> 1) & 1u);
            __syncwarp();
            if (b + 1 < nbatch) issue(b + 1);        // the other buffer was released by the __syncwarp below
            cur.q0 = srec[lane * 3]; cur.q1 = srec[lane * 3 + 1]; cur.q2 = srec[lane * 3 + 2];
        } else {
            cur = nxt;
            __syncwarp();
            s_rec[0][lane * 3] = cur.q0; s_rec[0][lane * 3 + 1] = cur.q1; s_rec[0][lane * 3 + 2] = cur.q2;
            __syncwarp();
            if (base + 32 + lane < total) nxt = load_rec(rec, point_list[range.x + base + 32 + lane]);
        }
        const uint32_t mymask = lane < cnt ? slot_mask<SPW>(cur, (float)X0, (float)Yr, live) : 0u;
        SCGR_STAT_ADD(batches, 1); SCGR_STAT_ADD(scanned, cnt);

        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            SCGR_STAT_ADD(hit, 1); SCGR_STAT_ADD(slots, __popc(mj));
            const float4 q0 = srec[j * 3];
            const float4 q1 = srec[j * 3 + 1];
            const float4 q2 = srec[j * 3 + 2];
            // power(dx, dy) = cA dx^2 + dy (cB dx + cC dy): the dx-only terms are shared by the slots of
            // a column and hoisted, leaving 2 FFMA per pixel
            // (dx of the second column is x - (px + 8), dy of row r is dy of row r - 1 minus 4: the same roundings as the
            // packed kernels below, so that forward and backward see bit-identical exponents)
            const float dxa = q0.x - pxf, dxb = q0.x - (pxf + 8.f);
            const float ax[2] = {(q0.z * dxa) * dxa, (q0.z * dxb) * dxb};
            const float bx[2] = {q0.w * dxa, q0.w * dxb};
            const float dy0 = q0.y - pyf, dy1 = dy0 - 4.f, dy2_ = dy1 - 4.f;
            const float dys[4] = {dy0, dy1, dy2_, dy2_ - 4.f};
            const uint32_t idx = (uint32_t)(base + j + 1);
            auto blend = [&](const int i) {
                const float dy = dys[i >> 1];
                const float power = fmaf(dy, fmaf(q1.x, dy, bx[i & 1]), ax[i & 1]);
                // The reference's skip chain (A.8), written so that it compiles to predicated instructions
                // (no divergent branch, no selects: the compare/select pipe is the busy one in this loop).
                // (alpha >= 1/255 implies power >= pmin2 - margin: the slot bound needs no per-pixel twin.)
                const float araw = fminf(ALPHA_MAX, q1.y * ex2(power));
                const float w0 = araw * Tt[i];
                const float test_T = Tt[i] - w0;               // T (1 - alpha); negative for a finished pixel
                const bool cand = araw >= ALPHA_MIN && power <= 0.f;
                const bool go = cand && test_T >= T_EPS;       // open pixel, blended
                const bool sat = cand && test_T < T_EPS;       // saturates right here (NOT blended), or was finished
                if (go) {
                    Cr[i] = fmaf(q2.x, w0, Cr[i]); Cg[i] = fmaf(q2.y, w0, Cg[i]); Cb[i] = fmaf(q2.z, w0, Cb[i]);
                    Dd[i] = fmaf(q1.z, w0, Dd[i]);
                    Tt[i] = test_T;
                    last[i] = idx;
                }
                if (sat) Tt[i] = -fabsf(Tt[i]);                // finished pixels keep -|T|: test_T < 0 from now on
                SCGR_STAT_ADD(cand, cand ? 1 : 0); SCGR_STAT_ADD(go, go ? 1 : 0);
            };
#pragma unroll
            for (int i = 0; i < SPW; i++)
                if (mj & (1u << i)) blend(i);      // warp-uniform branch
        }
        __syncwarp();      // every lane is done with this stage before it is refilled
    }
    if (TMA == 1 && b < nbatch) mbar_wait(&bars[b & 1], (uint32_t)(b >> 1) & 1u);   // drain the copy in flight before exiting
    if (TMA == 2) cp_async_wait_all();
    SCGR_STAT_FLUSH_WARP(0, batches); SCGR_STAT_FLUSH_WARP(1, scanned); SCGR_STAT_FLUSH_WARP(2, hit);
    SCGR_STAT_FLUSH_WARP(3, slots); SCGR_STAT_FLUSH_LANES(4, cand); SCGR_STAT_FLUSH_LANES(5, go);
#ifdef SCGR_STATS
    if (lane == 0) { atomicAdd(&d_render_stats[6], 1ull); atomicAdd(&d_render_stats[7], (unsigned long long)total); }
#endif
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;
    {
        // depth of the backward's walk over this region (max n_contrib): the backward dispatches tiles
        // longest first.  Every work item of a tile owns its own words of the tile's uint4: no init pass.
        uint32_t deepest = 0u;
#pragma unroll
        for (int i = 0; i < SPW; i++) deepest = max(deepest, last[i]);
        deepest = __reduce_max_sync(0xffffffffu, deepest);
        if (lane == 0) {
            uint32_t* w = reinterpret_cast<uint32_t*>(tile_todo + tile);
            if (SPW == 8) tile_todo[tile] = make_uint4(deepest, 0u, 0u, 0u);
            else if (SPW == 4) { w[k0 >> 1] = deepest; w[(k0 >> 1) + 1] = 0u; }
            else w[k0 >> 1] = deepest;
        }
    }
#pragma unroll
    for (int i = 0; i < SPW; i++) {
        const int px = X0 + ((i & 1) << 3) + lx, py = Yr + ((i >> 1) << 2) + ly;
        if (px < W && py < H) {
            const size_t pid = (size_t)py * W + px;
            const float T = fabsf(Tt[i]);
            out_color[pid] = Cr[i] + T * bg0;
            out_color[N + pid] = Cg[i] + T * bg1;
            out_color[2 * N + pid] = Cb[i] + T * bg2;
            out_depth[pid] = Dd[i];
            out_alpha[pid] = 1.f - T;         // == sum alpha_i T_i (telescoping), A.8
            n_contrib[pid] = last[i];
            final_T[pid] = T;
        }
    }
}

// The same region walk with the blend done for the lane's two pixels of a tile row at once (packed fp32: FFMA2 /
// FMUL2 on register pairs, see fma2 above).  Per pixel the operations and their rounding are exactly those of
// forward_region: the two variants produce bit-identical images.
template <int SPW>
__device__ __forceinline__ void
forward_region_packed(const int tile, const int tiles_x, const int k0, float4* s_dup,
                      const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const Record* __restrict__ rec, const int W, const int H, const float* __restrict__ bg,
                      float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      uint32_t* __restrict__ n_contrib, float* __restrict__ final_T, uint4* __restrict__ tile_todo) {
    constexpr int NR = SPW / 2;
    const int lane = threadIdx.x & 31;
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = (tile % tiles_x) * TILE, Yr = (tile / tiles_x) * TILE + ((k0 >> 1) << 2);
    const uint2 range = ranges[tile];
    const int total = range.y > range.x ? (int)(range.y - range.x) : 0;   // empty tiles hold (0xffffffff, 0)
    // (sign convention of Tt as in forward_region: a finished pixel keeps its final transmittance negated)
    float2 Tt2[NR], Cr2[NR], Cg2[NR], Cb2[NR], Dd2[NR];
    uint32_t last[SPW];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        Cr2[r] = Cg2[r] = Cb2[r] = Dd2[r] = make_float2(0.f, 0.f);
        last[2 * r] = last[2 * r + 1] = 0u;
        const int py = Yr + (r << 2) + ly;
        Tt2[r] = make_float2((X0 + lx >= W || py >= H) ? -1.f : 1.f, (X0 + 8 + lx >= W || py >= H) ? -1.f : 1.f);
    }
    const float2 npx2 = make_float2(-(float)(X0 + lx), -(float)(X0 + lx + 8));
    const float2 npy2 = make_float2(-(float)(Yr + ly), -(float)(Yr + ly));
    const float2 kNegOne2 = make_float2(-1.f, -1.f), kNegFour2 = make_float2(-4.f, -4.f);
    Rec nxt;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nbatch = (total + 31) >> 5;
    if (lane < total) nxt = load_rec(rec, point_list[range.x + lane]);
    for (int b = 0; b < nbatch; b++) {
        const int base = 32 * b;
        uint32_t live = 0u;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            if (__any_sync(0xffffffffu, Tt2[r].x > 0.f)) live |= 1u << (2 * r);
            if (__any_sync(0xffffffffu, Tt2[r].y > 0.f)) live |= 2u << (2 * r);
        }
        if (live == 0u) break;
        const int cnt = min(32, total - base);
        const Rec cur = nxt;
        __syncwarp();
        stage_dup(s_dup, lane, cur);
        __syncwarp();
        if (base + 32 + lane < total) nxt = load_rec(rec, point_list[range.x + base + 32 + lane]);
        const uint32_t mymask = lane < cnt ? slot_mask<SPW>(cur, (float)X0, (float)Yr, live) : 0u;
        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            const float4* const dj = s_dup + j * DUP_F4;
            const float4 d0 = dj[0], d1 = dj[1], d2 = dj[2], d3 = dj[3], d4 = dj[4];
            const float2 dx2 = add2(make_float2(d0.x, d0.y), npx2);
            const float2 ax2 = mul2(mul2(make_float2(d0.z, d0.w), dx2), dx2);
            const float2 bx2 = mul2(make_float2(d1.x, d1.y), dx2);
            const float2 cC2 = make_float2(d1.z, d1.w), op2 = make_float2(d2.x, d2.y), dep2 = make_float2(d2.z, d2.w);
            const float2 qr2 = make_float2(d3.x, d3.y), qg2 = make_float2(d3.z, d3.w), qb2 = make_float2(d4.x, d4.y);
            float2 dy2 = add2(make_float2(d4.z, d4.w), npy2);
            const uint32_t idx = (uint32_t)(base + j + 1);
            auto blend_row = [&](const int r) {
                const float2 power = fma2(dy2, fma2(cC2, dy2, bx2), ax2);
                const float2 oe = mul2(op2, make_float2(ex2(power.x), ex2(power.y)));
                const float2 araw = make_float2(fminf(ALPHA_MAX, oe.x), fminf(ALPHA_MAX, oe.y));
                const float2 w0 = mul2(araw, Tt2[r]);
                const float2 test_T = fma2(w0, kNegOne2, Tt2[r]);      // T (1 - alpha); negative for a finished pixel
                // the reference's skip chain (A.8), per pixel: a slot the Gaussian cannot reach fails `cand` by itself
                const bool cand0 = araw.x >= ALPHA_MIN && power.x <= 0.f, cand1 = araw.y >= ALPHA_MIN && power.y <= 0.f;
                const bool go0 = cand0 && test_T.x >= T_EPS, go1 = cand1 && test_T.y >= T_EPS;
                const bool sat0 = cand0 && test_T.x < T_EPS, sat1 = cand1 && test_T.y < T_EPS;
                const float2 wm = make_float2(go0 ? w0.x : 0.f, go1 ? w0.y : 0.f);     // a masked weight adds exactly nothing
                Cr2[r] = fma2(qr2, wm, Cr2[r]); Cg2[r] = fma2(qg2, wm, Cg2[r]); Cb2[r] = fma2(qb2, wm, Cb2[r]);
                Dd2[r] = fma2(dep2, wm, Dd2[r]);
                if (go0) { Tt2[r].x = test_T.x; last[2 * r] = idx; }
                if (go1) { Tt2[r].y = test_T.y; last[2 * r + 1] = idx; }
                if (sat0) Tt2[r].x = -fabsf(Tt2[r].x);               // finished pixels keep -|T|: test_T < 0 from now on
                if (sat1) Tt2[r].y = -fabsf(Tt2[r].y);
            };
#pragma unroll
            for (int r = 0; r < NR; r++) {
                if ((mj >> (2 * r)) & 3u) blend_row(r);               // warp-uniform branch
                if (r + 1 < NR) dy2 = add2(dy2, kNegFour2);
            }
        }
        __syncwarp();      // every lane is done with this stage before it is refilled
    }
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;
    {
        uint32_t deepest = 0u;
#pragma unroll
        for (int i = 0; i < SPW; i++) deepest = max(deepest, last[i]);
        deepest = __reduce_max_sync(0xffffffffu, deepest);
        if (lane == 0) {
            uint32_t* w = reinterpret_cast<uint32_t*>(tile_todo + tile);
            if (SPW == 8) tile_todo[tile] = make_uint4(deepest, 0u, 0u, 0u);
            else if (SPW == 4) { w[k0 >> 1] = deepest; w[(k0 >> 1) + 1] = 0u; }
            else w[k0 >> 1] = deepest;
        }
    }
#pragma unroll
    for (int i = 0; i < SPW; i++) {
        const int r = i >> 1;
        const int px = X0 + ((i & 1) << 3) + lx, py = Yr + (r << 2) + ly;
        if (px < W && py < H) {
            const size_t pid = (size_t)py * W + px;
            const bool hi = i & 1;
            const float T = fabsf(hi ? Tt2[r].y : Tt2[r].x);
            out_color[pid] = (hi ? Cr2[r].y : Cr2[r].x) + T * bg0;
            out_color[N + pid] = (hi ? Cg2[r].y : Cg2[r].x) + T * bg1;
            out_color[2 * N + pid] = (hi ? Cb2[r].y : Cb2[r].x) + T * bg2;
            out_depth[pid] = hi ? Dd2[r].y : Dd2[r].x;
            out_alpha[pid] = 1.f - T;         // == sum alpha_i T_i (telescoping), A.8
            n_contrib[pid] = last[i];
            final_T[pid] = T;
        }
    }
}

template <int MINB, int MODE>      // MODE 0: register-prefetched gathers, 1: TMA staging, 2: packed-fp32 blend, 3: cp.async staging
__global__ void __launch_bounds__(32, MINB)
render_forward_kernel(const WorkSplit ws, const int tiles_x, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, const Record* __restrict__ rec, int W, int H,
                      const float* __restrict__ bg, const int64_t* __restrict__ status, int64_t capacity,
                      float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      uint32_t* __restrict__ n_contrib, float* __restrict__ final_T, uint4* __restrict__ tile_todo) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    constexpr int TMA = MODE == 1 ? 1 : (MODE == 3 ? 2 : 0);
    __shared__ __align__(16) float4 s_rec[MODE == 2 ? 1 : 2][MODE == 2 ? 32 * DUP_F4 : 96];   // record staging (MODE 2: packed-blend layout)
    __shared__ __align__(8) unsigned long long s_bar[2];
    int b = blockIdx.x;
    if (status[0] > capacity) {
        // Binning overflowed: the caller re-runs stage 2 with a larger buffer.  Until then the images hold NaN,
        // so a caller that missed the overflow (it is told by the return value / status word) cannot mistake
        // uninitialised memory for a rendering.
        int tile, k0, spw;
        if (b < ws.n8) { tile = b; k0 = 0; spw = 8; }
        else if ((b -= ws.n8) < 2 * ws.n4) { tile = ws.n8 + (b >> 1); k0 = (b & 1) << 2; spw = 4; }
        else { b -= 2 * ws.n4; tile = ws.n8 + ws.n4 + (b >> 2); k0 = (b & 3) << 1; spw = 2; }
        const int lane = threadIdx.x & 31, lx = lane & 7, ly = lane >> 3;
        const int X0 = (tile % tiles_x) * TILE, Yr = (tile / tiles_x) * TILE + ((k0 >> 1) << 2);
        const size_t N = (size_t)W * H;
        const float nan = __uint_as_float(0x7fc00000u);
        for (int i = 0; i < spw; i++) {
            const int px = X0 + ((i & 1) << 3) + lx, py = Yr + ((i >> 1) << 2) + ly;
            if (px < W && py < H) {
                const size_t pid = (size_t)py * W + px;
                out_color[pid] = nan; out_color[N + pid] = nan; out_color[2 * N + pid] = nan;
                out_depth[pid] = nan; out_alpha[pid] = nan;
            }
        }
        return;
    }
    if constexpr (MODE == 2) {
#define SCGR_ARGS &s_rec[0][0], ranges, point_list, rec, W, H, bg, out_color, out_depth, out_alpha, n_contrib, final_T, tile_todo
        if (b < ws.n8) { forward_region_packed<8>(b, tiles_x, 0, SCGR_ARGS); return; }
        b -= ws.n8;
        if (b < 2 * ws.n4) { forward_region_packed<4>(ws.n8 + (b >> 1), tiles_x, (b & 1) << 2, SCGR_ARGS); return; }
        b -= 2 * ws.n4;
        forward_region_packed<2>(ws.n8 + ws.n4 + (b >> 2), tiles_x, (b & 3) << 1, SCGR_ARGS);
#undef SCGR_ARGS
    } else {
        float4 (*s_rec2)[96] = reinterpret_cast<float4 (*)[96]>(&s_rec[0][0]);
#define SCGR_ARGS s_rec2, s_bar, ranges, point_list, rec, W, H, bg, out_color, out_depth, out_alpha, n_contrib, final_T, tile_todo
        if (b < ws.n8) { forward_region<8, TMA>(b, tiles_x, 0, SCGR_ARGS); return; }
        b -= ws.n8;
        if (b < 2 * ws.n4) { forward_region<4, TMA>(ws.n8 + (b >> 1), tiles_x, (b & 1) << 2, SCGR_ARGS); return; }
        b -= 2 * ws.n4;
        forward_region<2, TMA>(ws.n8 + ws.n4 + (b >> 2), tiles_x, (b & 3) << 1, SCGR_ARGS);
#undef SCGR_ARGS
    }
}

// ------------------------------------------------------------------------------------------
// backward (A.9)
// ------------------------------------------------------------------------------------------
// 10 per-lane partial sums -> lane (h, b8, b4, b2, 0) ends up holding the full 32-lane sum of ONE
// of them.  12 shuffles instead of 50.  Returns the value; `*slot` = float offset inside ScreenGrad
// (or -1 if this lane holds nothing).
// ScreenGrad float offset of the value lane `lane` holds after transpose_reduce10 (-1: none); lane-only,
// computed once per kernel
__device__ __forceinline__ int reduce10_slot(const int lane) {
    const bool h = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    int idx = -1;
    if (!(lane & 1)) {
        if (!b8) idx = b4 ? (b2 ? -1 : 2) : (b2 ? 1 : 0);
        else idx = b4 ? -1 : (b2 ? 4 : 3);
        if (idx >= 0 && h) idx += 5;
    }
    // value order {mx, my, A, B, C, opacity, depth, r, g, b} -> ScreenGrad float offsets
    return idx < 0 ? -1 : (idx < 7 ? idx : idx + 1);
}

__device__ __forceinline__ float transpose_reduce10(const float v[10], const int lane) {
    const bool h = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    float a[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float keep = h ? v[5 + i] : v[i];
        const float send = h ? v[i] : v[5 + i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float b[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float up = i < 2 ? a[3 + i] : 0.f;
        const float keep = b8 ? up : a[i];
        const float send = b8 ? a[i] : up;
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float c[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float up = i < 1 ? b[2] : 0.f;
        const float keep = b4 ? up : b[i];
        const float send = b4 ? b[i] : up;
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = b2 ? c[1] : c[0];
    const float send = b2 ? c[0] : c[1];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// TMA: how the records of a batch reach shared memory -- 0: register-prefetched gathers, 1: one bulk copy per lane +
// mbarrier, 2: three 16-byte cp.async per lane (one instruction serves the warp; a per-lane bulk copy is issued lane by
// lane, 256 instructions per batch).
template <int SPW, int TMA>
__device__ __forceinline__ void
backward_region(const int tile, const int tiles_x, const int k0, float4 (*s_rec)[96], uint32_t (*s_idb)[32],
                unsigned long long* bars, float4* s_g4, const uint2* __restrict__ ranges,
                const uint32_t* __restrict__ point_list, const Record* __restrict__ rec, const int W, const int H,
                const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                ScreenGrad* __restrict__ screen_grad) {
    const int lane = threadIdx.x & 31;
    if (TMA == 1) {
        if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
        mbar_fence_init();
        __syncwarp();
    }
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = (tile % tiles_x) * TILE, Yr = (tile / tiles_x) * TILE + ((k0 >> 1) << 2);
    const uint2 range = ranges[tile];
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;

    float T[SPW], Bs[SPW], tfb[SPW];
    int lc[SPW];
    int slot_lc[SPW];
    int toDo = 0;
#pragma unroll
    for (int i = 0; i < SPW; i++) {
        const int px = X0 + ((i & 1) << 3) + lx, py = Yr + ((i >> 1) << 2) + ly;
        float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f, Tf = 0.f;
        lc[i] = 0;
        if (px < W && py < H) {
            const size_t pid = (size_t)py * W + px;
            Tf = final_T[pid];
            lc[i] = (int)n_contrib[pid];
            gr = dL_dcolor[pid]; gg = dL_dcolor[N + pid]; gb = dL_dcolor[2 * N + pid];
            gd = dL_ddepth[pid];
            ga = dL_dalpha[pid];
        }
        s_g4[i * 32 + lane] = make_float4(gr, gg, gb, gd);
        T[i] = Tf;
        tfb[i] = Tf * (ga - (bg0 * gr + bg1 * gg + bg2 * gb));
        Bs[i] = 0.f;
        slot_lc[i] = __reduce_max_sync(0xffffffffu, lc[i]);
        toDo = max(toDo, slot_lc[i]);
    }
    const float pxf = (float)(X0 + lx), pyf = (float)(Yr + ly);
    // (each lane only ever reads back the s_g4 entries it wrote itself: no barrier needed)
    const int slot = reduce10_slot(lane);                         // which of the 10 sums this lane deposits
    float* const my_grad = reinterpret_cast<float*>(screen_grad) + (slot >= 0 ? slot : 0);

    // batch entry j  <->  0-based list position  pos = toDo - 1 - (base + j)   (back to front)
    // stage b = batch entries [32 b, 32 b + 32) of the back-to-front walk; buffer b & 1
    auto issue = [&](const int b) {
        const int c = min(32, toDo - 32 * b);
        if (TMA == 1) {
            if (lane == 0) mbar_expect_tx(&bars[b & 1], 48u * (uint32_t)c);
            __syncwarp();
        }
        if (lane < c) {
            const uint32_t id = point_list[range.x + (toDo - 1 - (32 * b + lane))];
            s_idb[b & 1][lane] = id;
            if (TMA == 1) {
                bulk_g2s(&s_rec[b & 1][lane * 3], rec + id, 48u, &bars[b & 1]);
            } else {
                const float4* src = reinterpret_cast<const float4*>(rec + id);
                float4* dst = &s_rec[b & 1][lane * 3];
                cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
            }
        }
        if (TMA == 2) cp_async_commit();
    };
    Rec nxt;
    uint32_t nxt_id = 0u;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nbatch = (toDo + 31) >> 5;
    if (TMA) { if (nbatch > 0) issue(0); }
    else if (lane < toDo) {
        nxt_id = point_list[range.x + (toDo - 1 - lane)];
        nxt = load_rec(rec, nxt_id);
    }
    SCGR_STAT_DECL(batches); SCGR_STAT_DECL(scanned); SCGR_STAT_DECL(hit); SCGR_STAT_DECL(slots);
    SCGR_STAT_DECL(ok); SCGR_STAT_DECL(red); SCGR_STAT_DECL(atom);
    for (int b = 0; b < nbatch; b++) {
        const int base = 32 * b;
        const int cnt = min(32, toDo - base);
        SCGR_STAT_ADD(batches, 1); SCGR_STAT_ADD(scanned, cnt);
        const float4* const srec = s_rec[TMA ? (b & 1) : 0];
        const uint32_t* const s_id = s_idb[TMA ? (b & 1) : 0];
        Rec cur;
        if (TMA == 2) {
            // the other buffer was released by the __syncwarp that ends the previous iteration: refill it, then wait for
            // this thread's copies of the current one; the warp barrier makes every lane's copies visible to all
            if (b + 1 < nbatch) { issue(b + 1); cp_async_wait_all_but_one(); } else cp_async_wait_all();
            __syncwarp();
            cur.q0 = srec[lane * 3]; cur.q1 = srec[lane * 3 + 1]; cur.q2 = srec[lane * 3 + 2];
        } else if (TMA) {
            mbar_wait(&bars[b & 1], (uint32_t)(b >> 1) & 1u);
            __syncwarp();                            // (also orders the generic s_id stores of this stage)
            if (b + 1 < nbatch) issue(b + 1);        // the other buffer was released by the __syncwarp below
            cur.q0 = srec[lane * 3]; cur.q1 = srec[lane * 3 + 1]; cur.q2 = srec[lane * 3 + 2];
        } else {
            cur = nxt;
            __syncwarp();
            s_rec[0][lane * 3] = cur.q0; s_rec[0][lane * 3 + 1] = cur.q1; s_rec[0][lane * 3 + 2] = cur.q2;
            s_idb[0][lane] = nxt_id;
            __syncwarp();
            if (base + 32 + lane < toDo) {
                nxt_id = point_list[range.x + (toDo - 1 - (base + 32 + lane))];
                nxt = load_rec(rec, nxt_id);
            }
        }
        uint32_t mymask = 0u;
        if (lane < cnt) {
            const int mypos = toDo - 1 - (base + lane);
            uint32_t live = 0u;
#pragma unroll
            for (int i = 0; i < SPW; i++)
                if (mypos < slot_lc[i]) live |= 1u << i;
            mymask = slot_mask<SPW>(cur, (float)X0, (float)Yr, live);
        }
        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            SCGR_STAT_ADD(hit, 1); SCGR_STAT_ADD(slots, __popc(mj));
            const int pos = toDo - 1 - (base + j);
            const float4 q0 = srec[j * 3];
            const float4 q1 = srec[j * 3 + 1];
            const float4 q2 = srec[j * 3 + 2];
            // hoisted per-column / per-row terms of the exponent, as in the forward
            const float dxa = q0.x - pxf, dxb = q0.x - (pxf + 8.f);      // (same roundings as the forward)
            const float dxs[2] = {dxa, dxb};
            const float ax[2] = {(q0.z * dxa) * dxa, (q0.z * dxb) * dxb};
            const float bx[2] = {q0.w * dxa, q0.w * dxb};
            const float dy0 = q0.y - pyf, dy1 = dy0 - 4.f, dy2_ = dy1 - 4.f;
            const float dys[4] = {dy0, dy1, dy2_, dy2_ - 4.f};
            float v[10];
#pragma unroll
            for (int i = 0; i < 10; i++) v[i] = 0.f;
            auto pair_grad = [&](const int i) {
                const float dx = dxs[i & 1], dy = dys[i >> 1];
                const float power = fmaf(dy, fmaf(q1.x, dy, bx[i & 1]), ax[i & 1]);
                // straight-line: a pair that fails the reference's tests (A.9) runs with og = alpha = 0, which
                // makes every term below vanish and leaves the pixel state untouched.
                // min(0.99, og) >= 1/255  <=>  og >= 1/255
                const float ograw = q1.y * ex2(power);
                const float og = gate_pair(ograw, power, pos, lc[i]);   // opacity * G, un-capped
                SCGR_STAT_ADD(ok, og > 0.f ? 1 : 0);
                const float alpha = fminf(ALPHA_MAX, og);
                const float ra = rcp_approx(1.f - alpha);    // 1 - alpha >= 0.01
                T[i] *= ra;                                  // transmittance in front of this Gaussian
                const float4 g4 = s_g4[i * 32 + lane];       // upstream dL/d{r, g, b, depth} of this pixel
                // Only the upstream-weighted sum over channels of the suffix blend is needed:
                //   Bs = sum_ch g_ch A_ch,  A_ch <- alpha c_ch + (1 - alpha) A_ch   =>   Bs <- Bs + alpha (g.c - Bs)
                const float cg = fmaf(q1.z, g4.w, fmaf(q2.z, g4.z, fmaf(q2.y, g4.y, q2.x * g4.x)));
                const float e = cg - Bs[i];
                Bs[i] = fmaf(alpha, e, Bs[i]);
                // the alpha output and the background enter as T_final / (1 - alpha) * (dL/dalpha_pix - bg . dL/dC)
                const float dL_dalpha_ = fmaf(e, T[i], tfb[i] * ra);
                const float w = alpha * T[i];
                const float uG = og * dL_dalpha_;            // G * dL/dG; propagated even when alpha was capped (A.9)
                const float ux = uG * dx, uy = uG * dy;
                v[0] += ux;                                  // preprocess-backward rebuilds dL/dmean from these two
                v[1] += uy;
                v[2] = fmaf(ux, dx, v[2]);                   // * -0.5 = dL/dconic_A
                v[3] = fmaf(ux, dy, v[3]);                   // * -1   = dL/dconic_B
                v[4] = fmaf(uy, dy, v[4]);                   // * -0.5 = dL/dconic_C
                v[5] += uG;                                  // / opacity = dL/dopacity
                v[6] = fmaf(w, g4.w, v[6]);                  // dL/ddepth
                v[7] = fmaf(w, g4.x, v[7]); v[8] = fmaf(w, g4.y, v[8]); v[9] = fmaf(w, g4.z, v[9]);
            };
#pragma unroll
            for (int i = 0; i < SPW; i++)
                if (mj & (1u << i)) pair_grad(i);  // warp-uniform branch
            // (97 % of the pairs that get here have a contributing pixel: reduce unconditionally)
            const float sum = transpose_reduce10(v, lane);
            SCGR_STAT_ADD(red, 1); SCGR_STAT_ADD(atom, (slot >= 0 && sum != 0.f) ? 1 : 0);
            if (slot >= 0 && sum != 0.f)
                atomicAdd(my_grad + (size_t)s_id[j] * (sizeof(ScreenGrad) / sizeof(float)), sum);   // RED.E.ADD.F32
        }
        __syncwarp();      // every lane is done with this stage before it is refilled
    }
    SCGR_STAT_FLUSH_WARP(8, batches); SCGR_STAT_FLUSH_WARP(9, scanned); SCGR_STAT_FLUSH_WARP(10, hit);
    SCGR_STAT_FLUSH_WARP(11, slots); SCGR_STAT_FLUSH_LANES(12, ok); SCGR_STAT_FLUSH_WARP(13, red);
    SCGR_STAT_FLUSH_LANES(14, atom);
#ifdef SCGR_STATS
    if (lane == 0) { atomicAdd(&d_render_stats[15], 1ull); atomicAdd(&d_render_stats[16], (unsigned long long)toDo); }
#endif
}

// The same walk with the blend done for the lane's two pixels of a tile row at once (packed fp32, see fma2 above).
// MEASURED SLOWER than the scalar blend on B200 (profiles/r02_packed_fp32.md): kept as a switch (SCGR_BWD_PACKED=1).
template <int SPW, bool TMA>
__device__ __forceinline__ void
backward_region_packed(const int tile, const int tiles_x, const int k0, float4 (*s_rec)[96], float4* s_dup, uint32_t (*s_idb)[32],
                unsigned long long* bars, float4* s_g4, const uint2* __restrict__ ranges,
                const uint32_t* __restrict__ point_list, const Record* __restrict__ rec, const int W, const int H,
                const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                ScreenGrad* __restrict__ screen_grad) {
    constexpr int NR = SPW / 2;         // tile rows of this region; a lane owns one pixel in each of the row's two columns
    const int lane = threadIdx.x & 31;
    if (TMA) {
        if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
        mbar_fence_init();
        __syncwarp();
    }
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = (tile % tiles_x) * TILE, Yr = (tile / tiles_x) * TILE + ((k0 >> 1) << 2);
    const uint2 range = ranges[tile];
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;

    // pixel state, packed over the two columns of a row: transmittance T, suffix blend Bs, background / alpha term tfb
    float2 T2[NR], Bs2[NR], tfb2[NR];
    int lc[SPW];
    int slot_lc[SPW];
    int toDo = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        float gr[2], gg[2], gb[2], gd[2], tf[2], tb[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int i = 2 * r + c;
            const int px = X0 + (c << 3) + lx, py = Yr + (r << 2) + ly;
            float ga = 0.f;
            gr[c] = gg[c] = gb[c] = gd[c] = tf[c] = 0.f;
            lc[i] = 0;
            if (px < W && py < H) {
                const size_t pid = (size_t)py * W + px;
                tf[c] = final_T[pid];
                lc[i] = (int)n_contrib[pid];
                gr[c] = dL_dcolor[pid]; gg[c] = dL_dcolor[N + pid]; gb[c] = dL_dcolor[2 * N + pid];
                gd[c] = dL_ddepth[pid];
                ga = dL_dalpha[pid];
            }
            tb[c] = tf[c] * (ga - (bg0 * gr[c] + bg1 * gg[c] + bg2 * gb[c]));
            slot_lc[i] = __reduce_max_sync(0xffffffffu, lc[i]);
            toDo = max(toDo, slot_lc[i]);
        }
        // upstream dL/d{r, g, b, depth} of the lane's two pixels of this row, column-interleaved
        s_g4[(2 * r) * 32 + lane] = make_float4(gr[0], gr[1], gg[0], gg[1]);
        s_g4[(2 * r + 1) * 32 + lane] = make_float4(gb[0], gb[1], gd[0], gd[1]);
        T2[r] = make_float2(tf[0], tf[1]);
        tfb2[r] = make_float2(tb[0], tb[1]);
        Bs2[r] = make_float2(0.f, 0.f);
    }
    // (each lane only ever reads back the s_g4 entries it wrote itself: no barrier needed)
    const float2 npx2 = make_float2(-(float)(X0 + lx), -(float)(X0 + lx + 8));    // pixel abscissae of the two columns
    const float2 npy2 = make_float2(-(float)(Yr + ly), -(float)(Yr + ly));
    const float2 kOne2 = make_float2(1.f, 1.f), kNegOne2 = make_float2(-1.f, -1.f), kNegFour2 = make_float2(-4.f, -4.f);
    const int slot = reduce10_slot(lane);                         // which of the 10 sums this lane deposits
    float* const my_grad = reinterpret_cast<float*>(screen_grad) + (slot >= 0 ? slot : 0);

    // batch entry j  <->  0-based list position  pos = toDo - 1 - (base + j)   (back to front)
    // stage b = batch entries [32 b, 32 b + 32) of the back-to-front walk; buffer b & 1
    auto issue = [&](const int b) {
        const int c = min(32, toDo - 32 * b);
        if (lane == 0) mbar_expect_tx(&bars[b & 1], 48u * (uint32_t)c);
        __syncwarp();
        if (lane < c) {
            const uint32_t id = point_list[range.x + (toDo - 1 - (32 * b + lane))];
            s_idb[b & 1][lane] = id;
            bulk_g2s(&s_rec[b & 1][lane * 3], rec + id, 48u, &bars[b & 1]);
        }
    };
    Rec nxt;
    uint32_t nxt_id = 0u;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nbatch = (toDo + 31) >> 5;
    if (TMA) { if (nbatch > 0) issue(0); }
    else if (lane < toDo) {
        nxt_id = point_list[range.x + (toDo - 1 - lane)];
        nxt = load_rec(rec, nxt_id);
    }
    SCGR_STAT_DECL(batches); SCGR_STAT_DECL(scanned); SCGR_STAT_DECL(hit); SCGR_STAT_DECL(slots);
    SCGR_STAT_DECL(ok); SCGR_STAT_DECL(red); SCGR_STAT_DECL(atom);
    for (int b = 0; b < nbatch; b++) {
        const int base = 32 * b;
        const int cnt = min(32, toDo - base);
        SCGR_STAT_ADD(batches, 1); SCGR_STAT_ADD(scanned, cnt);
        const uint32_t* const s_id = s_idb[TMA ? (b & 1) : 0];
        Rec cur;
        if (TMA) {
            mbar_wait(&bars[b & 1], (uint32_t)(b >> 1) & 1u);
            __syncwarp();                            // (also orders the generic s_id stores of this stage)
            if (b + 1 < nbatch) issue(b + 1);        // the other buffer was released by the __syncwarp below
            const float4* const srec = s_rec[b & 1];
            cur.q0 = srec[lane * 3]; cur.q1 = srec[lane * 3 + 1]; cur.q2 = srec[lane * 3 + 2];
        } else {
            cur = nxt;
            s_idb[0][lane] = nxt_id;
            if (base + 32 + lane < toDo) {
                nxt_id = point_list[range.x + (toDo - 1 - (base + 32 + lane))];
                nxt = load_rec(rec, nxt_id);
            }
        }
        stage_dup(s_dup, lane, cur);                 // (the previous batch's readers passed the __syncwarp at the loop's end)
        __syncwarp();
        uint32_t mymask = 0u;
        if (lane < cnt) {
            const int mypos = toDo - 1 - (base + lane);
            uint32_t live = 0u;
#pragma unroll
            for (int i = 0; i < SPW; i++)
                if (mypos < slot_lc[i]) live |= 1u << i;
            mymask = slot_mask<SPW>(cur, (float)X0, (float)Yr, live);
        }
        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            SCGR_STAT_ADD(hit, 1); SCGR_STAT_ADD(slots, __popc(mj));
            const int pos = toDo - 1 - (base + j);
            const float4* const dj = s_dup + j * DUP_F4;
            const float4 d0 = dj[0], d1 = dj[1], d2 = dj[2], d3 = dj[3], d4 = dj[4];
            // per-column terms of the exponent, hoisted out of the rows:  power = cA dx^2 + dy (cB dx + cC dy)
            const float2 dx2 = add2(make_float2(d0.x, d0.y), npx2);
            const float2 ax2 = mul2(mul2(make_float2(d0.z, d0.w), dx2), dx2);
            const float2 bx2 = mul2(make_float2(d1.x, d1.y), dx2);
            const float2 cC2 = make_float2(d1.z, d1.w), op2 = make_float2(d2.x, d2.y), dep2 = make_float2(d2.z, d2.w);
            const float2 qr2 = make_float2(d3.x, d3.y), qg2 = make_float2(d3.z, d3.w), qb2 = make_float2(d4.x, d4.y);
            float2 dy2 = add2(make_float2(d4.z, d4.w), npy2);
            float2 v2[10];
#pragma unroll
            for (int i = 0; i < 10; i++) v2[i] = make_float2(0.f, 0.f);
            auto row_grad = [&](const int r) {
                const float2 power = fma2(dy2, fma2(cC2, dy2, bx2), ax2);
                // straight-line: a pixel that fails the reference's tests (A.9) -- or whose slot the Gaussian cannot
                // reach at all -- runs with og = alpha = 0, which makes every term below vanish and leaves its state
                // untouched.     min(0.99, og) >= 1/255  <=>  og >= 1/255
                const float2 ograw = mul2(op2, make_float2(ex2(power.x), ex2(power.y)));
                float2 og;                                              // opacity * G, un-capped
                og.x = gate_pair(ograw.x, power.x, pos, lc[2 * r]);
                og.y = gate_pair(ograw.y, power.y, pos, lc[2 * r + 1]);
                SCGR_STAT_ADD(ok, (og.x > 0.f ? 1 : 0) + (og.y > 0.f ? 1 : 0));
                const float2 alpha = make_float2(fminf(ALPHA_MAX, og.x), fminf(ALPHA_MAX, og.y));
                const float2 om = fma2(alpha, kNegOne2, kOne2);          // 1 - alpha >= 0.01
                const float2 ra = make_float2(rcp_approx(om.x), rcp_approx(om.y));
                T2[r] = mul2(T2[r], ra);                                 // transmittance in front of this Gaussian
                const float4 gA = s_g4[(2 * r) * 32 + lane], gB = s_g4[(2 * r + 1) * 32 + lane];
                const float2 gr2 = make_float2(gA.x, gA.y), gg2 = make_float2(gA.z, gA.w);
                const float2 gb2 = make_float2(gB.x, gB.y), gd2 = make_float2(gB.z, gB.w);
                // Only the upstream-weighted sum over channels of the suffix blend is needed:
                //   Bs = sum_ch g_ch A_ch,  A_ch <- alpha c_ch + (1 - alpha) A_ch   =>   Bs <- Bs + alpha (g.c - Bs)
                const float2 cg = fma2(dep2, gd2, fma2(qb2, gb2, fma2(qg2, gg2, mul2(qr2, gr2))));
                const float2 e = fma2(Bs2[r], kNegOne2, cg);
                Bs2[r] = fma2(alpha, e, Bs2[r]);
                // the alpha output and the background enter as T_final / (1 - alpha) * (dL/dalpha_pix - bg . dL/dC)
                const float2 dL_dalpha_ = fma2(e, T2[r], mul2(tfb2[r], ra));
                const float2 w = mul2(alpha, T2[r]);
                const float2 uG = mul2(og, dL_dalpha_);                  // G * dL/dG; propagated even when alpha was capped (A.9)
                const float2 ux = mul2(uG, dx2), uy = mul2(uG, dy2);
                v2[0] = add2(v2[0], ux);                                 // preprocess-backward rebuilds dL/dmean from these two
                v2[1] = add2(v2[1], uy);
                v2[2] = fma2(ux, dx2, v2[2]);                            // * -0.5 = dL/dconic_A
                v2[3] = fma2(ux, dy2, v2[3]);                            // * -1   = dL/dconic_B
                v2[4] = fma2(uy, dy2, v2[4]);                            // * -0.5 = dL/dconic_C
                v2[5] = add2(v2[5], uG);                                 // / opacity = dL/dopacity
                v2[6] = fma2(w, gd2, v2[6]);                             // dL/ddepth
                v2[7] = fma2(w, gr2, v2[7]); v2[8] = fma2(w, gg2, v2[8]); v2[9] = fma2(w, gb2, v2[9]);
            };
#pragma unroll
            for (int r = 0; r < NR; r++) {
                if ((mj >> (2 * r)) & 3u) row_grad(r);                  // warp-uniform branch
                if (r + 1 < NR) dy2 = add2(dy2, kNegFour2);
            }
            float v[10];
#pragma unroll
            for (int i = 0; i < 10; i++) v[i] = v2[i].x + v2[i].y;
            // (97 % of the pairs that get here have a contributing pixel: reduce unconditionally)
            const float sum = transpose_reduce10(v, lane);
            SCGR_STAT_ADD(red, 1); SCGR_STAT_ADD(atom, (slot >= 0 && sum != 0.f) ? 1 : 0);
            if (slot >= 0 && sum != 0.f)
                atomicAdd(my_grad + (size_t)s_id[j] * (sizeof(ScreenGrad) / sizeof(float)), sum);   // RED.E.ADD.F32
        }
        __syncwarp();      // every lane is done with this stage before it is refilled
    }
    SCGR_STAT_FLUSH_WARP(8, batches); SCGR_STAT_FLUSH_WARP(9, scanned); SCGR_STAT_FLUSH_WARP(10, hit);
    SCGR_STAT_FLUSH_WARP(11, slots); SCGR_STAT_FLUSH_LANES(12, ok); SCGR_STAT_FLUSH_WARP(13, red);
    SCGR_STAT_FLUSH_LANES(14, atom);
#ifdef SCGR_STATS
    if (lane == 0) { atomicAdd(&d_render_stats[15], 1ull); atomicAdd(&d_render_stats[16], (unsigned long long)toDo); }
#endif
}

template <int MINB, int TMA, bool PK>
__global__ void __launch_bounds__(32, MINB)
render_backward_kernel(const WorkSplit ws, const int tiles_x, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, const Record* __restrict__ rec, int W, int H,
                       const float* __restrict__ bg, const int64_t* __restrict__ status, int64_t capacity,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T,
                       const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                       const float* __restrict__ dL_dalpha, ScreenGrad* __restrict__ screen_grad,
                       const uint32_t* __restrict__ tile_order) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    __shared__ __align__(16) float4 s_rec[2][96];        // 2 stages x 32 records x {q0, q1, q2}
    __shared__ __align__(16) float4 s_dup[PK ? 32 * DUP_F4 : 1];  // packed blend: the batch's records in its layout (stage_dup)
    __shared__ uint32_t s_id[2][32];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ float4 s_g4[TILE_PIX];       // upstream gradients of the region, column-interleaved per tile row
    if (status[0] > capacity) return;
    int b = blockIdx.x;
    // work item -> tile through the longest-first permutation built from the forward's per-tile depths
    // work item -> (position in the longest-first order, first slot, slots): whole tiles first and the cut ones last, or
    // (deep_first) the deepest tiles cut in quarters / halves first: their walks are the critical path of the launch
    int ord, k0, spw;
    if (!ws.deep_first) {
        if (b < ws.n8) { ord = b; k0 = 0; spw = 8; }
        else if ((b -= ws.n8) < 2 * ws.n4) { ord = ws.n8 + (b >> 1); k0 = (b & 1) << 2; spw = 4; }
        else { b -= 2 * ws.n4; ord = ws.n8 + ws.n4 + (b >> 2); k0 = (b & 3) << 1; spw = 2; }
    } else {
        if (b < 4 * ws.n2) { ord = b >> 2; k0 = (b & 3) << 1; spw = 2; }
        else if ((b -= 4 * ws.n2) < 2 * ws.n4) { ord = ws.n2 + (b >> 1); k0 = (b & 1) << 2; spw = 4; }
        else { b -= 2 * ws.n4; ord = ws.n2 + ws.n4 + b; k0 = 0; spw = 8; }
    }
    const int tile = (int)tile_order[ord];
    if constexpr (PK) {
#define SCGR_ARGS s_rec, s_dup, s_id, s_bar, s_g4, ranges, point_list, rec, W, H, bg, n_contrib, final_T, dL_dcolor, dL_ddepth, \
                  dL_dalpha, screen_grad
        if (spw == 8) backward_region_packed<8, TMA != 0>(tile, tiles_x, k0, SCGR_ARGS);
        else if (spw == 4) backward_region_packed<4, TMA != 0>(tile, tiles_x, k0, SCGR_ARGS);
        else backward_region_packed<2, TMA != 0>(tile, tiles_x, k0, SCGR_ARGS);
#undef SCGR_ARGS
    } else {
#define SCGR_ARGS s_rec, s_id, s_bar, s_g4, ranges, point_list, rec, W, H, bg, n_contrib, final_T, dL_dcolor, dL_ddepth, \
                  dL_dalpha, screen_grad
        if (spw == 8) backward_region<8, TMA>(tile, tiles_x, k0, SCGR_ARGS);
        else if (spw == 4) backward_region<4, TMA>(tile, tiles_x, k0, SCGR_ARGS);
        else backward_region<2, TMA>(tile, tiles_x, k0, SCGR_ARGS);
#undef SCGR_ARGS
    }
}

// Backward prologue, one launch: CTA 0 orders the tiles by descending depth of their backward
// walk (counting sort on 1024 buckets: the grid is dispatched in block order, so the long tiles
// start first and the tail of the launch is made of short ones -- longest-processing-time-first;
// order inside a bucket is arbitrary); all other CTAs zero the per-Gaussian gradient accumulators.
constexpr int ORDER_THREADS = 1024;
constexpr int ORDER_BINS = 1024;
constexpr int ORDER_CACHE = 8192;          // tile depths kept in shared memory (larger images re-read them)
constexpr int ZERO_F4_PER_CTA = 4096;      // 64 KB of zeros per CTA
__global__ void __launch_bounds__(ORDER_THREADS)
backward_prologue_kernel(const uint4* __restrict__ tile_todo, const int n_tiles, uint32_t* __restrict__ tile_order,
                         float4* __restrict__ zero_dst, const size_t zero_f4) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    if (blockIdx.x > 0) {
        const size_t base = (size_t)(blockIdx.x - 1) * ZERO_F4_PER_CTA;
#pragma unroll
        for (int k = 0; k < ZERO_F4_PER_CTA / ORDER_THREADS; k++) {
            const size_t i = base + (size_t)k * ORDER_THREADS + threadIdx.x;
            if (i < zero_f4) zero_dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    __shared__ uint32_t s_bin[ORDER_BINS];
    __shared__ uint32_t s_key[ORDER_CACHE];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_max;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    auto depth_of = [&](const int i) {
        const uint4 d = tile_todo[i];
        return max(max(d.x, d.y), max(d.z, d.w));
    };
    uint32_t mx = 0u;
    for (int i = t; i < n_tiles; i += ORDER_THREADS) {
        const uint32_t k = depth_of(i);
        if (i < ORDER_CACHE) s_key[i] = k;
        mx = max(mx, k);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) s_warp[w] = mx;
    s_bin[t] = 0u;
    __syncthreads();
    if (w == 0) {
        const uint32_t m = __reduce_max_sync(0xffffffffu, s_warp[lane]);
        if (lane == 0) s_max = m;
    }
    __syncthreads();
    int shift = 0;
    while ((s_max >> shift) >= (uint32_t)ORDER_BINS) shift++;
    // bucket 0 = deepest
    auto bucket = [&](const int i) { return (ORDER_BINS - 1) - (int)((i < ORDER_CACHE ? s_key[i] : depth_of(i)) >> shift); };
    for (int i = t; i < n_tiles; i += ORDER_THREADS) atomicAdd(&s_bin[bucket(i)], 1u);
    __syncthreads();
    // exclusive scan of the 1024 buckets (one per thread)
    const uint32_t c = s_bin[t];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = s_warp[lane], xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += n;
        }
        s_warp[lane] = xi - x;
    }
    __syncthreads();
    s_bin[t] = s_warp[w] + inc - c;
    __syncthreads();
    for (int i = t; i < n_tiles; i += ORDER_THREADS) tile_order[atomicAdd(&s_bin[bucket(i)], 1u)] = (uint32_t)i;
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// tiles rendered whole / in halves / in quarters: SCGR_SPLIT="h,q" = percent of the tiles (by index,
// from the end) cut in halves and in quarters
// Small images do not have enough tiles to fill 148 SMs with one warp each (504x378, the reference's
// own LLFF resolution, has 768): below ~16 warps per SM every tile is cut in halves, below ~8 in quarters,
// trading per-(tile, Gaussian) overhead for parallelism in what is then a latency-bound launch.
WorkSplit make_split(const int n_tiles, const char* env, const int dflt_half, const int dflt_quarter) {
    int ph = dflt_half, pq = dflt_quarter;
    static const int sm_count = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    if (n_tiles < 8 * sm_count) { ph = 0; pq = 100; }
    else if (n_tiles < 16 * sm_count) { ph = 100; pq = 0; }
    if (const char* e = getenv(env)) sscanf(e, "%d,%d", &ph, &pq);
    WorkSplit ws;
    ws.deep_first = 0;
    ws.n2 = (int)((int64_t)n_tiles * pq / 100);
    ws.n4 = (int)((int64_t)n_tiles * ph / 100);
    if (ws.n2 + ws.n4 > n_tiles) { ws.n2 = 0; ws.n4 = n_tiles; }
    ws.n8 = n_tiles - ws.n4 - ws.n2;
    return ws;
}

}  // namespace

void launch_render_forward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                           const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                           float* out_color, float* out_depth, float* out_alpha, const Launch& L) {
    const int tx = (v.image_width + TILE - 1) / TILE, ty = (v.image_height + TILE - 1) / TILE;
    if (tx == 0 || ty == 0) return;
    static const int minb = env_int("SCGR_FWD_MINB", 20);
    // staging / blend variant: 0 register-prefetched gathers + scalar blend, 1 TMA staging + scalar blend, 2 packed-fp32 blend,
    // 3 cp.async staging + scalar blend
    static const int tma = env_int("SCGR_FWD_PACKED", 0) ? 2 : env_int("SCGR_TMA_FWD", env_int("SCGR_TMA", 3));
    const WorkSplit ws = make_split(tx * ty, "SCGR_FWD_SPLIT", 10, 5);
    begin_kernel("render_forward", L);
#define SCGR_FWD(M_, T_) chain(render_forward_kernel<M_, T_>, dim3(ws.items()), dim3(32), 0, L)(ws, tx, B.ranges, point_list, G.rec, \
        v.image_width, v.image_height, v.bg, G.status, capacity, out_color, out_depth, out_alpha, I.n_contrib, I.final_T, \
        I.tile_todo)
    if (tma == 0) { if (minb == 20) SCGR_FWD(20, 0); else if (minb == 24) SCGR_FWD(24, 0); else SCGR_FWD(1, 0); }
    else if (tma == 1) { if (minb == 20) SCGR_FWD(20, 1); else if (minb == 24) SCGR_FWD(24, 1); else SCGR_FWD(1, 1); }
    else if (tma == 3) {      // cp.async staging
        if (minb == 20) SCGR_FWD(20, 3); else if (minb == 24) SCGR_FWD(24, 3); else if (minb == 22) SCGR_FWD(22, 3);
        else if (minb == 26) SCGR_FWD(26, 3); else if (minb == 28) SCGR_FWD(28, 3); else SCGR_FWD(1, 3);
    }
    else if (minb == 20) SCGR_FWD(20, 2); else if (minb == 24) SCGR_FWD(24, 2); else if (minb == 16) SCGR_FWD(16, 2); else if (minb == 18) SCGR_FWD(18, 2); else SCGR_FWD(1, 2);
#undef SCGR_FWD
    check_launch("render_forward", L);
}

void launch_render_backward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                            const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                            const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                            int32_t P, const Launch& L) {
    const int tx = (v.image_width + TILE - 1) / TILE, ty = (v.image_height + TILE - 1) / TILE;
    if (tx == 0 || ty == 0) {
        cudaMemsetAsync(G.screen_grad, 0, (size_t)(P > 0 ? P : 0) * sizeof(ScreenGrad), L.stream);
        return;
    }
    // defaults measured on B200 (config 3): asynchronously staged records free the 13 prefetch registers, which lets 18
    // warps per SM fit without rematerialisation (round 1: 450 us with per-lane bulk copies vs 464 us for register prefetch
    // at 16 warps); round 2: cp.async instead of bulk copies (a per-lane bulk copy is issued lane by lane, 256 instructions
    // per batch of an issue-bound kernel): 430 -> 416 us
    static const int minb = env_int("SCGR_BWD_MINB", 18);
    static const int tma = env_int("SCGR_TMA_BWD", env_int("SCGR_TMA", 2));
    WorkSplit ws = make_split(tx * ty, "SCGR_BWD_SPLIT", 0, 0);
    // SCGR_BWD_DEEP="h,q": cut the DEEPEST q % of the tiles in quarters and the next h % in halves instead
    if (const char* e = getenv("SCGR_BWD_DEEP")) {
        int ph = 0, pq = 0;
        sscanf(e, "%d,%d", &ph, &pq);
        const int n = tx * ty;
        ws.n2 = (int)((int64_t)n * pq / 100);
        ws.n4 = (int)((int64_t)n * ph / 100);
        if (ws.n2 + ws.n4 > n) { ws.n2 = 0; ws.n4 = n; }
        ws.n8 = n - ws.n4 - ws.n2;
        ws.deep_first = 1;
    }
    {
        const size_t zero_f4 = (size_t)(P > 0 ? P : 0) * (sizeof(ScreenGrad) / sizeof(float4));
        const unsigned zero_ctas = (unsigned)((zero_f4 + ZERO_F4_PER_CTA - 1) / ZERO_F4_PER_CTA);
        begin_kernel("backward_prologue", L);
        chain(backward_prologue_kernel, dim3(1 + zero_ctas), dim3(ORDER_THREADS), 0, L)(
            I.tile_todo, tx * ty, I.tile_order, reinterpret_cast<float4*>(G.screen_grad), zero_f4);
        check_launch("backward_prologue", L);
    }
    begin_kernel("render_backward", L);
#define SCGR_BWD(M_, T_, P_) chain(render_backward_kernel<M_, T_, P_>, dim3(ws.items()), dim3(32), 0, L)(ws, tx, B.ranges, point_list, \
        G.rec, v.image_width, v.image_height, v.bg, G.status, capacity, I.n_contrib, I.final_T, dL_dcolor, dL_ddepth, \
        dL_dalpha, G.screen_grad, I.tile_order)
    static const int packed = env_int("SCGR_BWD_PACKED", 0);      // packed-fp32 blend: measured slower, off by default
    if (packed) {
        if (!tma) { if (minb == 16) SCGR_BWD(16, false, true); else SCGR_BWD(1, false, true); }
        else if (minb == 16) SCGR_BWD(16, true, true); else if (minb == 14) SCGR_BWD(14, true, true); else if (minb == 18) SCGR_BWD(18, true, true); else SCGR_BWD(1, true, true);
    }
    else if (!tma) { if (minb == 16) SCGR_BWD(16, false, false); else if (minb == 18) SCGR_BWD(18, false, false); else SCGR_BWD(1, false, false); }
    else if (tma == 2) {      // cp.async staging
        if (minb == 16) SCGR_BWD(16, 2, false); else if (minb == 20) SCGR_BWD(20, 2, false); else if (minb == 19) SCGR_BWD(19, 2, false);
        else if (minb == 21) SCGR_BWD(21, 2, false); else SCGR_BWD(18, 2, false);
    }
    else if (minb == 16) SCGR_BWD(16, true, false); else if (minb == 14) SCGR_BWD(14, true, false); else if (minb == 20) SCGR_BWD(20, true, false); else if (minb == 18) SCGR_BWD(18, true, false); else if (minb == 19) SCGR_BWD(19, true, false); else if (minb == 17) SCGR_BWD(17, true, false); else SCGR_BWD(1, true, false);
#undef SCGR_BWD
    check_launch("render_backward", L);
}

}  // namespace scgr

#ifdef SCGR_STATS
// tools-only entry point of the -DSCGR_STATS build (not part of include/scgr.h)
extern "C" int scgr_debug_render_stats(unsigned long long* out32, int reset) {
    if (cudaMemcpyFromSymbol(out32, scgr::d_render_stats, 32 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(scgr::d_render_stats, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif
