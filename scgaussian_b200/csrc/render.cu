// render.cu -- per-tile alpha compositing, forward and backward (sm_100a).
//
// Replaces the external rasterizer's FORWARD::renderCUDA / BACKWARD::renderCUDA (SURVEY.md
// section 2c, section 8a rows a14/a15; semantics Appendix A.8 / A.9) behind
// reference gaussian_renderer/__init__.py:100-108.
//
// Design (differs from the reference's one-thread-per-pixel, 256-thread CTA):
//   * ONE WARP per 16x16 tile.  The tile is cut into 8 "slots" of 8x4 pixels; lane l owns pixel
//     (l % 8, l / 8) of every slot, i.e. 8 pixels per thread held in registers.
//   * The tile's depth-ordered list is consumed 32 Gaussians at a time: lane l gathers the packed
//     48-byte record of the l-th one (3 x 128-bit loads; the NEXT batch is prefetched into
//     registers while the current one is processed) and computes, for its Gaussian, an exact
//     closed-form bound of the maximum of the Gaussian's exponent over each slot rectangle.
//     Slots whose bound says alpha < 1/255 everywhere are skipped without touching a pixel
//     (conservative: the per-pixel test that follows is the reference's, so no output changes).
//   * Backward: every lane accumulates its 10 per-Gaussian gradient terms over its 8 pixels in
//     registers; one 12-shuffle transposing butterfly per (tile, Gaussian) leaves each term summed
//     over all 256 pixels in one lane, and those 10 lanes issue one coalesced `red.global.add.f32`
//     -- instead of the reference's 10 global atomics per contributing (pixel, Gaussian) pair.
//   * exp() is one MUFU.EX2: the conic is stored pre-multiplied by -0.5*log2(e).
// No block-level barriers at all (a warp never waits for another); no tensor cores (blend is not a
// contraction).
#include "common.cuh"

namespace scgr {

namespace {

constexpr int SLOTS = 8;
constexpr float PREFILTER_MARGIN = CULL_MARGIN;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Rec {
    float4 q0, q1, q2;
};

__device__ __forceinline__ Rec load_rec(const Record* __restrict__ rec, uint32_t id) {
    const float4* p = reinterpret_cast<const float4*>(rec + id);
    Rec r;
    r.q0 = __ldg(p);
    r.q1 = __ldg(p + 1);
    r.q2 = __ldg(p + 2);
    return r;
}

// slot mask of one Gaussian: bit k set <=> slot k may receive a contribution
__device__ __forceinline__ uint32_t slot_mask(const Rec& r, const float X0, const float Y0, const uint32_t live_slots) {
    const CullParams c = make_cull(r.q0, r.q1);
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        const float x0 = X0 + (float)((k & 1) << 3), y0 = Y0 + (float)((k >> 1) << 2);
        const float mp = max_power_over_rect(c.cA, c.cB, c.cC, c.kx, c.ky, c.mx, c.my, x0, x0 + 7.f, y0, y0 + 3.f);
        if (mp >= c.thr) m |= 1u << k;
    }
    return m & live_slots;
}

// ------------------------------------------------------------------------------------------
// forward (A.8)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const Record* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const int64_t* __restrict__ status, int64_t capacity, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
    __shared__ float4 s_q0[32], s_q1[32], s_q2[32];
    if (status[0] > capacity) return;   // binning overflowed: caller re-runs with a larger buffer
    const int lane = threadIdx.x;
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = blockIdx.x * TILE, Y0 = blockIdx.y * TILE;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int total = (int)(range.y - range.x);

    float T[SLOTS], Cr[SLOTS], Cg[SLOTS], Cb[SLOTS], Dd[SLOTS];
    uint32_t last[SLOTS];
    uint32_t done = 0u;    // bit k: this lane's pixel of slot k is finished
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        T[k] = 1.f; Cr[k] = 0.f; Cg[k] = 0.f; Cb[k] = 0.f; Dd[k] = 0.f; last[k] = 0u;
        const int px = X0 + ((k & 1) << 3) + lx, py = Y0 + ((k >> 1) << 2) + ly;
        if (px >= W || py >= H) done |= 1u << k;
    }
    const float pxf = (float)(X0 + lx), pyf = (float)(Y0 + ly);

    Rec nxt;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < total) nxt = load_rec(rec, point_list[range.x + lane]);

    for (int base = 0; base < total; base += 32) {
        // slots in which some pixel is still open
        uint32_t live = 0u;
#pragma unroll
        for (int k = 0; k < SLOTS; k++)
            if (!__all_sync(0xffffffffu, (done >> k) & 1u)) live |= 1u << k;
        if (live == 0u) break;
        const int cnt = min(32, total - base);
        const Rec cur = nxt;
        __syncwarp();
        s_q0[lane] = cur.q0; s_q1[lane] = cur.q1; s_q2[lane] = cur.q2;
        __syncwarp();
        if (base + 32 + lane < total) nxt = load_rec(rec, point_list[range.x + base + 32 + lane]);
        const uint32_t mymask = lane < cnt ? slot_mask(cur, (float)X0, (float)Y0, live) : 0u;

        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            const float4 q0 = s_q0[j];
            const float4 q1 = s_q1[j];
            const float4 q2 = s_q2[j];
            const float dx0 = q0.x - pxf, dy0 = q0.y - pyf;
#pragma unroll
            for (int k = 0; k < SLOTS; k++) {
                if (!(mj & (1u << k))) continue;      // warp-uniform
                const float dx = dx0 - (float)((k & 1) << 3), dy = dy0 - (float)((k >> 1) << 2);
                const float power = dx * (q0.z * dx + q0.w * dy) + q1.x * dy * dy;
                if (((done >> k) & 1u) || power > 0.f || power < q1.w - PREFILTER_MARGIN) continue;
                const float alpha = fminf(ALPHA_MAX, q1.y * ex2(power));
                if (alpha < ALPHA_MIN) continue;
                const float test_T = T[k] * (1.f - alpha);
                if (test_T < T_EPS) { done |= 1u << k; continue; }
                const float w = alpha * T[k];
                Cr[k] += q2.x * w; Cg[k] += q2.y * w; Cb[k] += q2.z * w;
                Dd[k] += q1.z * w;
                T[k] = test_T;
                last[k] = (uint32_t)(base + j + 1);
            }
        }
    }
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        const int px = X0 + ((k & 1) << 3) + lx, py = Y0 + ((k >> 1) << 2) + ly;
        if (px < W && py < H) {
            const size_t pid = (size_t)py * W + px;
            out_color[pid] = Cr[k] + T[k] * bg0;
            out_color[N + pid] = Cg[k] + T[k] * bg1;
            out_color[2 * N + pid] = Cb[k] + T[k] * bg2;
            out_depth[pid] = Dd[k];
            out_alpha[pid] = 1.f - T[k];      // == sum alpha_i T_i (telescoping), A.8
            n_contrib[pid] = last[k];
            final_T[pid] = T[k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward (A.9)
// ------------------------------------------------------------------------------------------
// 10 per-lane partial sums -> lane (h, b8, b4, b2, 0) ends up holding the full 32-lane sum of ONE
// of them.  12 shuffles instead of 50.  Returns the value; `*slot` = float offset inside ScreenGrad
// (or -1 if this lane holds nothing).
__device__ __forceinline__ float transpose_reduce10(const float v[10], const int lane, int* slot) {
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
    // which of the 10 values does this lane hold?
    int idx = -1;
    if (!(lane & 1)) {
        if (!b8) idx = b4 ? (b2 ? -1 : 2) : (b2 ? 1 : 0);
        else idx = b4 ? -1 : (b2 ? 4 : 3);
        if (idx >= 0 && h) idx += 5;
    }
    // value order {mx, my, A, B, C, opacity, depth, r, g, b} -> ScreenGrad float offsets
    *slot = idx < 0 ? -1 : (idx < 7 ? idx : idx + 1);
    return d;
}

__global__ void __launch_bounds__(32)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const Record* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                       const int64_t* __restrict__ status, int64_t capacity,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T,
                       const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                       const float* __restrict__ dL_dalpha, ScreenGrad* __restrict__ screen_grad) {
    __shared__ float4 s_q0[32], s_q1[32], s_q2[32];
    __shared__ uint32_t s_id[32];
    __shared__ float s_g[5][TILE_PIX];      // upstream gradients of the tile: r, g, b, depth, alpha
    if (status[0] > capacity) return;
    const int lane = threadIdx.x;
    const int lx = lane & 7, ly = lane >> 3;
    const int X0 = blockIdx.x * TILE, Y0 = blockIdx.y * TILE;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const size_t N = (size_t)W * H;

    float T[SLOTS], Ar[SLOTS], Ag[SLOTS], Ab[SLOTS], Ad[SLOTS], Aa[SLOTS], tfb[SLOTS];
    int lc[SLOTS];
    int slot_lc[SLOTS];
    int toDo = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        const int px = X0 + ((k & 1) << 3) + lx, py = Y0 + ((k >> 1) << 2) + ly;
        float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f, Tf = 0.f;
        lc[k] = 0;
        if (px < W && py < H) {
            const size_t pid = (size_t)py * W + px;
            Tf = final_T[pid];
            lc[k] = (int)n_contrib[pid];
            gr = dL_dcolor[pid]; gg = dL_dcolor[N + pid]; gb = dL_dcolor[2 * N + pid];
            gd = dL_ddepth[pid];
            ga = dL_dalpha[pid];
        }
        s_g[0][k * 32 + lane] = gr; s_g[1][k * 32 + lane] = gg; s_g[2][k * 32 + lane] = gb;
        s_g[3][k * 32 + lane] = gd; s_g[4][k * 32 + lane] = ga;
        T[k] = Tf;
        tfb[k] = -Tf * (bg0 * gr + bg1 * gg + bg2 * gb);
        Ar[k] = 0.f; Ag[k] = 0.f; Ab[k] = 0.f; Ad[k] = 0.f; Aa[k] = 0.f;
        slot_lc[k] = __reduce_max_sync(0xffffffffu, lc[k]);
        toDo = max(toDo, slot_lc[k]);
    }
    const float pxf = (float)(X0 + lx), pyf = (float)(Y0 + ly);
    // (each lane only ever reads back the s_g entries it wrote itself: no barrier needed)

    // batch entry j  <->  0-based list position  pos = toDo - 1 - (base + j)   (back to front)
    Rec nxt;
    uint32_t nxt_id = 0u;
    nxt.q0 = nxt.q1 = nxt.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < toDo) {
        nxt_id = point_list[range.x + (toDo - 1 - lane)];
        nxt = load_rec(rec, nxt_id);
    }
    for (int base = 0; base < toDo; base += 32) {
        const int cnt = min(32, toDo - base);
        const Rec cur = nxt;
        __syncwarp();
        s_q0[lane] = cur.q0; s_q1[lane] = cur.q1; s_q2[lane] = cur.q2; s_id[lane] = nxt_id;
        __syncwarp();
        if (base + 32 + lane < toDo) {
            nxt_id = point_list[range.x + (toDo - 1 - (base + 32 + lane))];
            nxt = load_rec(rec, nxt_id);
        }
        uint32_t mymask = 0u;
        if (lane < cnt) {
            const int mypos = toDo - 1 - (base + lane);
            uint32_t live = 0u;
#pragma unroll
            for (int k = 0; k < SLOTS; k++)
                if (mypos < slot_lc[k]) live |= 1u << k;
            mymask = slot_mask(cur, (float)X0, (float)Y0, live);
        }
        for (int j = 0; j < cnt; j++) {
            const uint32_t mj = __shfl_sync(0xffffffffu, mymask, j);
            if (mj == 0u) continue;
            const int pos = toDo - 1 - (base + j);
            const float4 q0 = s_q0[j];
            const float4 q1 = s_q1[j];
            const float4 q2 = s_q2[j];
            const float dx0 = q0.x - pxf, dy0 = q0.y - pyf;
            float v[10];
#pragma unroll
            for (int i = 0; i < 10; i++) v[i] = 0.f;
            bool touched = false;
#pragma unroll
            for (int k = 0; k < SLOTS; k++) {
                if (!(mj & (1u << k))) continue;      // warp-uniform
                const float dx = dx0 - (float)((k & 1) << 3), dy = dy0 - (float)((k >> 1) << 2);
                const float power = dx * (q0.z * dx + q0.w * dy) + q1.x * dy * dy;
                if (pos >= lc[k] || power > 0.f || power < q1.w - PREFILTER_MARGIN) continue;
                const float G = ex2(power);
                const float alpha = fminf(ALPHA_MAX, q1.y * G);
                if (alpha < ALPHA_MIN) continue;
                touched = true;
                const float ra = 1.f / (1.f - alpha);
                T[k] *= ra;                                  // transmittance in front of this Gaussian
                const float gr = s_g[0][k * 32 + lane], gg = s_g[1][k * 32 + lane], gb = s_g[2][k * 32 + lane];
                const float gd = s_g[3][k * 32 + lane], ga = s_g[4][k * 32 + lane];
                // suffix-blended values behind this Gaussian: A* = alpha_{j+1} c_{j+1} + (1-alpha_{j+1}) A*
                const float er = q2.x - Ar[k], eg = q2.y - Ag[k], eb = q2.z - Ab[k], ed = q1.z - Ad[k], ea = 1.f - Aa[k];
                float dL_dalpha_ = er * gr + eg * gg + eb * gb + ed * gd + ea * ga;
                Ar[k] += alpha * er; Ag[k] += alpha * eg; Ab[k] += alpha * eb; Ad[k] += alpha * ed; Aa[k] += alpha * ea;
                dL_dalpha_ = dL_dalpha_ * T[k] + tfb[k] * ra;
                const float w = alpha * T[k];
                const float u = q1.y * dL_dalpha_;           // dL/dG, propagated even when alpha was capped (A.9)
                const float gdx = G * dx, gdy = G * dy;
                v[0] += u * (2.f * q0.z * gdx + q0.w * gdy);  // * ln2  = dL/dmean_x (pixel units)
                v[1] += u * (2.f * q1.x * gdy + q0.w * gdx);  // * ln2  = dL/dmean_y
                const float ux = u * gdx;
                v[2] += ux * dx;                             // * -0.5 = dL/dconic_A
                v[3] += ux * dy;                             // * -1   = dL/dconic_B
                v[4] += u * gdy * dy;                        // * -0.5 = dL/dconic_C
                v[5] += G * dL_dalpha_;                      // dL/dopacity
                v[6] += w * gd;                              // dL/ddepth
                v[7] += w * gr; v[8] += w * gg; v[9] += w * gb;
            }
            if (!__any_sync(0xffffffffu, touched)) continue;
            int slot;
            const float sum = transpose_reduce10(v, lane, &slot);
            if (slot >= 0 && sum != 0.f)
                atomicAdd(reinterpret_cast<float*>(screen_grad + s_id[j]) + slot, sum);   // RED.E.ADD.F32
        }
    }
}

}  // namespace

void launch_render_forward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                           const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                           float* out_color, float* out_depth, float* out_alpha, const Launch& L) {
    const dim3 grid((v.image_width + TILE - 1) / TILE, (v.image_height + TILE - 1) / TILE);
    if (grid.x == 0 || grid.y == 0) return;
    begin_kernel("render_forward", L);
    render_forward_kernel<<<grid, 32, 0, L.stream>>>(B.ranges, point_list, G.rec, v.image_width,
                                                     v.image_height, v.bg, G.status, capacity, out_color,
                                                     out_depth, out_alpha, I.n_contrib, I.final_T);
    check_launch("render_forward", L);
}

void launch_render_backward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                            const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                            const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                            int32_t P, const Launch& L) {
    const dim3 grid((v.image_width + TILE - 1) / TILE, (v.image_height + TILE - 1) / TILE);
    cudaMemsetAsync(G.screen_grad, 0, (size_t)(P > 0 ? P : 0) * sizeof(ScreenGrad), L.stream);
    if (grid.x == 0 || grid.y == 0) return;
    begin_kernel("render_backward", L);
    render_backward_kernel<<<grid, 32, 0, L.stream>>>(B.ranges, point_list, G.rec, v.image_width,
                                                      v.image_height, v.bg, G.status, capacity, I.n_contrib,
                                                      I.final_T, dL_dcolor, dL_ddepth, dL_dalpha, G.screen_grad);
    check_launch("render_backward", L);
}

}  // namespace scgr
