// render.cu -- per-tile alpha compositing, forward and backward (sm_100a).
//
// Replaces the external rasterizer's FORWARD::renderCUDA / BACKWARD::renderCUDA (SURVEY.md
// section 2c, section 8a rows a14/a15; semantics Appendix A.8 / A.9) behind
// reference gaussian_renderer/__init__.py:100-108.
//
// One CTA per 16x16 tile, one thread per pixel; a warp owns an 8x4 pixel block (compact
// footprint -> more warp-uniform skips than the reference's 16x2 rows).  The tile's depth-ordered
// Gaussian list is consumed in batches of 256: each thread gathers ONE packed 48-byte record
// {xy, conic, opacity, depth, rgb} with three 128-bit loads (the reference gathers five separate
// arrays and re-reads colour/depth from global memory per pixel per Gaussian) into shared
// memory, from where all 256 pixels read it by broadcast.
//
// Backward: the per-(pixel, Gaussian) gradients are reduced hierarchically -- warp shuffle
// butterfly over the 32 pixels of a warp, shared-memory float atomics across the 8 warps of the
// tile, then ONE set of three 128-bit vector reductions (red.global.add.v4.f32) per (Gaussian,
// tile) -- instead of the reference's 10 global atomics per contributing pair.
#include "common.cuh"

namespace scgr {

namespace {

constexpr int BATCH = 256;

__device__ __forceinline__ void pixel_of_thread(int& lx, int& ly) {
    // warp w covers the 8x4 block at (w%2 * 8, w/2 * 4); lane -> (lane%8, lane/8)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    lx = ((w & 1) << 3) + (lane & 7);
    ly = ((w >> 1) << 2) + (lane >> 3);
}

__global__ void __launch_bounds__(TILE_PIX)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const Record* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const int64_t* __restrict__ status, int64_t capacity, float* __restrict__ out_color,
                      float* __restrict__ out_depth, float* __restrict__ out_alpha,
                      uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
    __shared__ float4 s_q0[BATCH], s_q1[BATCH], s_q2[BATCH];
    if (status[0] > capacity) return;   // binning overflowed: caller re-runs with a larger buffer
    int lx, ly;
    pixel_of_thread(lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int total = (int)(range.y - range.x);

    bool done = !inside;
    float T = 1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dsum = 0.f, Wsum = 0.f;
    uint32_t last = 0u;

    for (int base = 0; base < total; base += BATCH) {
        if (__syncthreads_count(done) == TILE_PIX) break;
        const int cnt = min(BATCH, total - base);
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = point_list[range.x + base + threadIdx.x];
            const Record* r = rec + id;
            s_q0[threadIdx.x] = r->q0;
            s_q1[threadIdx.x] = r->q1;
            s_q2[threadIdx.x] = r->q2;
        }
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            if (__all_sync(0xffffffffu, done)) break;
            const float4 q0 = s_q0[j];
            const float4 q1 = s_q1[j];
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            const float alpha = fminf(ALPHA_MAX, q1.y * __expf(power));
            if (done || power > 0.f || alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < T_EPS) { done = true; continue; }
            const float4 q2 = s_q2[j];
            const float w = alpha * T;
            Cr += q2.x * w; Cg += q2.y * w; Cb += q2.z * w;
            Dsum += q1.z * w;
            Wsum += w;
            T = test_T;
            last = (uint32_t)(base + j + 1);
        }
    }
    if (inside) {
        const size_t pid = (size_t)py * W + px, N = (size_t)W * H;
        out_color[pid] = Cr + T * __ldg(bg);
        out_color[N + pid] = Cg + T * __ldg(bg + 1);
        out_color[2 * N + pid] = Cb + T * __ldg(bg + 2);
        out_depth[pid] = Dsum;
        out_alpha[pid] = Wsum;
        n_contrib[pid] = last;
        final_T[pid] = T;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(TILE_PIX)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const Record* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                       const int64_t* __restrict__ status, int64_t capacity,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T,
                       const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                       const float* __restrict__ dL_dalpha, ScreenGrad* __restrict__ screen_grad) {
    __shared__ float4 s_q0[BATCH], s_q1[BATCH], s_q2[BATCH];
    __shared__ uint32_t s_id[BATCH];
    __shared__ float s_acc[10][BATCH];
    __shared__ int s_max[TILE_PIX / 32];
    if (status[0] > capacity) return;
    int lx, ly;
    pixel_of_thread(lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int lane = threadIdx.x & 31;

    float T_final = 0.f, gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
    int last_contributor = 0;
    if (inside) {
        const size_t pid = (size_t)py * W + px, N = (size_t)W * H;
        T_final = final_T[pid];
        last_contributor = (int)n_contrib[pid];
        gr = dL_dcolor[pid]; gg = dL_dcolor[N + pid]; gb = dL_dcolor[2 * N + pid];
        gd = dL_ddepth[pid];
        ga = dL_dalpha[pid];
    }
    const float bgdot = __ldg(bg) * gr + __ldg(bg + 1) * gg + __ldg(bg + 2) * gb;

    // nothing behind the deepest contributor of the whole tile matters
    int m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    int toDo = 0;
#pragma unroll
    for (int k = 0; k < TILE_PIX / 32; k++) toDo = max(toDo, s_max[k]);

    float T = T_final;
    float arec_r = 0.f, arec_g = 0.f, arec_b = 0.f, drec = 0.f, alrec = 0.f;
    float last_alpha = 0.f, last_r = 0.f, last_g = 0.f, last_b = 0.f, last_d = 0.f;

    for (int base = 0; base < toDo; base += BATCH) {
        // batch entry j  <->  0-based list position  pos = toDo - 1 - (base + j)
        const int cnt = min(BATCH, toDo - base);
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = point_list[range.x + (toDo - 1 - (base + threadIdx.x))];
            const Record* r = rec + id;
            s_id[threadIdx.x] = id;
            s_q0[threadIdx.x] = r->q0;
            s_q1[threadIdx.x] = r->q1;
            s_q2[threadIdx.x] = r->q2;
        }
#pragma unroll
        for (int k = 0; k < 10; k++) s_acc[k][threadIdx.x] = 0.f;
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            const int pos = toDo - 1 - (base + j);
            const float4 q0 = s_q0[j];
            const float4 q1 = s_q1[j];
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            const float G = __expf(power);
            const float alpha = fminf(ALPHA_MAX, q1.y * G);
            const bool ok = (pos < last_contributor) && (power <= 0.f) && (alpha >= ALPHA_MIN);
            if (!__any_sync(0xffffffffu, ok)) continue;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f, v9 = 0.f;
            if (ok) {
                const float4 q2 = s_q2[j];
                T = T / (1.f - alpha);
                const float w = alpha * T;
                float dL_dalpha_ = 0.f;
                arec_r = last_alpha * last_r + (1.f - last_alpha) * arec_r; last_r = q2.x;
                dL_dalpha_ += (q2.x - arec_r) * gr;
                arec_g = last_alpha * last_g + (1.f - last_alpha) * arec_g; last_g = q2.y;
                dL_dalpha_ += (q2.y - arec_g) * gg;
                arec_b = last_alpha * last_b + (1.f - last_alpha) * arec_b; last_b = q2.z;
                dL_dalpha_ += (q2.z - arec_b) * gb;
                drec = last_alpha * last_d + (1.f - last_alpha) * drec; last_d = q1.z;
                dL_dalpha_ += (q1.z - drec) * gd;
                alrec = last_alpha + (1.f - last_alpha) * alrec;
                dL_dalpha_ += (1.f - alrec) * ga;
                dL_dalpha_ *= T;
                last_alpha = alpha;
                dL_dalpha_ += (-T_final / (1.f - alpha)) * bgdot;
                const float dL_dG = q1.y * dL_dalpha_;   // propagated even when alpha was capped (A.9)
                const float gdx = G * dx, gdy = G * dy;
                v0 = dL_dG * (-gdx * q0.z - gdy * q0.w);      // dL/dmean_x, pixel units
                v1 = dL_dG * (-gdy * q1.x - gdx * q0.w);      // dL/dmean_y
                v2 = -0.5f * gdx * dx * dL_dG;                // dL/dconic_A
                v3 = -gdx * dy * dL_dG;                       // dL/dconic_B (true derivative)
                v4 = -0.5f * gdy * dy * dL_dG;                // dL/dconic_C
                v5 = G * dL_dalpha_;                          // dL/dopacity
                v6 = w * gd;                                  // dL/ddepth
                v7 = w * gr; v8 = w * gg; v9 = w * gb;        // dL/drgb
            }
            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
            v5 = warp_sum(v5); v6 = warp_sum(v6); v7 = warp_sum(v7); v8 = warp_sum(v8); v9 = warp_sum(v9);
            if (lane == 0) {
                atomicAdd(&s_acc[0][j], v0); atomicAdd(&s_acc[1][j], v1); atomicAdd(&s_acc[2][j], v2);
                atomicAdd(&s_acc[3][j], v3); atomicAdd(&s_acc[4][j], v4); atomicAdd(&s_acc[5][j], v5);
                atomicAdd(&s_acc[6][j], v6); atomicAdd(&s_acc[7][j], v7); atomicAdd(&s_acc[8][j], v8);
                atomicAdd(&s_acc[9][j], v9);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const int j = threadIdx.x;
            const float4 a0 = make_float4(s_acc[0][j], s_acc[1][j], s_acc[2][j], s_acc[3][j]);
            const float4 a1 = make_float4(s_acc[4][j], s_acc[5][j], s_acc[6][j], 0.f);
            const float4 a2 = make_float4(s_acc[7][j], s_acc[8][j], s_acc[9][j], 0.f);
            const bool any = a0.x != 0.f || a0.y != 0.f || a0.z != 0.f || a0.w != 0.f || a1.x != 0.f ||
                             a1.y != 0.f || a1.z != 0.f || a2.x != 0.f || a2.y != 0.f || a2.z != 0.f;
            if (any) {
                ScreenGrad* dst = screen_grad + s_id[j];
                atomicAdd(&dst->a0, a0);   // red.global.add.v4.f32 (sm_90+)
                atomicAdd(&dst->a1, a1);
                atomicAdd(&dst->a2, a2);
            }
        }
        __syncthreads();
    }
}

}  // namespace

void launch_render_forward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                           const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                           float* out_color, float* out_depth, float* out_alpha, const Launch& L) {
    const dim3 grid((v.image_width + TILE - 1) / TILE, (v.image_height + TILE - 1) / TILE);
    if (grid.x == 0 || grid.y == 0) return;
    begin_kernel("render_forward", L);
    render_forward_kernel<<<grid, TILE_PIX, 0, L.stream>>>(B.ranges, point_list, G.rec, v.image_width,
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
    render_backward_kernel<<<grid, TILE_PIX, 0, L.stream>>>(B.ranges, point_list, G.rec, v.image_width,
                                                            v.image_height, v.bg, G.status, capacity, I.n_contrib,
                                                            I.final_T, dL_dcolor, dL_ddepth, dL_dalpha, G.screen_grad);
    check_launch("render_backward", L);
}

}  // namespace scgr
