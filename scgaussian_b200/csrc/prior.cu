// prior.cu -- the two small loss terms SCGaussian adds around the rasterizer's depth and alpha outputs (sm_100a).
// SURVEY.md section 8(f) row f1 (second half) and BASELINE config 5's loss path.
//
//   * match-prior loss on the RENDERED DEPTH (reference scene/gaussian_model.py:241-282
//     `GaussianModel.get_matchloss_from_renderdepth`, called at reference train.py:164-165 with weight 0.3): for every
//     other view, bilinear `grid_sample` of the depth image at the matched pixels, back-projection along the matched
//     rays, re-projection into the other view, L1 distance to the matched pixels there, masked mean.  ~20 torch
//     launches per view pair in the reference; here ONE launch forward and ONE backward for all pairs of a view, the
//     backward scattering straight into dL/d(rendered depth) -- the tensor the rasterizer's backward consumes.
//   * the DTU background term (reference train.py:150-158, :167-168): `bg_mask` = dark pixels of the ground truth,
//     eroded by the reference's 49-step shift loop (a pixel stays masked when the 49 pixels above it in its column are
//     dark too); the ground truth is zeroed there and `rendered_alpha[bg_mask].mean()` joins the loss.  One kernel for
//     mask + zeroing + count (instead of ~100 launches of the Python loop), one for the masked mean, one for its backward.
//
// All of it is latency-bound bookkeeping on <= 2000 matches per pair / one image: no tensor cores, no tiling games; what
// matters is the launch count and that nothing synchronises with the host.
#include <cstdint>

#include "common.cuh"

namespace scgr {

namespace {

constexpr int PRIOR_THREADS = 1024;

struct MatchTable {
    ScgrMatchPair pair[SCGR_MATCH_MAX_PAIRS];
    int n_pairs;
};

struct MatchGeom {      // everything one match contributes, shared by forward and backward
    float w[4];         // bilinear weights nw, ne, sw, se
    int ix, iy;         // north-west tap
    float depth;        // sampled depth
    float x, y, z;      // projection into the other view: pixel (x, y), depth z (before the divide: xyz[2])
    float dx, dy;       // d(x)/d(sampled depth), d(y)/d(sampled depth)
    bool inside;        // mask_0to1
};

__device__ __forceinline__ float tap(const float* __restrict__ depth, const int H, const int W, const int ix, const int iy) {
    return (ix >= 0 && ix < W && iy >= 0 && iy < H) ? __ldg(depth + (size_t)iy * W + ix) : 0.f;      // padding_mode="zeros"
}

__device__ __forceinline__ MatchGeom match_geometry(const ScgrMatchPair& p, const int m, const float* __restrict__ depth,
                                                    const int H, const int W, const float width, const float height) {
    MatchGeom g;
    // F.grid_sample(..., mode="bilinear") with its defaults padding_mode="zeros", align_corners=False (reference :256-259)
    const float u = __ldg(p.uv0 + 2 * m), v = __ldg(p.uv0 + 2 * m + 1);
    const float nx = (u / width) * 2.f - 1.f, ny = (v / height) * 2.f - 1.f;
    const float fx = ((nx + 1.f) * W - 1.f) / 2.f, fy = ((ny + 1.f) * H - 1.f) / 2.f;
    const float x0 = floorf(fx), y0 = floorf(fy);
    g.ix = (int)x0; g.iy = (int)y0;
    const float x1 = x0 + 1.f, y1 = y0 + 1.f;
    g.w[0] = (x1 - fx) * (y1 - fy); g.w[1] = (fx - x0) * (y1 - fy);
    g.w[2] = (x1 - fx) * (fy - y0); g.w[3] = (fx - x0) * (fy - y0);
    g.depth = tap(depth, H, W, g.ix, g.iy) * g.w[0] + tap(depth, H, W, g.ix + 1, g.iy) * g.w[1] +
              tap(depth, H, W, g.ix, g.iy + 1) * g.w[2] + tap(depth, H, W, g.ix + 1, g.iy + 1) * g.w[3];
    // zval = depth / cam_rays_d.z;  world = rays_o + rays_d * zval  (:261-263)
    const float cz = __ldg(p.cam_rays_d + 3 * m + 2);
    const float zval = g.depth / cz;
    const float rd[3] = {__ldg(p.rays_d + 3 * m), __ldg(p.rays_d + 3 * m + 1), __ldg(p.rays_d + 3 * m + 2)};
    const float wp[3] = {__ldg(p.rays_o + 3 * m) + rd[0] * zval, __ldg(p.rays_o + 3 * m + 1) + rd[1] * zval,
                         __ldg(p.rays_o + 3 * m + 2) + rd[2] * zval};
    // cam = w2c1 [world; 1];  xyz = intr1 cam;  xy = xyz[:2] / (xyz[2] + 1e-8)  (:267-270)
    float cam[3], dcam[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        cam[r] = p.w2c1[4 * r] * wp[0] + p.w2c1[4 * r + 1] * wp[1] + p.w2c1[4 * r + 2] * wp[2] + p.w2c1[4 * r + 3];
        dcam[r] = (p.w2c1[4 * r] * rd[0] + p.w2c1[4 * r + 1] * rd[1] + p.w2c1[4 * r + 2] * rd[2]) / cz;      // d cam / d depth
    }
    float xyz[3], dxyz[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        xyz[r] = p.intr1[3 * r] * cam[0] + p.intr1[3 * r + 1] * cam[1] + p.intr1[3 * r + 2] * cam[2];
        dxyz[r] = p.intr1[3 * r] * dcam[0] + p.intr1[3 * r + 1] * dcam[1] + p.intr1[3 * r + 2] * dcam[2];
    }
    const float den = xyz[2] + 1e-8f;
    g.x = xyz[0] / den; g.y = xyz[1] / den; g.z = xyz[2];
    g.dx = (dxyz[0] - g.x * dxyz[2]) / den;
    g.dy = (dxyz[1] - g.y * dxyz[2]) / den;
    g.inside = g.x > 0.f && g.x < width && g.y > 0.f && g.y < height;      // mask_0to1 (:271)
    return g;
}

__device__ __forceinline__ float block_sum(float v, float* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_warp[w] = v;
    __syncthreads();
    float t = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_warp[0] = t;
    }
    __syncthreads();
    return s_warp[0];
}

// One CTA walks the pairs in order (deterministic sums).  out[0] = sum over pairs of the masked mean; inv_den[p] is kept
// for the backward.
__global__ void __launch_bounds__(PRIOR_THREADS)
match_loss_forward_kernel(const __grid_constant__ MatchTable t, const float* __restrict__ depth, const int H, const int W,
                          const float width, const float height, float* __restrict__ out, float* __restrict__ inv_den) {
    __shared__ float s_warp[32];
    float total = 0.f;
    for (int p = 0; p < t.n_pairs; p++) {
        const ScgrMatchPair& P = t.pair[p];
        float num = 0.f, den = 0.f;
        for (int m = threadIdx.x; m < P.n; m += PRIOR_THREADS) {
            const MatchGeom g = match_geometry(P, m, depth, H, W, width, height);
            const float valid = P.valid ? (__ldg(P.valid + m) > 0.f ? 1.f : 0.f) : 1.f;
            const float wgt = (g.inside ? 1.f : 0.f) * valid;
            const float cur = 0.5f * (fabsf(g.x - __ldg(P.uv1 + 2 * m)) / width + fabsf(g.y - __ldg(P.uv1 + 2 * m + 1)) / height);
            num += cur * wgt;
            den += wgt;
        }
        num = block_sum(num, s_warp);
        den = block_sum(den, s_warp);
        const float inv = 1.f / (den + 1e-8f);
        total += num * inv;
        if (threadIdx.x == 0) inv_den[p] = inv;
    }
    if (threadIdx.x == 0) out[0] = total;
}

// thread per match (all pairs flattened): dL/d(sampled depth) scattered to the four taps of the depth image
__global__ void __launch_bounds__(256)
match_loss_backward_kernel(const __grid_constant__ MatchTable t, const float* __restrict__ depth, const int H, const int W,
                           const float width, const float height, const float* __restrict__ inv_den,
                           const float* __restrict__ upstream, float* __restrict__ dL_ddepth) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int p = 0;
    while (p < t.n_pairs && m >= t.pair[p].n) { m -= t.pair[p].n; p++; }
    if (p >= t.n_pairs) return;
    const ScgrMatchPair& P = t.pair[p];
    const MatchGeom g = match_geometry(P, m, depth, H, W, width, height);
    const float valid = P.valid ? (__ldg(P.valid + m) > 0.f ? 1.f : 0.f) : 1.f;
    if (!g.inside || valid == 0.f) return;
    const float ex = g.x - __ldg(P.uv1 + 2 * m), ey = g.y - __ldg(P.uv1 + 2 * m + 1);
    const float sx = ex > 0.f ? 1.f : (ex < 0.f ? -1.f : 0.f), sy = ey > 0.f ? 1.f : (ey < 0.f ? -1.f : 0.f);   // d|x| = sign(x), 0 at 0
    const float up = upstream ? __ldg(upstream) : 1.f;
    const float gd = up * __ldg(inv_den + p) * 0.5f * (sx * g.dx / width + sy * g.dy / height);
    if (gd == 0.f) return;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ix = g.ix + (k & 1), iy = g.iy + (k >> 1);
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) atomicAdd(dL_ddepth + (size_t)iy * W + ix, g.w[k] * gd);
    }
}

// ---- DTU background mask (reference train.py:150-158) ----
// bg(y, x) = max_c gt[c, y, x] < threshold;  mask(y, x) = bg(y, x) & bg(y-1, x) & ... & bg(y-(window-1), x) over the rows
// that exist (the reference's `for i in range(1, 50): bg_mask[:, i:] *= bg_mask_clone[:, :-i]`).  A thread owns one
// column of a band of BAND rows and warms its run counter up on the window-1 rows above the band.
constexpr int MASK_BAND = 64;
__global__ void __launch_bounds__(256)
bg_mask_kernel(float* __restrict__ gt, const int C, const int H, const int W, const float threshold, const int window,
               uint8_t* __restrict__ mask, float* __restrict__ count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y_begin = blockIdx.y * MASK_BAND, y_end = min(H, y_begin + MASK_BAND);
    float n_masked = 0.f;
    if (x < W) {
        const size_t N = (size_t)H * W;
        int run = 0;
        for (int y = max(0, y_begin - (window - 1)); y < y_end; y++) {
            float mx = __ldg(gt + (size_t)y * W + x);
            for (int c = 1; c < C; c++) mx = fmaxf(mx, __ldg(gt + c * N + (size_t)y * W + x));
            run = mx < threshold ? run + 1 : 0;
            if (y >= y_begin) {
                const bool m = run >= min(y + 1, window);
                mask[(size_t)y * W + x] = m ? 1 : 0;
                n_masked += m ? 1.f : 0.f;
            }
        }
    }
    // the ground truth is zeroed where masked AFTER every thread of the band has read what it needs: the warm-up rows of
    // this band belong to the band above, so the zeroing is a second launch-wide phase (kernel below)
    __shared__ float s_warp[32];
    const float tot = block_sum(n_masked, s_warp);
    if (threadIdx.x == 0 && tot != 0.f) atomicAdd(count, tot);     // integer-valued partial sums: exact in any order
}

__global__ void apply_mask_kernel(float* __restrict__ gt, const int C, const size_t N, const uint8_t* __restrict__ mask) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && mask[i])
        for (int c = 0; c < C; c++) gt[c * N + i] = 0.f;
}

// ---- values[mask].mean() and its backward (reference train.py:167-168 on rendered_alpha) ----
constexpr int MEAN_PER_CTA = 4096;
__global__ void __launch_bounds__(256)
masked_mean_forward_kernel(const float* __restrict__ values, const uint8_t* __restrict__ mask, const size_t n,
                           float* __restrict__ partial, unsigned int* __restrict__ ticket, float* __restrict__ out2) {
    __shared__ float s_warp[32];
    __shared__ bool s_last;
    float s = 0.f, c = 0.f;
    const size_t base = (size_t)blockIdx.x * MEAN_PER_CTA;
    for (int k = threadIdx.x; k < MEAN_PER_CTA; k += 256) {
        const size_t i = base + k;
        if (i < n && mask[i]) { s += values[i]; c += 1.f; }
    }
    s = block_sum(s, s_warp);
    c = block_sum(c, s_warp);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = c;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last CTA adds the partials in a fixed order: deterministic
    float ts = 0.f, tc = 0.f;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += 256) { ts += partial[2 * k]; tc += partial[2 * k + 1]; }
    ts = block_sum(ts, s_warp);
    tc = block_sum(tc, s_warp);
    if (threadIdx.x == 0) { out2[0] = ts / tc; out2[1] = tc; }      // an empty mask gives NaN, as torch's mean of nothing does
}

__global__ void masked_mean_backward_kernel(const uint8_t* __restrict__ mask, const size_t n, const float* __restrict__ out2,
                                            const float* __restrict__ upstream, float* __restrict__ dL_dvalues) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = (upstream ? __ldg(upstream) : 1.f) / __ldg(out2 + 1);
    dL_dvalues[i] = mask[i] ? g : 0.f;
}

MatchTable make_table(const ScgrMatchPair* pairs, int n_pairs) {
    MatchTable t{};
    int k = 0;
    for (int i = 0; i < n_pairs; i++)
        if (pairs[i].n > 0) t.pair[k++] = pairs[i];
    t.n_pairs = k;
    return t;
}

}  // namespace

void launch_match_loss_forward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs,
                               int n_pairs, float* scratch, float* out, const Launch& L) {
    const MatchTable t = make_table(pairs, n_pairs);
    begin_kernel("match_loss_forward", L);
    match_loss_forward_kernel<<<1, PRIOR_THREADS, 0, L.stream>>>(t, depth, H, W, width, height, out, scratch);
    check_launch("match_loss_forward", L);
}

void launch_match_loss_backward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs,
                                int n_pairs, const float* scratch, const float* upstream, float* dL_ddepth, const Launch& L) {
    const MatchTable t = make_table(pairs, n_pairs);
    long long total = 0;
    for (int i = 0; i < t.n_pairs; i++) total += t.pair[i].n;
    cudaMemsetAsync(dL_ddepth, 0, (size_t)H * W * sizeof(float), L.stream);
    if (total == 0) return;
    begin_kernel("match_loss_backward", L);
    match_loss_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(t, depth, H, W, width, height, scratch, upstream,
                                                                                    dL_ddepth);
    check_launch("match_loss_backward", L);
}

void launch_bg_mask(float* gt, int C, int H, int W, float threshold, int window, uint8_t* mask, float* count, const Launch& L) {
    cudaMemsetAsync(count, 0, sizeof(float), L.stream);
    begin_kernel("bg_mask", L);
    bg_mask_kernel<<<dim3((W + 255) / 256, (H + MASK_BAND - 1) / MASK_BAND), 256, 0, L.stream>>>(gt, C, H, W, threshold, window, mask, count);
    check_launch("bg_mask", L);
    const size_t N = (size_t)H * W;
    begin_kernel("bg_mask_apply", L);
    apply_mask_kernel<<<(unsigned)((N + 255) / 256), 256, 0, L.stream>>>(gt, C, N, mask);
    check_launch("bg_mask_apply", L);
}

size_t masked_mean_scratch_bytes(size_t n) { return align_up(((n + MEAN_PER_CTA - 1) / MEAN_PER_CTA + 1) * 2 * sizeof(float) + 16); }

void launch_masked_mean_forward(const float* values, const uint8_t* mask, size_t n, void* scratch, float* out2, const Launch& L) {
    const unsigned blocks = (unsigned)((n + MEAN_PER_CTA - 1) / MEAN_PER_CTA);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
    float* partial = reinterpret_cast<float*>(scratch) + 4;
    cudaMemsetAsync(ticket, 0, 16, L.stream);
    begin_kernel("masked_mean_forward", L);
    masked_mean_forward_kernel<<<blocks, 256, 0, L.stream>>>(values, mask, n, partial, ticket, out2);
    check_launch("masked_mean_forward", L);
}

void launch_masked_mean_backward(const uint8_t* mask, size_t n, const float* out2, const float* upstream, float* dL_dvalues,
                                 const Launch& L) {
    begin_kernel("masked_mean_backward", L);
    masked_mean_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(mask, n, out2, upstream, dL_dvalues);
    check_launch("masked_mean_backward", L);
}

}  // namespace scgr
