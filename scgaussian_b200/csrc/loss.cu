// loss.cu -- fused photometric loss, forward + backward (sm_100a).  SURVEY.md section 8(f), row f1.
//
// Replaces what reference train.py:160-161 computes with a dozen torch kernels per direction:
//     Ll1  = l1_loss(image, gt)                                  reference utils/loss_utils.py:40-41
//     loss = (1 - lambda) * Ll1 + lambda * (1 - ssim(image, gt)) reference utils/loss_utils.py:56-94
// ssim(): 11x11 Gaussian window (sigma 1.5, reference :46-54), zero padding 5, depthwise, C1 = 0.01^2,
// C2 = 0.03^2, mean over every element.  The 2-D window is the outer product of the 1-D one (:52), so
// the five windowed moments E[x], E[y], E[x^2], E[y^2], E[xy] are computed separably.
//
// Forward (one kernel): a CTA owns a 32x32 pixel tile of one channel; both images are staged with a
// 5-pixel halo in shared memory (42x42), filtered horizontally (5 moments x 42 rows x 32 columns), then
// vertically; each thread produces 4 adjacent outputs per pass from 14 loaded values (register blocking:
// shared-memory loads are the bottleneck of a separable 11-tap filter).  The thread evaluates |x - y| and the SSIM value,
// and -- when a gradient is wanted -- the three partial derivatives of its SSIM value with respect to
// the windowed moments that depend on x:  dS/dE[x], dS/dE[x^2], dS/dE[xy].  Block-reduced sums go
// to a partial array; the last CTA to finish adds them up in a fixed order (deterministic) and
// writes {Ll1, ssim, loss}.
// Backward (one kernel): dL/dx(q) = (G * dS/dE[x])(q) + 2 x(q) (G * dS/dE[x^2])(q) + y(q) (G * dS/dE[xy])(q)
// (G symmetric, zero padding = sum over valid pixels), i.e. three more separable filters of the
// stored derivative maps, plus the L1 term sign(x - y); everything scaled by the weights of the two
// means and by the upstream scalar read from device memory (no host synchronisation).
// HBM-bound streaming work on fp32: no tensor cores.
#include <climits>
#include <cmath>

#include "common.cuh"

namespace scgr {

namespace {

constexpr int LT = 32;               // tile edge: a CTA of 256 threads owns 32x32 pixels, 4 per thread
constexpr int LTHREADS = 256;
constexpr int LPT = 4;               // outputs per thread along the filtered direction (register blocking)
constexpr int LHALO = 5;             // window 11
constexpr int LWIN = 2 * LHALO + 1;
constexpr int LE = LT + 2 * LHALO;   // 42: staged edge
constexpr int LSX = 45;              // staged row stride: = 1 mod 4, conflict-free for the 4-column groups
constexpr int LSH = LT + 1;          // filtered row stride
constexpr int LTILES = 2;            // consecutive tiles of a tile row walked by one CTA (register-prefetched)
constexpr int LNONE = INT_MIN;       // "no image row" marker of the staging offsets
constexpr int LSTAGE = (LE * LE + LTHREADS - 1) / LTHREADS;   // staged elements per thread
constexpr float SSIM_C1 = 0.01f * 0.01f;
constexpr float SSIM_C2 = 0.03f * 0.03f;

struct Window {
    float g[LWIN];
};

// reference utils/loss_utils.py:46-48: exp(-(x - 5)^2 / (2 sigma^2)) as float32, normalised in float32
Window make_window() {
    Window w;
    float sum = 0.f;
    for (int x = 0; x < LWIN; x++) {
        w.g[x] = (float)std::exp(-(double)((x - LWIN / 2) * (x - LWIN / 2)) / (2.0 * 1.5 * 1.5));
        sum += w.g[x];
    }
    for (int x = 0; x < LWIN; x++) w.g[x] /= sum;
    return w;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Shared memory is the scarce resource of a separable 11-tap filter (one LDS per tap and output
// otherwise): every thread produces LPT = 4 adjacent outputs from 14 loaded inputs, in both passes.
template <bool WANT_GRAD>
__global__ void __launch_bounds__(LTHREADS, 3)
photometric_forward_kernel(const float* __restrict__ img, const float* __restrict__ gt, const int H, const int W,
                           const Window win, float* __restrict__ d_mu, float* __restrict__ d_e11,
                           float* __restrict__ d_e12, float2* __restrict__ partial, unsigned int* __restrict__ counter,
                           const float lambda, const double inv_count, float* __restrict__ out3) {
    __shared__ float s_x[LE][LSX], s_y[LE][LSX];
    __shared__ float s_h[4][LE][LSH];       // E[x], E[y], E[x^2 + y^2], E[xy]: SSIM only needs sigma_1^2 + sigma_2^2
    __shared__ double s_red[2][LTHREADS / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int y0 = blockIdx.y * LT;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* __restrict__ ip = img + plane;
    const float* __restrict__ gp = gt + plane;
    const int tiles_x = (W + LT - 1) / LT;
    const int t_first = blockIdx.x * LTILES, t_end = min(tiles_x, t_first + LTILES);

    // A CTA walks LTILES consecutive tiles of a tile row.  The global loads of the NEXT tile are issued
    // into registers before the current tile is filtered, so their latency hides behind the arithmetic.
    float pre_a[LSTAGE], pre_b[LSTAGE];
    // staged element `it` of this thread: its place in the staging arrays and in the image row do not
    // depend on the tile -- computed once (soff < 0: nothing to stage / row outside the image)
    int soff[LSTAGE], goff[LSTAGE], scol[LSTAGE];
#pragma unroll
    for (int it = 0; it < LSTAGE; it++) {
        const int idx = tid + it * LTHREADS;
        const int r = idx / LE, c = idx - r * LE;
        const int gy = y0 + r - LHALO;
        scol[it] = c - LHALO;
        soff[it] = idx < LE * LE ? r * LSX + c : -1;
        goff[it] = (idx < LE * LE && gy >= 0 && gy < H) ? gy * W + c - LHALO : LNONE;   // (may be negative in row 0)
    }
    auto fetch = [&](const int t) {
        const int x0 = t * LT;
#pragma unroll
        for (int it = 0; it < LSTAGE; it++) {
            const int gx = x0 + scol[it];
            float a = 0.f, b = 0.f;
            if (goff[it] != LNONE && gx >= 0 && gx < W) {
                a = __ldg(ip + goff[it] + x0);
                b = __ldg(gp + goff[it] + x0);
            }
            pre_a[it] = a;
            pre_b[it] = b;
        }
    };
    float l1 = 0.f, ss_sum = 0.f;
    fetch(t_first);
    for (int t = t_first; t < t_end; t++) {
    const int x0 = t * LT;
#pragma unroll
    for (int it = 0; it < LSTAGE; it++) {
        if (soff[it] >= 0) {
            (&s_x[0][0])[soff[it]] = pre_a[it];
            (&s_y[0][0])[soff[it]] = pre_b[it];
        }
    }
    __syncthreads();
    if (t + 1 < t_end) fetch(t + 1);
    // horizontal: item = (row r, group of 4 columns)
    for (int idx = tid; idx < LE * (LT / LPT); idx += LTHREADS) {
        const int r = idx / (LT / LPT), c0 = (idx - r * (LT / LPT)) * LPT;
        // the 14 inputs of this item and their products x^2 + y^2, x y, formed once: 4 FMA per tap and output
        float xa[LWIN + LPT - 1], ya[LWIN + LPT - 1], qa[LWIN + LPT - 1], pa[LWIN + LPT - 1];
#pragma unroll
        for (int k = 0; k < LWIN + LPT - 1; k++) {
            xa[k] = s_x[r][c0 + k]; ya[k] = s_y[r][c0 + k];
            qa[k] = fmaf(xa[k], xa[k], ya[k] * ya[k]); pa[k] = xa[k] * ya[k];
        }
#pragma unroll
        for (int j = 0; j < LPT; j++) {
            float m1 = 0.f, m2 = 0.f, ess = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) {
                const float g = win.g[k];
                m1 = fmaf(g, xa[j + k], m1); m2 = fmaf(g, ya[j + k], m2);
                ess = fmaf(g, qa[j + k], ess); e12 = fmaf(g, pa[j + k], e12);
            }
            s_h[0][r][c0 + j] = m1; s_h[1][r][c0 + j] = m2; s_h[2][r][c0 + j] = ess; s_h[3][r][c0 + j] = e12;
        }
    }
    __syncthreads();
    // vertical: thread = (column tx, group of 4 rows)
    const int tx = tid & (LT - 1), r0 = (tid >> 5) * LPT;
    float acc[4][LPT];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        float col[LWIN + LPT - 1];
#pragma unroll
        for (int k = 0; k < LWIN + LPT - 1; k++) col[k] = s_h[m][r0 + k][tx];
#pragma unroll
        for (int j = 0; j < LPT; j++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) a = fmaf(win.g[k], col[j + k], a);
            acc[m][j] = a;
        }
    }
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < LPT; j++) {
        const int py = y0 + r0 + j;
        if (px < W && py < H) {
            const float x = s_x[r0 + j + LHALO][tx + LHALO], y = s_y[r0 + j + LHALO][tx + LHALO];
            l1 += fabsf(x - y);
            // reference utils/loss_utils.py:76-90
            const float mu1 = acc[0][j], mu2 = acc[1][j];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s12 = acc[3][j] - mu12;
            const float A1 = 2.f * mu12 + SSIM_C1, A2 = 2.f * s12 + SSIM_C2;
            const float B1 = mu1_sq + mu2_sq + SSIM_C1, B2 = (acc[2][j] - mu1_sq - mu2_sq) + SSIM_C2;   // sigma_1^2 + sigma_2^2 + C2
            const float iB1 = __frcp_rn(B1), iB2 = __frcp_rn(B2);   // B1 >= C1, B2 >= C2 up to rounding: well away from 0
            const float inv = iB1 * iB2;
            const float ss = A1 * A2 * inv;
            ss_sum += ss;
            if (WANT_GRAD) {
                // partials of S(mu1, E[x^2], E[xy]) with sigma_1^2 = E[x^2] - mu1^2, sigma_12 = E[xy] - mu1 mu2
                const size_t o = plane + (size_t)py * W + px;
                d_mu[o] = 2.f * (mu2 * (A2 - A1) * inv + mu1 * ss * (iB2 - iB1));
                d_e11[o] = -ss * iB2;
                d_e12[o] = 2.f * A1 * inv;
            }
        }
    }
    __syncthreads();      // the staged tile is dead: the next iteration overwrites it
    }   // tiles of this CTA
    // CTA sums -> partial[]; the last CTA to finish adds the partials up in a fixed order (deterministic)
    const int lane = tid & 31, wid = tid >> 5;
    double dl1 = warp_sum((double)l1), dss = warp_sum((double)ss_sum);
    if (lane == 0) { s_red[0][wid] = dl1; s_red[1][wid] = dss; }
    __syncthreads();
    const unsigned int n_ctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int me = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < LTHREADS / 32; k++) { a += s_red[0][k]; b += s_red[1][k]; }
        partial[me] = make_float2((float)a, (float)b);
        __threadfence();
        s_last = atomicAdd(counter, 1u) == n_ctas - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int i = tid; i < n_ctas; i += LTHREADS) {
        const float2 p = __ldcg(partial + i);
        a += (double)p.x;
        b += (double)p.y;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { s_red[0][wid] = a; s_red[1][wid] = b; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int k = 0; k < LTHREADS / 32; k++) { ta += s_red[0][k]; tb += s_red[1][k]; }
        const float ll1 = (float)(ta * inv_count), ssim = (float)(tb * inv_count);
        out3[0] = ll1;
        out3[1] = ssim;
        out3[2] = (1.f - lambda) * ll1 + lambda * (1.f - ssim);   // reference train.py:161
        *counter = 0u;                                              // ready for the next launch
    }
}

__global__ void __launch_bounds__(LTHREADS, 3)
photometric_backward_kernel(const float* __restrict__ img, const float* __restrict__ gt, const int H, const int W,
                            const Window win, const float* __restrict__ d_mu, const float* __restrict__ d_e11,
                            const float* __restrict__ d_e12, const float w_l1, const float w_ssim,
                            const float* __restrict__ upstream, float* __restrict__ dL_dimg) {
    __shared__ float s_m[3][LE][LSX];
    __shared__ float s_h[3][LE][LSH];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.y * LT;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tiles_x = (W + LT - 1) / LT;
    const int t_first = blockIdx.x * LTILES, t_end = min(tiles_x, t_first + LTILES);
    const float up = upstream ? __ldg(upstream) : 1.f;
    float pre[3][LSTAGE];
    int soff[LSTAGE], goff[LSTAGE], scol[LSTAGE];      // as in the forward: tile-independent staging indices
#pragma unroll
    for (int it = 0; it < LSTAGE; it++) {
        const int idx = tid + it * LTHREADS;
        const int r = idx / LE, c = idx - r * LE;
        const int gy = y0 + r - LHALO;
        scol[it] = c - LHALO;
        soff[it] = idx < LE * LE ? r * LSX + c : -1;
        goff[it] = (idx < LE * LE && gy >= 0 && gy < H) ? gy * W + c - LHALO : LNONE;   // (may be negative in row 0)
    }
    const float* __restrict__ m0 = d_mu + plane;
    const float* __restrict__ m1 = d_e11 + plane;
    const float* __restrict__ m2 = d_e12 + plane;
    auto fetch = [&](const int t) {
        const int x0 = t * LT;
#pragma unroll
        for (int it = 0; it < LSTAGE; it++) {
            const int gx = x0 + scol[it];
            float a = 0.f, b = 0.f, d = 0.f;
            if (goff[it] != LNONE && gx >= 0 && gx < W) {
                a = __ldg(m0 + goff[it] + x0); b = __ldg(m1 + goff[it] + x0); d = __ldg(m2 + goff[it] + x0);
            }
            pre[0][it] = a; pre[1][it] = b; pre[2][it] = d;
        }
    };
    fetch(t_first);
    for (int t = t_first; t < t_end; t++) {
    const int x0 = t * LT;
#pragma unroll
    for (int it = 0; it < LSTAGE; it++) {
        if (soff[it] >= 0) {
            (&s_m[0][0][0])[soff[it]] = pre[0][it];
            (&s_m[1][0][0])[soff[it]] = pre[1][it];
            (&s_m[2][0][0])[soff[it]] = pre[2][it];
        }
    }
    __syncthreads();
    if (t + 1 < t_end) fetch(t + 1);
    for (int idx = tid; idx < LE * (LT / LPT); idx += LTHREADS) {
        const int r = idx / (LT / LPT), c0 = (idx - r * (LT / LPT)) * LPT;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float in[LWIN + LPT - 1];
#pragma unroll
            for (int k = 0; k < LWIN + LPT - 1; k++) in[k] = s_m[m][r][c0 + k];
#pragma unroll
            for (int j = 0; j < LPT; j++) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < LWIN; k++) a = fmaf(win.g[k], in[j + k], a);
                s_h[m][r][c0 + j] = a;
            }
        }
    }
    __syncthreads();
    const int tx = tid & (LT - 1), r0 = (tid >> 5) * LPT;
    float acc[3][LPT];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        float col[LWIN + LPT - 1];
#pragma unroll
        for (int k = 0; k < LWIN + LPT - 1; k++) col[k] = s_h[m][r0 + k][tx];
#pragma unroll
        for (int j = 0; j < LPT; j++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < LWIN; k++) a = fmaf(win.g[k], col[j + k], a);
            acc[m][j] = a;
        }
    }
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < LPT; j++) {
        const int py = y0 + r0 + j;
        if (px < W && py < H) {
            const size_t o = plane + (size_t)py * W + px;
            const float x = __ldg(img + o), y = __ldg(gt + o);
            const float diff = x - y;
            const float sgn = (diff > 0.f ? 1.f : 0.f) - (diff < 0.f ? 1.f : 0.f);   // d|u|/du, 0 at 0 (torch.abs)
            dL_dimg[o] = up * (w_ssim * (acc[0][j] + 2.f * x * acc[1][j] + y * acc[2][j]) + w_l1 * sgn);
        }
    }
    __syncthreads();      // the staged tile is dead: the next iteration overwrites it
    }   // tiles of this CTA
}

}  // namespace

size_t loss_scratch_bytes(int64_t elems, int64_t ctas) {
    return align_up((size_t)elems * 4) * 3 + align_up((size_t)ctas * sizeof(float2)) + 256;
}

struct LossLayout {
    float *d_mu, *d_e11, *d_e12;
    float2* partial;
    unsigned int* counter;
};
static LossLayout carve_loss(void* base, int64_t elems, int64_t ctas) {
    LossLayout L;
    char* b = (char*)base;
    size_t o = 0;
    auto take = [&](size_t n) { char* p = b + o; o += align_up(n); return (void*)p; };
    L.d_mu = (float*)take((size_t)elems * 4);
    L.d_e11 = (float*)take((size_t)elems * 4);
    L.d_e12 = (float*)take((size_t)elems * 4);
    L.partial = (float2*)take((size_t)ctas * sizeof(float2));
    L.counter = (unsigned int*)take(256);
    return L;
}

static dim3 loss_grid(int C, int H, int W) {
    const int tiles_x = (W + LT - 1) / LT;
    return dim3((tiles_x + LTILES - 1) / LTILES, (H + LT - 1) / LT, C);
}

size_t photometric_scratch_bytes(int C, int H, int W) {
    const dim3 g = loss_grid(C, H, W);
    return loss_scratch_bytes((int64_t)C * H * W, (int64_t)g.x * g.y * g.z);
}

void launch_photometric_forward(const float* img, const float* gt, int C, int H, int W, float lambda, void* scratch,
                                bool want_grad, float* out3, const Launch& L) {
    static const Window win = make_window();
    const dim3 g = loss_grid(C, H, W);
    const LossLayout S = carve_loss(scratch, (int64_t)C * H * W, (int64_t)g.x * g.y * g.z);
    cudaMemsetAsync(S.counter, 0, sizeof(unsigned int), L.stream);
    const double inv_count = 1.0 / ((double)C * H * W);
    begin_kernel("photometric_forward", L);
    if (want_grad)
        photometric_forward_kernel<true><<<g, LTHREADS, 0, L.stream>>>(img, gt, H, W, win, S.d_mu, S.d_e11, S.d_e12, S.partial,
                                                                     S.counter, lambda, inv_count, out3);
    else
        photometric_forward_kernel<false><<<g, LTHREADS, 0, L.stream>>>(img, gt, H, W, win, S.d_mu, S.d_e11, S.d_e12, S.partial,
                                                                      S.counter, lambda, inv_count, out3);
    check_launch("photometric_forward", L);
}

void launch_photometric_backward(const float* img, const float* gt, int C, int H, int W, float lambda,
                                 const void* scratch, const float* upstream, float* dL_dimg, const Launch& L) {
    static const Window win = make_window();
    const dim3 g = loss_grid(C, H, W);
    const LossLayout S = carve_loss(const_cast<void*>(scratch), (int64_t)C * H * W, (int64_t)g.x * g.y * g.z);
    const float inv_count = (float)(1.0 / ((double)C * H * W));
    begin_kernel("photometric_backward", L);
    // loss = (1 - lambda) mean|x - y| + lambda (1 - mean S)
    photometric_backward_kernel<<<g, LTHREADS, 0, L.stream>>>(img, gt, H, W, win, S.d_mu, S.d_e11, S.d_e12,
                                                             (1.f - lambda) * inv_count, -lambda * inv_count, upstream, dL_dimg);
    check_launch("photometric_backward", L);
}

}  // namespace scgr
