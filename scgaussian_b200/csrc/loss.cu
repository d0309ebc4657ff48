// loss.cu -- fused photometric loss, forward + backward (sm_100a).  SURVEY.md section 8(f), row f1.
//
// Replaces what reference train.py:160-161 computes with a dozen torch kernels per direction:
//     Ll1  = l1_loss(image, gt)                                  reference utils/loss_utils.py:40-41
//     loss = (1 - lambda) * Ll1 + lambda * (1 - ssim(image, gt)) reference utils/loss_utils.py:56-94
// ssim(): 11x11 Gaussian window (sigma 1.5, reference :46-54), zero padding 5, depthwise, C1 = 0.01^2,
// C2 = 0.03^2, mean over every element.  The 2-D window is the outer product of the 1-D one (:52), so
// the five windowed moments E[x], E[y], E[x^2], E[y^2], E[xy] are computed separably.
//
// Two kernel designs compute the same thing: the streaming column-strip kernels further down (the default,
// SCGR_LOSS_VARIANT=1: 56 / 49 us at 1080p on B200) and the round-1 tile kernels described here (SCGR_LOSS_VARIANT=0:
// 88 / 78 us), kept as the measured alternative and as a second implementation for the tests.
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
// Streaming fp32 work bounded by instruction issue (the separable 11-tap filter costs 88 / 66 FMA per pixel): no tensor cores.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>

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


// =====================================================================================================================
// Streaming variant (SCGR_LOSS_VARIANT=1).  The tile kernels above filter a 42x42 staged patch for every 32x32 outputs:
// (42/32)^2 = 1.7x the loads and 1.3x the horizontal work, every vertical tap a shared-memory load, ~330 instructions
// per output pixel, and they are bound by the issue rate (73 % issue-active).  Here a CTA owns a column strip
// (SWO = 116 output columns + 5 halo columns each side: one staged column per thread) of a band of image rows and walks
// DOWN it.  Vertical pass: the thread of a column keeps the 11 partial sums that the row it has just loaded takes part
// in -- a ring of 11 accumulators per moment in registers, the loop over rows unrolled by 11 so that every ring index is
// a compile-time constant -- i.e. one FMA per tap and moment with NO shared-memory traffic and a vertical halo only at
// the two ends of the band.  Every 11 rows the finished sums of those rows go through shared memory (double-buffered:
// one barrier per batch) for the horizontal pass, where a thread produces 4 adjacent outputs of one row from 14 values
// fetched as 128-bit loads, evaluates SSIM / the gradient sum, and reads / writes the image rows as float4.  The global
// loads of the next 11 rows are issued before the horizontal pass of the current ones.
constexpr int ST = 128;                  // threads per CTA = staged columns (126 in use)
constexpr int SWO = 116;                 // output columns per strip: 29 groups of 4
constexpr int SWS = SWO + 2 * LHALO;     // staged columns
constexpr int SRB = LWIN;                // rows per batch = period of the accumulator ring
constexpr int SVS = 132;                 // row stride of the vertically filtered rows (a multiple of 4 floats)
constexpr int SG = SWO / LPT;            // column groups of the horizontal pass
constexpr int SOCC = 4;                  // CTAs per SM the kernels are built for (__launch_bounds__)
static_assert(SWO % LPT == 0 && SWS <= ST && SVS >= ST && SVS % 4 == 0, "strip geometry");

struct SsimPoint {
    float ss, d_mu, d_e11, d_e12;
};
// reference utils/loss_utils.py:76-90 at one pixel from the windowed moments E[x], E[y], E[x^2 + y^2], E[xy], and the
// partial derivatives of the value with respect to E[x], E[x^2], E[xy] (sigma_1^2 = E[x^2] - mu1^2, sigma_12 = E[xy] - mu1 mu2)
__device__ __forceinline__ SsimPoint ssim_point(const float mu1, const float mu2, const float ess, const float e12) {
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s12 = e12 - mu12;
    const float A1 = 2.f * mu12 + SSIM_C1, A2 = 2.f * s12 + SSIM_C2;
    const float B1 = mu1_sq + mu2_sq + SSIM_C1, B2 = (ess - mu1_sq - mu2_sq) + SSIM_C2;
    // one IEEE reciprocal instead of two: B1 B2 >= C1 C2 = 9e-8 up to rounding, far from underflow
    const float inv = __frcp_rn(B1 * B2);
    const float iB1 = inv * B2, iB2 = inv * B1;
    SsimPoint o;
    o.ss = A1 * A2 * inv;
    o.d_mu = 2.f * (mu2 * (A2 - A1) * inv + mu1 * o.ss * (iB2 - iB1));
    o.d_e11 = -o.ss * iB2;
    o.d_e12 = 2.f * A1 * inv;
    return o;
}

// 14 consecutive floats starting at a 16-byte aligned shared-memory address
__device__ __forceinline__ void load14(const float* __restrict__ src, float (&in)[LWIN + LPT - 1]) {
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    const float4 c = *reinterpret_cast<const float4*>(src + 8);
    const float2 d = *reinterpret_cast<const float2*>(src + 12);
    in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w; in[4] = b.x; in[5] = b.y; in[6] = b.z; in[7] = b.w;
    in[8] = c.x; in[9] = c.y; in[10] = c.z; in[11] = c.w; in[12] = d.x; in[13] = d.y;
}

// 4 adjacent elements of an image row: one 128-bit access when the row layout allows it (`vec`), else element-wise
__device__ __forceinline__ void load4(const float* __restrict__ p, const int n_valid, const bool vec, float (&v)[LPT]) {
    if (vec && n_valid == LPT) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < LPT; j++) v[j] = j < n_valid ? __ldg(p + j) : 0.f;
    }
}
__device__ __forceinline__ void store4(float* __restrict__ p, const int n_valid, const bool vec, const float (&v)[LPT]) {
    if (vec && n_valid == LPT) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int j = 0; j < LPT; j++)
            if (j < n_valid) p[j] = v[j];
    }
}

// Batch b of a band starting at image row ya covers the input rows ya - 6 + 11 b + r, r = 0..10 (the first one of batch 0
// is not needed: the band's first output row ya needs ya - 5 .. ya + 5), and -- for b >= 1 -- finishes the output rows
// ya + 11 (b - 1) + r: the ring slot of the output that row r completes is (r + 1) % 11, the slot the row opens is r.
template <bool WANT_GRAD>
__global__ void __launch_bounds__(ST, SOCC)
photometric_forward_stream_kernel(const float* __restrict__ img, const float* __restrict__ gt, const int H, const int W,
                                  const int band, const int vec, const Window win, float* __restrict__ d_mu,
                                  float* __restrict__ d_e11, float* __restrict__ d_e12, float2* __restrict__ partial,
                                  unsigned int* __restrict__ counter, const float lambda, const double inv_count,
                                  float* __restrict__ out3) {
    __shared__ __align__(16) float s_v[2][4][SRB][SVS];      // E[x], E[y], E[x^2 + y^2], E[xy] after the vertical pass
    __shared__ double s_red[2][ST / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SWO;
    const int ya = blockIdx.y * band, yb = min(H, ya + band);
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* __restrict__ ip = img + plane;
    const float* __restrict__ gp = gt + plane;
    const int gx = x0 - LHALO + tid;
    const bool col_ok = tid < SWS && gx >= 0 && gx < W;
    const int nb = (yb - ya + SRB - 1) / SRB;

    float acc[4][SRB];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
        for (int k = 0; k < SRB; k++) acc[m][k] = 0.f;
    float pa[SRB], pb[SRB];
    auto fetch = [&](const int b) {
#pragma unroll
        for (int r = 0; r < SRB; r++) {
            const int gy = ya - (LHALO + 1) + SRB * b + r;
            float a = 0.f, c = 0.f;
            if (col_ok && gy >= 0 && gy < H && gy >= ya - LHALO && gy < yb + LHALO) {
                a = __ldg(ip + (size_t)gy * W + gx);
                c = __ldg(gp + (size_t)gy * W + gx);
            }
            pa[r] = a;
            pb[r] = c;
        }
    };
    float l1 = 0.f, ss_sum = 0.f;
    fetch(0);
    for (int b = 0; b <= nb; b++) {
        float (*sv)[SRB][SVS] = s_v[b & 1];
#pragma unroll
        for (int r = 0; r < SRB; r++) {
            const float x = pa[r], y = pb[r];
            const float v[4] = {x, y, fmaf(x, x, y * y), x * y};
#pragma unroll
            for (int m = 0; m < 4; m++) {
                acc[m][r] = win.g[0] * v[m];
#pragma unroll
                for (int k = 1; k < LWIN; k++) acc[m][(r - k + SRB) % SRB] = fmaf(win.g[k], v[m], acc[m][(r - k + SRB) % SRB]);
            }
            if (b > 0) {
#pragma unroll
                for (int m = 0; m < 4; m++) sv[m][r][tid] = acc[m][(r + 1) % SRB];
            }
        }
        if (b < nb) fetch(b + 1);
        if (b == 0) continue;
        __syncthreads();
        // horizontal: item = (row r of the batch, group of 4 output columns)
        for (int it = tid; it < SRB * SG; it += ST) {
            const int r = it / SG, g4 = (it - r * SG) * LPT;
            const int py = ya + SRB * (b - 1) + r, px = x0 + g4;
            if (py >= yb || px >= W) continue;
            // the centre pixels come from L2 (the rows left L1 a batch ago): requested before the filter, used after it
            const int n_valid = min(LPT, W - px);
            const size_t off = (size_t)py * W + px;
            float xs[LPT], ys[LPT];
            load4(ip + off, n_valid, vec != 0, xs);
            load4(gp + off, n_valid, vec != 0, ys);
            float o[4][LPT];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                float in[LWIN + LPT - 1];
                load14(&sv[m][r][g4], in);
#pragma unroll
                for (int j = 0; j < LPT; j++) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < LWIN; k++) a = fmaf(win.g[k], in[j + k], a);
                    o[m][j] = a;
                }
            }
            float g_mu[LPT], g_e11[LPT], g_e12[LPT];
#pragma unroll
            for (int j = 0; j < LPT; j++) {
                const SsimPoint s = ssim_point(o[0][j], o[1][j], o[2][j], o[3][j]);
                if (j < n_valid) {
                    l1 += fabsf(xs[j] - ys[j]);
                    ss_sum += s.ss;
                }
                g_mu[j] = s.d_mu; g_e11[j] = s.d_e11; g_e12[j] = s.d_e12;
            }
            if (WANT_GRAD) {
                store4(d_mu + plane + off, n_valid, vec != 0, g_mu);
                store4(d_e11 + plane + off, n_valid, vec != 0, g_e11);
                store4(d_e12 + plane + off, n_valid, vec != 0, g_e12);
            }
        }
        // no second barrier: the next batch fills the other buffer, and the barrier behind it orders this pass
        // before the batch after that overwrites this one
    }
    // CTA sums -> partial[]; the last CTA to finish adds the partials up in a fixed order (deterministic)
    const int lane = tid & 31, wid = tid >> 5;
    double dl1 = warp_sum((double)l1), dss = warp_sum((double)ss_sum);
    if (lane == 0) { s_red[0][wid] = dl1; s_red[1][wid] = dss; }
    __syncthreads();
    const unsigned int n_ctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int me = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double a = 0.0, c = 0.0;
#pragma unroll
        for (int k = 0; k < ST / 32; k++) { a += s_red[0][k]; c += s_red[1][k]; }
        partial[me] = make_float2((float)a, (float)c);
        __threadfence();
        s_last = atomicAdd(counter, 1u) == n_ctas - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double a = 0.0, c = 0.0;
    for (unsigned int i = tid; i < n_ctas; i += ST) {
        const float2 p = __ldcg(partial + i);
        a += (double)p.x;
        c += (double)p.y;
    }
    a = warp_sum(a);
    c = warp_sum(c);
    __syncthreads();
    if (lane == 0) { s_red[0][wid] = a; s_red[1][wid] = c; }
    __syncthreads();
    if (tid == 0) {
        double ta = 0.0, tc = 0.0;
#pragma unroll
        for (int k = 0; k < ST / 32; k++) { ta += s_red[0][k]; tc += s_red[1][k]; }
        const float ll1 = (float)(ta * inv_count), ssim = (float)(tc * inv_count);
        out3[0] = ll1;
        out3[1] = ssim;
        out3[2] = (1.f - lambda) * ll1 + lambda * (1.f - ssim);   // reference train.py:161
        *counter = 0u;
    }
}

__global__ void __launch_bounds__(ST, SOCC)
photometric_backward_stream_kernel(const float* __restrict__ img, const float* __restrict__ gt, const int H, const int W,
                                   const int band, const int vec, const Window win, const float* __restrict__ d_mu,
                                   const float* __restrict__ d_e11, const float* __restrict__ d_e12, const float w_l1,
                                   const float w_ssim, const float* __restrict__ upstream, float* __restrict__ dL_dimg) {
    __shared__ __align__(16) float s_v[2][3][SRB][SVS];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SWO;
    const int ya = blockIdx.y * band, yb = min(H, ya + band);
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* __restrict__ m0 = d_mu + plane;
    const float* __restrict__ m1 = d_e11 + plane;
    const float* __restrict__ m2 = d_e12 + plane;
    const int gx = x0 - LHALO + tid;
    const bool col_ok = tid < SWS && gx >= 0 && gx < W;
    const int nb = (yb - ya + SRB - 1) / SRB;
    const float up = upstream ? __ldg(upstream) : 1.f;

    float acc[3][SRB];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int k = 0; k < SRB; k++) acc[m][k] = 0.f;
    float pre[3][SRB];
    auto fetch = [&](const int b) {
#pragma unroll
        for (int r = 0; r < SRB; r++) {
            const int gy = ya - (LHALO + 1) + SRB * b + r;
            float a = 0.f, c = 0.f, d = 0.f;
            if (col_ok && gy >= 0 && gy < H && gy >= ya - LHALO && gy < yb + LHALO) {
                const size_t o = (size_t)gy * W + gx;
                a = __ldg(m0 + o); c = __ldg(m1 + o); d = __ldg(m2 + o);
            }
            pre[0][r] = a; pre[1][r] = c; pre[2][r] = d;
        }
    };
    fetch(0);
    for (int b = 0; b <= nb; b++) {
        float (*sv)[SRB][SVS] = s_v[b & 1];
#pragma unroll
        for (int r = 0; r < SRB; r++) {
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const float v = pre[m][r];
                acc[m][r] = win.g[0] * v;
#pragma unroll
                for (int k = 1; k < LWIN; k++) acc[m][(r - k + SRB) % SRB] = fmaf(win.g[k], v, acc[m][(r - k + SRB) % SRB]);
            }
            if (b > 0) {
#pragma unroll
                for (int m = 0; m < 3; m++) sv[m][r][tid] = acc[m][(r + 1) % SRB];
            }
        }
        if (b < nb) fetch(b + 1);
        if (b == 0) continue;
        // The centre pixels of an item come from DRAM (nothing else in this kernel reads the images): those of the thread's
        // first item are requested before the barrier, those of every further item before the previous item is filtered.
        float xs[LPT] = {0.f, 0.f, 0.f, 0.f}, ys[LPT] = {0.f, 0.f, 0.f, 0.f};
        auto centre = [&](const int it) {
            const int r = it / SG, g4 = (it - r * SG) * LPT;
            const int py = ya + SRB * (b - 1) + r, px = x0 + g4;
            if (it < SRB * SG && py < yb && px < W) {
                const size_t off = plane + (size_t)py * W + px;
                load4(img + off, min(LPT, W - px), vec != 0, xs);
                load4(gt + off, min(LPT, W - px), vec != 0, ys);
            }
        };
        centre(tid);
        __syncthreads();
        for (int it = tid; it < SRB * SG; it += ST) {
            const int r = it / SG, g4 = (it - r * SG) * LPT;
            const int py = ya + SRB * (b - 1) + r, px = x0 + g4;
            float xc[LPT], yc[LPT];
#pragma unroll
            for (int j = 0; j < LPT; j++) { xc[j] = xs[j]; yc[j] = ys[j]; }
            centre(it + ST);
            if (py >= yb || px >= W) continue;
            const int n_valid = min(LPT, W - px);
            const size_t off = plane + (size_t)py * W + px;
            float out[LPT];
            float o[3][LPT];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                float in[LWIN + LPT - 1];
                load14(&sv[m][r][g4], in);
#pragma unroll
                for (int j = 0; j < LPT; j++) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < LWIN; k++) a = fmaf(win.g[k], in[j + k], a);
                    o[m][j] = a;
                }
            }
#pragma unroll
            for (int j = 0; j < LPT; j++) {
                const float diff = xc[j] - yc[j];
                const float sgn = (diff > 0.f ? 1.f : 0.f) - (diff < 0.f ? 1.f : 0.f);   // d|u|/du, 0 at 0 (torch.abs)
                out[j] = up * (w_ssim * (o[0][j] + 2.f * xc[j] * o[1][j] + yc[j] * o[2][j]) + w_l1 * sgn);
            }
            store4(dL_dimg + off, n_valid, vec != 0, out);
        }
    }
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

// Streaming variant: grid.x = column strips, grid.y = bands of `band` rows (a multiple of 11), grid.z = planes.  The
// bands are as tall as they can be with every CTA resident at once (SOCC per SM): the vertical halo (10 rows per band)
// is the only redundant work, and a band below 33 rows is never worth it.
struct StreamGrid {
    dim3 grid;
    int band;
};
static StreamGrid stream_grid(int C, int H, int W) {
    static const int sm_count = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    const int strips = (W + SWO - 1) / SWO;
    const int64_t columns = (int64_t)C * strips;
    int want = (int)std::max<int64_t>(1, (int64_t)sm_count * SOCC / columns);           // bands that fit in one wave
    if (const char* e = getenv("SCGR_LOSS_BANDS")) want = std::max(1, atoi(e));           // tuning switch
    // measured on B200: one full wave (1080p: 11 bands of 99 rows) and several waves of 99-row bands (4K: 22 bands) beat
    // everything in between (1.1 - 1.4 waves: tail) and taller bands (fewer CTAs than slots: idle SMs)
    const int batches = std::min(9, std::max(3, (H + want * SRB - 1) / (want * SRB)));
    StreamGrid g;
    g.band = batches * SRB;
    g.grid = dim3(strips, (H + g.band - 1) / g.band, C);
    return g;
}

// 1 (default): streaming column strips; 0: 32x32 tiles.  Read on every launch (tests switch it between calls).
// B200, 1080p, end-to-end step of bench.py: 1.351 -> 1.291 ms with the streaming kernels (4K: forward 352 -> 235 us,
// backward 294 -> 169 us).
static int loss_variant() {
    const char* e = getenv("SCGR_LOSS_VARIANT");
    return e ? atoi(e) : 1;
}
static bool rows_vectorisable(int W, const void* a, const void* b, const void* c) {
    return W % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

// CTAs of the larger of the two grids: the scratch layout is the same whichever variant runs
static int64_t loss_ctas(int C, int H, int W) {
    const dim3 g = loss_grid(C, H, W), gs = stream_grid(C, H, W).grid;
    return std::max((int64_t)g.x * g.y * g.z, (int64_t)gs.x * gs.y * gs.z);
}

size_t photometric_scratch_bytes(int C, int H, int W) {
    return loss_scratch_bytes((int64_t)C * H * W, loss_ctas(C, H, W));
}

void launch_photometric_forward(const float* img, const float* gt, int C, int H, int W, float lambda, void* scratch,
                                bool want_grad, float* out3, const Launch& L) {
    static const Window win = make_window();
    const dim3 g = loss_grid(C, H, W);
    const StreamGrid sg = stream_grid(C, H, W);
    const LossLayout S = carve_loss(scratch, (int64_t)C * H * W, loss_ctas(C, H, W));
    cudaMemsetAsync(S.counter, 0, sizeof(unsigned int), L.stream);
    const double inv_count = 1.0 / ((double)C * H * W);
    begin_kernel("photometric_forward", L);
    if (loss_variant() == 1) {
        const int vec = rows_vectorisable(W, img, gt, S.d_mu);
        if (want_grad)
            photometric_forward_stream_kernel<true><<<sg.grid, ST, 0, L.stream>>>(img, gt, H, W, sg.band, vec, win, S.d_mu, S.d_e11,
                                                                                S.d_e12, S.partial, S.counter, lambda, inv_count, out3);
        else
            photometric_forward_stream_kernel<false><<<sg.grid, ST, 0, L.stream>>>(img, gt, H, W, sg.band, vec, win, S.d_mu, S.d_e11,
                                                                                 S.d_e12, S.partial, S.counter, lambda, inv_count, out3);
    } else if (want_grad)
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
    const StreamGrid sg = stream_grid(C, H, W);
    const LossLayout S = carve_loss(const_cast<void*>(scratch), (int64_t)C * H * W, loss_ctas(C, H, W));
    const float inv_count = (float)(1.0 / ((double)C * H * W));
    begin_kernel("photometric_backward", L);
    // loss = (1 - lambda) mean|x - y| + lambda (1 - mean S)
    if (loss_variant() == 1)
        photometric_backward_stream_kernel<<<sg.grid, ST, 0, L.stream>>>(img, gt, H, W, sg.band,
                                                                        (int)(rows_vectorisable(W, img, gt, dL_dimg) && rows_vectorisable(W, S.d_mu, S.d_e11, S.d_e12)),
                                                                        win, S.d_mu, S.d_e11, S.d_e12, (1.f - lambda) * inv_count,
                                                                        -lambda * inv_count, upstream, dL_dimg);
    else
        photometric_backward_kernel<<<g, LTHREADS, 0, L.stream>>>(img, gt, H, W, win, S.d_mu, S.d_e11, S.d_e12,
                                                                 (1.f - lambda) * inv_count, -lambda * inv_count, upstream, dL_dimg);
    check_launch("photometric_backward", L);
}

}  // namespace scgr
