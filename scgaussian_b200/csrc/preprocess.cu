// preprocess.cu -- per-Gaussian stages of the rasterizer, forward and backward (sm_100a).
//
// Replaces the external rasterizer's FORWARD::preprocessCUDA, BACKWARD::computeCov2DCUDA,
// BACKWARD::preprocessCUDA and checkFrustum, i.e. the
// per-Gaussian half of what reference gaussian_renderer/__init__.py:100-108 invokes.  Semantics:
// SURVEY.md Appendix A.1-A.5 (forward), A.10 (backward).  The SH polynomial is the one of
// reference utils/sh_utils.py:74-100; R(q) that of reference utils/general_utils.py:96-104.
//
// B200 design: both kernels are HBM-bound streaming passes (311 B / 579 B of compulsory traffic
// per Gaussian).  One thread per Gaussian, 128-thread CTAs; the 192-byte SH row of each Gaussian
// (and its gradient) is moved between HBM and shared memory by the whole CTA with 128-bit,
// fully-coalesced accesses into a padded (13 x float4 per row) layout that is bank-conflict-free
// for the per-thread 128-bit reads/writes; the backward fuses cov2D-, projection-, depth-, SH-
// and cov3D-backward in one pass and writes every gradient tensor in full (zeros for culled
// Gaussians), so the host never memsets them.
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "common.cuh"

namespace scgr {

namespace {

constexpr int PRE_THREADS = 128;
#ifndef SCGR_PREB_THREADS
#define SCGR_PREB_THREADS 64
#endif
#ifndef SCGR_PREB_GAUSS
#define SCGR_PREB_GAUSS 128
#endif
constexpr int PB_G = SCGR_PREB_GAUSS;      // backward: Gaussians classified per CTA
constexpr int PB_T = SCGR_PREB_THREADS;    // backward: threads per CTA = live Gaussians processed per round
constexpr int PB_K = PB_G / PB_T;
static_assert(PB_G % PB_T == 0 && PB_T % 32 == 0 && PB_G <= 256, "backward block shape");
constexpr int SH_ROW_F4 = 12;       // 48 floats = 12 float4 per Gaussian at M = 16
constexpr int SH_ROW_F4_PAD = 13;   // padded row stride (float4 units): conflict-free LDS.128/STS.128
constexpr int SH_ROW_PAD = 4 * SH_ROW_F4_PAD;      // the same in floats
constexpr int SH_DC = 3, SH_REST = 45;              // floats per Gaussian in features_dc / features_rest at 16 coefficients

// ---- split SH layout (ScgrGaussians.sh_dc / sh_rest: the hybrid model's own four arrays, include/scgr.h) ----
// Rows [row0, row0 + nrows) of the virtual [P,16,3] array are up to two runs, one per set; inside a set features_dc and
// features_rest rows are contiguous, so the CTA walks them as flat streams, one float per lane on consecutive addresses
// (rows of 3 / 45 floats have no 16-byte alignment), and lands them in the padded staging layout the SH_FAST path uses.
template <typename F>
__device__ __forceinline__ void for_each_split_run(const ScgrGaussians& g, const int row0, const int nrows, F&& f) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int lo = k == 0 ? row0 : max(row0, g.sh_n0), hi = k == 0 ? min(row0 + nrows, g.sh_n0) : row0 + nrows;
        if (hi > lo) f(k, lo - (k ? g.sh_n0 : 0), hi - lo, lo - row0);      // set (compile-time after unrolling), first row inside the set, rows, first local row
    }
}

__constant__ const float kC0 = 0.28209479177387814f;
__constant__ const float kC1 = 0.4886025119029199f;
__device__ const float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};

struct Camera {
    float V[16];   // viewmatrix tensor, row-major: p_view_j = sum_k hom_k V[k*4+j]
    float PM[16];  // projmatrix tensor, same convention
    float cam[3];
};

__device__ __forceinline__ void load_camera(const ScgrView& v, Camera& c) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        c.V[i] = __ldg(v.viewmatrix + i);
        c.PM[i] = __ldg(v.projmatrix + i);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) c.cam[i] = __ldg(v.campos + i);
}

// Everything the EWA projection of one Gaussian produces; shared by forward and backward.
struct Proj {
    float3 t;        // view-space point
    float tx, ty;    // after the 1.3*tanfov clamp
    bool xin, yin;   // clamp inactive
    float3 M0, M1;   // rows of J * W3
    float a, b, c;   // dilated 2D covariance
    float fx, fy;
};

__device__ __forceinline__ void quat_to_R(const float4 q, float R[9]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = (R S)(R S)^T as 6 floats xx,xy,xz,yy,yz,zz  (A.3)
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 s, const float mod, const float4 q,
                                                     float c6[6]) {
    float R[9];
    quat_to_R(q, R);
    const float sx = mod * s.x, sy = mod * s.y, sz = mod * s.z;
    float L[9];
#pragma unroll
    for (int i = 0; i < 3; i++) { L[i * 3] = R[i * 3] * sx; L[i * 3 + 1] = R[i * 3 + 1] * sy; L[i * 3 + 2] = R[i * 3 + 2] * sz; }
    c6[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
    c6[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
    c6[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
    c6[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
    c6[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
    c6[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
}

__device__ __forceinline__ void project_cov(const Camera& cam, const ScgrView& v, const float3 p,
                                            const float c6[6], Proj& o) {
    const float* V = cam.V;
    o.t.x = p.x * V[0] + p.y * V[4] + p.z * V[8] + V[12];
    o.t.y = p.x * V[1] + p.y * V[5] + p.z * V[9] + V[13];
    o.t.z = p.x * V[2] + p.y * V[6] + p.z * V[10] + V[14];
    o.fx = v.image_width / (2.f * v.tanfovx);
    o.fy = v.image_height / (2.f * v.tanfovy);
    const float limx = 1.3f * v.tanfovx, limy = 1.3f * v.tanfovy;
    const float tz = o.t.z;
    const float txtz = o.t.x / tz, tytz = o.t.y / tz;
    o.xin = (txtz >= -limx) && (txtz <= limx);
    o.yin = (tytz >= -limy) && (tytz <= limy);
    o.tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    o.ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float j00 = o.fx / tz, j02 = -o.fx * o.tx / (tz * tz);
    const float j11 = o.fy / tz, j12 = -o.fy * o.ty / (tz * tz);
    // W3[r][k] = V[k*4+r];  M = J W3
    o.M0 = make_float3(j00 * V[0] + j02 * V[2], j00 * V[4] + j02 * V[6], j00 * V[8] + j02 * V[10]);
    o.M1 = make_float3(j11 * V[1] + j12 * V[2], j11 * V[5] + j12 * V[6], j11 * V[9] + j12 * V[10]);
    const float s0x = c6[0] * o.M0.x + c6[1] * o.M0.y + c6[2] * o.M0.z;
    const float s0y = c6[1] * o.M0.x + c6[3] * o.M0.y + c6[4] * o.M0.z;
    const float s0z = c6[2] * o.M0.x + c6[4] * o.M0.y + c6[5] * o.M0.z;
    const float s1x = c6[0] * o.M1.x + c6[1] * o.M1.y + c6[2] * o.M1.z;
    const float s1y = c6[1] * o.M1.x + c6[3] * o.M1.y + c6[4] * o.M1.z;
    const float s1z = c6[2] * o.M1.x + c6[4] * o.M1.y + c6[5] * o.M1.z;
    o.a = o.M0.x * s0x + o.M0.y * s0y + o.M0.z * s0z + DILATION;
    o.b = o.M0.x * s1x + o.M0.y * s1y + o.M0.z * s1z;
    o.c = o.M1.x * s1x + o.M1.y * s1y + o.M1.z * s1z + DILATION;
}

// SH basis b[k] at unit direction d, for k < (D+1)^2  (reference utils/sh_utils.py:74-100)
__device__ __forceinline__ void sh_basis(const int D, const float3 d, float b[16]) {
    b[0] = kC0;
    if (D < 1) return;
    const float x = d.x, y = d.y, z = d.z;
    b[1] = -kC1 * y; b[2] = kC1 * z; b[3] = -kC1 * x;
    if (D < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = kC2[0] * xy; b[5] = kC2[1] * yz; b[6] = kC2[2] * (2.f * zz - xx - yy);
    b[7] = kC2[3] * xz; b[8] = kC2[4] * (xx - yy);
    if (D < 3) return;
    b[9] = kC3[0] * y * (3.f * xx - yy); b[10] = kC3[1] * xy * z;
    b[11] = kC3[2] * y * (4.f * zz - xx - yy); b[12] = kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[13] = kC3[4] * x * (4.f * zz - xx - yy); b[14] = kC3[5] * z * (xx - yy);
    b[15] = kC3[6] * x * (xx - 3.f * yy);
}

// d(basis)/d(direction): gx[k], gy[k], gz[k]
__device__ __forceinline__ void sh_basis_grad(const int D, const float3 d, float gx[16], float gy[16],
                                              float gz[16]) {
#pragma unroll
    for (int k = 0; k < 16; k++) { gx[k] = 0.f; gy[k] = 0.f; gz[k] = 0.f; }
    if (D < 1) return;
    const float x = d.x, y = d.y, z = d.z;
    gy[1] = -kC1; gz[2] = kC1; gx[3] = -kC1;
    if (D < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    gx[4] = kC2[0] * y; gy[4] = kC2[0] * x;
    gy[5] = kC2[1] * z; gz[5] = kC2[1] * y;
    gx[6] = kC2[2] * -2.f * x; gy[6] = kC2[2] * -2.f * y; gz[6] = kC2[2] * 4.f * z;
    gx[7] = kC2[3] * z; gz[7] = kC2[3] * x;
    gx[8] = kC2[4] * 2.f * x; gy[8] = kC2[4] * -2.f * y;
    if (D < 3) return;
    gx[9] = kC3[0] * 6.f * xy; gy[9] = kC3[0] * (3.f * xx - 3.f * yy);
    gx[10] = kC3[1] * yz; gy[10] = kC3[1] * xz; gz[10] = kC3[1] * xy;
    gx[11] = kC3[2] * -2.f * xy; gy[11] = kC3[2] * (4.f * zz - xx - 3.f * yy); gz[11] = kC3[2] * 8.f * yz;
    gx[12] = kC3[3] * -6.f * xz; gy[12] = kC3[3] * -6.f * yz; gz[12] = kC3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
    gx[13] = kC3[4] * (4.f * zz - 3.f * xx - yy); gy[13] = kC3[4] * -2.f * xy; gz[13] = kC3[4] * 8.f * xz;
    gx[14] = kC3[5] * 2.f * xz; gy[14] = kC3[5] * -2.f * yz; gz[14] = kC3[5] * (xx - yy);
    gx[15] = kC3[6] * (3.f * xx - 3.f * yy); gy[15] = kC3[6] * -6.f * xy;
}

__device__ __forceinline__ float3 load3(const float* p, int i) {
    return make_float3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2));
}

// View-space depth (A.1).  Explicit fused operations: the depth-key kernel and the preprocess
// kernel -- two kernels on two streams -- must derive bit-identical depths from the same point.
__device__ __forceinline__ float view_depth(const float3 p, const float v2, const float v6, const float v10,
                                            const float v14) {
    return __fmaf_rn(p.x, v2, __fmaf_rn(p.y, v6, __fmaf_rn(p.z, v10, v14)));
}

// ------------------------------------------------------------------------------------------
// depth keys (the sort input) + the digit histograms of the four radix passes.  Reads 12 B and
// writes 8 B per Gaussian; runs on the auxiliary stream together with the depth sort while the
// main stream does the heavy preprocess.  Only the near-plane cull (A.1) is known here: Gaussians
// dropped later (degenerate covariance, empty tile rectangle) keep their depth key and simply own
// zero instances.
// ------------------------------------------------------------------------------------------
constexpr int KEY_THREADS = 256;
constexpr int KEY_ITEMS = 8;
__global__ void __launch_bounds__(KEY_THREADS)
depth_key_kernel(const float* __restrict__ means3D, const int P, const float* __restrict__ V,
                 uint32_t* __restrict__ depth_key, uint32_t* __restrict__ order_init,
                 uint32_t* __restrict__ sweep, const size_t pass_words) {
    pdl_trigger();                  // lets the next kernel of the chain become resident early (common.cuh); it waits for this grid to finish
    __shared__ uint32_t s_hist[4][RADIX_BINS];
#pragma unroll
    for (int p = 0; p < 4; p++) s_hist[p][threadIdx.x] = 0u;
    __syncthreads();
    const float v2 = __ldg(V + 2), v6 = __ldg(V + 6), v10 = __ldg(V + 10), v14 = __ldg(V + 14);
    const int base = blockIdx.x * KEY_THREADS * KEY_ITEMS;
    uint32_t kk[KEY_ITEMS];
#pragma unroll
    for (int it = 0; it < KEY_ITEMS; it++) {
        const int i = base + it * KEY_THREADS + threadIdx.x;
        kk[it] = CULLED_KEY;
        if (i < P) {
            const float zv = view_depth(load3(means3D, i), v2, v6, v10, v14);
            if (zv > NEAR_Z) kk[it] = __float_as_uint(zv);   // zv > 0.2 => bit order == numeric order (A.6)
        }
    }
#pragma unroll
    for (int it = 0; it < KEY_ITEMS; it++) {
        const int i = base + it * KEY_THREADS + threadIdx.x;
        if (i < P) {
            const uint32_t k = kk[it];
            depth_key[i] = k;
            order_init[i] = (uint32_t)i;      // value array of the depth sort
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

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// SH_FAST: M == 16 and shs 16-byte aligned -> CTA-cooperative 128-bit staging through smem.
// SPLIT (implies SH_FAST): the SH rows come from the model's features_dc / features_rest arrays of both sets.
// TMA (SH_FAST, not SPLIT): the CTA's 128 x 192-byte SH block arrives by 16 bulk-async copies of 8 rows each (one lane per
// copy: a bulk copy takes its addresses from uniform registers, so per-lane copies are issued one lane at a time -- 32
// rounds of 8 instructions per warp when every thread fetched its own row), completion on one mbarrier per CTA.  Nothing
// waits for the rows until the geometry (projection, covariance, radius, tile rectangle) of the Gaussian is done, where
// the register-staged variant spends 44 % of its warp time stalled in front of the barrier that closes the staging
// (profiles/r02_preprocess_sort.md).  Chunks are 8 x 192 + 16 bytes apart, and the threads of a warp take the rows in the order
// 8 (lane & 3) + 2 (lane >> 3) + ((lane >> 2) & 1): the 8 rows a quarter-warp reads with one LDS.128 then start in 8
// different groups of 4 banks -- conflict-free without padding every row.
constexpr int TMA_CHUNK_ROWS = 8;
constexpr int TMA_CHUNK_F4 = TMA_CHUNK_ROWS * SH_ROW_F4 + 1;      // 97 float4 = 1552 bytes
static_assert(PRE_THREADS / TMA_CHUNK_ROWS * TMA_CHUNK_F4 <= PRE_THREADS * SH_ROW_F4_PAD, "staging buffer");
template <bool SH_FAST, bool SPLIT, bool TMA>
__global__ void __launch_bounds__(PRE_THREADS, 8)
preprocess_forward_kernel(const ScgrView v, const ScgrGaussians g, Record* __restrict__ rec,
                          uint32_t* __restrict__ tiles_touched,
                          uint2* __restrict__ rect, unsigned long long* __restrict__ tile_mask,
                          int32_t* __restrict__ radii) {
    __shared__ __align__(16) float4 s_sh[SH_FAST ? PRE_THREADS * SH_ROW_F4_PAD : 1];
    __shared__ __align__(8) unsigned long long s_bar;
    const int P = g.P;
    const bool use_sh = SPLIT || g.shs != nullptr;
    const bool tma = TMA && use_sh;
    const int lane = threadIdx.x & 31;
    // local row of the CTA's block this thread processes (TMA: permuted inside the warp, see above)
    const int lrow = tma ? (int)(threadIdx.x & ~31u) + 8 * (lane & 3) + 2 * (lane >> 3) + ((lane >> 2) & 1) : (int)threadIdx.x;
    const int i = blockIdx.x * PRE_THREADS + lrow;
    if (tma) {
        const int nrows = min(PRE_THREADS, P - (int)blockIdx.x * PRE_THREADS);
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&s_bar, (uint32_t)nrows * (uint32_t)(SH_ROW_F4 * sizeof(float4)));      // the one arrival of the phase
        }
        __syncthreads();
        if (lane < 32 / TMA_CHUNK_ROWS) {
            const int chunk = (int)(threadIdx.x >> 5) * (32 / TMA_CHUNK_ROWS) + lane;
            const int first = chunk * TMA_CHUNK_ROWS;
            const int rows = min(TMA_CHUNK_ROWS, nrows - first);
            if (rows > 0)
                bulk_g2s(&s_sh[chunk * TMA_CHUNK_F4], g.shs + ((size_t)blockIdx.x * PRE_THREADS + first) * (SH_ROW_F4 * 4),
                         (uint32_t)rows * (uint32_t)(SH_ROW_F4 * sizeof(float4)), &s_bar);
        }
    }

    // the thread's own small inputs first: their loads are in flight together with the CTA's SH block below (a load cannot
    // be moved across the barrier that closes the staging, so left after it they would start a second latency period)
    float3 p = make_float3(0.f, 0.f, 0.f), sc_in = make_float3(0.f, 0.f, 0.f);
    float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
    float opac = 0.f;
    const bool has_sr = g.cov3D_precomp == nullptr;
    if (i < P) {
        p = load3(g.means3D, i);
        opac = __ldg(g.opacities + i);
        if (has_sr) {
            q_in = __ldg(reinterpret_cast<const float4*>(g.rotations) + i);
            sc_in = load3(g.scales, i);
        }
    }

    const float* my_dc = nullptr;      // SPLIT: this thread's rows inside the staging buffer
    const float* my_rest = nullptr;
    if (SPLIT) {
        // The rows of a set are contiguous in features_dc (3 floats each) and features_rest (45 floats each): every run is
        // copied as ONE flat stream -- 128-bit loads between a scalar head and tail, the shared-memory copy placed at
        // the same offset modulo 16 bytes as its source -- and kept in that layout: a thread then reads its own row with
        // a stride of 45 / 3 words, which is odd and therefore free of bank conflicts.
        float* const sf = reinterpret_cast<float*>(s_sh);
        float* const rest_s = sf;                                    // [128 * 45 + 16]
        float* const dc_s = sf + PRE_THREADS * SH_REST + 16;         // [128 * 3 + 16]
        const int row0 = blockIdx.x * PRE_THREADS;
        int cur_rest = 0, cur_dc = 0;
        auto stage = [&](const float* __restrict__ src, const int n, float* dst, int& cursor) {
            const int ms = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);
            const int off = ((cursor + 3) & ~3) + ms;
            const int head = min(n, (4 - ms) & 3);
            const int body4 = (n - head) >> 2;
            if ((int)threadIdx.x < head) dst[off + threadIdx.x] = __ldg(src + threadIdx.x);
            const float4* s4 = reinterpret_cast<const float4*>(src + head);
            float4* d4 = reinterpret_cast<float4*>(dst + off + head);
            for (int q = threadIdx.x; q < body4; q += PRE_THREADS) d4[q] = __ldg(s4 + q);
            const int done = head + 4 * body4;
            if ((int)threadIdx.x < n - done) dst[off + done + threadIdx.x] = __ldg(src + done + threadIdx.x);
            cursor = off + n;
            return off;
        };
        for_each_split_run(g, row0, min(PRE_THREADS, P - row0), [&](const int k, const int j0, const int cnt, const int r0) {
            const int o_dc = stage(g.sh_dc[k] + (size_t)j0 * SH_DC, cnt * SH_DC, dc_s, cur_dc);
            const int o_rest = stage(g.sh_rest[k] + (size_t)j0 * SH_REST, cnt * SH_REST, rest_s, cur_rest);
            const int r = (int)threadIdx.x - r0;
            if (r >= 0 && r < cnt) { my_dc = dc_s + o_dc + r * SH_DC; my_rest = rest_s + o_rest + r * SH_REST; }
        });
        __syncthreads();
    } else if (SH_FAST && use_sh && !TMA) {
        // rows [block0, block0 + 128) are one contiguous run of 128*12 float4 in HBM.  (Fetching only the
        // rows of Gaussians in front of the near plane was tried: the test needs means3D first, and the
        // serialised load latencies cost 25 % on an all-visible scene.)
        const int row0 = blockIdx.x * PRE_THREADS;
        const int nrows = min(PRE_THREADS, P - row0);
        const float4* src = reinterpret_cast<const float4*>(g.shs) + (size_t)row0 * SH_ROW_F4;
        const int nf4 = nrows * SH_ROW_F4;
        for (int f = threadIdx.x; f < nf4; f += PRE_THREADS) {
            const int r = f / SH_ROW_F4, c = f - r * SH_ROW_F4;
            s_sh[r * SH_ROW_F4_PAD + c] = __ldg(src + f);
        }
        __syncthreads();
    }
    if (i >= P) return;

    Camera cam;
    load_camera(v, cam);
    const float zv = view_depth(p, cam.V[2], cam.V[6], cam.V[10], cam.V[14]);

    bool alive = zv > NEAR_Z;   // A.1
    if (!alive && v.prefiltered) {
        printf("scgr: point %d is filtered although prefiltered is set\n", i);
        __trap();
    }
    Proj pr;
    float det = 0.f, px = 0.f, py = 0.f;
    int rad = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    const int gx = (v.image_width + TILE - 1) / TILE, gy = (v.image_height + TILE - 1) / TILE;
    if (alive) {
        const float* PM = cam.PM;
        const float h0 = p.x * PM[0] + p.y * PM[4] + p.z * PM[8] + PM[12];
        const float h1 = p.x * PM[1] + p.y * PM[5] + p.z * PM[9] + PM[13];
        const float h3 = p.x * PM[3] + p.y * PM[7] + p.z * PM[11] + PM[15];
        const float pw = 1.f / (h3 + 1e-7f);                          // A.2
        float c6[6];
        if (!has_sr) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = __ldg(g.cov3D_precomp + 6 * (size_t)i + k);
        } else {
            cov3d_from_scale_rot(sc_in, v.scale_modifier, q_in, c6);   // A.3
        }
        project_cov(cam, v, p, c6, pr);                               // A.4
        det = pr.a * pr.c - pr.b * pr.b;
        alive = det != 0.f;
        if (alive) {
            const float mid = 0.5f * (pr.a + pr.c);
            const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
            const float lam = fmaxf(mid + disc, mid - disc);
            rad = (int)ceilf(3.f * sqrtf(lam));
            px = ((h0 * pw + 1.f) * v.image_width - 1.f) * 0.5f;
            py = ((h1 * pw + 1.f) * v.image_height - 1.f) * 0.5f;
            x0 = min(gx, max(0, (int)((px - rad) / TILE)));
            y0 = min(gy, max(0, (int)((py - rad) / TILE)));
            x1 = min(gx, max(0, (int)((px + rad + TILE - 1) / TILE)));
            y1 = min(gy, max(0, (int)((py + rad + TILE - 1) / TILE)));
            alive = (x1 - x0) * (y1 - y0) != 0;
        }
    }
    if (!alive) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        rect[i] = make_uint2(0u, 0u);
        tile_mask[i] = 0ull;
        rec[i].q2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));   // radius 0 marks "culled" for the backward
        if (tma) mbar_wait(&s_bar, 0u);      // the CTA's shared memory must outlive the copies in flight
        return;
    }
    if (tma) mbar_wait(&s_bar, 0u);

    // A.5 colour
    float3 rgb;
    uint32_t flags = 0;
    if (use_sh) {
        float3 d = make_float3(p.x - cam.cam[0], p.y - cam.cam[1], p.z - cam.cam[2]);
        const float inv = 1.f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        d.x *= inv; d.y *= inv; d.z *= inv;
        float b[16];
        sh_basis(v.sh_degree, d, b);
        const int nk = (v.sh_degree + 1) * (v.sh_degree + 1);
        float acc[3] = {0.f, 0.f, 0.f};
        if (SPLIT) {
            // (same order of accumulation per channel as the assembled layout below: bit-identical colours)
#pragma unroll
            for (int c = 0; c < 3; c++) acc[c] += b[0] * my_dc[c];
#pragma unroll
            for (int k = 1; k < 16; k++)
                if (k < nk) {
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[c] += b[k] * my_rest[3 * (k - 1) + c];
                }
        } else if (SH_FAST) {
            const float4* row = tma ? s_sh + (lrow / TMA_CHUNK_ROWS) * TMA_CHUNK_F4 + (lrow % TMA_CHUNK_ROWS) * SH_ROW_F4
                                    : s_sh + threadIdx.x * SH_ROW_F4_PAD;
#pragma unroll
            for (int cc = 0; cc < SH_ROW_F4; cc++) {
                const float4 q = row[cc];
                const float in[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int k = (4 * cc + e) / 3, c = (4 * cc + e) - 3 * k;   // compile-time
                    if (k < nk) acc[c] += b[k] * in[e];
                }
            }
        } else {
            const float* row = g.shs + (size_t)i * g.sh_coeffs * 3;
            for (int k = 0; k < nk; k++) {
                acc[0] += b[k] * __ldg(row + 3 * k); acc[1] += b[k] * __ldg(row + 3 * k + 1); acc[2] += b[k] * __ldg(row + 3 * k + 2);
            }
        }
        rgb = make_float3(acc[0] + 0.5f, acc[1] + 0.5f, acc[2] + 0.5f);
        if (rgb.x < 0.f) flags |= 1u;
        if (rgb.y < 0.f) flags |= 2u;
        if (rgb.z < 0.f) flags |= 4u;
        rgb.x = fmaxf(rgb.x, 0.f); rgb.y = fmaxf(rgb.y, 0.f); rgb.z = fmaxf(rgb.z, 0.f);
    } else {
        rgb = load3(g.colors_precomp, i);
    }

    const float dinv = 1.f / det;
    // render-ready conic: power2 = cA dx^2 + cB dx dy + cC dy^2 = log2(e) * (-0.5 (A dx^2 + C dy^2) - B dx dy)
    constexpr float LOG2E = 1.4426950408889634f;
    Record r;
    r.q0 = make_float4(px, py, -0.5f * LOG2E * (pr.c * dinv), LOG2E * (pr.b * dinv));
    // q1.w: alpha >= 1/255  <=>  power2 >= -log2(255 * opacity)
    r.q1 = make_float4(-0.5f * LOG2E * (pr.a * dinv), opac, zv, -log2f(255.f * opac));
    r.q2 = make_float4(rgb.x, rgb.y, rgb.z, __uint_as_float((uint32_t)rad | (flags << 28)));
    rec[i] = r;
    radii[i] = rad;
    // Output-preserving culling: of the tiles in the reference's rect (A.4) keep only those in which
    // some pixel can pass the reference's own alpha >= 1/255 test (exact closed-form bound).  The
    // dropped (Gaussian, tile) pairs would be evaluated and skipped pixel by pixel in A.8.
    // For rects of <= 64 tiles the survivors are also recorded as a bit mask, so that the emission
    // kernel expands bits instead of repeating the test.
    const CullParams cp = make_cull(r.q0, r.q1);
    const SpanParams sp = make_span(cp);
    uint32_t touched = 0u;
    unsigned long long mask = 0ull;
    const int rw = x1 - x0;
    const bool small_rect = rw * (y1 - y0) <= 64;
    if (sp.robust) {
        // one closed-form column span per tile row (common.cuh)
        for (int ty = y0; ty < y1; ty++) {
            int first;
            const int n = row_span(sp, ty, x0, x1, &first);
            touched += (uint32_t)n;
            if (small_rect && n > 0)
                mask |= (n >= 64 ? ~0ull : ((1ull << n) - 1ull)) << ((ty - y0) * rw + (first - x0));
        }
    } else {
        for (int ty = y0; ty < y1; ty++)
            for (int tx = x0; tx < x1; tx++)
                if (tile_may_contribute(cp, tx, ty)) {
                    touched++;
                    if (small_rect) mask |= 1ull << ((ty - y0) * rw + (tx - x0));
                }
    }
    tiles_touched[i] = touched;
    tile_mask[i] = mask;
    rect[i] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
}

// ------------------------------------------------------------------------------------------
// backward (A.10), fused: conic->cov2D->{Sigma, t}, NDC mean, depth, SH, Sigma->{scale, rot}
// ------------------------------------------------------------------------------------------
// ACC: gradient accumulation over the views of a batch (ScgrGrads.accumulate): every parameter gradient is added to
// what the array holds, Gaussians without gradient are not touched at all; dL/dmean2D stays per view.
// TMA (SH_FAST, not SPLIT): the SH row of every live Gaussian of a round arrives by one bulk-async copy issued by the thread
// that will process it; the wait sits in front of step (4), behind the conic / covariance / mean chain of steps (1)-(3).
// (TMA 2: the same with twelve 16-byte cp.async per thread instead of one bulk copy -- a thread waits for its own copies
// and reads only its own row, so neither an mbarrier nor a block barrier is involved, and the copies are not issued lane
// by lane as per-lane bulk copies are.)
template <bool SH_FAST, int MINB, bool ACC, bool SPLIT, int TMA>
__global__ void __launch_bounds__(PB_T, MINB)
preprocess_backward_kernel(const ScgrView v, const ScgrGaussians g, const Record* __restrict__ rec,
                           const ScreenGrad* __restrict__ sg, const ScgrGrads out) {
    pdl_wait(); pdl_trigger();      // programmatic dependent launch (common.cuh): nothing is read or written before this
    __shared__ float4 s_sh[SH_FAST ? PB_T * SH_ROW_F4_PAD : 1];   // SH rows (then gradient rows) of one round, compact
    __shared__ float4 s_acc[PB_G][3];             // screen-space gradient sums of the live Gaussians
    __shared__ uint8_t s_live[PB_G];
    __shared__ uint8_t s_list[PB_G];              // local indices of the live Gaussians, compacted
    __shared__ int s_wcnt[PB_K][PB_T / 32];
    __shared__ uint32_t s_wbal[PB_K][PB_T / 32];   // live mask of local rows [32 (k PB_T / 32 + w), + 32)
    __shared__ __align__(8) unsigned long long s_bar;
    if (TMA == 1 && threadIdx.x == 0) {
        mbar_init(&s_bar, PB_T);      // every thread arrives once per round, the live ones with the bytes of their row
        mbar_fence_init();
    }
    const int P = g.P;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool use_sh = SPLIT || g.shs != nullptr;
    const int row0 = blockIdx.x * PB_G;
    const int nrows = min(PB_G, P - row0);

    // ---- phase 1: which Gaussians of this block of PB_G receive any gradient?  A Gaussian contributes
    // only if it survived the forward's culling AND render-backward deposited something for it: culled
    // ones, and the many that sit behind saturated pixels, have an all-zero accumulator -- every output
    // is then exactly zero and none of their inputs is read.  Each of the PB_T threads classifies PB_K
    // Gaussians; the live ones are compacted and processed PB_T at a time, so that the arithmetic runs
    // in full warps and a CTA's registers and staging buffer are sized for the live fraction.
    uint32_t bal[PB_K];
    bool own_live[PB_K];
#pragma unroll
    for (int k = 0; k < PB_K; k++) {
        const int local = k * PB_T + tid, own = row0 + local;
        own_live[k] = false;
        if (own < P) {
            // (a culled Gaussian owns no instances, so nothing was ever added to its accumulator: the
            // all-zero test covers it, and the record is not read here)
            const ScreenGrad A0 = sg[own];
            own_live[k] = A0.a0.x != 0.f || A0.a0.y != 0.f || A0.a0.z != 0.f || A0.a0.w != 0.f || A0.a1.x != 0.f ||
                          A0.a1.y != 0.f || A0.a1.z != 0.f || A0.a2.x != 0.f || A0.a2.y != 0.f || A0.a2.z != 0.f;
            if (own_live[k]) { s_acc[local][0] = A0.a0; s_acc[local][1] = A0.a1; s_acc[local][2] = A0.a2; }
        }
        s_live[local] = own_live[k] ? 1 : 0;
        bal[k] = __ballot_sync(0xffffffffu, own_live[k]);
        if (lane == 0) { s_wcnt[k][wid] = __popc(bal[k]); s_wbal[k][wid] = bal[k]; }
    }
    __syncthreads();
    int n_live = 0;
#pragma unroll
    for (int k = 0; k < PB_K; k++) {
        int before = n_live;
#pragma unroll
        for (int w = 0; w < PB_T / 32; w++) {
            if (w < wid) before += s_wcnt[k][w];
            n_live += s_wcnt[k][w];
        }
        if (own_live[k]) s_list[before + __popc(bal[k] & ((1u << lane) - 1u))] = (uint8_t)(k * PB_T + tid);
    }
    // Gaussians without gradient: zeros.  The small tensors by the classifying thread, the 192-byte
    // dL/dSH rows by the whole CTA (coalesced).
#pragma unroll
    for (int k = 0; k < PB_K; k++) {
        const int own = row0 + k * PB_T + tid;
        if (own < P && !own_live[k]) {
            out.dL_dmeans2D[3 * (size_t)own] = 0.f; out.dL_dmeans2D[3 * (size_t)own + 1] = 0.f; out.dL_dmeans2D[3 * (size_t)own + 2] = 0.f;
            if (out.densification_stats) {      // {|dL/dmean2D| visible, visible}: visible without gradient counts in the denominator
                const float vis = __ldg(out.radii + own) > 0 ? 1.f : 0.f;
                float2* st = reinterpret_cast<float2*>(out.densification_stats) + own;
                if (ACC) { if (vis != 0.f) { float2 o = *st; o.y += vis; *st = o; } }
                else *st = make_float2(0.f, vis);
            }
            if (out.live_count && !ACC) out.live_count[own] = 0.f;
            if (ACC) continue;
            out.dL_dmeans3D[3 * (size_t)own] = 0.f; out.dL_dmeans3D[3 * (size_t)own + 1] = 0.f; out.dL_dmeans3D[3 * (size_t)own + 2] = 0.f;
            out.dL_dopacities[own] = 0.f;
            if (out.dL_dcolors_precomp) {
                out.dL_dcolors_precomp[3 * (size_t)own] = 0.f; out.dL_dcolors_precomp[3 * (size_t)own + 1] = 0.f; out.dL_dcolors_precomp[3 * (size_t)own + 2] = 0.f;
            }
            if (out.dL_dscales) {
                out.dL_dscales[3 * (size_t)own] = 0.f; out.dL_dscales[3 * (size_t)own + 1] = 0.f; out.dL_dscales[3 * (size_t)own + 2] = 0.f;
            }
            if (out.dL_drotations) reinterpret_cast<float4*>(out.dL_drotations)[own] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (out.dL_dcov3D_precomp) {
#pragma unroll
                for (int c = 0; c < 6; c++) out.dL_dcov3D_precomp[6 * (size_t)own + c] = 0.f;
            }
            if (use_sh && !SH_FAST) {
                float* orow = out.dL_dshs + (size_t)own * g.sh_coeffs * 3;
                for (int c = 0; c < 3 * g.sh_coeffs; c++) orow[c] = 0.f;
            }
        }
    }
    // SPLIT: a row of the model's arrays is 3 + 45 floats at 4-byte alignment.  16 lanes share a row (lane l moves floats l,
    // l + 16, l + 32 of features_rest, lanes 0..2 also one float of features_dc), so the row's set and address are
    // worked out once per lane and row, not once per float.
    constexpr int RL = 16, RG = PB_T / RL;        // lanes per row, rows in flight per CTA
    const int rl = tid & (RL - 1), rgrp = tid / RL;
    // Zero rows of dL/dSH for the Gaussians without gradient: every group of RL lanes walks the dead rows of its residue class
    // (row % RG) by bit scanning the live masks -- no memory access decides a branch.
    if (SH_FAST && use_sh && !ACC) {
#pragma unroll
        for (int wd = 0; wd < PB_G / 32; wd++) {
            const int base = 32 * wd;
            if (base >= nrows) break;
            const uint32_t valid = nrows - base >= 32 ? 0xffffffffu : ((1u << (nrows - base)) - 1u);
            const uint32_t livew = (&s_wbal[0][0])[wd];
            constexpr uint32_t CLASS = RG == 4 ? 0x11111111u : (RG == 2 ? 0x55555555u : (RG == 8 ? 0x01010101u : 0xffffffffu));
            static_assert(RG == 1 || RG == 2 || RG == 4 || RG == 8, "rows in flight per CTA");
            uint32_t bits = ~livew & valid & (CLASS << rgrp);
            while (bits) {
                const int r = base + __ffs(bits) - 1;
                bits &= bits - 1u;
                const int gi = row0 + r;
                if (SPLIT) {
                    const int k = gi >= g.sh_n0;
                    const size_t j = (size_t)(gi - (k ? g.sh_n0 : 0));
                    float* rest = (k ? out.dL_dsh_rest[1] : out.dL_dsh_rest[0]) + j * SH_REST;
                    rest[rl] = 0.f; rest[rl + RL] = 0.f;
                    if (rl + 2 * RL < SH_REST) rest[rl + 2 * RL] = 0.f;
                    if (rl < SH_DC) ((k ? out.dL_dsh_dc[1] : out.dL_dsh_dc[0]) + j * SH_DC)[rl] = 0.f;
                } else if (rl < SH_ROW_F4) {
                    reinterpret_cast<float4*>(out.dL_dshs)[(size_t)gi * SH_ROW_F4 + rl] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2..4, PB_T live Gaussians per round: thread t takes the t-th of the round ----
    for (int first = 0; first < n_live; first += PB_T) {
    const int in_round = min(PB_T, n_live - first);
    const bool live = tid < in_round;
    const int jl = live ? (int)s_list[first + tid] : 0;      // local index of the Gaussian this thread processes
    const int i = row0 + jl;
    const bool tma = TMA != 0 && use_sh;
    if (tma && TMA == 2) {
        if (live) {
            const float4* src = reinterpret_cast<const float4*>(g.shs) + (size_t)i * SH_ROW_F4;
            float4* dst = &s_sh[tid * SH_ROW_F4_PAD];
#pragma unroll
            for (int c = 0; c < SH_ROW_F4; c++) cp_async16(dst + c, src + c);
        }
        cp_async_commit();
    } else if (tma) {
        if (live) {
            fence_proxy_async();      // the previous round read and rewrote this row through the generic proxy
            mbar_expect_tx(&s_bar, SH_ROW_F4 * sizeof(float4));
            bulk_g2s(&s_sh[tid * SH_ROW_F4_PAD], g.shs + (size_t)i * (SH_ROW_F4 * 4), SH_ROW_F4 * sizeof(float4), &s_bar);
        } else {
            mbar_arrive(&s_bar);
        }
    } else if (SPLIT) {
        // the 48 floats of every live row of the round, from whichever set the row belongs to
        float* const sf = reinterpret_cast<float*>(s_sh);
        // four rows per lane at a time: all their loads are issued before the first one is stored (a row's address comes
        // from s_list, a byte array the compiler must assume the staging stores alias -- row by row the loads serialise)
        constexpr int RU = 4;
        for (int rb = rgrp; rb < in_round; rb += RU * RG) {
            const float* rest[RU];
            const float* dc[RU];
            bool ok[RU];
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const int r = rb + u * RG;
                ok[u] = r < in_round;
                const int gi = row0 + (int)s_list[first + (ok[u] ? r : rb)];
                const int k = gi >= g.sh_n0;
                const size_t j = (size_t)(gi - (k ? g.sh_n0 : 0));
                // (ternaries, not g.sh_dc[k]: a run-time index into the kernel parameters would spill them to local memory)
                rest[u] = (k ? g.sh_rest[1] : g.sh_rest[0]) + j * SH_REST;
                dc[u] = (k ? g.sh_dc[1] : g.sh_dc[0]) + j * SH_DC;
            }
            float x[RU][4];
#pragma unroll
            for (int u = 0; u < RU; u++) {
                x[u][0] = ok[u] ? __ldg(rest[u] + rl) : 0.f;
                x[u][1] = ok[u] ? __ldg(rest[u] + rl + RL) : 0.f;
                x[u][2] = ok[u] && rl + 2 * RL < SH_REST ? __ldg(rest[u] + rl + 2 * RL) : 0.f;
                x[u][3] = ok[u] && rl < SH_DC ? __ldg(dc[u] + rl) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < RU; u++) {
                if (!ok[u]) continue;
                float* const row = sf + (rb + u * RG) * SH_ROW_PAD;
                row[SH_DC + rl] = x[u][0];
                row[SH_DC + rl + RL] = x[u][1];
                if (rl + 2 * RL < SH_REST) row[SH_DC + rl + 2 * RL] = x[u][2];
                if (rl < SH_DC) row[rl] = x[u][3];
            }
        }
    } else if (SH_FAST && use_sh) {
        const float4* src = reinterpret_cast<const float4*>(g.shs) + (size_t)row0 * SH_ROW_F4;
        // (four loads in flight per thread before the first store, for the reason given above)
        constexpr int FU = 4;
        const int nf = in_round * SH_ROW_F4;
        for (int fb = tid; fb < nf; fb += FU * PB_T) {
            float4 x[FU];
            int dst[FU];
#pragma unroll
            for (int u = 0; u < FU; u++) {
                const int f = fb + u * PB_T;
                const int fc = f < nf ? f : fb;
                const int r = fc / SH_ROW_F4, c = fc - r * SH_ROW_F4;
                dst[u] = f < nf ? r * SH_ROW_F4_PAD + c : -1;
                x[u] = __ldg(src + (int)s_list[first + r] * SH_ROW_F4 + c);
            }
#pragma unroll
            for (int u = 0; u < FU; u++)
                if (dst[u] >= 0) s_sh[dst[u]] = x[u];
        }
    }
    // (the per-Gaussian inputs below are fetched while the SH rows are in flight)
    float3 p_in = make_float3(0.f, 0.f, 0.f), sc_in = make_float3(0.f, 0.f, 0.f);
    float4 q_in = make_float4(1.f, 0.f, 0.f, 0.f), rq0 = make_float4(0.f, 0.f, 0.f, 0.f), rq1 = rq0;
    uint32_t rbits = 0u;                              // record word {radius | flags << 28}
    if (live) {
        p_in = load3(g.means3D, i);
        rq0 = rec[i].q0;
        rq1 = rec[i].q1;
        rbits = __float_as_uint(rec[i].q2.w);
        if (!g.cov3D_precomp) {
            q_in = __ldg(reinterpret_cast<const float4*>(g.rotations) + i);
            sc_in = load3(g.scales, i);
        }
    }
    if (SH_FAST && use_sh && !tma) __syncthreads();      // (SPLIT implies SH_FAST)

    float dmean[3] = {0.f, 0.f, 0.f};
    float dm2x = 0.f, dm2y = 0.f, dop = 0.f;
    float d6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float drot[4] = {0.f, 0.f, 0.f, 0.f};
    float dcol[3] = {0.f, 0.f, 0.f};
    if (live) {
        Camera cam;
        load_camera(v, cam);
        const uint32_t flags = rbits >> 28;
        ScreenGrad A;
        A.a0 = s_acc[jl][0]; A.a1 = s_acc[jl][1]; A.a2 = s_acc[jl][2];
        const float3 p = p_in;
        const float* V = cam.V;
        const float* PM = cam.PM;
        // render-backward stores raw sums (render.cu): scale them into true derivatives here
        constexpr float LN2 = 0.6931471805599453f;
        const float Sx = A.a0.x, Sy = A.a0.y;            // sum u G dx, sum u G dy
        dm2x = LN2 * (2.f * rq0.z * Sx + rq0.w * Sy) * (0.5f * v.image_width);    // NDC units (A.9)
        dm2y = LN2 * (2.f * rq1.x * Sy + rq0.w * Sx) * (0.5f * v.image_height);
        dop = A.a1.y / rq1.y;                            // sum (opacity G) dL/dalpha  ->  sum G dL/dalpha
        // (1) conic -> cov2D
        float c6[6];
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        float3 sc = make_float3(0.f, 0.f, 0.f);
        if (g.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = __ldg(g.cov3D_precomp + 6 * (size_t)i + k);
        } else {
            q = q_in;
            sc = sc_in;
            cov3d_from_scale_rot(sc, v.scale_modifier, q, c6);
        }
        Proj pr;
        project_cov(cam, v, p, c6, pr);
        const float a = pr.a, b = pr.b, c = pr.c;
        const float den = a * c - b * b;
        const float k2 = 1.f / (den * den + 1e-7f);
        const float dA = -0.5f * A.a0.z, dB = -A.a0.w, dC = -0.5f * A.a1.x;
        const float dLa = k2 * (-c * c * dA + b * c * dB - b * b * dC);
        const float dLb = k2 * (2.f * b * c * dA - (den + 2.f * b * b) * dB + 2.f * a * b * dC);
        const float dLc = k2 * (-b * b * dA + a * b * dB - a * a * dC);
        const float3 M0 = pr.M0, M1 = pr.M1;
        d6[0] = dLa * M0.x * M0.x + dLb * M0.x * M1.x + dLc * M1.x * M1.x;
        d6[3] = dLa * M0.y * M0.y + dLb * M0.y * M1.y + dLc * M1.y * M1.y;
        d6[5] = dLa * M0.z * M0.z + dLb * M0.z * M1.z + dLc * M1.z * M1.z;
        d6[1] = 2.f * dLa * M0.x * M0.y + dLb * (M0.x * M1.y + M0.y * M1.x) + 2.f * dLc * M1.x * M1.y;
        d6[2] = 2.f * dLa * M0.x * M0.z + dLb * (M0.x * M1.z + M0.z * M1.x) + 2.f * dLc * M1.x * M1.z;
        d6[4] = 2.f * dLa * M0.y * M0.z + dLb * (M0.y * M1.z + M0.z * M1.y) + 2.f * dLc * M1.y * M1.z;
        // cov2D -> M -> J -> t -> mean
        const float s0x = c6[0] * M0.x + c6[1] * M0.y + c6[2] * M0.z;
        const float s0y = c6[1] * M0.x + c6[3] * M0.y + c6[4] * M0.z;
        const float s0z = c6[2] * M0.x + c6[4] * M0.y + c6[5] * M0.z;
        const float s1x = c6[0] * M1.x + c6[1] * M1.y + c6[2] * M1.z;
        const float s1y = c6[1] * M1.x + c6[3] * M1.y + c6[4] * M1.z;
        const float s1z = c6[2] * M1.x + c6[4] * M1.y + c6[5] * M1.z;
        const float dM0x = 2.f * dLa * s0x + dLb * s1x, dM0y = 2.f * dLa * s0y + dLb * s1y, dM0z = 2.f * dLa * s0z + dLb * s1z;
        const float dM1x = dLb * s0x + 2.f * dLc * s1x, dM1y = dLb * s0y + 2.f * dLc * s1y, dM1z = dLb * s0z + 2.f * dLc * s1z;
        const float dJ00 = dM0x * V[0] + dM0y * V[4] + dM0z * V[8];
        const float dJ02 = dM0x * V[2] + dM0y * V[6] + dM0z * V[10];
        const float dJ11 = dM1x * V[1] + dM1y * V[5] + dM1z * V[9];
        const float dJ12 = dM1x * V[2] + dM1y * V[6] + dM1z * V[10];
        const float tzi = 1.f / pr.t.z, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = pr.xin ? -pr.fx * tz2 * dJ02 : 0.f;
        const float dty = pr.yin ? -pr.fy * tz2 * dJ12 : 0.f;
        const float dtz = -pr.fx * tz2 * dJ00 - pr.fy * tz2 * dJ11 + 2.f * pr.fx * pr.tx * tz3 * dJ02 +
                          2.f * pr.fy * pr.ty * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] = V[k * 4] * dtx + V[k * 4 + 1] * dty + V[k * 4 + 2] * dtz;
        // (2) NDC mean -> mean3D
        const float h0 = p.x * PM[0] + p.y * PM[4] + p.z * PM[8] + PM[12];
        const float h1 = p.x * PM[1] + p.y * PM[5] + p.z * PM[9] + PM[13];
        const float h3 = p.x * PM[3] + p.y * PM[7] + p.z * PM[11] + PM[15];
        const float pw = 1.f / (h3 + 1e-7f);
        const float mx = h0 * pw * pw, my = h1 * pw * pw;
#pragma unroll
        for (int k = 0; k < 3; k++)
            dmean[k] += dm2x * (PM[k * 4] * pw - mx * PM[k * 4 + 3]) + dm2y * (PM[k * 4 + 1] * pw - my * PM[k * 4 + 3]);
        // (3) depth -> mean3D
        const float dd = A.a1.z;
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] += V[k * 4 + 2] * dd;
        // (4) colour
        const float gr = (flags & 1u) ? 0.f : A.a2.x;
        const float gg = (flags & 2u) ? 0.f : A.a2.y;
        const float gb = (flags & 4u) ? 0.f : A.a2.z;
        if (use_sh) {
            float3 d = make_float3(p.x - cam.cam[0], p.y - cam.cam[1], p.z - cam.cam[2]);
            const float inv = 1.f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
            d.x *= inv; d.y *= inv; d.z *= inv;
            const int D = v.sh_degree;
            const int nk = (D + 1) * (D + 1);
            float bb[16], bx[16], by[16], bz[16];
            sh_basis(D, d, bb);
            sh_basis_grad(D, d, bx, by, bz);
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            if (tma && TMA == 2) cp_async_wait_all();
            else if (tma) mbar_wait(&s_bar, (uint32_t)(first / PB_T) & 1u);
            if (SH_FAST) {
                // in place: the staged input row of the Gaussian becomes its gradient row
                float4* row = s_sh + tid * SH_ROW_F4_PAD;
                const float gch[3] = {gr, gg, gb};
#pragma unroll
                for (int cc = 0; cc < SH_ROW_F4; cc++) {
                    const float4 qq = row[cc];
                    const float in[4] = {qq.x, qq.y, qq.z, qq.w};
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int k = (4 * cc + e) / 3, c = (4 * cc + e) - 3 * k;   // compile-time
                        if (k < nk) {
                            o[e] = bb[k] * gch[c];
                            const float w = gch[c] * in[e];
                            ddx += bx[k] * w; ddy += by[k] * w; ddz += bz[k] * w;
                        } else {
                            o[e] = 0.f;
                        }
                    }
                    row[cc] = make_float4(o[0], o[1], o[2], o[3]);
                }
            } else {
                const float* row = g.shs + (size_t)i * g.sh_coeffs * 3;
                float* orow = out.dL_dshs + (size_t)i * g.sh_coeffs * 3;
                for (int k = 0; k < g.sh_coeffs; k++) {
                    if (k < nk) {
                        if (ACC) { orow[3 * k] += bb[k] * gr; orow[3 * k + 1] += bb[k] * gg; orow[3 * k + 2] += bb[k] * gb; }
                        else { orow[3 * k] = bb[k] * gr; orow[3 * k + 1] = bb[k] * gg; orow[3 * k + 2] = bb[k] * gb; }
                        const float w = gr * __ldg(row + 3 * k) + gg * __ldg(row + 3 * k + 1) + gb * __ldg(row + 3 * k + 2);
                        ddx += bx[k] * w; ddy += by[k] * w; ddz += bz[k] * w;
                    } else if (!ACC) {
                        orow[3 * k] = 0.f; orow[3 * k + 1] = 0.f; orow[3 * k + 2] = 0.f;
                    }
                }
            }
            const float dot = d.x * ddx + d.y * ddy + d.z * ddz;   // through dir = v / |v|
            dmean[0] += (ddx - d.x * dot) * inv;
            dmean[1] += (ddy - d.y * dot) * inv;
            dmean[2] += (ddz - d.z * dot) * inv;
        } else {
            dcol[0] = A.a2.x; dcol[1] = A.a2.y; dcol[2] = A.a2.z;
        }
        // (5) Sigma -> scale, rotation
        if (!g.cov3D_precomp) {
            float R[9];
            quat_to_R(q, R);
            const float mod = v.scale_modifier;
            const float sp[3] = {mod * sc.x, mod * sc.y, mod * sc.z};
            const float Sg[9] = {2.f * d6[0], d6[1], d6[2], d6[1], 2.f * d6[3], d6[4], d6[2], d6[4], 2.f * d6[5]};
            float Lm[9], dLm[9], Dm[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int j = 0; j < 3; j++) Lm[r * 3 + j] = R[r * 3 + j] * sp[j];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int j = 0; j < 3; j++)
                    dLm[r * 3 + j] = Sg[r * 3] * Lm[j] + Sg[r * 3 + 1] * Lm[3 + j] + Sg[r * 3 + 2] * Lm[6 + j];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                dscale[j] = mod * (dLm[j] * R[j] + dLm[3 + j] * R[3 + j] + dLm[6 + j] * R[6 + j]);
#pragma unroll
                for (int r = 0; r < 3; r++) Dm[r * 3 + j] = dLm[r * 3 + j] * sp[j];
            }
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot[0] = 2.f * (-z * Dm[1] + y * Dm[2] + z * Dm[3] - x * Dm[5] - y * Dm[6] + x * Dm[7]);
            drot[1] = 2.f * (y * Dm[1] + z * Dm[2] + y * Dm[3] - 2.f * x * Dm[4] - r * Dm[5] + z * Dm[6] + r * Dm[7] - 2.f * x * Dm[8]);
            drot[2] = 2.f * (-2.f * y * Dm[0] + x * Dm[1] + r * Dm[2] + x * Dm[3] + z * Dm[5] - r * Dm[6] + z * Dm[7] - 2.f * y * Dm[8]);
            drot[3] = 2.f * (-2.f * z * Dm[0] - r * Dm[1] + x * Dm[2] + r * Dm[3] - 2.f * z * Dm[4] + y * Dm[5] + x * Dm[6] + y * Dm[7]);
        }
    }

    if (live) {
        auto put = [](float* ptr, const float val) { if (ACC) *ptr += val; else *ptr = val; };
        put(out.dL_dmeans3D + 3 * (size_t)i, dmean[0]); put(out.dL_dmeans3D + 3 * (size_t)i + 1, dmean[1]); put(out.dL_dmeans3D + 3 * (size_t)i + 2, dmean[2]);
        out.dL_dmeans2D[3 * (size_t)i] = dm2x; out.dL_dmeans2D[3 * (size_t)i + 1] = dm2y; out.dL_dmeans2D[3 * (size_t)i + 2] = 0.f;
        put(out.dL_dopacities + i, dop);
        if (out.densification_stats) {      // a Gaussian with gradient is visible
            float2* st = reinterpret_cast<float2*>(out.densification_stats) + i;
            const float nrm = sqrtf(dm2x * dm2x + dm2y * dm2y);
            if (ACC) { float2 o = *st; o.x += nrm; o.y += 1.f; *st = o; }
            else *st = make_float2(nrm, 1.f);
        }
        if (out.live_count) put(out.live_count + i, 1.f);
        if (out.dL_dcolors_precomp) {
            put(out.dL_dcolors_precomp + 3 * (size_t)i, dcol[0]); put(out.dL_dcolors_precomp + 3 * (size_t)i + 1, dcol[1]); put(out.dL_dcolors_precomp + 3 * (size_t)i + 2, dcol[2]);
        }
        if (out.dL_dscales) {
            put(out.dL_dscales + 3 * (size_t)i, dscale[0]); put(out.dL_dscales + 3 * (size_t)i + 1, dscale[1]); put(out.dL_dscales + 3 * (size_t)i + 2, dscale[2]);
        }
        if (out.dL_drotations) {
            float4* pr4 = reinterpret_cast<float4*>(out.dL_drotations) + i;
            float4 o = make_float4(drot[0], drot[1], drot[2], drot[3]);
            if (ACC) { const float4 h = *pr4; o.x += h.x; o.y += h.y; o.z += h.z; o.w += h.w; }
            *pr4 = o;
        }
        if (out.dL_dcov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) put(out.dL_dcov3D_precomp + 6 * (size_t)i + k, d6[k]);
        }
    }
    if (SPLIT) {
        // the gradient rows of the round sit in smem (written in place above): back to the model's own arrays
        __syncthreads();
        const float* const sf = reinterpret_cast<const float*>(s_sh);
        for (int r = rgrp; r < in_round; r += RG) {
            const int gi = row0 + (int)s_list[first + r];
            const int k = gi >= g.sh_n0;
            const size_t j = (size_t)(gi - (k ? g.sh_n0 : 0));
            float* rest = (k ? out.dL_dsh_rest[1] : out.dL_dsh_rest[0]) + j * SH_REST;
            const float* const row = sf + r * SH_ROW_PAD;
            auto put1 = [&](float* d1, const float x) { if (ACC) *d1 += x; else *d1 = x; };
            put1(rest + rl, row[SH_DC + rl]);
            put1(rest + rl + RL, row[SH_DC + rl + RL]);
            if (rl + 2 * RL < SH_REST) put1(rest + rl + 2 * RL, row[SH_DC + rl + 2 * RL]);
            if (rl < SH_DC) put1((k ? out.dL_dsh_dc[1] : out.dL_dsh_dc[0]) + j * SH_DC + rl, row[rl]);
        }
        __syncthreads();      // the staging buffer is reused by the next round
    } else if (SH_FAST && use_sh) {
        // the 192-byte gradient rows of the round sit in smem (written in place above): stream them out,
        // 12 consecutive 128-bit stores per row
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(out.dL_dshs) + (size_t)row0 * SH_ROW_F4;
        for (int f = tid; f < in_round * SH_ROW_F4; f += PB_T) {
            const int r = f / SH_ROW_F4, c = f - r * SH_ROW_F4;
            float4* d4 = dst + (int)s_list[first + r] * SH_ROW_F4 + c;
            float4 o = s_sh[r * SH_ROW_F4_PAD + c];
            if (ACC) { const float4 h = *d4; o.x += h.x; o.y += h.y; o.z += h.z; o.w += h.w; }
            *d4 = o;
        }
        __syncthreads();      // the staging buffer is reused by the next round
    }
    }   // rounds
}

__global__ void mark_visible_kernel(const float* __restrict__ means3D, int P, const float* __restrict__ V,
                                    uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 p = load3(means3D, i);
    const float zv = view_depth(p, __ldg(V + 2), __ldg(V + 6), __ldg(V + 10), __ldg(V + 14));
    present[i] = zv > NEAR_Z ? 1 : 0;
}

inline bool sh_split(const ScgrGaussians& g) { return g.shs == nullptr && (g.sh_dc[0] != nullptr || g.sh_dc[1] != nullptr); }

inline bool sh_fast_ok(const ScgrGaussians& g, const void* dsh) {
    return g.shs != nullptr && g.sh_coeffs == 16 && (reinterpret_cast<uintptr_t>(g.shs) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(dsh) & 15) == 0;
}

}  // namespace

void launch_preprocess_forward(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G,
                               int32_t* radii, const Launch& L) {
    if (g.P <= 0) return;
    const int blocks = (g.P + PRE_THREADS - 1) / PRE_THREADS;
    begin_kernel("preprocess_forward", L);
    static const int tma = getenv("SCGR_TMA_PRE") ? atoi(getenv("SCGR_TMA_PRE")) : 1;
    if (sh_split(g))
        preprocess_forward_kernel<true, true, false><<<blocks, PRE_THREADS, 0, L.stream>>>(v, g, G.rec, G.tiles_touched, G.rect, G.tile_mask, radii);
    else if (sh_fast_ok(g, nullptr) && tma)
        preprocess_forward_kernel<true, false, true><<<blocks, PRE_THREADS, 0, L.stream>>>(v, g, G.rec, G.tiles_touched, G.rect, G.tile_mask, radii);
    else if (sh_fast_ok(g, nullptr))
        preprocess_forward_kernel<true, false, false><<<blocks, PRE_THREADS, 0, L.stream>>>(v, g, G.rec, G.tiles_touched, G.rect, G.tile_mask, radii);
    else
        preprocess_forward_kernel<false, false, false><<<blocks, PRE_THREADS, 0, L.stream>>>(v, g, G.rec, G.tiles_touched, G.rect, G.tile_mask, radii);
    check_launch("preprocess_forward", L);
}

void launch_depth_keys(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G, const Launch& L) {
    if (g.P <= 0) return;
    const int per_cta = KEY_THREADS * KEY_ITEMS;
    begin_kernel("depth_keys", L);
    depth_key_kernel<<<(g.P + per_cta - 1) / per_cta, KEY_THREADS, 0, L.stream>>>(
        g.means3D, g.P, v.viewmatrix, G.depth_key, G.sort_vals[0], G.sweep, sweep_pass_words(g.P));
    check_launch("depth_keys", L);
}

void launch_preprocess_backward(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G,
                                const ScgrGrads& out, const Launch& L) {
    if (g.P <= 0) return;
    const int blocks = (g.P + PB_G - 1) / PB_G;
    begin_kernel("preprocess_backward", L);
    static const int minb = getenv("SCGR_PREB_MINB") ? atoi(getenv("SCGR_PREB_MINB")) : 1;
    const bool acc = out.accumulate != 0;
    static const int tma = getenv("SCGR_TMA_PREB") ? atoi(getenv("SCGR_TMA_PREB")) : 2;
    if (sh_split(g)) {
        if (acc) chain(preprocess_backward_kernel<true, 1, true, true, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else chain(preprocess_backward_kernel<true, 1, false, true, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
    } else if (sh_fast_ok(g, out.dL_dshs)) {
        if (acc && tma == 2) chain(preprocess_backward_kernel<true, 1, true, false, 2>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (acc && tma) chain(preprocess_backward_kernel<true, 1, true, false, 1>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (acc) chain(preprocess_backward_kernel<true, 1, true, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (minb == 12) chain(preprocess_backward_kernel<true, 12, false, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (minb == 10) chain(preprocess_backward_kernel<true, 10, false, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (tma == 2) chain(preprocess_backward_kernel<true, 1, false, false, 2>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else if (tma) chain(preprocess_backward_kernel<true, 1, false, false, 1>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
        else chain(preprocess_backward_kernel<true, 1, false, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
    } else if (acc) {
        chain(preprocess_backward_kernel<false, 1, true, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
    } else {
        chain(preprocess_backward_kernel<false, 1, false, false, 0>, dim3(blocks), dim3(PB_T), 0, L)(v, g, G.rec, G.screen_grad, out);
    }
    check_launch("preprocess_backward", L);
}

void launch_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present,
                         const Launch& L) {
    if (P <= 0) return;
    begin_kernel("mark_visible", L);
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, L.stream>>>(means3D, P, viewmatrix, present);
    check_launch("mark_visible", L);
}

}  // namespace scgr
