// model.cu -- the elementwise / data-movement passes either side of the rasterizer in a training step (sm_100a).
// SURVEY.md section 8(f) rows f2, f3, f4 and 8(a) row a17.  All of them are pure HBM streams with no reuse; shared
// memory appears once, as the layout converter of the SH copy.  Measured: profiles/r01_model_passes.md.
//
// (f2) assemble_forward / assemble_backward: what SCGaussian's `GaussianModel.get_xyz / get_scaling /
//      get_rotation / get_opacity / get_features` compute on every render() call
//      (reference scene/gaussian_model.py:105-152) -- xyz = rayo + rayd * zval for the ray-based set,
//      exp / normalize / sigmoid activations, the cat of the ray-based and the free ("bg_") set and the
//      cat of features_dc with features_rest -- in ONE launch each way instead of ~20 torch kernels
//      (each activation, each cat and every autograd node is its own launch and its own round trip through
//      HBM: ~2x the bytes).  Output layout = exactly what GaussianRasterizer takes: [P,3] [P,3] [P,4] [P,1]
//      [P,K,3] with P = n_ray + n_bg, ray-based set first.
//      algorithmic bytes / Gaussian (K = 16): forward 252 in (ray set; 240 free set) + 236 out;
//      backward 236 in + 44 raw re-read + 228 out.
// (f3) adam: `torch.optim.Adam(groups, lr=0.0, eps=1e-15).step()` as the reference runs it twice per
//      iteration (reference scene/gaussian_model.py:491-510, train.py:204-208), all parameter groups of
//      both optimizers in ONE launch; per element 16 B read (param, grad, exp_avg, exp_avg_sq) + 12 B
//      written = 28 B.
// (a17) densification_stats: reference train.py:192-193 -> scene/gaussian_model.py:932-934, one launch, no host sync.
// (f4) gather_rows / copy_segments: the prune compaction (reference scene/gaussian_model.py:777-820) and the
//      densification append (:822-862) over every per-Gaussian array of the model, one launch per index list / per
//      append.
// tests/test_kernel_emulation.py compiles THIS FILE for the host and runs every kernel thread for thread.
#include <cstdlib>

#include "common.cuh"

namespace scgr {

namespace {

constexpr int MODEL_THREADS = 256;
constexpr int SH_PER_THREAD = 8;                                        // SH floats per thread, stride MODEL_THREADS
constexpr int SH_PER_BLOCK = MODEL_THREADS * SH_PER_THREAD;

__device__ __forceinline__ float sigmoid_f(const float x) { return 1.0f / (1.0f + expf(-x)); }

// the per-Gaussian arrays of one of the two sets, picked field by field (static indices into the kernel
// parameters: they stay in constant memory)
struct RawSet {
    const float *xyz, *rayo, *rayd, *zval, *scaling, *rotation, *opacity;
};
__device__ __forceinline__ RawSet select_set(const ScgrModel& m, const bool first) {
    RawSet s;
    s.xyz = first ? m.set[0].xyz : m.set[1].xyz;
    s.rayo = first ? m.set[0].rayo : m.set[1].rayo;
    s.rayd = first ? m.set[0].rayd : m.set[1].rayd;
    s.zval = first ? m.set[0].zval : m.set[1].zval;
    s.scaling = first ? m.set[0].scaling : m.set[1].scaling;
    s.rotation = first ? m.set[0].rotation : m.set[1].rotation;
    s.opacity = first ? m.set[0].opacity : m.set[1].opacity;
    return s;
}
__device__ __forceinline__ ScgrModelSetGrads select_grads(const ScgrModelGrads& g, const bool first) {
    ScgrModelSetGrads d;
    d.dL_dxyz = first ? g.set[0].dL_dxyz : g.set[1].dL_dxyz;
    d.dL_dzval = first ? g.set[0].dL_dzval : g.set[1].dL_dzval;
    d.dL_dscaling = first ? g.set[0].dL_dscaling : g.set[1].dL_dscaling;
    d.dL_drotation = first ? g.set[0].dL_drotation : g.set[1].dL_drotation;
    d.dL_dopacity = first ? g.set[0].dL_dopacity : g.set[1].dL_dopacity;
    d.dL_dfeatures_dc = first ? g.set[0].dL_dfeatures_dc : g.set[1].dL_dfeatures_dc;
    d.dL_dfeatures_rest = first ? g.set[0].dL_dfeatures_rest : g.set[1].dL_dfeatures_rest;
    return d;
}

// element k (0 .. 3K-1) of the [K,3] SH block of Gaussian i of the assembled model
__device__ __forceinline__ float load_sh(const ScgrModel& m, const uint32_t i, const uint32_t k, const uint32_t n3k) {
    const bool ray = i < (uint32_t)m.set[0].n;
    const uint32_t ii = ray ? i : i - (uint32_t)m.set[0].n;
    const float* dc = ray ? m.set[0].features_dc : m.set[1].features_dc;
    const float* rest = ray ? m.set[0].features_rest : m.set[1].features_rest;
    return k < 3 ? __ldg(dc + (size_t)ii * 3 + k) : __ldg(rest + (size_t)ii * (n3k - 3) + (k - 3));
}

// ---- the SH copy staged through shared memory (K = 16 or 4 coefficients: rows of 48 / 12 floats, a multiple of 16 bytes).
// A CTA owns 64 Gaussians of ONE set: their features_rest rows are one contiguous run of the source and their SH rows
// one contiguous, 16-byte aligned run of the assembled array.  The interleaved side (dc | rest, rows of 3 / N3K-3
// floats, no 16-byte alignment) is walked one float per lane on consecutive addresses; the assembled side moves as
// float4; shared memory converts between the two (both sides of it conflict-free: consecutive lanes on consecutive words).
constexpr int STAGE_G = 64;

template <int N3K>
__device__ __forceinline__ void staged_sh_forward(const ScgrModel& m, float* __restrict__ shs, const uint32_t blocks0) {
    constexpr int REST = N3K - 3;
    constexpr int REST_ITERS = (STAGE_G * REST + MODEL_THREADS - 1) / MODEL_THREADS;
    constexpr int OUT_ITERS = (STAGE_G * N3K / 4 + MODEL_THREADS - 1) / MODEL_THREADS;
    __shared__ __align__(16) float s[STAGE_G * N3K];
    const bool first = blockIdx.x < blocks0;
    const uint32_t n = (uint32_t)(first ? m.set[0].n : m.set[1].n);
    const uint32_t ii0 = (first ? blockIdx.x : blockIdx.x - blocks0) * STAGE_G;
    const uint32_t cnt = min((uint32_t)STAGE_G, n - ii0);
    const float* dc = (first ? m.set[0].features_dc : m.set[1].features_dc) + (size_t)ii0 * 3;
    const float* rest = (first ? m.set[0].features_rest : m.set[1].features_rest) + (size_t)ii0 * REST;
    float* out = shs + ((size_t)(first ? 0 : m.set[0].n) + ii0) * N3K;
    float v[REST_ITERS];
#pragma unroll
    for (int u = 0; u < REST_ITERS; u++) {
        const uint32_t r = u * MODEL_THREADS + threadIdx.x;
        v[u] = r < cnt * REST ? __ldg(rest + r) : 0.f;
    }
    const float d = threadIdx.x < cnt * 3 ? __ldg(dc + threadIdx.x) : 0.f;
#pragma unroll
    for (int u = 0; u < REST_ITERS; u++) {
        const uint32_t r = u * MODEL_THREADS + threadIdx.x;
        if (r < cnt * REST) { const uint32_t gi = r / REST; s[gi * N3K + 3 + (r - gi * REST)] = v[u]; }
    }
    if (threadIdx.x < cnt * 3) { const uint32_t gi = threadIdx.x / 3; s[gi * N3K + (threadIdx.x - gi * 3)] = d; }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < OUT_ITERS; u++) {
        const uint32_t q = u * MODEL_THREADS + threadIdx.x;
        if (q < cnt * (N3K / 4)) reinterpret_cast<float4*>(out)[q] = reinterpret_cast<const float4*>(s)[q];
    }
}

template <int N3K>
__device__ __forceinline__ void staged_sh_backward(const ScgrModel& m, const float* __restrict__ dL_dshs,
                                                   const ScgrModelGrads& g, const uint32_t blocks0) {
    constexpr int REST = N3K - 3;
    constexpr int REST_ITERS = (STAGE_G * REST + MODEL_THREADS - 1) / MODEL_THREADS;
    constexpr int IN_ITERS = (STAGE_G * N3K / 4 + MODEL_THREADS - 1) / MODEL_THREADS;
    __shared__ __align__(16) float s[STAGE_G * N3K];
    const bool first = blockIdx.x < blocks0;
    const uint32_t n = (uint32_t)(first ? m.set[0].n : m.set[1].n);
    const uint32_t ii0 = (first ? blockIdx.x : blockIdx.x - blocks0) * STAGE_G;
    const uint32_t cnt = min((uint32_t)STAGE_G, n - ii0);
    float* dc = (first ? g.set[0].dL_dfeatures_dc : g.set[1].dL_dfeatures_dc) + (size_t)ii0 * 3;
    float* rest = (first ? g.set[0].dL_dfeatures_rest : g.set[1].dL_dfeatures_rest) + (size_t)ii0 * REST;
    const float* in = dL_dshs + ((size_t)(first ? 0 : m.set[0].n) + ii0) * N3K;
    float4 v[IN_ITERS];
#pragma unroll
    for (int u = 0; u < IN_ITERS; u++) {
        const uint32_t q = u * MODEL_THREADS + threadIdx.x;
        v[u] = q < cnt * (N3K / 4) ? __ldg(reinterpret_cast<const float4*>(in) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < IN_ITERS; u++) {
        const uint32_t q = u * MODEL_THREADS + threadIdx.x;
        if (q < cnt * (N3K / 4)) reinterpret_cast<float4*>(s)[q] = v[u];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < REST_ITERS; u++) {
        const uint32_t r = u * MODEL_THREADS + threadIdx.x;
        if (r < cnt * REST) { const uint32_t gi = r / REST; rest[r] = s[gi * N3K + 3 + (r - gi * REST)]; }
    }
    if (threadIdx.x < cnt * 3) { const uint32_t gi = threadIdx.x / 3; dc[threadIdx.x] = s[gi * N3K + (threadIdx.x - gi * 3)]; }
}

template <int STAGED>      // 0: flat copy for any K; 48 / 12: staged_sh_* above
__global__ void __launch_bounds__(MODEL_THREADS)
assemble_forward_kernel(const __grid_constant__ ScgrModel m, const __grid_constant__ ScgrActivated o,
                        const uint32_t sh_blocks, const uint32_t n3k, const uint32_t sh_total) {
    const uint32_t step_i = MODEL_THREADS / n3k, step_k = MODEL_THREADS - step_i * n3k;   // (i, k) of element e + 256
    if constexpr (STAGED != 0) {
        if (blockIdx.x < sh_blocks) {       // get_features (reference scene/gaussian_model.py:131-140)
            staged_sh_forward<STAGED>(m, o.shs, ((uint32_t)m.set[0].n + STAGE_G - 1) / STAGE_G);
            return;
        }
    }
    if (blockIdx.x < sh_blocks) {
        // ---- get_features (reference scene/gaussian_model.py:131-140): flat over the P*K*3 output floats; consecutive
        // lanes read consecutive source floats and write consecutive output floats (the [n,K-1,3] source rows are 4
        // bytes off 16-byte alignment for every other Gaussian, so wider accesses would need staging), 8 per thread in flight
        float r[SH_PER_THREAD];
        const uint32_t e = blockIdx.x * SH_PER_BLOCK + threadIdx.x;
        uint32_t i = e / n3k, k = e - i * n3k;
#pragma unroll
        for (int u = 0; u < SH_PER_THREAD; u++) {
            r[u] = (e + u * MODEL_THREADS < sh_total) ? load_sh(m, i, k, n3k) : 0.f;
            k += step_k; i += step_i;
            if (k >= n3k) { k -= n3k; i++; }
        }
#pragma unroll
        for (int u = 0; u < SH_PER_THREAD; u++)
            if (e + u * MODEL_THREADS < sh_total) o.shs[e + u * MODEL_THREADS] = r[u];
        return;
    }
    // ---- a CTA per 256 Gaussians.  Position and scale are [P,3] arrays: walked as flat streams of 768 floats
    // (consecutive lanes on consecutive floats, in and out) rather than 3 strided floats per thread
    const uint32_t P = (uint32_t)m.set[0].n + (uint32_t)m.set[1].n;
    const uint32_t n0 = (uint32_t)m.set[0].n;
    const uint32_t g0 = (blockIdx.x - sh_blocks) * MODEL_THREADS;
#pragma unroll
    for (int u = 0; u < 3; u++) {
        const uint32_t e = 3 * g0 + u * MODEL_THREADS + threadIdx.x;       // flat index into the [P,3] outputs
        if (e < 3 * P) {
            const uint32_t j = e / 3;
            const bool jr = j < n0;
            const size_t ee = jr ? e : e - 3 * n0, jj = jr ? j : j - n0;
            const RawSet sj = select_set(m, jr);
            float v;
            if (sj.rayo)    // get_xyz, reference :124: rayo + rayd * zval (product rounded, then the sum: two torch kernels)
                v = __fadd_rn(__ldg(sj.rayo + ee), __fmul_rn(__ldg(sj.rayd + ee), __ldg(sj.zval + jj)));
            else
                v = __ldg(sj.xyz + ee);
            o.means3D[e] = v;
            o.scales[e] = expf(__ldg(sj.scaling + ee));                     // get_scaling, reference :105-113
        }
    }
    const uint32_t i = g0 + threadIdx.x;
    if (i >= P) return;
    const bool ray = i < n0;
    const size_t ii = ray ? i : i - n0;
    const RawSet s = select_set(m, ray);

    // get_rotation, reference :115-122: torch.nn.functional.normalize = q / max(|q|, 1e-12)
    const float4 q = __ldg(reinterpret_cast<const float4*>(s.rotation) + ii);
    const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float d = fmaxf(nrm, 1e-12f);
    reinterpret_cast<float4*>(o.rotations)[i] = make_float4(q.x / d, q.y / d, q.z / d, q.w / d);

    // get_opacity, reference :142-150: sigmoid
    o.opacities[i] = sigmoid_f(__ldg(s.opacity + ii));
}

template <int STAGED>
__global__ void __launch_bounds__(MODEL_THREADS)
assemble_backward_kernel(const __grid_constant__ ScgrModel m, const __grid_constant__ ScgrActivatedGrads g,
                         const __grid_constant__ ScgrModelGrads out, const uint32_t sh_blocks, const uint32_t n3k,
                         const uint32_t sh_total) {
    const uint32_t n0 = (uint32_t)m.set[0].n;
    const uint32_t step_i = MODEL_THREADS / n3k, step_k = MODEL_THREADS - step_i * n3k;   // (i, k) of element e + 256
    if constexpr (STAGED != 0) {
        if (blockIdx.x < sh_blocks) {
            staged_sh_backward<STAGED>(m, g.dL_dshs, out, (n0 + STAGE_G - 1) / STAGE_G);
            return;
        }
    }
    if (blockIdx.x < sh_blocks) {
        // ---- dL/dshs [P,K,3] split into dL/dfeatures_dc [n,1,3] and dL/dfeatures_rest [n,K-1,3] of the two sets:
        // the mirror image of the forward copy, coalesced 4-byte accesses on both sides
        float v[SH_PER_THREAD];
        const uint32_t e = blockIdx.x * SH_PER_BLOCK + threadIdx.x;
#pragma unroll
        for (int u = 0; u < SH_PER_THREAD; u++)
            v[u] = (e + u * MODEL_THREADS < sh_total) ? __ldg(g.dL_dshs + e + u * MODEL_THREADS) : 0.f;
        uint32_t i = e / n3k, k = e - i * n3k;
#pragma unroll
        for (int u = 0; u < SH_PER_THREAD; u++) {
            if (e + u * MODEL_THREADS < sh_total) {
                const bool ray = i < n0;
                const size_t ii = ray ? i : i - n0;
                float* dc = ray ? out.set[0].dL_dfeatures_dc : out.set[1].dL_dfeatures_dc;
                float* rest = ray ? out.set[0].dL_dfeatures_rest : out.set[1].dL_dfeatures_rest;
                if (k < 3) dc[ii * 3 + k] = v[u];
                else rest[ii * (n3k - 3) + (k - 3)] = v[u];
            }
            k += step_k; i += step_i;
            if (k >= n3k) { k -= n3k; i++; }
        }
        return;
    }
    // ---- a CTA per 256 Gaussians; the [n,3] gradients (free positions, log scales) as flat streams of 768 floats
    const uint32_t P = n0 + (uint32_t)m.set[1].n;
    const uint32_t g0 = (blockIdx.x - sh_blocks) * MODEL_THREADS;
#pragma unroll
    for (int u = 0; u < 3; u++) {
        const uint32_t e = 3 * g0 + u * MODEL_THREADS + threadIdx.x;
        if (e < 3 * P) {
            const uint32_t j = e / 3;
            const bool jr = j < n0;
            const size_t ee = jr ? e : e - 3 * n0;
            const RawSet sj = select_set(m, jr);
            const ScgrModelSetGrads dj = select_grads(out, jr);
            if (!sj.rayo) dj.dL_dxyz[ee] = __ldg(g.dL_dmeans3D + e);                                  // free position
            dj.dL_dscaling[ee] = __ldg(g.dL_dscales + e) * expf(__ldg(sj.scaling + ee));              // exp: dL/draw = dL/dscale * scale
        }
    }
    const uint32_t i = g0 + threadIdx.x;
    if (i >= P) return;
    const bool ray = i < n0;
    const size_t ii = ray ? i : i - n0;
    const RawSet s = select_set(m, ray);
    const ScgrModelSetGrads d = select_grads(out, ray);

    // ray-based position: d zval = <rayd, dL/dxyz> (rayo, rayd are fixed ray geometry: reference :493 optimises zval only)
    if (s.rayo) {
        const float px = __fmul_rn(__ldg(g.dL_dmeans3D + 3 * (size_t)i + 0), __ldg(s.rayd + 3 * ii + 0)),
                    py = __fmul_rn(__ldg(g.dL_dmeans3D + 3 * (size_t)i + 1), __ldg(s.rayd + 3 * ii + 1)),
                    pz = __fmul_rn(__ldg(g.dL_dmeans3D + 3 * (size_t)i + 2), __ldg(s.rayd + 3 * ii + 2));
        d.dL_dzval[ii] = __fadd_rn(__fadd_rn(px, py), pz);
    }

    // normalize: y = q / max(|q|, eps);  dL/dq = (g - y <y, g>) / |q|  (clamp active: dL/dq = g / eps)
    const float4 q = __ldg(reinterpret_cast<const float4*>(s.rotation) + ii);
    const float4 gr = __ldg(reinterpret_cast<const float4*>(g.dL_drotations) + i);
    const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float4 dq;
    if (nrm > 1e-12f) {
        const float inv = 1.0f / nrm;
        const float4 y = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
        const float dot = y.x * gr.x + y.y * gr.y + y.z * gr.z + y.w * gr.w;
        dq = make_float4((gr.x - y.x * dot) * inv, (gr.y - y.y * dot) * inv, (gr.z - y.z * dot) * inv,
                         (gr.w - y.w * dot) * inv);
    } else {
        dq = make_float4(gr.x * 1e12f, gr.y * 1e12f, gr.z * 1e12f, gr.w * 1e12f);
    }
    reinterpret_cast<float4*>(d.dL_drotation)[ii] = dq;

    // sigmoid: dL/draw = dL/dopacity * (1 - y) * y
    const float y = sigmoid_f(__ldg(s.opacity + ii));
    d.dL_dopacity[ii] = __ldg(g.dL_dopacities + i) * (1.0f - y) * y;
}

// ---- what the training step does with the operator's outputs every iteration before the optimizer step
// (reference train.py:192-193 -> scene/gaussian_model.py:932-934):
//   max_radii2D[vis] = max(max_radii2D[vis], radii[vis]);  xyz_gradient_accum[vis] += |dL/dmean2D[vis, :2]|;  denom[vis] += 1
// The reference's boolean-mask indexing costs a nonzero + host synchronisation per statement (6 per iteration);
// here it is one elementwise launch with no host involvement.  36 B read + 12 B written per visible Gaussian.
__global__ void __launch_bounds__(MODEL_THREADS)
densification_stats_kernel(const float* __restrict__ g2d, const uint8_t* __restrict__ filter,
                           const int32_t* __restrict__ radii, const int32_t P, float* __restrict__ accum,
                           float* __restrict__ denom, float* __restrict__ max_radii) {
    const int i = blockIdx.x * MODEL_THREADS + threadIdx.x;
    if (i >= P) return;
    const int r = radii ? __ldg(radii + i) : 0;
    const bool vis = filter ? __ldg(filter + i) != 0 : r > 0;
    if (!vis) return;
    if (radii && max_radii) max_radii[i] = fmaxf(max_radii[i], (float)r);
    if (accum) {
        const float gx = __ldg(g2d + 3 * (size_t)i), gy = __ldg(g2d + 3 * (size_t)i + 1);
        accum[i] += sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
        denom[i] += 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------------
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;                                    // float4 per thread per array
constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC * 4;        // elements per CTA

struct AdamSeg {
    float* p;
    const float* g;
    float* m;
    float* v;
    uint32_t n;
    uint32_t first_block;
    float step_size;      // lr / (1 - beta1^step)
    float bc2_sqrt;       // 1 / sqrt(1 - beta2^step)
};
struct AdamTable {
    AdamSeg seg[SCGR_ADAM_MAX_GROUPS];
    int n_seg;
    float beta2, w1, w2, eps;   // w1 = 1 - beta1, w2 = 1 - beta2
};

// torch/optim/adam.py _single_tensor_adam, in its order of operations:
//   exp_avg.lerp_(grad, 1 - beta1);  exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
//   denom = exp_avg_sq.sqrt() / sqrt(bias_correction2) + eps;  param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
// The square root and the quotient use the hardware approximations (sqrt.approx: 1 ulp, div.approx: 2 ulp): IEEE
// sqrtf / division are multi-instruction sequences with slow-path branches for zeros, denormals and extreme exponents --
// exactly what the moments of the many Gaussians that receive no gradient in a view hold (exp_avg_sq == 0, exp_avg
// decaying into the denormals).  In a training step that made this kernel issue-bound: 458 us and 230 M warp
// instructions for 1.6 GB, against 254 us on dense random data (profiles/r02_model_kernels.md).  The approximations
// are branch-free for every input and move the update by < 3e-7 of its own size.
__device__ __forceinline__ float sqrt_approx(const float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float div_approx(const float a, const float b) {
    float y;
    asm("div.approx.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
    return y;
}
__device__ __forceinline__ void adam_update(float& p, const float g, float& m, float& v, const float beta2,
                                            const float w1, const float w2, const float eps, const float step_size,
                                            const float inv_bc2_sqrt) {
    m = m + w1 * (g - m);
    v = v * beta2 + w2 * g * g;
    const float denom = sqrt_approx(v) * inv_bc2_sqrt + eps;
    p = p - step_size * div_approx(m, denom);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const __grid_constant__ AdamTable t) {
    // which parameter group does this CTA work on?  (static indices only: the table stays in constant memory)
    float* p = t.seg[0].p; const float* g = t.seg[0].g; float* m = t.seg[0].m; float* v = t.seg[0].v;
    uint32_t n = t.seg[0].n, first = 0;
    float step_size = t.seg[0].step_size, bc2_sqrt = t.seg[0].bc2_sqrt;
#pragma unroll
    for (int k = 1; k < SCGR_ADAM_MAX_GROUPS; k++) {
        if (k < t.n_seg && blockIdx.x >= t.seg[k].first_block) {
            p = t.seg[k].p; g = t.seg[k].g; m = t.seg[k].m; v = t.seg[k].v;
            n = t.seg[k].n; first = t.seg[k].first_block;
            step_size = t.seg[k].step_size; bc2_sqrt = t.seg[k].bc2_sqrt;
        }
    }
    const uint32_t base = (blockIdx.x - first) * (uint32_t)ADAM_CHUNK;
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15u) == 0;
    if (aligned && base + ADAM_CHUNK <= n) {
        // full chunk: all loads of the 4 rounds in flight before the first store
        float4 P4[ADAM_VEC], G4[ADAM_VEC], M4[ADAM_VEC], V4[ADAM_VEC];
#pragma unroll
        for (int u = 0; u < ADAM_VEC; u++) {
            const uint32_t e = base + (u * ADAM_THREADS + threadIdx.x) * 4u;
            P4[u] = *reinterpret_cast<const float4*>(p + e);
            G4[u] = __ldg(reinterpret_cast<const float4*>(g + e));
            M4[u] = *reinterpret_cast<const float4*>(m + e);
            V4[u] = *reinterpret_cast<const float4*>(v + e);
        }
#pragma unroll
        for (int u = 0; u < ADAM_VEC; u++) {
            const uint32_t e = base + (u * ADAM_THREADS + threadIdx.x) * 4u;
            adam_update(P4[u].x, G4[u].x, M4[u].x, V4[u].x, t.beta2, t.w1, t.w2, t.eps, step_size, bc2_sqrt);
            adam_update(P4[u].y, G4[u].y, M4[u].y, V4[u].y, t.beta2, t.w1, t.w2, t.eps, step_size, bc2_sqrt);
            adam_update(P4[u].z, G4[u].z, M4[u].z, V4[u].z, t.beta2, t.w1, t.w2, t.eps, step_size, bc2_sqrt);
            adam_update(P4[u].w, G4[u].w, M4[u].w, V4[u].w, t.beta2, t.w1, t.w2, t.eps, step_size, bc2_sqrt);
            *reinterpret_cast<float4*>(p + e) = P4[u];
            *reinterpret_cast<float4*>(m + e) = M4[u];
            *reinterpret_cast<float4*>(v + e) = V4[u];
        }
        return;
    }
    // ragged last chunk of a group, or a group that is not 16-byte aligned (a view at an odd offset)
    const uint32_t end = min(n, base + (uint32_t)ADAM_CHUNK);
    for (uint32_t e = base + threadIdx.x; e < end; e += ADAM_THREADS) {
        float pe = p[e], me = m[e], ve = v[e];
        adam_update(pe, __ldg(g + e), me, ve, t.beta2, t.w1, t.w2, t.eps, step_size, bc2_sqrt);
        p[e] = pe; m[e] = me; v[e] = ve;
    }
}

// ---- prune compaction (reference scene/gaussian_model.py:777-820 `_prune_optimizer` / `prune_points`; SURVEY.md section
// 8f row f4): dst[a][j, :] = src[a][index[j], :] for EVERY per-Gaussian array a of the model -- the 12 parameters, their
// exp_avg / exp_avg_sq, the ray geometry and the densification statistics -- in one launch from one index list, where
// the reference evaluates `t[mask]` (a nonzero + host synchronisation + gather) ~45 times.  Flat over the output floats
// of each array: coalesced stores, loads coalesced within a row.
constexpr int GATHER_PER_THREAD = 8;
constexpr int GATHER_PER_BLOCK = MODEL_THREADS * GATHER_PER_THREAD;

struct GatherSeg {
    const float* src;
    float* dst;
    uint32_t row;           // floats per row
    uint32_t first_block;
};
struct GatherTable {
    GatherSeg seg[SCGR_GATHER_MAX_ARRAYS];
    int n_seg;
};

__global__ void __launch_bounds__(MODEL_THREADS)
gather_rows_kernel(const __grid_constant__ GatherTable t, const int64_t* __restrict__ index, const uint32_t n_out) {
    const float* src = t.seg[0].src; float* dst = t.seg[0].dst;
    uint32_t row = t.seg[0].row, first = 0;
#pragma unroll
    for (int k = 1; k < SCGR_GATHER_MAX_ARRAYS; k++) {
        if (k < t.n_seg && blockIdx.x >= t.seg[k].first_block) {
            src = t.seg[k].src; dst = t.seg[k].dst; row = t.seg[k].row; first = t.seg[k].first_block;
        }
    }
    const uint64_t total = (uint64_t)n_out * row;
    const uint64_t e = (uint64_t)(blockIdx.x - first) * GATHER_PER_BLOCK + threadIdx.x;
    uint32_t j = (uint32_t)(e / row), c = (uint32_t)(e - (uint64_t)j * row);
    const uint32_t step_j = MODEL_THREADS / row, step_c = MODEL_THREADS - step_j * row;
    float v[GATHER_PER_THREAD];
#pragma unroll
    for (int u = 0; u < GATHER_PER_THREAD; u++) {
        v[u] = 0.f;
        if (e + (uint64_t)u * MODEL_THREADS < total) v[u] = __ldg(src + (size_t)__ldg(index + j) * row + c);
        c += step_c; j += step_j;
        if (c >= row) { c -= row; j++; }
    }
#pragma unroll
    for (int u = 0; u < GATHER_PER_THREAD; u++)
        if (e + (uint64_t)u * MODEL_THREADS < total) dst[e + (uint64_t)u * MODEL_THREADS] = v[u];
}

// ---- densification append (reference scene/gaussian_model.py:822-862 `cat_tensors_to_optimizer` /
// `densification_postfix`): the new Gaussians are appended to the 6 parameter groups of
// the free set, their moments are extended with zeros and the statistics are reset -- 18 torch.cat + 15 zero fills in
// the reference.  Here: ONE launch over a table of flat segments {dst, src or NULL (= zeros), n}.
constexpr int COPY_THREADS = 256;
constexpr int COPY_CHUNK = COPY_THREADS * 4 * 4;     // floats per CTA: 4 float4 per thread

struct CopySeg {
    float* dst;
    const float* src;
    uint64_t n;
    uint32_t first_block;
};
struct CopyTable {
    CopySeg seg[SCGR_COPY_MAX_SEGMENTS];
    int n_seg;
};

__global__ void __launch_bounds__(COPY_THREADS)
copy_segments_kernel(const __grid_constant__ CopyTable t) {
    float* dst = t.seg[0].dst; const float* src = t.seg[0].src;
    uint64_t n = t.seg[0].n;
    uint32_t first = 0;
#pragma unroll
    for (int k = 1; k < SCGR_COPY_MAX_SEGMENTS; k++) {
        if (k < t.n_seg && blockIdx.x >= t.seg[k].first_block) {
            dst = t.seg[k].dst; src = t.seg[k].src; n = t.seg[k].n; first = t.seg[k].first_block;
        }
    }
    const uint64_t base = (uint64_t)(blockIdx.x - first) * COPY_CHUNK;
    const bool aligned = ((((uintptr_t)dst) | ((uintptr_t)src)) & 15u) == 0;     // a NULL src counts as aligned
    if (aligned && base + COPY_CHUNK <= n) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t e = base + (uint64_t)(u * COPY_THREADS + threadIdx.x) * 4u;
            v[u] = src ? __ldg(reinterpret_cast<const float4*>(src + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t e = base + (uint64_t)(u * COPY_THREADS + threadIdx.x) * 4u;
            *reinterpret_cast<float4*>(dst + e) = v[u];
        }
        return;
    }
    const uint64_t end = min(n, base + (uint64_t)COPY_CHUNK);
    for (uint64_t e = base + threadIdx.x; e < end; e += COPY_THREADS) dst[e] = src ? __ldg(src + e) : 0.f;
}

uint32_t sh_block_count(uint64_t sh_total) { return (uint32_t)((sh_total + SH_PER_BLOCK - 1) / SH_PER_BLOCK); }

}  // namespace

// SCGR_ASSEMBLE_STAGED=0 keeps the flat copy for every K (A/B switch)
static int staged_rows(uint32_t n3k) {
    static const bool off = getenv("SCGR_ASSEMBLE_STAGED") && atoi(getenv("SCGR_ASSEMBLE_STAGED")) == 0;
    return (!off && (n3k == 48 || n3k == 12)) ? (int)n3k : 0;
}

void launch_assemble_forward(const ScgrModel& m, const ScgrActivated& out, const Launch& L) {
    const uint64_t P = (uint64_t)m.set[0].n + (uint64_t)m.set[1].n;
    if (P == 0) return;
    const uint32_t n3k = 3u * (uint32_t)(m.sh_rest + 1);
    const uint64_t sh_total = P * n3k;
    const int staged = staged_rows(n3k);
    // out.shs == NULL: the operator reads the SH rows straight from features_dc / features_rest (ScgrGaussians.sh_dc /
    // sh_rest, SURVEY 8f row f2 second half) -- no SH CTAs, 192 of the 236 assembled bytes per Gaussian never exist
    const uint32_t sh_blocks = !out.shs ? 0u
                               : staged ? (uint32_t)((m.set[0].n + STAGE_G - 1) / STAGE_G + (m.set[1].n + STAGE_G - 1) / STAGE_G)
                                        : sh_block_count(sh_total);
    const uint32_t blocks = sh_blocks + (uint32_t)((P + MODEL_THREADS - 1) / MODEL_THREADS);
    begin_kernel("assemble_forward", L);
    if (staged == 48)
        assemble_forward_kernel<48><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, out, sh_blocks, n3k, (uint32_t)sh_total);
    else if (staged == 12)
        assemble_forward_kernel<12><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, out, sh_blocks, n3k, (uint32_t)sh_total);
    else
        assemble_forward_kernel<0><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, out, sh_blocks, n3k, (uint32_t)sh_total);
    check_launch("assemble_forward", L);
}

void launch_assemble_backward(const ScgrModel& m, const ScgrActivatedGrads& g, const ScgrModelGrads& out,
                              const Launch& L) {
    const uint64_t P = (uint64_t)m.set[0].n + (uint64_t)m.set[1].n;
    if (P == 0) return;
    const uint32_t n3k = 3u * (uint32_t)(m.sh_rest + 1);
    const uint64_t sh_total = P * n3k;
    const int staged = staged_rows(n3k);
    const uint32_t sh_blocks = !g.dL_dshs ? 0u      // split SH layout: scgr_backward wrote dL/dfeatures_* itself
                               : staged ? (uint32_t)((m.set[0].n + STAGE_G - 1) / STAGE_G + (m.set[1].n + STAGE_G - 1) / STAGE_G)
                                        : sh_block_count(sh_total);
    const uint32_t blocks = sh_blocks + (uint32_t)((P + MODEL_THREADS - 1) / MODEL_THREADS);
    begin_kernel("assemble_backward", L);
    if (staged == 48)
        assemble_backward_kernel<48><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, g, out, sh_blocks, n3k, (uint32_t)sh_total);
    else if (staged == 12)
        assemble_backward_kernel<12><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, g, out, sh_blocks, n3k, (uint32_t)sh_total);
    else
        assemble_backward_kernel<0><<<blocks, MODEL_THREADS, 0, L.stream>>>(m, g, out, sh_blocks, n3k, (uint32_t)sh_total);
    check_launch("assemble_backward", L);
}

void launch_densification_stats(const float* dL_dmeans2D, const uint8_t* update_filter, const int32_t* radii, int32_t P,
                                float* xyz_gradient_accum, float* denom, float* max_radii2D, const Launch& L) {
    if (P <= 0) return;
    begin_kernel("densification_stats", L);
    densification_stats_kernel<<<(P + MODEL_THREADS - 1) / MODEL_THREADS, MODEL_THREADS, 0, L.stream>>>(
        dL_dmeans2D, update_filter, radii, P, xyz_gradient_accum, denom, max_radii2D);
    check_launch("densification_stats", L);
}

void launch_gather_rows(const ScgrRowGather* arrays, int32_t n_arrays, const int64_t* index, int64_t n_out,
                        const Launch& L) {
    GatherTable t{};
    uint64_t blocks = 0;
    int k = 0;
    for (int a = 0; a < n_arrays; a++) {
        if (arrays[a].row_floats == 0) continue;
        GatherSeg& s = t.seg[k++];
        s.src = arrays[a].src; s.dst = arrays[a].dst;
        s.row = (uint32_t)arrays[a].row_floats;
        s.first_block = (uint32_t)blocks;
        blocks += ((uint64_t)n_out * s.row + GATHER_PER_BLOCK - 1) / GATHER_PER_BLOCK;
    }
    if (k == 0 || n_out == 0) return;
    t.n_seg = k;
    begin_kernel("gather_rows", L);
    gather_rows_kernel<<<(uint32_t)blocks, MODEL_THREADS, 0, L.stream>>>(t, index, (uint32_t)n_out);
    check_launch("gather_rows", L);
}

void launch_copy_segments(const ScgrSegmentCopy* segs, int32_t n_segs, const Launch& L) {
    CopyTable t{};
    uint64_t blocks = 0;
    int k = 0;
    for (int i = 0; i < n_segs; i++) {
        if (segs[i].n_floats == 0) continue;
        CopySeg& c = t.seg[k++];
        c.dst = segs[i].dst; c.src = segs[i].src;
        c.n = (uint64_t)segs[i].n_floats;
        c.first_block = (uint32_t)blocks;
        blocks += (c.n + COPY_CHUNK - 1) / COPY_CHUNK;
    }
    if (k == 0) return;
    t.n_seg = k;
    begin_kernel("copy_segments", L);
    copy_segments_kernel<<<(uint32_t)blocks, COPY_THREADS, 0, L.stream>>>(t);
    check_launch("copy_segments", L);
}

void launch_adam(const ScgrAdamGroup* groups, int32_t n_groups, double beta1, double beta2, double eps,
                 const Launch& L) {
    AdamTable t{};
    uint64_t blocks = 0;
    int k = 0;
    for (int i = 0; i < n_groups; i++) {
        const ScgrAdamGroup& G = groups[i];
        if (G.n == 0) continue;
        const double bc1 = 1.0 - pow(beta1, (double)G.step);
        const double bc2 = 1.0 - pow(beta2, (double)G.step);
        AdamSeg& s = t.seg[k++];
        s.p = G.param; s.g = G.grad; s.m = G.exp_avg; s.v = G.exp_avg_sq;
        s.n = (uint32_t)G.n;
        s.first_block = (uint32_t)blocks;
        s.step_size = (float)(G.lr / bc1);
        s.bc2_sqrt = (float)(1.0 / sqrt(bc2));
        blocks += ((uint64_t)G.n + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    if (k == 0) return;
    t.n_seg = k;
    t.beta2 = (float)beta2;
    t.w1 = (float)(1.0 - beta1);
    t.w2 = (float)(1.0 - beta2);
    t.eps = (float)eps;
    begin_kernel("adam", L);
    adam_kernel<<<(uint32_t)blocks, ADAM_THREADS, 0, L.stream>>>(t);
    check_launch("adam", L);
}

}  // namespace scgr
