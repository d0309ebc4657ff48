// standin.cu -- BASELINE / TEST INFRASTRUCTURE, not the product and NOT the reference.
//
// A deliberately naive CUDA restatement of the published 3D-Gaussian-splatting rasterizer algorithm (SURVEY.md
// Appendix A.1-A.10), structured the way generic implementations of it are: one thread per Gaussian for the
// per-Gaussian stages, an inclusive scan + BLOCKING device-to-host read of the instance count, one 64-bit
// (tile | depth) key per instance sorted by cub::DeviceRadixSort::SortPairs on 32 + bits(tiles) key bits, one
// 256-thread CTA per 16x16 tile that stages 256 list entries at a time in shared memory, and TEN global atomics per
// contributing (pixel, Gaussian) pair in the backward.  It answers "what do generic kernels cost on the same B200"
// (SURVEY.md section 8d option 2-ii); bench.py reports it as `gpu_standin_baseline`.
// The reference's own rasterizer (external package, reference README.md:23-25) is not obtainable offline.
// The arithmetic is that of oracle/scg_oracle.c (the same author's CPU restatement), against which
// tests/test_standin.py checks it: a third implementation of Appendix A.
//
// Never imported by scgaussian_b200/.  Keeps its scratch in grow-only device buffers owned by the library (the
// reference lets torch's caching allocator own them: neither pays a cudaMalloc in steady state).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

namespace standin {

constexpr int BLK = 16;
constexpr int BLOCK_SIZE = BLK * BLK;

__constant__ float kSH_C0 = 0.28209479177387814f;
__constant__ float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                    -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                    0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f};

struct Params {
    int P, M, D, W, H;
    float tanfovx, tanfovy, scale_modifier;
    const float *bg, *viewmatrix, *projmatrix, *campos;
    const float *means3D, *opacities, *shs, *scales, *rotations;
};

struct Buffers {   // grow-only scratch
    void* ptr[24] = {};
    size_t cap[24] = {};
};
static Buffers g_buf;
template <typename T>
static T* take(int slot, size_t n) {
    const size_t bytes = (n > 0 ? n : 1) * sizeof(T);
    if (g_buf.cap[slot] < bytes) {
        if (g_buf.ptr[slot]) cudaFree(g_buf.ptr[slot]);
        cudaMalloc(&g_buf.ptr[slot], bytes + bytes / 4);
        g_buf.cap[slot] = bytes + bytes / 4;
    }
    return reinterpret_cast<T*>(g_buf.ptr[slot]);
}

struct State {
    Params p;
    int gx, gy;
    float* depths; float2* xy; float4* conic_opacity; float* rgb; float* cov3D; int* radii_i; bool* clamped;
    uint32_t* tiles_touched; uint32_t* offsets;
    uint64_t *keys, *keys_sorted; uint32_t *vals, *vals_sorted; uint2* ranges; void* cub_tmp; size_t cub_tmp_bytes;
    uint32_t* n_contrib; float* final_T;
    int64_t R;
    // backward internals
    float *dL_dmean2D, *dL_dconic, *dL_dcolors, *dL_ddepths;
};
static State g;

__device__ inline float3 tp43(const float3 p, const float* m) {   // hom @ M (row-vector convention), xyz
    return make_float3(p.x * m[0] + p.y * m[4] + p.z * m[8] + m[12], p.x * m[1] + p.y * m[5] + p.z * m[9] + m[13],
                       p.x * m[2] + p.y * m[6] + p.z * m[10] + m[14]);
}

__device__ inline void quat_R(const float* q, float R[9]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z); R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y); R[7] = 2.f * (y * z + r * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

__device__ inline void cov3d_of(const float* scale, float mod, const float* q, float* c6) {
    float R[9], L[9];
    quat_R(q, R);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L[i * 3 + j] = R[i * 3 + j] * (mod * scale[j]);
    float S[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        float a = 0; for (int k = 0; k < 3; k++) a += L[i * 3 + k] * L[j * 3 + k]; S[i * 3 + j] = a;
    }
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

struct Proj {
    float t[3], tx, ty; int xin, yin; float M0[3], M1[3], a, b, c, fx, fy;
};
__device__ inline void project_cov(const Params& in, const float* p, const float* c6, Proj* o) {
    const float* V = in.viewmatrix;
    for (int j = 0; j < 3; j++) o->t[j] = p[0] * V[0 * 4 + j] + p[1] * V[1 * 4 + j] + p[2] * V[2 * 4 + j] + V[3 * 4 + j];
    o->fx = in.W / (2 * in.tanfovx);
    o->fy = in.H / (2 * in.tanfovy);
    const float limx = 1.3f * in.tanfovx, limy = 1.3f * in.tanfovy, tz = o->t[2];
    const float txtz = o->t[0] / tz, tytz = o->t[1] / tz;
    o->xin = (txtz >= -limx && txtz <= limx);
    o->yin = (tytz >= -limy && tytz <= limy);
    o->tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    o->ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float J[6] = {o->fx / tz, 0, -o->fx * o->tx / (tz * tz), 0, o->fy / tz, -o->fy * o->ty / (tz * tz)};
    for (int k = 0; k < 3; k++) {
        o->M0[k] = J[0] * V[k * 4 + 0] + J[1] * V[k * 4 + 1] + J[2] * V[k * 4 + 2];
        o->M1[k] = J[3] * V[k * 4 + 0] + J[4] * V[k * 4 + 1] + J[5] * V[k * 4 + 2];
    }
    const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float SM0[3], SM1[3];
    for (int i = 0; i < 3; i++) {
        SM0[i] = S[i * 3] * o->M0[0] + S[i * 3 + 1] * o->M0[1] + S[i * 3 + 2] * o->M0[2];
        SM1[i] = S[i * 3] * o->M1[0] + S[i * 3 + 1] * o->M1[1] + S[i * 3 + 2] * o->M1[2];
    }
    o->a = o->M0[0] * SM0[0] + o->M0[1] * SM0[1] + o->M0[2] * SM0[2] + 0.3f;
    o->b = o->M0[0] * SM1[0] + o->M0[1] * SM1[1] + o->M0[2] * SM1[2];
    o->c = o->M1[0] * SM1[0] + o->M1[1] * SM1[1] + o->M1[2] * SM1[2] + 0.3f;
}

__device__ inline void sh_basis(int D, float x, float y, float z, float* b, float (*db)[3]) {
    for (int k = 0; k < 16; k++) { b[k] = 0; db[k][0] = db[k][1] = db[k][2] = 0; }
    b[0] = kSH_C0;
    if (D < 1) return;
    b[1] = -kSH_C1 * y; db[1][1] = -kSH_C1;
    b[2] = kSH_C1 * z;  db[2][2] = kSH_C1;
    b[3] = -kSH_C1 * x; db[3][0] = -kSH_C1;
    if (D < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = kSH_C2[0] * xy; db[4][0] = kSH_C2[0] * y; db[4][1] = kSH_C2[0] * x;
    b[5] = kSH_C2[1] * yz; db[5][1] = kSH_C2[1] * z; db[5][2] = kSH_C2[1] * y;
    b[6] = kSH_C2[2] * (2 * zz - xx - yy); db[6][0] = kSH_C2[2] * -2 * x; db[6][1] = kSH_C2[2] * -2 * y; db[6][2] = kSH_C2[2] * 4 * z;
    b[7] = kSH_C2[3] * xz; db[7][0] = kSH_C2[3] * z; db[7][2] = kSH_C2[3] * x;
    b[8] = kSH_C2[4] * (xx - yy); db[8][0] = kSH_C2[4] * 2 * x; db[8][1] = kSH_C2[4] * -2 * y;
    if (D < 3) return;
    b[9] = kSH_C3[0] * y * (3 * xx - yy); db[9][0] = kSH_C3[0] * 6 * xy; db[9][1] = kSH_C3[0] * (3 * xx - 3 * yy);
    b[10] = kSH_C3[1] * xy * z; db[10][0] = kSH_C3[1] * yz; db[10][1] = kSH_C3[1] * xz; db[10][2] = kSH_C3[1] * xy;
    b[11] = kSH_C3[2] * y * (4 * zz - xx - yy);
    db[11][0] = kSH_C3[2] * -2 * xy; db[11][1] = kSH_C3[2] * (4 * zz - xx - 3 * yy); db[11][2] = kSH_C3[2] * 8 * yz;
    b[12] = kSH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
    db[12][0] = kSH_C3[3] * -6 * xz; db[12][1] = kSH_C3[3] * -6 * yz; db[12][2] = kSH_C3[3] * (6 * zz - 3 * xx - 3 * yy);
    b[13] = kSH_C3[4] * x * (4 * zz - xx - yy);
    db[13][0] = kSH_C3[4] * (4 * zz - 3 * xx - yy); db[13][1] = kSH_C3[4] * -2 * xy; db[13][2] = kSH_C3[4] * 8 * xz;
    b[14] = kSH_C3[5] * z * (xx - yy); db[14][0] = kSH_C3[5] * 2 * xz; db[14][1] = kSH_C3[5] * -2 * yz; db[14][2] = kSH_C3[5] * (xx - yy);
    b[15] = kSH_C3[6] * x * (xx - 3 * yy); db[15][0] = kSH_C3[6] * (3 * xx - 3 * yy); db[15][1] = kSH_C3[6] * -6 * xy;
}

// ---- A.1-A.5: one thread per Gaussian ----
__global__ void preprocess_kernel(Params in, int gx, int gy, float* depths, float2* xy, float4* conic_opacity, float* rgb,
                                  float* cov3D, int* radii, bool* clamped, uint32_t* tiles_touched) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.P) return;
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float* p = in.means3D + 3 * i;
    const float *V = in.viewmatrix, *PM = in.projmatrix;
    const float zv = p[0] * V[2] + p[1] * V[6] + p[2] * V[10] + V[14];
    if (zv <= 0.2f) return;
    float hom[4];
    for (int j = 0; j < 4; j++) hom[j] = p[0] * PM[j] + p[1] * PM[4 + j] + p[2] * PM[8 + j] + PM[12 + j];
    const float pw = 1.f / (hom[3] + 1e-7f);
    const float ndcx = hom[0] * pw, ndcy = hom[1] * pw;
    float* c6 = cov3D + 6 * i;
    cov3d_of(in.scales + 3 * i, in.scale_modifier, in.rotations + 4 * i, c6);
    Proj pr;
    project_cov(in, p, c6, &pr);
    const float det = pr.a * pr.c - pr.b * pr.b;
    if (det == 0.f) return;
    const float dinv = 1.f / det;
    const float mid = 0.5f * (pr.a + pr.c);
    const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lam = fmaxf(mid + disc, mid - disc);
    const int rad = (int)ceilf(3.f * sqrtf(lam));
    const float px = ((ndcx + 1) * in.W - 1) * 0.5f, py = ((ndcy + 1) * in.H - 1) * 0.5f;
    const int x0 = min(gx, max(0, (int)((px - rad) / BLK))), y0 = min(gy, max(0, (int)((py - rad) / BLK)));
    const int x1 = min(gx, max(0, (int)((px + rad + BLK - 1) / BLK))), y1 = min(gy, max(0, (int)((py + rad + BLK - 1) / BLK)));
    if ((x1 - x0) * (y1 - y0) == 0) return;
    {
        float dx = p[0] - in.campos[0], dy = p[1] - in.campos[1], dz = p[2] - in.campos[2];
        const float n = sqrtf(dx * dx + dy * dy + dz * dz);
        float b[16], db[16][3];
        sh_basis(in.D, dx / n, dy / n, dz / n, b, db);
        const int nk = (in.D + 1) * (in.D + 1);
        for (int ch = 0; ch < 3; ch++) {
            float acc = 0;
            for (int k = 0; k < nk; k++) acc += b[k] * in.shs[((size_t)i * in.M + k) * 3 + ch];
            const float v = acc + 0.5f;
            clamped[3 * i + ch] = v < 0;
            rgb[3 * i + ch] = fmaxf(v, 0.f);
        }
    }
    depths[i] = zv;
    radii[i] = rad;
    xy[i] = make_float2(px, py);
    conic_opacity[i] = make_float4(pr.c * dinv, -pr.b * dinv, pr.a * dinv, in.opacities[i]);
    tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
}

// ---- A.6: one thread per Gaussian writes all of its (tile | depth) keys ----
__global__ void duplicate_with_keys(int P, const float2* xy, const float* depths, const uint32_t* offsets, const int* radii,
                                    int gx, int gy, uint64_t* keys, uint32_t* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || radii[i] <= 0) return;
    uint32_t off = i == 0 ? 0 : offsets[i - 1];
    const float px = xy[i].x, py = xy[i].y;
    const int rad = radii[i];
    const int x0 = min(gx, max(0, (int)((px - rad) / BLK))), y0 = min(gy, max(0, (int)((py - rad) / BLK)));
    const int x1 = min(gx, max(0, (int)((px + rad + BLK - 1) / BLK))), y1 = min(gy, max(0, (int)((py + rad + BLK - 1) / BLK)));
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            uint64_t key = (uint64_t)(y * gx + x);
            key <<= 32;
            key |= (uint64_t)__float_as_uint(depths[i]);
            keys[off] = key;
            vals[off] = (uint32_t)i;
            off++;
        }
}

// ---- A.7 ----
__global__ void identify_tile_ranges(int64_t L, const uint64_t* keys, uint2* ranges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[t].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (t != prev) { ranges[prev].y = (uint32_t)i; ranges[t].x = (uint32_t)i; }
    }
    if (i == L - 1) ranges[t].y = (uint32_t)L;
}

// ---- A.8: one 256-thread CTA per tile, one thread per pixel ----
__global__ void __launch_bounds__(BLOCK_SIZE)
render_forward_kernel(const uint2* ranges, const uint32_t* point_list, int W, int H, const float2* xy, const float* rgb,
                      const float* depths, const float4* conic_opacity, const float* bg, uint32_t* n_contrib, float* final_T,
                      float* out_color, float* out_depth, float* out_alpha) {
    const int gx = (W + BLK - 1) / BLK;
    const int px = blockIdx.x * BLK + threadIdx.x, py = blockIdx.y * BLK + threadIdx.y;
    const bool inside = px < W && py < H;
    const size_t pid = (size_t)py * W + px;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BLOCK_SIZE - 1) / BLOCK_SIZE;
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    __shared__ int s_id[BLOCK_SIZE];
    __shared__ float2 s_xy[BLOCK_SIZE];
    __shared__ float4 s_co[BLOCK_SIZE];
    float T = 1.f, C[3] = {0, 0, 0}, Dsum = 0.f, Wsum = 0.f;
    uint32_t contributor = 0, last = 0;
    const int tid = threadIdx.y * BLK + threadIdx.x;
    for (int r = 0; r < rounds; r++, todo -= BLOCK_SIZE) {
        if (__syncthreads_count(done) == BLOCK_SIZE) break;
        const int progress = r * BLOCK_SIZE + tid;
        if (range.x + progress < range.y) {
            const int id = (int)point_list[range.x + progress];
            s_id[tid] = id; s_xy[tid] = xy[id]; s_co[tid] = conic_opacity[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK_SIZE, todo); j++) {
            contributor++;
            const float dx = s_xy[j].x - (float)px, dy = s_xy[j].y - (float)py;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.f / 255.f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float w = alpha * T;
            const int id = s_id[j];
            for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * id + ch] * w;
            Dsum += depths[id] * w;
            Wsum += w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t N = (size_t)W * H;
        final_T[pid] = T;
        n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ch++) out_color[ch * N + pid] = C[ch] + T * bg[ch];
        out_depth[pid] = Dsum;
        out_alpha[pid] = Wsum;
    }
}

// ---- A.9: one CTA per tile, back to front, ten global atomics per contributing (pixel, Gaussian) pair ----
__global__ void __launch_bounds__(BLOCK_SIZE)
render_backward_kernel(const uint2* ranges, const uint32_t* point_list, int W, int H, const float* bg, const float2* xy,
                       const float4* conic_opacity, const float* rgb, const float* depths, const float* final_Ts,
                       const uint32_t* n_contrib, const float* gC, const float* gD, const float* gA, float* dL_dmean2D,
                       float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_ddepths) {
    const int gx = (W + BLK - 1) / BLK;
    const int px = blockIdx.x * BLK + threadIdx.x, py = blockIdx.y * BLK + threadIdx.y;
    const bool inside = px < W && py < H;
    const size_t pid = (size_t)py * W + px, N = (size_t)W * H;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BLOCK_SIZE - 1) / BLOCK_SIZE;
    bool done = !inside;
    int todo = (int)(range.y - range.x);
    __shared__ int s_id[BLOCK_SIZE];
    __shared__ float2 s_xy[BLOCK_SIZE];
    __shared__ float4 s_co[BLOCK_SIZE];
    __shared__ float s_rgb[3 * BLOCK_SIZE];
    __shared__ float s_dep[BLOCK_SIZE];
    const float T_final = inside ? final_Ts[pid] : 0.f;
    float T = T_final;
    uint32_t contributor = (uint32_t)todo;
    const int last_contributor = inside ? (int)n_contrib[pid] : 0;
    float g3[3] = {0, 0, 0}, gd = 0.f, ga = 0.f;
    if (inside) { for (int ch = 0; ch < 3; ch++) g3[ch] = gC[ch * N + pid]; gd = gD[pid]; ga = gA[pid]; }
    const float bgdot = bg[0] * g3[0] + bg[1] * g3[1] + bg[2] * g3[2];
    float arec[3] = {0, 0, 0}, drec = 0.f, alrec = 0.f, last_alpha = 0.f, last_c[3] = {0, 0, 0}, last_d = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int tid = threadIdx.y * BLK + threadIdx.x;
    for (int r = 0; r < rounds; r++, todo -= BLOCK_SIZE) {
        __syncthreads();
        const int progress = r * BLOCK_SIZE + tid;
        if (range.x + progress < range.y) {
            const int id = (int)point_list[range.y - progress - 1];
            s_id[tid] = id; s_xy[tid] = xy[id]; s_co[tid] = conic_opacity[id];
            for (int ch = 0; ch < 3; ch++) s_rgb[ch * BLOCK_SIZE + tid] = rgb[3 * id + ch];
            s_dep[tid] = depths[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK_SIZE, todo); j++) {
            contributor--;
            if ((int)contributor >= last_contributor) continue;
            const float dx = s_xy[j].x - (float)px, dy = s_xy[j].y - (float)py;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.f / 255.f) continue;
            T = T / (1.f - alpha);
            const float w = alpha * T;
            const int id = s_id[j];
            float dL_dalpha = 0.f;
            for (int ch = 0; ch < 3; ch++) {
                const float c = s_rgb[ch * BLOCK_SIZE + j];
                arec[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * arec[ch];
                last_c[ch] = c;
                dL_dalpha += (c - arec[ch]) * g3[ch];
                atomicAdd(&dL_dcolors[3 * id + ch], w * g3[ch]);                       // atomics 1-3
            }
            const float dpt = s_dep[j];
            drec = last_alpha * last_d + (1.f - last_alpha) * drec;
            last_d = dpt;
            dL_dalpha += (dpt - drec) * gd;
            atomicAdd(&dL_ddepths[id], w * gd);                                         // 4
            alrec = last_alpha + (1.f - last_alpha) * alrec;
            dL_dalpha += (1.f - alrec) * ga;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            atomicAdd(&dL_dmean2D[2 * id], dL_dG * (-gdx * co.x - gdy * co.y) * ddelx_dx);      // 5
            atomicAdd(&dL_dmean2D[2 * id + 1], dL_dG * (-gdy * co.z - gdx * co.y) * ddely_dy);  // 6
            atomicAdd(&dL_dconic[3 * id], -0.5f * gdx * dx * dL_dG);                    // 7
            atomicAdd(&dL_dconic[3 * id + 1], -gdx * dy * dL_dG);                       // 8 (true dL/dB)
            atomicAdd(&dL_dconic[3 * id + 2], -0.5f * gdy * dy * dL_dG);                // 9
            atomicAdd(&dL_dopacity[id], G * dL_dalpha);                                 // 10
        }
    }
}

// ---- A.10: one thread per Gaussian ----
__global__ void preprocess_backward_kernel(Params in, const int* radii, const float* cov3D, const bool* clamped,
                                           const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolors,
                                           const float* dL_ddepths, float* dmeans3D, float* dmeans2D, float* dsh,
                                           float* dscales, float* drots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.P || radii[i] <= 0) return;
    const float* p = in.means3D + 3 * i;
    const float *V = in.viewmatrix, *PM = in.projmatrix;
    float dmean[3] = {0, 0, 0};
    const float dm2x = dL_dmean2D[2 * i], dm2y = dL_dmean2D[2 * i + 1];
    dmeans2D[3 * i] = dm2x; dmeans2D[3 * i + 1] = dm2y;
    const float* c6 = cov3D + 6 * i;
    Proj pr;
    project_cov(in, p, c6, &pr);
    const float a = pr.a, b = pr.b, c = pr.c;
    const float den = a * c - b * b;
    const float k = 1.f / (den * den + 1e-7f);
    const float dA = dL_dconic[3 * i], dB = dL_dconic[3 * i + 1], dC = dL_dconic[3 * i + 2];
    const float dLa = k * (-c * c * dA + b * c * dB - b * b * dC);
    const float dLb = k * (2 * b * c * dA - (den + 2 * b * b) * dB + 2 * a * b * dC);
    const float dLc = k * (-b * b * dA + a * b * dB - a * a * dC);
    const float *M0 = pr.M0, *M1 = pr.M1;
    float d6[6];
    d6[0] = dLa * M0[0] * M0[0] + dLb * M0[0] * M1[0] + dLc * M1[0] * M1[0];
    d6[3] = dLa * M0[1] * M0[1] + dLb * M0[1] * M1[1] + dLc * M1[1] * M1[1];
    d6[5] = dLa * M0[2] * M0[2] + dLb * M0[2] * M1[2] + dLc * M1[2] * M1[2];
    d6[1] = 2 * dLa * M0[0] * M0[1] + dLb * (M0[0] * M1[1] + M0[1] * M1[0]) + 2 * dLc * M1[0] * M1[1];
    d6[2] = 2 * dLa * M0[0] * M0[2] + dLb * (M0[0] * M1[2] + M0[2] * M1[0]) + 2 * dLc * M1[0] * M1[2];
    d6[4] = 2 * dLa * M0[1] * M0[2] + dLb * (M0[1] * M1[2] + M0[2] * M1[1]) + 2 * dLc * M1[1] * M1[2];
    const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float SM0[3], SM1[3], dM0[3], dM1[3];
    for (int r = 0; r < 3; r++) {
        SM0[r] = S[r * 3] * M0[0] + S[r * 3 + 1] * M0[1] + S[r * 3 + 2] * M0[2];
        SM1[r] = S[r * 3] * M1[0] + S[r * 3 + 1] * M1[1] + S[r * 3 + 2] * M1[2];
    }
    for (int r = 0; r < 3; r++) { dM0[r] = 2 * dLa * SM0[r] + dLb * SM1[r]; dM1[r] = dLb * SM0[r] + 2 * dLc * SM1[r]; }
    const float dJ00 = dM0[0] * V[0] + dM0[1] * V[4] + dM0[2] * V[8];
    const float dJ02 = dM0[0] * V[2] + dM0[1] * V[6] + dM0[2] * V[10];
    const float dJ11 = dM1[0] * V[1] + dM1[1] * V[5] + dM1[2] * V[9];
    const float dJ12 = dM1[0] * V[2] + dM1[1] * V[6] + dM1[2] * V[10];
    const float tz = pr.t[2], tzi = 1.f / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
    const float dtx = pr.xin ? -pr.fx * tz2 * dJ02 : 0.f;
    const float dty = pr.yin ? -pr.fy * tz2 * dJ12 : 0.f;
    const float dtz = -pr.fx * tz2 * dJ00 - pr.fy * tz2 * dJ11 + 2 * pr.fx * pr.tx * tz3 * dJ02 + 2 * pr.fy * pr.ty * tz3 * dJ12;
    for (int kk = 0; kk < 3; kk++) dmean[kk] += V[kk * 4 + 0] * dtx + V[kk * 4 + 1] * dty + V[kk * 4 + 2] * dtz;
    float hom[4];
    for (int j = 0; j < 4; j++) hom[j] = p[0] * PM[j] + p[1] * PM[4 + j] + p[2] * PM[8 + j] + PM[12 + j];
    const float pw = 1.f / (hom[3] + 1e-7f);
    for (int kk = 0; kk < 3; kk++)
        dmean[kk] += dm2x * (PM[kk * 4 + 0] * pw - hom[0] * pw * pw * PM[kk * 4 + 3]) +
                     dm2y * (PM[kk * 4 + 1] * pw - hom[1] * pw * pw * PM[kk * 4 + 3]);
    for (int kk = 0; kk < 3; kk++) dmean[kk] += V[kk * 4 + 2] * dL_ddepths[i];
    {
        const float vx = p[0] - in.campos[0], vy = p[1] - in.campos[1], vz = p[2] - in.campos[2];
        const float n = sqrtf(vx * vx + vy * vy + vz * vz);
        const float dir[3] = {vx / n, vy / n, vz / n};
        float bb[16], db[16][3];
        sh_basis(in.D, dir[0], dir[1], dir[2], bb, db);
        const int nk = (in.D + 1) * (in.D + 1);
        float ddir[3] = {0, 0, 0};
        for (int ch = 0; ch < 3; ch++) {
            const float gch = clamped[3 * i + ch] ? 0.f : dL_dcolors[3 * i + ch];
            for (int kq = 0; kq < nk; kq++) {
                const size_t idx = ((size_t)i * in.M + kq) * 3 + ch;
                dsh[idx] = bb[kq] * gch;
                for (int ax = 0; ax < 3; ax++) ddir[ax] += gch * db[kq][ax] * in.shs[idx];
            }
        }
        const float dot = dir[0] * ddir[0] + dir[1] * ddir[1] + dir[2] * ddir[2];
        for (int ax = 0; ax < 3; ax++) dmean[ax] += (ddir[ax] - dir[ax] * dot) / n;
    }
    {
        const float *sc = in.scales + 3 * i, *q = in.rotations + 4 * i;
        float R[9];
        quat_R(q, R);
        const float sp[3] = {in.scale_modifier * sc[0], in.scale_modifier * sc[1], in.scale_modifier * sc[2]};
        const float Sg[9] = {2 * d6[0], d6[1], d6[2], d6[1], 2 * d6[3], d6[4], d6[2], d6[4], 2 * d6[5]};
        float dLm[9], L[9], Dm[9];
        for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++) L[r * 3 + j] = R[r * 3 + j] * sp[j];
        for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++)
            dLm[r * 3 + j] = Sg[r * 3] * L[j] + Sg[r * 3 + 1] * L[3 + j] + Sg[r * 3 + 2] * L[6 + j];
        for (int j = 0; j < 3; j++) {
            const float ds = dLm[j] * R[j] + dLm[3 + j] * R[3 + j] + dLm[6 + j] * R[6 + j];
            dscales[3 * i + j] = in.scale_modifier * ds;
            for (int r = 0; r < 3; r++) Dm[r * 3 + j] = dLm[r * 3 + j] * sp[j];
        }
        const float r = q[0], x = q[1], y = q[2], z = q[3];
        drots[4 * i + 0] = 2 * (-z * Dm[1] + y * Dm[2] + z * Dm[3] - x * Dm[5] - y * Dm[6] + x * Dm[7]);
        drots[4 * i + 1] = 2 * (y * Dm[1] + z * Dm[2] + y * Dm[3] - 2 * x * Dm[4] - r * Dm[5] + z * Dm[6] + r * Dm[7] - 2 * x * Dm[8]);
        drots[4 * i + 2] = 2 * (-2 * y * Dm[0] + x * Dm[1] + r * Dm[2] + x * Dm[3] + z * Dm[5] - r * Dm[6] + z * Dm[7] - 2 * y * Dm[8]);
        drots[4 * i + 3] = 2 * (-2 * z * Dm[0] - r * Dm[1] + x * Dm[2] + r * Dm[3] - 2 * z * Dm[4] + y * Dm[5] + x * Dm[6] + y * Dm[7]);
    }
    for (int kk = 0; kk < 3; kk++) dmeans3D[3 * i + kk] = dmean[kk];
}

}  // namespace standin

using namespace standin;

extern "C" {

// Forward.  All pointers are device fp32 / int32 except num_rendered (host).  Blocks the host once (D2H of R), like
// generic implementations do.  Returns 0 on success.
int standin_forward(int P, int M, int D, int W, int H, float tanfovx, float tanfovy, float scale_modifier, const float* bg,
                    const float* viewmatrix, const float* projmatrix, const float* campos, const float* means3D,
                    const float* opacities, const float* shs, const float* scales, const float* rotations,
                    float* out_color, float* out_depth, float* out_alpha, int* radii, long long* num_rendered,
                    void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    g.p = Params{P, M, D, W, H, tanfovx, tanfovy, scale_modifier, bg, viewmatrix, projmatrix, campos, means3D, opacities, shs, scales, rotations};
    g.gx = (W + BLK - 1) / BLK; g.gy = (H + BLK - 1) / BLK;
    const size_t N = (size_t)W * H, Tn = (size_t)g.gx * g.gy;
    g.depths = take<float>(0, P); g.xy = take<float2>(1, P); g.conic_opacity = take<float4>(2, P); g.rgb = take<float>(3, 3 * (size_t)P);
    g.cov3D = take<float>(4, 6 * (size_t)P); g.clamped = take<bool>(5, 3 * (size_t)P); g.tiles_touched = take<uint32_t>(6, P);
    g.offsets = take<uint32_t>(7, P); g.ranges = take<uint2>(8, Tn); g.n_contrib = take<uint32_t>(9, N); g.final_T = take<float>(10, N);
    g.radii_i = radii;
    cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), st);       // generic implementations hand back zero-filled outputs
    cudaMemsetAsync(out_depth, 0, N * sizeof(float), st);
    cudaMemsetAsync(out_alpha, 0, N * sizeof(float), st);
    if (P == 0) { *num_rendered = 0; g.R = 0; return 0; }
    preprocess_kernel<<<(P + 255) / 256, 256, 0, st>>>(g.p, g.gx, g.gy, g.depths, g.xy, g.conic_opacity, g.rgb, g.cov3D, radii,
                                                      g.clamped, g.tiles_touched);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, g.tiles_touched, g.offsets, P, st);
    void* scan_tmp = take<char>(11, scan_bytes);
    cub::DeviceScan::InclusiveSum(scan_tmp, scan_bytes, g.tiles_touched, g.offsets, P, st);
    uint32_t R32 = 0;
    cudaMemcpyAsync(&R32, g.offsets + P - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;        // the blocking read of num_rendered
    const int64_t R = R32;
    g.R = R;
    *num_rendered = R;
    g.keys = take<uint64_t>(12, R); g.keys_sorted = take<uint64_t>(13, R); g.vals = take<uint32_t>(14, R); g.vals_sorted = take<uint32_t>(15, R);
    cudaMemsetAsync(g.ranges, 0, Tn * sizeof(uint2), st);
    if (R > 0) {
        duplicate_with_keys<<<(P + 255) / 256, 256, 0, st>>>(P, g.xy, g.depths, g.offsets, radii, g.gx, g.gy, g.keys, g.vals);
        int bit = 0;
        while ((1u << bit) <= (uint32_t)Tn) bit++;                   // msb of the tile count
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, g.keys, g.keys_sorted, g.vals, g.vals_sorted, (int)R, 0, 32 + bit, st);
        void* sort_tmp = take<char>(16, sort_bytes);
        cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, g.keys, g.keys_sorted, g.vals, g.vals_sorted, (int)R, 0, 32 + bit, st);
        identify_tile_ranges<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(R, g.keys_sorted, g.ranges);
    }
    render_forward_kernel<<<dim3(g.gx, g.gy), dim3(BLK, BLK), 0, st>>>(g.ranges, g.vals_sorted, W, H, g.xy, g.rgb, g.depths,
                                                                      g.conic_opacity, bg, g.n_contrib, g.final_T, out_color,
                                                                      out_depth, out_alpha);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// Backward of the last standin_forward.  Gradient arrays are zero-filled here (generic implementations allocate them
// with zeros).
int standin_backward(const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, float* dmeans3D, float* dmeans2D,
                     float* dsh, float* dopac, float* dscales, float* drots, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    const int P = g.p.P, W = g.p.W, H = g.p.H;
    if (P == 0) return 0;
    g.dL_dmean2D = take<float>(17, 2 * (size_t)P); g.dL_dconic = take<float>(18, 3 * (size_t)P);
    g.dL_dcolors = take<float>(19, 3 * (size_t)P); g.dL_ddepths = take<float>(20, P);
    cudaMemsetAsync(g.dL_dmean2D, 0, 2 * (size_t)P * 4, st); cudaMemsetAsync(g.dL_dconic, 0, 3 * (size_t)P * 4, st);
    cudaMemsetAsync(g.dL_dcolors, 0, 3 * (size_t)P * 4, st); cudaMemsetAsync(g.dL_ddepths, 0, (size_t)P * 4, st);
    cudaMemsetAsync(dmeans3D, 0, 3 * (size_t)P * 4, st); cudaMemsetAsync(dmeans2D, 0, 3 * (size_t)P * 4, st);
    cudaMemsetAsync(dsh, 0, 3 * (size_t)g.p.M * P * 4, st); cudaMemsetAsync(dopac, 0, (size_t)P * 4, st);
    cudaMemsetAsync(dscales, 0, 3 * (size_t)P * 4, st); cudaMemsetAsync(drots, 0, 4 * (size_t)P * 4, st);
    render_backward_kernel<<<dim3(g.gx, g.gy), dim3(BLK, BLK), 0, st>>>(g.ranges, g.vals_sorted, W, H, g.p.bg, g.xy, g.conic_opacity,
                                                                       g.rgb, g.depths, g.final_T, g.n_contrib, dL_dcolor, dL_ddepth,
                                                                       dL_dalpha, g.dL_dmean2D, g.dL_dconic, dopac, g.dL_dcolors,
                                                                       g.dL_ddepths);
    preprocess_backward_kernel<<<(P + 255) / 256, 256, 0, st>>>(g.p, g.radii_i, g.cov3D, g.clamped, g.dL_dmean2D, g.dL_dconic,
                                                               g.dL_dcolors, g.dL_ddepths, dmeans3D, dmeans2D, dsh, dscales, drots);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

void standin_release(void) {
    for (int i = 0; i < 24; i++) { if (g_buf.ptr[i]) cudaFree(g_buf.ptr[i]); g_buf.ptr[i] = nullptr; g_buf.cap[i] = 0; }
}

}  // extern "C"
