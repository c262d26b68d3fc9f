// Dense optical flow on the device (stages K0-K5 of SURVEY.md section 2.2): what the reference gets
// from cv::cvtColor(BGR2GRAY) and cv::calcOpticalFlowFarneback(prev, next, flow, 0.5, 3, 15, 3, 5,
// 1.2, 0) at cpp/src/segment.cpp:97-101 (and :222-226 in the video loop).  OpenCV is not part of the
// reference tree and not version-pinned by it (cpp/CMakeLists.txt:14); the arithmetic below restates
// the published algorithm of OpenCV 4.x modules/video/src/optflowgf.cpp (polynomial expansion,
// update-matrices, box-filtered 2x2 solve, coarse-to-fine over a Gaussian pyramid) and is checked
// against cv2 4.13 with an end-point-error tolerance (tests/test_gpu_parity.py, tests/test_gpu_fullsize.py).
//
// Layout in HBM (all row-major, one image after the other):
//   gray   u8  [image][H][W]
//   I      f32 [image][Hk][Wk]            pyramid level k of every image
//   R      f32 [image][Hk][Wk][5]         polynomial expansion coefficients (OpenCV's channel order)
//   M      f32 [pair][Hk][Wk][5]          G11, G12, G22, h1, h2 products
//   flow   f32 [pair][Hk][Wk][2]
// Every kernel takes the image / pair index from blockIdx.z (or .y), so one launch covers the batch.
#pragma once
#include <math.h>

#include "dofs_common.cuh"

// ---------------------------------------------------------------------------------------------
// K0  cv::cvtColor(BGR2GRAY) on 8-bit data: (R*9798 + G*19235 + B*3735 + 2^14) >> 15
// ---------------------------------------------------------------------------------------------
DOFS_D u32 gray_of(u32 b, u32 g, u32 r) { return (r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15; }

// four pixels per thread: three aligned 32-bit loads, one 32-bit store
__global__ void __launch_bounds__(256)
k_bgr2gray(const u8* __restrict__ bgr, u8* __restrict__ gray, size_t n_px) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = q * 4;
    if (p >= n_px) return;
    if (p + 4 <= n_px) {
        const u32* in = reinterpret_cast<const u32*>(bgr) + q * 3;
        const u32 w0 = in[0], w1 = in[1], w2 = in[2];
        const u32 g0 = gray_of(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        const u32 g1 = gray_of(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        const u32 g2 = gray_of((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        const u32 g3 = gray_of((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        reinterpret_cast<u32*>(gray)[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (size_t i = p; i < n_px; ++i) gray[i] = (u8)gray_of(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
    }
}

// ---------------------------------------------------------------------------------------------
// host-side configuration
// ---------------------------------------------------------------------------------------------
#define FLOW_MAX_LEVELS 8
#define FLOW_MAX_SMOOTH_RADIUS 24
#define FLOW_MAX_POLY_N 7

struct FlowConfig {
    double pyr_scale = 0.5;
    int levels = 3;
    int winsize = 15;
    int iters = 3;
    int poly_n = 5;
    double poly_sigma = 1.2;
};

struct SmoothTaps {  // Gaussian of one pyramid level, cv::getGaussianKernel(smooth_sz, sigma, CV_32F)
    int radius;
    float k[2 * FLOW_MAX_SMOOTH_RADIUS + 1];
};

struct PolyCoef {  // FarnebackPrepareGaussian
    int n;
    float g[FLOW_MAX_POLY_N + 1], xg[FLOW_MAX_POLY_N + 1], xxg[FLOW_MAX_POLY_N + 1];
    double ig11, ig03, ig33, ig55;
};

struct FlowLevel {
    int w, h;
    double scale;  // pyr_scale^k
    SmoothTaps taps;
};

struct FlowBuffers {
    int W = 0, H = 0, F = 0;
    FlowConfig cfg;
    int n_levels = 0;  // number of scales (coarsest index = n_levels - 1)
    FlowLevel level[FLOW_MAX_LEVELS];
    PolyCoef poly;
    float* I = nullptr;      // [2F][N]
    float* R = nullptr;      // [2F][N][5]
    float* M = nullptr;      // [F][N][5]
    float* M2 = nullptr;     // [F][N][5] second M buffer of the fused box-filter + update-matrices iterations
    bool fuse_um = false;    // A/B knob (DOFS3D_FLOW_FUSE=1)
    bool bs7_float = true;   // A/B knob (DOFS3D_BS7_FLOAT=0: double window sums in shared memory, k_box_solve7)
    float2* flowA = nullptr; // [F][N]
    float2* flowB = nullptr; // [F][N]
    bool pyr_generic = false;            // A/B knob (DOFS3D_PYR_GENERIC=1): the run-time-radius pyramid kernel for every level
    bool pyr_untiled = true;             // A/B knob (DOFS3D_PYR_TILED=1 selects k_pyr_level_tiled): measured on the B200, the
                                         // tiled kernel is 29 % faster alone (pyramid 1.34 -> 0.95 ms per 32 pairs) and costs
                                         // 10 % of the throughput when six contexts share the GPU (964 -> 862 pairs/s): its
                                         // 40 KB of shared memory per block keeps the other contexts' kernels off the SM while a
                                         // latency-bound kernel runs; the per-thread kernel shares the SM with them
    float* R_carry = nullptr;            // polynomial expansion of ONE frame at every level (streaming: the last frame of a chunk)
    size_t carry_off[FLOW_MAX_LEVELS];   // float offset of level k in R_carry
};

struct FlowLaunchStats {
    long long launches = 0;
    void (*mark)(void* user, const char* name) = nullptr;  // optional: called after the launches of every stage (timing marks)
    void* user = nullptr;
};
#define FLOW_MARK(st, name)                              \
    do {                                                  \
        if ((st)->mark) (st)->mark((st)->user, name);     \
    } while (0)

inline int flow_cv_round(double v) { return (int)lrint(v); }  // round half to even, like cvRound

inline void flow_gaussian_kernel(int ksize, double sigma, float* out) {
    if (sigma <= 0 && ksize == 3) {  // OpenCV's fixed small kernel
        out[0] = 0.25f;
        out[1] = 0.5f;
        out[2] = 0.25f;
        return;
    }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        out[i] = (float)exp(scale2x * x * x);
        sum += out[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < ksize; ++i) out[i] = (float)(out[i] * sum);
}

// FarnebackPrepareGaussian: separable basis taps and the four entries of inv(G) that are used.
// G is 6x6 with the sparsity  [a . . b b .; . b . . . .; . . b . . .; b . . c d .; b . . d c .; . . . . . d]
// so the needed entries of its inverse have closed forms (G is symmetric positive definite).
inline void flow_poly_coef(int n, double sigma, PolyCoef* pc) {
    if (sigma < 1.1920929e-07) sigma = n * 0.3;
    pc->n = n;
    float g[2 * FLOW_MAX_POLY_N + 1];
    double s = 0;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = (float)exp(-x * x / (2 * sigma * sigma));
        s += g[x + n];
    }
    s = 1.0 / s;
    for (int x = -n; x <= n; ++x) g[x + n] = (float)(g[x + n] * s);
    for (int x = 0; x <= n; ++x) {
        pc->g[x] = g[x + n];
        pc->xg[x] = (float)(x * g[x + n]);
        pc->xxg[x] = (float)(x * x * g[x + n]);
    }
    double a = 0, b = 0, c = 0, d = 0;
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            // OpenCV forms these products in float (float * float * int ...) and accumulates in double
            const volatile float gg = g[y + n] * g[x + n];
            const volatile float gx2 = (float)(gg * (float)x) * (float)x;
            const volatile float gx4 = (float)((float)(gx2 * (float)x) * (float)x);
            const volatile float gx2y2 = (float)((float)(gx2 * (float)y) * (float)y);
            a += gg;
            b += gx2;
            c += gx4;
            d += gx2y2;
        }
    // rows/cols {0,3,4}: [a b b; b c d; b d c]; {1},{2}: b; {5}: d
    const double det3 = a * (c * c - d * d) - b * (b * c - b * d) + b * (b * d - b * c);
    pc->ig11 = 1.0 / b;
    pc->ig03 = -(b * c - b * d) / det3;  // cofactor(3,0)/det, symmetric
    pc->ig33 = (a * c - b * b) / det3;
    pc->ig55 = 1.0 / d;
}

inline const char* farneback_error(int rc) {
    return rc == 1 ? "unsupported Farneback parameters" : rc == 2 ? "device allocation failed (flow buffers)"
                   : rc == 3 ? "cudaFuncSetAttribute failed (box filter shared memory)" : "flow error";
}

// Dynamic shared-memory opt-in of the box-filter kernels (above 48 KB).  Function attributes are per device, so every
// context sets them after cudaSetDevice, when it allocates its flow buffers; the generic kernel gets the size of the
// largest window the configuration check admits, so contexts with different windows can share a device.
inline int farneback_set_attributes();

inline int farneback_alloc(FlowBuffers* fb, int W, int H, int F, const FlowConfig& cfg, size_t* bytes) {
    fb->W = W;
    fb->H = H;
    fb->F = F;
    fb->cfg = cfg;
    if (cfg.poly_n < 1 || cfg.poly_n > FLOW_MAX_POLY_N || cfg.winsize < 1 || cfg.winsize > 63 || cfg.iters < 1 ||
        cfg.levels < 1 || cfg.levels >= FLOW_MAX_LEVELS || !(cfg.pyr_scale > 0 && cfg.pyr_scale < 1))
        return 1;
    // number of scales: OpenCV stops when the next level would be smaller than 32 px
    int k = 0;
    double scale = 1;
    for (; k < cfg.levels; ++k) {
        scale *= cfg.pyr_scale;
        if (W * scale < 32 || H * scale < 32) break;
    }
    fb->n_levels = k + 1;
    for (int l = 0; l <= k; ++l) {
        double sc = 1;
        for (int i = 0; i < l; ++i) sc *= cfg.pyr_scale;
        FlowLevel& L = fb->level[l];
        L.scale = sc;
        const double sigma = (1.0 / sc - 1) * 0.5;
        int smooth = flow_cv_round(sigma * 5) | 1;
        smooth = smooth < 3 ? 3 : smooth;
        if (smooth / 2 > FLOW_MAX_SMOOTH_RADIUS) return 1;
        L.taps.radius = smooth / 2;
        flow_gaussian_kernel(smooth, sigma, L.taps.k);
        L.w = flow_cv_round(W * sc);
        L.h = flow_cv_round(H * sc);
        if (L.w < 2 || L.h < 2) return 1;
    }
    flow_poly_coef(cfg.poly_n, cfg.poly_sigma, &fb->poly);
    const size_t N = (size_t)W * H;
    size_t total = 0;
    auto alloc = [&](void** p, size_t b) {
        total += b;
        return cudaMalloc(p, b) == cudaSuccess;
    };
    if (!alloc((void**)&fb->I, 2 * F * N * sizeof(float)) || !alloc((void**)&fb->R, 2 * F * N * 5 * sizeof(float)) ||
        !alloc((void**)&fb->M, F * N * 5 * sizeof(float)) ||
        (fb->fuse_um && !alloc((void**)&fb->M2, F * N * 5 * sizeof(float))) || !alloc((void**)&fb->flowA, F * N * sizeof(float2)) ||
        !alloc((void**)&fb->flowB, F * N * sizeof(float2)))
        return 2;
    size_t carry = 0;
    for (int l = 0; l < fb->n_levels; ++l) {
        fb->carry_off[l] = carry;
        carry += (size_t)fb->level[l].w * fb->level[l].h * 5;
    }
    if (!alloc((void**)&fb->R_carry, carry * sizeof(float))) return 2;
    if (bytes) *bytes = total;
    return farneback_set_attributes();
}

inline void farneback_free(FlowBuffers* fb) {
    cudaFree(fb->I);
    cudaFree(fb->R);
    cudaFree(fb->M);
    cudaFree(fb->M2);
    fb->M2 = nullptr;
    cudaFree(fb->flowA);
    cudaFree(fb->flowB);
    cudaFree(fb->R_carry);
    fb->R_carry = nullptr;
    fb->I = fb->R = fb->M = nullptr;
    fb->flowA = fb->flowB = nullptr;
}

// ---------------------------------------------------------------------------------------------
// K1  pyramid level: convertTo(CV_32F) -> GaussianBlur(full resolution, smooth_sz, sigma,
// BORDER_REFLECT_101) -> resize(level size, INTER_LINEAR), fused: one thread per level pixel evaluates
// the (at most four) blurred full-resolution samples its bilinear footprint needs, rows first and then
// columns like OpenCV's separable filter, and interpolates.  The blurred full-resolution image never
// exists in HBM.
// ---------------------------------------------------------------------------------------------
struct ImageSet {  // image s of the batch lives at (s < split ? base0 + s*N : base1 + (s-split)*N)
    const u8* base0;
    const u8* base1;
    int split;
    int count;
};

DOFS_D int flow_reflect101(int i, int n) {
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// cv::resize INTER_LINEAR source coordinate: s = (d + 0.5) * (src/dst) - 0.5, clamped like OpenCV
DOFS_D void flow_linear_coord(int d, double scale, int src_n, int* i0, float* frac) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    f -= (float)i;
    if (i < 0) {
        i = 0;
        f = 0.f;
    }
    if (i >= src_n - 1) {
        i = src_n - 1;
        f = 0.f;
    }
    *i0 = i;
    *frac = f;
}

// RT = the filter radius at compile time (1, 4, 9: the three coarser levels of pyr_scale 0.5), 0 = run time.  With RT
// the row filter is unrolled: per tap a byte load, a conversion and two fma with the tap as a constant-bank operand,
// instead of a counted loop that also fetches the tap (the kernel is bound by instruction issue: 19 x 20 taps per
// output at the coarsest level).  Same fma chains in the same order: bit-identical.
template <int RT>
__global__ void __launch_bounds__(256)
k_pyr_level(ImageSet imgs, float* __restrict__ I, int W, int H, int Wk, int Hk, SmoothTaps taps) {
    const int img = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wk || y >= Hk) return;
    const size_t N = (size_t)W * H;
    const u8* src = img < imgs.split ? imgs.base0 + (size_t)img * N : imgs.base1 + (size_t)(img - imgs.split) * N;
    const int r = RT ? RT : taps.radius;
    int sx, sy;
    float fx, fy;
    if (Wk == W && Hk == H) {  // resize to the same size is a copy
        sx = x;
        sy = y;
        fx = fy = 0.f;
    } else {
        flow_linear_coord(x, (double)W / Wk, W, &sx, &fx);
        flow_linear_coord(y, (double)H / Hk, H, &sy, &fy);
    }
    const int nx = fx != 0.f ? 2 : 1, ny = fy != 0.f ? 2 : 1;
    // column (vertical) filter of the row-filtered samples, for the nx x ny blurred samples
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [dy][dx]
    const bool interior = sx - r >= 0 && sx + r + 1 < W;
#pragma unroll 1
    for (int j = -r; j <= r + ny - 1; ++j) {
        const int yy = flow_reflect101(sy + j, H);
        const u8* row = src + yy * W;  // (an image has fewer than 2^31 pixels)
        float h0 = 0.f, h1 = 0.f;  // row filter at columns sx and sx+1
        if (interior) {
            const u8* p = row + (sx - r);
            if (RT) {
                float v[2 * (RT ? RT : 1) + 2];
#pragma unroll
                for (int i = 0; i < 2 * RT + 2; ++i) v[i] = (float)p[i];
#pragma unroll
                for (int i = 0; i < 2 * RT + 1; ++i) {
                    h0 = fmaf(taps.k[i], v[i], h0);
                    h1 = fmaf(taps.k[i], v[i + 1], h1);
                }
            } else {
                float prev = (float)p[0];
                for (int i = -r; i <= r; ++i) {
                    const float nxt = (float)p[i + r + 1];
                    h0 = fmaf(taps.k[i + r], prev, h0);
                    h1 = fmaf(taps.k[i + r], nxt, h1);
                    prev = nxt;
                }
            }
        } else {
            for (int i = -r; i <= r; ++i) {
                h0 = fmaf(taps.k[i + r], (float)row[flow_reflect101(sx + i, W)], h0);
                if (nx == 2) h1 = fmaf(taps.k[i + r], (float)row[flow_reflect101(sx + 1 + i, W)], h1);
            }
        }
        if (j <= r) {
            acc[0][0] = fmaf(taps.k[j + r], h0, acc[0][0]);
            acc[0][1] = fmaf(taps.k[j + r], h1, acc[0][1]);
        }
        if (ny == 2 && j >= -r + 1) {
            acc[1][0] = fmaf(taps.k[j - 1 + r], h0, acc[1][0]);
            acc[1][1] = fmaf(taps.k[j - 1 + r], h1, acc[1][1]);
        }
    }
    // bilinear: horizontal, then vertical
    const float top = nx == 2 ? fmaf(acc[0][1], fx, acc[0][0] * (1.f - fx)) : acc[0][0];
    float v = top;
    if (ny == 2) {
        const float bot = nx == 2 ? fmaf(acc[1][1], fx, acc[1][0] * (1.f - fx)) : acc[1][0];
        v = fmaf(bot, fy, top * (1.f - fy));
    }
    I[((size_t)img * Hk + y) * Wk + x] = v;
}

// The same level from a shared-memory tile: the block's source window is staged once as bytes (whole 32-bit words with
// several loads in flight per thread when the window lies inside the image, reflected byte by byte otherwise),
// row-filtered at the two source columns every output needs, then column-filtered and interpolated.  The fmaf chains
// are those of k_pyr_level in the same order, so the result is bit-identical to it.  What changes: k_pyr_level is bound by
// latency — every tap is a dependent global byte load -> convert -> fma chain with a run-time trip count (18 x 18 of
// them per output at the coarsest level) — while here a source byte comes from HBM/L2 once per block.
// Dynamic shared memory: row-filtered values [TH][64] floats, then the byte tile [TH][TWB].
__global__ void __launch_bounds__(256)
k_pyr_level_tiled(ImageSet imgs, float* __restrict__ I, int W, int H, int Wk, int Hk, SmoothTaps taps, int TWB, int TH) {
    extern __shared__ __align__(16) float pyr_smem[];
    float* s_h = pyr_smem;                                   // [TH][64]
    u8* s_t = reinterpret_cast<u8*>(pyr_smem + TH * 64);     // [TH][TWB], TWB a multiple of 4
    const int img = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int xb = blockIdx.x * 32, yb = blockIdx.y * 8;
    const size_t N = (size_t)W * H;
    const u8* src = img < imgs.split ? imgs.base0 + (size_t)img * N : imgs.base1 + (size_t)(img - imgs.split) * N;
    const int r = taps.radius;
    const double scx = (double)W / Wk, scy = (double)H / Hk;
    int sx, sy, t0;
    float fx, fy, tf;
    flow_linear_coord(min(xb + tx, Wk - 1), scx, W, &sx, &fx);
    flow_linear_coord(min(yb + ty, Hk - 1), scy, H, &sy, &fy);
    // window of the block: columns X0 .. X0+tw-1, rows Y0 .. Y0+th-1 (source coordinates, may leave the image)
    flow_linear_coord(xb, scx, W, &t0, &tf);
    const int X0 = t0 - r;
    flow_linear_coord(min(xb + 31, Wk - 1), scx, W, &t0, &tf);
    const int tw = t0 + 1 + r - X0 + 1;
    flow_linear_coord(yb, scy, H, &t0, &tf);
    const int Y0 = t0 - r;
    flow_linear_coord(min(yb + 7, Hk - 1), scy, H, &t0, &tf);
    const int th = t0 + 1 + r - Y0 + 1;
    int shift = 0;  // tile column of source column X0
    if (X0 >= 0 && Y0 >= 0 && X0 + tw <= W && Y0 + th <= H && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const int Xa = X0 & ~3;
        shift = X0 - Xa;
        const int nw = (shift + tw + 3) >> 2;  // words per tile row; Xa + 4 nw <= W because W is a multiple of 4
        const int total = th * nw;
        u32* s_w = reinterpret_cast<u32*>(s_t);
        for (int base = 0; base < total; base += 256 * 4) {
            u32 v[4];
            int at[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 256 + (int)threadIdx.x;
                const int ry = idx / nw, c = idx - ry * nw;
                at[u] = idx < total ? ry * (TWB >> 2) + c : -1;
                v[u] = idx < total ? *reinterpret_cast<const u32*>(src + (size_t)(Y0 + ry) * W + Xa + 4 * c) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (at[u] >= 0) s_w[at[u]] = v[u];
        }
    } else {
        for (int ry = ty; ry < th; ry += 8) {
            const u8* row = src + (size_t)flow_reflect101(Y0 + ry, H) * W;
            for (int cx = tx; cx < tw; cx += 32) s_t[ry * TWB + cx] = row[flow_reflect101(X0 + cx, W)];
        }
    }
    __syncthreads();
    // row filter at columns sx and sx+1 of every output column of the block, for every row of the window
    {
        const int lx = sx - X0 + shift;  // tile column of the centre tap
        for (int ry = ty; ry < th; ry += 8) {
            const u8* t = s_t + ry * TWB + lx;
            float h0 = 0.f, h1 = 0.f;
            float prev = (float)t[-r];
            for (int i = -r; i <= r; ++i) {
                const float nxt = (float)t[i + 1];
                h0 = fmaf(taps.k[i + r], prev, h0);
                h1 = fmaf(taps.k[i + r], nxt, h1);
                prev = nxt;
            }
            s_h[ry * 64 + 2 * tx] = h0;
            s_h[ry * 64 + 2 * tx + 1] = h1;
        }
    }
    __syncthreads();
    const int x = xb + tx, y = yb + ty;
    if (x >= Wk || y >= Hk) return;
    const int nx = fx != 0.f ? 2 : 1, ny = fy != 0.f ? 2 : 1;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [dy][dx]
    const float* hcol = s_h + (sy - Y0) * 64 + 2 * tx;
    for (int j = -r; j <= r + ny - 1; ++j) {
        const float h0 = hcol[j * 64], h1 = hcol[j * 64 + 1];
        if (j <= r) {
            acc[0][0] = fmaf(taps.k[j + r], h0, acc[0][0]);
            acc[0][1] = fmaf(taps.k[j + r], h1, acc[0][1]);
        }
        if (ny == 2 && j >= -r + 1) {
            acc[1][0] = fmaf(taps.k[j - 1 + r], h0, acc[1][0]);
            acc[1][1] = fmaf(taps.k[j - 1 + r], h1, acc[1][1]);
        }
    }
    const float top = nx == 2 ? fmaf(acc[0][1], fx, acc[0][0] * (1.f - fx)) : acc[0][0];
    float v = top;
    if (ny == 2) {
        const float bot = nx == 2 ? fmaf(acc[1][1], fx, acc[1][0] * (1.f - fx)) : acc[1][0];
        v = fmaf(bot, fy, top * (1.f - fy));
    }
    I[((size_t)img * Hk + y) * Wk + x] = v;
}
#define PYR_TILED_MAX_SMEM (48 * 1024)  // no opt-in: the level falls back to k_pyr_level when its window needs more
// window bound of a block of 32 x 8 outputs: the source step of 31 (7) outputs, the two interpolation taps, the filter
// radius on both sides, and slack for the rounding of the coordinate map
inline void pyr_tile_dims(int W, int H, int Wk, int Hk, int r, int* TWB, int* TH) {
    const int tw = (int)ceil(31.0 * W / Wk) + 2 * r + 4;
    *TWB = (tw + 3 + 3) & ~3;  // + the alignment shift of the word loads, rounded to whole words
    *TH = (int)ceil(7.0 * H / Hk) + 2 * r + 4;
}

// Level 0 of the pyramid (same size as the frame, 3 taps): the general kernel's arithmetic, without its loops
// (same fmaf chains in the same order, so the result is bit-identical to k_pyr_level).
__global__ void __launch_bounds__(256)
k_pyr_level0(ImageSet imgs, float* __restrict__ I, int W, int H, SmoothTaps taps) {
    const int img = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t N = (size_t)W * H;
    const u8* src = img < imgs.split ? imgs.base0 + (size_t)img * N : imgs.base1 + (size_t)(img - imgs.split) * N;
    const int xm = flow_reflect101(x - 1, W), xp = flow_reflect101(x + 1, W);
    const float k0 = taps.k[0], k1 = taps.k[1], k2 = taps.k[2];
    float acc = 0.f;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
        const u8* row = src + (size_t)flow_reflect101(y + j, H) * W;
        float h = fmaf(k0, (float)row[xm], 0.f);
        h = fmaf(k1, (float)row[x], h);
        h = fmaf(k2, (float)row[xp], h);
        acc = fmaf(j == -1 ? k0 : j == 0 ? k1 : k2, h, acc);
    }
    I[(size_t)img * N + (size_t)y * W + x] = acc;
}

// ---------------------------------------------------------------------------------------------
// K2  FarnebackPolyExp: 3 vertical moment sums in float (rows clamped), 6 horizontal sums in double
// (columns clamped), projected through inv(G).  Block = 32x8 pixels; the vertical pass of the tile
// plus a halo of n columns is staged in shared memory.
// ---------------------------------------------------------------------------------------------
#define PE_TW 32
#define PE_TH 8
#ifndef PE_STAGE_RAW
#define PE_STAGE_RAW 0  // 1: stage the tile's source window in shared memory first.  Measured: 1.54 ms per 32 pairs against 1.40
                        // without (the extra barrier and the second pass cost more than the address arithmetic they save)
#endif

// NT = poly_n at compile time (5 for the reference: loops unrolled, coefficients in registers), 0 = run time
template <int NT>
__global__ void __launch_bounds__(PE_TW * PE_TH)
k_polyexp(const float* __restrict__ I, float* __restrict__ R, int Wk, int Hk, PolyCoef pc) {
    __shared__ float s_row[3][PE_TH][PE_TW + 2 * FLOW_MAX_POLY_N];
#if PE_STAGE_RAW
    // the tile's source window (rows and columns clamped) is staged once: the vertical pass then reads shared memory at
    // compile-time offsets instead of 2 n + 1 clamped global addresses per value (address arithmetic was a quarter of
    // this kernel's instructions, and it is bound by instruction issue)
    __shared__ float s_raw[PE_TH + 2 * FLOW_MAX_POLY_N][PE_TW + 2 * FLOW_MAX_POLY_N + 1];
#endif
    const int img = blockIdx.z;
    const int n = NT ? NT : pc.n;
    const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
    const float* src = I + (size_t)img * Wk * Hk;
    const int span = PE_TW + 2 * n;
#if PE_STAGE_RAW
    for (int idx = threadIdx.x; idx < span * (PE_TH + 2 * n); idx += PE_TW * PE_TH) {
        const int ry = idx / span, cx = idx - ry * span;
        const int y = min(max(y0 + ry - n, 0), Hk - 1);
        const int xs = min(max(x0 + cx - n, 0), Wk - 1);
        s_raw[ry][cx] = src[y * Wk + xs];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < span * PE_TH; idx += PE_TW * PE_TH) {
        const int ty = idx / span, cx = idx - ty * span;
        const float* col = &s_raw[ty + n][cx];
        constexpr int RP = PE_TW + 2 * FLOW_MAX_POLY_N + 1;
        const float c = col[0];
        float t0 = xfmul(c, pc.g[0]), t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 1; k <= n; ++k) {
            const float a = col[-k * RP];
            const float b = col[k * RP];
            const float p = xfadd(a, b);
            t0 = xfadd(t0, xfmul(pc.g[k], p));
            t1 = xfadd(t1, xfmul(pc.xg[k], xfsub(b, a)));
            t2 = xfadd(t2, xfmul(pc.xxg[k], p));
        }
        s_row[0][ty][cx] = t0;
        s_row[1][ty][cx] = t1;
        s_row[2][ty][cx] = t2;
    }
#else
    for (int idx = threadIdx.x; idx < span * PE_TH; idx += PE_TW * PE_TH) {
        const int ty = idx / span, cx = idx - ty * span;
        const int y = min(y0 + ty, Hk - 1);
        const int xs = min(max(x0 + cx - n, 0), Wk - 1);
        const float c = src[y * Wk + xs];  // (32-bit offsets inside an image: a third fewer instructions per load)
        float t0 = xfmul(c, pc.g[0]), t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 1; k <= n; ++k) {
            const float a = src[max(y - k, 0) * Wk + xs];
            const float b = src[min(y + k, Hk - 1) * Wk + xs];
            const float p = xfadd(a, b);
            t0 = xfadd(t0, xfmul(pc.g[k], p));
            t1 = xfadd(t1, xfmul(pc.xg[k], xfsub(b, a)));
            t2 = xfadd(t2, xfmul(pc.xxg[k], p));
        }
        s_row[0][ty][cx] = t0;
        s_row[1][ty][cx] = t1;
        s_row[2][ty][cx] = t2;
    }
#endif
    __syncthreads();
    const int tx = threadIdx.x & (PE_TW - 1), ty = threadIdx.x / PE_TW;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= Wk || y >= Hk) return;
    const float* r0 = &s_row[0][ty][tx + n];
    const float* r1 = &s_row[1][ty][tx + n];
    const float* r2 = &s_row[2][ty][tx + n];
    double b1 = (double)xfmul(r0[0], pc.g[0]), b2 = 0, b3 = (double)xfmul(r1[0], pc.g[0]), b4 = 0;
    double b5 = (double)xfmul(r2[0], pc.g[0]), b6 = 0;
#pragma unroll
    for (int k = 1; k <= n; ++k) {
        const double tg = (double)xfadd(r0[k], r0[-k]);
        const float g0 = pc.g[k];
        b1 = xdadd(b1, xdmul(tg, (double)g0));
        b4 = xdadd(b4, xdmul(tg, (double)pc.xxg[k]));
        b2 = xdadd(b2, (double)xfmul(xfsub(r0[k], r0[-k]), pc.xg[k]));
        b3 = xdadd(b3, (double)xfmul(xfadd(r1[k], r1[-k]), g0));
        b6 = xdadd(b6, (double)xfmul(xfsub(r1[k], r1[-k]), pc.xg[k]));
        b5 = xdadd(b5, (double)xfmul(xfadd(r2[k], r2[-k]), g0));
    }
    float* out = R + (((size_t)img * Hk + y) * Wk + x) * 5;
    out[0] = (float)xdmul(b3, pc.ig11);
    out[1] = (float)xdmul(b2, pc.ig11);
    out[2] = (float)xdadd(xdmul(b1, pc.ig03), xdmul(b5, pc.ig33));
    out[3] = (float)xdadd(xdmul(b1, pc.ig03), xdmul(b4, pc.ig33));
    out[4] = (float)xdmul(b6, pc.ig55);
}

// ---------------------------------------------------------------------------------------------
// K3  flow of the next finer level: resize(prev_flow, size, INTER_LINEAR) * (1 / pyr_scale) — evaluated inside the
// first k_update_matrices of the level, never stored
// ---------------------------------------------------------------------------------------------
// one pixel of resize(prev_flow, (Wk, Hk), INTER_LINEAR) * mul
DOFS_D float2 flow_upsampled(const float2* __restrict__ src /* [Hp][Wp] of the pair */, int x, int y, int Wp, int Hp, int Wk,
                             int Hk, double mul, double scx, double scy) {
    int sx, sy;
    float fx, fy;
    flow_linear_coord(x, scx, Wp, &sx, &fx);
    flow_linear_coord(y, scy, Hp, &sy, &fy);
    const int sx1 = min(sx + 1, Wp - 1), sy1 = min(sy + 1, Hp - 1);
    const float2 a = src[(size_t)sy * Wp + sx], b = src[(size_t)sy * Wp + sx1];
    const float2 c = src[(size_t)sy1 * Wp + sx], d = src[(size_t)sy1 * Wp + sx1];
    const float wx0 = 1.f - fx, wy0 = 1.f - fy;
    const float tx = xfadd(xfmul(a.x, wx0), xfmul(b.x, fx)), ty = xfadd(xfmul(a.y, wx0), xfmul(b.y, fx));
    const float bx = xfadd(xfmul(c.x, wx0), xfmul(d.x, fx)), by = xfadd(xfmul(c.y, wx0), xfmul(d.y, fx));
    const float vx = xfadd(xfmul(tx, wy0), xfmul(bx, fy)), vy = xfadd(xfmul(ty, wy0), xfmul(by, fy));
    return make_float2((float)xdmul((double)vx, mul), (float)xdmul((double)vy, mul));
}

// ---------------------------------------------------------------------------------------------
// K4  FarnebackUpdateMatrices: warp R1 by the current flow (bilinear gather), average with R0, build
// the five products; attenuate within 5 px of the border.
// ---------------------------------------------------------------------------------------------
struct PairSlots {  // pair p uses polynomial expansions of images (p + first0) and (p + first1)
    int first0, first1;
};

DOFS_D float flow_border_scale(int x, int y, int w, int h) {
    const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    float s = 1.f;
    s = xfmul(s, x < 5 ? border[x] : 1.f);
    s = xfmul(s, x >= w - 5 ? border[w - x - 1] : 1.f);
    s = xfmul(s, y < 5 ? border[y] : 1.f);
    s = xfmul(s, y >= h - 5 ? border[h - y - 1] : 1.f);
    return s;
}

// The flow a level starts from — zero at the coarsest level, the upsampled flow of the coarser level otherwise — is only
// ever read here (the first box filter + solve overwrites it), so the first launch of a level forms it on the fly.
struct FlowStart {
    const float2* prev;  // flow of the coarser level, or nullptr: the level starts from zero
    int Wp, Hp;
    double mul;
    double scx, scy;  // (double)Wp / Wk, (double)Hp / Hk: divided once on the host, not twice per pixel
};
enum { UM_FLOW = 0, UM_START = 1 };

// one pixel of FarnebackUpdateMatrices: the five products for pixel (x, y) of `pair` given its current flow d
DOFS_D void update_matrices_pixel(const float* __restrict__ R, float* __restrict__ M, int Wk, int Hk, PairSlots ps, int pair,
                                  int x, int y, float2 d) {
    const size_t npx = (size_t)Wk * Hk;
    const float* R0 = R + ((size_t)(pair + ps.first0) * npx + (size_t)y * Wk + x) * 5;
    const float* R1 = R + (size_t)(pair + ps.first1) * npx * 5;
    const float dx = d.x, dy = d.y;
    float fx = xfadd((float)x, dx), fy = xfadd((float)y, dy);
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx = xfsub(fx, (float)x1);
    fy = xfsub(fy, (float)y1);
    const float q0 = R0[0], q1 = R0[1], q2 = R0[2], q3 = R0[3], q4 = R0[4];
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(Wk - 1) && (unsigned)y1 < (unsigned)(Hk - 1)) {
        const float a00 = xfmul(xfsub(1.f, fx), xfsub(1.f, fy)), a01 = xfmul(fx, xfsub(1.f, fy));
        const float a10 = xfmul(xfsub(1.f, fx), fy), a11 = xfmul(fx, fy);
        const float* p = R1 + ((size_t)y1 * Wk + x1) * 5;
        const float* q = p + (size_t)Wk * 5;
#define BIL(c) xfadd(xfadd(xfadd(xfmul(a00, p[c]), xfmul(a01, p[5 + c])), xfmul(a10, q[c])), xfmul(a11, q[5 + c]))
        r2 = BIL(0);
        r3 = BIL(1);
        r4 = BIL(2);
        r5 = BIL(3);
        r6 = BIL(4);
#undef BIL
        r4 = xfmul(xfadd(q2, r4), 0.5f);
        r5 = xfmul(xfadd(q3, r5), 0.5f);
        r6 = xfmul(xfadd(q4, r6), 0.25f);
    } else {
        r2 = r3 = 0.f;
        r4 = q2;
        r5 = q3;
        r6 = xfmul(q4, 0.5f);
    }
    r2 = xfmul(xfsub(q0, r2), 0.5f);
    r3 = xfmul(xfsub(q1, r3), 0.5f);
    r2 = xfadd(r2, xfadd(xfmul(r4, dy), xfmul(r6, dx)));
    r3 = xfadd(r3, xfadd(xfmul(r6, dy), xfmul(r5, dx)));
    if ((unsigned)(x - 5) >= (unsigned)(Wk - 10) || (unsigned)(y - 5) >= (unsigned)(Hk - 10)) {
        const float s = flow_border_scale(x, y, Wk, Hk);
        r2 = xfmul(r2, s);
        r3 = xfmul(r3, s);
        r4 = xfmul(r4, s);
        r5 = xfmul(r5, s);
        r6 = xfmul(r6, s);
    }
    float* out = M + ((size_t)pair * npx + (size_t)y * Wk + x) * 5;
    out[0] = xfadd(xfmul(r4, r4), xfmul(r6, r6));
    out[1] = xfmul(xfadd(r4, r5), r6);
    out[2] = xfadd(xfmul(r5, r5), xfmul(r6, r6));
    out[3] = xfadd(xfmul(r4, r2), xfmul(r6, r3));
    out[4] = xfadd(xfmul(r6, r2), xfmul(r5, r3));
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_update_matrices(const float* __restrict__ R, const float2* __restrict__ flow, float* __restrict__ M, int Wk, int Hk,
                  PairSlots ps, FlowStart fs) {
    const int pair = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wk || y >= Hk) return;
    const size_t npx = (size_t)Wk * Hk;
    float2 d;
    if (MODE == UM_FLOW) d = flow[(size_t)pair * npx + (size_t)y * Wk + x];
    else if (fs.prev) d = flow_upsampled(fs.prev + (size_t)pair * fs.Wp * fs.Hp, x, y, fs.Wp, fs.Hp, Wk, Hk, fs.mul, fs.scx, fs.scy);
    else d = make_float2(0.f, 0.f);
    update_matrices_pixel(R, M, Wk, Hk, ps, pair, x, y, d);
}

// ---------------------------------------------------------------------------------------------
// K5  FarnebackUpdateFlow_Blur: winsize x winsize box mean of M (replicated borders) and the 2x2
// solve, in double.  Block = 32x8 pixels; the M tile with its halo is staged in shared memory, the
// vertical window sums are formed once per column and reused by the horizontal window.
// ---------------------------------------------------------------------------------------------
#ifndef BS_COLS
#define BS_COLS 64   // output columns per block
#endif
#ifndef BS_ROWS
#define BS_ROWS 64   // rows per block (one marching segment)
#endif
#ifndef BS_BATCH
#define BS_BATCH 8   // rows of vertical sums staged per horizontal phase
#endif
#define BS_GROUP 8   // consecutive outputs one thread produces with a sliding horizontal window

// Each thread owns one (column, channel) of a strip of BS_COLS + 2m columns and marches down BS_ROWS
// rows keeping the vertical window sum in a register (first window summed directly, then + entering
// row - leaving row, in double): every M element is read straight from global memory / L1, coalesced
// across the strip, with all the loads of a batch of rows in flight together — no halo rows are
// re-staged.  Every BS_BATCH rows the vertical sums go through shared memory to the horizontal
// sliding windows and the solve.
__global__ void __launch_bounds__(640)
k_box_solve(const float* __restrict__ M, float2* __restrict__ flow, int Wk, int Hk, int m /* winsize/2 */) {
    extern __shared__ __align__(16) unsigned char bx_smem[];
    const int cols = BS_COLS + 2 * m;
    const int cw = cols * 5;
    double* s_v = reinterpret_cast<double*>(bx_smem);   // [BS_BATCH][cols][5] vertical window sums
    double* s_h = s_v + (size_t)BS_BATCH * cw;           // [BS_BATCH][BS_COLS][5] full window sums
    const int pair = blockIdx.z;
    const int x0 = blockIdx.x * BS_COLS;
    const int y_begin = blockIdx.y * BS_ROWS, y_end = min(y_begin + BS_ROWS, Hk);
    const float* src = M + (size_t)pair * Wk * Hk * 5;
    const size_t pitch = (size_t)Wk * 5;
    const int e = threadIdx.x;
    const bool owner = e < cw;
    const float* col = src;
    if (owner) {
        const int cx = e / 5, c = e - cx * 5;
        col = src + (size_t)min(max(x0 + cx - m, 0), Wk - 1) * 5 + c;  // replicated border columns
    }
    double s = 0;
    if (owner)
        for (int j = -m; j <= m; ++j) s += (double)col[(size_t)min(max(y_begin + j, 0), Hk - 1) * pitch];
    const int bs = 2 * m + 1;
    const double scale = 1.0 / (double)(bs * bs);
    for (int yb = y_begin; yb < y_end; yb += BS_BATCH) {
        const int nb = min(BS_BATCH, y_end - yb);
        if (owner) {
            float vin[BS_BATCH], vout[BS_BATCH];
#pragma unroll
            for (int r = 0; r < BS_BATCH; ++r) {
                const int y = yb + r;
                vin[r] = col[(size_t)min(y + m, Hk - 1) * pitch];
                vout[r] = col[(size_t)max(y - m - 1, 0) * pitch];
            }
#pragma unroll
            for (int r = 0; r < BS_BATCH; ++r) {
                if (r < nb) {
                    if (yb + r != y_begin) s += (double)vin[r] - (double)vout[r];
                    s_v[r * cw + e] = s;
                }
            }
        }
        __syncthreads();
        for (int w = threadIdx.x; w < nb * 5 * (BS_COLS / BS_GROUP); w += blockDim.x) {
            const int c = w % 5, g = (w / 5) % (BS_COLS / BS_GROUP), r = w / (5 * (BS_COLS / BS_GROUP));
            const double* v = s_v + r * cw + c;  // column cx of the strip at v[cx * 5]
            const int xs = g * BS_GROUP;
            double t = 0;
            for (int i = 0; i <= 2 * m; ++i) t += v[(xs + i) * 5];
            s_h[(r * BS_COLS + xs) * 5 + c] = t;
#pragma unroll
            for (int k = 1; k < BS_GROUP; ++k) {
                t += v[(xs + k + 2 * m) * 5] - v[(xs + k - 1) * 5];
                s_h[(r * BS_COLS + xs + k) * 5 + c] = t;
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < nb * BS_COLS; o += blockDim.x) {
            const int tx = o % BS_COLS, r = o / BS_COLS;
            const int x = x0 + tx, y = yb + r;
            if (x >= Wk) continue;
            const double* h = s_h + (size_t)o * 5;
            const double g11 = xdmul(h[0], scale), g12 = xdmul(h[1], scale), g22 = xdmul(h[2], scale);
            const double h1 = xdmul(h[3], scale), h2 = xdmul(h[4], scale);
            const double idet = xddiv(1.0, xdadd(xdsub(xdmul(g11, g22), xdmul(g12, g12)), 1e-3));
            const float u = (float)xdmul(xdsub(xdmul(g11, h2), xdmul(g12, h1)), idet);
            const float w = (float)xdmul(xdsub(xdmul(g22, h1), xdmul(g12, h2)), idet);
            flow[((size_t)pair * Hk + y) * Wk + x] = make_float2(u, w);
        }
        __syncthreads();
    }
}

// The same kernel specialised for the reference's window (winsize 15 -> m = 7), with shared-memory layouts that keep
// its phases free of bank conflicts (the kernel is bound by the L1/shared-memory pipeline, not by HBM): vertical sums
// as [row][channel][column + column/8] (channel pitch 99 doubles spreads the five channels of a column over distinct
// banks, row pitch 504 = 8 mod 16), window sums as [row][channel][x + x/8] (pitches 72 and 360).  A half-warp of the
// horizontal phase = 8 groups of 8 outputs (9 doubles apart) x 2 rows: 16 distinct 64-bit banks.
// Summation order is that of the generic kernel, so the results are bit-identical to it.
#define BS7_M 7
#define BS7_WIN 15
#define BS7_COLS 64
#define BS7_ROWS 64                      // rows per strip (one marching segment)
#define BS7_SPAN (BS7_COLS + 2 * BS7_M)  // 78 strip columns
#define BS7_VC 99                        // doubles per (row, channel) of vertical sums
#define BS7_VR 504
#define BS7_HC 72                        // doubles per (row, channel) of window sums
#define BS7_HR (5 * BS7_HC)
#define BS7_SUB 8                        // rows per shared-memory phase (a window of 15 rows = 8 + 7)
#define BS7_THREADS 416                  // >= 78 * 5 owners, 13 warps
#ifndef BS7_BLOCKS
#define BS7_BLOCKS 3
#endif
DOFS_D constexpr int bs7_pad(int i) { return i + (i >> 3); }

// horizontal sliding windows and the 2x2 solve for the nb rows whose vertical sums are in s_v (rows yb .. yb+nb-1)
// FUSE: the solved flow of a pixel is not stored; FarnebackUpdateMatrices of the NEXT iteration is applied to it at once
// (it only needs the pixel's own new flow) and its five products go to the other M buffer.
struct Bs7Fuse {
    const float* R;  // polynomial expansions
    float* M_out;    // the M buffer the next iteration reads
    PairSlots ps;
};

template <bool FUSE>
DOFS_D void bs7_rows(const double* s_v, double* s_h, float2* __restrict__ flow, int pair, int x0, int yb, int nb, int Wk, int Hk,
                     double scale, const Bs7Fuse& fu) {
    constexpr int m = BS7_M;
    const int e = threadIdx.x;
    const int hg = e & 7, hr = (e >> 3) & 7, hc = e >> 6;  // work item = (group of 8 outputs, row, channel), groups fastest
    __syncthreads();
    if (hc < 5 && hr < nb) {
        const double* v = s_v + hr * BS7_VR + hc * BS7_VC + 9 * hg;  // column 8*hg of the strip
        double* h = s_h + hr * BS7_HR + hc * BS7_HC + 9 * hg;
        double t = 0, lead[7];  // the first seven columns leave the window one by one: keep them instead of reading them twice
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            lead[i] = v[bs7_pad(i)];
            t += lead[i];
        }
#pragma unroll
        for (int i = 7; i <= 2 * m; ++i) t += v[bs7_pad(i)];
        h[0] = t;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            t += v[bs7_pad(k + 2 * m)] - lead[k - 1];
            h[k] = t;
        }
    }
    __syncthreads();
    for (int o = e; o < nb * BS7_COLS; o += BS7_THREADS) {
        const int tx = o & (BS7_COLS - 1), r = o >> 6;
        const int x = x0 + tx, y = yb + r;
        if (x >= Wk) continue;
        const double* h = s_h + r * BS7_HR + bs7_pad(tx);
        const double g11 = xdmul(h[0], scale), g12 = xdmul(h[BS7_HC], scale), g22 = xdmul(h[2 * BS7_HC], scale);
        const double h1 = xdmul(h[3 * BS7_HC], scale), h2 = xdmul(h[4 * BS7_HC], scale);
        const double idet = xddiv(1.0, xdadd(xdsub(xdmul(g11, g22), xdmul(g12, g12)), 1e-3));
        const float u = (float)xdmul(xdsub(xdmul(g11, h2), xdmul(g12, h1)), idet);
        const float w = (float)xdmul(xdsub(xdmul(g22, h1), xdmul(g12, h2)), idet);
        if (FUSE) update_matrices_pixel(fu.R, fu.M_out, Wk, Hk, fu.ps, pair, x, y, make_float2(u, w));
        else flow[((size_t)pair * Hk + y) * Wk + x] = make_float2(u, w);
    }
    __syncthreads();
}

template <bool FUSE>
__global__ void __launch_bounds__(BS7_THREADS, BS7_BLOCKS)
k_box_solve7(const float* __restrict__ M, float2* __restrict__ flow, int Wk, int Hk, Bs7Fuse fu) {
    extern __shared__ __align__(16) unsigned char bx7_smem[];
    double* s_v = reinterpret_cast<double*>(bx7_smem);  // [BS7_SUB][5][BS7_VC] (+ padding to BS7_VR)
    double* s_h = s_v + BS7_SUB * BS7_VR;               // [BS7_SUB][5][BS7_HC]
    constexpr int m = BS7_M;
    const int pair = blockIdx.z;
    const int x0 = blockIdx.x * BS7_COLS;
    const int y_begin = blockIdx.y * BS7_ROWS, y_end = min(y_begin + BS7_ROWS, Hk);
    const float* src = M + (size_t)pair * Wk * Hk * 5;
    const size_t pitch = (size_t)Wk * 5;
    const int e = threadIdx.x;
    const bool owner = e < BS7_SPAN * 5;
    const float* col = src;
    int v_at = 0;
    if (owner) {
        const int cx = e / 5, c = e - cx * 5;
        col = src + (size_t)min(max(x0 + cx - m, 0), Wk - 1) * 5 + c;  // replicated border columns
        v_at = c * BS7_VC + bs7_pad(cx);
    }
    double s = 0;
    if (owner) {
#pragma unroll
        for (int j = -m; j <= m; ++j) s += (double)col[(size_t)min(max(y_begin + j, 0), Hk - 1) * pitch];
    }
    const double scale = 1.0 / (double)(BS7_WIN * BS7_WIN);
    for (int yb = y_begin; yb < y_end; yb += BS7_SUB) {
        const int nb = min(BS7_SUB, y_end - yb);
        if (owner) {
            float vin[BS7_SUB], vout[BS7_SUB];
#pragma unroll
            for (int r = 0; r < BS7_SUB; ++r) {
                const int y = yb + r;
                vin[r] = col[(size_t)min(y + m, Hk - 1) * pitch];
                vout[r] = col[(size_t)max(y - m - 1, 0) * pitch];
            }
#pragma unroll
            for (int r = 0; r < BS7_SUB; ++r) {
                if (yb + r != y_begin) s += (double)vin[r] - (double)vout[r];
                s_v[r * BS7_VR + v_at] = s;
            }
        }
        bs7_rows<FUSE>(s_v, s_h, flow, pair, x0, yb, nb, Wk, Hk, scale, fu);
    }
}

// ---- the same kernel with FLOAT window sums in shared memory ----------------------------------------------------------
// k_box_solve7 is bound by 64-bit shared-memory wavefronts (ncu: l1tex 76 %, DRAM 22 %): five channels of double window sums
// cross shared memory three times.  Here the vertical 15-row sum is still accumulated in a double register (first window
// direct, then + entering - leaving, exactly as above) but rounded to float ONCE when it is handed over; the horizontal
// windows are balanced float trees (<= 5 roundings each); only the 2x2 solve converts back to double.  Every shared-memory access is 32 bits wide:
// half the wavefronts, a third of the footprint (25.6 KB instead of 55.3 KB per block).
// Accuracy: OpenCV keeps these sums in double; the float hand-over perturbs a window sum by <= 2^-24 relative and the
// 15 + 7 float additions of a horizontal window by about 1e-6 relative.  Measured with oracle/farneback_np.py on the
// repo's pair against cv2: EPE max 1.6e-4 / mean 1.37e-6 with double sums, max 2.2e-4 / mean 1.46e-6 with float sums in BOTH
// directions (the stated bar is 1e-3 / 1e-5) — the flow tolerance does not notice; tests/test_gpu_fullsize.py checks the
// 1080p and 4K fields on the device.
#define BS7F_VC 87   // floats per (row, channel) of vertical sums: conflict-free sliding-window reads (brute-forced layout)
#define BS7F_VR 440
#define BS7F_HC 71   // floats per (row, channel) of window sums
#define BS7F_HR 360
#ifndef BS7F_BLOCKS
#define BS7F_BLOCKS 3
#endif
#ifndef BS7F_TREE
#define BS7F_TREE 1  // 1: every window as a balanced tree (default: the quietest); 0: one tree, the others slid outwards from it
                     // (7 % faster alone, same throughput in the bench, but the border-band scatter of the 256x144 video test
                     // grows from below 1 % to 1.2 % of the pixels above 1e-3 px)
#endif
#ifndef BS7F_SUB
#define BS7F_SUB 8    // rows per shared-memory phase
#endif
#ifndef BS7F_ROWS
#define BS7F_ROWS 64  // rows per strip
#endif
#define BS7F_SMEM ((size_t)BS7F_SUB * (BS7F_VR + BS7F_HR) * sizeof(float))

template <bool FUSE>
DOFS_D void bs7f_rows(const float* s_v, float* s_h, float2* __restrict__ flow, int pair, int x0, int yb, int nb, int Wk, int Hk,
                      double scale, const Bs7Fuse& fu) {
    constexpr int m = BS7_M;
    const int e = threadIdx.x;
    __syncthreads();
    for (int wi = e; wi < 5 * 8 * BS7F_SUB; wi += BS7_THREADS) {
        const int hg = wi & 7, hr = (wi >> 3) % BS7F_SUB, hc = wi / (8 * BS7F_SUB);  // (group of 8 outputs, row, channel), groups fastest
        if (hr >= nb) continue;
        const float* v = s_v + hr * BS7F_VR + hc * BS7F_VC + 9 * hg;  // column 8*hg of the strip
        float* h = s_h + hr * BS7F_HR + hc * BS7F_HC + 9 * hg;
        float a[2 * m + 8];
#pragma unroll
        for (int i = 0; i < 2 * m + 8; ++i) a[i] = v[bs7_pad(i)];
#if BS7F_TREE
        // the eight 15-wide windows of the group as balanced trees over shared partial sums (pairs, quads, octets): at
        // most five float roundings per window
        float p2[2 * m + 7], p4[2 * m + 5], p8[2 * m + 1];
#pragma unroll
        for (int i = 0; i < 2 * m + 7; ++i) p2[i] = xfadd(a[i], a[i + 1]);
#pragma unroll
        for (int i = 0; i < 2 * m + 5; ++i) p4[i] = xfadd(p2[i], p2[i + 2]);
#pragma unroll
        for (int i = 0; i < 2 * m + 1; ++i) p8[i] = xfadd(p4[i], p4[i + 4]);
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = xfadd(p8[k], xfadd(p4[k + 8], xfadd(p2[k + 12], a[k + 14])));
#else
        // window 4 (of 0..7) as a balanced tree (4 roundings), the others by sliding outwards from it: + the entering value
        // - the leaving one, at most 4 steps = 8 more roundings (a window slid all the way from a sequential sum of
        // window 0 carries 15 + 14; that form was measurably noisier at the ill-conditioned pixels of the 4K field)
        const float q0 = xfadd(xfadd(a[4], a[5]), xfadd(a[6], a[7])), q1 = xfadd(xfadd(a[8], a[9]), xfadd(a[10], a[11]));
        const float q2 = xfadd(xfadd(a[12], a[13]), xfadd(a[14], a[15])), q3 = xfadd(xfadd(a[16], a[17]), a[18]);
        float t = xfadd(xfadd(q0, q1), xfadd(q2, q3));  // a[4..18]
        h[4] = t;
        float up = t;
#pragma unroll
        for (int k = 5; k < 8; ++k) {
            up = xfadd(up, xfsub(a[k + 14], a[k - 1]));
            h[k] = up;
        }
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            t = xfadd(t, xfsub(a[k], a[k + 15]));
            h[k] = t;
        }
#endif
    }
    __syncthreads();
    for (int o = e; o < nb * BS7_COLS; o += BS7_THREADS) {
        const int tx = o & (BS7_COLS - 1), r = o >> 6;
        const int x = x0 + tx, y = yb + r;
        if (x >= Wk) continue;
        const float* h = s_h + r * BS7F_HR + bs7_pad(tx);
        const double g11 = xdmul((double)h[0], scale), g12 = xdmul((double)h[BS7F_HC], scale), g22 = xdmul((double)h[2 * BS7F_HC], scale);
        const double h1 = xdmul((double)h[3 * BS7F_HC], scale), h2 = xdmul((double)h[4 * BS7F_HC], scale);
        const double idet = xddiv(1.0, xdadd(xdsub(xdmul(g11, g22), xdmul(g12, g12)), 1e-3));
        const float u = (float)xdmul(xdsub(xdmul(g11, h2), xdmul(g12, h1)), idet);
        const float w = (float)xdmul(xdsub(xdmul(g22, h1), xdmul(g12, h2)), idet);
        if (FUSE) update_matrices_pixel(fu.R, fu.M_out, Wk, Hk, fu.ps, pair, x, y, make_float2(u, w));
        else flow[((size_t)pair * Hk + y) * Wk + x] = make_float2(u, w);
    }
    __syncthreads();
}

template <bool FUSE>
__global__ void __launch_bounds__(BS7_THREADS, BS7F_BLOCKS)
k_box_solve7f(const float* __restrict__ M, float2* __restrict__ flow, int Wk, int Hk, Bs7Fuse fu) {
    extern __shared__ __align__(16) unsigned char bx7f_smem[];
    float* s_v = reinterpret_cast<float*>(bx7f_smem);  // [BS7_SUB][5][BS7F_VC] (+ padding to BS7F_VR)
    float* s_h = s_v + BS7F_SUB * BS7F_VR;              // [BS7_SUB][5][BS7F_HC]
    constexpr int m = BS7_M;
    const int pair = blockIdx.z;
    const int x0 = blockIdx.x * BS7_COLS;
    const int y_begin = blockIdx.y * BS7F_ROWS, y_end = min(y_begin + BS7F_ROWS, Hk);
    const float* src = M + (size_t)pair * Wk * Hk * 5;
    const int pitch = Wk * 5;  // (row * pitch stays an int product + one wide multiply-add per address: the vertical phase is
                               //  half of this kernel's instructions, and 64-bit index arithmetic was a third of those)
    const int e = threadIdx.x;
    const bool owner = e < BS7_SPAN * 5;
    const float* col = src;
    int v_at = 0;
    if (owner) {
        const int cx = e / 5, c = e - cx * 5;
        col = src + (size_t)min(max(x0 + cx - m, 0), Wk - 1) * 5 + c;  // replicated border columns
        v_at = c * BS7F_VC + bs7_pad(cx);
    }
    double s = 0;
    if (owner) {
#pragma unroll
        for (int j = -m; j <= m; ++j) s += (double)col[(long long)(min(max(y_begin + j, 0), Hk - 1) * pitch)];
    }
    const double scale = 1.0 / (double)(BS7_WIN * BS7_WIN);
    for (int yb = y_begin; yb < y_end; yb += BS7F_SUB) {
        const int nb = min(BS7F_SUB, y_end - yb);
        if (owner) {
#pragma unroll
            for (int r0 = 0; r0 < BS7F_SUB; r0 += 4) {  // four rows of loads in flight at a time
                float vin[4], vout[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int y = yb + r0 + r;
                    vin[r] = col[(long long)(min(y + m, Hk - 1) * pitch)];
                    vout[r] = col[(long long)(max(y - m - 1, 0) * pitch)];
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (yb + r0 + r != y_begin) s += (double)vin[r] - (double)vout[r];
                    s_v[(r0 + r) * BS7F_VR + v_at] = (float)s;
                }
            }
        }
        bs7f_rows<FUSE>(s_v, s_h, flow, pair, x0, yb, nb, Wk, Hk, scale, fu);
    }
}

#define BS7_SMEM ((size_t)BS7_SUB * (BS7_VR + BS7_HR) * sizeof(double))

inline size_t box_solve_smem(int m) {
    const size_t cols = BS_COLS + 2 * m;
    return (size_t)BS_BATCH * cols * 5 * 8 + (size_t)BS_BATCH * BS_COLS * 5 * 8;
}
inline int box_solve_threads(int m) { return (((BS_COLS + 2 * m) * 5 + 31) / 32) * 32; }

inline int farneback_set_attributes() {
    if (cudaFuncSetAttribute(k_box_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)box_solve_smem(63 / 2)) != cudaSuccess)
        return 3;
    if (cudaFuncSetAttribute(k_box_solve7<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS7_SMEM) != cudaSuccess) return 3;
    if (cudaFuncSetAttribute(k_box_solve7<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS7_SMEM) != cudaSuccess) return 3;
    if (BS7F_SMEM > 48 * 1024) {
        if (cudaFuncSetAttribute(k_box_solve7f<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS7F_SMEM) != cudaSuccess) return 3;
        if (cudaFuncSetAttribute(k_box_solve7f<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS7F_SMEM) != cudaSuccess) return 3;
    }
    return 0;
}

// the same preferred shared-memory carveout for every kernel of the flow stage (see LAUNCH in dofs3d.cu)
inline void farneback_set_carveout(int pct) {
    const void* ks[] = {(const void*)k_bgr2gray, (const void*)k_pyr_level0, (const void*)k_pyr_level<0>, (const void*)k_pyr_level<1>, (const void*)k_pyr_level<4>,
                        (const void*)k_pyr_level<9>, (const void*)k_pyr_level_tiled,
                        (const void*)k_polyexp<5>, (const void*)k_polyexp<0>, (const void*)k_update_matrices<UM_FLOW>,
                        (const void*)k_update_matrices<UM_START>, (const void*)k_box_solve, (const void*)k_box_solve7<false>,
                        (const void*)k_box_solve7<true>, (const void*)k_box_solve7f<false>, (const void*)k_box_solve7f<true>};
    for (const void* k : ks) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// ---------------------------------------------------------------------------------------------
// driver: n pairs; images of pair p are gray0 + p*N and gray1 + p*N.  When gray1 == gray0 + N the
// batch is a video (pair p = frames p, p+1) and every frame is expanded once instead of twice.
// Streaming (video batches only): carry_in = frame 0 was the last frame of the previous batch and its polynomial
// expansion is in R_carry (it is not expanded again); carry_out = keep the expansion of the last frame in R_carry.
// Returns 0, or non-zero after a launch error.
// ---------------------------------------------------------------------------------------------
inline int farneback_run(FlowBuffers& fb, const u8* d_gray0, const u8* d_gray1, int n, float2* d_flow_out,
                         cudaStream_t stream, FlowLaunchStats* st, bool carry_in = false, bool carry_out = false) {
    const size_t N = (size_t)fb.W * fb.H;
    const bool video = d_gray1 == d_gray0 + N;
    ImageSet imgs;
    imgs.base0 = d_gray0;
    imgs.base1 = d_gray1;
    imgs.split = video ? n + 1 : n;
    imgs.count = video ? n + 1 : 2 * n;
    PairSlots ps;
    ps.first0 = 0;
    ps.first1 = video ? 1 : n;
    const int m = fb.cfg.winsize / 2;
    const size_t bx_smem = box_solve_smem(m);
    if ((carry_in || carry_out) && !video) return 3;
    const int skip = carry_in ? 1 : 0;  // images not expanded here
    ImageSet fresh = imgs;
    fresh.base0 = imgs.base0 + (size_t)skip * N;
    fresh.split -= skip;
    fresh.count -= skip;
    float2* cur = nullptr;   // flow of the level being refined
    float2* prev = nullptr;  // flow of the coarser level
    int Wp = 0, Hp = 0;
    for (int k = fb.n_levels - 1; k >= 0; --k) {
        const FlowLevel& L = fb.level[k];
        const dim3 blk(256);
        const dim3 g_img((L.w + 31) / 32, (L.h + 7) / 8, fresh.count);
        const size_t npx = (size_t)L.w * L.h;
        float* I_fresh = fb.I + (size_t)skip * npx;
        float* R_fresh = fb.R + (size_t)skip * npx * 5;
        const dim3 g_pair((L.w + 31) / 32, (L.h + 7) / 8, n);
        const dim3 g_box((L.w + BS_COLS - 1) / BS_COLS, (L.h + BS_ROWS - 1) / BS_ROWS, n);
        cur = (k == 0) ? d_flow_out : (prev == fb.flowA ? fb.flowB : fb.flowA);
        FlowStart start;
        start.prev = prev;
        start.Wp = Wp;
        start.Hp = Hp;
        start.mul = 1.0 / fb.cfg.pyr_scale;
        start.scx = Wp > 0 ? (double)Wp / L.w : 1.0;
        start.scy = Hp > 0 ? (double)Hp / L.h : 1.0;
        if (L.w == fb.W && L.h == fb.H && L.taps.radius == 1)
            k_pyr_level0<<<g_img, blk, 0, stream>>>(fresh, I_fresh, fb.W, fb.H, L.taps);
        else {
            int TW, TH;
            pyr_tile_dims(fb.W, fb.H, L.w, L.h, L.taps.radius, &TW, &TH);
            const size_t smem = (size_t)TH * 64 * sizeof(float) + (size_t)TH * TW;
            if (smem <= PYR_TILED_MAX_SMEM && !fb.pyr_untiled)
                k_pyr_level_tiled<<<g_img, blk, smem, stream>>>(fresh, I_fresh, fb.W, fb.H, L.w, L.h, L.taps, TW, TH);
            else if (L.taps.radius == 1 && !fb.pyr_generic)
                k_pyr_level<1><<<g_img, blk, 0, stream>>>(fresh, I_fresh, fb.W, fb.H, L.w, L.h, L.taps);
            else if (L.taps.radius == 4 && !fb.pyr_generic)
                k_pyr_level<4><<<g_img, blk, 0, stream>>>(fresh, I_fresh, fb.W, fb.H, L.w, L.h, L.taps);
            else if (L.taps.radius == 9 && !fb.pyr_generic)
                k_pyr_level<9><<<g_img, blk, 0, stream>>>(fresh, I_fresh, fb.W, fb.H, L.w, L.h, L.taps);
            else
                k_pyr_level<0><<<g_img, blk, 0, stream>>>(fresh, I_fresh, fb.W, fb.H, L.w, L.h, L.taps);
        }
        FLOW_MARK(st, "flow.pyramid");
        if (fb.poly.n == 5)
            k_polyexp<5><<<g_img, blk, 0, stream>>>(I_fresh, R_fresh, L.w, L.h, fb.poly);
        else
            k_polyexp<0><<<g_img, blk, 0, stream>>>(I_fresh, R_fresh, L.w, L.h, fb.poly);
        if (carry_in)
            cudaMemcpyAsync(fb.R, fb.R_carry + fb.carry_off[k], npx * 5 * sizeof(float), cudaMemcpyDeviceToDevice, stream);
        if (carry_out)
            cudaMemcpyAsync(fb.R_carry + fb.carry_off[k], fb.R + (size_t)n * npx * 5, npx * 5 * sizeof(float),
                            cudaMemcpyDeviceToDevice, stream);
        FLOW_MARK(st, "flow.polyexp");
        k_update_matrices<UM_START><<<g_pair, blk, 0, stream>>>(fb.R, cur, fb.M, L.w, L.h, ps, start);
        FLOW_MARK(st, "flow.update_matrices");
        st->launches += 3;
        float* M_in = fb.M;
        float* M_next = fb.M2;
        for (int it = 0; it < fb.cfg.iters; ++it) {
            const bool last = it == fb.cfg.iters - 1;
            const int rows7 = fb.bs7_float ? BS7F_ROWS : BS7_ROWS;
            const dim3 g7((L.w + BS7_COLS - 1) / BS7_COLS, (L.h + rows7 - 1) / rows7, n);
            if (m == BS7_M && fb.fuse_um && !last) {
                // box filter + solve + the next iteration's update-matrices in one kernel: the flow of this iteration is
                // never stored, its M goes to the other buffer (blocks still read the halo of the current one)
                Bs7Fuse fu;
                fu.R = fb.R;
                fu.M_out = M_next;
                fu.ps = ps;
                if (fb.bs7_float) k_box_solve7f<true><<<g7, BS7_THREADS, BS7F_SMEM, stream>>>(M_in, cur, L.w, L.h, fu);
                else k_box_solve7<true><<<g7, BS7_THREADS, BS7_SMEM, stream>>>(M_in, cur, L.w, L.h, fu);
                st->launches++;
                FLOW_MARK(st, k == 0 ? "flow.box_solve.L0" : "flow.box_solve");
                float* t = M_in;
                M_in = M_next;
                M_next = t;
                continue;
            }
            if (m == BS7_M && fb.bs7_float)
                k_box_solve7f<false><<<g7, BS7_THREADS, BS7F_SMEM, stream>>>(M_in, cur, L.w, L.h, Bs7Fuse());
            else if (m == BS7_M)
                k_box_solve7<false><<<g7, BS7_THREADS, BS7_SMEM, stream>>>(M_in, cur, L.w, L.h, Bs7Fuse());
            else
                k_box_solve<<<g_box, box_solve_threads(m), bx_smem, stream>>>(M_in, cur, L.w, L.h, m);
            st->launches++;
            FLOW_MARK(st, k == 0 ? "flow.box_solve.L0" : "flow.box_solve");
            if (!last) {
                k_update_matrices<UM_FLOW><<<g_pair, blk, 0, stream>>>(fb.R, cur, M_in, L.w, L.h, ps, start);
                st->launches++;
                FLOW_MARK(st, "flow.update_matrices");
            }
        }
        prev = cur;
        Wp = L.w;
        Hp = L.h;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
