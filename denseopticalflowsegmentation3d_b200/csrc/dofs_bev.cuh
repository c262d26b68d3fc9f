// Bird's-eye-view warp: the reference's `transform` (cpp/src/lifting_3d.cpp:516-522) =
//     cv::warpPerspective(img, result, mat, Size(2500, 14000), INTER_CUBIC, BORDER_REPLICATE)
// on 8-bit images (called at segment.cpp:143,196 to build the `bev` image the Forest carries).
// OpenCV is not part of the reference tree; this restates the published algorithm of OpenCV 4.x imgwarp.cpp (the same
// restatement as oracle/warp_np.py, which is bit-exact against cv2 4.13):
//   * the 3x3 matrix is inverted in double on the host (adjugate / determinant, cv::invert);
//   * per destination pixel, in double and in OpenCV's evaluation order (blocks of 64 columns: the block's first column
//     goes through M0*x + M1*y + M2, the offset inside the block is added as M0*x1): X = cvRound(X0 / W0 * 32) clamped to
//     the int range (W0 == 0 -> 0); integer part X >> 5 saturated to short, fraction X & 31;
//   * 4 x 4 bicubic taps (a = -0.75) weighted by a 32 x 32 table of 15-bit fixed-point products whose 16 entries sum to
//     2^15 exactly (initInterTab2D), source coordinates clamped to the image (BORDER_REPLICATE);
//   * result = saturate_u8((sum + 2^14) >> 15).
// One thread per destination pixel; the source image (0.7 MB for the reference's 640 x 360 frame) and the 32 KB weight
// table stay in L1/L2, so the kernel is bound by its output stream: out_w * out_h * channels bytes written once
// (105 MB for the reference's BEV), staged per block row in shared memory and stored as 32-bit words.
#pragma once
#include <math.h>

#include <vector>

#include "dofs_common.cuh"

#define WARP_TAB 32
#define WARP_COEF_BITS 15

struct WarpMatrix {
    double m[9];  // inverse of the caller's matrix (destination -> source)
};

// interpolateCubic + initInterTab2D of OpenCV (float products, rounded to short, sum forced to 2^15)
inline void warp_cubic_table(std::vector<short>* out) {
    float tab1[WARP_TAB][4];
    const float scale = 1.f / WARP_TAB;
    for (int i = 0; i < WARP_TAB; ++i) {
        const volatile float x = i * scale;
        const float A = -0.75f;
        volatile float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        volatile float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
        volatile float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        volatile float c3 = 1.f - c0 - c1 - c2;
        tab1[i][0] = c0, tab1[i][1] = c1, tab1[i][2] = c2, tab1[i][3] = c3;
    }
    out->assign((size_t)WARP_TAB * WARP_TAB * 16, 0);
    for (int i = 0; i < WARP_TAB; ++i)
        for (int j = 0; j < WARP_TAB; ++j) {
            short* it = out->data() + ((size_t)i * WARP_TAB + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; ++k1)
                for (int k2 = 0; k2 < 4; ++k2) {
                    const volatile float v = tab1[i][k1] * tab1[j][k2];
                    long r = lrintf(v * (float)(1 << WARP_COEF_BITS));
                    r = r < -32768 ? -32768 : r > 32767 ? 32767 : r;
                    it[k1 * 4 + k2] = (short)r;
                    isum += (int)r;
                }
            if (isum != (1 << WARP_COEF_BITS)) {
                const int diff = isum - (1 << WARP_COEF_BITS);
                int Mk = 2 * 4 + 2, mk = 2 * 4 + 2;
                for (int k1 = 2; k1 < 4; ++k1)
                    for (int k2 = 2; k2 < 4; ++k2) {
                        if (it[k1 * 4 + k2] < it[mk]) mk = k1 * 4 + k2;
                        else if (it[k1 * 4 + k2] > it[Mk]) Mk = k1 * 4 + k2;
                    }
                if (diff < 0) it[Mk] = (short)(it[Mk] - diff);
                else it[mk] = (short)(it[mk] - diff);
            }
        }
}

// cv::invert of a 3x3 in double; all zeros when singular
inline void warp_invert(const float* mat9, WarpMatrix* out) {
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = (double)mat9[i];
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) {
        for (int i = 0; i < 9; ++i) out->m[i] = 0;
        return;
    }
    d = 1.0 / d;
    out->m[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    out->m[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    out->m[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    out->m[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    out->m[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    out->m[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    out->m[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    out->m[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    out->m[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

#define WARP_BW 64  // OpenCV's block width: the evaluation order of the coordinates depends on it
#define WARP_BH 16

template <int C>
__global__ void __launch_bounds__(256)
k_warp_perspective_cubic(const u8* __restrict__ src, int W, int H, u8* __restrict__ dst, int out_w, int out_h, WarpMatrix M,
                         const short* __restrict__ itab) {
    __shared__ __align__(16) u8 s_out[4][WARP_BW * C];
    const int tx = threadIdx.x & (WARP_BW - 1), tr = threadIdx.x >> 6;  // 4 rows per step
    const int xb = blockIdx.x * WARP_BW, yb = blockIdx.y * WARP_BH;
    const int x = xb + tx;
    const bool words = ((out_w * C) & 3) == 0 && ((xb * C) & 3) == 0 && xb + WARP_BW <= out_w &&
                       (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    for (int r0 = 0; r0 < WARP_BH; r0 += 4) {
        const int y = yb + r0 + tr;
        u8 px[C];
#pragma unroll
        for (int c = 0; c < C; ++c) px[c] = 0;
        if (x < out_w && y < out_h) {
            // plain multiplies and adds, one rounding each, like the host code (no FMA contraction)
            const double X0 = xdadd(xdadd(xdmul(M.m[0], (double)xb), xdmul(M.m[1], (double)y)), M.m[2]);
            const double Y0 = xdadd(xdadd(xdmul(M.m[3], (double)xb), xdmul(M.m[4], (double)y)), M.m[5]);
            const double W0 = xdadd(xdadd(xdmul(M.m[6], (double)xb), xdmul(M.m[7], (double)y)), M.m[8]);
            double Wv = xdadd(W0, xdmul(M.m[6], (double)tx));
            Wv = Wv != 0.0 ? xddiv((double)WARP_TAB, Wv) : 0.0;
            const double fX = fmax(-2147483648.0, fmin(2147483647.0, xdmul(xdadd(X0, xdmul(M.m[0], (double)tx)), Wv)));
            const double fY = fmax(-2147483648.0, fmin(2147483647.0, xdmul(xdadd(Y0, xdmul(M.m[3], (double)tx)), Wv)));
            const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
            const int sx = max(-32768, min(32767, X >> 5)) - 1, sy = max(-32768, min(32767, Y >> 5)) - 1;
            const short* wt = itab + (size_t)((Y & (WARP_TAB - 1)) * WARP_TAB + (X & (WARP_TAB - 1))) * 16;
            int acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = 0;
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const u8* row = src + (size_t)min(max(sy + ky, 0), H - 1) * W * C;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const u8* p = row + (size_t)min(max(sx + kx, 0), W - 1) * C;
                    const int w = wt[ky * 4 + kx];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] += w * (int)p[c];
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) px[c] = (u8)min(max((acc[c] + (1 << (WARP_COEF_BITS - 1))) >> WARP_COEF_BITS, 0), 255);
        }
        if (words) {
#pragma unroll
            for (int c = 0; c < C; ++c) s_out[tr][tx * C + c] = px[c];
            __syncthreads();
            constexpr int NW = WARP_BW * C / 4;  // words per block row
            for (int i = threadIdx.x; i < 4 * NW; i += 256) {
                const int rr = i / NW, wi = i - rr * NW;
                const int yy = yb + r0 + rr;
                if (yy < out_h)
                    reinterpret_cast<u32*>(dst + ((size_t)yy * out_w + xb) * C)[wi] = reinterpret_cast<const u32*>(s_out[rr])[wi];
            }
            __syncthreads();
        } else if (x < out_w && y < out_h) {
#pragma unroll
            for (int c = 0; c < C; ++c) dst[((size_t)y * out_w + x) * C + c] = px[c];
        }
    }
}
