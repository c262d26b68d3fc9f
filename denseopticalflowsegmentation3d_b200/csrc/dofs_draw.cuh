// Display side of the path: what plot_best_segments_simple (cpp/src/draw.cpp:102-160) returns for a
// frame — segments with score > min_score painted in ascending root order in their class colour (draw.cpp:120-147), the
// wireframe cube of each (draw_cube, draw.cpp:85-100: twelve cv::line calls, colour (255, 0, 0), thickness 1) drawn on
// both the frame and the painted copy, and the two blended with cv::addWeighted(frame, 1 - 2/5, seg, 2/5) (draw.cpp:157-158).
// OpenCV's drawing is restated as OpenCV 4.x does it (oracle/draw_np.py is the same restatement, bit-exact against cv2):
//   * Point2f -> Point rounds half to even (saturate_cast<int>);
//   * cv::line, LINE_8, thickness 1 = clipLine on 64-bit integers (the clipped end points, not the discarded pixels of the
//     unclipped line, define the raster) + the 8-connected LineIterator started from the left end;
//   * addWeighted on 8-bit data: float products and sum, round half to even, saturate.
// Painting order: segment i's cube is drawn right after segment i is painted, so on the painted copy a cube pixel survives
// unless a LATER painted segment covers it, i.e. iff i >= painted[p] (the largest painting index at p, from k_paint); on
// the frame cubes are never painted over.  All cubes share one colour, so the lines need no order among themselves.
#pragma once
#include "dofs_common.cuh"

DOFS_D long long draw_round(float v) {  // saturate_cast<int>(float): cvRound; NaN and out-of-range values -> INT_MIN
    if (!(v >= -2147483648.f && v <= 2147483520.f)) return -2147483648ll;
    return (long long)__float2int_rn(v);
}

DOFS_D long long draw_scaled(long long a, long long b, long long c) {  // (int64)((double)a * b / c)
    return __double2ll_rz(xddiv(xdmul((double)a, (double)b), (double)c));
}

// cv::clipLine(Size(w, h), pt1, pt2)
DOFS_D bool draw_clip_line(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += draw_scaled(a - y1, x2 - x1, y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += draw_scaled(a - y2, x2 - x1, y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += draw_scaled(a - x1, y2 - y1, x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += draw_scaled(a - x2, y2 - y1, x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// one thread per cube edge: 12 edges per box
template <typename Box>
__global__ void __launch_bounds__(128)
k_cube_lines(const Box* __restrict__ boxes, const int* __restrict__ n_boxes, int box_cap, const int* __restrict__ painted,
             u8* __restrict__ frame, u8* __restrict__ seg, int W, int H, double min_score) {
    const int fr = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t / 12, e = t - b * 12;
    if (b >= min(n_boxes[fr], box_cap)) return;
    const Box& bx = boxes[(size_t)fr * box_cap + b];
    if (!(bx.score > min_score)) return;
    const int i = e & 3, kind = e >> 2;  // kind 0: lower face edge, 1: upper face edge, 2: vertical edge
    const float* pa = kind == 1 ? bx.upper_face : bx.lower_face;
    const float* pb = kind == 0 ? bx.lower_face : bx.upper_face;
    const int ia = i, ib = kind == 2 ? i : (i + 1) & 3;
    long long x1 = draw_round(pa[2 * ia]), y1 = draw_round(pa[2 * ia + 1]);
    long long x2 = draw_round(pb[2 * ib]), y2 = draw_round(pb[2 * ib + 1]);
    if ((unsigned long long)x1 >= (unsigned long long)W || (unsigned long long)x2 >= (unsigned long long)W ||
        (unsigned long long)y1 >= (unsigned long long)H || (unsigned long long)y2 >= (unsigned long long)H) {
        if (!draw_clip_line(W, H, x1, y1, x2, y2)) return;
    }
    long long dx = x2 - x1, dy = y2 - y1;
    int step_y = 1;
    if (dx < 0) {  // leftToRight: start from the left end
        dx = -dx;
        dy = -dy;
        x1 = x2;
        y1 = y2;
    }
    if (dy < 0) {
        dy = -dy;
        step_y = -1;
    }
    const bool vert = dy > dx;
    if (vert) {
        const long long s = dx;
        dx = dy;
        dy = s;
    }
    long long err = dx - (dy + dy);
    const long long plus_delta = dx + dx, minus_delta = -(dy + dy);
    int x = (int)x1, y = (int)y1;
    const size_t fo = (size_t)fr * W * H;
    for (long long k = 0; k <= dx; ++k) {
        const size_t p = fo + (size_t)y * W + x;
        frame[3 * p] = 255, frame[3 * p + 1] = 0, frame[3 * p + 2] = 0;
        if (b >= painted[p]) seg[3 * p] = 255, seg[3 * p + 1] = 0, seg[3 * p + 2] = 0;
        const bool both = err < 0;
        err += minus_delta + (both ? plus_delta : 0);
        if (vert) {
            y += step_y;
            if (both) x += 1;
        } else {
            x += 1;
            if (both) y += step_y;
        }
    }
}

// cv::addWeighted(frame, alpha, seg, beta, 0, frame) on 8-bit data, four bytes per thread
__global__ void __launch_bounds__(256)
k_add_weighted(u8* __restrict__ frame, const u8* __restrict__ seg, size_t n_bytes, float alpha, float beta) {
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q >= n_bytes) return;
    const int cnt = (int)min((size_t)4, n_bytes - q);
    for (int k = 0; k < cnt; ++k) {
        const float t = xfadd(xfmul((float)frame[q + k], alpha), xfmul((float)seg[q + k], beta));
        frame[q + k] = (u8)min(max(__float2int_rn(t), 0), 255);
    }
}
