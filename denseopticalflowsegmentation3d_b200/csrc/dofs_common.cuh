// Common definitions of the sm_100a hot path (kernels are in the dofs_*.cuh headers, the context,
// orchestration and C ABI in dofs3d.cu).
//
// Exact arithmetic: the segmentation and lifting stages must reproduce the reference's host
// arithmetic bit for bit (plain x86-64 build: every float/double operation rounds once, no FMA
// contraction).  All such operations go through the *_rn helpers below, which nvcc never contracts.
#pragma once

#include <stdint.h>

#include <cuda_runtime.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define DOFS_INF32 0xFFFFFFFFu

#define DOFS_HD __host__ __device__ __forceinline__
#define DOFS_D __device__ __forceinline__

// x / d and x % d for a run-time divisor that stays the same for a whole launch (the image width): one multiply-high
// and one correction step instead of the ~25-instruction integer division.  m = fastdiv_magic(d) = floor((2^32 - 1) / d);
// floor(x m / 2^32) is floor(x / d) or one less for every x < 2^31.
DOFS_HD u32 fastdiv_magic(u32 d) { return 0xFFFFFFFFu / d; }
DOFS_D int fastdiv(int x, int d, u32 m, int* rem) {
    int q = (int)__umulhi((u32)x, m);
    int r = x - q * d;
    if (r >= d) {
        ++q;
        r -= d;
    }
    *rem = r;
    return q;
}

// float ops, one rounding each
DOFS_D float xfadd(float a, float b) { return __fadd_rn(a, b); }
DOFS_D float xfsub(float a, float b) { return __fsub_rn(a, b); }
DOFS_D float xfmul(float a, float b) { return __fmul_rn(a, b); }
DOFS_D float xfdiv(float a, float b) { return __fdiv_rn(a, b); }
// double ops, one rounding each
DOFS_D double xdadd(double a, double b) { return __dadd_rn(a, b); }
DOFS_D double xdsub(double a, double b) { return __dsub_rn(a, b); }
DOFS_D double xdmul(double a, double b) { return __dmul_rn(a, b); }
DOFS_D double xddiv(double a, double b) { return __ddiv_rn(a, b); }
DOFS_D double xdsqrt(double a) { return __dsqrt_rn(a); }

// sqrt((double)x*x + (double)y*y): cv::norm(Point2f) / cv::norm(Vec2f) and diff (segment.cpp:25-29).
// The two products of floats are exact in double, so the sum rounds once either way.
DOFS_D double norm2d(float x, float y) {
    double dx = (double)x, dy = (double)y;
    return xdsqrt(xdadd(xdmul(dx, dx), xdmul(dy, dy)));
}

struct Homographies {
    float persp[9];
    float inv[9];
    float upper[3][9];
};

struct SegParams {
    Homographies hg;
    int cls_size[3][2];
    double cls_min_convexity[3];
    double score_threshold;
    int min_size;
    int neighbors;
};

// One merge that passed the size / row / move gates of Forest::new_merge (graph.cpp:280-300).
struct __align__(16) Candidate {
    u32 root;
    u32 time;      // index of the merge in the reference's sequence of accepted edges
    int size;
    float fx, fy;  // mean flow of the merged set
    u16 bbox[4];   // xmin, ymin, xmax, ymax
    u32 pad;
};
