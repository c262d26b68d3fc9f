// Synthetic traffic video: a static textured background and rigid textured
// rectangles that move up / down the image by an integer number of pixels per frame.  Everything is
// integer arithmetic on 32-bit hashes, so this device generator and the numpy generator
// denseopticalflowsegmentation3d_b200/synth.py produce bit-identical frames (tests/test_synth.py).
// Not part of the reference; it only feeds the benchmark and the parity tests.
#pragma once
#include "dofs_common.cuh"

#define SYNTH_MAX_OBJECTS 64
#define SYNTH_PERIOD 8  // frames of travel before an object turns back

struct SynthObject {
    int x0, y0, w, h, dx, dy;
    u32 salt;
    int pad;
};

struct SynthScene {
    int n, s;  // objects, resolution scale (H / 360, at least 1)
    u32 seed;
    int pad;
    SynthObject o[SYNTH_MAX_OBJECTS];
};

__host__ __device__ inline u32 synth_hash(u32 a) {
    a ^= a >> 16;
    a *= 0x7feb352du;
    a ^= a >> 15;
    a *= 0x846ca68bu;
    a ^= a >> 16;
    return a;
}

__host__ __device__ inline u32 synth_lattice(u32 ix, u32 iy, u32 salt) {
    return synth_hash(ix * 0x9E3779B1u ^ synth_hash(iy * 0x85EBCA77u ^ salt)) & 255u;
}

// integer value noise: bilinear interpolation of lattice values, cell x cell pixels per lattice cell
__host__ __device__ inline u32 synth_noise(u32 x, u32 y, u32 cell, u32 salt) {
    const u32 ix = x / cell, iy = y / cell, fx = x % cell, fy = y % cell;
    const u32 v00 = synth_lattice(ix, iy, salt), v10 = synth_lattice(ix + 1, iy, salt);
    const u32 v01 = synth_lattice(ix, iy + 1, salt), v11 = synth_lattice(ix + 1, iy + 1, salt);
    const u32 top = v00 * (cell - fx) + v10 * fx;
    const u32 bot = v01 * (cell - fx) + v11 * fx;
    return (top * (cell - fy) + bot * fy) / (cell * cell);
}

// one colour channel of the texture `salt` at (x, y): two octaves
__host__ __device__ inline u32 synth_texture(u32 x, u32 y, u32 s, u32 salt, u32 channel) {
    const u32 coarse = synth_noise(x, y, 16u * s, salt);
    const u32 fine = synth_noise(x, y, 4u * s, salt * 3u + channel + 1u);
    return (2u * coarse + fine) / 3u;
}

inline u32 synth_rand(u32 seed, u32 i, u32 j) {
    return synth_hash(seed * 0x9E3779B1u + i * 0x85EBCA77u + j * 0xC2B2AE3Du + 12345u);
}

inline void synth_make_scene(u32 seed, int n_objects, int W, int H, SynthScene* sc) {
    const int s = H / 360 > 0 ? H / 360 : 1;
    sc->n = n_objects;
    sc->s = s;
    sc->seed = seed;
    sc->pad = 0;
    const int P = SYNTH_PERIOD;
    const int m_max = 2 + 4 * s;
    for (int i = 0; i < n_objects; ++i) {
        SynthObject& o = sc->o[i];
        o.w = s * (40 + (int)(synth_rand(seed, i, 0) % 71u));
        o.h = s * (30 + (int)(synth_rand(seed, i, 1) % 41u));
        if (o.w > W / 2) o.w = W / 2;
        if (o.h > H / 4) o.h = H / 4;
        const int x_lo = P, x_hi = W - o.w - P;
        const int y_lo = H / 8 + m_max * P, y_hi = H - o.h - m_max * P;
        o.x0 = x_lo + (int)(synth_rand(seed, i, 2) % (u32)(x_hi > x_lo ? x_hi - x_lo : 1));
        o.y0 = y_lo + (int)(synth_rand(seed, i, 3) % (u32)(y_hi > y_lo ? y_hi - y_lo : 1));
        const int m = 2 + (4 * s * (o.y0 + o.h)) / H;  // faster towards the bottom of the image (closer)
        o.dy = (synth_rand(seed, i, 4) & 1u) ? m : -m;
        o.dx = (int)(synth_rand(seed, i, 5) % 3u) - 1;
        o.salt = synth_hash(seed ^ (0xA5A5u + (u32)i * 977u));
        o.pad = 0;
    }
}

__host__ __device__ inline int synth_travel(int frame) {
    const int ph = frame % (2 * SYNTH_PERIOD);
    return ph < SYNTH_PERIOD ? ph : 2 * SYNTH_PERIOD - ph;
}

// frame `first_frame + blockIdx.y`: BGR u8 [frame][H][W][3]
__global__ void __launch_bounds__(256)
k_synth_frames(SynthScene sc, int first_frame, u8* __restrict__ out, int W, int H) {
    const int f = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int y = p / W, x = p - y * W;
    const int travel = synth_travel(first_frame + f);
    u32 tx = (u32)x, ty = (u32)y, salt = synth_hash(sc.seed ^ 0xBACC0001u);
    for (int i = sc.n - 1; i >= 0; --i) {
        const SynthObject& o = sc.o[i];
        const int ox = o.x0 + o.dx * travel, oy = o.y0 + o.dy * travel;
        if (x >= ox && x < ox + o.w && y >= oy && y < oy + o.h) {
            tx = (u32)(x - ox);
            ty = (u32)(y - oy);
            salt = o.salt;
            break;
        }
    }
    u8* dst = out + ((size_t)f * W * H + p) * 3;
    dst[0] = (u8)synth_texture(tx, ty, (u32)sc.s, salt, 0u);
    dst[1] = (u8)synth_texture(tx, ty, (u32)sc.s, salt, 1u);
    dst[2] = (u8)synth_texture(tx, ty, (u32)sc.s, salt, 2u);
}
