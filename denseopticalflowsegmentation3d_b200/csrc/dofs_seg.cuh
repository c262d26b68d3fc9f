// Flow-graph clustering on the device (stages K6-K12 of SURVEY.md section 2.2).
//
// The reference (cpp/src/graph.cpp) is a strictly sequential Kruskal loop with union by rank, a
// running float mean per set and a per-merge scoring hook.  This file computes the SAME merge
// sequence and per-merge state in parallel, using two structural facts (checked against the reference
// in tests/ and tools/proto_parallel.py):
//
//   1. The loop only acts on the edges it accepts, and those are the minimum spanning forest under the strict
//      order (weight, insertion sequence) whatever else is in the sorted list.  So the 4N edge slots are never
//      sorted here: Boruvka compares edges directly, and only the <= N-1 accepted edges are ordered afterwards
//      (their positions are the merge times).
//   2. Boruvka levels under that order are exactly the union-by-rank ranks.
//      In level k every current component S (all have rank k) picks its minimum outgoing edge m_k(S).
//        * a pick that is not mutual: S loses at m_k(S) to whatever component holds the other
//          endpoint at that time (it has rank > k);
//        * a mutual pick (S and S' pick the same edge): a rank tie; the component holding `edge.end`
//          survives (graph.cpp:177-182) and becomes a level k+1 component.
//      So every root id loses exactly once, at the time of its edge, and `up[c]` = root of the next-level
//      component it is contracted into.  The root of any pixel's set at time t is found by climbing `up`
//      while loss time < t (<= max rank hops).
//
// Merge events are then grouped into per-root chains ordered by time, and the chains are replayed
// in waves of increasing final rank (a chain only absorbs roots of strictly lower final rank), which
// reproduces sizes, bounding boxes and the order-dependent float mean flow (graph.cpp:184-190)
// bit for bit.
#pragma once
#include "dofs_common.cuh"
#include "dofs_lift.cuh"

#define SEG_THREADS 256

// ---------------------------------------------------------------------------------------------
// K6  cv::GaussianBlur(flow, flow, Size(0,0), sigma) (segment.cpp:52): separable, BORDER_REFLECT_101.
// taps: 2*radius+1 float coefficients (host-computed like cv::getGaussianKernel, CV_32F).
// ---------------------------------------------------------------------------------------------
#define BLUR_MAX_RADIUS 32
struct BlurTaps {
    int radius;
    float k[2 * BLUR_MAX_RADIUS + 1];
};

DOFS_D int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// horizontal pass: one thread per pixel (float2), rows are contiguous so neighbouring loads coalesce/L1-hit
__global__ void __launch_bounds__(SEG_THREADS)
k_blur_rows(const float2* __restrict__ src, float2* __restrict__ dst, int W, int H, BlurTaps taps) {
    const int frame = blockIdx.y;
    const int N = W * H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int y = p / W, x = p - y * W;
    const float2* row = src + (size_t)frame * N + (size_t)y * W;
    float sx = 0.f, sy = 0.f;
    const int r = taps.radius;
    if (x >= r && x + r < W) {
        for (int k = -r; k <= r; ++k) {
            float2 v = row[x + k];
            float c = taps.k[k + r];
            sx = fmaf(c, v.x, sx);
            sy = fmaf(c, v.y, sy);
        }
    } else {
        for (int k = -r; k <= r; ++k) {
            float2 v = row[reflect101(x + k, W)];
            float c = taps.k[k + r];
            sx = fmaf(c, v.x, sx);
            sy = fmaf(c, v.y, sy);
        }
    }
    dst[(size_t)frame * N + p] = make_float2(sx, sy);
}

__global__ void __launch_bounds__(SEG_THREADS)
k_blur_cols(const float2* __restrict__ src, float2* __restrict__ dst, int W, int H, BlurTaps taps) {
    const int frame = blockIdx.y;
    const int N = W * H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int y = p / W, x = p - y * W;
    const float2* img = src + (size_t)frame * N;
    float sx = 0.f, sy = 0.f;
    const int r = taps.radius;
    for (int k = -r; k <= r; ++k) {
        int yy = y + k;
        if (yy < 0 || yy >= H) yy = reflect101(yy, H);
        float2 v = img[(size_t)yy * W + x];
        float c = taps.k[k + r];
        sx = fmaf(c, v.x, sx);
        sy = fmaf(c, v.y, sy);
    }
    dst[(size_t)frame * N + p] = make_float2(sx, sy);
}

// Fused form of the two passes above for the reference's radius (sigma 3 -> 25 taps): a 32x32 output tile
// with its halo is staged once in shared memory (reflected borders resolved while loading), filtered along
// rows into a second shared buffer and along columns into the output.  Each thread produces 4 consecutive
// outputs from one run of 2R+4 shared-memory reads, with exactly the accumulation order of k_blur_rows /
// k_blur_cols (so the results are bit-identical to the two-pass kernels).  HBM: 8 B read + 8 B written per pixel
// (the halo comes out of L2) instead of 32.
#define FB_T 32
template <int R>
__global__ void __launch_bounds__(256)
k_blur_fused(const float2* __restrict__ src, float2* __restrict__ dst, int W, int H, BlurTaps taps) {
    constexpr int C = FB_T + 2 * R;  // tile columns / rows with halo
    __shared__ float2 s_in[C][C + 1];
    __shared__ float2 s_h[C][FB_T + 1];
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * FB_T, y0 = blockIdx.y * FB_T;
    const float2* img = src + (size_t)frame * W * H;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (x0 >= R && y0 >= R && x0 + FB_T + R <= W && y0 + FB_T + R <= H) {  // no border in reach: plain rows (block-uniform)
        for (int ry = wrp; ry < C; ry += 8) {
            const float2* row = img + (size_t)(y0 + ry - R) * W + (x0 - R);
            for (int cx = lane; cx < C; cx += 32) s_in[ry][cx] = row[cx];
        }
    } else {
        for (int ry = wrp; ry < C; ry += 8) {
            const int yy = reflect101(y0 + ry - R, H);
            const float2* row = img + (size_t)yy * W;
            for (int cx = lane; cx < C; cx += 32) s_in[ry][cx] = row[reflect101(x0 + cx - R, W)];
        }
    }
    __syncthreads();
    // rows: (tile row, group of 4 columns)
    for (int w = threadIdx.x; w < C * (FB_T / 4); w += 256) {
        const int ry = w / (FB_T / 4), g = w % (FB_T / 4);
        float2 v[2 * R + 4];
#pragma unroll
        for (int i = 0; i < 2 * R + 4; ++i) v[i] = s_in[ry][4 * g + i];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                sx = fmaf(taps.k[k], v[o + k].x, sx);
                sy = fmaf(taps.k[k], v[o + k].y, sy);
            }
            s_h[ry][4 * g + o] = make_float2(sx, sy);
        }
    }
    __syncthreads();
    // columns: (tile column, group of 4 rows)
    {
        const int tx = threadIdx.x & 31, g = threadIdx.x >> 5;  // 8 groups of 4 rows
        float2 v[2 * R + 4];
#pragma unroll
        for (int i = 0; i < 2 * R + 4; ++i) v[i] = s_h[4 * g + i][tx];
        const int x = x0 + tx;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                sx = fmaf(taps.k[k], v[o + k].x, sx);
                sy = fmaf(taps.k[k], v[o + k].y, sy);
            }
            const int y = y0 + 4 * g + o;
            if (x < W && y < H) dst[(size_t)frame * W * H + (size_t)y * W + x] = make_float2(sx, sy);
        }
    }
}

// The same kernel with the tile staged by the TMA engine instead of LDG -> STS through the LSU pipe: for a tile whose halo
// lies inside the image, one elected thread arms an mbarrier with the tile's byte count and issues one bulk copy
// (cp.async.bulk, SASS: UBLKCP) per tile row — 56 x 448 contiguous bytes — straight into the padded shared-memory rows;
// the block waits on the barrier's phase.  Border tiles resolve their reflections with ordinary loads into the same
// layout.  Needs 16-byte aligned rows (even width, aligned base); the compute phases and their arithmetic are those of
// k_blur_fused, so the results are bit-identical.
DOFS_D u32 dofs_smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
DOFS_D void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dofs_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DOFS_D void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dofs_smem_addr(bar)), "r"(bytes) : "memory");
}
DOFS_D void bulk_copy_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dofs_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(dofs_smem_addr(bar))
                 : "memory");
}
DOFS_D bool mbar_try_wait(u64* bar, u32 phase) {
    u32 ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(dofs_smem_addr(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}

#ifndef BLUR_TMA_WARP_ISSUE
#define BLUR_TMA_WARP_ISSUE 1
#endif
template <int R>
__global__ void __launch_bounds__(256)
k_blur_fused_tma(const float2* __restrict__ src, float2* __restrict__ dst, int W, int H, BlurTaps taps) {
    constexpr int C = FB_T + 2 * R;  // tile columns / rows with halo
    constexpr int P = (C + 2) & ~1;  // row pitch in float2: a multiple of 16 bytes for the bulk copies
    __shared__ __align__(16) float2 s_in[C][P];
    __shared__ float2 s_h[C][FB_T + 1];
    __shared__ __align__(8) u64 s_bar;
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * FB_T, y0 = blockIdx.y * FB_T;
    const float2* img = src + (size_t)frame * W * H;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (x0 >= R && y0 >= R && x0 + FB_T + R <= W && y0 + FB_T + R <= H) {  // no border in reach (block-uniform)
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
#if BLUR_TMA_WARP_ISSUE
        if (wrp == 0) {  // the first warp issues the row copies, two rows a lane (one thread issuing all 56 was a third of
                         // the kernel's stall samples); the byte count is armed before any copy can complete
            if (lane == 0) mbar_expect_tx(&s_bar, (u32)(C * C * sizeof(float2)));
            __syncwarp();
            const float2* row = img + (size_t)(y0 - R) * W + (x0 - R);
            for (int ry = lane; ry < C; ry += 32) bulk_copy_g2s(&s_in[ry][0], row + (size_t)ry * W, (u32)(C * sizeof(float2)), &s_bar);
        }
#else
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, (u32)(C * C * sizeof(float2)));
            const float2* row = img + (size_t)(y0 - R) * W + (x0 - R);
            for (int ry = 0; ry < C; ++ry) bulk_copy_g2s(&s_in[ry][0], row + (size_t)ry * W, (u32)(C * sizeof(float2)), &s_bar);
        }
#endif
        while (!mbar_try_wait(&s_bar, 0u)) {
        }
    } else {
        for (int ry = wrp; ry < C; ry += 8) {
            const int yy = reflect101(y0 + ry - R, H);
            const float2* row = img + (size_t)yy * W;
            for (int cx = lane; cx < C; cx += 32) s_in[ry][cx] = row[reflect101(x0 + cx - R, W)];
        }
        __syncthreads();
    }
    // rows: (tile row, group of 4 columns)
    for (int w = threadIdx.x; w < C * (FB_T / 4); w += 256) {
        const int ry = w / (FB_T / 4), g = w % (FB_T / 4);
        float2 v[2 * R + 4];
#pragma unroll
        for (int i = 0; i < 2 * R + 4; ++i) v[i] = s_in[ry][4 * g + i];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                sx = fmaf(taps.k[k], v[o + k].x, sx);
                sy = fmaf(taps.k[k], v[o + k].y, sy);
            }
            s_h[ry][4 * g + o] = make_float2(sx, sy);
        }
    }
    __syncthreads();
    // columns: (tile column, group of 4 rows)
    {
        const int tx = threadIdx.x & 31, g = threadIdx.x >> 5;  // 8 groups of 4 rows
        float2 v[2 * R + 4];
#pragma unroll
        for (int i = 0; i < 2 * R + 4; ++i) v[i] = s_h[4 * g + i][tx];
        const int x = x0 + tx;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                sx = fmaf(taps.k[k], v[o + k].x, sx);
                sy = fmaf(taps.k[k], v[o + k].y, sy);
            }
            const int y = y0 + 4 * g + o;
            if (x < W && y < H) dst[(size_t)frame * W * H + (size_t)y * W + x] = make_float2(sx, sy);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7  edge weights: build_graph's enumeration (graph.cpp:62-93) with diff (segment.cpp:20-32).
// Slot 4*p+d of pixel p=(x,y): d=0 left (x-1,y), d=1 up (x,y-1), d=2 up-left (x-1,y-1),
// d=3 down-left (x-1,y+1); the slot index IS the reference's insertion sequence number.
// Non-existent border edges (and slots 2,3 in 4-neighbour mode) get +inf so they sort last.
// key = bit pattern of the f64 weight (non-negative doubles order like unsigned integers).
// ---------------------------------------------------------------------------------------------
DOFS_D u64 edge_key(float2 a, float2 b) {
    float dx = xfsub(a.x, b.x), dy = xfsub(a.y, b.y);
    return (u64)__double_as_longlong(norm2d(dx, dy));
}

#define EDGE_KEY_INVALID 0x7FF0000000000000ull
#define EDGE_PREFIX_INVALID 0xFFFFFFFFu

// Order-preserving 32-bit prefix of a weight: exponent rebased to 2^-200 (9 bits cover every weight a
// float flow field can produce) followed by the top 23 mantissa bits.  Monotone in the weight; weights
// that share a prefix are put in exact order afterwards (k_prefix_repair_*).
DOFS_D u32 edge_prefix(u64 key) {
    const u64 base = (u64)(1023 - 200) << 52;
    if (key >= EDGE_KEY_INVALID) return EDGE_PREFIX_INVALID;
    const u64 k = key > base ? key - base : 0ull;
    return (u32)min(k >> 29, (u64)0xFFFFFFFEu);
}

__global__ void __launch_bounds__(SEG_THREADS)
k_edge_keys(const float2* __restrict__ flow, u64* __restrict__ keys, u32* __restrict__ prefix, size_t key_stride, int W,
            int H, int neighbors8) {
    const int frame = blockIdx.y;
    const int N = W * H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int y = p / W, x = p - y * W;
    const float2* f = flow + (size_t)frame * N;
    const float2 c = f[p];
    u64 k0 = EDGE_KEY_INVALID, k1 = EDGE_KEY_INVALID, k2 = EDGE_KEY_INVALID, k3 = EDGE_KEY_INVALID;
    if (x > 0) k0 = edge_key(c, f[p - 1]);
    if (y > 0) k1 = edge_key(c, f[p - W]);
    if (neighbors8) {
        if (x > 0 && y > 0) k2 = edge_key(c, f[p - W - 1]);
        if (x > 0 && y < H - 1) k3 = edge_key(c, f[p + W - 1]);
    }
    ulonglong2* out = reinterpret_cast<ulonglong2*>(keys + (size_t)frame * key_stride + 4 * (size_t)p);
    out[0] = make_ulonglong2(k0, k1);
    out[1] = make_ulonglong2(k2, k3);
    *reinterpret_cast<uint4*>(prefix + (size_t)frame * key_stride + 4 * (size_t)p) =
        make_uint4(edge_prefix(k0), edge_prefix(k1), edge_prefix(k2), edge_prefix(k3));
}

DOFS_D int edge_other(int s, int d, int W) {
    return d == 0 ? s - 1 : d == 1 ? s - W : d == 2 ? s - W - 1 : s + W - 1;
}

// The segmentation itself never sorts the 4N edge slots: Kruskal only acts on the edges it accepts, and those are
// the N-1 edges of the minimum spanning forest under the strict order (weight, insertion sequence).  Boruvka finds
// them by comparing edges directly (below), and only they are sorted afterwards (k_time_*).  So the per-slot state
// is just the 32-bit prefix; the exact weight of a slot is recomputed from the flow field when two prefixes tie.
__global__ void __launch_bounds__(SEG_THREADS)
k_edge_prefix(const float2* __restrict__ flow, u32* __restrict__ prefix, size_t stride, int W, int H, int neighbors8,
              u32 wm /* fastdiv_magic(W) */) {
    const int frame = blockIdx.y;
    const int N = W * H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    int x;
    const int y = fastdiv(p, W, wm, &x);
    const float2* f = flow + (size_t)frame * N;
    const float2 c = f[p];
    u32 k0 = EDGE_PREFIX_INVALID, k1 = EDGE_PREFIX_INVALID, k2 = EDGE_PREFIX_INVALID, k3 = EDGE_PREFIX_INVALID;
    if (x > 0) k0 = edge_prefix(edge_key(c, f[p - 1]));
    if (y > 0) k1 = edge_prefix(edge_key(c, f[p - W]));
    if (neighbors8) {
        if (x > 0 && y > 0) k2 = edge_prefix(edge_key(c, f[p - W - 1]));
        if (x > 0 && y < H - 1) k3 = edge_prefix(edge_key(c, f[p + W - 1]));
    }
    *reinterpret_cast<uint4*>(prefix + (size_t)frame * stride + 4 * (size_t)p) = make_uint4(k0, k1, k2, k3);
}

// weight (bit pattern of the f64) of an existing slot, from the blurred flow of its frame
DOFS_D u64 slot_weight(const float2* __restrict__ f, u32 slot, int W) {
    const int s = (int)(slot >> 2);
    return edge_key(f[s], f[edge_other(s, (int)(slot & 3u), W)]);
}

// A pick = (prefix << 32) | slot of an existing edge; PICK_NONE (prefix of a non-existent slot) is larger than any pick.
#define PICK_NONE 0xFFFFFFFFFFFFFFFFull
DOFS_D u64 make_pick(u32 prefix, u32 slot) { return ((u64)prefix << 32) | slot; }

__device__ __noinline__ bool pick_less_tie(u32 sa, u32 sb, const float2* __restrict__ f, int W) {
    const u64 wa = slot_weight(f, sa, W), wb = slot_weight(f, sb, W);
    return wa != wb ? wa < wb : sa < sb;
}
// the reference's edge order (graph.cpp:55-60: weight, then insertion sequence) between two picks of one frame
DOFS_D bool pick_less(u64 a, u64 b, const float2* __restrict__ f, int W) {
    const u32 pa = (u32)(a >> 32), pb = (u32)(b >> 32);
    if (pa != pb) return pa < pb;
    if (a == b) return false;
    if (pa == 0u) return (u32)a < (u32)b;  // prefix 0 is the weight 0 exactly (any other weight of a float field is >= 2^-149)
    return pick_less_tie((u32)a, (u32)b, f, W);
}
// lock-free minimum under that order; `seen` is a (possibly stale) read of *best, which only ever decreases
DOFS_D void pick_offer(u64* best, u64 cand, u64 seen, const float2* __restrict__ f, int W) {
    while (pick_less(cand, seen, f, W)) {
        const u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(best), (unsigned long long)seen, (unsigned long long)cand);
        if (old == seen) return;
        seen = old;
    }
}

// rank[seq] = position in the sorted list (INF for the non-existent slots, which sort last); only used
// after the full 64-bit fallback sort
__global__ void __launch_bounds__(SEG_THREADS)
k_rank_scatter(const u32* __restrict__ sorted_seq, u32* __restrict__ rank, size_t stride, int n_slots, int n_edges,
               const int* __restrict__ enable) {
    if (enable && *enable == 0) return;
    const int frame = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        u32 seq = sorted_seq[(size_t)frame * stride + i];
        rank[(size_t)frame * stride + seq] = i < n_edges ? (u32)i : DOFS_INF32;
    }
}

// parity hook: sorted (start, end, weight) from the sorted sequence numbers
__global__ void __launch_bounds__(SEG_THREADS)
k_edges_decode(const u32* __restrict__ sorted_seq, const float2* __restrict__ flow, int* __restrict__ start,
               int* __restrict__ end, u64* __restrict__ weight, int W, int n_edges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    u32 seq = sorted_seq[i];
    int s = (int)(seq >> 2);
    int e = edge_other(s, (int)(seq & 3u), W);
    start[i] = s;
    end[i] = e;
    weight[i] = edge_key(flow[s], flow[e]);
}

// ---------------------------------------------------------------------------------------------
// K8b  exact order inside runs of equal 32-bit prefix.  After the 4-pass sort on the prefix the
// sequence numbers are in (prefix, insertion) order; the reference order is (weight, insertion).
// Runs are short (a handful of edges: two weights must agree to 2^-23 relative to share a prefix):
//   k_prefix_repair_short  every run of at most REPAIR_SHORT edges is insertion-sorted by its head
//                          thread on the 64-bit weights; also writes rank[seq] = final position.
//                          Longer runs are pushed to a list.
//   k_prefix_repair_long   one block per listed run: a run of identical weights (zero-weight ties of a
//                          static scene, flat ramps) is already in order; otherwise it is sorted in
//                          shared memory (<= REPAIR_SMEM edges) or, beyond that, *need_full is raised
//                          and the full 64-bit radix sort that is enqueued behind runs instead.
// ---------------------------------------------------------------------------------------------
#define REPAIR_SHORT 16
#define REPAIR_SMEM 2048

struct RepairArgs {
    const u32* prefix;   // [F][S] sorted prefixes
    u32* seq;            // [F][S] sequence numbers in (prefix, insertion) order -> repaired in place
    const u64* keys;     // [F][S] weights by slot (= by sequence number)
    u32* rank;           // [F][S] out: rank[seq] = position, INF for non-existent slots
    uint2* long_list;    // (frame, start)
    int* long_count;
    int* need_full;
    int list_cap;
    size_t stride;
    int n_slots;
};

__global__ void __launch_bounds__(SEG_THREADS)
k_prefix_repair_short(RepairArgs A) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_slots) return;
    const size_t fo = (size_t)frame * A.stride;
    const u32* pre = A.prefix + fo;
    const u32 k = pre[i];
    if (k == EDGE_PREFIX_INVALID) {  // non-existent border slot
        A.rank[fo + A.seq[fo + i]] = DOFS_INF32;
        return;
    }
    const bool prev_same = i > 0 && pre[i - 1] == k;
    const bool next_same = i + 1 < A.n_slots && pre[i + 1] == k;
    if (!prev_same && !next_same) {  // a run of one
        A.rank[fo + A.seq[fo + i]] = (u32)i;
        return;
    }
    if (prev_same) return;  // the head of the run does the work
    int len = 2;
    while (len <= REPAIR_SHORT && i + len < A.n_slots && pre[i + len] == k) ++len;
    if (len > REPAIR_SHORT) {
        const int slot = atomicAdd(A.long_count, 1);
        if (slot < A.list_cap) A.long_list[slot] = make_uint2((u32)frame, (u32)i);
        else atomicExch(A.need_full, 1);
        return;
    }
    u32 sq[REPAIR_SHORT];
    u64 kk[REPAIR_SHORT];
    for (int j = 0; j < len; ++j) {
        sq[j] = A.seq[fo + i + j];
        kk[j] = A.keys[fo + sq[j]];
    }
    for (int j = 1; j < len; ++j) {  // stable insertion sort by weight
        const u32 s = sq[j];
        const u64 w = kk[j];
        int m = j - 1;
        while (m >= 0 && kk[m] > w) {
            sq[m + 1] = sq[m];
            kk[m + 1] = kk[m];
            --m;
        }
        sq[m + 1] = s;
        kk[m + 1] = w;
    }
    for (int j = 0; j < len; ++j) {
        A.seq[fo + i + j] = sq[j];
        A.rank[fo + sq[j]] = (u32)(i + j);
    }
}

__global__ void __launch_bounds__(256)
k_prefix_repair_long(RepairArgs A) {
    __shared__ u64 s_key[REPAIR_SMEM];
    __shared__ u32 s_seq[REPAIR_SMEM];
    __shared__ int s_flag, s_len;
    const int n_list = min(*A.long_count, A.list_cap);
    for (int item = blockIdx.x; item < n_list; item += gridDim.x) {
        const uint2 it = A.long_list[item];
        const size_t fo = (size_t)it.x * A.stride;
        const int i0 = (int)it.y;
        const u32* pre = A.prefix + fo;
        const u32 k = pre[i0];
        // length of the run
        if (threadIdx.x == 0) s_len = A.n_slots - i0;
        __syncthreads();
        for (int base = 0; base < A.n_slots - i0; base += 256) {
            const int j = base + threadIdx.x;
            if (j < A.n_slots - i0 && pre[i0 + j] != k) atomicMin(&s_len, j);
            __syncthreads();
            const int seen = s_len;
            __syncthreads();
            if (seen <= base + 256) break;
        }
        const int len = s_len;
        // identical weights?
        const u64 w0 = A.keys[fo + A.seq[fo + i0]];
        if (threadIdx.x == 0) s_flag = 0;
        __syncthreads();
        int differs = 0;
        for (int j = threadIdx.x; j < len; j += 256) differs |= A.keys[fo + A.seq[fo + i0 + j]] != w0;
        if (differs) s_flag = 1;
        __syncthreads();
        const bool trivial = s_flag == 0;
        __syncthreads();
        if (!trivial && len > REPAIR_SMEM) {
            if (threadIdx.x == 0) atomicExch(A.need_full, 1);
            continue;  // block-uniform
        }
        if (!trivial) {
            // odd-even transposition sort in shared memory: stable, len rounds of disjoint compare-exchanges
            for (int j = threadIdx.x; j < len; j += 256) {
                s_seq[j] = A.seq[fo + i0 + j];
                s_key[j] = A.keys[fo + s_seq[j]];
            }
            __syncthreads();
            for (int round = 0; round < len; ++round) {
                for (int j = 2 * threadIdx.x + (round & 1); j + 1 < len; j += 512) {
                    if (s_key[j] > s_key[j + 1]) {
                        const u64 tk = s_key[j];
                        s_key[j] = s_key[j + 1];
                        s_key[j + 1] = tk;
                        const u32 ts = s_seq[j];
                        s_seq[j] = s_seq[j + 1];
                        s_seq[j + 1] = ts;
                    }
                }
                __syncthreads();
            }
            for (int j = threadIdx.x; j < len; j += 256) A.seq[fo + i0 + j] = s_seq[j];
            __syncthreads();
        }
        for (int j = threadIdx.x; j < len; j += 256) A.rank[fo + A.seq[fo + i0 + j]] = (u32)(i0 + j);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K9a  Boruvka levels.  No host round trip: the host enqueues the guaranteed bound of levels
// (components at least halve per level) and every kernel of a level returns at once for a frame that
// is already one component.  Work shrinks with the forest:
//   * the pixel kernel skips pixels whose incident edges have all become internal (one byte of mask each:
//     an edge that became internal stays internal); from level 2 on it first packs the live pixels of a
//     chunk so that every lane has work (k_bor_pixel_packed);
//   * the root kernels run over the list of live roots, rebuilt every level (k_bor_contract chases the
//     hooks of four roots per thread in lockstep and appends the survivors with one atomic per block);
//   * `comp` (root of every pixel) is written by the level-0 contraction and refreshed by one streaming
//     pass per later level: one hop through the `up` link its root received in the contraction, for the
//     pixels that still have an external edge (nobody reads the others again).
// Kernels are grid-stride with a small fixed grid, so a level with nothing left costs microseconds.
// (The late levels sit behind a CUDA-graph IF node, see launch_conditional in dofs3d.cu.)
// ---------------------------------------------------------------------------------------------
#define EV_MAX_WAVES 32
// Measured knobs of the level kernels (both on; 32 pairs alone: Boruvka 9.13 -> 8.45 ms, bench 1 033 -> 1 058 pairs/s):
//   BOR_L0_COMP       the level-0 contraction writes `comp` itself (every pixel is a root there): no level-0 relabel pass
//   BOR_RELABEL_MASK  the relabel pass skips pixels whose edges have all become internal (four mask bytes per load)
#ifndef BOR_L0_COMP
#define BOR_L0_COMP 1
#endif
#ifndef BOR_CHASE_LOCKSTEP
#define BOR_CHASE_LOCKSTEP 1
#endif
#ifndef BOR_RELABEL_MASK
#define BOR_RELABEL_MASK 1
#endif
#ifndef BOR_PIXEL_BLOCKS
#define BOR_PIXEL_BLOCKS 8  // resident blocks per SM the pixel kernel is compiled for: it is bound by memory latency, so full
                            // occupancy (32 registers, a few spilled words) beats a spill-free 48-register build by 10 %
#endif

// Per-root state of Forest::merge (size, mean flow, bounding box) as one 32-byte record = one DRAM sector: the replay
// gathers the state of absorbed roots at random, and three separate arrays cost three sectors per gather.
struct __align__(32) RootState {
    int size;
    float fx, fy;
    u32 pad0;
    ushort4 bbox;  // xmin, ymin, xmax, ymax
    u32 pad1, pad2;
};
DOFS_D RootState root_load(const RootState* p) {
    const uint4 a = reinterpret_cast<const uint4*>(p)[0];
    const uint2 b = reinterpret_cast<const uint2*>(p)[2];
    RootState r;
    r.size = (int)a.x;
    r.fx = __uint_as_float(a.y);
    r.fy = __uint_as_float(a.z);
    r.pad0 = 0;
    r.bbox = make_ushort4((u16)(b.x & 0xFFFFu), (u16)(b.x >> 16), (u16)(b.y & 0xFFFFu), (u16)(b.y >> 16));
    r.pad1 = r.pad2 = 0;
    return r;
}
DOFS_D void root_store(RootState* p, int size, float2 f, ushort4 bb) {
    reinterpret_cast<uint4*>(p)[0] = make_uint4((u32)size, __float_as_uint(f.x), __float_as_uint(f.y), 0u);
    reinterpret_cast<uint4*>(p)[1] = make_uint4((u32)bb.x | ((u32)bb.y << 16), (u32)bb.z | ((u32)bb.w << 16), 0u, 0u);
}

struct BorState {
    u32* comp;       // [F][N] current root of each pixel
    u64* best;       // [F][N] per root: its minimum outgoing edge in this level, as a pick (prefix << 32 | slot)
    u32* newp;       // [F][N] per root: hook target in this level
    u32* loss_time;  // [F][N] per root id: the slot of the edge at which it loses (INF: never loses).  The kernels after
                     //         K8 get a copy of this struct whose loss_time is the time array instead: the position of that
                     //         edge in the reference's merge sequence
    u32* up;         // [F][N] per root id: root of the next-level component it is contracted into (itself while live)
    u8* lvl;         // [F][N] per root id: level at which it loses == its final union-find rank
    u8* mask;        // [F][N] per pixel: which of its 8 incident edges still join different components
    u32* roots[2];   // [F][N] live roots, ping-pong by level parity (level 0: every pixel, implicit)
    int* n_roots;    // [EV_MAX_WAVES][F] number of roots after each level (zeroed per call)
    int* levels;     // [F] number of levels the frame needed (written by k_bor_finish)
    int* final_root; // [F]
    int F;
};

DOFS_D bool bor_done(const BorState& S, int level, int frame) {
    return level > 0 && S.n_roots[(level - 1) * S.F + frame] == 1;
}

#define GRID_STRIDE(p, N) for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (N); p += gridDim.x * blockDim.x)

// The per-pixel Boruvka kernels read the rows above and below a pixel (neighbour components, neighbour edge slots).
// BOR_TILE_W > 0: a block visits BOR_TILE_W x (SEG_THREADS / BOR_TILE_W) tiles of the image instead of 256 consecutive
// pixels of a row, so the neighbour rows are its own lines in L1.  `continue` in the body moves on to the next tile.
#ifndef BOR_TILE_W
#define BOR_TILE_W 0
#endif
#if BOR_TILE_W
#define BOR_TILE_H (SEG_THREADS / BOR_TILE_W)
#define PIXEL_TILES(p, W, H)                                                                                              \
    for (int t_ = blockIdx.x, tnx_ = ((W) + BOR_TILE_W - 1) / BOR_TILE_W, tn_ = tnx_ * (((H) + BOR_TILE_H - 1) / BOR_TILE_H); \
         t_ < tn_; t_ += gridDim.x)                                                                                       \
        for (int px_ = (t_ % tnx_) * BOR_TILE_W + (int)(threadIdx.x % BOR_TILE_W),                                          \
                 py_ = (t_ / tnx_) * BOR_TILE_H + (int)(threadIdx.x / BOR_TILE_W), p = py_ * (W) + px_, once_ = 1;         \
             once_; once_ = 0)                                                                                            \
            if (px_ < (W) && py_ < (H))
#else
#define PIXEL_TILES(p, W, H) for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (W) * (H); p += gridDim.x * blockDim.x)
#endif

// Incident edge e of pixel p: e < 4 its own back-edge slots 4p+e; e = 4..7 the back-edges of the right, lower-right,
// upper-right and lower neighbours that point at p.
DOFS_D int incident_pixel(int p, int e, int W) {
    return e < 4 ? edge_other(p, e, W) : e == 4 ? p + 1 : e == 5 ? p + W + 1 : e == 6 ? p - W + 1 : p + W;
}
DOFS_D u32 incident_slot(int p, int e, int W) {
    return e < 4 ? 4u * (u32)p + e
                 : e == 4 ? 4u * (u32)(p + 1) : e == 5 ? 4u * (u32)(p + W + 1) + 2u : e == 6 ? 4u * (u32)(p - W + 1) + 3u
                                                                                              : 4u * (u32)(p + W) + 1u;
}
// which of the eight incident edges of (x, y) exist
DOFS_D u32 incident_mask(int x, int y, int W, int H, int neighbors8) {
    const bool l = x > 0, r = x + 1 < W, u = y > 0, d = y + 1 < H;
    u32 m = (l ? 1u : 0u) | (u ? 2u : 0u) | (r ? 16u : 0u) | (d ? 128u : 0u);
    if (neighbors8) m |= (l && u ? 4u : 0u) | (l && d ? 8u : 0u) | (r && d ? 32u : 0u) | (r && u ? 64u : 0u);
    return m;
}


// append to a per-frame list with ONE atomic per block and BOR_APPEND_ROUNDS items per thread (all threads of the
// block must call it; bit r of `want` selects value[r]).  One atomic per warp is not enough here: a level-0 contraction
// appends from every warp of a frame to the same counter, and same-address atomics that return a value serialise in
// L2.  Batching the rounds divides the three barriers and the atomic by the batch.
#ifndef BOR_APPEND_ROUNDS
#define BOR_APPEND_ROUNDS 4
#endif
DOFS_D void list_append_block(u32* list, int* counter, u32 want, const u32 (&value)[BOR_APPEND_ROUNDS]) {
    __shared__ int s_count[SEG_THREADS / 32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int mine = __popc(want);
    int incl = mine;  // inclusive scan of the per-thread counts inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_count[wrp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < SEG_THREADS / 32; ++w) {
            const int c = s_count[w];
            s_count[w] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    int pos = s_base + s_count[wrp] + incl - mine;
#pragma unroll
    for (int r = 0; r < BOR_APPEND_ROUNDS; ++r)
        if ((want >> r) & 1u) list[pos++] = value[r];
    __syncthreads();
}

// Every pixel offers the smallest of its (up to eight) incident edges that leave its component to ITS OWN component:
// one atomic per boundary pixel, and neighbouring pixels mostly address the same word.  Pixels are visited in image
// order (neighbouring component ids share cache lines); a pixel whose incident edges have all become internal costs
// one byte of mask.  Incident edge e of pixel p: e < 4 its own back-edge slots 4p+e; e = 4..7 the back-edges of the
// right, lower-right, upper-right and lower neighbours that point at p.
//
// The minimum is taken with one native 64-bit atomicMin per offer on (prefix << 32 | slot), i.e. in (prefix, slot)
// order L.  That is the reference's (weight, slot) order E unless two offers to the same component share a prefix with
// different weights (prefix 0 is the weight 0 exactly, so there L == E).  An offer that meets another offer x of its
// own prefix — in the value it read or in the value its atomicMin returned — settles the pair under E with a CAS loop
// (pick_offer): it offers itself, and it offers x again if its atomicMin displaced x.  The E-minimum m of a component
// is therefore installed when its own offer completes (either it is L-smaller than what is there, or it sees its
// prefix and installs itself under E), nothing can remove it under E, and whoever removes it under L puts it back.
// A pixel whose own candidates share the prefix of its L-smallest one offers them all under E.
#define PIX_SENT 1u       // the atomicMin was issued
#define PIX_TIE_SEEN 2u   // the offer met another offer of its prefix
#define PIX_TIE_LOCAL 4u  // two of the pixel's own candidates share the smallest prefix

__device__ __noinline__ void bor_pixel_settle(u64* best, const u32* __restrict__ pre, const float2* __restrict__ f, int W, int p,
                                              u32 cp, u32 out, u64 mine, u64 seen, u32 flags) {
    u64* b = &best[cp];
    if (flags & PIX_TIE_SEEN) {
        pick_offer(b, mine, *reinterpret_cast<volatile u64*>(b), f, W);
        if ((flags & PIX_SENT) && mine < seen) pick_offer(b, seen, *reinterpret_cast<volatile u64*>(b), f, W);
    }
    if (flags & PIX_TIE_LOCAL) {
#pragma unroll 1
        for (int e = 0; e < 8; ++e) {
            if (!((out >> e) & 1u)) continue;
            const u32 slot = incident_slot(p, e, W);
            if (pre[slot] != EDGE_PREFIX_INVALID) pick_offer(b, make_pick(pre[slot], slot), *reinterpret_cast<volatile u64*>(b), f, W);
        }
    }
}

// an offer `cand` met `seen` in best: same prefix, another edge, and not the exact weight 0
DOFS_D bool pick_meets(u64 cand, u64 seen) {
    return (u32)(seen >> 32) == (u32)(cand >> 32) && seen != cand && (u32)(cand >> 32) != 0u;
}

// one pixel of a level (see above); `m` = its edge mask, known to be non-zero
DOFS_D void bor_pixel_one(const BorState& S, size_t fo, u32* comp, const u32* up, const u32* __restrict__ pre, u64* best,
                          const float2* __restrict__ flow, int W, int p, u32 m, int fold) {
    // fold != 0 (A/B knob, DOFS3D_BOR_FOLD=1): there is no separate relabel pass after level 0 and `comp` of a live
    // pixel is one contraction behind: the contraction left the current root of every root of the previous level in
    // `up` (a survivor points at itself), so one hop refreshes it.  A neighbour's entry may already have been
    // refreshed by its own thread: one hop from a current root is the root itself.  Pixels whose edges are all
    // internal are never read again and stay stale.  Measured on the B200 (32 pairs, alone): the extra random
    // gathers cost more than the streaming relabel pass they replace (Boruvka 9.4 -> 11.2 ms), so it is off.
    // (three row pointers and compile-time column offsets: the eight neighbour reads and the four neighbour-slot reads
    //  share their address arithmetic; the minimum is taken without branches.  Line-level ncu counts had put half of this
    //  kernel's instructions in these two loops.)
    const u32* cc = comp + p;
    const u32* cu = cc - W;
    const u32* cd = cc + W;
    const u32 c0 = cc[0];
    const u32 cp = fold ? up[c0] : c0;
    if (cp != c0) comp[p] = cp;
    const u64 seen0 = best[cp];
    // independent loads first (the kernel is bound by memory latency): neighbour components, then their prefixes
    // incident pixel of edge e: 0 left, 1 up, 2 up-left, 3 down-left, 4 right, 5 down-right, 6 up-right, 7 down
    u32 q[8];
    q[0] = (m & 1u) ? cc[-1] : c0;
    q[1] = (m & 2u) ? cu[0] : c0;
    q[2] = (m & 4u) ? cu[-1] : c0;
    q[3] = (m & 8u) ? cd[-1] : c0;
    q[4] = (m & 16u) ? cc[1] : c0;
    q[5] = (m & 32u) ? cd[1] : c0;
    q[6] = (m & 64u) ? cu[1] : c0;
    q[7] = (m & 128u) ? cd[0] : c0;
    u32 out = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (q[e] != c0 && (!fold || up[q[e]] != cp)) out |= 1u << e;  // equal stale roots are equal current roots
    if (out != m) S.mask[fo + p] = (u8)out;  // an edge that became internal stays internal
    if (out == 0) return;
    // prefixes of the outgoing edges (EDGE_PREFIX_INVALID = the largest value: a NaN / infinite weight is no edge, and
    // neither is an edge that is not outgoing)
    const u32* pp = pre + 4 * (size_t)p;
    const u32* pu = pp - 4 * W;
    const u32* pd = pp + 4 * W;
    u32 pr[8];
    {
        uint4 r4 = make_uint4(EDGE_PREFIX_INVALID, EDGE_PREFIX_INVALID, EDGE_PREFIX_INVALID, EDGE_PREFIX_INVALID);
        if (out & 15u) r4 = *reinterpret_cast<const uint4*>(pp);
        pr[0] = (out & 1u) ? r4.x : EDGE_PREFIX_INVALID;
        pr[1] = (out & 2u) ? r4.y : EDGE_PREFIX_INVALID;
        pr[2] = (out & 4u) ? r4.z : EDGE_PREFIX_INVALID;
        pr[3] = (out & 8u) ? r4.w : EDGE_PREFIX_INVALID;
    }
    pr[4] = (out & 16u) ? pp[4] : EDGE_PREFIX_INVALID;       // right neighbour's left edge
    pr[5] = (out & 32u) ? pd[4 + 2] : EDGE_PREFIX_INVALID;   // down-right: its up-left edge
    pr[6] = (out & 64u) ? pu[4 + 3] : EDGE_PREFIX_INVALID;   // up-right: its down-left edge
    pr[7] = (out & 128u) ? pd[1] : EDGE_PREFIX_INVALID;      // lower neighbour's up edge
    // smallest (prefix, slot) among the outgoing edges; ties on the prefix go to the smallest slot, i.e. the first in
    // ascending slot order 6, 0, 1, 2, 3, 4, 7, 5
    const u32 pmin = min(min(min(pr[0], pr[1]), min(pr[2], pr[3])), min(min(pr[4], pr[5]), min(pr[6], pr[7])));
    if (pmin == EDGE_PREFIX_INVALID) return;
    int emin = 5;
    u32 same = 0;  // how many other candidates share the smallest prefix
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        constexpr int order[8] = {6, 0, 1, 2, 3, 4, 7, 5};
        const int e = order[k];
        const bool hit = pr[e] == pmin;
        emin = hit ? e : emin;
        same += hit ? 1u : 0u;
    }
    same -= 1u;
    const u32 smin = incident_slot(p, emin, W);
    const u64 mine = make_pick(pmin, smin);
    u32 flags = (same != 0 && pmin != 0u) ? PIX_TIE_LOCAL : 0u;
    u64 seen = seen0;
    if (mine < seen) {
        seen = atomicMin(reinterpret_cast<unsigned long long*>(&best[cp]), (unsigned long long)mine);
        flags |= PIX_SENT;
    }
    if (pick_meets(mine, seen)) flags |= PIX_TIE_SEEN;
    if (flags & (PIX_TIE_SEEN | PIX_TIE_LOCAL)) bor_pixel_settle(best, pre, flow + fo, W, p, cp, out, mine, seen, flags);
}

// BOR_COMPACT_FROM: from this level on (k_bor_pixel_packed) a block first packs the pixels of a 1024-pixel chunk that
// still have an external edge into shared memory (four mask bytes per thread) and then visits only those, so every
// lane of a warp has work; below it (most pixels still live) every thread visits its own pixel (k_bor_pixel).
// Measured alone, 32 pairs: Boruvka 8.40 -> 8.07 ms from level 1, 2 or 4 alike.
#ifndef BOR_COMPACT_FROM
#define BOR_COMPACT_FROM 2
#endif
#define BOR_CHUNK (4 * SEG_THREADS)
__global__ void __launch_bounds__(SEG_THREADS, BOR_PIXEL_BLOCKS)
k_bor_pixel(BorState S, const u32* __restrict__ prefix, size_t prefix_stride, const float2* __restrict__ flow, int W, int N,
            int level, int fold) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const size_t fo = (size_t)frame * N;
    u32* comp = S.comp + fo;
    const u32* up = S.up + fo;
    const u32* pre = prefix + (size_t)frame * prefix_stride;
    u64* best = S.best + fo;
    PIXEL_TILES(p, W, N / W) {
        const u32 m = S.mask[fo + p];
        if (m == 0) continue;
        bor_pixel_one(S, fo, comp, up, pre, best, flow, W, p, m, fold);
    }
}

// the same level step for N % 4 == 0, N <= 2^24 (bor_pixel_packed_ok): packs, then visits
DOFS_HD bool bor_pixel_packed_ok(int N) { return (N & 3) == 0 && N <= (1 << 24); }
__global__ void __launch_bounds__(SEG_THREADS, BOR_PIXEL_BLOCKS)
k_bor_pixel_packed(BorState S, const u32* __restrict__ prefix, size_t prefix_stride, const float2* __restrict__ flow, int W,
                   int N, int level, int fold) {
    __shared__ u32 s_list[BOR_CHUNK];
    __shared__ int s_count[SEG_THREADS / 32 + 1];
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const size_t fo = (size_t)frame * N;
    u32* comp = S.comp + fo;
    const u32* up = S.up + fo;
    const u32* pre = prefix + (size_t)frame * prefix_stride;
    u64* best = S.best + fo;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const u32* mask4 = reinterpret_cast<const u32*>(S.mask + fo);
    for (int base = blockIdx.x * BOR_CHUNK; base < N; base += gridDim.x * BOR_CHUNK) {
        const int p4 = base + 4 * (int)threadIdx.x;
        const u32 m4 = p4 < N ? mask4[p4 >> 2] : 0u;
        const int mine = ((m4 & 0xFFu) ? 1 : 0) + ((m4 & 0xFF00u) ? 1 : 0) + ((m4 & 0xFF0000u) ? 1 : 0) + ((m4 >> 24) ? 1 : 0);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_count[wrp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < SEG_THREADS / 32; ++w) {
                const int c = s_count[w];
                s_count[w] = tot;
                tot += c;
            }
            s_count[SEG_THREADS / 32] = tot;
        }
        __syncthreads();
        int pos = s_count[wrp] + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 mk = (m4 >> (8 * k)) & 0xFFu;
            if (mk) s_list[pos++] = (u32)(p4 + k) | (mk << 24);
        }
        __syncthreads();
        const int total = s_count[SEG_THREADS / 32];
        for (int j = threadIdx.x; j < total; j += SEG_THREADS) {
            const u32 e = s_list[j];
            bor_pixel_one(S, fo, comp, up, pre, best, flow, W, (int)(e & 0xFFFFFFu), e >> 24, fold);
        }
        __syncthreads();
    }
}

// Level 0 without atomics: every pixel is its own component, so its minimum outgoing edge is the minimum among its
// (up to) eight incident edges — its four back-edges and the back-edges of the four neighbours that point at it.
// k_bor_level0_pick stores that pick in `best`; k_bor_level0_root applies the same mutual-pick rule as k_bor_root (the
// edge's `end` side survives a tie; the pixel owning the slot is `start`).
#ifndef BOR_LEVEL0_BLOCKS
#define BOR_LEVEL0_BLOCKS 6
#endif
__global__ void __launch_bounds__(SEG_THREADS, BOR_LEVEL0_BLOCKS)
k_bor_level0_pick(BorState S, const u32* __restrict__ prefix, size_t prefix_stride, const float2* __restrict__ flow, int W,
                  int H, int N) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    const u32* pre = prefix + (size_t)frame * prefix_stride;
    const float2* f = flow + fo;
    const u32 wm = fastdiv_magic((u32)W);  // (once per thread: the loop below visits about a hundred pixels)
    PIXEL_TILES(p, W, H) {
        int x;
        const int y = fastdiv(p, W, wm, &x);
        const uint4 r4 = *reinterpret_cast<const uint4*>(pre + 4 * (size_t)p);
        u64 b = PICK_NONE;
#define L0_CONSIDER(pr, slot)                                     \
    do {                                                          \
        const u32 pr_ = (pr);                                     \
        if (pr_ != EDGE_PREFIX_INVALID) {                         \
            const u64 c_ = make_pick(pr_, (u32)(slot));           \
            if (pick_less(c_, b, f, W)) b = c_;                   \
        }                                                         \
    } while (0)
        L0_CONSIDER(r4.x, 4 * p);
        L0_CONSIDER(r4.y, 4 * p + 1);
        L0_CONSIDER(r4.z, 4 * p + 2);
        L0_CONSIDER(r4.w, 4 * p + 3);
        if (x + 1 < W) {
            L0_CONSIDER(pre[4 * (size_t)(p + 1)], 4 * (p + 1));                                   // right neighbour's left edge
            if (y + 1 < H) L0_CONSIDER(pre[4 * (size_t)(p + W + 1) + 2], 4 * (p + W + 1) + 2);    // down-right: its up-left edge
            if (y > 0) L0_CONSIDER(pre[4 * (size_t)(p - W + 1) + 3], 4 * (p - W + 1) + 3);        // up-right: its down-left edge
        }
        if (y + 1 < H) L0_CONSIDER(pre[4 * (size_t)(p + W) + 1], 4 * (p + W) + 1);                // lower neighbour's up edge
#undef L0_CONSIDER
        S.best[fo + p] = b;
    }
}

// Also the only initialisation the per-root arrays get: every pixel is a root here, so hook target, loss slot, rank
// and edge mask are written for all of them (no separate init pass); `up` and `comp` follow in the level-0 contraction
// and relabel.
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_level0_root(BorState S, int W, int H, int N, int neighbors8) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    const u32 wm = fastdiv_magic((u32)W);
    GRID_STRIDE(p, N) {
        int x;
        const int y = fastdiv(p, W, wm, &x);
        S.mask[fo + p] = (u8)incident_mask(x, y, W, H, neighbors8);
        S.lvl[fo + p] = 0;
        const u64 t = S.best[fo + p];
        u32 hook = (u32)p, loss = DOFS_INF32;
        if (t != PICK_NONE) {  // (PICK_NONE: a pixel without a finite-weight edge)
            const u32 slot = (u32)t;
            const int s = (int)(slot >> 2), e = edge_other(s, (int)(slot & 3u), W);
            const int q = s == p ? e : s;
            const bool mutual = S.best[fo + q] == t;
            if (!(mutual && p == e)) {  // rank tie: the `end` side survives (graph.cpp:177-182, 210-213)
                hook = (u32)q;
                loss = slot;
            }
        }
        S.newp[fo + p] = hook;
        S.loss_time[fo + p] = loss;
    }
}

// per live root: classify its pick (mutual winner / loser), record the loss
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_root(BorState S, int W, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const size_t fo = (size_t)frame * N;
    const u32* list = S.roots[level & 1] + fo;
    const int count = level == 0 ? N : S.n_roots[(level - 1) * S.F + frame];
    GRID_STRIDE(i, count) {
        const u32 c = level == 0 ? (u32)i : list[i];
        const u64 t = S.best[fo + c];
        if (t == PICK_NONE) {  // the last component
            S.newp[fo + c] = c;
            continue;
        }
        const u32 slot = (u32)t;
        const int s = (int)(slot >> 2);
        const int e = edge_other(s, (int)(slot & 3u), W);
        const u32 cs = S.comp[fo + s], ce = S.comp[fo + e];
        const u32 other = (cs == c) ? ce : cs;
        const bool mutual = S.best[fo + other] == t;
        if (mutual && ce == c) {
            S.newp[fo + c] = c;  // rank tie: the `end` side survives (graph.cpp:177-182, 210-213)
        } else {
            S.newp[fo + c] = other;
            S.loss_time[fo + c] = slot;
            S.lvl[fo + c] = (u8)level;
        }
    }
}

// contract: every live root follows the hooks to its group root; losers link to it in `up`, survivors form the next list
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_contract(BorState S, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) S.n_roots[level * S.F + frame] = 1;
        return;
    }
    const size_t fo = (size_t)frame * N;
    volatile u32* newp = S.newp + fo;
    const u32* list = S.roots[level & 1] + fo;
    u32* next = S.roots[(level + 1) & 1] + fo;
    const int count = level == 0 ? N : S.n_roots[(level - 1) * S.F + frame];
    const int rounds = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (int it0 = 0; it0 < rounds; it0 += BOR_APPEND_ROUNDS) {
        u32 want = 0;
        u32 live[BOR_APPEND_ROUNDS];
#if BOR_CHASE_LOCKSTEP
        // The chases of the thread's BOR_APPEND_ROUNDS roots advance in lockstep: every trip of the loop issues the next
        // link of ALL unfinished chases before any is consumed, so a thread has that many loads in flight instead of one
        // (the kernel is bound by the latency of these dependent loads).  Path halving as below.
        u32 g[BOR_APPEND_ROUNDS];
        u32 busy = 0;
#pragma unroll
        for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) {
            const int i = ((it0 + r) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
            live[r] = 0;
            g[r] = 0;
            if (it0 + r < rounds && i < count) {
                live[r] = level == 0 ? (u32)i : list[i];
                g[r] = live[r];
                busy |= 1u << r;
            }
        }
        const u32 valid = busy;
        while (busy) {
            u32 nx[BOR_APPEND_ROUNDS], nn[BOR_APPEND_ROUNDS];
#pragma unroll
            for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) nx[r] = ((busy >> r) & 1u) ? newp[g[r]] : 0u;
#pragma unroll
            for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) nn[r] = ((busy >> r) & 1u) ? newp[nx[r]] : 0u;
#pragma unroll
            for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) {
                if (!((busy >> r) & 1u)) continue;
                if (nx[r] == g[r]) {
                    busy &= ~(1u << r);
                } else if (nn[r] == nx[r]) {
                    g[r] = nx[r];
                    busy &= ~(1u << r);
                } else {
                    newp[g[r]] = nn[r];
                    g[r] = nn[r];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) {
            if (!((valid >> r) & 1u)) continue;
            const u32 c = live[r];
            const bool survives = g[r] == c;
            if (survives) S.best[fo + c] = PICK_NONE;  // only live roots collect offers in the next level
            if (!survives || level == 0) S.up[fo + c] = g[r];  // (level 0 initialises `up`: a survivor points at itself)
#if BOR_L0_COMP
            if (level == 0) S.comp[fo + c] = g[r];  // every pixel is a root here: its component, without a relabel pass
#endif
            if (survives) want |= 1u << r;
        }
#else
#pragma unroll
        for (int r = 0; r < BOR_APPEND_ROUNDS; ++r) {
            const int i = ((it0 + r) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
            live[r] = 0;
            if (it0 + r < rounds && i < count) {
                const u32 c = level == 0 ? (u32)i : list[i];
                // follow the hooks with path halving: concurrent writers only ever replace a pointer by one of its
                // ancestors, so any interleaving still ends at the same root
                u32 g = c;
                for (;;) {
                    const u32 nx = newp[g];
                    if (nx == g) break;
                    const u32 nn = newp[nx];
                    if (nn == nx) {
                        g = nx;
                        break;
                    }
                    newp[g] = nn;
                    g = nn;
                }
                const bool survives = g == c;
                if (survives) S.best[fo + c] = PICK_NONE;  // only live roots collect offers in the next level
                if (!survives || level == 0) S.up[fo + c] = g;  // (level 0 initialises `up`: a survivor points at itself)
#if BOR_L0_COMP
                if (level == 0) S.comp[fo + c] = g;  // every pixel is a root here: its component, without a relabel pass
#endif
                live[r] = c;
                if (survives) want |= 1u << r;
            }
        }
#endif
        list_append_block(next, &S.n_roots[level * S.F + frame], want, live);
    }
}

// every pixel follows its root's new link (one hop: the contraction stored the group root itself).  With the fold knob
// it is only launched after level 0, where it is the first write of `comp`.
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_relabel(BorState S, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const size_t fo = (size_t)frame * N;
    if (level == 0) {  // every pixel is its own component: this is where `comp` is first written
        GRID_STRIDE(p, N) S.comp[fo + p] = S.up[fo + p];
        return;
    }
#if BOR_RELABEL_MASK
    // a pixel whose incident edges are all internal is never looked at again (k_bor_pixel skips it and its neighbours
    // reach it only through an edge that is still external, which it would see too): its entry may go stale
    if ((N & 3) == 0) {
        const u32* mask4 = reinterpret_cast<const u32*>(S.mask + fo);
        uint4* comp4 = reinterpret_cast<uint4*>(S.comp + fo);
        GRID_STRIDE(i, N / 4) {
            const u32 m4 = mask4[i];
            if (m4 == 0) continue;
            uint4 c = comp4[i];
            const uint4 c0 = c;
            if (m4 & 0x000000FFu) c.x = S.up[fo + c.x];
            if (m4 & 0x0000FF00u) c.y = S.up[fo + c.y];
            if (m4 & 0x00FF0000u) c.z = S.up[fo + c.z];
            if (m4 & 0xFF000000u) c.w = S.up[fo + c.w];
            if (c.x != c0.x || c.y != c0.y || c.z != c0.z || c.w != c0.w) comp4[i] = c;
        }
        return;
    }
#endif
    GRID_STRIDE(p, N) {
        const u32 c = S.comp[fo + p];
        const u32 g = S.up[fo + c];
        if (g != c) S.comp[fo + p] = g;
    }
}

// final root: it loses never; its chain is replayed in the last wave
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_finish(BorState S, int N, int max_levels) {
    const int frame = blockIdx.y;
    int levels = max_levels;
    for (int l = 0; l < max_levels; ++l)
        if (S.n_roots[l * S.F + frame] == 1) {
            levels = l + 1;
            break;
        }
    GRID_STRIDE(p, N) {
        const size_t g = (size_t)frame * N + p;
        if (S.loss_time[g] == DOFS_INF32) {
            S.lvl[g] = (u8)max_levels;  // every frame's last chain runs in the same (last) wave, side by side
            S.final_root[frame] = p;    // if the frame did not converge several pixels land here; n_roots says so
            S.levels[frame] = levels;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K8  merge times.  The reference's loop accepts exactly the edges Boruvka picked (the minimum spanning forest
// under (weight, sequence)); their relative order in the sorted edge list is all that the merge sequence depends
// on.  So only those <= N-1 edges are sorted, on ONE 32-bit key per losing root:
//     weight == 0 (prefix 0)  ->  key = slot            (exact order among them; below every other key, see TIME_KEY_MIN_PREFIX)
//     otherwise               ->  key = prefix
// Four 8-bit passes (payload = the root).  Keys below TIME_KEY_MIN_PREFIX are unique, so their position is their
// time.  Runs of an equal prefix are rare and short; they are put in exact (weight, slot) order:
//   k_time_keys           one key per root id (roots that never lose: TIME_KEY_NONE, sorted last)
//   k_time_repair_short   runs of <= REPAIR_SHORT edges: insertion sort in the head thread
//   k_time_repair_long    longer runs, a block each: <= REPAIR_SMEM edges are sorted in shared memory; beyond that
//                         *need_full enables the exact fallback: a stable 64-bit radix sort by slot, then by weight
//                         (k_time_fallback_slots / k_time_fallback_weights / k_time_fallback_rank)
// time[c] = position = index of the merge in the reference's sequence of accepted edges.
// ---------------------------------------------------------------------------------------------
#define TIME_KEY_NONE 0xFFFFFFFFu
// the smallest prefix of a non-zero weight: a float flow field has no difference below 2^-149, i.e. (200 - 149) << 23
#define TIME_KEY_MIN_PREFIX 0x19800000u

__global__ void __launch_bounds__(SEG_THREADS)
k_time_keys(BorState S, const u32* __restrict__ prefix, size_t prefix_stride, u32* __restrict__ tkey, int N) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(c, N) {
        const u32 slot = S.loss_time[fo + c];
        u32 k = TIME_KEY_NONE;
        if (slot != DOFS_INF32) {
            const u32 pr = prefix[(size_t)frame * prefix_stride + slot];
            k = pr == 0u ? slot : pr;
        }
        tkey[fo + c] = k;
    }
}

struct TimeRepairArgs {
    const u32* key;      // [F][N] sorted keys
    u32* comp;           // [F][N] sorted payload (losing roots); runs that share a prefix are put in exact order in place
                         //         (the event sort below is a stable sort of this list)
    const u32* slot;     // [F][N] by root id: the slot of the edge at which it loses
    const float2* flow;  // [F][N] blurred flow
    u32* time;           // [F][N] out: time[c] = position, INF for roots that never lose
    uint2* long_list;    // (frame, start)
    int* long_count;
    int* need_full;
    int list_cap;
    int N, W;
};

// (weight, slot) order of two edges that share a prefix
DOFS_D bool run_less(u64 wa, u32 sa, u64 wb, u32 sb) { return wa != wb ? wa < wb : sa < sb; }

__global__ void __launch_bounds__(SEG_THREADS)
k_time_repair_short(TimeRepairArgs A) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.N) return;
    const size_t fo = (size_t)frame * A.N;
    const u32* key = A.key + fo;
    const u32 k = key[i];
    if (k == TIME_KEY_NONE) {
        A.time[fo + A.comp[fo + i]] = DOFS_INF32;
        return;
    }
    const bool prev_same = k >= TIME_KEY_MIN_PREFIX && i > 0 && key[i - 1] == k;
    const bool next_same = k >= TIME_KEY_MIN_PREFIX && i + 1 < A.N && key[i + 1] == k;
    if (!prev_same && !next_same) {
        A.time[fo + A.comp[fo + i]] = (u32)i;
        return;
    }
    if (prev_same) return;  // the head of the run does the work
    int len = 2;
    while (len <= REPAIR_SHORT && i + len < A.N && key[i + len] == k) ++len;
    if (len > REPAIR_SHORT) {
        const int slot = atomicAdd(A.long_count, 1);
        if (slot < A.list_cap) A.long_list[slot] = make_uint2((u32)frame, (u32)i);
        else atomicExch(A.need_full, 1);
        return;
    }
    u32 cc[REPAIR_SHORT], ss[REPAIR_SHORT];
    u64 kk[REPAIR_SHORT];
    for (int j = 0; j < len; ++j) {
        cc[j] = A.comp[fo + i + j];
        ss[j] = A.slot[fo + cc[j]];
        kk[j] = slot_weight(A.flow + fo, ss[j], A.W);
    }
    for (int j = 1; j < len; ++j) {  // insertion sort by (weight, slot)
        const u32 c = cc[j], sl = ss[j];
        const u64 w = kk[j];
        int m = j - 1;
        while (m >= 0 && run_less(w, sl, kk[m], ss[m])) {
            cc[m + 1] = cc[m];
            ss[m + 1] = ss[m];
            kk[m + 1] = kk[m];
            --m;
        }
        cc[m + 1] = c;
        ss[m + 1] = sl;
        kk[m + 1] = w;
    }
    for (int j = 0; j < len; ++j) {
        A.time[fo + cc[j]] = (u32)(i + j);
        A.comp[fo + i + j] = cc[j];
    }
}

__global__ void __launch_bounds__(256)
k_time_repair_long(TimeRepairArgs A) {
    __shared__ u64 s_key[REPAIR_SMEM];
    __shared__ u32 s_cmp[REPAIR_SMEM];
    __shared__ u32 s_slot[REPAIR_SMEM];
    __shared__ int s_len;
    const int n_list = min(*A.long_count, A.list_cap);
    for (int item = blockIdx.x; item < n_list; item += gridDim.x) {
        const uint2 it = A.long_list[item];
        const size_t fo = (size_t)it.x * A.N;
        const int i0 = (int)it.y;
        const u32* key = A.key + fo;
        const float2* f = A.flow + fo;
        const u32 k = key[i0];
        // length of the run
        if (threadIdx.x == 0) s_len = A.N - i0;
        __syncthreads();
        for (int base = 0; base < A.N - i0; base += 256) {
            const int j = base + threadIdx.x;
            if (j < A.N - i0 && key[i0 + j] != k) atomicMin(&s_len, j);
            __syncthreads();
            const int seen = s_len;
            __syncthreads();
            if (seen <= base + 256) break;
        }
        const int len = s_len;
        __syncthreads();
        if (len > REPAIR_SMEM) {
            if (threadIdx.x == 0) atomicExch(A.need_full, 1);
            continue;  // block-uniform
        }
        // odd-even transposition sort in shared memory by (weight, slot): len rounds of disjoint compare-exchanges
        for (int j = threadIdx.x; j < len; j += 256) {
            const u32 c = A.comp[fo + i0 + j];
            s_cmp[j] = c;
            s_slot[j] = A.slot[fo + c];
            s_key[j] = slot_weight(f, s_slot[j], A.W);
        }
        __syncthreads();
        for (int round = 0; round < len; ++round) {
            for (int j = 2 * threadIdx.x + (round & 1); j + 1 < len; j += 512) {
                if (run_less(s_key[j + 1], s_slot[j + 1], s_key[j], s_slot[j])) {
                    const u64 tk = s_key[j];
                    s_key[j] = s_key[j + 1];
                    s_key[j + 1] = tk;
                    const u32 ts = s_cmp[j];
                    s_cmp[j] = s_cmp[j + 1];
                    s_cmp[j + 1] = ts;
                    const u32 tl = s_slot[j];
                    s_slot[j] = s_slot[j + 1];
                    s_slot[j + 1] = tl;
                }
            }
            __syncthreads();
        }
        for (int j = threadIdx.x; j < len; j += 256) {
            A.time[fo + s_cmp[j]] = (u32)(i0 + j);
            A.comp[fo + i0 + j] = s_cmp[j];
        }
        __syncthreads();
    }
}

// fallback (enabled on the device by *enable): LSD order (weight, slot) = stable sort by slot, then stable sort by weight
__global__ void __launch_bounds__(SEG_THREADS)
k_time_fallback_slots(const u32* __restrict__ comp, const u32* __restrict__ slot, u64* __restrict__ skey, int N,
                      const int* __restrict__ enable) {
    if (*enable == 0) return;
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(i, N) {
        const u32 sl = slot[fo + comp[fo + i]];
        skey[fo + i] = sl == DOFS_INF32 ? 0xFFFFFFFFFFFFFFFFull : (u64)sl;
    }
}

__global__ void __launch_bounds__(SEG_THREADS)
k_time_fallback_weights(const u32* __restrict__ comp, const u32* __restrict__ slot, const float2* __restrict__ flow,
                        u64* __restrict__ wkey, int N, int W, const int* __restrict__ enable) {
    if (*enable == 0) return;
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(i, N) {
        const u32 sl = slot[fo + comp[fo + i]];
        wkey[fo + i] = sl == DOFS_INF32 ? 0xFFFFFFFFFFFFFFFFull : slot_weight(flow + fo, sl, W);
    }
}

__global__ void __launch_bounds__(SEG_THREADS)
k_time_fallback_rank(const u64* __restrict__ wkey, const u32* comp, u32* __restrict__ time, u32* order_out, int N,
                     const int* __restrict__ enable) {
    if (*enable == 0) return;
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(i, N) {
        const u32 c = comp[fo + i];
        time[fo + c] = wkey[fo + i] == 0xFFFFFFFFFFFFFFFFull ? DOFS_INF32 : (u32)i;
        order_out[fo + i] = c;  // (may be `comp` itself) the list the event sort starts from
    }
}

// ---------------------------------------------------------------------------------------------
// K9b  winner of every merge event (one event per losing root) and the event sort.  Events are wanted in
//      (wave of the winner, winner, time) order.  The losing roots are already listed in time order (K8), so a
//      STABLE sort of that list on the 32-bit key
//          key = lvl[winner] << wb | winner        (wb = bits of a pixel id; 1080p: 5 + 21 bits -> 4 radix passes)
//      is enough: the time never enters the key (it is read back from time[loser] by the few merges that pass the
//      gates).  k_event_keys writes the key by root id (coalesced, next to the walk that finds the winner),
//      k_event_gather lines the keys up with the time-ordered list.
// ---------------------------------------------------------------------------------------------
#define EV_KEY_NONE 0xFFFFFFFFu  // a root that never loses; real keys are below 2^(wb + 5) <= 2^31

struct EvBits {
    int wb;
};
DOFS_D u32 ev_chain(u32 key, EvBits) { return key; }  // wave | winner
DOFS_D u32 ev_winner(u32 key, EvBits b) { return key & ((1u << b.wb) - 1u); }
DOFS_D int ev_wave(u32 key, EvBits b) { return key == EV_KEY_NONE ? EV_MAX_WAVES - 1 : (int)(key >> b.wb); }

#ifndef EV_KEYS_ILP
#define EV_KEYS_ILP 4
#endif
__global__ void __launch_bounds__(SEG_THREADS)
k_event_keys(BorState S, u32* __restrict__ win, u32* __restrict__ key_by_root, int N, EvBits eb) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
#if EV_KEYS_ILP > 1
    // EV_KEYS_ILP roots per thread climb in lockstep: the climb is a chain of dependent gathers (time of the candidate,
    // then its `up` link), and one climb per thread leaves the memory system idle between them
    const u32* tm = S.loss_time + fo;
    const u32* up = S.up + fo;
    const int stride = gridDim.x * blockDim.x;
    for (int c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < N; c0 += EV_KEYS_ILP * stride) {
        u32 t[EV_KEYS_ILP], cur[EV_KEYS_ILP];
        u32 busy = 0;
#pragma unroll
        for (int k = 0; k < EV_KEYS_ILP; ++k) {
            const int c = c0 + k * stride;
            t[k] = DOFS_INF32;
            cur[k] = 0;
            if (c < N) {
                t[k] = tm[c];
                cur[k] = up[c];
                if (t[k] != DOFS_INF32) busy |= 1u << k;
            }
        }
        while (busy) {
            u32 tc[EV_KEYS_ILP], nx[EV_KEYS_ILP];
#pragma unroll
            for (int k = 0; k < EV_KEYS_ILP; ++k) {
                tc[k] = ((busy >> k) & 1u) ? tm[cur[k]] : DOFS_INF32;
                nx[k] = ((busy >> k) & 1u) ? up[cur[k]] : 0u;  // (speculative: used only if the climb goes on)
            }
#pragma unroll
            for (int k = 0; k < EV_KEYS_ILP; ++k) {
                if (!((busy >> k) & 1u)) continue;
                if (tc[k] < t[k]) cur[k] = nx[k];
                else busy &= ~(1u << k);
            }
        }
#pragma unroll
        for (int k = 0; k < EV_KEYS_ILP; ++k) {
            const int c = c0 + k * stride;
            if (c >= N) continue;
            if (t[k] == DOFS_INF32) {
                win[fo + c] = (u32)c;
                key_by_root[fo + c] = EV_KEY_NONE;
            } else {
                win[fo + c] = cur[k];
                key_by_root[fo + c] = ((u32)S.lvl[fo + cur[k]] << eb.wb) | cur[k];
            }
        }
    }
#else
    GRID_STRIDE(c, N) {
        const u32 t = S.loss_time[fo + c];
        if (t == DOFS_INF32) {
            win[fo + c] = (u32)c;
            key_by_root[fo + c] = EV_KEY_NONE;
            continue;
        }
        u32 cur = S.up[fo + c];
        while (S.loss_time[fo + cur] < t) cur = S.up[fo + cur];
        win[fo + c] = cur;
        key_by_root[fo + c] = ((u32)S.lvl[fo + cur] << eb.wb) | cur;
    }
#endif
}

// position i of the time-ordered list of losing roots -> the key of that root (roots that never lose are its tail)
__global__ void __launch_bounds__(SEG_THREADS)
k_event_gather(const u32* __restrict__ order, const u32* __restrict__ key_by_root, u32* __restrict__ ev_key, int N) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(i, N) ev_key[fo + i] = key_by_root[fo + order[fo + i]];
}

// first event index of every wave (events are sorted by key)
__global__ void __launch_bounds__(SEG_THREADS)
k_wave_starts(const u32* __restrict__ ev_key, int* __restrict__ wave_start /* [F][EV_MAX_WAVES+1] */, int N,
              EvBits eb) {
    const int frame = blockIdx.y;
    const u32* k = ev_key + (size_t)frame * N;
    int* ws = wave_start + frame * (EV_MAX_WAVES + 1);
    GRID_STRIDE(i, N) {
        const int wi = ev_wave(k[i], eb);
        const int wp = i == 0 ? -1 : ev_wave(k[i - 1], eb);
        // waves without events keep the start of the next non-empty wave
        for (int w = wp + 1; w <= wi; ++w) ws[w] = i;
        if (i == N - 1)
            for (int w = wi + 1; w <= EV_MAX_WAVES; ++w) ws[w] = N;
    }
}

// ---------------------------------------------------------------------------------------------
// K9c + K10  chain replay of one wave.  Forest::merge (graph.cpp:170-218) state update per absorbed
// root — size, bounding box, and the ORDER-DEPENDENT float mean flow — then the size / row / move
// gates of Forest::new_merge (graph.cpp:280-300); survivors are queued for lifting.
//
// A chain = the events won by one root, in time order (contiguous in the sorted event array).  Chains
// of one wave are independent (every absorbed root has a lower final rank, so its state is final).
//   k_replay_short     every chain of at most REPLAY_SHORT events (almost all of them) is replayed whole by one thread:
//                      state update and gates in a single pass over its events.  Longer chains are flagged and listed.
// For the long chains only the mean flow is a true recurrence; sizes and boxes are prefix sums / prefix min-max along
// the chain.  So they take five steps, each of which skips the events of short chains (and whole tiles without long ones):
//   k_replay_scan      per event: gather the absorbed root's state; segmented scan (segments = chains)
//                      of (size, box) inside tiles of 256 events; one aggregate per tile
//   k_replay_carry     per frame: running (size, box) across the tiles of the wave
//   k_replay_operands  per event: size before / after, box after, the operands of the recurrence
//                      (loser_mean*loser_size, float(size_before), 1.0/size_after)
//   k_replay_serial_long  the recurrence itself, nothing else: a warp per chain (operands stream in coalesced,
//                      prefetched, broadcast through shared memory)
//   k_replay_gates     per event: gates, candidate queue; the last event of a chain stores the root's state
// The only serial work left is ~10 instructions per event of the longest chain.
// ---------------------------------------------------------------------------------------------
#define REPLAY_SHORT 32
#define REPLAY_TILE 256

struct ScanTuple {  // absorbed size and bounding box since the start of the chain (or of the tile)
    int s;
    int x0, y0, x1, y1;
};
DOFS_D ScanTuple scan_combine(const ScanTuple& a, const ScanTuple& b) {  // a then b
    ScanTuple r;
    r.s = a.s + b.s;
    r.x0 = min(a.x0, b.x0);
    r.y0 = min(a.y0, b.y0);
    r.x1 = max(a.x1, b.x1);
    r.y1 = max(a.y1, b.y1);
    return r;
}
DOFS_D ScanTuple scan_identity() {
    ScanTuple r;
    r.s = 0;
    r.x0 = 65535;
    r.y0 = 65535;
    r.x1 = 0;
    r.y1 = 0;
    return r;
}
DOFS_D ScanTuple scan_shfl_up(const ScanTuple& t, int o) {
    ScanTuple r;
    r.s = __shfl_up_sync(0xffffffffu, t.s, o);
    r.x0 = __shfl_up_sync(0xffffffffu, t.x0, o);
    r.y0 = __shfl_up_sync(0xffffffffu, t.y0, o);
    r.x1 = __shfl_up_sync(0xffffffffu, t.x1, o);
    r.y1 = __shfl_up_sync(0xffffffffu, t.y1, o);
    return r;
}

struct TileAgg {  // what a tile passes on: the scan value of its last event, and whether a chain started inside it
    int s;
    ushort4 bb;
    int has_head;
};

struct ReplayArgs {
    const u32* ev_key;     // [F][N] sorted: wave | winner
    const u32* ev_loser;   // [F][N] sorted payload, in time order inside a chain
    const u32* time;       // [F][N] per root id: position of its loss in the merge sequence
    const int* wave_start; // [F][EV_MAX_WAVES+1]
    RootState* rstate;     // [F][N] size, mean flow and box of every root that has won a merge (written when its chain ends)
    const float2* flow;    // [F][N] blurred flow: the state of a root that never won is its pixel (Forest::Forest, graph.cpp:129-148)
    const u8* lvl;         // [F][N] final rank of every root: 0 = never won a merge
    // per event (index = position in the sorted event array)
    float4* ev_op;         // [F][N] (loser mean * loser size).xy, float(size before), float(size after)
    double* ev_inv;        // [F][N] 1.0 / size after
    int* ev_size;          // [F][N] k_replay_scan: tile-local scan | flag bit 31; k_replay_operands: size after
    ushort4* ev_bbox;      // [F][N] tile-local scan, then box after
    float2* ev_prod;       // [F][N] loser mean * loser size
    int* ev_sa;            // [F][N] loser size
    float2* ev_flow;       // [F][N] mean flow after the event
    TileAgg* tile_agg;     // [F][tiles]
    TileAgg* tile_carry;   // [F][tiles] state carried into the tile
    int tiles_cap;
    Candidate* cand;       // [F][cand_cap]
    int* n_cand;           // [F]
    int* longest_chain;    // [F]
    u8* long_flag;         // [F][N] per root: its chain has more than REPLAY_SHORT events (set by k_replay_short; zeroed per call)
    uint2* long_list;      // [list_cap] (frame, index of the chain's first event)
    int* long_count;       // [EV_MAX_WAVES + 1]
    int list_cap;
    int cand_cap;
    int W, H, N;
    u32 wm;                // fastdiv_magic(W)
    int min_size;
    EvBits eb;
    const double* rcp;     // [RCP_TABLE] 1.0 / n, correctly rounded (filled on the host: IEEE division), for the small set
                           // sizes of the early waves
};
#define RCP_TABLE 4096
// 1.0 / n in double, as Vec2f / int computes it (graph.cpp:188): the table entry for small n, the division otherwise
DOFS_D double size_reciprocal(const ReplayArgs& A, int n) {
    return n < RCP_TABLE ? __ldg(A.rcp + n) : xddiv(1.0, (double)n);
}

// Forest::merge's size-weighted mean (graph.cpp:184-190) with OpenCV's Vec2f rounding:
// Vec2f * int -> float products, float sum, Vec2f / int -> multiply by the double reciprocal.
DOFS_D float merge_mean(float fa_times_sa, float f, float sb, double inv) {
    return (float)xdmul((double)xfadd(fa_times_sa, xfmul(f, sb)), inv);
}

DOFS_D void push_candidate(const ReplayArgs& A, int frame, u32 root, u32 time, int size, float2 f, ushort4 bb) {
    const int slot = atomicAdd(&A.n_cand[frame], 1);
    if (slot >= A.cand_cap) return;
    Candidate c;
    c.root = root;
    c.time = time;
    c.size = size;
    c.fx = f.x;
    c.fy = f.y;
    c.bbox[0] = bb.x;
    c.bbox[1] = bb.y;
    c.bbox[2] = bb.z;
    c.bbox[3] = bb.w;
    c.pad = 0;
    A.cand[(size_t)frame * A.cand_cap + slot] = c;
}

// State of root c before its own chain: the singleton set of pixel c.  A root wins merges only in the wave of its
// final rank — all of them in one chain — so this is also its state at the head of that chain.
DOFS_D RootState root_initial(const ReplayArgs& A, size_t fo, u32 c) {
    RootState r;
    const float2 f = A.flow[fo + c];
    int x;
    const int y = fastdiv((int)c, A.W, A.wm, &x);
    r.size = 1;
    r.fx = f.x;
    r.fy = f.y;
    r.pad0 = r.pad1 = r.pad2 = 0;
    r.bbox = make_ushort4((u16)x, (u16)y, (u16)x, (u16)y);
    return r;
}
// State of an absorbed root: final (stored when its chain ended, in an earlier wave) if it ever won, else its pixel.
// In wave 1 every absorbed root has rank 0.
DOFS_D RootState root_absorbed(const ReplayArgs& A, size_t fo, u32 a, int wave) {
    if (wave == 1 || A.lvl[fo + a] == 0) return root_initial(A, fo, a);
    return root_load(&A.rstate[fo + a]);
}

#define EV_FLAG_STARTED 0x80000000u  // in ev_size after k_replay_scan: the event's chain started inside its tile

DOFS_D void replay_gate(const ReplayArgs& A, int frame, u32 r, u32 loser, int s, float2 f, ushort4 bb) {
    int x_;
    const int y = fastdiv((int)r, A.W, A.wm, &x_);
    if (s >= A.min_size && !(y < A.H / 10)) {                                       // graph.cpp:280, 288
        const double move = norm2d(f.x, f.y);
        if (!(move < xddiv((double)(3 * (y + 1)), (double)A.H)))                     // graph.cpp:296
            push_candidate(A, frame, r, A.time[(size_t)frame * A.N + loser], s, f, bb);
    }
}

// Chains of at most REPLAY_SHORT events — all but a few thousand of the ~2 M of a 1080p frame — are replayed whole by
// the thread of their first event: Forest::merge's state update (graph.cpp:184-208) and the gates of every merge, one
// pass over the events instead of five.  The head of a longer chain flags its root and queues the chain for the
// scan / operands / serial / gates kernels below, which skip everything else.
#ifndef REPLAY_PIPELINED
#define REPLAY_PIPELINED 0
#endif
#ifndef REPLAY_SHORT_BLOCKS
#define REPLAY_SHORT_BLOCKS 8  // bound by the latency of the random state gathers: full occupancy (32 registers)
#endif
__global__ void __launch_bounds__(SEG_THREADS, REPLAY_SHORT_BLOCKS)
k_replay_short(ReplayArgs A, int wave) {
    const int frame = blockIdx.y;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const size_t fo = (size_t)frame * A.N;
    const u32* key = A.ev_key + fo;
    for (int i = w0 + blockIdx.x * blockDim.x + threadIdx.x; i < w1; i += gridDim.x * blockDim.x) {
        u32 k = key[i];
        const u32 chain = ev_chain(k, A.eb);
        if (i > w0 && ev_chain(key[i - 1], A.eb) == chain) continue;  // not a head
        const u32 r = ev_winner(k, A.eb);
        if (i + REPLAY_SHORT < w1 && ev_chain(key[i + REPLAY_SHORT], A.eb) == chain) {
            A.long_flag[fo + r] = 1;
            const int slot = atomicAdd(&A.long_count[wave], 1);
            if (slot < A.list_cap) A.long_list[slot] = make_uint2((u32)frame, (u32)i);
            continue;
        }
        const RootState r0 = root_initial(A, fo, r);
        int s = r0.size;
        float2 f = make_float2(r0.fx, r0.fy);
        ushort4 bb = r0.bbox;
        int j = i;
#if REPLAY_PIPELINED
        // the state of the NEXT absorbed root is requested before the arithmetic of this event: the chain of dependent
        // loads (key -> loser id -> its state, a random gather) then runs under the division-free mean update and the gates
        u32 la = A.ev_loser[fo + j];
        RootState ra = root_absorbed(A, fo, la, wave);
        for (;;) {
            const int jn = j + 1;
            u32 kn = 0, ln = 0;
            bool same = false;
            RootState rn = ra;
            if (jn < w1) {
                kn = key[jn];
                same = ev_chain(kn, A.eb) == chain;
                if (same) {
                    ln = A.ev_loser[fo + jn];
                    rn = root_absorbed(A, fo, ln, wave);
                }
            }
            const int sa = ra.size;
            const float fsa = (float)sa;
            const int s_after = s + sa;
            const double inv = size_reciprocal(A, s_after);
            f.x = merge_mean(xfmul(ra.fx, fsa), f.x, (float)s, inv);
            f.y = merge_mean(xfmul(ra.fy, fsa), f.y, (float)s, inv);
            s = s_after;
            bb.x = min(bb.x, ra.bbox.x);
            bb.y = min(bb.y, ra.bbox.y);
            bb.z = max(bb.z, ra.bbox.z);
            bb.w = max(bb.w, ra.bbox.w);
            replay_gate(A, frame, r, la, s, f, bb);
            j = jn;
            if (!same) break;
            ra = rn;
            la = ln;
            k = kn;
        }
#else
        for (;;) {
            const u32 la = A.ev_loser[fo + j];
            const RootState ra = root_absorbed(A, fo, la, wave);
            const int sa = ra.size;
            const float2 fa = make_float2(ra.fx, ra.fy);
            const ushort4 ba = ra.bbox;
            const float fsa = (float)sa;
            const int s_after = s + sa;
            const double inv = size_reciprocal(A, s_after);
            f.x = merge_mean(xfmul(fa.x, fsa), f.x, (float)s, inv);
            f.y = merge_mean(xfmul(fa.y, fsa), f.y, (float)s, inv);
            s = s_after;
            bb.x = min(bb.x, ba.x);
            bb.y = min(bb.y, ba.y);
            bb.z = max(bb.z, ba.z);
            bb.w = max(bb.w, ba.w);
            replay_gate(A, frame, r, la, s, f, bb);
            ++j;
            if (j >= w1) break;
            k = key[j];
            if (ev_chain(k, A.eb) != chain) break;
        }
#endif
        root_store(&A.rstate[fo + r], s, f, bb);
        if (j - i > A.longest_chain[frame]) atomicMax(&A.longest_chain[frame], j - i);
    }
}

__global__ void __launch_bounds__(REPLAY_TILE)
k_replay_scan(ReplayArgs A, int wave) {
    __shared__ ScanTuple s_tot[REPLAY_TILE / 32];
    __shared__ int s_head[REPLAY_TILE / 32];
    const int frame = blockIdx.y;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const size_t fo = (size_t)frame * A.N;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n_tiles = (w1 - w0 + REPLAY_TILE - 1) / REPLAY_TILE;
    if (A.long_count[wave] == 0) return;  // nothing but short chains in this wave (any frame)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int j = w0 + tile * REPLAY_TILE + threadIdx.x;
        const bool valid = j < w1;
        ScanTuple t = scan_identity();
        bool head = false;
        u32 key = 0;
        bool is_long = false;
        if (valid) {
            key = A.ev_key[fo + j];
            is_long = A.long_flag[fo + ev_winner(key, A.eb)] != 0;
        }
        if (!__syncthreads_or(is_long)) {  // a tile without events of long chains: nothing crosses it
            if (threadIdx.x == 0) {
                TileAgg g;
                g.s = 0;
                g.bb = make_ushort4(65535, 65535, 0, 0);
                g.has_head = 1;
                A.tile_agg[(size_t)frame * A.tiles_cap + tile] = g;
            }
            continue;
        }
        if (valid) head = j == w0 || ev_chain(A.ev_key[fo + j - 1], A.eb) != ev_chain(key, A.eb);
        if (is_long) {
            const RootState ra = root_absorbed(A, fo, A.ev_loser[fo + j], wave);
            const int sa = ra.size;
            const float2 fa = make_float2(ra.fx, ra.fy);
            const ushort4 ba = ra.bbox;
            const float fsa = (float)sa;
            A.ev_prod[fo + j] = make_float2(xfmul(fa.x, fsa), xfmul(fa.y, fsa));
            A.ev_sa[fo + j] = sa;
            t.s = sa;
            t.x0 = ba.x;
            t.y0 = ba.y;
            t.x1 = ba.z;
            t.y1 = ba.w;
        }
        // segmented inclusive scan in the warp: `started` = a head at or before this lane (within the warp)
        int started = head ? 1 : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const ScanTuple u = scan_shfl_up(t, o);
            const int us = __shfl_up_sync(0xffffffffu, started, o);
            if (lane >= o) {
                if (!started) t = scan_combine(u, t);
                started |= us;
            }
        }
        if (lane == 31) {
            s_tot[wrp] = t;
            s_head[wrp] = started;
        }
        __syncthreads();
        // carry from the previous warps of the tile
        ScanTuple c = scan_identity();
        int c_started = 0;
        for (int w = 0; w < wrp; ++w) {
            if (s_head[w]) {
                c = s_tot[w];
                c_started = 1;
            } else {
                c = scan_combine(c, s_tot[w]);
            }
        }
        if (!started) t = scan_combine(c, t);
        started |= c_started;
        if (is_long) {
            A.ev_size[fo + j] = (int)((u32)t.s | (started ? EV_FLAG_STARTED : 0u));
            A.ev_bbox[fo + j] = make_ushort4((u16)t.x0, (u16)t.y0, (u16)t.x1, (u16)t.y1);
        }
        if (threadIdx.x == REPLAY_TILE - 1) {  // padding lanes carry the identity, so this is the last valid event's value
            TileAgg g;
            g.s = t.s;
            g.bb = make_ushort4((u16)t.x0, (u16)t.y0, (u16)t.x1, (u16)t.y1);
            g.has_head = started;
            A.tile_agg[(size_t)frame * A.tiles_cap + tile] = g;
        }
        __syncthreads();
    }
}

// one warp per frame: the state carried into every tile of the wave (segmented exclusive scan of the tile aggregates)
__global__ void __launch_bounds__(32)
k_replay_carry(ReplayArgs A, int wave) {
    const int frame = blockIdx.x, lane = threadIdx.x;
    if (A.long_count[wave] == 0) return;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const int n_tiles = (w1 - w0 + REPLAY_TILE - 1) / REPLAY_TILE;
    ScanTuple carry = scan_identity();  // state after all tiles before this round
    for (int base = 0; base < n_tiles; base += 32) {
        const int t = base + lane;
        ScanTuple v = scan_identity();
        int started = 0;
        if (t < n_tiles) {
            const TileAgg g = A.tile_agg[(size_t)frame * A.tiles_cap + t];
            v.s = g.s;
            v.x0 = g.bb.x;
            v.y0 = g.bb.y;
            v.x1 = g.bb.z;
            v.y1 = g.bb.w;
            started = g.has_head;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const ScanTuple u = scan_shfl_up(v, o);
            const int us = __shfl_up_sync(0xffffffffu, started, o);
            if (lane >= o) {
                if (!started) v = scan_combine(u, v);
                started |= us;
            }
        }
        if (!started) v = scan_combine(carry, v);  // inclusive value after tile t
        // exclusive: what enters tile t is the inclusive value after tile t-1
        ScanTuple in = scan_shfl_up(v, 1);
        if (lane == 0) in = carry;
        if (t < n_tiles) {
            TileAgg c;
            c.s = in.s;
            c.bb = make_ushort4((u16)in.x0, (u16)in.y0, (u16)in.x1, (u16)in.y1);
            c.has_head = 0;
            A.tile_carry[(size_t)frame * A.tiles_cap + t] = c;
        }
        carry.s = __shfl_sync(0xffffffffu, v.s, 31);
        carry.x0 = __shfl_sync(0xffffffffu, v.x0, 31);
        carry.y0 = __shfl_sync(0xffffffffu, v.y0, 31);
        carry.x1 = __shfl_sync(0xffffffffu, v.x1, 31);
        carry.y1 = __shfl_sync(0xffffffffu, v.y1, 31);
    }
}

__global__ void __launch_bounds__(SEG_THREADS)
k_replay_operands(ReplayArgs A, int wave) {
    const int frame = blockIdx.y;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const size_t fo = (size_t)frame * A.N;
    if (A.long_count[wave] == 0) return;
    for (int j = w0 + blockIdx.x * blockDim.x + threadIdx.x; j < w1; j += gridDim.x * blockDim.x) {
        const u32 key = A.ev_key[fo + j];
        const u32 r = ev_winner(key, A.eb);
        if (A.long_flag[fo + r] == 0) continue;  // replayed by k_replay_short
        const u32 raw = (u32)A.ev_size[fo + j];
        int s = (int)(raw & ~EV_FLAG_STARTED);
        ushort4 bb = A.ev_bbox[fo + j];
        if (!(raw & EV_FLAG_STARTED)) {  // the chain started in an earlier tile
            const TileAgg c = A.tile_carry[(size_t)frame * A.tiles_cap + (j - w0) / REPLAY_TILE];
            s += c.s;
            bb.x = min(bb.x, c.bb.x);
            bb.y = min(bb.y, c.bb.y);
            bb.z = max(bb.z, c.bb.z);
            bb.w = max(bb.w, c.bb.w);
        }
        // the root's own state before this wave
        const RootState rr = root_initial(A, fo, r);
        const int s_after = s + rr.size;
        const ushort4 rb = rr.bbox;
        bb.x = min(bb.x, rb.x);
        bb.y = min(bb.y, rb.y);
        bb.z = max(bb.z, rb.z);
        bb.w = max(bb.w, rb.w);
        const int sa = A.ev_sa[fo + j];
        const float2 pr = A.ev_prod[fo + j];
        A.ev_op[fo + j] = make_float4(pr.x, pr.y, (float)(s_after - sa), (float)s_after);
        A.ev_inv[fo + j] = size_reciprocal(A, s_after);
        A.ev_size[fo + j] = s_after;
        A.ev_bbox[fo + j] = bb;
    }
}

// The update without double arithmetic on the critical path.  With n = size after the merge (an integer
// <= 2^24, exact in float) the reference's (float)((double)u * (1.0 / n)) equals the correctly rounded float
// quotient u / n: u / n is never closer than 2^-49 (relative) to a midpoint between two floats when u has 24 and n
// at most 25 significant bits, while u * RN(1/n) rounded to double is within 2^-52 of it, so both round to the same
// float.  q0 = u*r with r = RN(1/n), then one correction by the residual (two FMAs) lands on that float except
// when u / n is within ~2^-46 of a midpoint.  Nothing here is trusted: every result is re-derived with merge_mean
// by the lane that owns the event (in parallel, off the serial path) and a chunk with any mismatch is replayed exactly.
DOFS_D float merge_mean_fast(float fa_times_sa, float f, float sb, float nf, float r) {
    const float u = xfadd(fa_times_sa, xfmul(f, sb));
    const float q0 = xfmul(u, r);
    return __fmaf_rn(__fmaf_rn(-nf, q0, u), r, q0);
}

// the recurrence of a long chain: one warp per chain, 32 events per round.  Operands arrive coalesced, one round
// ahead, and are broadcast through shared memory in groups of REPLAY_G so that no step waits for them.
#define REPLAY_WARPS 4
#define REPLAY_G 8

__global__ void __launch_bounds__(32 * REPLAY_WARPS)
k_replay_serial_long(ReplayArgs A, int wave) {
    __shared__ float4 s_op[REPLAY_WARPS][32];
    __shared__ float s_rcp[REPLAY_WARPS][32];
    __shared__ double s_inv[REPLAY_WARPS][32];
    __shared__ float2 s_f[REPLAY_WARPS][33];  // [0] = mean before the round, [k+1] = after event k
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_list = min(A.long_count[wave], A.list_cap);
    const unsigned FULL = 0xffffffffu;
    int redone = 0;
    for (int item = warp; item < n_list; item += n_warps) {
        const uint2 it = A.long_list[item];
        const int frame = (int)it.x, i0 = (int)it.y;
        const size_t fo = (size_t)frame * A.N;
        const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
        const u32 k0 = A.ev_key[fo + i0];
        const u32 chain = ev_chain(k0, A.eb);
        float2 f;
        {
            f = A.flow[fo + ev_winner(k0, A.eb)];  // the root's own pixel (root_initial)
        }
        int j0 = i0;
        // rounds t+1 and t+2 are in flight while round t is replayed; a load never waits for the chain test of its
        // event (operands of any event position are readable), the test happens when the round is consumed
        float4 n_op[2];
        double n_inv[2];
        u32 n_key[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = j0 + 32 * q + lane;
            n_key[q] = EV_KEY_NONE;
            n_op[q] = make_float4(0.f, 0.f, 0.f, 1.f);
            n_inv[q] = 1.0;
            if (j < w1) {
                n_key[q] = A.ev_key[fo + j];
                n_op[q] = A.ev_op[fo + j];
                n_inv[q] = A.ev_inv[fo + j];
            }
        }
        for (;;) {
            const float4 c_op = n_op[0];
            const double c_inv = n_inv[0];
            const bool c_valid = n_key[0] != EV_KEY_NONE && ev_chain(n_key[0], A.eb) == chain;
            n_op[0] = n_op[1];
            n_inv[0] = n_inv[1];
            n_key[0] = n_key[1];
            {
                const int j = j0 + 64 + lane;
                n_key[1] = EV_KEY_NONE;
                n_op[1] = make_float4(0.f, 0.f, 0.f, 1.f);
                n_inv[1] = 1.0;
                if (j < w1) {
                    n_key[1] = A.ev_key[fo + j];
                    n_op[1] = A.ev_op[fo + j];
                    n_inv[1] = A.ev_inv[fo + j];
                }
            }
            const int cnt = __popc(__ballot_sync(FULL, c_valid));  // valid lanes are a prefix
            if (cnt == 0) break;
            s_op[wib][lane] = c_op;
            s_inv[wib][lane] = c_inv;
            s_rcp[wib][lane] = __frcp_rn(c_op.w);
            if (lane == 0) s_f[wib][0] = f;
            __syncwarp();
            // sizes beyond 2^24 are not exact in float: such chunks take the double path directly
            bool exact_needed = cnt < 32 || __shfl_sync(FULL, c_op.w, 31) > 16777216.f;
            const float2 f_start = f;
            if (!exact_needed) {
                float4 o4[REPLAY_G];
                float rc[REPLAY_G];
#pragma unroll
                for (int k = 0; k < REPLAY_G; ++k) {
                    o4[k] = s_op[wib][k];
                    rc[k] = s_rcp[wib][k];
                }
#pragma unroll
                for (int g = 0; g < 32 / REPLAY_G; ++g) {
                    float4 n4[REPLAY_G];
                    float nr[REPLAY_G];
                    if (g + 1 < 32 / REPLAY_G) {
#pragma unroll
                        for (int k = 0; k < REPLAY_G; ++k) {
                            n4[k] = s_op[wib][(g + 1) * REPLAY_G + k];
                            nr[k] = s_rcp[wib][(g + 1) * REPLAY_G + k];
                        }
                    }
                    float2 fh[REPLAY_G];
#pragma unroll
                    for (int k = 0; k < REPLAY_G; ++k) {
                        f.x = merge_mean_fast(o4[k].x, f.x, o4[k].z, o4[k].w, rc[k]);
                        f.y = merge_mean_fast(o4[k].y, f.y, o4[k].z, o4[k].w, rc[k]);
                        fh[k] = f;
                    }
#pragma unroll
                    for (int k = 0; k < REPLAY_G; ++k) s_f[wib][g * REPLAY_G + k + 1] = fh[k];
                    if (g + 1 < 32 / REPLAY_G) {
#pragma unroll
                        for (int k = 0; k < REPLAY_G; ++k) {
                            o4[k] = n4[k];
                            rc[k] = nr[k];
                        }
                    }
                }
                __syncwarp();
                // every lane re-derives "its" step exactly from the mean before it
                const float2 before = s_f[wib][lane], after = s_f[wib][lane + 1];
                const float ex = merge_mean(c_op.x, before.x, c_op.z, c_inv), ey = merge_mean(c_op.y, before.y, c_op.z, c_inv);
                const bool same = __float_as_uint(ex) == __float_as_uint(after.x) && __float_as_uint(ey) == __float_as_uint(after.y);
                if (__ballot_sync(FULL, same) != FULL) {
                    exact_needed = true;
                    f = f_start;
                    ++redone;
                }
                __syncwarp();
            }
            if (exact_needed) {
                for (int k = 0; k < cnt; ++k) {
                    const float4 o = s_op[wib][k];
                    const double inv = s_inv[wib][k];
                    f.x = merge_mean(o.x, f.x, o.z, inv);
                    f.y = merge_mean(o.y, f.y, o.z, inv);
                    s_f[wib][k + 1] = f;
                }
                __syncwarp();
            }
            if (c_valid) A.ev_flow[fo + j0 + lane] = s_f[wib][lane + 1];
            __syncwarp();
            j0 += cnt;
            if (cnt < 32) break;
        }
        if (lane == 0) atomicMax(&A.longest_chain[frame], j0 - i0);
    }
    if (lane == 0 && redone) atomicAdd(&A.long_count[0], redone);  // slot 0 is no wave's counter: chunks replayed exactly
}

// gates of every event (graph.cpp:280-300); the last event of a chain leaves the root's state behind
__global__ void __launch_bounds__(SEG_THREADS)
k_replay_gates(ReplayArgs A, int wave) {
    const int frame = blockIdx.y;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const size_t fo = (size_t)frame * A.N;
    if (A.long_count[wave] == 0) return;
    for (int j = w0 + blockIdx.x * blockDim.x + threadIdx.x; j < w1; j += gridDim.x * blockDim.x) {
        const u32 key = A.ev_key[fo + j];
        const u32 r = ev_winner(key, A.eb);
        if (A.long_flag[fo + r] == 0) continue;  // replayed by k_replay_short
        const int s = A.ev_size[fo + j];
        const float2 f = A.ev_flow[fo + j];
        const ushort4 bb = A.ev_bbox[fo + j];
        replay_gate(A, frame, r, A.ev_loser[fo + j], s, f, bb);
        if (j + 1 >= w1 || ev_chain(A.ev_key[fo + j + 1], A.eb) != ev_chain(key, A.eb)) {
            root_store(&A.rstate[fo + r], s, f, bb);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K11 + K12  lifting of the queued merges, per-root first-maximum selection, box records, labels
// ---------------------------------------------------------------------------------------------
// Scores are compared through an order-preserving map of the double to u64 (any admissible score, negative ones included,
// maps above 0, which stands for "no score yet": segment_history starts at -1 and keeps any score above the caller's
// threshold, graph.cpp:348-352).
DOFS_D u64 score_key(double s) {
    const u64 b = (u64)__double_as_longlong(s);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
#define WIN_HAS_BOX 0x80000000u  // in win[root]: the root emitted a box (set by k_sort_boxes, read by first_box_above)
#define CAND_KEPT 1u  // Candidate::pad bit 0: passed the convexity and score gates (graph.cpp:341-346)

struct SelectArgs {
    Candidate* cand;        // [F][cand_cap]
    const int* n_cand;      // [F]
    double* cand_score;     // [F][cand_cap]  get_score of the merge (-1: no rectangle), whatever the later gates say
    u64* best_score;        // [F][N] score_key of the best kept score per root (0 = none)
    u32* sel_time;          // [F][N] time of the first merge reaching the best score
    int* sel_box;           // [F][N] index of the root's box in the sorted box list
    int* n_scored;          // [F]
    int* n_boxes;           // [F]
    int cand_cap;
    int N;
};

// The per-root selection state is dense ([F][N]) but only the roots of queued merges ever use it: they are reset here,
// nothing else is initialised (a root "has a box" through WIN_HAS_BOX in its `win` word, see first_box_above).
__global__ void __launch_bounds__(SEG_THREADS)
k_select_reset(SelectArgs A) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(A.n_cand[frame], A.cand_cap);
    if (i >= n) return;
    const size_t g = (size_t)frame * A.N + A.cand[(size_t)frame * A.cand_cap + i].root;
    A.best_score[g] = 0;  // "no score yet"
    A.sel_time[g] = DOFS_INF32;
}

// graph.cpp:302-348: convexity, get_score, class-dependent convexity gate, score threshold
__global__ void __launch_bounds__(128)
k_lift_score(SelectArgs A, SegParams P) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(A.n_cand[frame], A.cand_cap);
    if (i >= n) return;
    const Candidate c = A.cand[(size_t)frame * A.cand_cap + i];
    const int xmin = c.bbox[0], ymin = c.bbox[1], xmax = c.bbox[2], ymax = c.bbox[3];
    const double rect_area = (double)((xmax - xmin + 1) * (ymax - ymin + 1));
    const double convexity = xddiv((double)c.size, rect_area);
    LiftSolution sol;
    const double score = lift_get_score(c.fx, c.fy, xmin, ymin, xmax, ymax, P, &sol);
    bool kept = false;
    if (score != -1.0) {
        const double min_convexity = P.cls_min_convexity[sol.cls];
        kept = !(convexity < min_convexity) && score > P.score_threshold;
    }
    A.cand_score[(size_t)frame * A.cand_cap + i] = score;
    A.cand[(size_t)frame * A.cand_cap + i].pad = kept ? CAND_KEPT : 0u;
    if (kept) {
        atomicMax((unsigned long long*)&A.best_score[(size_t)frame * A.N + c.root], (unsigned long long)score_key(score));
        atomicAdd(&A.n_scored[frame], 1);
    }
}

// history keeps the FIRST merge (in time) that reaches the maximum score (strict '<', graph.cpp:352)
__global__ void __launch_bounds__(SEG_THREADS)
k_select_time(SelectArgs A) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(A.n_cand[frame], A.cand_cap);
    if (i >= n) return;
    const Candidate c = A.cand[(size_t)frame * A.cand_cap + i];
    if (!(c.pad & CAND_KEPT)) return;
    const double sc = A.cand_score[(size_t)frame * A.cand_cap + i];
    if (score_key(sc) == A.best_score[(size_t)frame * A.N + c.root])
        atomicMin(&A.sel_time[(size_t)frame * A.N + c.root], c.time);
}

struct BoxRec;  // == dofs3d_box (include/dofs3d.h); defined in dofs3d.cu

template <typename Box>
DOFS_D void fill_box(Box* b, const Candidate& c, double score, const LiftSolution& s) {
    b->root = (int)c.root;
    b->size = c.size;
    b->cls = s.cls;
    b->parent_box = -1;
    b->bbox[0] = c.bbox[0];
    b->bbox[1] = c.bbox[1];
    b->bbox[2] = c.bbox[2];
    b->bbox[3] = c.bbox[3];
    b->time = c.time;
    b->mean_flow[0] = c.fx;
    b->mean_flow[1] = c.fy;
    b->pad_ = 0.f;
    b->score = score;
    b->move = norm2d(c.fx, c.fy);
    b->orient = s.orient;
    b->w_error = s.w_error;
    b->h_error = s.h_error;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        b->ps_bev[2 * k] = s.ps_bev[k].x;
        b->ps_bev[2 * k + 1] = s.ps_bev[k].y;
        b->rectangle[2 * k] = s.rect[k].x;
        b->rectangle[2 * k + 1] = s.rect[k].y;
        b->lower_face[2 * k] = s.lower[k].x;
        b->lower_face[2 * k + 1] = s.lower[k].y;
        b->upper_face[2 * k] = s.upper[k].x;
        b->upper_face[2 * k + 1] = s.upper[k].y;
    }
}

// the selected merge of every root emits its box (unsorted) — lifting recomputed for the few winners
template <typename Box>
__global__ void __launch_bounds__(128)
k_emit_boxes(SelectArgs A, SegParams P, Box* __restrict__ tmp_boxes, int box_cap) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(A.n_cand[frame], A.cand_cap);
    if (i >= n) return;
    const Candidate c = A.cand[(size_t)frame * A.cand_cap + i];
    if (!(c.pad & CAND_KEPT)) return;
    const double sc = A.cand_score[(size_t)frame * A.cand_cap + i];
    const size_t g = (size_t)frame * A.N + c.root;
    if (score_key(sc) != A.best_score[g] || c.time != A.sel_time[g]) return;
    const int slot = atomicAdd(&A.n_boxes[frame], 1);
    if (slot >= box_cap) return;
    LiftSolution sol;
    const double score = lift_get_score(c.fx, c.fy, c.bbox[0], c.bbox[1], c.bbox[2], c.bbox[3], P, &sol);
    fill_box(&tmp_boxes[(size_t)frame * box_cap + slot], c, score, sol);
}

// order boxes by ascending root (the order of the reference's history vector) — counting rank, the
// list is a few hundred entries; then the nesting parent of every box
template <typename Box>
__global__ void __launch_bounds__(SEG_THREADS)
k_sort_boxes(const Box* __restrict__ tmp_boxes, Box* __restrict__ boxes, const int* __restrict__ n_boxes, int box_cap,
             int* __restrict__ sel_box, u32* __restrict__ win, int N) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(n_boxes[frame], box_cap);
    if (i >= n) return;
    const Box* src = tmp_boxes + (size_t)frame * box_cap;
    const int root = src[i].root;
    int pos = 0;
    for (int j = 0; j < n; ++j) pos += src[j].root < root ? 1 : 0;
    boxes[(size_t)frame * box_cap + pos] = src[i];
    sel_box[(size_t)frame * N + root] = pos;
    win[(size_t)frame * N + root] |= WIN_HAS_BOX;  // (one box per root: no other thread touches this word)
}

// Walk the union-find link forest (parent = winner at loss time, <= max-rank hops): a pixel / set that
// entered root a's set at time t_in belongs to a's snapshot taken at sel_time[a] iff t_in <= sel_time[a].
// `win` carries WIN_HAS_BOX for the roots that emitted a box: only their sel_time / sel_box entries are defined.
DOFS_D int first_box_above(const u32* loss_time, const u32* win, const u32* sel_time, const int* sel_box, u32 cur,
                           bool include_self) {
    u32 t_in = 0;
    bool first = include_self;
    if (!include_self) {
        t_in = loss_time[cur];
        if (t_in == DOFS_INF32) return -1;
        cur = win[cur] & ~WIN_HAS_BOX;
    }
    for (;;) {
        const u32 w = win[cur];
        if ((w & WIN_HAS_BOX) && (first || sel_time[cur] >= t_in)) return sel_box[cur];
        first = false;
        t_in = loss_time[cur];
        if (t_in == DOFS_INF32) return -1;
        cur = w & ~WIN_HAS_BOX;
    }
}

template <typename Box>
__global__ void __launch_bounds__(SEG_THREADS)
k_box_parents(Box* __restrict__ boxes, const int* __restrict__ n_boxes, int box_cap, const u32* __restrict__ loss_time,
              const u32* __restrict__ win, const u32* __restrict__ sel_time, const int* __restrict__ sel_box, int N) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(n_boxes[frame], box_cap);
    if (i >= n) return;
    const size_t fo = (size_t)frame * N;
    Box* b = boxes + (size_t)frame * box_cap + i;
    b->parent_box = first_box_above(loss_time + fo, win + fo, sel_time + fo, sel_box + fo, (u32)b->root, false);
}

// Label image: the smallest kept snapshot containing each pixel.  T = int32 (-1: none) or u16 (0xFFFF: none; a frame has
// fewer than 4096 boxes).  When run_count is given, every block also counts the label runs that start in its 256 pixels
// (a run starts where the label differs from the previous pixel's, in raster order) for the run-length output.
template <typename T>
DOFS_D T label_cast(int box) { return (T)box; }  // -1 -> 0xFFFF for u16

template <typename T>
__global__ void __launch_bounds__(SEG_THREADS)
k_labels(T* __restrict__ labels, const u32* __restrict__ loss_time, const u32* __restrict__ win,
         const u32* __restrict__ sel_time, const int* __restrict__ sel_box, int N, int* __restrict__ run_count, int run_blocks) {
    __shared__ int s_cnt[SEG_THREADS / 32];
    const int frame = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t fo = (size_t)frame * N;
    int lab = -2;
    if (p < N) {
        lab = first_box_above(loss_time + fo, win + fo, sel_time + fo, sel_box + fo, (u32)p, true);
        labels[fo + p] = label_cast<T>(lab);
    }
    if (!run_count) return;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    int prev = __shfl_up_sync(0xffffffffu, lab, 1);
    if (lane == 0) prev = (p > 0 && p < N) ? first_box_above(loss_time + fo, win + fo, sel_time + fo, sel_box + fo, (u32)(p - 1), true) : -2;
    const bool start = p < N && (p == 0 || prev != lab);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    if (lane == 0) s_cnt[wrp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < SEG_THREADS / 32; ++w) t += s_cnt[w];
        run_count[(size_t)frame * run_blocks + blockIdx.x] = t;
    }
}

// per frame: exclusive scan of the per-block run counts (in place) and the frame's number of runs
__global__ void __launch_bounds__(1024)
k_run_scan(int* __restrict__ run_count, int run_blocks, int* __restrict__ n_runs, int max_runs, int* __restrict__ sticky) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int frame = blockIdx.x;
    int* c = run_count + (size_t)frame * run_blocks;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < run_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < run_blocks ? c[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[wrp] = x;
        __syncthreads();
        if (wrp == 0) {
            int t = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            s_w[lane] = t;
        }
        __syncthreads();
        const int carry = s_carry;
        const int incl = x + (wrp ? s_w[wrp - 1] : 0);
        if (i < run_blocks) c[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        n_runs[frame] = s_carry;
        if (s_carry > max_runs) atomicOr(sticky, 8 /* STICKY_RUNS */);
    }
}

// run i of a frame = {first pixel, label}; it ends where run i+1 starts (the last one at W*H).  Runs are in raster order.
template <typename T, typename Run>
__global__ void __launch_bounds__(SEG_THREADS)
k_run_write(const T* __restrict__ labels, const int* __restrict__ run_base, int run_blocks, Run* __restrict__ runs, int max_runs,
            int N) {
    __shared__ int s_cnt[SEG_THREADS / 32];
    const int frame = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t fo = (size_t)frame * N;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    T lab = 0, prev = 0;
    if (p < N) {
        lab = labels[fo + p];
        if (p > 0) prev = labels[fo + p - 1];
    }
    const bool start = p < N && (p == 0 || prev != lab);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    if (lane == 0) s_cnt[wrp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wrp; ++w) before += s_cnt[w];
    if (start) {
        const int pos = run_base[(size_t)frame * run_blocks + blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
        if (pos < max_runs) {
            Run r;
            r.start = (u32)p;
            r.label = sizeof(T) == 2 ? (lab == (T)0xFFFF ? -1 : (int)lab) : (int)lab;
            runs[(size_t)frame * max_runs + pos] = r;
        }
    }
}

// dofs3d_lift: get_bottom_variants for n independent problems
template <typename Box>
__global__ void __launch_bounds__(128)
k_lift_problems(const float2* __restrict__ dir, const int4* __restrict__ bbox, const int* __restrict__ cls, int n,
                SegParams P, Box* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cls[i];
    const int4 b = bbox[i];
    LiftSolution s;
    lift_bottom_variants(dir[i].x, dir[i].y, b.x, b.y, b.z, b.w, P.hg.persp, P.hg.inv, P.hg.upper[c], c,
                         P.cls_size[c][0], P.cls_size[c][1], &s);
    Candidate cd;
    cd.root = 0;
    cd.time = 0;
    cd.size = s.has_rect;
    cd.fx = dir[i].x;
    cd.fy = dir[i].y;
    cd.bbox[0] = (u16)b.x;
    cd.bbox[1] = (u16)b.y;
    cd.bbox[2] = (u16)b.z;
    cd.bbox[3] = (u16)b.w;
    cd.pad = 0;
    fill_box(&out[i], cd, xddiv(xdadd(s.w_error, s.h_error), 2.0), s);
    if (!s.has_rect) out[i].cls = c;
}

// The boxes of a batch as one dense list (frame after frame, each frame's boxes in root order) for the gather of
// per-frame results over NCCL: one block per frame finds its offset from the counts of the frames before it.
template <typename Box>
__global__ void __launch_bounds__(128)
k_pack_boxes(const Box* __restrict__ boxes, const int* __restrict__ n_boxes, int box_cap, int n_frames, Box* __restrict__ out,
             int capacity, int* __restrict__ total_out) {
    const int frame = blockIdx.x;
    int before = 0, total = 0;
    for (int f = 0; f < n_frames; ++f) {
        const int c = min(n_boxes[f], box_cap);
        if (f < frame) before += c;
        total += c;
    }
    if (frame == 0 && threadIdx.x == 0) *total_out = total;
    const int mine = min(n_boxes[frame], box_cap);
    // a box is 27 8-byte words
    constexpr int WORDS = sizeof(Box) / 8;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(boxes + (size_t)frame * box_cap);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out + before);
    const int keep = max(0, min(mine, capacity - before));
    for (int i = threadIdx.x; i < keep * WORDS; i += blockDim.x) dst[i] = src[i];
}

// Predicates of the conditional graph nodes that skip work a batch does not need (dofs3d.cu: launch_conditional):
//   COND_LEVELS  some frame of the batch is not one component yet after level `arg - 1`
//   COND_WAVES   some frame has merge events in the waves arg+1 .. last-1 (the last wave is always replayed)
//   COND_FLAG    *flag != 0 (the exact fallback of the merge-time sort)
enum { COND_LEVELS = 0, COND_WAVES = 1, COND_FLAG = 2 };
__global__ void k_set_condition(cudaGraphConditionalHandle handle, int kind, const int* __restrict__ data, int n_frames, int F,
                                int arg, int last) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    int any = 0;
    if (kind == COND_FLAG) {
        any = threadIdx.x == 0 && *data != 0;
    } else {
        for (int f = threadIdx.x; f < n_frames; f += blockDim.x) {
            if (kind == COND_LEVELS) any |= data[(arg - 1) * F + f] != 1;                                   // n_roots[level][frame]
            else any |= data[f * (EV_MAX_WAVES + 1) + arg + 1] != data[f * (EV_MAX_WAVES + 1) + last];     // wave_start[frame][wave]
        }
    }
    if (any) s_any = 1;
    __syncthreads();
    if (threadIdx.x == 0) cudaGraphSetConditional(handle, s_any ? 1u : 0u);
}

// the largest number of Boruvka levels any frame of the call needed: the next call enqueues that many (+1) unconditionally
__global__ void k_call_levels(const int* __restrict__ levels, int n_frames, int* __restrict__ out) {
    int m = 0;
    for (int f = 0; f < n_frames; ++f) m = max(m, levels[f]);
    *out = m;
}

#define STICKY_INTERNAL 1    // Boruvka did not converge (non-finite flow) or a sort look-back timed out
#define STICKY_CANDIDATES 2  // candidate queue overflow
#define STICKY_BOXES 4       // more boxes than the caller's max_boxes / the box list
#define STICKY_RUNS 8        // more label runs than the caller's max_runs

// per-frame work counters -> the public stats record (include/dofs3d.h), on the device so that the
// whole call stays asynchronous
template <typename StatsT>
__global__ void k_stats(StatsT* __restrict__ out, BorState S, const int* __restrict__ n_cand, const int* __restrict__ n_scored,
                        const int* __restrict__ n_boxes, const int* __restrict__ longest_chain, int n_frames, int N,
                        int n_edges, int max_levels, const int* __restrict__ need_full, const int* __restrict__ replay_redone,
                        const int* __restrict__ sweep_timeout, int* __restrict__ sticky, int cand_cap, int box_cap, int max_boxes) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int levels = S.levels[f];
    StatsT st;
    st.n_edges = n_edges;
    st.n_levels = levels;
    const int roots = S.n_roots[(min(max(levels, 1), max_levels) - 1) * S.F + f];
    st.n_merges = N - roots;
    st.n_candidates = n_cand[f];
    st.n_scored = n_scored[f];
    st.n_boxes = n_boxes[f];
    st.longest_chain = longest_chain[f];
    st.final_root = (roots == 1 && *sweep_timeout == 0) ? S.final_root[f] : -1;  // -1: the call is reported as failed
    st.sort_fallback = *need_full;
    st.replay_exact_chunks = *replay_redone;
    out[f] = st;
    // deferred failures of asynchronous calls accumulate here until dofs3d_sync reports them (several calls may be
    // enqueued before one sync)
    int bad = 0;
    if (st.final_root < 0) bad |= STICKY_INTERNAL;
    if (st.n_candidates > cand_cap) bad |= STICKY_CANDIDATES;
    if (st.n_boxes > box_cap || (max_boxes >= 0 && st.n_boxes > max_boxes)) bad |= STICKY_BOXES;
    if (bad) atomicOr(sticky, bad);
}

// What the reference's display loop leaves in every pixel (draw.cpp:120-147): segments are painted in ascending root
// order when score > min_score, later ones over earlier ones — i.e. the containing box with the largest index wins.
// painted[p] = that box index or -1; bgr (optional) = the class colour draw.cpp uses (cls 0/2 (0,255,255), cls 1 (0,255,0)).
template <typename Box, typename T>
__global__ void __launch_bounds__(SEG_THREADS)
k_paint(const T* __restrict__ labels, const Box* __restrict__ boxes, int box_cap, int N, double min_score,
        int* __restrict__ painted, u8* __restrict__ bgr) {
    const int frame = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const Box* bx = boxes + (size_t)frame * box_cap;
    const T raw = labels[(size_t)frame * N + p];
    int b = sizeof(T) == 2 ? (raw == (T)0xFFFF ? -1 : (int)raw) : (int)raw, best = -1;
    while (b >= 0) {
        if (bx[b].score > min_score && b > best) best = b;
        b = bx[b].parent_box;
    }
    painted[(size_t)frame * N + p] = best;
    if (bgr && best >= 0) {
        u8* px = bgr + ((size_t)frame * N + p) * 3;
        const int cls = bx[best].cls;
        px[0] = 0;
        px[1] = 255;
        px[2] = cls == 1 ? 0 : 255;
    }
}
