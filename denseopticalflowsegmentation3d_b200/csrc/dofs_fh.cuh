// Felzenszwalb-style flow segmentation: the adaptive-threshold mode of the reference's Python twin
// (/root/reference/graph.py), SURVEY.md section 8f.4:
//   build_graph             graph.py:77-96    edges L, U, UL, DL per pixel in raster order (the slot order of K7)
//   diff / threshold        main.py:310-315   weight = sqrt(sum((a - b)^2)) in float32; threshold(size) = K / size
//   segment_graph_flow      graph.py:156-177  sorted(edges) (stable), then three passes over the sorted list:
//       1. merge when w <= thr[a] and w <= thr[b]; thr[root] = w + K / size            (:163-172)
//       2. remove_small_components: merge when either side has fewer than min_size pixels   (:98-106)
//       3. merge_components: merge when |mean flow a - mean flow b| < 5 and w < 5              (:108-130)
//   Forest.find / merge     graph.py:27-61    union by rank (rank[a] > rank[b] keeps a, else b), float32 size-weighted mean
// Arithmetic is the reference's under NumPy 2 promotion rules (see oracle/fh_oracle.cpp, which restates it on the CPU and
// is pinned to outputs of the reference's own Python).
//
// Unlike the reference's C++ path, what this loop accepts is NOT the minimum spanning forest: whether an edge merges
// depends on thresholds that every earlier accepted merge rewrites, so the Boruvka reformulation of dofs_seg.cuh does not
// apply and the edge list is really walked in order.  What is parallel: the weights and the stable radix sort of all 4N
// slots (k_fh_edge_keys + the one-sweep sort), and inside the walk a warp takes 32 consecutive edges at a time — every lane
// chases the roots of its edge (the dependent loads that dominate a scalar walk) and evaluates its test; an edge is settled
// in that round — merged or rejected for good — unless an earlier edge of the batch that touches one of its components
// merges in the round or is itself held back (merges of disjoint components commute); the others look again in the next
// round, so the outcome is the sequential one.
#pragma once
#include "dofs_common.cuh"
#include "dofs_seg.cuh"

#define FH_KEY_INVALID 0xFFFFFFFFu

// float32 weight of main.py's diff on a float32 field: bits of a non-negative float order like unsigned integers
DOFS_D u32 fh_edge_key(float2 a, float2 b) {
    const float dx = xfsub(a.x, b.x), dy = xfsub(a.y, b.y);
    const float w = __fsqrt_rn(xfadd(xfmul(dx, dx), xfmul(dy, dy)));
    const u32 k = __float_as_uint(w);
    return k < 0x7F800000u ? k : FH_KEY_INVALID;  // a non-finite weight is no edge
}

__global__ void __launch_bounds__(SEG_THREADS)
k_fh_edge_keys(const float2* __restrict__ f, u32* __restrict__ keys, int W, int H, int neighbors8) {
    const int N = W * H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int y = p / W, x = p - y * W;
    const float2 c = f[p];
    u32 k0 = FH_KEY_INVALID, k1 = FH_KEY_INVALID, k2 = FH_KEY_INVALID, k3 = FH_KEY_INVALID;
    if (x > 0) k0 = fh_edge_key(c, f[p - 1]);
    if (y > 0) k1 = fh_edge_key(c, f[p - W]);
    if (neighbors8) {
        if (x > 0 && y > 0) k2 = fh_edge_key(c, f[p - W - 1]);
        if (x > 0 && y < H - 1) k3 = fh_edge_key(c, f[p + W - 1]);
    }
    *reinterpret_cast<uint4*>(keys + 4 * (size_t)p) = make_uint4(k0, k1, k2, k3);
}

struct FhState {
    int* parent;    // [N]
    u8* rank;       // [N]
    int* size;      // [N]
    float2* color;  // [N] mean flow of the set (Node.color)
    double* thr;    // [N] threshold of the set: K / 1 (a Python float) until its first merge, a float32 afterwards
};

__global__ void __launch_bounds__(SEG_THREADS)
k_fh_init(FhState S, const float2* __restrict__ flow, int N, double K) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    S.parent[p] = p;
    S.rank[p] = 0;
    S.size[p] = 1;
    S.color[p] = flow[p];
    S.thr[p] = K;
}

DOFS_D int fh_root(volatile int* parent, int x) {
    for (;;) {  // path halving: every write replaces a pointer by one of its ancestors, benign under concurrency
        const int px = parent[x];
        if (px == x) return x;
        const int gp = parent[px];
        if (gp == px) return px;
        parent[x] = gp;
        x = gp;
    }
}

// Forest.merge (graph.py:44-61) of roots a (of the edge's first endpoint) and b; returns the surviving root.  All state is
// read and written through volatile pointers: other lanes of the warp read it in the next round.
DOFS_D int fh_merge(const FhState& S, int a, int b) {
    volatile u8* rank = S.rank;
    volatile int* size = S.size;
    volatile float* col = reinterpret_cast<volatile float*>(S.color);
    const int ra = rank[a], rb = rank[b];
    const int keep = ra > rb ? a : b, gone = ra > rb ? b : a;
    const int sa = size[a], sb = size[b];
    const float ax = col[2 * a], ay = col[2 * a + 1], bx = col[2 * b], by = col[2 * b + 1];
    const float tot = (float)(sa + sb);
    col[2 * keep] = xfdiv(xfadd(xfmul((float)sa, ax), xfmul((float)sb, bx)), tot);
    col[2 * keep + 1] = xfdiv(xfadd(xfmul((float)sa, ay), xfmul((float)sb, by)), tot);
    size[keep] = sa + sb;
    if (!(ra > rb) && ra == rb) rank[b] = (u8)(rb + 1);
    reinterpret_cast<volatile int*>(S.parent)[gone] = keep;
    return keep;
}

// One warp walks the sorted edge list of one frame, 32 edges per step; pass = 1, 2, 3 as listed at the top of the file.
__global__ void __launch_bounds__(32)
k_fh_walk(FhState S, const u32* __restrict__ sorted_key, const u32* __restrict__ sorted_slot, int n_edges, int W, int pass,
          double K, int min_size, double flow_dist, double edge_dist) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    volatile int* parent = S.parent;
    volatile int* vsize = S.size;
    volatile double* vthr = S.thr;
    volatile float* vcol = reinterpret_cast<volatile float*>(S.color);
    for (int base = 0; base < n_edges; base += 32) {
        const int k = base + lane;
        int ra = 0, rb = 0;
        float w = 0.f;
        bool pending = false;
        if (k < n_edges) {
            const u32 slot = sorted_slot[k];
            const int s = (int)(slot >> 2);
            w = __uint_as_float(sorted_key[k]);
            ra = fh_root(parent, s);
            rb = fh_root(parent, edge_other(s, (int)(slot & 3u), W));
            pending = ra != rb;
        }
        unsigned m_pending = __ballot_sync(FULL, pending);
        while (m_pending) {
            bool want = false;
            if (pending) {  // roots may have moved in the previous round
                ra = fh_root(parent, ra);
                rb = fh_root(parent, rb);
                if (ra == rb) pending = false;
                else if (pass == 1) want = (double)w <= vthr[ra] && (double)w <= vthr[rb];
                else if (pass == 2) want = vsize[ra] < min_size || vsize[rb] < min_size;
                else {
                    const float dx = xfsub(vcol[2 * ra], vcol[2 * rb]), dy = xfsub(vcol[2 * ra + 1], vcol[2 * rb + 1]);
                    const float d = __fsqrt_rn(xfadd(xfmul(dx, dx), xfmul(dy, dy)));
                    want = (double)d < flow_dist && (double)w < edge_dist;
                }
            }
            // In list order: an edge is settled this round (merged, or rejected for good) unless an earlier edge of the
            // batch that touches one of its components merges now or is itself held back — then it must look again
            // after that edge.  Lane j's status is final when the loop reaches j: it only depends on lanes below j.
            bool blocked = false;
            unsigned mm = __ballot_sync(FULL, pending);
            while (mm) {
                const int j = __ffs(mm) - 1;
                mm &= mm - 1;
                const int ja = __shfl_sync(FULL, ra, j), jb = __shfl_sync(FULL, rb, j);
                const int act = __shfl_sync(FULL, (int)(want || blocked), j);
                if (act && j < lane && pending && (ja == ra || ja == rb || jb == ra || jb == rb)) blocked = true;
            }
            __syncwarp();  // every decision of the round was read before any merge of the round is written
            if (pending && !blocked) {
                if (want) {
                    const int r = fh_merge(S, ra, rb);
                    if (pass == 1) vthr[r] = (double)xfadd(w, (float)(K * 1.0 / (double)vsize[r]));
                }
                pending = false;
            }
            __threadfence_block();
            __syncwarp();
            m_pending = __ballot_sync(FULL, pending);
        }
    }
}

// labels[p] = Forest.find(p); n_components = number of roots
__global__ void __launch_bounds__(SEG_THREADS)
k_fh_labels(FhState S, int* __restrict__ labels, int* __restrict__ n_components, int N) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    int x = p;
    while (S.parent[x] != x) x = S.parent[x];
    labels[p] = x;
    if (x == p) atomicAdd(n_components, 1);
}
