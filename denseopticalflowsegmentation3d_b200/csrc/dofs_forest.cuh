// The reference's incremental Forest (cpp/inc/graph.hpp:72-114) on the device, ONE call at a time, for callers that
// drive the merge loop themselves the way the reference's segment_graph does (graph.cpp:520-531):
//   Forest::Forest      graph.cpp:129-148   N singleton sets: parent = id, rank 0, size 1, flow_value = flow[i], own pixel box
//   Forest::find        graph.cpp:150-157   root, full path compression
//   Forest::merge       graph.cpp:170-218   union by rank, size-weighted float mean, pixel-set union, box union
//   Forest::new_merge   graph.cpp:272-384   merge + size / row / move gates + get_score + convexity gate + history update
// Every call is a one-thread kernel over the forest's device state followed by a wait: tens of microseconds per call by
// construction.  The batch path (dofs3d_segment: Boruvka + chain replay) computes the same merge sequence for a whole
// Kruskal pass at once; this is the same arithmetic (merge_mean, replay_gate's conditions, lift_get_score) one merge at a
// time.  A set's pixels are an append-only linked list (the absorbed set's list is appended to the survivor's), so the
// snapshot the history keeps for a root is the first `snap_size` pixels of its list.
#pragma once
#include "dofs_common.cuh"
#include "dofs_lift.cuh"
#include "dofs_seg.cuh"

struct ForestState {
    int* parent;
    u8* rank;
    int* size;
    float2* flow;       // Node::flow_value
    ushort4* bbox;      // xmin, ymin, xmax, ymax; cleared (0xFFFF...) when the root is absorbed (graph.cpp:207)
    int* next;          // pixel list: next pixel of the set, -1 at the end
    int* tail;          // per root: last pixel of its list
    double* last_score; // segment_scores (graph.cpp:326)
    double* best_score; // segment_history[root].score, -1 = empty
    int* snap_size;     // |seg| of the kept snapshot
    int* box_slot;      // index of the root's record in `boxes`, -1 = none
    int* counters;      // [0] num_sets, [1] merges so far, [2] boxes used, [3] last call's result
    int W, H, N, box_cap;
};

__global__ void __launch_bounds__(256)
k_forest_init(ForestState S, const float2* __restrict__ flow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        S.counters[0] = S.N;
        S.counters[1] = 0;
        S.counters[2] = 0;
        S.counters[3] = 0;
    }
    if (i >= S.N) return;
    const int y = i / S.W, x = i - y * S.W;
    S.parent[i] = i;
    S.rank[i] = 0;
    S.size[i] = 1;
    S.flow[i] = flow[i];
    S.bbox[i] = make_ushort4((u16)x, (u16)y, (u16)x, (u16)y);
    S.next[i] = -1;
    S.tail[i] = i;
    S.last_score[i] = 0.0;
    S.best_score[i] = -1.0;
    S.snap_size[i] = 0;
    S.box_slot[i] = -1;
}

DOFS_D int forest_find(const ForestState& S, int n) {
    int r = n;
    while (S.parent[r] != r) r = S.parent[r];
    while (S.parent[n] != r) {  // full compression, like the recursive reference
        const int nx = S.parent[n];
        S.parent[n] = r;
        n = nx;
    }
    return r;
}

__global__ void k_forest_find(ForestState S, int n) { S.counters[3] = forest_find(S, n); }

// mode 0: Forest::merge; mode 1: Forest::new_merge
template <typename Box>
__global__ void k_forest_merge(ForestState S, int a, int b, int mode, double score_threshold, int min_size, SegParams P,
                               Box* __restrict__ boxes) {
    int pa = forest_find(S, a), pb = forest_find(S, b);
    if (pa != pb) {
        if (S.rank[pa] > S.rank[pb]) {
            const int t = pa;
            pa = pb;
            pb = t;
        }
        S.parent[pa] = pb;
        const int sa = S.size[pa], sb = S.size[pb];
        const float2 fa = S.flow[pa], fb = S.flow[pb];
        const double inv = xddiv(1.0, (double)(sa + sb));
        float2 m;
        m.x = merge_mean(xfmul(fa.x, (float)sa), fb.x, (float)sb, inv);
        m.y = merge_mean(xfmul(fa.y, (float)sa), fb.y, (float)sb, inv);
        S.flow[pb] = m;
        S.next[S.tail[pb]] = pa;  // segments[parent_b].insert(segments[parent_a])
        S.tail[pb] = S.tail[pa];
        S.size[pb] = sa + sb;
        S.size[pa] = 0;
        const ushort4 ba = S.bbox[pa], bb = S.bbox[pb];
        S.bbox[pb] = make_ushort4(min(ba.x, bb.x), min(ba.y, bb.y), max(ba.z, bb.z), max(ba.w, bb.w));
        S.bbox[pa] = make_ushort4(65535, 65535, 65535, 65535);  // bboxes[parent_a].clear()
        if (S.rank[pa] == S.rank[pb]) S.rank[pb] += 1;
        S.counters[0] -= 1;
        S.counters[1] += 1;
    }
    S.counters[3] = pb;
    if (mode == 0) return;
    // Forest::new_merge after its merge(a, b): the gates of graph.cpp:280-300 in their order
    const int s = S.size[pb];
    if (s < min_size) return;
    const int y = pb / S.W;
    if (y < S.H / 10) return;
    const float2 f = S.flow[pb];
    const double move = norm2d(f.x, f.y);
    if (move < xddiv((double)(3 * (y + 1)), (double)S.H)) return;
    const ushort4 bx = S.bbox[pb];
    const double rect_area = (double)((bx.z - bx.x + 1) * (bx.w - bx.y + 1));
    const double convexity = xddiv((double)s, rect_area);
    LiftSolution sol;
    const double score = lift_get_score(f.x, f.y, bx.x, bx.y, bx.z, bx.w, P, &sol);
    if (score == -1.0) return;
    S.last_score[pb] = score;                                  // graph.cpp:326
    if (convexity < P.cls_min_convexity[sol.cls]) return;      // graph.cpp:328-346
    if (!(score > score_threshold)) return;
    if (!(S.best_score[pb] < score)) return;                   // graph.cpp:352
    int slot = S.box_slot[pb];
    if (slot < 0) {
        slot = S.counters[2];
        if (slot >= S.box_cap) {
            S.counters[2] = slot + 1;  // reported as an overflow by the caller
            return;
        }
        S.counters[2] = slot + 1;
        S.box_slot[pb] = slot;
    }
    S.best_score[pb] = score;
    S.snap_size[pb] = s;
    Candidate c;
    c.root = (u32)pb;
    c.time = (u32)(S.counters[1] - 1);  // index of this merge among the merges so far
    c.size = s;
    c.fx = f.x;
    c.fy = f.y;
    c.bbox[0] = bx.x;
    c.bbox[1] = bx.y;
    c.bbox[2] = bx.z;
    c.bbox[3] = bx.w;
    c.pad = 0;
    fill_box(&boxes[slot], c, score, sol);
}

// the first `count` pixels of a root's list (its kept snapshot when count = snap_size[root])
__global__ void k_forest_pixels(ForestState S, int root, int count, int* __restrict__ out) {
    int p = root;
    for (int i = 0; i < count && p >= 0; ++i) {
        out[i] = p;
        p = S.next[p];
    }
}
