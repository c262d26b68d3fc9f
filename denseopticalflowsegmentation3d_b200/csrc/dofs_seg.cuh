// Flow-graph clustering on the device (stages K6-K12 of SURVEY.md section 2.2).
//
// The reference (cpp/src/graph.cpp) is a strictly sequential Kruskal loop with union by rank, a
// running float mean per set and a per-merge scoring hook.  This file computes the SAME merge
// sequence and per-merge state in parallel, using one structural fact (checked against the reference
// in tests/ and tools/proto_parallel.py):
//
//   Boruvka levels on the (weight, insertion-order)-ranked edges are exactly the union-by-rank ranks.
//   In level k every current component S (all have rank k) picks its minimum outgoing edge m_k(S).
//     * a pick that is not mutual: S loses at time m_k(S) to whatever component holds the other
//       endpoint at that time (it has rank > k);
//     * a mutual pick (S and S' pick the same edge): a rank tie; the component holding `edge.end`
//       survives (graph.cpp:177-182) and becomes a level k+1 component.
//   So every root id loses exactly once, at loss_time[c] (position of its edge in the sorted list),
//   and `up[c]` = root of the next-level component it is contracted into.  The root of any pixel's
//   set at time t is found by climbing `up` while loss_time < t (<= max rank hops).
//
// Merge events are then grouped into per-root chains ordered by time, and the chains are replayed
// in waves of increasing final rank (a chain only absorbs roots of strictly lower final rank), which
// reproduces sizes, bounding boxes and the order-dependent float mean flow (graph.cpp:184-190)
// bit for bit.
#pragma once
#include "dofs_common.cuh"
#include "dofs_lift.cuh"

#define SEG_THREADS 256

struct SegFrame {  // per-frame strided views: element i of frame f at [f * stride + i]
    int W, H, N;   // N = W*H pixels; 4N edge slots
};

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
// is already one component.  Kernels are grid-stride with a small fixed grid, so a level that has
// nothing left to do costs a few microseconds.
// ---------------------------------------------------------------------------------------------
struct BorState {
    u32* comp;       // [F][N] current component root of each pixel
    u32* best;       // [F][N] per root: minimum rank of an outgoing edge in this level
    u32* newp;       // [F][N] per root: hook target in this level
    u32* loss_time;  // [F][N] per root id: sorted position of the edge at which it loses (INF: never)
    u32* up;         // [F][N] per root id: root of the next-level component it is contracted into
    u8* lvl;         // [F][N] per root id: level at which it loses == its final union-find rank
    int* n_roots;    // [EV_MAX_WAVES][F] number of roots after each level (zeroed per call)
    int* levels;     // [F] number of levels the frame needed (written by k_bor_finish)
    int* final_root; // [F]
    int F;
};

#define EV_MAX_WAVES 32

DOFS_D bool bor_done(const BorState& S, int level, int frame) {
    return level > 0 && S.n_roots[(level - 1) * S.F + frame] == 1;
}

#define GRID_STRIDE(p, N) for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (N); p += gridDim.x * blockDim.x)

__global__ void __launch_bounds__(SEG_THREADS)
k_bor_init(BorState S, const float2* __restrict__ flow, int* __restrict__ rsize, ushort4* __restrict__ rbbox,
           float2* __restrict__ rflow, u64* __restrict__ best_score, u32* __restrict__ sel_time,
           int* __restrict__ sel_box, int W, int N) {
    const int frame = blockIdx.y;
    GRID_STRIDE(p, N) {
        const size_t g = (size_t)frame * N + p;
        S.comp[g] = (u32)p;
        S.best[g] = DOFS_INF32;
        S.newp[g] = (u32)p;
        S.loss_time[g] = DOFS_INF32;
        S.up[g] = (u32)p;
        S.lvl[g] = 0;
        // Forest::Forest (graph.cpp:129-148): singleton sets
        rsize[g] = 1;
        const int y = p / W, x = p - y * W;
        rbbox[g] = make_ushort4((u16)x, (u16)y, (u16)x, (u16)y);
        rflow[g] = flow[g];
        best_score[g] = 0ull;
        sel_time[g] = DOFS_INF32;
        sel_box[g] = -1;
    }
}

// every edge whose endpoints are in different components offers its rank to both components
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_pixel(BorState S, const u32* __restrict__ rank, size_t rank_stride, int W, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const u32* comp = S.comp + (size_t)frame * N;
    u32* best = S.best + (size_t)frame * N;
    GRID_STRIDE(p, N) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(rank + (size_t)frame * rank_stride + 4 * (size_t)p);
        const u32 r[4] = {r4.x, r4.y, r4.z, r4.w};
        const u32 cp = comp[p];
        u32 mine = DOFS_INF32;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (r[d] != DOFS_INF32) {
                u32 cq = comp[edge_other(p, d, W)];
                if (cq != cp) {
                    mine = min(mine, r[d]);
                    if (r[d] < best[cq]) atomicMin(&best[cq], r[d]);  // best only decreases: a stale read is safe
                }
            }
        }
        if (mine != DOFS_INF32 && mine < best[cp]) atomicMin(&best[cp], mine);
    }
}

// per root: classify its pick (mutual winner / loser), record the loss
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_root(BorState S, const u32* __restrict__ sorted_seq, size_t seq_stride, int W, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) return;
    const size_t fo = (size_t)frame * N;
    const u32* comp = S.comp + fo;
    GRID_STRIDE(c, N) {
        if (comp[c] != (u32)c) continue;
        const u32 t = S.best[fo + c];
        if (t == DOFS_INF32) {  // the last component
            S.newp[fo + c] = (u32)c;
            continue;
        }
        const u32 seq = sorted_seq[(size_t)frame * seq_stride + t];
        const int s = (int)(seq >> 2);
        const int e = edge_other(s, (int)(seq & 3u), W);
        const u32 cs = comp[s], ce = comp[e];
        const u32 other = (cs == (u32)c) ? ce : cs;
        const bool mutual = S.best[fo + other] == t;
        if (mutual && ce == (u32)c) {
            S.newp[fo + c] = (u32)c;  // rank tie: the `end` side survives (graph.cpp:177-182, 210-213)
        } else {
            S.newp[fo + c] = other;
            S.loss_time[fo + c] = t;
            S.lvl[fo + c] = (u8)level;
        }
    }
}

// contract: every pixel follows the hooks to the group root; losers remember it in `up`
__global__ void __launch_bounds__(SEG_THREADS)
k_bor_relabel(BorState S, int N, int level) {
    const int frame = blockIdx.y;
    if (bor_done(S, level, frame)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) S.n_roots[level * S.F + frame] = 1;
        return;
    }
    const size_t fo = (size_t)frame * N;
    volatile u32* newp = S.newp + fo;
    int roots = 0;
    GRID_STRIDE(p, N) {
        const u32 c = S.comp[fo + p];
        // follow the hooks to the group root with path halving: concurrent writers only ever replace a
        // pointer by one of its ancestors, so any interleaving still ends at the same root
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
        if (c == (u32)p) {  // p was a root in this level
            S.best[fo + p] = DOFS_INF32;
            if (g != (u32)p) S.up[fo + p] = g;
            else ++roots;
        }
        if (g != c) S.comp[fo + p] = g;
    }
    // one atomic per warp
    for (int o = 16; o > 0; o >>= 1) roots += __shfl_down_sync(0xffffffffu, roots, o);
    if ((threadIdx.x & 31) == 0 && roots) atomicAdd(&S.n_roots[level * S.F + frame], roots);
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
            S.final_root[frame] = p;  // if the frame did not converge several pixels land here; n_roots says so
            S.levels[frame] = levels;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K9b  winner of every merge event (one event per losing root) and the event sort key
//      key = lvl[winner] << (tb + wb) | winner << tb | time ; payload = loser
//      tb = bits of a sorted edge position (< 4N), wb = bits of a pixel id (< N): as few radix
//      passes as the frame size allows (1080p: 23 + 21 + 5 = 49 bits -> 7 passes instead of 8)
// ---------------------------------------------------------------------------------------------
#define EV_KEY_NONE 0xFFFFFFFFFFFFFFFFull

struct EvBits {
    int tb, wb;
};
DOFS_D u64 ev_chain(u64 key, EvBits b) { return key >> b.tb; }  // wave | winner
DOFS_D u32 ev_winner(u64 key, EvBits b) { return (u32)(key >> b.tb) & ((1u << b.wb) - 1u); }
DOFS_D u32 ev_time(u64 key, EvBits b) { return (u32)key & ((1u << b.tb) - 1u); }
DOFS_D int ev_wave(u64 key, EvBits b) { return key == EV_KEY_NONE ? EV_MAX_WAVES - 1 : (int)(key >> (b.tb + b.wb)); }

__global__ void __launch_bounds__(SEG_THREADS)
k_event_keys(BorState S, u32* __restrict__ win, u64* __restrict__ ev_key, int N, EvBits eb) {
    const int frame = blockIdx.y;
    const size_t fo = (size_t)frame * N;
    GRID_STRIDE(c, N) {
        const u32 t = S.loss_time[fo + c];
        if (t == DOFS_INF32) {
            win[fo + c] = (u32)c;
            ev_key[fo + c] = EV_KEY_NONE;
            continue;
        }
        u32 cur = S.up[fo + c];
        while (S.loss_time[fo + cur] < t) cur = S.up[fo + cur];
        win[fo + c] = cur;
        ev_key[fo + c] = ((u64)S.lvl[fo + cur] << (eb.tb + eb.wb)) | ((u64)cur << eb.tb) | (u64)t;
    }
}

// first event index of every wave (events are sorted by key)
__global__ void __launch_bounds__(SEG_THREADS)
k_wave_starts(const u64* __restrict__ ev_key, int* __restrict__ wave_start /* [F][EV_MAX_WAVES+1] */, int N,
              EvBits eb) {
    const int frame = blockIdx.y;
    const u64* k = ev_key + (size_t)frame * N;
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
// A chain = the events won by one root, in time order.  Chains of one wave are independent (every
// absorbed root has a lower final rank, so its state is final).  Two kernels per wave:
//   k_replay_short  one thread per chain of at most REPLAY_SHORT events; longer chains are pushed
//                   to a work list
//   k_replay_long   one warp per listed chain, 32 events at a time: the lanes gather the absorbed
//                   roots' states together (coalesced event reads, 32 gathers in flight, next chunk
//                   prefetched), sizes and boxes come from warp scans, and only the 3-operation
//                   float/double recurrence of the mean flow runs serially, fed by shuffles; each
//                   lane then applies the gates to "its" event.
// ---------------------------------------------------------------------------------------------
#define REPLAY_SHORT 32

struct ReplayArgs {
    const u64* ev_key;     // [F][N] sorted
    const u32* ev_loser;   // [F][N] sorted payload
    const int* wave_start; // [F][EV_MAX_WAVES+1]
    int* rsize;            // [F][N]
    ushort4* rbbox;        // [F][N]
    float2* rflow;         // [F][N]
    Candidate* cand;       // [F][cand_cap]
    int* n_cand;           // [F]
    int* longest_chain;    // [F]
    uint2* long_list;      // [list_cap] (frame, index of the chain's first event)
    int* long_count;       // [EV_MAX_WAVES + 1]
    int list_cap;
    int cand_cap;
    int W, H, N;
    int min_size;
    EvBits eb;
};

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

__global__ void __launch_bounds__(SEG_THREADS)
k_replay_short(ReplayArgs A, int wave) {
    const int frame = blockIdx.y;
    const int w0 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave];
    const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
    const size_t fo = (size_t)frame * A.N;
    const u64* key = A.ev_key + fo;
    int longest = 0;
    for (int i = w0 + blockIdx.x * blockDim.x + threadIdx.x; i < w1; i += gridDim.x * blockDim.x) {
        const u64 k0 = key[i];
        const u64 chain = ev_chain(k0, A.eb);  // wave | winner
        if (i > w0 && ev_chain(key[i - 1], A.eb) == chain) continue;  // not the head of its chain
        if (i + REPLAY_SHORT < w1 && ev_chain(key[i + REPLAY_SHORT], A.eb) == chain) {  // long chain: a warp takes it
            const int slot = atomicAdd(&A.long_count[wave], 1);
            if (slot < A.list_cap) A.long_list[slot] = make_uint2((u32)frame, (u32)i);
            continue;
        }
        const u32 r = ev_winner(k0, A.eb);
        int s = A.rsize[fo + r];
        float2 f = A.rflow[fo + r];
        ushort4 bb = A.rbbox[fo + r];
        const int y = (int)r / A.W;
        const bool row_ok = !(y < A.H / 10);                                  // graph.cpp:288
        const double move_min = xddiv((double)(3 * (y + 1)), (double)A.H);    // graph.cpp:296
        int j = i;
        u64 kj = k0;
        for (;;) {
            const u32 a = A.ev_loser[fo + j];
            const int sa = A.rsize[fo + a];
            const float2 fa = A.rflow[fo + a];
            const ushort4 ba = A.rbbox[fo + a];
            const float fsa = (float)sa, fsb = (float)s;
            const double inv = xddiv(1.0, (double)(sa + s));
            f.x = merge_mean(xfmul(fa.x, fsa), f.x, fsb, inv);
            f.y = merge_mean(xfmul(fa.y, fsa), f.y, fsb, inv);
            s += sa;
            bb.x = min(bb.x, ba.x);
            bb.y = min(bb.y, ba.y);
            bb.z = max(bb.z, ba.z);
            bb.w = max(bb.w, ba.w);
            if (s >= A.min_size && row_ok) {
                const double move = norm2d(f.x, f.y);
                if (!(move < move_min)) push_candidate(A, frame, r, ev_time(kj, A.eb), s, f, bb);
            }
            ++j;
            if (j >= w1) break;
            kj = key[j];
            if (ev_chain(kj, A.eb) != chain) break;
        }
        A.rsize[fo + r] = s;
        A.rflow[fo + r] = f;
        A.rbbox[fo + r] = bb;
        longest = max(longest, j - i);
    }
    if (longest) atomicMax(&A.longest_chain[frame], longest);
}

// REPLAY_Q chunks of 32 events form one software-pipeline stage: while stage t is replayed, the gathers
// of stage t+1 and the event reads of stage t+2 are in flight (nothing is unpacked or tested before
// its use, so the warp never waits at the issue point of a load).
#define REPLAY_Q 4
#define REPLAY_WARPS 4
#define REPLAY_G 8

struct ReplayEvent {  // stage 0: straight reads of the sorted event arrays
    u64 key;
    u32 a;
};
struct ReplayOperand {  // stage 1: state of the absorbed root, raw
    int sa;
    float2 fa;
    uint2 ba;  // ushort4 bounding box, still packed
};

DOFS_D ReplayEvent replay_read(const ReplayArgs& A, size_t fo, int j, int w1, u32 r) {
    ReplayEvent e;
    e.key = EV_KEY_NONE;
    e.a = r;  // any valid index
    if (j < w1) {
        e.key = A.ev_key[fo + j];
        e.a = A.ev_loser[fo + j];
    }
    return e;
}

DOFS_D ReplayOperand replay_gather(const ReplayArgs& A, size_t fo, u32 a) {
    ReplayOperand o;
    o.sa = A.rsize[fo + a];
    o.fa = A.rflow[fo + a];
    o.ba = *reinterpret_cast<const uint2*>(&A.rbbox[fo + a]);
    return o;
}

__global__ void __launch_bounds__(32 * REPLAY_WARPS, 1)
k_replay_long(ReplayArgs A, int wave) {
    __shared__ float4 s_op[REPLAY_WARPS][32];   // (loser mean * loser size).xy, size before the event, unused
    __shared__ double s_inv[REPLAY_WARPS][32];  // 1 / size after the event
    __shared__ float2 s_f[REPLAY_WARPS][32];    // mean flow after the event
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_list = min(A.long_count[wave], A.list_cap);
    const unsigned FULL = 0xffffffffu;
    for (int item = warp; item < n_list; item += n_warps) {
        const uint2 it = A.long_list[item];
        const int frame = (int)it.x, i0 = (int)it.y;
        const size_t fo = (size_t)frame * A.N;
        const int w1 = A.wave_start[frame * (EV_MAX_WAVES + 1) + wave + 1];
        const u64 k0 = A.ev_key[fo + i0];
        const u64 chain = ev_chain(k0, A.eb);
        const u32 r = ev_winner(k0, A.eb);
        int s = A.rsize[fo + r];
        float2 f = A.rflow[fo + r];
        ushort4 bb = A.rbbox[fo + r];
        const int y = (int)r / A.W;
        const bool row_ok = !(y < A.H / 10);
        const double move_min = xddiv((double)(3 * (y + 1)), (double)A.H);
        int j0 = i0;
        bool more = true;
        ReplayEvent ev_cur[REPLAY_Q], ev_nxt[REPLAY_Q], ev_far[REPLAY_Q];
        ReplayOperand op_cur[REPLAY_Q], op_nxt[REPLAY_Q];
#pragma unroll
        for (int q = 0; q < REPLAY_Q; ++q) ev_nxt[q] = replay_read(A, fo, j0 + 32 * q + lane, w1, r);
#pragma unroll
        for (int q = 0; q < REPLAY_Q; ++q) ev_far[q] = replay_read(A, fo, j0 + 32 * (REPLAY_Q + q) + lane, w1, r);
#pragma unroll
        for (int q = 0; q < REPLAY_Q; ++q) op_nxt[q] = replay_gather(A, fo, ev_nxt[q].a);
        while (more) {
#pragma unroll
            for (int q = 0; q < REPLAY_Q; ++q) {
                ev_cur[q] = ev_nxt[q];
                op_cur[q] = op_nxt[q];
                ev_nxt[q] = ev_far[q];
            }
#pragma unroll
            for (int q = 0; q < REPLAY_Q; ++q) op_nxt[q] = replay_gather(A, fo, ev_nxt[q].a);
#pragma unroll
            for (int q = 0; q < REPLAY_Q; ++q) ev_far[q] = replay_read(A, fo, j0 + 32 * (2 * REPLAY_Q + q) + lane, w1, r);
#pragma unroll
            for (int q = 0; q < REPLAY_Q; ++q) {
                const bool valid = ev_cur[q].key != EV_KEY_NONE && ev_chain(ev_cur[q].key, A.eb) == chain;
                const unsigned vmask = __ballot_sync(FULL, valid);  // valid lanes are a prefix (events are sorted)
                const int n_valid = __popc(vmask);
                if (n_valid == 0) {
                    more = false;
                    break;
                }
                const int sa = valid ? op_cur[q].sa : 0;
                const float2 fa = op_cur[q].fa;
                // sizes and boxes after every event: inclusive warp scans
                int s_inc = sa;
                int bx0 = valid ? (int)(op_cur[q].ba.x & 0xffffu) : 65535, by0 = valid ? (int)(op_cur[q].ba.x >> 16) : 65535;
                int bx1 = valid ? (int)(op_cur[q].ba.y & 0xffffu) : 0, by1 = valid ? (int)(op_cur[q].ba.y >> 16) : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, s_inc, o);
                    const int t0 = __shfl_up_sync(FULL, bx0, o), t1 = __shfl_up_sync(FULL, by0, o);
                    const int t2 = __shfl_up_sync(FULL, bx1, o), t3 = __shfl_up_sync(FULL, by1, o);
                    if (lane >= o) {
                        s_inc += t;
                        bx0 = min(bx0, t0);
                        by0 = min(by0, t1);
                        bx1 = max(bx1, t2);
                        by1 = max(by1, t3);
                    }
                }
                const int s_after = s + s_inc, s_before = s_after - sa;
                ushort4 bb_after;
                bb_after.x = (u16)min((int)bb.x, bx0);
                bb_after.y = (u16)min((int)bb.y, by0);
                bb_after.z = (u16)max((int)bb.z, bx1);
                bb_after.w = (u16)max((int)bb.w, by1);
                const float fsa = (float)sa;
                // operands of the recurrence, broadcast through shared memory
                s_op[wib][lane] = make_float4(xfmul(fa.x, fsa), xfmul(fa.y, fsa), (float)s_before, 0.f);
                s_inv[wib][lane] = xddiv(1.0, (double)max(s_after, 1));
                __syncwarp();
                // the serial part: mean flow after each event; every lane runs the same recurrence
                if (n_valid == 32) {
                    // groups of REPLAY_G events: operands of the next group are read from shared memory before the
                    // results of this group are stored, so no step waits for a shared-memory round trip
                    float4 o4[REPLAY_G];
                    double kv[REPLAY_G];
#pragma unroll
                    for (int k = 0; k < REPLAY_G; ++k) {
                        o4[k] = s_op[wib][k];
                        kv[k] = s_inv[wib][k];
                    }
#pragma unroll
                    for (int g = 0; g < 32 / REPLAY_G; ++g) {
                        float4 n4[REPLAY_G];
                        double nv[REPLAY_G];
                        if (g + 1 < 32 / REPLAY_G) {
#pragma unroll
                            for (int k = 0; k < REPLAY_G; ++k) {
                                n4[k] = s_op[wib][(g + 1) * REPLAY_G + k];
                                nv[k] = s_inv[wib][(g + 1) * REPLAY_G + k];
                            }
                        }
                        float2 fh[REPLAY_G];
#pragma unroll
                        for (int k = 0; k < REPLAY_G; ++k) {
                            f.x = merge_mean(o4[k].x, f.x, o4[k].z, kv[k]);
                            f.y = merge_mean(o4[k].y, f.y, o4[k].z, kv[k]);
                            fh[k] = f;
                        }
#pragma unroll
                        for (int k = 0; k < REPLAY_G; ++k) s_f[wib][g * REPLAY_G + k] = fh[k];
                        if (g + 1 < 32 / REPLAY_G) {
#pragma unroll
                            for (int k = 0; k < REPLAY_G; ++k) {
                                o4[k] = n4[k];
                                kv[k] = nv[k];
                            }
                        }
                    }
                } else {
                    for (int k = 0; k < n_valid; ++k) {
                        const float4 o4 = s_op[wib][k];
                        const double kinv = s_inv[wib][k];
                        f.x = merge_mean(o4.x, f.x, o4.z, kinv);
                        f.y = merge_mean(o4.y, f.y, o4.z, kinv);
                        s_f[wib][k] = f;
                    }
                }
                __syncwarp();
                // gates of the event this lane holds
                if (valid && s_after >= A.min_size && row_ok) {
                    const float2 mine = s_f[wib][lane];
                    const double move = norm2d(mine.x, mine.y);
                    if (!(move < move_min))
                        push_candidate(A, frame, r, ev_time(ev_cur[q].key, A.eb), s_after, mine, bb_after);
                }
                __syncwarp();
                // carry = state after the last valid event
                const int last = n_valid - 1;
                s = __shfl_sync(FULL, s_after, last);
                bb.x = (u16)__shfl_sync(FULL, (int)bb_after.x, last);
                bb.y = (u16)__shfl_sync(FULL, (int)bb_after.y, last);
                bb.z = (u16)__shfl_sync(FULL, (int)bb_after.z, last);
                bb.w = (u16)__shfl_sync(FULL, (int)bb_after.w, last);
                j0 += n_valid;
                if (n_valid < 32) {
                    more = false;
                    break;
                }
            }
        }
        if (lane == 0) {
            A.rsize[fo + r] = s;
            A.rflow[fo + r] = f;
            A.rbbox[fo + r] = bb;
            atomicMax(&A.longest_chain[frame], j0 - i0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K11 + K12  lifting of the queued merges, per-root first-maximum selection, box records, labels
// ---------------------------------------------------------------------------------------------
struct SelectArgs {
    const Candidate* cand;  // [F][cand_cap]
    const int* n_cand;      // [F]
    double* cand_score;     // [F][cand_cap]  score if the merge passed every gate, else -1
    u64* best_score;        // [F][N] bit pattern of the best score per root (0 = none)
    u32* sel_time;          // [F][N] time of the first merge reaching the best score
    int* sel_box;           // [F][N] index of the root's box in the sorted box list
    int* n_scored;          // [F]
    int* n_boxes;           // [F]
    int cand_cap;
    int N;
};

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
    double kept = -1.0;
    if (score != -1.0) {
        const double min_convexity = P.cls_min_convexity[sol.cls];
        if (!(convexity < min_convexity) && score > P.score_threshold) kept = score;
    }
    A.cand_score[(size_t)frame * A.cand_cap + i] = kept;
    if (kept > 0.0) {
        atomicMax((unsigned long long*)&A.best_score[(size_t)frame * A.N + c.root],
                  (unsigned long long)__double_as_longlong(kept));
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
    const double sc = A.cand_score[(size_t)frame * A.cand_cap + i];
    if (!(sc > 0.0)) return;
    const Candidate c = A.cand[(size_t)frame * A.cand_cap + i];
    if ((u64)__double_as_longlong(sc) == A.best_score[(size_t)frame * A.N + c.root])
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
    const double sc = A.cand_score[(size_t)frame * A.cand_cap + i];
    if (!(sc > 0.0)) return;
    const Candidate c = A.cand[(size_t)frame * A.cand_cap + i];
    const size_t g = (size_t)frame * A.N + c.root;
    if ((u64)__double_as_longlong(sc) != A.best_score[g] || c.time != A.sel_time[g]) return;
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
             int* __restrict__ sel_box, int N) {
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
}

// Walk the union-find link forest (parent = winner at loss time, <= max-rank hops): a pixel / set that
// entered root a's set at time t_in belongs to a's snapshot taken at sel_time[a] iff t_in <= sel_time[a].
DOFS_D int first_box_above(const u32* loss_time, const u32* win, const u32* sel_time, const int* sel_box, u32 cur,
                           bool include_self) {
    u32 t_in = 0;
    bool first = include_self;
    if (!include_self) {
        t_in = loss_time[cur];
        if (t_in == DOFS_INF32) return -1;
        cur = win[cur];
    }
    for (;;) {
        const u32 st = sel_time[cur];
        if (st != DOFS_INF32 && (first || st >= t_in)) return sel_box[cur];
        first = false;
        t_in = loss_time[cur];
        if (t_in == DOFS_INF32) return -1;
        cur = win[cur];
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

__global__ void __launch_bounds__(SEG_THREADS)
k_labels(int* __restrict__ labels, const u32* __restrict__ loss_time, const u32* __restrict__ win,
         const u32* __restrict__ sel_time, const int* __restrict__ sel_box, int N) {
    const int frame = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const size_t fo = (size_t)frame * N;
    labels[fo + p] = first_box_above(loss_time + fo, win + fo, sel_time + fo, sel_box + fo, (u32)p, true);
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

// per-frame work counters -> the public stats record (include/dofs3d.h), on the device so that the
// whole call stays asynchronous
template <typename StatsT>
__global__ void k_stats(StatsT* __restrict__ out, BorState S, const int* __restrict__ n_cand, const int* __restrict__ n_scored,
                        const int* __restrict__ n_boxes, const int* __restrict__ longest_chain, int n_frames, int N,
                        int n_edges, int max_levels, const int* __restrict__ need_full) {
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
    st.final_root = roots == 1 ? S.final_root[f] : -1;
    st.sort_fallback = *need_full;
    st.pad_ = 0;
    out[f] = st;
}
