// dofs3d.cu — context, orchestration and the C ABI (include/dofs3d.h) of the sm_100a hot path.
//
// One context per GPU.  A call processes a batch of up to max_pairs independent frame pairs; every
// kernel takes the frame index from blockIdx.y, so one launch covers the whole batch and the grid is
// many multiples of the 148 SMs.  All per-frame arrays are [frame][element] in HBM.
//
// There is no CPU path in this file: without a CUDA device dofs3d_create fails.
#include "../../include/dofs3d.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "dofs_bev.cuh"
#include "dofs_common.cuh"
#include "dofs_draw.cuh"
#include "dofs_flow.cuh"
#include "dofs_lift.cuh"
#include "dofs_fh.cuh"
#include "dofs_forest.cuh"
#include "dofs_seg.cuh"
#include "dofs_sort.cuh"
#include "dofs_synth.cuh"

static_assert(sizeof(dofs3d_box) == 216, "dofs3d_box layout");

namespace {

// ---------------------------------------------------------------------------------------------
// host-side setup arithmetic (once per context)
// ---------------------------------------------------------------------------------------------

// cv::getPerspectiveTransform as the reference uses it (lifting_3d.cpp:479,510-511): the 8x8 system
// of the four point correspondences solved by LU with partial pivoting in double, M[8] = 1, result
// rounded to float.
void perspective_transform(const double src[4][2], const double dst[4][2], float out[9]) {
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
        const double x = src[i][0], y = src[i][1], u = dst[i][0], v = dst[i][1];
        const double top[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
        const double bot[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
        memcpy(a[i], top, sizeof top);
        memcpy(a[i + 4], bot, sizeof bot);
    }
    for (int col = 0; col < 8; ++col) {
        int piv = col;
        for (int r = col + 1; r < 8; ++r)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (piv != col)
            for (int c = col; c < 9; ++c) std::swap(a[col][c], a[piv][c]);
        const double d = -1.0 / a[col][col];
        for (int r = col + 1; r < 8; ++r) {
            const double f = a[r][col] * d;
            for (int c = col + 1; c < 9; ++c) a[r][c] += f * a[col][c];
        }
    }
    double sol[8];
    for (int r = 7; r >= 0; --r) {
        double s = a[r][8];
        for (int c = r + 1; c < 8; ++c) s -= a[r][c] * sol[c];
        sol[r] = s / a[r][r];
    }
    for (int i = 0; i < 8; ++i) out[i] = (float)sol[i];
    out[8] = 1.0f;
}

// cv::getGaussianKernel(ksize, sigma, CV_32F): taps rounded to float, normalised by the double sum
// of the rounded taps.
void gaussian_taps(int ksize, double sigma, float* out) {
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

int ceil_log2(unsigned long long v) {
    int b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

struct StageTimer {
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    struct Row {
        std::string name;
        float ms;
        int count;
    };
    std::vector<Row> result;
    bool enabled = false;
};

}  // namespace

struct dofs3d_ctx {
    int device = 0, W = 0, H = 0, N = 0, F = 0;
    size_t S = 0;  // edge slots per frame = 4N
    dofs3d_params prm;
    SegParams seg;
    BlurTaps taps;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;
    long long bytes = 0;
    std::vector<void*> allocs;
    StageTimer timer;

    // segmentation buffers
    float2 *flow_in = nullptr, *flow_tmp = nullptr, *flow_blur = nullptr;
    u64 *keysA = nullptr, *keysB = nullptr;
    u32 *valsA = nullptr, *valsB = nullptr;
    u32 *tile_hist = nullptr, *digit_tot = nullptr;
    u32* sweep_hist = nullptr;   // [2][F][RS_MAX_PASSES][256] digit histograms and bases of the one-sweep sorts
    int* sweep_ticket = nullptr; // [RS_MAX_PASSES] tile tickets + [1] look-back time-out flag
    int num_tiles = 0;
    BorState bor;
    u32* win = nullptr;
    int* wave_start = nullptr;
    uint2* long_list = nullptr;  // chains longer than REPLAY_SHORT events, per wave
    int* long_count = nullptr;
    int* repair_flags = nullptr;  // [0] number of long prefix runs, [1] need the full 64-bit edge sort
    // per merge event (chain replay)
    float4* ev_op = nullptr;
    double* ev_inv = nullptr;
    int *ev_size = nullptr, *ev_sa = nullptr;
    ushort4* ev_bbox = nullptr;
    float2 *ev_prod = nullptr, *ev_flow = nullptr;
    TileAgg *tile_agg = nullptr, *tile_carry = nullptr;
    int tiles_cap = 0;
    int list_cap = 0;
    RootState* rstate = nullptr;
    u64* best_score = nullptr;
    u32* sel_time = nullptr;
    int* sel_box = nullptr;
    Candidate* cand = nullptr;
    double* cand_score = nullptr;
    int cand_cap = 0, box_cap = 0;
    int* counters = nullptr;    // [6][F]: n_cand, longest_chain, n_scored, n_boxes, n_roots, final_root
    int max_levels = 0;               // guaranteed bound of Boruvka levels for this frame size
    bool force_time_fallback = false; // test knob (DOFS3D_FORCE_TIME_FALLBACK=1): always take the exact 64-bit fallback of K8
    dofs3d_box *boxes_tmp = nullptr, *boxes = nullptr;
    int32_t* labels = nullptr;        // [F][N] int32, or [F][N] u16 in the same buffer (labels_fmt of the last call)
    int labels_fmt = DOFS3D_LABELS_I32;
    int* run_count = nullptr;         // [F][run_blocks] label runs starting in every block of 256 pixels, then their scan
    int* n_runs = nullptr;            // [F]
    int run_blocks = 0;
    dofs3d_run* runs = nullptr;       // [F][runs_cap], allocated on the first run-length call
    int runs_cap = 0;
    dofs3d_stats* stats = nullptr;
    // Conditional graph nodes (CUDA 12.4+): the Boruvka levels and replay waves beyond what the previous call needed, and
    // the exact fallback of the merge-time sort, are enqueued as the body of an IF node whose condition a one-block kernel
    // sets from device state — a batch that does not need them pays one graph launch instead of ~150 empty kernels, a
    // batch that does runs exactly the launches it would have got.  Cached per (kind, batch size, first level).
    struct CondGraph {
        int kind, n, arg;
        cudaGraph_t graph;
        cudaGraphExec_t exec;
    };
    std::vector<CondGraph> cond_graphs;
    int soft_levels_forced = 0;       // test knob DOFS3D_SOFT_LEVELS: levels enqueued unconditionally (the IF bodies then run)
    bool use_cond_graphs = true;      // DOFS3D_COND_GRAPHS=0: enqueue every launch unconditionally (round-1 behaviour)
    int* d_call_levels = nullptr;     // device: levels the last call needed
    int* h_call_levels = nullptr;     // pinned copy, read (possibly one call late: it is only a hint) by the next call
    int stride_blocks_per_sm = 16;    // grid of the grid-stride kernels (DOFS3D_STRIDE_BLOCKS): blocks per SM over the whole batch
    int carveout = -1;                // DOFS3D_CARVEOUT: preferred shared-memory carveout (percent) of every kernel; -1 = the driver's choice
    std::vector<const void*> carveout_done;
    bool bor_fold = false;            // A/B knob: DOFS3D_BOR_FOLD=1 folds the Boruvka relabel pass into the pixel kernel
    bool blur_tma = true;             // A/B knob: DOFS3D_BLUR_TMA=0 stages the blur tiles with LDG -> STS instead of bulk copies
    u8* render_seg = nullptr;         // [F][N][3] the painted copy of dofs3d_render, allocated on first use
    double* rcp_table = nullptr;      // [RCP_TABLE] 1.0 / n for the replay of small sets
    int* sticky = nullptr;            // device: STICKY_* bits of every call since the last dofs3d_sync
    int* h_sticky = nullptr;          // pinned copy, refreshed after every call

    // streaming (dofs3d_stream_*): copy stream, two staging buffers, the outstanding chunks
    cudaStream_t copy_stream = nullptr;
    u8* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};   // recorded when the gray conversion has consumed the buffer
    cudaEvent_t stage_ready[2] = {nullptr, nullptr};  // recorded when the upload has landed
    struct Chunk {
        int n_pairs;
        cudaEvent_t done;
    };
    std::vector<Chunk> chunks;        // outstanding, oldest first (at most 2)
    long long chunks_submitted = 0;
    bool have_carry = false;          // gray slot 0 and fb.R_carry hold the last frame of the previous chunk

    // flow buffers
    FlowBuffers fb;
    u8 *bgr = nullptr, *gray = nullptr;  // [F+1] frames
    float2* flow_raw = nullptr;          // [F][N] Farneback output
};

namespace {

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return e_ == cudaErrorMemoryAllocation ? DOFS3D_ERR_NOMEM : DOFS3D_ERR_CUDA;               \
        }                                                                                             \
    } while (0)

// Every kernel of a context can be given the same preferred shared-memory carveout (DOFS3D_CARVEOUT=<percent>): with
// several contexts interleaving kernels on one GPU, kernels that ask for different L1 / shared-memory splits make the SMs
// drain and reconfigure between them.  Applied once per kernel and context.
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
    do {                                                                   \
        if ((ctx)->carveout >= 0) prefer_carveout((ctx), (const void*)(kernel)); \
        kernel<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);         \
        (ctx)->launches++;                                                 \
    } while (0)

void prefer_carveout(dofs3d_ctx* ctx, const void* kernel) {
    if (std::find(ctx->carveout_done.begin(), ctx->carveout_done.end(), kernel) != ctx->carveout_done.end()) return;
    ctx->carveout_done.push_back(kernel);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->carveout);
}

template <typename T>
int dalloc(dofs3d_ctx* ctx, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CK(cudaMalloc(&q, bytes));
    ctx->allocs.push_back(q);
    ctx->bytes += (long long)bytes;
    *p = static_cast<T*>(q);
    return 0;
}

#define DA(ptr, count)                               \
    do {                                             \
        int r_ = dalloc(ctx, &(ptr), (count));       \
        if (r_) return r_;                           \
    } while (0)

inline dim3 grid1(size_t n, int threads, int frames) { return dim3((unsigned)((n + threads - 1) / threads), frames); }
// small fixed grid for the grid-stride kernels: about 16 blocks per SM over the whole batch
inline dim3 grid_stride(const dofs3d_ctx* ctx, int frames) {
    const unsigned total = 148u * (unsigned)ctx->stride_blocks_per_sm;
    const unsigned per_frame = std::max(1u, std::min(total / (unsigned)frames, (unsigned)((ctx->N + SEG_THREADS - 1) / SEG_THREADS)));
    return dim3(per_frame, frames);
}

void mark(dofs3d_ctx* ctx, const char* name) {
    if (!ctx->timer.enabled) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, ctx->stream);
    ctx->timer.marks.push_back({name, ev});
}

void timer_begin(dofs3d_ctx* ctx) {
    for (auto& m : ctx->timer.marks) cudaEventDestroy(m.second);
    ctx->timer.marks.clear();
    mark(ctx, "begin");
}

void timer_collect(dofs3d_ctx* ctx) {
    if (!ctx->timer.enabled) return;
    cudaStreamSynchronize(ctx->stream);
    ctx->timer.result.clear();
    for (size_t i = 1; i < ctx->timer.marks.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->timer.marks[i - 1].second, ctx->timer.marks[i].second);
        bool found = false;
        for (auto& r : ctx->timer.result)
            if (r.name == ctx->timer.marks[i].first) {
                r.ms += ms;
                r.count += 1;
                found = true;
            }
        if (!found) ctx->timer.result.push_back({ctx->timer.marks[i].first, ms, 1});
    }
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of `frames` independent arrays of n (u64 key, u32 payload) pairs on key bits
// [0, key_bits).  Input in (kA, vA) (payload = index when iota), ping-pong with (kB, vB).
// Returns 0 if the result is in A, 1 if in B.
// ---------------------------------------------------------------------------------------------
template <typename K>
int radix_sort(dofs3d_ctx* ctx, K* kA, u32* vA, K* kB, u32* vB, size_t stride, int n, int frames, int key_bits, bool iota,
               const char* tag_hist, const char* tag_scan, const char* tag_scatter, const int* enable = nullptr) {
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    // a conditional sort is enqueued with a small persistent grid: it costs next to nothing when disabled
    const int gx = enable ? std::min(tiles, std::max(1, 148 * 2 / frames)) : tiles;
    int side = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        const K* kin = side ? kB : kA;
        const u32* vin = side ? vB : vA;
        K* kout = side ? kA : kB;
        u32* vout = side ? vA : vB;
        LAUNCH(ctx, k_radix_hist<K>, dim3(gx, frames), RS_THREADS, 0, kin, stride, ctx->tile_hist, n, shift, tiles, enable);
        mark(ctx, tag_hist);
        LAUNCH(ctx, k_radix_scan, dim3(RS_BINS, frames), RS_THREADS, 0, ctx->tile_hist, ctx->digit_tot, tiles, enable);
        mark(ctx, tag_scan);
        LAUNCH(ctx, k_radix_scatter<K>, dim3(gx, frames), RS_THREADS, rs_smem_bytes<K>(), kin, vin, kout, vout, stride,
               ctx->tile_hist, ctx->digit_tot, n, shift, tiles, (iota && shift == 0) ? 1 : 0, enable);
        mark(ctx, tag_scatter);
        side ^= 1;
    }
    return side;
}

// the same sort in one-sweep form: one histogram pass for all digits, then one scatter kernel per digit whose tiles
// find their offsets by decoupled look-back (dofs_sort.cuh).  Used by the unconditional sorts.
template <typename K>
int radix_sort_onesweep(dofs3d_ctx* ctx, K* kA, u32* vA, K* kB, u32* vB, size_t stride, int n, int frames, int key_bits,
                        bool iota, const char* tag_hist, const char* tag_scatter) {
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    const int passes = (key_bits + 7) / 8;
    if (passes > RS_MAX_PASSES) return -1;
    u32* ghist = ctx->sweep_hist;
    u32* gbase = ghist + (size_t)ctx->F * RS_MAX_PASSES * RS_BINS;
    cudaMemsetAsync(ghist, 0, sizeof(u32) * (size_t)frames * RS_MAX_PASSES * RS_BINS, ctx->stream);
    cudaMemsetAsync(ctx->sweep_ticket, 0, sizeof(int) * (RS_MAX_PASSES + 1), ctx->stream);
    LAUNCH(ctx, k_radix_hist_all<K>, dim3(std::max(1, std::min(tiles, 148 * 8 / frames)), frames), RS_THREADS, 0, kA, stride, ghist,
           n, passes);
    LAUNCH(ctx, k_radix_bases, dim3(passes, frames), RS_THREADS, 0, ghist, gbase);
    mark(ctx, tag_hist);
    int side = 0;
    for (int p = 0; p < passes; ++p) {
        const K* kin = side ? kB : kA;
        const u32* vin = side ? vB : vA;
        K* kout = side ? kA : kB;
        u32* vout = side ? vA : vB;
        cudaMemsetAsync(ctx->tile_hist, 0, sizeof(u32) * (size_t)frames * tiles * RS_BINS, ctx->stream);  // status words
        LAUNCH(ctx, k_radix_onesweep<K>, dim3((unsigned)tiles * frames), RS_THREADS, rs_smem_bytes<K>(), kin, vin, kout, vout,
               stride, gbase + (size_t)p * RS_BINS, ctx->tile_hist, ctx->sweep_ticket + p, ctx->sweep_ticket + RS_MAX_PASSES, n,
               8 * p, tiles, frames, (iota && p == 0) ? 1 : 0);
        mark(ctx, tag_scatter);
        side ^= 1;
    }
    return side;
}

int n_edges_of(int W, int H, int neighbors) {
    return neighbors == 8 ? 4 * W * H - 3 * W - 3 * H + 2 : 2 * W * H - W - H;
}

// K7 + K8: blurred flow (ctx->flow_blur) -> weights by slot in keysA, sorted sequence numbers in valsA, rank of every
// slot in valsB.  4 radix passes on a 32-bit order-preserving prefix of the weight, exact repair of the runs that share
// a prefix, and — only if a frame has a long run the repair kernels cannot order — the full 64-bit sort.
int build_sorted_edges(dofs3d_ctx* ctx, int n) {
    const int N = ctx->N;
    const int S = (int)ctx->S;
    u32* preA = reinterpret_cast<u32*>(ctx->keysB);  // the prefix buffers live in keysB, which only the fallback needs
    u32* preB = preA + (size_t)n * ctx->S;  // n == 1 (the parity hook): both fit the S u64 of keysB
    LAUNCH(ctx, k_edge_keys, grid1(N, SEG_THREADS, n), SEG_THREADS, 0, ctx->flow_blur, ctx->keysA, preA, ctx->S, ctx->W,
           ctx->H, ctx->seg.neighbors == 8 ? 1 : 0);
    mark(ctx, "edge_keys");
    int side = radix_sort_onesweep<u32>(ctx, preA, ctx->valsA, preB, ctx->valsB, ctx->S, S, n, 32, true, "edge_sort.hist",
                                        "edge_sort.scatter");
    if (side != 0) {
        ctx->err = "internal: odd number of sort passes";
        return DOFS3D_ERR_INTERNAL;
    }
    CK(cudaMemsetAsync(ctx->repair_flags, 0, 2 * sizeof(int), ctx->stream));
    RepairArgs R;
    R.prefix = preA;
    R.seq = ctx->valsA;
    R.keys = ctx->keysA;
    R.rank = ctx->valsB;
    R.long_list = ctx->long_list;
    R.long_count = ctx->repair_flags;
    R.need_full = ctx->repair_flags + 1;
    R.list_cap = ctx->list_cap;
    R.stride = ctx->S;
    R.n_slots = S;
    LAUNCH(ctx, k_prefix_repair_short, grid1(ctx->S, SEG_THREADS, n), SEG_THREADS, 0, R);
    LAUNCH(ctx, k_prefix_repair_long, dim3(148 * 2), 256, 0, R);
    mark(ctx, "edge_repair");
    // fallback, enabled on the device by *need_full
    const int* enable = ctx->repair_flags + 1;
    side = radix_sort<u64>(ctx, ctx->keysA, ctx->valsA, ctx->keysB, ctx->valsB, ctx->S, S, n, 64, true, "edge_fallback",
                           "edge_fallback", "edge_fallback", enable);
    LAUNCH(ctx, k_rank_scatter, dim3(std::max(1, 148 * 4 / n), n), SEG_THREADS, 0, ctx->valsA, ctx->valsB, ctx->S, S,
           n_edges_of(ctx->W, ctx->H, ctx->seg.neighbors), enable);
    mark(ctx, "edge_fallback");
    return 0;
}

enum { CNT_CAND = 0, CNT_CHAIN = 1, CNT_SCORED = 2, CNT_BOXES = 3, CNT_ROOTS = 4, CNT_FINAL = 5, CNT_KINDS = 6 };

// GaussianBlur(flow, sigma) (segment.cpp:52): the fused tile kernel for the reference's 25 taps, two passes otherwise
void blur_launch(dofs3d_ctx* ctx, const float2* src, float2* dst, int n) {
    const int W = ctx->W, H = ctx->H;
    if (ctx->taps.radius == 12) {
        const dim3 grid((W + FB_T - 1) / FB_T, (H + FB_T - 1) / FB_T, n);
        // TMA-staged tiles need 16-byte aligned rows of float2: an even width and an aligned base
        if (ctx->blur_tma && (W & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
            LAUNCH(ctx, k_blur_fused_tma<12>, grid, 256, 0, src, dst, W, H, ctx->taps);
        else
            LAUNCH(ctx, k_blur_fused<12>, grid, 256, 0, src, dst, W, H, ctx->taps);
    } else {
        const dim3 gN = grid1(ctx->N, SEG_THREADS, n);
        LAUNCH(ctx, k_blur_rows, gN, SEG_THREADS, 0, src, ctx->flow_tmp, W, H, ctx->taps);
        LAUNCH(ctx, k_blur_cols, gN, SEG_THREADS, 0, ctx->flow_tmp, dst, W, H, ctx->taps);
    }
}

// get_segmented_array (segment.cpp:34-72) for n frames whose (unblurred or blurred) flow is at d_flow.
// Enqueues `body` as the body of an IF node: kind / data / arg / last are k_set_condition's arguments.  The graph is built
// once per (kind, n, arg) by capturing the body's launches from the context's stream; kernel arguments of a context are
// the same on every call (its buffers), so the instantiated graph is replayed as it is.
template <typename Body>
int launch_conditional(dofs3d_ctx* ctx, int kind, const int* data, int n, int arg, int last, Body body) {
    if (!ctx->use_cond_graphs) {
        body();
        return 0;
    }
    for (const auto& g : ctx->cond_graphs)
        if (g.kind == kind && g.n == n && g.arg == arg) {
            CK(cudaGraphLaunch(g.exec, ctx->stream));
            ctx->launches++;
            return 0;
        }
    dofs3d_ctx::CondGraph cg;
    cg.kind = kind;
    cg.n = n;
    cg.arg = arg;
    CK(cudaGraphCreate(&cg.graph, 0));
    cudaGraphConditionalHandle handle;
    CK(cudaGraphConditionalHandleCreate(&handle, cg.graph, 0, cudaGraphCondAssignDefault));
    cudaGraphNode_t set_node, if_node;
    cudaKernelNodeParams kp = {};
    int F = ctx->F;
    void* args[] = {&handle, &kind, &data, &n, &F, &arg, &last};
    kp.func = (void*)k_set_condition;
    kp.gridDim = dim3(1);
    kp.blockDim = dim3(128);
    kp.kernelParams = args;
    CK(cudaGraphAddKernelNode(&set_node, cg.graph, nullptr, 0, &kp));
    cudaGraphNodeParams cp = {};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeIf;
    cp.conditional.size = 1;
    CK(cudaGraphAddNode(&if_node, cg.graph, &set_node, 1, &cp));
    const bool timing = ctx->timer.enabled;
    ctx->timer.enabled = false;  // no event records inside the captured body
    const long long launches_before = ctx->launches;
    CK(cudaStreamBeginCaptureToGraph(ctx->stream, cp.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    body();
    cudaGraph_t captured = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &captured);
    ctx->timer.enabled = timing;
    ctx->launches = launches_before;
    if (e != cudaSuccess) {
        ctx->err = std::string("stream capture of a conditional body failed: ") + cudaGetErrorString(e);
        return DOFS3D_ERR_CUDA;
    }
    CK(cudaGraphInstantiate(&cg.exec, cg.graph, 0));
    ctx->cond_graphs.push_back(cg);
    CK(cudaGraphLaunch(cg.exec, ctx->stream));
    ctx->launches++;
    return 0;
}

// what a call must leave in the context for export_results: label format, run-length capacity, the box capacity the
// caller announced (for the deferred overflow check)
struct OutSpec {
    int fmt = DOFS3D_LABELS_I32;
    int max_runs = 0;
    int max_boxes = -1;
};

int ensure_runs(dofs3d_ctx* ctx, int max_runs);

int segment_dev(dofs3d_ctx* ctx, const float* d_flow, int already_blurred, int n, const OutSpec& spec = OutSpec()) {
    const int N = ctx->N, W = ctx->W, H = ctx->H, F = ctx->F;
    const dim3 gN = grid1(N, SEG_THREADS, n);
    const float2* src = reinterpret_cast<const float2*>(d_flow);
    if (already_blurred) {
        CK(cudaMemcpyAsync(ctx->flow_blur, src, (size_t)n * N * sizeof(float2), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        blur_launch(ctx, src, ctx->flow_blur, n);
    }
    mark(ctx, "flow_blur");

    // K7: one 32-bit order-preserving prefix of the weight per edge slot (lives in keysB until the events need it)
    u32* prefix = reinterpret_cast<u32*>(ctx->keysB);
    LAUNCH(ctx, k_edge_prefix, gN, SEG_THREADS, 0, ctx->flow_blur, prefix, ctx->S, W, H, ctx->seg.neighbors == 8 ? 1 : 0, fastdiv_magic((u32)ctx->W));
    mark(ctx, "edge_keys");

    // K9a Boruvka levels (== union-by-rank ranks) on direct edge comparisons: the guaranteed bound of levels is
    // enqueued, finished frames skip
    BorState& B = ctx->bor;
    const dim3 gS = grid_stride(ctx, n);
    // (the per-root arrays are written by the level-0 kernels; the selection state of a root is reset when it queues a
    // merge, k_select_reset)
    CK(cudaMemsetAsync(ctx->counters, 0, sizeof(int) * CNT_KINDS * F, ctx->stream));
    CK(cudaMemsetAsync(B.n_roots, 0, sizeof(int) * EV_MAX_WAVES * F, ctx->stream));
    const int levels = ctx->max_levels;
    // levels the previous call needed + 1 are enqueued unconditionally, the rest of the guaranteed bound as an IF body
    int soft = levels;
    if (ctx->use_cond_graphs) soft = std::min(levels, std::max(2, (*ctx->h_call_levels > 0 ? *ctx->h_call_levels : 11) + 1));
    if (ctx->use_cond_graphs && ctx->soft_levels_forced > 0) soft = std::min(levels, std::max(1, ctx->soft_levels_forced));
    auto bor_level = [&](int level) {
        if (level == 0) {
            LAUNCH(ctx, k_bor_level0_pick, gS, SEG_THREADS, 0, B, prefix, ctx->S, ctx->flow_blur, W, H, N);
            LAUNCH(ctx, k_bor_level0_root, gS, SEG_THREADS, 0, B, W, H, N, ctx->seg.neighbors == 8 ? 1 : 0);
        } else {
            if (level >= BOR_COMPACT_FROM && bor_pixel_packed_ok(N))
                LAUNCH(ctx, k_bor_pixel_packed, gS, SEG_THREADS, 0, B, prefix, ctx->S, ctx->flow_blur, W, N, level, ctx->bor_fold ? 1 : 0);
            else
                LAUNCH(ctx, k_bor_pixel, gS, SEG_THREADS, 0, B, prefix, ctx->S, ctx->flow_blur, W, N, level, ctx->bor_fold ? 1 : 0);
            LAUNCH(ctx, k_bor_root, gS, SEG_THREADS, 0, B, W, N, level);
        }
        LAUNCH(ctx, k_bor_contract, gS, SEG_THREADS, 0, B, N, level);
        if ((level == 0 && !BOR_L0_COMP) || (level > 0 && !ctx->bor_fold)) LAUNCH(ctx, k_bor_relabel, gS, SEG_THREADS, 0, B, N, level);
    };
    for (int level = 0; level < soft; ++level) bor_level(level);
    if (soft < levels) {
        // (a skipped level never writes its n_roots entry: the entries of the call were zeroed above, and k_bor_finish /
        // k_stats read "1" at the first level that reports one component, which is at or below `soft - 1` then)
        int rc = launch_conditional(ctx, COND_LEVELS, B.n_roots, n, soft, levels, [&] {
            for (int level = soft; level < levels; ++level) bor_level(level);
        });
        if (rc) return rc;
    }
    LAUNCH(ctx, k_bor_finish, gS, SEG_THREADS, 0, B, N, levels);
    mark(ctx, "boruvka");

    // K8 merge times: only the edges Boruvka picked are sorted (<= N-1 of the 4N slots), on a 32-bit key
    u32* tkA = reinterpret_cast<u32*>(ctx->keysA);  // keysA holds F*4N u64: two u32 key buffers, then two u64 fallback buffers
    u32* tkB = tkA + (size_t)F * N;
    u32* tvA = ctx->valsA;
    u32* tvB = ctx->valsB;
    LAUNCH(ctx, k_time_keys, gS, SEG_THREADS, 0, B, prefix, ctx->S, tkA, N);
    mark(ctx, "time_keys");
    int side = radix_sort_onesweep<u32>(ctx, tkA, tvA, tkB, tvB, (size_t)N, N, n, 32, true, "time_sort.hist", "time_sort.scatter");
    if (side < 0) {
        ctx->err = "internal: time key too wide";
        return DOFS3D_ERR_INTERNAL;
    }
    const u32* tk = side ? tkB : tkA;  // sorted keys
    u32* order = side ? tvB : tvA;     // losing roots in key order
    u32* order_other = side ? tvA : tvB;
    u32* times = B.newp;               // the hook targets are dead: per root id, the position of its loss in the merge sequence
    CK(cudaMemsetAsync(ctx->repair_flags, 0, 2 * sizeof(int), ctx->stream));
    if (ctx->force_time_fallback) CK(cudaMemsetAsync(ctx->repair_flags + 1, 1, 1, ctx->stream));
    {
        TimeRepairArgs R;
        R.key = tk;
        R.comp = order;
        R.slot = B.loss_time;
        R.flow = ctx->flow_blur;
        R.time = times;
        R.long_list = ctx->long_list;
        R.long_count = ctx->repair_flags;
        R.need_full = ctx->repair_flags + 1;
        R.list_cap = ctx->list_cap;
        R.N = N;
        R.W = W;
        LAUNCH(ctx, k_time_repair_short, gN, SEG_THREADS, 0, R);
        LAUNCH(ctx, k_time_repair_long, dim3(148 * 2), 256, 0, R);
        mark(ctx, "time_repair");
        // fallback, enabled on the device by *need_full: stable 64-bit sorts by slot, then by weight.  Its 36 launches
        // are the body of an IF node on that flag (launch_conditional): a batch without a long tie run skips them.
        const int* enable = ctx->repair_flags + 1;
        const dim3 gF(std::max(1, 148 * 4 / n), n);
        u64* fA = ctx->keysA + (size_t)F * N;
        u64* fB = ctx->keysA + 2 * (size_t)F * N;
        int rc_fb = launch_conditional(ctx, COND_FLAG, enable, n, 0, 0, [&] {
            LAUNCH(ctx, k_time_fallback_slots, gF, SEG_THREADS, 0, order, B.loss_time, fA, N, enable);
            int fs = radix_sort<u64>(ctx, fA, order, fB, order_other, (size_t)N, N, n, ceil_log2(ctx->S), false, "time_fallback",
                                     "time_fallback", "time_fallback", enable);
            u32* o1 = fs ? order_other : order;  // roots in slot order
            u32* o1_other = fs ? order : order_other;
            LAUNCH(ctx, k_time_fallback_weights, gF, SEG_THREADS, 0, o1, B.loss_time, ctx->flow_blur, fA, N, W, enable);
            fs = radix_sort<u64>(ctx, fA, o1, fB, o1_other, (size_t)N, N, n, 64, false, "time_fallback", "time_fallback",
                                 "time_fallback", enable);
            LAUNCH(ctx, k_time_fallback_rank, gF, SEG_THREADS, 0, fs ? fB : fA, fs ? o1_other : o1, times, order, N, enable);
        });
        if (rc_fb) return rc_fb;
        mark(ctx, "time_fallback");
    }
    BorState BT = B;  // from here on loss_time means time
    BT.loss_time = times;

    // K9b events: winner of every loss, in (wave, winner, time) order = a stable 32-bit sort by (wave, winner) of the
    // time-ordered list of losing roots.  Buffers alias the dead edge prefixes (keysB holds 4 F N u32).
    EvBits eb;
    eb.wb = ceil_log2((unsigned long long)N);
    if (eb.wb + 5 > 31) {
        ctx->err = "internal: event key too wide";
        return DOFS3D_ERR_INTERNAL;
    }
    u32* evR = reinterpret_cast<u32*>(ctx->keysB);  // key by root id
    u32* evA = evR + (size_t)F * N;
    u32* evB = evA + (size_t)F * N;
    LAUNCH(ctx, k_event_keys, gS, SEG_THREADS, 0, BT, ctx->win, evR, N, eb);
    LAUNCH(ctx, k_event_gather, gS, SEG_THREADS, 0, order, evR, evA, N);
    mark(ctx, "event_keys");
    side = radix_sort_onesweep<u32>(ctx, evA, order, evB, order_other, (size_t)N, N, n, eb.wb + 5, false, "event_sort.hist",
                                    "event_sort.scatter");
    if (side < 0) {
        ctx->err = "internal: event key too wide";
        return DOFS3D_ERR_INTERNAL;
    }
    const u32* ev_key = side ? evB : evA;
    const u32* ev_loser = side ? order_other : order;
    LAUNCH(ctx, k_wave_starts, gS, SEG_THREADS, 0, ev_key, ctx->wave_start, N, eb);
    mark(ctx, "event_waves");

    // K9c + K10 chain replay, wave by wave
    ReplayArgs R;
    R.ev_key = ev_key;
    R.ev_loser = ev_loser;
    R.time = times;
    R.wave_start = ctx->wave_start;
    R.rstate = ctx->rstate;
    R.flow = ctx->flow_blur;
    R.lvl = B.lvl;
    R.cand = ctx->cand;
    R.n_cand = ctx->counters + CNT_CAND * F;
    R.longest_chain = ctx->counters + CNT_CHAIN * F;
    R.cand_cap = ctx->cand_cap;
    R.W = W;
    R.H = H;
    R.N = N;
    R.wm = fastdiv_magic((u32)W);
    R.min_size = ctx->seg.min_size;
    R.eb = eb;
    R.long_list = ctx->long_list;
    R.long_count = ctx->long_count;
    R.list_cap = ctx->list_cap;
    CK(cudaMemsetAsync(ctx->long_count, 0, sizeof(int) * (EV_MAX_WAVES + 1), ctx->stream));
    R.long_flag = B.mask;  // the edge masks are dead after the Boruvka levels
    CK(cudaMemsetAsync(R.long_flag, 0, (size_t)n * N, ctx->stream));
    R.ev_op = ctx->ev_op;
    R.ev_inv = ctx->ev_inv;
    R.ev_size = ctx->ev_size;
    R.ev_bbox = ctx->ev_bbox;
    R.ev_prod = ctx->ev_prod;
    R.ev_sa = ctx->ev_sa;
    R.ev_flow = ctx->ev_flow;
    R.tile_agg = ctx->tile_agg;
    R.tile_carry = ctx->tile_carry;
    R.tiles_cap = ctx->tiles_cap;
    R.rcp = ctx->rcp_table;
    auto replay_wave = [&](int wave) {
        LAUNCH(ctx, k_replay_short, gS, SEG_THREADS, 0, R, wave);
        LAUNCH(ctx, k_replay_scan, gS, REPLAY_TILE, 0, R, wave);
        LAUNCH(ctx, k_replay_carry, dim3(n), 32, 0, R, wave);
        LAUNCH(ctx, k_replay_operands, gS, SEG_THREADS, 0, R, wave);
        LAUNCH(ctx, k_replay_serial_long, dim3(148 * 4), 32 * REPLAY_WARPS, 0, R, wave);
        LAUNCH(ctx, k_replay_gates, gS, SEG_THREADS, 0, R, wave);
    };
    // a root's wave is the level at which it loses, so the frames of the previous call's depth have no events beyond wave
    // `soft`; the waves between it and the last one (the final roots' chains) are an IF body like the late levels
    const int soft_wave = std::min(soft, levels - 1);
    for (int wave = 1; wave <= soft_wave; ++wave) replay_wave(wave);
    if (soft_wave + 1 <= levels - 1) {
        int rc = launch_conditional(ctx, COND_WAVES, ctx->wave_start, n, soft_wave, levels, [&] {
            for (int wave = soft_wave + 1; wave <= levels - 1; ++wave) replay_wave(wave);
        });
        if (rc) return rc;
    }
    replay_wave(levels);
    mark(ctx, "chain_replay");

    // K11 + K12 lifting, selection, boxes, labels
    SelectArgs A;
    A.cand = ctx->cand;
    A.n_cand = ctx->counters + CNT_CAND * F;
    A.cand_score = ctx->cand_score;
    A.best_score = ctx->best_score;
    A.sel_time = ctx->sel_time;
    A.sel_box = ctx->sel_box;
    A.n_scored = ctx->counters + CNT_SCORED * F;
    A.n_boxes = ctx->counters + CNT_BOXES * F;
    A.cand_cap = ctx->cand_cap;
    A.N = N;
    const dim3 gC128 = grid1(ctx->cand_cap, 128, n);
    const dim3 gC = grid1(ctx->cand_cap, SEG_THREADS, n);
    LAUNCH(ctx, k_select_reset, gC, SEG_THREADS, 0, A);
    LAUNCH(ctx, k_lift_score, gC128, 128, 0, A, ctx->seg);
    LAUNCH(ctx, k_select_time, gC, SEG_THREADS, 0, A);
    LAUNCH(ctx, k_emit_boxes<dofs3d_box>, gC128, 128, 0, A, ctx->seg, ctx->boxes_tmp, ctx->box_cap);
    mark(ctx, "lifting");
    const dim3 gB = grid1(ctx->box_cap, SEG_THREADS, n);
    LAUNCH(ctx, k_sort_boxes<dofs3d_box>, gB, SEG_THREADS, 0, ctx->boxes_tmp, ctx->boxes, A.n_boxes, ctx->box_cap,
           ctx->sel_box, ctx->win, N);
    LAUNCH(ctx, k_box_parents<dofs3d_box>, gB, SEG_THREADS, 0, ctx->boxes, A.n_boxes, ctx->box_cap, BT.loss_time,
           ctx->win, ctx->sel_time, ctx->sel_box, N);
    ctx->labels_fmt = spec.fmt == DOFS3D_LABELS_I32 ? DOFS3D_LABELS_I32 : DOFS3D_LABELS_U16;
    if (spec.fmt == DOFS3D_LABELS_I32) {
        LAUNCH(ctx, k_labels<int>, gN, SEG_THREADS, 0, ctx->labels, BT.loss_time, ctx->win, ctx->sel_time, ctx->sel_box, N,
               (int*)nullptr, 0);
    } else {
        const bool rle = spec.fmt == DOFS3D_LABELS_RLE;
        if (rle) {
            int rc = ensure_runs(ctx, spec.max_runs);
            if (rc) return rc;
        }
        u16* lab16 = reinterpret_cast<u16*>(ctx->labels);
        LAUNCH(ctx, k_labels<u16>, gN, SEG_THREADS, 0, lab16, BT.loss_time, ctx->win, ctx->sel_time, ctx->sel_box, N,
               rle ? ctx->run_count : (int*)nullptr, ctx->run_blocks);
        if (rle) {
            LAUNCH(ctx, k_run_scan, dim3(n), 1024, 0, ctx->run_count, ctx->run_blocks, ctx->n_runs, spec.max_runs, ctx->sticky);
            LAUNCH(ctx, (k_run_write<u16, dofs3d_run>), gN, SEG_THREADS, 0, lab16, ctx->run_count, ctx->run_blocks, ctx->runs,
                   ctx->runs_cap, N);
        }
    }
    mark(ctx, "labels");

    // counters -> stats record, on the device; a copy lands in pinned host memory for the host-pointer entry points
    LAUNCH(ctx, k_stats<dofs3d_stats>, dim3((n + 63) / 64), 64, 0, ctx->stats, B, ctx->counters + CNT_CAND * F,
           ctx->counters + CNT_SCORED * F, ctx->counters + CNT_BOXES * F, ctx->counters + CNT_CHAIN * F, n, N,
           n_edges_of(W, H, ctx->seg.neighbors), levels, ctx->repair_flags + 1, ctx->long_count,
           ctx->sweep_ticket + RS_MAX_PASSES, ctx->sticky, ctx->cand_cap, ctx->box_cap, spec.max_boxes);
    CK(cudaMemcpyAsync(ctx->h_sticky, ctx->sticky, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LAUNCH(ctx, k_call_levels, 1, 1, 0, B.levels, n, ctx->d_call_levels);
    CK(cudaMemcpyAsync(ctx->h_call_levels, ctx->d_call_levels, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

// run-length buffers, allocated (or grown) on the first call that asks for them
int ensure_runs(dofs3d_ctx* ctx, int max_runs) {
    if (max_runs < 1) {
        ctx->err = "max_runs must be positive for the run-length label format";
        return DOFS3D_ERR_ARG;
    }
    if (!ctx->run_count) {
        ctx->run_blocks = (ctx->N + SEG_THREADS - 1) / SEG_THREADS;
        DA(ctx->run_count, (size_t)ctx->F * ctx->run_blocks);
        DA(ctx->n_runs, (size_t)ctx->F);
    }
    if (max_runs > ctx->runs_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->runs) {
            cudaFree(ctx->runs);
            ctx->allocs.erase(std::find(ctx->allocs.begin(), ctx->allocs.end(), (void*)ctx->runs));
            ctx->bytes -= (long long)sizeof(dofs3d_run) * ctx->F * ctx->runs_cap;
            ctx->runs = nullptr;
        }
        DA(ctx->runs, (size_t)ctx->F * max_runs);
        ctx->runs_cap = max_runs;
    }
    return 0;
}

// the deferred conditions of every call since the last check, from the pinned copy of the sticky bits (the stream must
// have been synchronised); clears them
int check_sticky(dofs3d_ctx* ctx) {
    const int bits = *ctx->h_sticky;
    if (!bits) return 0;
    *ctx->h_sticky = 0;
    CK(cudaMemsetAsync(ctx->sticky, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bits & STICKY_INTERNAL) {
        ctx->err = "internal: Boruvka did not converge (non-finite flow?) or a sort look-back timed out";
        return DOFS3D_ERR_INTERNAL;
    }
    ctx->err = bits & STICKY_CANDIDATES ? "candidate queue overflow"
               : bits & STICKY_BOXES    ? "more boxes than max_boxes"
                                        : "more label runs than max_runs";
    return DOFS3D_ERR_OVERFLOW;
}

// copy results of the last segment_dev to caller memory (kind = cudaMemcpyDeviceToHost / DeviceToDevice); asynchronous
int export_results(dofs3d_ctx* ctx, int n, const dofs3d_outputs& o, cudaMemcpyKind kind) {
    const int F = ctx->F;
    if (o.labels) {
        if (o.label_format == DOFS3D_LABELS_RLE) {
            const int cols = std::min(o.max_runs, ctx->runs_cap);
            CK(cudaMemcpy2DAsync(o.labels, (size_t)o.max_runs * sizeof(dofs3d_run), ctx->runs,
                                 (size_t)ctx->runs_cap * sizeof(dofs3d_run), (size_t)cols * sizeof(dofs3d_run), n, kind,
                                 ctx->stream));
        } else {
            const size_t el = o.label_format == DOFS3D_LABELS_U16 ? sizeof(u16) : sizeof(int32_t);
            CK(cudaMemcpyAsync(o.labels, ctx->labels, (size_t)n * ctx->N * el, kind, ctx->stream));
        }
    }
    if (o.n_runs && o.label_format == DOFS3D_LABELS_RLE)
        CK(cudaMemcpyAsync(o.n_runs, ctx->n_runs, sizeof(int) * n, kind, ctx->stream));
    if (o.boxes && o.max_boxes > 0) {
        const int cols = std::min(o.max_boxes, ctx->box_cap);
        CK(cudaMemcpy2DAsync(o.boxes, (size_t)o.max_boxes * sizeof(dofs3d_box), ctx->boxes,
                             (size_t)ctx->box_cap * sizeof(dofs3d_box), (size_t)cols * sizeof(dofs3d_box), n, kind,
                             ctx->stream));
    }
    if (o.n_boxes) CK(cudaMemcpyAsync(o.n_boxes, ctx->counters + CNT_BOXES * F, sizeof(int) * n, kind, ctx->stream));
    if (o.stats) CK(cudaMemcpyAsync(o.stats, ctx->stats, sizeof(dofs3d_stats) * n, kind, ctx->stream));
    return 0;
}

int check_outputs(dofs3d_ctx* ctx, const dofs3d_outputs* o, OutSpec* spec) {
    if (!o) {
        ctx->err = "null dofs3d_outputs";
        return DOFS3D_ERR_ARG;
    }
    if (o->label_format < DOFS3D_LABELS_I32 || o->label_format > DOFS3D_LABELS_RLE || o->max_boxes < 0 ||
        (o->label_format == DOFS3D_LABELS_RLE && o->labels && o->max_runs < 1)) {
        ctx->err = "bad dofs3d_outputs (label_format / max_boxes / max_runs)";
        return DOFS3D_ERR_ARG;
    }
    // run-length output the caller does not collect is not computed
    spec->fmt = (o->label_format == DOFS3D_LABELS_RLE && !o->labels) ? DOFS3D_LABELS_U16 : o->label_format;
    spec->max_runs = o->max_runs;
    spec->max_boxes = o->boxes ? o->max_boxes : -1;
    return 0;
}

dofs3d_outputs legacy_outputs(int32_t* labels, dofs3d_box* boxes, int32_t* n_boxes, int max_boxes, dofs3d_stats* stats) {
    dofs3d_outputs o;
    memset(&o, 0, sizeof o);
    o.label_format = DOFS3D_LABELS_I32;
    o.labels = labels;
    o.boxes = boxes;
    o.n_boxes = n_boxes;
    o.max_boxes = max_boxes;
    o.stats = stats;
    return o;
}

// cvtColor + Farneback for n pairs out of n+1 consecutive gray frames already in ctx->gray
int flow_dev(dofs3d_ctx* ctx, const u8* d_gray0, const u8* d_gray1, int n, float2* d_flow_out, bool carry_in = false,
             bool carry_out = false) {
    FlowLaunchStats st;
    st.mark = [](void* u, const char* name) { mark(static_cast<dofs3d_ctx*>(u), name); };
    st.user = ctx;
    int rc = farneback_run(ctx->fb, d_gray0, d_gray1, n, d_flow_out, ctx->stream, &st, carry_in, carry_out);
    ctx->launches += st.launches;
    if (rc) {
        ctx->err = "farneback_run failed";
        return DOFS3D_ERR_CUDA;
    }
    return 0;
}

int check_batch(dofs3d_ctx* ctx, int n) {
    if (!ctx) return DOFS3D_ERR_ARG;
    if (n < 0 || n > ctx->F) {
        ctx->err = "n_pairs out of range (0..max_pairs)";
        return DOFS3D_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

// the reference's calibration quads are pixel coordinates of a 640x360 frame; (sx, sy) rescales them to another frame size
static void calibration(dofs3d_params* p, double sx, double sy) {
    // get_mat (lifting_3d.cpp:482-514): image quad <-> bird's-eye-view rectangle
    const double img[4][2] = {{215 * sx, 265 * sy}, {90 * sx, 121 * sy}, {294 * sx, 120 * sy}, {625 * sx, 265 * sy}};
    const double bev[4][2] = {{100, 13000}, {100, 6000}, {800, 6000}, {800, 13000}};
    perspective_transform(img, bev, p->persp);
    perspective_transform(bev, img, p->inv);
    // get_mat_upper (lifting_3d.cpp:441-480): the same BEV rectangle to the roof-height image quads
    const double roof_y[3][2] = {{176, 85}, {185, 80}, {140, 55}};
    for (int c = 0; c < 3; ++c) {
        const double roof[4][2] = {{215 * sx, roof_y[c][0] * sy}, {90 * sx, roof_y[c][1] * sy}, {294 * sx, roof_y[c][1] * sy},
                                   {625 * sx, roof_y[c][0] * sy}};
        perspective_transform(bev, roof, p->inv_upper[c]);
    }
}

void dofs3d_params_for_size(dofs3d_params* p, int width, int height) {
    if (!p) return;
    dofs3d_default_params(p);
    if (width > 0 && height > 0) calibration(p, width / 640.0, height / 360.0);
}

void dofs3d_default_params(dofs3d_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    calibration(p, 1.0, 1.0);
    p->pyr_scale = 0.5;
    p->levels = 3;
    p->winsize = 15;
    p->iters = 3;
    p->poly_n = 5;
    p->poly_sigma = 1.2;
    p->blur_sigma = 3.0;
    p->neighbors = 8;
    p->min_size = 500;
    p->score_threshold = 0.3;
    const int sizes[3][2] = {{258, 84}, {349, 165}, {370, 180}};
    memcpy(p->cls_size, sizes, sizeof sizes);
    p->cls_min_convexity[0] = 3.0 / 4.0;
    p->cls_min_convexity[1] = 1.0 / 2.0;
    p->cls_min_convexity[2] = 20.0 / 29.0;
}

int dofs3d_create(dofs3d_ctx** out, int device, int width, int height, int max_pairs, const dofs3d_params* params) {
    if (!out) return DOFS3D_ERR_ARG;
    *out = nullptr;
    if (width < 2 || height < 2 || width > 65535 || height > 65535 || max_pairs < 1) return DOFS3D_ERR_ARG;
    if ((unsigned long long)width * height > (1ull << 26)) return DOFS3D_ERR_ARG;  // 4N slots stay below TIME_KEY_MIN_PREFIX
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return DOFS3D_ERR_CUDA;
    dofs3d_ctx* ctx = new dofs3d_ctx();
    *out = ctx;  // returned even on failure so that dofs3d_last_error can be read; destroy it either way
    ctx->device = device;
    ctx->W = width;
    ctx->H = height;
    ctx->N = width * height;
    ctx->S = 4 * (size_t)ctx->N;
    ctx->F = max_pairs;
    if (params) ctx->prm = *params;
    else dofs3d_default_params(&ctx->prm);
    const dofs3d_params& p = ctx->prm;
    if (p.neighbors != 4 && p.neighbors != 8) {
        ctx->err = "neighbors must be 4 or 8";
        return DOFS3D_ERR_ARG;
    }
    memcpy(ctx->seg.hg.persp, p.persp, sizeof p.persp);
    memcpy(ctx->seg.hg.inv, p.inv, sizeof p.inv);
    memcpy(ctx->seg.hg.upper, p.inv_upper, sizeof p.inv_upper);
    memcpy(ctx->seg.cls_size, p.cls_size, sizeof p.cls_size);
    memcpy(ctx->seg.cls_min_convexity, p.cls_min_convexity, sizeof p.cls_min_convexity);
    ctx->seg.score_threshold = p.score_threshold;
    ctx->seg.min_size = p.min_size;
    ctx->seg.neighbors = p.neighbors;
    // GaussianBlur(flow, Size(0,0), sigma) on CV_32F: ksize = cvRound(sigma*4*2 + 1) | 1
    {
        int ksize = ((int)lrint(p.blur_sigma * 8 + 1)) | 1;
        if (ksize < 1 || ksize > 2 * BLUR_MAX_RADIUS + 1) {
            ctx->err = "blur_sigma out of range";
            return DOFS3D_ERR_ARG;
        }
        ctx->taps.radius = ksize / 2;
        gaussian_taps(ksize, p.blur_sigma, ctx->taps.k);
    }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaFuncSetAttribute(k_radix_scatter<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_smem_bytes<u64>()));
    CK(cudaFuncSetAttribute(k_radix_scatter<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_smem_bytes<u32>()));

    const size_t F = ctx->F, N = ctx->N, S = ctx->S;
    DA(ctx->flow_in, F * N);
    DA(ctx->flow_tmp, F * N);
    DA(ctx->flow_blur, F * N);
    // sort buffers: the segmentation needs per frame 3N u64 in keysA (two u32 time-key buffers + two u64 fallback buffers),
    // 2N u64 in keysB (the 4N u32 prefixes, later the event keys) and N u32 in each payload buffer; the parity hook
    // dofs3d_edges_sorted sorts all 4N slots of ONE frame (4N u64 / 4N u32 per buffer)
    DA(ctx->keysA, std::max(S, 3 * F * N));
    DA(ctx->keysB, std::max(S, 2 * F * N));
    DA(ctx->valsA, std::max(S, F * N));
    DA(ctx->valsB, std::max(S, F * N));
    ctx->num_tiles = (int)((S + RS_TILE - 1) / RS_TILE);
    DA(ctx->tile_hist, RS_BINS * std::max<size_t>(ctx->num_tiles, F * ((N + RS_TILE - 1) / RS_TILE)));  // F frames of N, or 1 of 4N
    DA(ctx->digit_tot, F * RS_BINS);
    DA(ctx->sweep_hist, 2 * F * RS_MAX_PASSES * RS_BINS);
    DA(ctx->sweep_ticket, RS_MAX_PASSES + 1);
    CK(cudaFuncSetAttribute(k_radix_onesweep<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_smem_bytes<u64>()));
    CK(cudaFuncSetAttribute(k_radix_onesweep<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_smem_bytes<u32>()));
    DA(ctx->bor.comp, F * N);
    DA(ctx->bor.best, F * N);
    DA(ctx->bor.newp, F * N);
    DA(ctx->bor.loss_time, F * N);
    DA(ctx->bor.up, F * N);
    DA(ctx->bor.lvl, F * N);
    DA(ctx->win, F * N);
    DA(ctx->wave_start, F * (EV_MAX_WAVES + 1));
    ctx->list_cap = (int)std::min<size_t>(F * (N / REPLAY_SHORT + 1), (size_t)1 << 28);
    DA(ctx->long_list, (size_t)ctx->list_cap);
    DA(ctx->long_count, EV_MAX_WAVES + 1);
    DA(ctx->repair_flags, 2);
    DA(ctx->ev_op, F * N);
    DA(ctx->ev_inv, F * N);
    DA(ctx->ev_size, F * N);
    DA(ctx->ev_sa, F * N);
    DA(ctx->ev_bbox, F * N);
    DA(ctx->ev_prod, F * N);
    DA(ctx->ev_flow, F * N);
    ctx->tiles_cap = (int)((N + REPLAY_TILE - 1) / REPLAY_TILE + 1);
    DA(ctx->tile_agg, F * ctx->tiles_cap);
    DA(ctx->tile_carry, F * ctx->tiles_cap);
    DA(ctx->rstate, F * N);
    DA(ctx->best_score, F * N);
    DA(ctx->sel_time, F * N);
    DA(ctx->sel_box, F * N);
    ctx->cand_cap = (int)std::max<size_t>(65536, N / 4);
    ctx->box_cap = 4096;
    DA(ctx->cand, F * ctx->cand_cap);
    DA(ctx->cand_score, F * ctx->cand_cap);
    DA(ctx->counters, (size_t)CNT_KINDS * F);
    DA(ctx->bor.n_roots, (size_t)EV_MAX_WAVES * F);
    DA(ctx->bor.mask, F * N);
    DA(ctx->bor.roots[0], F * N);
    DA(ctx->bor.roots[1], F * N);
    DA(ctx->bor.levels, F);
    ctx->bor.final_root = ctx->counters + CNT_FINAL * F;
    ctx->bor.F = (int)F;
    ctx->max_levels = std::min(EV_MAX_WAVES - 2, ceil_log2((unsigned long long)N) + 1);  // components at least halve per level
    if (const char* e = getenv("DOFS3D_MAX_LEVELS")) {  // experiment knob: fewer levels than the guaranteed bound (a frame that
        const int v = atoi(e);                           // needs more is reported as failed, never silently wrong)
        if (v >= 1 && v < ctx->max_levels) ctx->max_levels = v;
    }
    if (const char* e = getenv("DOFS3D_FORCE_TIME_FALLBACK")) ctx->force_time_fallback = atoi(e) != 0;
    if (const char* e = getenv("DOFS3D_BLUR_TMA")) ctx->blur_tma = atoi(e) != 0;
    if (const char* e = getenv("DOFS3D_BOR_FOLD")) ctx->bor_fold = atoi(e) != 0;
    if (const char* e = getenv("DOFS3D_STRIDE_BLOCKS")) ctx->stride_blocks_per_sm = std::max(1, std::min(64, atoi(e)));
    if (const char* e = getenv("DOFS3D_CARVEOUT")) ctx->carveout = std::max(-1, std::min(100, atoi(e)));
    CK(cudaMallocHost(&ctx->h_sticky, sizeof(int)));
    *ctx->h_sticky = 0;
    DA(ctx->sticky, 1);
    DA(ctx->d_call_levels, 1);
    CK(cudaMallocHost(&ctx->h_call_levels, sizeof(int)));
    *ctx->h_call_levels = 0;
    if (const char* e = getenv("DOFS3D_COND_GRAPHS")) ctx->use_cond_graphs = atoi(e) != 0;
    if (const char* e = getenv("DOFS3D_SOFT_LEVELS")) ctx->soft_levels_forced = atoi(e);
    {
        std::vector<double> rcp(RCP_TABLE, 0.0);
        for (int i = 1; i < RCP_TABLE; ++i) rcp[i] = 1.0 / (double)i;
        DA(ctx->rcp_table, RCP_TABLE);
        CK(cudaMemcpy(ctx->rcp_table, rcp.data(), sizeof(double) * RCP_TABLE, cudaMemcpyHostToDevice));
    }
    CK(cudaMemsetAsync(ctx->sticky, 0, sizeof(int), ctx->stream));
    DA(ctx->boxes_tmp, F * ctx->box_cap);
    DA(ctx->boxes, F * ctx->box_cap);
    DA(ctx->labels, F * N);
    DA(ctx->stats, F);
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// buffers of the gray / flow stages, allocated on first use (a segmentation-only context never pays for them)
static int ensure_flow(dofs3d_ctx* ctx) {
    if (ctx->flow_raw) return 0;
    const dofs3d_params& p = ctx->prm;
    const size_t F = ctx->F, N = ctx->N;
    const int width = ctx->W, height = ctx->H;
    DA(ctx->bgr, (F + 1) * N * 3);
    DA(ctx->gray, (F + 1) * N);
    {
        FlowConfig fc;
        fc.pyr_scale = p.pyr_scale;
        fc.levels = p.levels;
        fc.winsize = p.winsize;
        fc.iters = p.iters;
        fc.poly_n = p.poly_n;
        fc.poly_sigma = p.poly_sigma;
        size_t fbytes = 0;
        if (const char* e = getenv("DOFS3D_FLOW_FUSE")) ctx->fb.fuse_um = atoi(e) != 0;  // (its second M buffer is only allocated then)
        if (const char* e = getenv("DOFS3D_BS7_FLOAT")) ctx->fb.bs7_float = atoi(e) != 0;
        int rc = farneback_alloc(&ctx->fb, width, height, (int)F, fc, &fbytes);
        ctx->bytes += (long long)fbytes;
        if (const char* e = getenv("DOFS3D_PYR_TILED")) ctx->fb.pyr_untiled = atoi(e) == 0;
        if (const char* e = getenv("DOFS3D_PYR_GENERIC")) ctx->fb.pyr_generic = atoi(e) != 0;
        if (ctx->carveout >= 0) farneback_set_carveout(ctx->carveout);
        if (rc) {
            ctx->err = farneback_error(rc);
            return rc == 1 ? DOFS3D_ERR_ARG : rc == 3 ? DOFS3D_ERR_CUDA : DOFS3D_ERR_NOMEM;
        }
    }
    DA(ctx->flow_raw, F * N);
    return 0;
}

void dofs3d_destroy(dofs3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->allocs) cudaFree(p);
    farneback_free(&ctx->fb);
    if (ctx->h_sticky) cudaFreeHost(ctx->h_sticky);
    if (ctx->h_call_levels) cudaFreeHost(ctx->h_call_levels);
    for (auto& g : ctx->cond_graphs) {
        cudaGraphExecDestroy(g.exec);
        cudaGraphDestroy(g.graph);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage[i]) cudaFree(ctx->stage[i]);
        if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]);
        if (ctx->stage_ready[i]) cudaEventDestroy(ctx->stage_ready[i]);
    }
    for (auto& c : ctx->chunks) cudaEventDestroy(c.done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& m : ctx->timer.marks) cudaEventDestroy(m.second);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int dofs3d_sync(dofs3d_ctx* ctx) {
    if (!ctx) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_sticky(ctx);  // deferred overflow / non-convergence of every asynchronous call since the last sync
}

const char* dofs3d_last_error(const dofs3d_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* dofs3d_stream(dofs3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long dofs3d_launch_count(const dofs3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
long long dofs3d_device_bytes(const dofs3d_ctx* ctx) { return ctx ? ctx->bytes : 0; }

int dofs3d_set_timing(dofs3d_ctx* ctx, int enabled) {
    if (!ctx) return DOFS3D_ERR_ARG;
    ctx->timer.enabled = enabled != 0;
    return 0;
}

int dofs3d_get_timing(dofs3d_ctx* ctx, const char** names, float* ms, int* counts, int cap) {
    if (!ctx) return DOFS3D_ERR_ARG;
    timer_collect(ctx);  // waits for the stream; the timed call itself stays asynchronous
    int n = 0;
    for (auto& r : ctx->timer.result) {
        if (n >= cap) break;
        if (names) names[n] = r.name.c_str();
        if (ms) ms[n] = r.ms;
        if (counts) counts[n] = r.count;
        ++n;
    }
    return n;
}

// ------------------------------------------------------------------------------------- gray
int dofs3d_gray_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr, int n_frames, uint8_t* d_gray_out) {
    if (!ctx || !d_bgr || !d_gray_out || n_frames < 0) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_frames == 0) return 0;
    LAUNCH(ctx, k_bgr2gray, grid1((size_t)ctx->N * n_frames / 4 + 1, 256, 1), 256, 0, d_bgr, d_gray_out,
           (size_t)ctx->N * n_frames);
    CK(cudaGetLastError());
    return 0;
}

int dofs3d_gray(dofs3d_ctx* ctx, const uint8_t* bgr, int n_frames, uint8_t* gray_out) {
    if (!ctx || !bgr || !gray_out) return DOFS3D_ERR_ARG;
    if (n_frames < 0 || n_frames > ctx->F + 1) {
        ctx->err = "n_frames out of range (0..max_pairs+1)";
        return DOFS3D_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    if (int rc0 = ensure_flow(ctx)) return rc0;
    const size_t px = (size_t)ctx->N * n_frames;
    CK(cudaMemcpyAsync(ctx->bgr, bgr, px * 3, cudaMemcpyHostToDevice, ctx->stream));
    int rc = dofs3d_gray_dev(ctx, ctx->bgr, n_frames, ctx->gray);
    if (rc) return rc;
    CK(cudaMemcpyAsync(gray_out, ctx->gray, px, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- flow
int dofs3d_flow_dev(dofs3d_ctx* ctx, const uint8_t* d_gray0, const uint8_t* d_gray1, int n_pairs, float* d_flow_out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!d_gray0 || !d_gray1 || !d_flow_out) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    if ((rc = ensure_flow(ctx))) return rc;
    timer_begin(ctx);
    rc = flow_dev(ctx, d_gray0, d_gray1, n_pairs, reinterpret_cast<float2*>(d_flow_out));
    if (rc) return rc;
    CK(cudaGetLastError());
    return 0;
}

int dofs3d_flow(dofs3d_ctx* ctx, const uint8_t* gray0, const uint8_t* gray1, int n_pairs, float* flow_out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!gray0 || !gray1 || !flow_out) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    if ((rc = ensure_flow(ctx))) return rc;
    const size_t px = (size_t)ctx->N * n_pairs;
    // staging: gray0 frames in ctx->gray, gray1 frames in ctx->bgr (3x larger than needed)
    CK(cudaMemcpyAsync(ctx->gray, gray0, px, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->bgr, gray1, px, cudaMemcpyHostToDevice, ctx->stream));
    rc = dofs3d_flow_dev(ctx, ctx->gray, ctx->bgr, n_pairs, reinterpret_cast<float*>(ctx->flow_raw));
    if (rc) return rc;
    CK(cudaMemcpyAsync(flow_out, ctx->flow_raw, px * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- blur
int dofs3d_blur(dofs3d_ctx* ctx, const float* flow_in, int n_pairs, float* flow_out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!flow_in || !flow_out) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    const size_t px = (size_t)ctx->N * n_pairs;
    const dim3 gN = grid1(ctx->N, SEG_THREADS, n_pairs);
    CK(cudaMemcpyAsync(ctx->flow_in, flow_in, px * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    blur_launch(ctx, ctx->flow_in, ctx->flow_blur, n_pairs);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flow_out, ctx->flow_blur, px * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- segment
int dofs3d_segment_dev(dofs3d_ctx* ctx, const float* d_flow, int already_blurred, int n_pairs, int32_t* d_labels_out,
                       dofs3d_box* d_boxes_out, int32_t* d_n_boxes_out, int max_boxes, dofs3d_stats* d_stats_out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!d_flow || max_boxes < 0) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    const dofs3d_outputs o = legacy_outputs(d_labels_out, d_boxes_out, d_n_boxes_out, max_boxes, d_stats_out);
    OutSpec spec;
    if ((rc = check_outputs(ctx, &o, &spec))) return rc;
    timer_begin(ctx);
    rc = segment_dev(ctx, d_flow, already_blurred, n_pairs, spec);
    if (rc) return rc;
    rc = export_results(ctx, n_pairs, o, cudaMemcpyDeviceToDevice);
    CK(cudaGetLastError());
    return rc;
}

int dofs3d_segment_ex(dofs3d_ctx* ctx, const float* flow, int already_blurred, int n_pairs, const dofs3d_outputs* out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    OutSpec spec;
    if (!flow) return DOFS3D_ERR_ARG;
    if ((rc = check_outputs(ctx, out, &spec))) return rc;
    if (n_pairs == 0) return 0;
    const size_t px = (size_t)ctx->N * n_pairs;
    CK(cudaMemcpyAsync(ctx->flow_in, flow, px * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    timer_begin(ctx);
    rc = segment_dev(ctx, reinterpret_cast<const float*>(ctx->flow_in), already_blurred, n_pairs, spec);
    if (rc) return rc;
    if ((rc = export_results(ctx, n_pairs, *out, cudaMemcpyDeviceToHost))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return check_sticky(ctx);
}

int dofs3d_segment(dofs3d_ctx* ctx, const float* flow, int already_blurred, int n_pairs, int32_t* labels_out,
                   dofs3d_box* boxes_out, int32_t* n_boxes_out, int max_boxes, dofs3d_stats* stats_out,
                   float* flow_blurred_out) {
    if (!ctx || max_boxes < 0) return DOFS3D_ERR_ARG;
    const dofs3d_outputs o = legacy_outputs(labels_out, boxes_out, n_boxes_out, max_boxes, stats_out);
    int rc = dofs3d_segment_ex(ctx, flow, already_blurred, n_pairs, &o);
    if (flow_blurred_out && n_pairs > 0 && (rc == 0 || rc == DOFS3D_ERR_OVERFLOW)) {
        CK(cudaMemcpyAsync(flow_blurred_out, ctx->flow_blur, (size_t)ctx->N * n_pairs * sizeof(float2), cudaMemcpyDeviceToHost,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

// ------------------------------------------------------------------------------------- paint
int dofs3d_paint(dofs3d_ctx* ctx, int n_pairs, double min_score, int32_t* painted_out, uint8_t* bgr_inout) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!painted_out) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    const size_t px = (size_t)ctx->N * n_pairs;
    int32_t* d_painted = reinterpret_cast<int32_t*>(ctx->sel_box);  // dead after the labels were written
    u8* d_bgr = nullptr;
    if (bgr_inout) {
        if ((rc = ensure_flow(ctx))) return rc;
        d_bgr = ctx->bgr;
        CK(cudaMemcpyAsync(d_bgr, bgr_inout, px * 3, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (ctx->labels_fmt == DOFS3D_LABELS_I32)
        LAUNCH(ctx, (k_paint<dofs3d_box, int>), grid1(ctx->N, SEG_THREADS, n_pairs), SEG_THREADS, 0, ctx->labels, ctx->boxes,
               ctx->box_cap, ctx->N, min_score, d_painted, d_bgr);
    else
        LAUNCH(ctx, (k_paint<dofs3d_box, u16>), grid1(ctx->N, SEG_THREADS, n_pairs), SEG_THREADS, 0,
               reinterpret_cast<const u16*>(ctx->labels), ctx->boxes, ctx->box_cap, ctx->N, min_score, d_painted, d_bgr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(painted_out, d_painted, px * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (bgr_inout) CK(cudaMemcpyAsync(bgr_inout, d_bgr, px * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- render
int dofs3d_render(dofs3d_ctx* ctx, int n_pairs, double min_score, uint8_t* bgr_inout) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!bgr_inout) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) return 0;
    if ((rc = ensure_flow(ctx))) return rc;  // ctx->bgr: the frames
    const size_t px = (size_t)ctx->N * n_pairs;
    if (!ctx->render_seg) DA(ctx->render_seg, (size_t)ctx->F * ctx->N * 3);
    int32_t* d_painted = reinterpret_cast<int32_t*>(ctx->sel_box);  // dead after the labels were written
    CK(cudaMemcpyAsync(ctx->bgr, bgr_inout, px * 3, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->render_seg, ctx->bgr, px * 3, cudaMemcpyDeviceToDevice, ctx->stream));  // seg = frame.clone()
    timer_begin(ctx);
    const dim3 gN = grid1(ctx->N, SEG_THREADS, n_pairs);
    if (ctx->labels_fmt == DOFS3D_LABELS_I32)
        LAUNCH(ctx, (k_paint<dofs3d_box, int>), gN, SEG_THREADS, 0, ctx->labels, ctx->boxes, ctx->box_cap, ctx->N, min_score,
               d_painted, ctx->render_seg);
    else
        LAUNCH(ctx, (k_paint<dofs3d_box, u16>), gN, SEG_THREADS, 0, reinterpret_cast<const u16*>(ctx->labels), ctx->boxes,
               ctx->box_cap, ctx->N, min_score, d_painted, ctx->render_seg);
    LAUNCH(ctx, k_cube_lines<dofs3d_box>, dim3((ctx->box_cap * 12 + 127) / 128, n_pairs), 128, 0, ctx->boxes,
           ctx->counters + CNT_BOXES * ctx->F, ctx->box_cap, d_painted, ctx->bgr, ctx->render_seg, ctx->W, ctx->H, min_score);
    const double opacity = 2.0 / 5.0;  // draw.cpp:157
    LAUNCH(ctx, k_add_weighted, dim3((unsigned)((px * 3 / 4 + 256) / 256)), 256, 0, ctx->bgr, ctx->render_seg, px * 3,
           (float)(1.0 - opacity), (float)opacity);
    mark(ctx, "render");
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(bgr_inout, ctx->bgr, px * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- lift
int dofs3d_lift(dofs3d_ctx* ctx, const float* dir2, const int32_t* bbox4, const int32_t* cls, int n, dofs3d_box* out) {
    if (!ctx || !dir2 || !bbox4 || !cls || !out || n < 0) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    for (int i = 0; i < n; ++i)
        if (cls[i] < 0 || cls[i] > 2) {
            ctx->err = "cls must be 0, 1 or 2";
            return DOFS3D_ERR_ARG;
        }
    float2* d_dir = nullptr;
    int4* d_box = nullptr;
    int* d_cls = nullptr;
    dofs3d_box* d_out = nullptr;
    int rc = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&d_dir, sizeof(float2) * n)) != cudaSuccess || (e = cudaMalloc(&d_box, sizeof(int4) * n)) != cudaSuccess ||
        (e = cudaMalloc(&d_cls, sizeof(int) * n)) != cudaSuccess || (e = cudaMalloc(&d_out, sizeof(dofs3d_box) * n)) != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        rc = DOFS3D_ERR_NOMEM;
    }
    if (!rc) {
        cudaMemcpyAsync(d_dir, dir2, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_box, bbox4, sizeof(int4) * n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_cls, cls, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream);
        LAUNCH(ctx, k_lift_problems<dofs3d_box>, grid1(n, 128, 1), 128, 0, d_dir, d_box, d_cls, n, ctx->seg, d_out);
        cudaMemcpyAsync(out, d_out, sizeof(dofs3d_box) * n, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            ctx->err = cudaGetErrorString(e);
            rc = DOFS3D_ERR_CUDA;
        }
    }
    cudaFree(d_dir);
    cudaFree(d_box);
    cudaFree(d_cls);
    cudaFree(d_out);
    return rc;
}

// ------------------------------------------------------------------------------------- sorted edges
long long dofs3d_edges_sorted(dofs3d_ctx* ctx, const float* flow_blurred, int32_t* start, int32_t* end,
                              uint64_t* weight_bits) {
    if (!ctx || !flow_blurred) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int N = ctx->N;
    const int E = n_edges_of(ctx->W, ctx->H, ctx->seg.neighbors);
    CK(cudaMemcpyAsync(ctx->flow_blur, flow_blurred, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    int rc = build_sorted_edges(ctx, 1);
    if (rc) return rc;
    // decode (start, end, weight) into the dead B buffers
    int* d_start = reinterpret_cast<int*>(ctx->keysB);
    int* d_end = d_start + ctx->S;
    u64* d_w = ctx->keysA;  // the weights by slot are no longer needed
    LAUNCH(ctx, k_edges_decode, grid1(E, SEG_THREADS, 1), SEG_THREADS, 0, ctx->valsA, ctx->flow_blur, d_start, d_end, d_w,
           ctx->W, E);
    CK(cudaGetLastError());
    if (start) CK(cudaMemcpyAsync(start, d_start, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost, ctx->stream));
    if (end) CK(cudaMemcpyAsync(end, d_end, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost, ctx->stream));
    if (weight_bits) CK(cudaMemcpyAsync(weight_bits, d_w, sizeof(u64) * (size_t)E, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return E;
}

// ------------------------------------------------------------------------------------- whole path
static int process_frames_dev(dofs3d_ctx* ctx, const u8* d_bgr, int n_frames, const OutSpec& spec) {
    const int n = n_frames - 1;
    const size_t N = ctx->N;
    LAUNCH(ctx, k_bgr2gray, grid1(N * n_frames / 4 + 1, 256, 1), 256, 0, d_bgr, ctx->gray, N * n_frames);
    mark(ctx, "gray");
    int rc = flow_dev(ctx, ctx->gray, ctx->gray + N, n, ctx->flow_raw);  // pair i = frames (i, i+1)
    if (rc) return rc;
    return segment_dev(ctx, reinterpret_cast<const float*>(ctx->flow_raw), 0, n, spec);
}

int dofs3d_process_ex_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr_frames, int n_frames, const dofs3d_outputs* d_out) {
    if (!ctx || !d_bgr_frames || n_frames < 1) return DOFS3D_ERR_ARG;
    const int n = n_frames - 1;
    int rc = check_batch(ctx, n);
    if (rc) return rc;
    OutSpec spec;
    if ((rc = check_outputs(ctx, d_out, &spec))) return rc;
    if (n == 0) return 0;
    if ((rc = ensure_flow(ctx))) return rc;
    ctx->have_carry = false;
    timer_begin(ctx);
    if ((rc = process_frames_dev(ctx, d_bgr_frames, n_frames, spec))) return rc;
    rc = export_results(ctx, n, *d_out, cudaMemcpyDeviceToDevice);
    CK(cudaGetLastError());
    return rc;
}

int dofs3d_process_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr_frames, int n_frames, int32_t* d_labels_out,
                       dofs3d_box* d_boxes_out, int32_t* d_n_boxes_out, int max_boxes, dofs3d_stats* d_stats_out) {
    if (max_boxes < 0) return DOFS3D_ERR_ARG;
    const dofs3d_outputs o = legacy_outputs(d_labels_out, d_boxes_out, d_n_boxes_out, max_boxes, d_stats_out);
    return dofs3d_process_ex_dev(ctx, d_bgr_frames, n_frames, &o);
}

int dofs3d_process_ex(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, const dofs3d_outputs* out) {
    if (!ctx || !bgr_frames || n_frames < 1) return DOFS3D_ERR_ARG;
    const int n = n_frames - 1;
    int rc = check_batch(ctx, n);
    if (rc) return rc;
    OutSpec spec;
    if ((rc = check_outputs(ctx, out, &spec))) return rc;
    if (n == 0) return 0;
    if ((rc = ensure_flow(ctx))) return rc;
    ctx->have_carry = false;
    const size_t N = ctx->N;
    CK(cudaMemcpyAsync(ctx->bgr, bgr_frames, N * 3 * n_frames, cudaMemcpyHostToDevice, ctx->stream));
    timer_begin(ctx);
    if ((rc = process_frames_dev(ctx, ctx->bgr, n_frames, spec))) return rc;
    if ((rc = export_results(ctx, n, *out, cudaMemcpyDeviceToHost))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return check_sticky(ctx);
}

int dofs3d_process(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, int32_t* labels_out, dofs3d_box* boxes_out,
                   int32_t* n_boxes_out, int max_boxes, dofs3d_stats* stats_out) {
    if (max_boxes < 0) return DOFS3D_ERR_ARG;
    const dofs3d_outputs o = legacy_outputs(labels_out, boxes_out, n_boxes_out, max_boxes, stats_out);
    return dofs3d_process_ex(ctx, bgr_frames, n_frames, &o);
}

// ------------------------------------------------------------------------------------- streaming
int dofs3d_stream_begin(dofs3d_ctx* ctx) {
    if (!ctx) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->chunks.empty()) {
        ctx->err = "dofs3d_stream_begin with chunks outstanding: collect them first";
        return DOFS3D_ERR_ARG;
    }
    ctx->have_carry = false;
    return 0;
}

int dofs3d_stream_submit(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, const dofs3d_outputs* out) {
    if (!ctx || !bgr_frames || n_frames < 1) return DOFS3D_ERR_ARG;
    const int n = ctx->have_carry ? n_frames : n_frames - 1;  // pairs of this chunk
    int rc = check_batch(ctx, n);
    if (rc) return rc;
    OutSpec spec;
    if ((rc = check_outputs(ctx, out, &spec))) return rc;
    if (ctx->chunks.size() >= 2) {
        ctx->err = "two chunks are already outstanding: collect one first";
        return DOFS3D_ERR_ARG;
    }
    if ((rc = ensure_flow(ctx))) return rc;
    const size_t N = ctx->N;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaMalloc(&ctx->stage[i], (size_t)(ctx->F + 1) * N * 3));
            ctx->bytes += (long long)(ctx->F + 1) * N * 3;
            CK(cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->stage_ready[i], cudaEventDisableTiming));
        }
    }
    const int b = (int)(ctx->chunks_submitted & 1);
    // upload on the copy stream, under the kernels of the previous chunk; the buffer is free once the gray conversion
    // of the chunk that used it last has run
    if (ctx->chunks_submitted >= 2) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[b], 0));
    CK(cudaMemcpyAsync(ctx->stage[b], bgr_frames, N * 3 * n_frames, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->stage_ready[b], ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->stage_ready[b], 0));
    timer_begin(ctx);
    // gray slot 0 holds the carried frame; the chunk's frames follow it
    u8* gray_dst = ctx->gray + (ctx->have_carry ? N : 0);
    LAUNCH(ctx, k_bgr2gray, grid1(N * n_frames / 4 + 1, 256, 1), 256, 0, ctx->stage[b], gray_dst, N * n_frames);
    CK(cudaEventRecord(ctx->stage_free[b], ctx->stream));
    mark(ctx, "gray");
    ctx->chunks_submitted++;
    cudaEvent_t done;
    CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    if (n > 0) {
        if ((rc = flow_dev(ctx, ctx->gray, ctx->gray + N, n, ctx->flow_raw, ctx->have_carry, true))) return rc;
        if ((rc = segment_dev(ctx, reinterpret_cast<const float*>(ctx->flow_raw), 0, n, spec))) return rc;
        if ((rc = export_results(ctx, n, *out, cudaMemcpyDeviceToHost))) return rc;
        // the last frame becomes frame 0 of the next chunk
        CK(cudaMemcpyAsync(ctx->gray, ctx->gray + (size_t)n * N, N, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->have_carry = true;
    } else {
        // a first chunk of one frame: nothing to pair yet; it is expanded with the next chunk (gray slot 0 already holds it,
        // but its polynomial expansion does not exist yet, so the next chunk treats it as a fresh frame)
        ctx->err = "the first chunk of a stream needs at least two frames";
        cudaEventDestroy(done);
        return DOFS3D_ERR_ARG;
    }
    CK(cudaEventRecord(done, ctx->stream));
    CK(cudaGetLastError());
    ctx->chunks.push_back({n, done});
    return 0;
}

int dofs3d_stream_collect(dofs3d_ctx* ctx, int* n_pairs_out) {
    if (!ctx) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->chunks.empty()) {
        ctx->err = "no chunk outstanding";
        return DOFS3D_ERR_ARG;
    }
    const dofs3d_ctx::Chunk c = ctx->chunks.front();
    ctx->chunks.erase(ctx->chunks.begin());
    CK(cudaEventSynchronize(c.done));
    cudaEventDestroy(c.done);
    if (n_pairs_out) *n_pairs_out = c.n_pairs;
    // the sticky bits copied after this chunk's kernels are on the host now (a later chunk may add to them; it reports then)
    const int bits = *ctx->h_sticky;
    if (bits) {
        CK(cudaStreamSynchronize(ctx->stream));
        return check_sticky(ctx);
    }
    return 0;
}

// ------------------------------------------------------------------------------------- Felzenszwalb mode
void dofs3d_fh_default_params(dofs3d_fh_params* p) {
    if (!p) return;
    p->k = 10.0;
    p->min_size = 100;
    p->neighbors = 8;
    p->flow_dist = 5.0;
    p->edge_dist = 5.0;
    p->stage = 3;
}

int dofs3d_segment_fh(dofs3d_ctx* ctx, const float* flow, const dofs3d_fh_params* params, int32_t* labels_out,
                      int32_t* n_components_out) {
    if (!ctx || !flow) return DOFS3D_ERR_ARG;
    dofs3d_fh_params p;
    if (params) p = *params;
    else dofs3d_fh_default_params(&p);
    if ((p.neighbors != 4 && p.neighbors != 8) || p.stage < 1 || p.stage > 3 || !(p.k >= 0) || p.min_size < 0) {
        ctx->err = "bad dofs3d_fh_params";
        return DOFS3D_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    const int N = ctx->N, W = ctx->W, H = ctx->H;
    const int S = (int)ctx->S;
    CK(cudaMemcpyAsync(ctx->flow_blur, flow, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    timer_begin(ctx);
    // float32 weights of all 4N slots, stable radix sort: (weight, insertion order) like Python's sorted()
    u32* keyA = reinterpret_cast<u32*>(ctx->keysB);
    u32* keyB = keyA + ctx->S;
    LAUNCH(ctx, k_fh_edge_keys, grid1(N, SEG_THREADS, 1), SEG_THREADS, 0, ctx->flow_blur, keyA, W, H, p.neighbors == 8 ? 1 : 0);
    mark(ctx, "fh.edge_keys");
    const int side = radix_sort_onesweep<u32>(ctx, keyA, ctx->valsA, keyB, ctx->valsB, ctx->S, S, 1, 32, true, "fh.sort.hist",
                                              "fh.sort.scatter");
    if (side != 0) {
        ctx->err = "internal: odd number of sort passes";
        return DOFS3D_ERR_INTERNAL;
    }
    FhState st;
    st.parent = reinterpret_cast<int*>(ctx->bor.comp);
    st.rank = ctx->bor.lvl;
    st.size = ctx->ev_size;
    st.color = ctx->ev_flow;
    st.thr = ctx->ev_inv;
    LAUNCH(ctx, k_fh_init, grid1(N, SEG_THREADS, 1), SEG_THREADS, 0, st, ctx->flow_blur, N, p.k);
    const int E = n_edges_of(W, H, p.neighbors);  // non-finite weights sort behind them and are never walked
    for (int pass = 1; pass <= p.stage; ++pass) {
        LAUNCH(ctx, k_fh_walk, dim3(1), 32, 0, st, keyA, ctx->valsA, E, W, pass, p.k, p.min_size, p.flow_dist, p.edge_dist);
        mark(ctx, pass == 1 ? "fh.walk.threshold" : pass == 2 ? "fh.walk.small" : "fh.walk.merge");
    }
    int* d_count = ctx->counters;  // [0] of the per-frame counters: free between calls
    CK(cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_fh_labels, grid1(N, SEG_THREADS, 1), SEG_THREADS, 0, st, ctx->labels, d_count, N);
    mark(ctx, "fh.labels");
    CK(cudaGetLastError());
    if (labels_out) CK(cudaMemcpyAsync(labels_out, ctx->labels, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_components_out) CK(cudaMemcpyAsync(n_components_out, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------- BEV warp
int dofs3d_warp_perspective(dofs3d_ctx* ctx, const uint8_t* img, int width, int height, int channels, const float* mat9,
                            int out_w, int out_h, uint8_t* out) {
    if (!ctx || !img || !mat9 || !out || width < 1 || height < 1 || out_w < 1 || out_h < 1) return DOFS3D_ERR_ARG;
    if (channels != 1 && channels != 3 && channels != 4) {
        ctx->err = "channels must be 1, 3 or 4";
        return DOFS3D_ERR_ARG;
    }
    if (width > 32767 || height > 32767) {
        ctx->err = "source image too large for warpPerspective's 16-bit coordinates";
        return DOFS3D_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    WarpMatrix M;
    warp_invert(mat9, &M);
    std::vector<short> tab;
    warp_cubic_table(&tab);
    u8 *d_src = nullptr, *d_dst = nullptr;
    short* d_tab = nullptr;
    const size_t src_bytes = (size_t)width * height * channels, dst_bytes = (size_t)out_w * out_h * channels;
    int rc = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&d_src, src_bytes)) != cudaSuccess || (e = cudaMalloc(&d_dst, dst_bytes)) != cudaSuccess ||
        (e = cudaMalloc(&d_tab, tab.size() * sizeof(short))) != cudaSuccess) {
        ctx->err = cudaGetErrorString(e);
        rc = DOFS3D_ERR_NOMEM;
    }
    if (!rc) {
        cudaMemcpyAsync(d_src, img, src_bytes, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(short), cudaMemcpyHostToDevice, ctx->stream);
        timer_begin(ctx);
        const dim3 grid((out_w + WARP_BW - 1) / WARP_BW, (out_h + WARP_BH - 1) / WARP_BH);
        if (channels == 1) LAUNCH(ctx, k_warp_perspective_cubic<1>, grid, 256, 0, d_src, width, height, d_dst, out_w, out_h, M, d_tab);
        else if (channels == 3) LAUNCH(ctx, k_warp_perspective_cubic<3>, grid, 256, 0, d_src, width, height, d_dst, out_w, out_h, M, d_tab);
        else LAUNCH(ctx, k_warp_perspective_cubic<4>, grid, 256, 0, d_src, width, height, d_dst, out_w, out_h, M, d_tab);
        mark(ctx, "bev.warp");
        cudaMemcpyAsync(out, d_dst, dst_bytes, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            ctx->err = cudaGetErrorString(e);
            rc = DOFS3D_ERR_CUDA;
        }
    }
    cudaFree(d_src);
    cudaFree(d_dst);
    cudaFree(d_tab);
    return rc;
}

int dofs3d_bev_transform(dofs3d_ctx* ctx, const uint8_t* bgr_frame, uint8_t* bev_out) {
    if (!ctx) return DOFS3D_ERR_ARG;
    return dofs3d_warp_perspective(ctx, bgr_frame, ctx->W, ctx->H, 3, ctx->prm.persp, DOFS3D_BEV_WIDTH, DOFS3D_BEV_HEIGHT, bev_out);
}

int dofs3d_pack_boxes_dev(dofs3d_ctx* ctx, int n_pairs, dofs3d_box* d_out, int capacity, int32_t* d_total_out) {
    int rc = check_batch(ctx, n_pairs);
    if (rc) return rc;
    if (!d_out || !d_total_out || capacity < 0) return DOFS3D_ERR_ARG;
    if (n_pairs == 0) {
        CK(cudaMemsetAsync(d_total_out, 0, sizeof(int32_t), ctx->stream));
        return 0;
    }
    LAUNCH(ctx, k_pack_boxes<dofs3d_box>, dim3(n_pairs), 128, 0, ctx->boxes, ctx->counters + CNT_BOXES * ctx->F, ctx->box_cap,
           n_pairs, d_out, capacity, d_total_out);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------- incremental Forest
}  // extern "C"

struct dofs3d_forest {
    dofs3d_ctx* ctx = nullptr;
    ForestState S;
    dofs3d_box* boxes = nullptr;
    int* pixels = nullptr;  // scratch of dofs3d_forest_pixels, N entries
    std::vector<void*> allocs;
};

namespace {
template <typename T>
bool falloc(dofs3d_forest* f, T** p, size_t count) {
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 16)) != cudaSuccess) return false;
    f->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return true;
}
int forest_result(dofs3d_forest* f, int* out) {
    dofs3d_ctx* ctx = f->ctx;
    int v = 0;
    CK(cudaMemcpyAsync(&v, f->S.counters + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (out) *out = v;
    return 0;
}
bool forest_node_ok(dofs3d_forest* f, int n) { return f && n >= 0 && n < f->S.N; }
}  // namespace

extern "C" {

int dofs3d_forest_create(dofs3d_ctx* ctx, const float* flow, dofs3d_forest** out) {
    if (!ctx || !flow || !out) return DOFS3D_ERR_ARG;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    dofs3d_forest* f = new dofs3d_forest();
    f->ctx = ctx;
    ForestState& S = f->S;
    S.W = ctx->W;
    S.H = ctx->H;
    S.N = ctx->N;
    S.box_cap = ctx->box_cap;
    const size_t N = ctx->N;
    float2* d_flow = nullptr;
    if (!falloc(f, &S.parent, N) || !falloc(f, &S.rank, N) || !falloc(f, &S.size, N) || !falloc(f, &S.flow, N) ||
        !falloc(f, &S.bbox, N) || !falloc(f, &S.next, N) || !falloc(f, &S.tail, N) || !falloc(f, &S.last_score, N) ||
        !falloc(f, &S.best_score, N) || !falloc(f, &S.snap_size, N) || !falloc(f, &S.box_slot, N) ||
        !falloc(f, &S.counters, (size_t)4) || !falloc(f, &f->boxes, (size_t)S.box_cap) || !falloc(f, &f->pixels, N) ||
        !falloc(f, &d_flow, N)) {
        ctx->err = "device allocation failed (forest)";
        dofs3d_forest_destroy(f);
        return DOFS3D_ERR_NOMEM;
    }
    CK(cudaMemcpyAsync(d_flow, flow, N * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_forest_init, grid1(N, 256, 1), 256, 0, S, d_flow);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    *out = f;
    return 0;
}

void dofs3d_forest_destroy(dofs3d_forest* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    for (void* p : f->allocs) cudaFree(p);
    delete f;
}

int dofs3d_forest_find(dofs3d_forest* f, int n, int32_t* root_out) {
    if (!forest_node_ok(f, n)) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    LAUNCH(ctx, k_forest_find, 1, 1, 0, f->S, n);
    return forest_result(f, root_out);
}

int dofs3d_forest_merge(dofs3d_forest* f, int a, int b, int32_t* root_out) {
    if (!forest_node_ok(f, a) || !forest_node_ok(f, b)) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    LAUNCH(ctx, k_forest_merge<dofs3d_box>, 1, 1, 0, f->S, a, b, 0, 0.0, 0, ctx->seg, f->boxes);
    return forest_result(f, root_out);
}

int dofs3d_forest_new_merge(dofs3d_forest* f, int a, int b, double score_threshold, int min_size) {
    if (!forest_node_ok(f, a) || !forest_node_ok(f, b)) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    LAUNCH(ctx, k_forest_merge<dofs3d_box>, 1, 1, 0, f->S, a, b, 1, score_threshold, min_size, ctx->seg, f->boxes);
    return forest_result(f, nullptr);
}

int dofs3d_forest_num_sets(dofs3d_forest* f, int32_t* num_sets_out) {
    if (!f || !num_sets_out) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(num_sets_out, f->S.counters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int dofs3d_forest_last_score(dofs3d_forest* f, int node, double* score_out) {
    if (!forest_node_ok(f, node) || !score_out) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(score_out, f->S.last_score + node, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int dofs3d_forest_bbox(dofs3d_forest* f, int node, int32_t* bbox4_out) {
    if (!forest_node_ok(f, node) || !bbox4_out) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    ushort4 b;
    CK(cudaMemcpyAsync(&b, f->S.bbox + node, sizeof b, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (b.x == 65535 && b.z == 65535) return 0;  // cleared when the node was absorbed
    bbox4_out[0] = b.x, bbox4_out[1] = b.y, bbox4_out[2] = b.z, bbox4_out[3] = b.w;
    return 1;
}

int dofs3d_forest_boxes(dofs3d_forest* f, int max_boxes, dofs3d_box* boxes_out) {
    if (!f || max_boxes < 0 || (max_boxes > 0 && !boxes_out)) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    int used = 0;
    CK(cudaMemcpyAsync(&used, f->S.counters + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (used > f->S.box_cap || used > max_boxes) {
        ctx->err = "more boxes than max_boxes";
        return DOFS3D_ERR_OVERFLOW;
    }
    std::vector<dofs3d_box> tmp((size_t)used);
    if (used) {
        CK(cudaMemcpyAsync(tmp.data(), f->boxes, sizeof(dofs3d_box) * used, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    std::sort(tmp.begin(), tmp.end(), [](const dofs3d_box& l, const dofs3d_box& r) { return l.root < r.root; });
    for (int i = 0; i < used; ++i) boxes_out[i] = tmp[i];
    return used;
}

int dofs3d_forest_pixels(dofs3d_forest* f, int root, int cap, int32_t* pixels_out) {
    if (!forest_node_ok(f, root) || cap < 0 || (cap > 0 && !pixels_out)) return DOFS3D_ERR_ARG;
    dofs3d_ctx* ctx = f->ctx;
    CK(cudaSetDevice(ctx->device));
    int n = 0;
    CK(cudaMemcpyAsync(&n, f->S.snap_size + root, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int take = std::min(n, cap);
    if (take > 0) {
        LAUNCH(ctx, k_forest_pixels, 1, 1, 0, f->S, root, take, f->pixels);
        CK(cudaMemcpyAsync(pixels_out, f->pixels, sizeof(int) * take, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
    }
    return n;
}

void* dofs3d_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}
void dofs3d_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------- per-node state
int dofs3d_node_state(dofs3d_ctx* ctx, int pair, int node, int32_t* size_out, float* mean_flow2_out, int32_t* bbox4_out) {
    if (!ctx || pair < 0 || pair >= ctx->F || node < 0 || node >= ctx->N) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t g = (size_t)pair * ctx->N + node;
    u8 lvl = 0;
    CK(cudaMemcpyAsync(&lvl, ctx->bor.lvl + g, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (lvl == 0) {  // never won a merge: the singleton it started as (Forest::Forest, graph.cpp:143-147)
        float2 f;
        CK(cudaMemcpyAsync(&f, ctx->flow_blur + g, sizeof f, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int x = node % ctx->W, y = node / ctx->W;
        if (size_out) *size_out = 1;
        if (mean_flow2_out) mean_flow2_out[0] = f.x, mean_flow2_out[1] = f.y;
        if (bbox4_out) bbox4_out[0] = x, bbox4_out[1] = y, bbox4_out[2] = x, bbox4_out[3] = y;
        return 0;
    }
    RootState r;
    CK(cudaMemcpyAsync(&r, ctx->rstate + g, sizeof r, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (size_out) *size_out = r.size;
    if (mean_flow2_out) mean_flow2_out[0] = r.fx, mean_flow2_out[1] = r.fy;
    if (bbox4_out) bbox4_out[0] = r.bbox.x, bbox4_out[1] = r.bbox.y, bbox4_out[2] = r.bbox.z, bbox4_out[3] = r.bbox.w;
    return 0;
}

int dofs3d_scored_merges(dofs3d_ctx* ctx, int pair, int cap, int32_t* root_out, uint32_t* time_out, double* score_out,
                         uint8_t* kept_out) {
    if (!ctx || pair < 0 || pair >= ctx->F || cap < 0) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int n_cand = 0;
    CK(cudaMemcpyAsync(&n_cand, ctx->counters + CNT_CAND * ctx->F + pair, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    n_cand = std::min(n_cand, ctx->cand_cap);
    std::vector<Candidate> cand((size_t)n_cand);
    std::vector<double> score((size_t)n_cand);
    if (n_cand) {
        CK(cudaMemcpyAsync(cand.data(), ctx->cand + (size_t)pair * ctx->cand_cap, sizeof(Candidate) * n_cand,
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(score.data(), ctx->cand_score + (size_t)pair * ctx->cand_cap, sizeof(double) * n_cand,
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    int m = 0;
    for (int i = 0; i < n_cand; ++i) {
        if (score[i] == -1.0) continue;  // no class produced a rectangle (graph.cpp:318-322)
        if (m < cap) {
            if (root_out) root_out[m] = (int32_t)cand[i].root;
            if (time_out) time_out[m] = cand[i].time;
            if (score_out) score_out[m] = score[i];
            if (kept_out) kept_out[m] = (cand[i].pad & CAND_KEPT) ? 1 : 0;
        }
        ++m;
    }
    return m;
}

// ------------------------------------------------------------------------------------- synthetic video
int dofs3d_synth_frames_dev(dofs3d_ctx* ctx, uint32_t seed, int n_objects, int first_frame, int n_frames,
                            uint8_t* d_bgr_out) {
    if (!ctx || !d_bgr_out || n_frames < 0 || n_objects < 0 || n_objects > SYNTH_MAX_OBJECTS) return DOFS3D_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_frames == 0) return 0;
    SynthScene sc;
    synth_make_scene(seed, n_objects, ctx->W, ctx->H, &sc);
    LAUNCH(ctx, k_synth_frames, grid1(ctx->N, 256, n_frames), 256, 0, sc, first_frame, d_bgr_out, ctx->W, ctx->H);
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
