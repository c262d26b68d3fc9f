// C++ multi-GPU streaming driver (the reference's video loop is main1,
// cpp/src/segment.cpp:174-275): ONE process, one host thread and one set of contexts per visible GPU, the frames of a
// synthetic video sharded contiguously over the GPUs (pair i = frames i, i+1; nothing is exchanged on the per-frame path),
// every shard pushed through the streaming entry points of the C ABI (dofs3d_stream_submit / dofs3d_stream_collect:
// pinned host frames in, run-length labels + boxes out, the upload of a chunk under the kernels of the previous one,
// the boundary frame of a chunk carried on the device), and at the end the per-frame results of every GPU's last chunk
// gathered over NCCL (ncclCommInitAll + grouped ncclAllGather straight from the device buffers dofs3d_pack_boxes_dev
// fills — NVLink / NVSwitch when there is more than one GPU).
//
//   multi_gpu_driver <width> <height> <pairs_per_gpu> [chunk_pairs=32] [contexts_per_gpu=3] [objects=8]
//
// Build: python -m denseopticalflowsegmentation3d_b200.build --examples   (g++, links libdofs3d, cudart, nccl)
#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "dofs3d.h"

#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                   \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)
#define NCCL_OK(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #call, ncclGetErrorString(r_));                   \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)
#define DOFS_OK(ctx, call)                                                                     \
    do {                                                                                       \
        int r_ = (call);                                                                       \
        if (r_ != 0) {                                                                         \
            std::fprintf(stderr, "%s: status %d (%s)\n", #call, r_, dofs3d_last_error(ctx));   \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

namespace {

constexpr int MAX_BOXES = 256;

struct Lane {  // one context of a GPU: its part of the shard, two chunks in flight
    dofs3d_ctx* ctx = nullptr;
    uint8_t* frames = nullptr;  // pinned: the whole part of the shard, (pairs + 1) frames
    int pairs = 0, submitted_pairs = 0, collected_pairs = 0, chunks_submitted = 0, chunks_collected = 0;
    struct Slot {
        dofs3d_run* runs = nullptr;
        int32_t *n_runs = nullptr, *n_boxes = nullptr;
        dofs3d_box* boxes = nullptr;
        dofs3d_stats* stats = nullptr;
    } slot[2];
    long long boxes_found = 0, runs_found = 0;
};

struct Gpu {
    int device = 0;
    std::vector<Lane> lanes;
    double seconds = 0;
    // gather payload: the boxes of every lane's last chunk, dense, and their count
    dofs3d_box* d_packed = nullptr;  // [lanes][cap]
    int32_t* d_counts = nullptr;     // [lanes]
    int32_t* d_all_counts = nullptr; // [gpus][lanes]
    dofs3d_box* d_all_boxes = nullptr;
};

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <width> <height> <pairs_per_gpu> [chunk_pairs=32] [contexts_per_gpu=3] [objects=8]\n", argv[0]);
        return 2;
    }
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]), pairs_per_gpu = std::atoi(argv[3]);
    const int chunk = argc > 4 ? std::atoi(argv[4]) : 32;
    const int n_lanes = argc > 5 ? std::atoi(argv[5]) : 3;
    const int objects = argc > 6 ? std::atoi(argv[6]) : 8;
    const size_t N = (size_t)W * H;
    const int max_runs = 32 * H;
    int n_gpus = 0;
    CUDA_OK(cudaGetDeviceCount(&n_gpus));
    if (n_gpus < 1 || pairs_per_gpu < n_lanes || chunk < 1) {
        std::fprintf(stderr, "need a GPU, chunk >= 1 and at least one pair per context\n");
        return 2;
    }
    std::vector<Gpu> gpus(n_gpus);
    const int cap = chunk * 64;  // boxes one lane contributes to the gather

    // ---- set-up (untimed): contexts, the shard's frames rendered on the device and parked in pinned host memory
    for (int g = 0; g < n_gpus; ++g) {
        Gpu& G = gpus[g];
        G.device = g;
        CUDA_OK(cudaSetDevice(g));
        G.lanes.resize(n_lanes);
        const int base = pairs_per_gpu / n_lanes, extra = pairs_per_gpu % n_lanes;
        int first_pair = g * pairs_per_gpu;  // contiguous shard of the stream
        for (int l = 0; l < n_lanes; ++l) {
            Lane& L = G.lanes[l];
            L.pairs = base + (l < extra ? 1 : 0);
            int rc = dofs3d_create(&L.ctx, g, W, H, chunk, nullptr);
            if (rc != 0) {
                std::fprintf(stderr, "dofs3d_create on GPU %d: %d (%s)\n", g, rc, L.ctx ? dofs3d_last_error(L.ctx) : "no device");
                return 1;
            }
            const int n_frames = L.pairs + 1;
            L.frames = static_cast<uint8_t*>(dofs3d_pinned_alloc((size_t)n_frames * N * 3));
            uint8_t* d_tmp = nullptr;
            CUDA_OK(cudaMalloc(&d_tmp, (size_t)(chunk + 1) * N * 3));
            for (int f0 = 0; f0 < n_frames; f0 += chunk + 1) {
                const int nf = std::min(chunk + 1, n_frames - f0);
                DOFS_OK(L.ctx, dofs3d_synth_frames_dev(L.ctx, 1234, objects, first_pair + f0, nf, d_tmp));
                DOFS_OK(L.ctx, dofs3d_sync(L.ctx));
                CUDA_OK(cudaMemcpy(L.frames + (size_t)f0 * N * 3, d_tmp, (size_t)nf * N * 3, cudaMemcpyDeviceToHost));
            }
            CUDA_OK(cudaFree(d_tmp));
            for (auto& s : L.slot) {
                s.runs = static_cast<dofs3d_run*>(dofs3d_pinned_alloc(sizeof(dofs3d_run) * (size_t)chunk * max_runs));
                s.n_runs = static_cast<int32_t*>(dofs3d_pinned_alloc(sizeof(int32_t) * chunk));
                s.n_boxes = static_cast<int32_t*>(dofs3d_pinned_alloc(sizeof(int32_t) * chunk));
                s.boxes = static_cast<dofs3d_box*>(dofs3d_pinned_alloc(sizeof(dofs3d_box) * (size_t)chunk * MAX_BOXES));
                s.stats = static_cast<dofs3d_stats*>(dofs3d_pinned_alloc(sizeof(dofs3d_stats) * chunk));
            }
            first_pair += L.pairs;
        }
        CUDA_OK(cudaMalloc(&G.d_packed, sizeof(dofs3d_box) * (size_t)n_lanes * cap));
        CUDA_OK(cudaMalloc(&G.d_counts, sizeof(int32_t) * n_lanes));
        CUDA_OK(cudaMalloc(&G.d_all_counts, sizeof(int32_t) * (size_t)n_gpus * n_lanes));
        CUDA_OK(cudaMalloc(&G.d_all_boxes, sizeof(dofs3d_box) * (size_t)n_gpus * n_lanes * cap));
        CUDA_OK(cudaMemset(G.d_packed, 0, sizeof(dofs3d_box) * (size_t)n_lanes * cap));
    }
    std::vector<ncclComm_t> comms(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int g = 0; g < n_gpus; ++g) devs[g] = g;
    NCCL_OK(ncclCommInitAll(comms.data(), n_gpus, devs.data()));
    // warm-up (untimed): the first chunk of a context allocates its flow and staging buffers, the first collective of a
    // communicator sets up its channels
    for (Gpu& G : gpus) {
        CUDA_OK(cudaSetDevice(G.device));
        for (Lane& L : G.lanes) {
            dofs3d_outputs o;
            std::memset(&o, 0, sizeof o);
            o.label_format = DOFS3D_LABELS_RLE;
            o.labels = L.slot[0].runs;
            o.n_runs = L.slot[0].n_runs;
            o.max_runs = max_runs;
            o.boxes = L.slot[0].boxes;
            o.n_boxes = L.slot[0].n_boxes;
            o.max_boxes = MAX_BOXES;
            int got = 0;
            DOFS_OK(L.ctx, dofs3d_stream_begin(L.ctx));
            DOFS_OK(L.ctx, dofs3d_stream_submit(L.ctx, L.frames, std::min(L.pairs, chunk) + 1, &o));
            DOFS_OK(L.ctx, dofs3d_stream_collect(L.ctx, &got));
        }
    }
    NCCL_OK(ncclGroupStart());
    for (int g = 0; g < n_gpus; ++g) {
        CUDA_OK(cudaSetDevice(g));
        NCCL_OK(ncclAllGather(gpus[g].d_counts, gpus[g].d_all_counts, n_lanes, ncclInt32, comms[g], 0));
    }
    NCCL_OK(ncclGroupEnd());
    for (int g = 0; g < n_gpus; ++g) {
        CUDA_OK(cudaSetDevice(g));
        CUDA_OK(cudaStreamSynchronize(0));
    }

    // ---- the timed region: every GPU streams its shard, then the gather
    auto collect = [&](Lane& L) {
        int got = 0;
        DOFS_OK(L.ctx, dofs3d_stream_collect(L.ctx, &got));
        const Lane::Slot& s = L.slot[L.chunks_collected & 1];
        for (int i = 0; i < got; ++i) {
            L.boxes_found += s.n_boxes[i];
            L.runs_found += s.n_runs[i];
        }
        L.collected_pairs += got;
        L.chunks_collected++;
    };
    auto run_gpu = [&](Gpu& G) {
        CUDA_OK(cudaSetDevice(G.device));
        const auto t0 = std::chrono::steady_clock::now();
        for (Lane& L : G.lanes) DOFS_OK(L.ctx, dofs3d_stream_begin(L.ctx));
        bool busy = true;
        while (busy) {  // round-robin over the lanes: one host thread keeps every context fed
            busy = false;
            for (Lane& L : G.lanes) {
                if (L.submitted_pairs >= L.pairs) continue;
                busy = true;
                if (L.chunks_submitted - L.chunks_collected == 2) collect(L);
                const bool first = L.chunks_submitted == 0;
                const int np = std::min(chunk, L.pairs - L.submitted_pairs);
                const int nf = first ? np + 1 : np;
                const uint8_t* src = L.frames + (size_t)(first ? 0 : L.submitted_pairs + 1) * N * 3;
                Lane::Slot& s = L.slot[L.chunks_submitted & 1];
                dofs3d_outputs o;
                std::memset(&o, 0, sizeof o);
                o.label_format = DOFS3D_LABELS_RLE;
                o.labels = s.runs;
                o.n_runs = s.n_runs;
                o.max_runs = max_runs;
                o.boxes = s.boxes;
                o.n_boxes = s.n_boxes;
                o.max_boxes = MAX_BOXES;
                o.stats = s.stats;
                DOFS_OK(L.ctx, dofs3d_stream_submit(L.ctx, src, nf, &o));
                L.submitted_pairs += np;
                L.chunks_submitted++;
            }
        }
        for (Lane& L : G.lanes)
            while (L.chunks_collected < L.chunks_submitted) collect(L);
        // the payload of the gather: the boxes of every lane's last chunk, packed on the device
        for (size_t l = 0; l < G.lanes.size(); ++l) {
            Lane& L = G.lanes[l];
            const int last = L.pairs - (L.chunks_submitted - 1) * chunk;
            DOFS_OK(L.ctx, dofs3d_pack_boxes_dev(L.ctx, last, G.d_packed + l * cap, cap, G.d_counts + l));
            DOFS_OK(L.ctx, dofs3d_sync(L.ctx));
        }
        G.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (Gpu& G : gpus) threads.emplace_back(run_gpu, std::ref(G));
    for (auto& t : threads) t.join();
    const auto t1 = std::chrono::steady_clock::now();
    // gather over NCCL: counts, then the packed boxes (fixed capacity per lane), on each GPU's default stream
    NCCL_OK(ncclGroupStart());
    for (int g = 0; g < n_gpus; ++g) {
        CUDA_OK(cudaSetDevice(g));
        NCCL_OK(ncclAllGather(gpus[g].d_counts, gpus[g].d_all_counts, n_lanes, ncclInt32, comms[g], 0));
        NCCL_OK(ncclAllGather(gpus[g].d_packed, gpus[g].d_all_boxes, sizeof(dofs3d_box) * (size_t)n_lanes * cap, ncclChar, comms[g], 0));
    }
    NCCL_OK(ncclGroupEnd());
    for (int g = 0; g < n_gpus; ++g) {
        CUDA_OK(cudaSetDevice(g));
        CUDA_OK(cudaStreamSynchronize(0));
    }
    const auto t2 = std::chrono::steady_clock::now();

    // ---- report; every GPU must hold the same gathered counts
    std::vector<int32_t> ref_counts((size_t)n_gpus * n_lanes), counts((size_t)n_gpus * n_lanes);
    long long total_boxes = 0, total_pairs = 0, gathered = 0;
    for (int g = 0; g < n_gpus; ++g) {
        CUDA_OK(cudaSetDevice(g));
        CUDA_OK(cudaMemcpy(counts.data(), gpus[g].d_all_counts, sizeof(int32_t) * counts.size(), cudaMemcpyDeviceToHost));
        if (g == 0) ref_counts = counts;
        else if (counts != ref_counts) {
            std::fprintf(stderr, "GPU %d holds different gathered counts\n", g);
            return 1;
        }
        long long b = 0, r = 0;
        int p = 0;
        for (const Lane& L : gpus[g].lanes) b += L.boxes_found, r += L.runs_found, p += L.collected_pairs;
        std::printf("gpu %d: %d pairs in %.3f s = %.1f pairs/s, %lld boxes, %.1f label runs per frame\n", g, p, gpus[g].seconds,
                    p / gpus[g].seconds, b, p ? (double)r / p : 0.0);
        total_boxes += b;
        total_pairs += p;
    }
    for (int32_t c : ref_counts) gathered += c;
    // the first gathered box of GPU 0's first lane must be the one its last chunk reported on the host
    std::vector<dofs3d_box> head(1);
    if (ref_counts[0] > 0) {
        CUDA_OK(cudaSetDevice(n_gpus - 1));
        CUDA_OK(cudaMemcpy(head.data(), gpus[n_gpus - 1].d_all_boxes, sizeof(dofs3d_box), cudaMemcpyDeviceToHost));
        const Lane& L = gpus[0].lanes[0];
        const Lane::Slot& s = L.slot[(L.chunks_submitted - 1) & 1];
        int i = 0;
        while (s.n_boxes[i] == 0) ++i;
        if (std::memcmp(&head[0], &s.boxes[(size_t)i * MAX_BOXES], sizeof(dofs3d_box)) != 0) {
            std::fprintf(stderr, "gathered box differs from the box the stream reported\n");
            return 1;
        }
    }
    const double stream_s = std::chrono::duration<double>(t1 - t0).count();
    const double gather_ms = 1e3 * std::chrono::duration<double>(t2 - t1).count();
    std::printf("total: %lld pairs on %d GPU(s) in %.3f s = %.1f pairs/s end to end (host frames in, labels + boxes out); "
                "NCCL gather of %lld boxes of the last chunks: %.3f ms; %lld boxes in all\n",
                total_pairs, n_gpus, stream_s + gather_ms / 1e3, total_pairs / (stream_s + gather_ms / 1e3), gathered, gather_ms,
                total_boxes);
    for (int g = 0; g < n_gpus; ++g) {
        ncclCommDestroy(comms[g]);
        for (Lane& L : gpus[g].lanes) dofs3d_destroy(L.ctx);
    }
    return 0;
}
