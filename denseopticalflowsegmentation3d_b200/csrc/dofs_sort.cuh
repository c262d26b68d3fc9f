// Stable LSD radix sort of (key, u32 payload) pairs, batched over frames (blockIdx.y = frame), for
// 32-bit and 64-bit keys.
//
// Used for the merge times (dofs_seg.cuh: the <= N-1 accepted edges on a 32-bit key, 4 passes, then an exact repair of
// the short runs that share a prefix), for the merge events (a stable sort of the time-ordered losers on a 32-bit
// (wave, winner root) key, 4 passes at 1080p), for the exact 64-bit fallback of the merge times, and — parity hook dofs3d_edges_sorted only — for the reference's whole edge list
// in std::multiset order (graph.cpp:55-60): ascending f64 weight (non-negative doubles order like their bit patterns),
// equal weights in insertion order, i.e. a STABLE sort of the insertion sequence by weight.
//
// One 8-bit digit per pass, three kernels per pass:
//   k_radix_hist     per-tile digit histogram             reads keys
//   k_radix_scan     per-digit exclusive scan over tiles   (tiny)
//   k_radix_scatter  stable in-tile ranking (warp match-any), reorder through shared memory so that
//                    each digit run leaves the block as one coalesced segment; reads and writes
//                    keys + payload
// HBM-bound: (2 * sizeof(key) + 8) + sizeof(key) bytes per element per pass.
// Every kernel loops over its tiles with stride gridDim.x and returns at once when *enable == 0, so a
// conditional sort can be enqueued with a small grid at the cost of a few empty launches.
#pragma once
#include "dofs_common.cuh"

#define RS_THREADS 256
#ifndef RS_ITEMS
#define RS_ITEMS 12
#endif
#ifndef RS_BLOCKS
#define RS_BLOCKS 3
#endif
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)
#define RS_BINS 256

// exclusive scan of one value per thread across a 256-thread block; also returns the block total
DOFS_D u32 rs_block_excl_scan(u32 v, u32* s_warp /* >= 8 */, u32* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    u32 wsum = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        u32 t = s_warp[w];
        if (w < warp) wsum += t;
        tot += t;
    }
    __syncthreads();
    if (total) *total = tot;
    return wsum + inc - v;
}

// lanes of the warp holding the same 9-bit value (8-bit digit + the "out of range" flag), by ballots: one vote per bit
#ifndef RS_USE_MATCH_ANY
#define RS_USE_MATCH_ANY 0  // A/B: the MATCH.ANY instruction instead of nine votes
#endif
#ifndef RS_MATCH_PTX
#define RS_MATCH_PTX 1
#endif
template <int BITS = 9>
DOFS_D u32 rs_match9(u32 d) {
#if RS_USE_MATCH_ANY
    return __match_any_sync(0xffffffffu, d);
#endif
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
#if RS_MATCH_PTX
        // four instructions per bit (test -> predicate, vote, complement under the predicate, and) where the compiler's
        // select form takes six: the ranking is a third of the one-sweep kernel's instructions
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " .reg .b32 m;\n"
            " setp.ne.u32 p, %1, 0;\n"
            " vote.sync.ballot.b32 m, p, 0xffffffff;\n"
            " @!p not.b32 m, m;\n"
            " and.b32 %0, %0, m;\n"
            "}\n"
            : "+r"(peers)
            : "r"(d & (1u << b)));
#else
        const bool bit = (d >> b) & 1u;
        const u32 m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
#endif
    }
    return peers;
}

// tile_hist layout: [frame][digit][tile]
template <typename K>
__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const K* __restrict__ keys, size_t frame_stride, u32* __restrict__ tile_hist, int n, int shift,
             int num_tiles, const int* __restrict__ enable) {
    if (enable && *enable == 0) return;
    __shared__ u32 s_hist[RS_BINS];
    const int frame = blockIdx.y;
    keys += (size_t)frame * frame_stride;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        s_hist[threadIdx.x] = 0;
        __syncthreads();
        const int base = tile * RS_TILE;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            int idx = base + i * RS_THREADS + threadIdx.x;
            if (idx < n) atomicAdd(&s_hist[(u32)(keys[idx] >> shift) & 255u], 1u);
        }
        __syncthreads();
        tile_hist[((size_t)frame * RS_BINS + threadIdx.x) * num_tiles + tile] = s_hist[threadIdx.x];
        __syncthreads();
    }
}

// grid (256 digits, frames): exclusive scan of row [digit][0..num_tiles) in place; digit_tot[frame][digit]
__global__ void __launch_bounds__(RS_THREADS)
k_radix_scan(u32* __restrict__ tile_hist, u32* __restrict__ digit_tot, int num_tiles, const int* __restrict__ enable) {
    if (enable && *enable == 0) return;
    __shared__ u32 s_warp[RS_WARPS];
    const int digit = blockIdx.x, frame = blockIdx.y;
    u32* row = tile_hist + ((size_t)frame * RS_BINS + digit) * num_tiles;
    u32 carry = 0;
    for (int base = 0; base < num_tiles; base += RS_THREADS) {
        int i = base + threadIdx.x;
        u32 v = i < num_tiles ? row[i] : 0u;
        u32 tot;
        u32 ex = rs_block_excl_scan(v, s_warp, &tot);
        if (i < num_tiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) digit_tot[frame * RS_BINS + digit] = carry;
}

// dynamic shared memory: keys[RS_TILE] | vals[RS_TILE] u32 | whist[RS_WARPS][256] | dlocal[256] | dbase[256] | warp[8]
template <typename K>
constexpr int rs_smem_bytes() {
    return RS_TILE * (int)sizeof(K) + RS_TILE * 4 + RS_WARPS * RS_BINS * 4 + RS_BINS * 4 * 2 + 64;
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS, 4)
k_radix_scatter(const K* __restrict__ keys_in, const u32* __restrict__ vals_in, K* __restrict__ keys_out,
                u32* __restrict__ vals_out, size_t frame_stride, const u32* __restrict__ tile_offs,
                const u32* __restrict__ digit_tot, int n, int shift, int num_tiles, int iota_vals,
                const int* __restrict__ enable) {
    if (enable && *enable == 0) return;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    K* s_keys = reinterpret_cast<K*>(rs_smem);
    u32* s_vals = reinterpret_cast<u32*>(rs_smem + RS_TILE * sizeof(K));
    u32* s_whist = s_vals + RS_TILE;           // [warp][digit]
    u32* s_dlocal = s_whist + RS_WARPS * RS_BINS;  // exclusive prefix of the digit inside this tile
    u32* s_dbase = s_dlocal + RS_BINS;             // global position of the tile's first element of the digit
    u32* s_warp = s_dbase + RS_BINS;

    const int frame = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    keys_in += (size_t)frame * frame_stride;
    vals_in += (size_t)frame * frame_stride;
    keys_out += (size_t)frame * frame_stride;
    vals_out += (size_t)frame * frame_stride;
    // global base of every digit (exclusive scan of the digit totals), the same for all tiles
    const u32 gtot = digit_tot[frame * RS_BINS + tid];
    const u32 gbase = rs_block_excl_scan(gtot, s_warp, nullptr);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) s_whist[i] = 0;
        __syncthreads();

        // warp-striped load: element order inside the tile is (warp, item, lane) == memory order
        const int wbase = tile * RS_TILE + warp * (32 * RS_ITEMS);
        K key[RS_ITEMS];
        u32 val[RS_ITEMS];
        u32 rnk[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            int idx = wbase + i * 32 + lane;
            bool ok = idx < n;
            key[i] = ok ? keys_in[idx] : (K)~(K)0;
            val[i] = ok ? (iota_vals ? (u32)idx : vals_in[idx]) : 0u;
        }
        u32* my_hist = s_whist + warp * RS_BINS;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            int idx = wbase + i * 32 + lane;
            bool ok = idx < n;
            u32 d = ok ? ((u32)(key[i] >> shift) & 255u) : 256u;  // out-of-range lanes form their own group
            u32 peers = rs_match9(d);
            u32 below = __popc(peers & ((1u << lane) - 1u));
            int leader = __ffs(peers) - 1;
            u32 pre = 0;
            if (lane == leader && ok) {
                pre = my_hist[d];
                my_hist[d] = pre + __popc(peers);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rnk[i] = pre + below;
            __syncwarp();
        }
        __syncthreads();

        // digit = tid: exclusive scan over the warps of this digit, tile total of the digit
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            u32 c = s_whist[w * RS_BINS + tid];
            s_whist[w * RS_BINS + tid] = run;
            run += c;
        }
        u32 tile_total;
        u32 dl = rs_block_excl_scan(run, s_warp, &tile_total);
        s_dlocal[tid] = dl;
        s_dbase[tid] = gbase + tile_offs[((size_t)frame * RS_BINS + tid) * num_tiles + tile];
        __syncthreads();

#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            int idx = wbase + i * 32 + lane;
            if (idx < n) {
                u32 d = (u32)(key[i] >> shift) & 255u;
                u32 pos = s_dlocal[d] + my_hist[d] + rnk[i];
                s_keys[pos] = key[i];
                s_vals[pos] = val[i];
            }
        }
        __syncthreads();
        for (u32 j = tid; j < tile_total; j += RS_THREADS) {
            K k = s_keys[j];
            u32 d = (u32)(k >> shift) & 255u;
            u32 dst = s_dbase[d] + (j - s_dlocal[d]);
            keys_out[dst] = k;
            vals_out[dst] = s_vals[j];
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// One-sweep variant (decoupled look-back) used by the unconditional sorts: the digit histograms of ALL passes are
// taken in one read of the keys (k_radix_hist_all + k_radix_bases), and every scatter pass gets the position of its
// tile among the tiles of its frame from a chained look-back over per-(tile, digit) status words instead of from a
// separate histogram + scan pass:
//   status word = flag (bits 31:30: 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix) | count (30 bits)
// Tiles take a ticket from one global counter (frame = ticket % frames, tile = ticket / frames), so a tile only ever
// waits for tiles that started before it, and the tiles in flight are spread over all frames (short look-back chains).
// Every spin is bounded: on a time-out *err is raised and the tile carries on with what it has (no hang).
// ---------------------------------------------------------------------------------------------
#define RS_FLAG_AGG 0x40000000u
#define RS_FLAG_PREFIX 0x80000000u
#define RS_COUNT_MASK 0x3FFFFFFFu
#define RS_SPIN_LIMIT (1 << 22)
#define RS_MAX_PASSES 8

// ghist layout: [frame][pass][256]
template <typename K>
__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist_all(const K* __restrict__ keys, size_t frame_stride, u32* __restrict__ ghist, int n, int n_passes) {
    __shared__ u32 s_hist[RS_MAX_PASSES * RS_BINS];
    const int frame = blockIdx.y;
    keys += (size_t)frame * frame_stride;
    for (int i = threadIdx.x; i < n_passes * RS_BINS; i += RS_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (int base = blockIdx.x * RS_TILE; base < n; base += gridDim.x * RS_TILE) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int idx = base + i * RS_THREADS + threadIdx.x;
            if (idx < n) {
                const K k = keys[idx];
                for (int p = 0; p < n_passes; ++p) atomicAdd(&s_hist[p * RS_BINS + ((u32)(k >> (8 * p)) & 255u)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * RS_BINS; i += RS_THREADS) {
        const u32 c = s_hist[i];
        if (c) atomicAdd(&ghist[(size_t)frame * RS_MAX_PASSES * RS_BINS + i], c);
    }
}

// grid (n_passes, frames): exclusive scan of the 256 digit totals -> first output position of every digit
__global__ void __launch_bounds__(RS_THREADS)
k_radix_bases(const u32* __restrict__ ghist, u32* __restrict__ gbase) {
    __shared__ u32 s_warp[RS_WARPS];
    const size_t o = ((size_t)blockIdx.y * RS_MAX_PASSES + blockIdx.x) * RS_BINS + threadIdx.x;
    gbase[o] = rs_block_excl_scan(ghist[o], s_warp, nullptr);
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS, RS_BLOCKS)
k_radix_onesweep(const K* __restrict__ keys_in, const u32* __restrict__ vals_in, K* __restrict__ keys_out,
                 u32* __restrict__ vals_out, size_t frame_stride, const u32* __restrict__ gbase /* this pass: [frame][..] */,
                 u32* __restrict__ status /* [frame][tile][256], zeroed */, int* __restrict__ ticket /* zeroed */,
                 int* __restrict__ err, int n, int shift, int num_tiles, int frames, int iota_vals) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    K* s_keys = reinterpret_cast<K*>(rs_smem);
    u32* s_vals = reinterpret_cast<u32*>(rs_smem + RS_TILE * sizeof(K));
    u32* s_whist = s_vals + RS_TILE;               // [warp][digit]
    u32* s_dlocal = s_whist + RS_WARPS * RS_BINS;  // exclusive prefix of the digit inside this tile
    u32* s_dbase = s_dlocal + RS_BINS;             // global position of the tile's first element of the digit
    u32* s_warp = s_dbase + RS_BINS;
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1);
    for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) s_whist[i] = 0;
    __syncthreads();
    const int frame = s_ticket % frames, tile = s_ticket / frames;
    if (tile >= num_tiles) return;
    keys_in += (size_t)frame * frame_stride;
    vals_in += (size_t)frame * frame_stride;
    keys_out += (size_t)frame * frame_stride;
    vals_out += (size_t)frame * frame_stride;

    // warp-striped load: element order inside the tile is (warp, item, lane) == memory order
    const int wbase = tile * RS_TILE + warp * (32 * RS_ITEMS);
    const bool full = (tile + 1) * RS_TILE <= n;  // block-uniform: every tile of a frame but the last one
    K key[RS_ITEMS];
    u32 val[RS_ITEMS];
    u32 rnk[RS_ITEMS];
    u32* my_hist = s_whist + warp * RS_BINS;
    if (full) {  // no range checks, eight votes per item
        const K* kin = keys_in + wbase + lane;
        const u32* vin = vals_in + wbase + lane;
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            key[i] = kin[i * 32];
            val[i] = iota_vals ? (u32)(wbase + i * 32 + lane) : vin[i * 32];
        }
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const u32 d = (u32)(key[i] >> shift) & 255u;
            const u32 peers = rs_match9<8>(d);
            const u32 below = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            u32 pre = 0;
            if (lane == leader) {
                pre = my_hist[d];
                my_hist[d] = pre + __popc(peers);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rnk[i] = pre + below;
            __syncwarp();
        }
    } else {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int idx = wbase + i * 32 + lane;
            const bool ok = idx < n;
            key[i] = ok ? keys_in[idx] : (K)~(K)0;
            val[i] = ok ? (iota_vals ? (u32)idx : vals_in[idx]) : 0u;
        }
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const bool ok = wbase + i * 32 + lane < n;
            const u32 d = ok ? ((u32)(key[i] >> shift) & 255u) : 256u;  // out-of-range lanes form their own group
            const u32 peers = rs_match9<9>(d);
            const u32 below = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            u32 pre = 0;
            if (lane == leader && ok) {
                pre = my_hist[d];
                my_hist[d] = pre + __popc(peers);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rnk[i] = pre + below;
            __syncwarp();
        }
    }
    __syncthreads();

    // digit = tid: exclusive scan over the warps of this digit, tile total of the digit
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const u32 c = s_whist[w * RS_BINS + tid];
        s_whist[w * RS_BINS + tid] = run;
        run += c;
    }
    // publish the tile's count of this digit, then look back over the earlier tiles of the frame
    volatile u32* st = status + ((size_t)frame * num_tiles) * RS_BINS + tid;
    u32 before = 0;
    if (tile == 0) {
        st[0] = RS_FLAG_PREFIX | run;
    } else {
        st[(size_t)tile * RS_BINS] = RS_FLAG_AGG | run;
        for (int t = tile - 1; t >= 0; --t) {
            u32 w = 0;
            int spin = 0;
            do {
                w = st[(size_t)t * RS_BINS];
            } while ((w & ~RS_COUNT_MASK) == 0u && ++spin < RS_SPIN_LIMIT);
            if ((w & ~RS_COUNT_MASK) == 0u) {  // timed out: never expected, but never hang
                atomicExch(err, 1);
                break;
            }
            before += w & RS_COUNT_MASK;
            if (w & RS_FLAG_PREFIX) break;
        }
        st[(size_t)tile * RS_BINS] = RS_FLAG_PREFIX | (before + run);
    }
    u32 tile_total;
    const u32 dl = rs_block_excl_scan(run, s_warp, &tile_total);
    s_dlocal[tid] = dl;
    s_dbase[tid] = gbase[(size_t)frame * RS_MAX_PASSES * RS_BINS + tid] + before;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        if (full || wbase + i * 32 + lane < n) {
            const u32 d = (u32)(key[i] >> shift) & 255u;
            const u32 pos = s_dlocal[d] + my_hist[d] + rnk[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    for (u32 j = tid; j < tile_total; j += RS_THREADS) {
        const K k = s_keys[j];
        const u32 d = (u32)(k >> shift) & 255u;
        const u32 dst = s_dbase[d] + (j - s_dlocal[d]);
        keys_out[dst] = k;
        vals_out[dst] = s_vals[j];
    }
}
