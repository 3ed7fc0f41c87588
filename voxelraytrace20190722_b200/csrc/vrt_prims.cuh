// vrt_prims.cuh -- device-wide primitives used by the octree build:
// exclusive scan (uint32) and a stable LSD radix sort on 64-bit keys.
// Hand-written for sm_100a (no CUB/Thrust): 32-wide warps, __match_any_sync
// ranking, shared-memory digit counters, grid sizes derived from the tile count.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "vrt_internal.h"

namespace vrt {

// ---------------------------------------------------------------------------
// Exclusive scan of uint32 (in place allowed).  Tile = 256 threads x 8.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
k_scan_tile(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
            uint32_t* __restrict__ tile_sums, uint64_t n)
{
        __shared__ uint32_t warp_tot[kScanThreads / 32];
        const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
        uint32_t v[kScanItems];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
                uint64_t k = base + i;
                v[i] = (k < n) ? in[k] : 0u;
                sum += v[i];
        }
        // warp inclusive scan of per-thread sums
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                        inc += t;
        }
        if (lane == 31)
                warp_tot[w] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int i = 0; i < kScanThreads / 32; ++i) {
                uint32_t t = warp_tot[i];
                if (i < w)
                        woff += t;
                total += t;
        }
        uint32_t run = woff + inc - sum;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
                uint64_t k = base + i;
                if (k < n)
                        out[k] = run;
                run += v[i];
        }
        if (threadIdx.x == 0 && tile_sums)
                tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_add(uint32_t* __restrict__ data, const uint32_t* __restrict__ tile_offs, uint64_t n)
{
        const uint32_t off = tile_offs[blockIdx.x];
        const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
                uint64_t k = base + i;
                if (k < n)
                        data[k] += off;
        }
}

// Single-pass exclusive scan (round 2): decoupled look-back.  A tile takes a ticket (atomic counter, so
// tiles start in ticket order and every predecessor of a running tile is running or done), scans its
// 2048 elements, publishes its aggregate as (value | FLAG_AGG << 32), looks back over its predecessors'
// status words until it meets an inclusive prefix, publishes its own inclusive prefix (FLAG_INC) and
// writes the result: one kernel and 8 n bytes of traffic instead of tile scan + scan of the sums + add
// pass (up to 5 launches and 16 n bytes).  status[] and the ticket are zeroed by one memset per scan.
constexpr unsigned long long kScanAgg = 1ull << 32, kScanInc = 2ull << 32;

__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket)
{
        __shared__ uint32_t warp_tot[kScanThreads / 32];
        __shared__ uint32_t s_tile, s_prefix;
        if (threadIdx.x == 0)
                s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
        uint32_t v[kScanItems];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
                const uint64_t k = base + i;
                v[i] = (k < n) ? in[k] : 0u;
                sum += v[i];
        }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                        inc += t;
        }
        if (lane == 31)
                warp_tot[w] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int i = 0; i < kScanThreads / 32; ++i) {
                const uint32_t t = warp_tot[i];
                if (i < w)
                        woff += t;
                total += t;
        }
        // ---- publish the aggregate, look back, publish the inclusive prefix (warp 0) ----
        if (w == 0) {
                volatile unsigned long long* st = status;
                if (lane == 0) {
                        st[tile] = (tile == 0 ? kScanInc : kScanAgg) | total;
                        __threadfence();
                }
                uint32_t prefix = 0;
                if (tile != 0) {
                        int j = (int)tile - 1 - lane;  // the 32 predecessors closest to this tile, lane 0 = nearest
                        for (;;) {
                                unsigned long long sv = kScanInc;  // lanes before tile 0 read an empty inclusive prefix
                                if (j >= 0) {
                                        do {
                                                sv = st[j];
                                        } while ((sv >> 32) == 0ull);
                                }
                                const uint32_t has_inc = __ballot_sync(0xffffffffu, (sv >> 32) == 2ull);
                                // add everything up to and including the nearest predecessor with an inclusive prefix
                                const int stop = has_inc ? (__ffs((int)has_inc) - 1) : 31;
                                uint32_t add = (lane <= stop) ? (uint32_t)sv : 0u;
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1)
                                        add += __shfl_xor_sync(0xffffffffu, add, o);
                                prefix += add;
                                if (has_inc)
                                        break;
                                j -= 32;
                        }
                        if (lane == 0) {
                                __threadfence();
                                st[tile] = kScanInc | (unsigned long long)(uint32_t)(prefix + total);
                        }
                }
                if (lane == 0)
                        s_prefix = prefix;
        }
        __syncthreads();
        uint32_t run = s_prefix + woff + inc - sum;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
                const uint64_t k = base + i;
                if (k < n)
                        out[k] = run;
                run += v[i];
        }
}

// scratch must hold at least scan_scratch_elems(n) uint32.
inline uint64_t scan_scratch_elems(uint64_t n)
{
        const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
        return 2 * tiles + 16;  // ticket + padding to 8 bytes + one 64-bit status word per tile
}

// Exclusive scan (in place allowed); the total can be read from out[n-1] + in[n-1] by the caller.
inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch,
                               cudaStream_t s)
{
        if (n == 0)
                return;
        const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
        if (tiles == 1) {
                k_scan_tile<<<1, kScanThreads, 0, s>>>(in, out, nullptr, n);
                count_launch();
                return;
        }
        // scratch: [ticket][pad..] then the status words at the next 8-byte boundary
        uint32_t* ticket = scratch;
        unsigned long long* status =
                reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(scratch) + 8 + 7) & ~(uintptr_t)7);
        const size_t bytes = (size_t)(reinterpret_cast<char*>(status + tiles) - reinterpret_cast<char*>(scratch));
        cudaMemsetAsync(scratch, 0, bytes, s);
        k_scan_lookback<<<(unsigned)tiles, kScanThreads, 0, s>>>(in, out, n, status, ticket);
        count_launch();
}

// ---------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits, 64-bit keys.
// Tile = 256 threads x 16 keys; warp w owns the contiguous 512-key slice
// [w*512,(w+1)*512) of its tile and walks it 32 keys at a time, so the order
// inside a tile is (warp, round, lane) == memory order and the local ranking is
// stable without any cross-warp synchronisation inside the rounds.
// ---------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;  // 4096
constexpr int kSortWarpSlice = 32 * kSortRounds;       // 512

__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const unsigned long long* __restrict__ keys, uint64_t n, int shift,
            uint32_t* __restrict__ hist /* [256][tiles] */, uint32_t tiles)
{
        __shared__ uint32_t h[256];
        h[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll 4
        for (int r = 0; r < kSortRounds; ++r) {
                uint64_t k = base + (uint64_t)r * kSortThreads + threadIdx.x;
                if (k < n)
                        atomicAdd(&h[(uint32_t)(keys[k] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(uint64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// Scatter pass.  The tile is first reordered by digit in shared memory (stable: digit, then
// memory order), then written out position by position, so that the keys of one digit leave
// as one contiguous run (16 keys = 128 B on average) instead of one 8-byte store per sector.
__global__ void __launch_bounds__(kSortThreads)
k_sort_scatter(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
               uint64_t n, int shift, const uint32_t* __restrict__ offs /* scanned [256][tiles] */,
               uint32_t tiles)
{
        __shared__ uint32_t cnt[kSortWarps][256];
        __shared__ uint32_t gbase[256];  // global position of the digit's run minus its tile-local start
        __shared__ uint32_t wtot[kSortWarps];
        __shared__ unsigned long long stage[kSortTile];
        for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads)
                (&cnt[0][0])[i] = 0;
        __syncthreads();
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const uint64_t tbase = (uint64_t)blockIdx.x * kSortTile;
        const uint64_t wbase = tbase + (uint64_t)w * kSortWarpSlice;
        unsigned long long key[kSortRounds];
        uint32_t rank[kSortRounds];
#pragma unroll
        for (int r = 0; r < kSortRounds; ++r) {
                uint64_t k = wbase + (uint64_t)r * 32 + lane;
                bool valid = k < n;
                key[r] = valid ? in[k] : ~0ull;
                uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;
                uint32_t peers = __match_any_sync(0xffffffffu, d);
                uint32_t before = __popc(peers & ((1u << lane) - 1u));
                uint32_t old = 0;
                if (valid)
                        old = cnt[w][d];
                __syncwarp();
                if (valid && before == 0)
                        cnt[w][d] = old + __popc(peers);
                __syncwarp();
                rank[r] = old + before;
        }
        __syncthreads();
        // per digit d (= threadIdx.x): exclusive prefix over the warps, then the tile-local start of
        // the digit (exclusive scan of the digit totals over the 256 threads)
        {
                const uint32_t d = threadIdx.x;
                uint32_t run = 0;
#pragma unroll
                for (int i = 0; i < kSortWarps; ++i) {
                        uint32_t c = cnt[i][d];
                        cnt[i][d] = run;
                        run += c;
                }
                uint32_t inc = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o)
                                inc += t;
                }
                if (lane == 31)
                        wtot[w] = inc;
                __syncthreads();
                uint32_t woff = 0;
#pragma unroll
                for (int i = 0; i < kSortWarps; ++i)
                        woff += (i < w) ? wtot[i] : 0u;
                const uint32_t dstart = woff + inc - run;
#pragma unroll
                for (int i = 0; i < kSortWarps; ++i)
                        cnt[i][d] += dstart;
                gbase[d] = offs[(uint64_t)d * tiles + blockIdx.x] - dstart;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSortRounds; ++r) {
                uint64_t k = wbase + (uint64_t)r * 32 + lane;
                if (k < n) {
                        uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
                        stage[cnt[w][d] + rank[r]] = key[r];
                }
        }
        __syncthreads();
        const uint32_t tile_n = (uint32_t)((n - tbase < (uint64_t)kSortTile) ? (n - tbase) : (uint64_t)kSortTile);
#pragma unroll 4
        for (uint32_t j = threadIdx.x; j < tile_n; j += kSortThreads) {
                const unsigned long long kv = stage[j];
                const uint32_t d = (uint32_t)(kv >> shift) & 255u;
                out[gbase[d] + j] = kv;
        }
}

inline uint64_t sort_hist_elems(uint64_t n)
{
        uint64_t tiles = (n + kSortTile - 1) / kSortTile;
        return 256ull * tiles;
}

// Sorts keys on bits [lo_bit, hi_bit).  Result pointer is returned through
// *sorted (either a or b).  hist must hold sort_hist_elems(n) uint32, scan_tmp
// scan_scratch_elems(256*tiles) uint32.
inline void radix_sort_u64(unsigned long long* a, unsigned long long* b, uint64_t n, int lo_bit,
                           int hi_bit, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t s,
                           unsigned long long** sorted)
{
        *sorted = a;
        if (n < 2)
                return;
        uint32_t tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
        for (int shift = lo_bit; shift < hi_bit; shift += 8) {
                k_sort_hist<<<tiles, kSortThreads, 0, s>>>(a, n, shift, hist, tiles);
                count_launch();
                exclusive_scan_u32(hist, hist, 256ull * tiles, scan_tmp, s);
                k_sort_scatter<<<tiles, kSortThreads, 0, s>>>(a, b, n, shift, hist, tiles);
                count_launch();
                unsigned long long* t = a;
                a = b;
                b = t;
        }
        *sorted = a;
}

}  // namespace vrt
