// K2: tile binning and per-tile depth sort.
//
// Upstream (rasterizer_impl.cu) builds one global list of 64-bit keys
// (tile<<32 | depth bits), runs a device-wide radix sort and reads the
// instance count back to the host.  Here: K1 has already counted instances
// per tile; (a) one block scans the per-tile counts into tile ranges,
// (b) every surfel scatters (depth bits<<32 | surfel) into its tiles' ranges,
// (c) one warp per tile sorts its range in registers (tiles above 512 instances: one block, shared memory).  Because the keys
// are unique, the sorted order equals upstream's stable sort of the
// duplication order (ascending surfel index) -- bit-exact -- without a global
// sort, without a host read-back and with ~3 passes over 8-byte keys instead
// of ~10 over 12-byte pairs.
#include "raster_common.cuh"

#define SCAN_THREADS 1024
#define SORT_WARP_MAX 512          // tiles up to this many instances are sorted by a single warp in registers

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles_kernel(RasterDims d, RasterWs ws)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    __shared__ int s_big;
    const int n = d.NV * d.T;
    int big = 0;                                   // tiles of this thread that need the block-level sort
    if (threadIdx.x == 0) { s_carry = 0; s_big = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        uint32_t rep[GA_TILE_REPLICAS];
        uint32_t v = 0u;
        if (i < n) {
            const uint4 *src = reinterpret_cast<const uint4 *>(ws.tile_count + (size_t)i * GA_TILE_REPLICAS);
#pragma unroll
            for (int q = 0; q < GA_TILE_REPLICAS / 4; q++) {
                const uint4 c = src[q];
                rep[4 * q] = c.x; rep[4 * q + 1] = c.y; rep[4 * q + 2] = c.z; rep[4 * q + 3] = c.w;
            }
#pragma unroll
            for (int q = 0; q < GA_TILE_REPLICAS; q++) v += rep[q];
            big += v > SORT_WARP_MAX;
        }
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t wv = s_warp[lane], wx = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, wx, o);
                if (lane >= o) wx += y;
            }
            s_warp[lane] = wx - wv;   // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + x - v;
        if (i < n) {
            ws.tile_start[i] = excl;
            // every replica becomes the absolute fill cursor of its own sub-range of the tile's slots
            uint32_t run = excl;
            uint32_t cur[GA_TILE_REPLICAS];
#pragma unroll
            for (int q = 0; q < GA_TILE_REPLICAS; q++) { cur[q] = run; run += rep[q]; }
            uint4 *dst = reinterpret_cast<uint4 *>(ws.tile_count + (size_t)i * GA_TILE_REPLICAS);
#pragma unroll
            for (int q = 0; q < GA_TILE_REPLICAS / 4; q++)
                dst[q] = make_uint4(cur[4 * q], cur[4 * q + 1], cur[4 * q + 2], cur[4 * q + 3]);
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (big) atomicAdd(&s_big, big);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = s_carry;
        ws.tile_start[n] = total;
        ws.status[0] = (int32_t)total;
        ws.status[1] = ((int64_t)total > d.max_instances) ? 1 : 0;
        ws.status[3] = s_big;                      // 0: sort_tiles_kernel has nothing to do and exits at once
    }
}

__global__ void __launch_bounds__(256)
scatter_kernel(RasterDims d, RasterWs ws)
{
    const int view = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.P || ws.status[1]) return;
    const size_t vi = (size_t)view * d.P + i;
    const uint32_t r = ws.rect[vi];
    if (r == 0) return;
    const int x0 = r & 255, y0 = (r >> 8) & 255, x1 = (r >> 16) & 255, y1 = r >> 24;
    const unsigned long long key =
        ((unsigned long long)__float_as_uint(ws.depth[vi]) << 32) | (unsigned long long)(uint32_t)i;
    // same replica as in K1 (same launch geometry: 256 threads, surfel i -> thread i % 256), so every replica
    // receives exactly the instances it counted
    const int rep = (threadIdx.x >> 5) & (GA_TILE_REPLICAS - 1);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const size_t t = (size_t)view * d.T + y * d.gx + x;
            const uint32_t slot = atomicAdd(&ws.tile_count[t * GA_TILE_REPLICAS + rep], 1u);
            ws.keys[slot] = key;
        }
}

// Ascending-only bitonic network (flip + half-cleaners): every compare-exchange
// moves the minimum to the lower index, so virtual +inf padding beyond n never
// moves and n need not be a power of two.  Works on shared or global memory.
__device__ __forceinline__ void block_bitonic_sort(unsigned long long *a, int n)
{
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
        // flip step: partner = mirror inside the k-block
        for (int t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
            const int blk = t / (k / 2), off = t % (k / 2);
            const int i = blk * k + off, p = blk * k + k - 1 - off;
            if (p < n) {
                unsigned long long x = a[i], y = a[p];
                if (x > y) { a[i] = y; a[p] = x; }
            }
        }
        __syncthreads();
        for (int j = k / 4; j >= 1; j >>= 1) {
            for (int t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
                const int i = (t / j) * 2 * j + (t % j), p = i + j;
                if (p < n) {
                    unsigned long long x = a[i], y = a[p];
                    if (x > y) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

#define SORT_SMEM_KEYS 4096

// Warp-level bitonic sort of up to 32*KPL keys held in registers, element i = r*32 + lane (striped, so global
// loads/stores are coalesced): partner distances < 32 are shuffles, distances >= 32 stay inside the lane.
// No shared memory and no block barrier.  Only the in-lane stages (static register indices) are unrolled; the
// shuffle stages run as a loop over the distance -- fully unrolled, the four instantiations came to 27k
// instructions and the kernel spent half its time on instruction-cache misses (profiles/r01_raster_small.md).
template <int KPL>
__device__ __forceinline__ void warp_shuffle_stages(unsigned long long (&v)[KPL], int lane, int k, int jstart)
{
#pragma unroll 1
    for (int j = jstart; j >= 1; j >>= 1) {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < KPL; r++) {
            const bool asc = ((((r << 5) | lane) & k) == 0);
            const unsigned long long mine = v[r];
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, j);
            const bool keep_min = (lower == asc);
            v[r] = keep_min ? (mine < other ? mine : other) : (mine > other ? mine : other);
        }
    }
}

template <int KPL>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&v)[KPL], int lane)
{
    constexpr int N = 32 * KPL;
#pragma unroll 1
    for (int k = 2; k <= 32; k <<= 1) warp_shuffle_stages<KPL>(v, lane, k, k >> 1);
#pragma unroll
    for (int k = 64; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 32; j >>= 1) {
            const int jr = j >> 5;
#pragma unroll
            for (int r = 0; r < KPL; r++) {
                if ((r & jr) == 0) {
                    const bool asc = (((r << 5) & k) == 0);              // k >= 64 here: decided by r alone
                    unsigned long long a = v[r], b = v[r | jr];
                    const bool sw = asc ? (a > b) : (a < b);
                    v[r] = sw ? b : a;
                    v[r | jr] = sw ? a : b;
                }
            }
        }
        warp_shuffle_stages<KPL>(v, lane, k, 16);
    }
}

template <int KPL>
__device__ __forceinline__ void warp_sort_tile(unsigned long long *gk, uint32_t *gid, int n, int lane)
{
    unsigned long long v[KPL];
#pragma unroll
    for (int r = 0; r < KPL; r++) {
        const int i = r * 32 + lane;
        v[r] = i < n ? gk[i] : 0xffffffffffffffffull;
    }
    warp_bitonic_sort<KPL>(v, lane);
#pragma unroll
    for (int r = 0; r < KPL; r++) {
        const int i = r * 32 + lane;
        if (i < n) {
            gk[i] = v[r];
            gid[i] = (uint32_t)(v[r] & 0xffffffffull);
        }
    }
}

// one warp per tile (tiles with more than SORT_WARP_MAX instances are left to sort_tiles_kernel)
#ifndef SORT_WARP_THREADS
#define SORT_WARP_THREADS 256
#endif
#ifndef SORT_WARP_CTAS
#define SORT_WARP_CTAS (512 / SORT_WARP_THREADS)
#endif
__global__ void __launch_bounds__(SORT_WARP_THREADS, SORT_WARP_CTAS)
sort_tiles_warp_kernel(RasterDims d, RasterWs ws)
{
    if (ws.status[1]) return;
    const size_t t = (size_t)blockIdx.x * (SORT_WARP_THREADS / 32) + (threadIdx.x >> 5);
    if (t >= (size_t)d.NV * d.T) return;
    const int lane = threadIdx.x & 31;
    const uint32_t start = ws.tile_start[t], end = ws.tile_start[t + 1];
    const int n = (int)(end - start);
    if (n == 0 || n > SORT_WARP_MAX) return;
    unsigned long long *gk = ws.keys + start;
    uint32_t *gid = ws.ids + start;
    if (n <= 32) warp_sort_tile<1>(gk, gid, n, lane);
    else if (n <= 128) warp_sort_tile<4>(gk, gid, n, lane);
    else if (n <= 256) warp_sort_tile<8>(gk, gid, n, lane);
    else warp_sort_tile<16>(gk, gid, n, lane);
}

// tiles with more than SORT_WARP_MAX instances: one block per tile, a small persistent grid walks all tiles
__global__ void __launch_bounds__(256)
sort_tiles_kernel(RasterDims d, RasterWs ws)
{
    __shared__ unsigned long long s_keys[SORT_SMEM_KEYS];
    if (ws.status[1] || ws.status[3] == 0) return;          // overflow, or no tile above the warp-sort limit
    const size_t tiles = (size_t)d.NV * d.T;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t start = ws.tile_start[t], end = ws.tile_start[t + 1];
        const int n = (int)(end - start);
        if (n <= SORT_WARP_MAX) continue;               // done by sort_tiles_warp_kernel (block-uniform branch)
        unsigned long long *gk = ws.keys + start;
        if (n <= SORT_SMEM_KEYS) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = gk[i];
            __syncthreads();
            block_bitonic_sort(s_keys, n);
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned long long k = s_keys[i];
                gk[i] = k;
                ws.ids[start + i] = (uint32_t)(k & 0xffffffffull);
            }
            __syncthreads();                            // s_keys is reused by the next tile
        } else {
            // rare: a tile with more instances than fit in shared memory is sorted
            // in place in global memory (L2 resident) by the same network.
            if (threadIdx.x == 0) atomicAdd(&ws.status[2], 1);
            block_bitonic_sort(gk, n);
            for (int i = threadIdx.x; i < n; i += blockDim.x)
                ws.ids[start + i] = (uint32_t)(gk[i] & 0xffffffffull);
        }
    }
}

cudaError_t ga_launch_binning(const RasterDims &d, const RasterWs &w, cudaStream_t s, int32_t *status_host,
                              cudaEvent_t status_event)
{
    scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>(d, w);
    if (status_host) {
        cudaError_t e = cudaMemcpyAsync(status_host, w.status, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
    }
    if (status_event) {
        cudaError_t e = cudaEventRecord(status_event, s);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((d.P + 255) / 256, d.NV);
    scatter_kernel<<<grid, 256, 0, s>>>(d, w);
    const int big_grid = d.NV * d.T < 148 * 7 ? d.NV * d.T : 148 * 7;      // 32 KB of keys per block: 7 blocks per SM
    // the few tiles above the warp-sort limit are sorted by whole blocks (a long tail of a handful of CTAs): on the
    // side stream, beside the warp-per-tile kernel that fills the GPU, instead of after it
    GaSide *g = ga_side();
    if (g) {
        cudaEventRecord(g->fork, s);
        cudaStreamWaitEvent(g->st, g->fork, 0);
        sort_tiles_kernel<<<big_grid, 256, 0, g->st>>>(d, w);
        cudaEventRecord(g->join, g->st);
        sort_tiles_warp_kernel<<<(d.NV * d.T + SORT_WARP_THREADS / 32 - 1) / (SORT_WARP_THREADS / 32), SORT_WARP_THREADS, 0, s>>>(d, w);
        cudaStreamWaitEvent(s, g->join, 0);
    } else {
        sort_tiles_warp_kernel<<<(d.NV * d.T + SORT_WARP_THREADS / 32 - 1) / (SORT_WARP_THREADS / 32), SORT_WARP_THREADS, 0, s>>>(d, w);
        sort_tiles_kernel<<<big_grid, 256, 0, s>>>(d, w);
    }
    return cudaGetLastError();
}
