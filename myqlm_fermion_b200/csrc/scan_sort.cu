// Device-wide exclusive scan and a stable LSD radix sort for packed Pauli records.
// Hand-written for sm_100a; both are HBM-streaming integer kernels (coalesced 128-bit-friendly
// loads, shared-memory histograms, warp match/ballot ranking), used by the term canonicaliser.
#include "common.cuh"

namespace qfe {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
        s = warp_inclusive_scan(s);
        warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t base = warp ? warp_sums[warp - 1] : 0u;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums, int64_t n) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (j < n) s += in[j];
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// each thread owns SCAN_ITEMS consecutive elements so the scan order is the array order
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_apply(const uint32_t* in, uint32_t* out,  // may alias (in-place scan)
                                                                  const uint32_t* __restrict__ tile_offsets, int64_t n,
                                                                  uint32_t* __restrict__ d_total) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
    uint32_t off = (tile_offsets ? tile_offsets[blockIdx.x] : 0u) + ex;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = off;
        off += v[i];
    }
    if (d_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *d_total = off;
}

}  // namespace

static int scan_rec(qfe_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, uint32_t* d_total) {
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles <= 1) {
        scan_tile_apply<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, nullptr, n, d_total);
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
        return QFE_OK;
    }
    uint32_t* sums = scratch;
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, sums, n);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    // scan the tile sums in place (next level uses the scratch after them)
    QFE_TRY(scan_rec(ctx, sums, sums, tiles, scratch + ((tiles + 3) & ~int64_t(3)), nullptr));
    scan_tile_apply<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, sums, n, d_total);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

int exclusive_scan_u32(qfe_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, DevBuf& scratch, uint32_t* d_total) {
    if (n <= 0) {
        if (d_total) QFE_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), ctx->stream));
        return QFE_OK;
    }
    // geometric series of tile-sum levels: n/2048 + n/2048^2 + ... (+ alignment slack)
    size_t need = 0;
    for (int64_t m = n; m > 1;) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        need += (size_t)((m + 3) & ~int64_t(3));
        if (m <= 1) break;
    }
    QFE_TRY(scratch.reserve((need + 16) * sizeof(uint32_t)));
    return scan_rec(ctx, in, out, n, scratch.as<uint32_t>(), d_total);
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // elements per block
constexpr int RS_WARP_SPAN = RS_TILE / RS_WARPS;

struct SortKeys {
    const uint64_t* x;
    const uint64_t* z;
    const uint32_t* owner;
};

// which: 0 = z, 1 = x, 2 = owner
__device__ __forceinline__ uint32_t digit_of(const SortKeys& k, int64_t i, int which, int shift) {
    if (which == 0) return (uint32_t)(k.z[i] >> shift) & 255u;
    if (which == 1) return (uint32_t)(k.x[i] >> shift) & 255u;
    return (k.owner[i] >> shift) & 255u;
}

__device__ __forceinline__ void warp_histogram(const SortKeys& k, int64_t n, int which, int shift, uint32_t* wh /* [256] of this warp */) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int64_t start = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WARP_SPAN;
    for (int c = 0; c < RS_WARP_SPAN; c += 32) {
        int64_t i = start + c + lane;
        if (i < n) atomicAdd(&wh[digit_of(k, i, which, shift)], 1u);
    }
}

// hist layout: [digit][block] so that one exclusive scan yields global bucket bases
__global__ void __launch_bounds__(RS_THREADS) radix_hist(SortKeys k, int64_t n, int which, int shift, uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t wh[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0u;
    __syncthreads();
    warp_histogram(k, n, which, shift, wh[threadIdx.x >> 5]);
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) s += wh[w][threadIdx.x];
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = s;
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter(SortKeys k, const uint32_t* __restrict__ idx_in, int64_t n, int which, int shift,
                                                             const uint32_t* __restrict__ hist_scanned, int nblocks, uint64_t* __restrict__ x_out,
                                                             uint64_t* __restrict__ z_out, uint32_t* __restrict__ owner_out,
                                                             uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t wh[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    warp_histogram(k, n, which, shift, wh[warp]);
    __syncthreads();
    {
        // thread d turns the per-warp counts of digit d into per-warp global start offsets
        uint32_t run = hist_scanned[(int64_t)threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t cnt = wh[w][threadIdx.x];
            wh[w][threadIdx.x] = run;
            run += cnt;
        }
    }
    __syncthreads();
    uint32_t* off = wh[warp];
    const int64_t start = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WARP_SPAN;
    for (int c = 0; c < RS_WARP_SPAN; c += 32) {
        const int64_t i = start + c + lane;
        const bool live = i < n;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const uint32_t d = digit_of(k, i, which, shift);
            const unsigned peers = __match_any_sync(active, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            const uint32_t pos = off[d] + rank;
            __syncwarp(active);
            if (rank == 0) off[d] += __popc(peers);
            x_out[pos] = k.x[i];
            z_out[pos] = k.z[i];
            if (owner_out) owner_out[pos] = k.owner[i];
            idx_out[pos] = idx_in[i];
        }
        __syncwarp();
    }
}

}  // namespace

int radix_sort_terms(qfe_ctx* ctx, uint64_t* x, uint64_t* z, uint32_t* owner, uint32_t* idx, int64_t n, int n_bits_x, int n_bits_z,
                     int n_bits_owner, DevBuf& tmp) {
    if (n <= 1) return QFE_OK;
    QFE_REQUIRE(n < (int64_t(1) << 32), QFE_ERR_UNSUPPORTED, "radix_sort_terms: more than 2^32 records");
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    const size_t hist_elems = (size_t)256 * nblocks;
    // tmp layout: x2 | z2 | owner2 | idx2 | hist | scan scratch (separate DevBuf)
    const size_t off_x = 0, off_z = off_x + sizeof(uint64_t) * n, off_o = off_z + sizeof(uint64_t) * n;
    const size_t off_i = off_o + sizeof(uint32_t) * ((n + 3) & ~int64_t(3));
    const size_t off_h = off_i + sizeof(uint32_t) * ((n + 3) & ~int64_t(3));
    QFE_TRY(tmp.reserve(off_h + hist_elems * sizeof(uint32_t) + 64));
    char* base = tmp.as<char>();
    uint64_t* xb[2] = {x, reinterpret_cast<uint64_t*>(base + off_x)};
    uint64_t* zb[2] = {z, reinterpret_cast<uint64_t*>(base + off_z)};
    uint32_t* ob[2] = {owner, owner ? reinterpret_cast<uint32_t*>(base + off_o) : nullptr};
    uint32_t* ib[2] = {idx, reinterpret_cast<uint32_t*>(base + off_i)};
    uint32_t* hist = reinterpret_cast<uint32_t*>(base + off_h);
    DevBuf scan_scratch;

    struct Pass {
        int which, shift;
    };
    std::vector<Pass> passes;
    for (int s = 0; s < n_bits_z; s += 8) passes.push_back({0, s});
    for (int s = 0; s < n_bits_x; s += 8) passes.push_back({1, s});
    if (owner)
        for (int s = 0; s < n_bits_owner; s += 8) passes.push_back({2, s});

    int cur = 0;
    for (const Pass& p : passes) {
        SortKeys k{xb[cur], zb[cur], ob[cur]};
        radix_hist<<<nblocks, RS_THREADS, 0, ctx->stream>>>(k, n, p.which, p.shift, hist, nblocks);
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
        QFE_TRY(exclusive_scan_u32(ctx, hist, hist, (int64_t)hist_elems, scan_scratch, nullptr));
        radix_scatter<<<nblocks, RS_THREADS, 0, ctx->stream>>>(k, ib[cur], n, p.which, p.shift, hist, nblocks, xb[cur ^ 1], zb[cur ^ 1],
                                                                ob[cur ^ 1], ib[cur ^ 1]);
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    if (cur == 1) {
        QFE_CUDA(cudaMemcpyAsync(x, xb[1], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        QFE_CUDA(cudaMemcpyAsync(z, zb[1], sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        if (owner) QFE_CUDA(cudaMemcpyAsync(owner, ob[1], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        QFE_CUDA(cudaMemcpyAsync(idx, ib[1], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // scan_scratch is freed on return; make sure nothing queued still uses it
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return QFE_OK;
}

}  // namespace qfe
