// lj_sort.cu - device-wide exclusive scan and stable LSD radix sort of (cell key, atom slot) pairs.
//
// Replaces the std::vector<std::vector<unsigned>> subcells binning of build_neighbor_list
// (/root/reference/src/simulator/lennardjonessimulation.cpp:465-478): there atoms are appended to
// their cell in index order; here a stable sort by cell key yields the same per-cell order.
#include "lj_common.cuh"

namespace ljgpu {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems   = 4;
constexpr int kScanTile    = kScanThreads * kScanItems;   // 1024

// Exclusive scan of one tile per block; block total -> sums[blockIdx.x] (if sums != null).
__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        tsum += v[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = tsum;                                  // inclusive scan across the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t t = lane < kScanThreads / 32 ? warp_tot[lane] : 0u;
        uint32_t s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = s - t;     // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1 && sums) sums[blockIdx.x] = s;
    }
    __syncthreads();
    uint32_t run = warp_tot[w] + (inc - tsum);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ sums) {
    const uint32_t add = sums[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) data[base + k] += add;
}

constexpr int kSortThreads = 256;                 // == number of digits, relied upon below
constexpr int kSortItems   = 8;
constexpr int kSortTile    = kSortThreads * kSortItems;   // 2048 pairs per block
constexpr int kSortWarps   = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, int nblocks,
                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const size_t idx = base + (size_t)k * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];    // digit-major
}

// Stable scatter.  Warp w of a block owns 32*kSortItems consecutive pairs and walks them in kSortItems
// rounds of 32; within a round, lanes holding the same digit are ranked by lane id.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ hist_scanned, size_t n, int shift, int nblocks) {
    __shared__ uint32_t cnt[kSortWarps][256];
    for (int k = threadIdx.x; k < kSortWarps * 256; k += kSortThreads) (&cnt[0][0])[k] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)w * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t idx = wbase + (size_t)r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0u;
        val[r] = valid ? vals_in[idx] : 0u;
        const uint32_t d = valid ? ((key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        uint32_t old = 0;
        if (lane == leader && valid) { old = cnt[w][d]; cnt[w][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + below;
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive scan of the per-warp counts of digit d, seeded with the global base
        const int d = threadIdx.x;
        uint32_t run = hist_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) { const uint32_t t = cnt[ww][d]; cnt[ww][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t idx = wbase + (size_t)r * 32 + lane;
        if (idx < n) {
            const uint32_t pos = cnt[w][(key[r] >> shift) & 255u] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

// First CUDA error seen by the helpers of this file on the calling thread; the caller collects it with
// sort_take_error() and turns it into LJGPU_ECUDA (lj_sim.cu: CK).  After an error the helpers launch nothing more.
thread_local cudaError_t t_err = cudaSuccess;
#define SCK(x) do { const cudaError_t e_ = (x); if (e_ != cudaSuccess && t_err == cudaSuccess) t_err = e_; } while (0)

size_t scan_tmp_need(size_t n) {
    size_t need = 0, m = n;
    while (m > 1) { m = (m + kScanTile - 1) / kScanTile; need += m; }
    return need + 16;
}

void ensure(uint32_t*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    if (p) SCK(cudaFree(p));
    p = nullptr; cap = 0;
    size_t ncap = need + need / 4 + 64;
    SCK(cudaMalloc(&p, ncap * sizeof(uint32_t)));
    if (p) cap = ncap;
}

// Recursive scan; `tmp` has room for all levels.  Returns kernels launched.
int scan_rec(uint32_t* data, size_t n, uint32_t* tmp, uint32_t* total_out, cudaStream_t st) {
    const size_t nb = (n + kScanTile - 1) / kScanTile;
    int launches = 0;
    if (nb <= 1) {
        uint32_t* sums = total_out ? total_out : nullptr;
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(data, n, sums);
        SCK(cudaGetLastError());
        return 1;
    }
    scan_tile_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(data, n, tmp);
    SCK(cudaGetLastError());
    ++launches;
    launches += scan_rec(tmp, nb, tmp + nb, total_out, st);
    scan_add_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(data, n, tmp);
    SCK(cudaGetLastError());
    ++launches;
    return launches;
}

} // namespace

int exclusive_scan_u32(uint32_t* data, size_t n, uint32_t* total_out, SortScratch& scr, cudaStream_t st) {
    if (n == 0) {
        if (total_out) SCK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st));
        return 0;
    }
    ensure(scr.scan_tmp, scr.cap_scan, scan_tmp_need(n));
    if (t_err != cudaSuccess) return 0;
    return scan_rec(data, n, scr.scan_tmp, total_out, st);
}

int radix_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, int key_bits, SortScratch& scr,
                     cudaStream_t st) {
    if (n <= 1) return 0;
    if (scr.cap_items < n) {
        if (scr.keys_alt) SCK(cudaFree(scr.keys_alt));
        if (scr.vals_alt) SCK(cudaFree(scr.vals_alt));
        scr.keys_alt = scr.vals_alt = nullptr; scr.cap_items = 0;
        const size_t want = n + n / 8 + 1024;
        SCK(cudaMalloc(&scr.keys_alt, want * sizeof(uint32_t)));
        SCK(cudaMalloc(&scr.vals_alt, want * sizeof(uint32_t)));
        if (scr.keys_alt && scr.vals_alt) scr.cap_items = want;
    }
    const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
    ensure(scr.hist, scr.cap_hist, (size_t)256 * nblocks);
    if (t_err != cudaSuccess) return 0;
    int launches = 0;
    uint32_t *kin = keys, *vin = vals, *kout = scr.keys_alt, *vout = scr.vals_alt;
    const int passes = key_bits <= 0 ? 1 : (key_bits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, n, shift, nblocks, scr.hist);
        SCK(cudaGetLastError());
        ++launches;
        launches += exclusive_scan_u32(scr.hist, (size_t)256 * nblocks, nullptr, scr, st);
        radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, vin, kout, vout, scr.hist, n, shift, nblocks);
        SCK(cudaGetLastError());
        ++launches;
        if (t_err != cudaSuccess) return launches;
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (kin != keys) {
        SCK(cudaMemcpyAsync(keys, kin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        SCK(cudaMemcpyAsync(vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return launches;
}

int sort_take_error() {
    const cudaError_t e = t_err;
    t_err = cudaSuccess;
    return (int)e;
}

void sort_scratch_free(SortScratch& scr) {
    if (scr.keys_alt) cudaFree(scr.keys_alt);
    if (scr.vals_alt) cudaFree(scr.vals_alt);
    if (scr.hist) cudaFree(scr.hist);
    if (scr.scan_tmp) cudaFree(scr.scan_tmp);
    scr = SortScratch();
}

} // namespace ljgpu
