// k2_sort.cu -- ordering of the window edges of the far-capable (line, depth) pairs.
//
// Part of the K2 PREPARATION pass of the far-field scheme (not of the per-pixel hot loop): the pairs whose window
// starts (ends) strictly inside a given tile are found by two binary searches in ONE sorted array instead of by
// scanning whole class lists per tile.  k_build_records appends a 64-bit key
//     kind (0 start / 1 end) | depth | lmin | edge pixel | line index
// only for edges that exist (inside the grid) and can matter to this context (inside its extended pixel range), so a
// nu shard sorts ~1/R of what a whole-grid run sorts.  The key is a total order: the sorted array -- and every
// summation order derived from it -- is independent of the (atomic) append order.  The radix sort itself is the one
// CUDA toolkit utility in the library (CUB, header-only, compiled here for sm_100a); it needs the number of keys on the
// host, which costs one 8-byte device->host copy and stream synchronisation per preparation pass.
#include <cub/device/device_radix_sort.cuh>

#include "sd_internal.h"

namespace {
// off[(kind * D + d) * SD_FAR_LEVELS + m] = first sorted key of (kind, d, lmin = m); one more entry closes the array
__global__ void k_edge_offsets(FarGeom fg, const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ count,
                               int D, int *__restrict__ off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 2 * D * SD_FAR_LEVELS;
    if (i > n) return;
    const long long M = (long long)count[0];
    long long lo = 0, hi = M;
    if (i == n) lo = M;
    else {
        const int m = i % SD_FAR_LEVELS, kd = i / SD_FAR_LEVELS, kind = kd / D, d = kd - kind * D;
        const unsigned long long X = sd_edge_key(fg, kind, d, m, 0, 0);
        while (lo < hi) {
            const long long mid = lo + ((hi - lo) >> 1);
            if (keys[mid] < X) lo = mid + 1; else hi = mid;
        }
    }
    off[i] = (int)lo;
}
// Range tables (FarGeom::fc_tab / edge_tab): one thread per (row, level-0 tile boundary), a binary search each.
__global__ void __launch_bounds__(256) k_range_tables(FarGeom fg, int D, int64_t L, const int *__restrict__ line_idx,
                                                      const int *__restrict__ cls_list, const int *__restrict__ cls_off,
                                                      const unsigned long long *__restrict__ keys, int *__restrict__ fc_tab,
                                                      int *__restrict__ edge_tab) {
    const int nb = fg.n_tiles[0] + 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_fc = (long long)D * SD_FAR_LEVELS * nb;
    if (i < n_fc) {
        const int t = (int)(i % nb), row = (int)(i / nb), m = row % SD_FAR_LEVELS, d = row / SD_FAR_LEVELS;
        // first entry of the class list whose line index (= centre pixel unless the line lies behind the last pixel,
        // which belongs to the last tile) is below 64 t; the indices descend along the list
        int lo = cls_off[d * (SD_NCLS + 1) + SD_FC0 + m], hi = cls_off[d * (SD_NCLS + 1) + SD_FC0 + m + 1];
        const int *list_d = cls_list + (size_t)d * L;
        const long long X = (t == nb - 1) ? (1LL << 40) : ((long long)t << SD_FAR_TILE0_SHIFT);
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((long long)line_idx[list_d[mid]] < X) hi = mid; else lo = mid + 1;
        }
        fc_tab[i] = lo;
    } else if (i < 3 * n_fc) {
        const long long e = i - n_fc;
        const int t = (int)(e % nb), row = (int)(e / nb);   // row = (kind * D + d) * SD_FAR_LEVELS + m
        int lo = fg.edge_off[row], hi = fg.edge_off[row + 1];
        const int m = row % SD_FAR_LEVELS, kd = row / SD_FAR_LEVELS, kind = kd / D, d = kd - kind * D;
        const unsigned long long X = sd_edge_key(fg, kind, d, m, (long long)t << SD_FAR_TILE0_SHIFT, 0);
        if (t == nb - 1) lo = hi;  // every listed edge lies inside the grid (the key field may not hold 64 t here)
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if (keys[mid] < X) lo = mid + 1; else hi = mid;
        }
        edge_tab[e] = lo;
    }
}
}  // namespace

int sd_sort_edges(sd_ctx *c) {
    FarGeom &fg = c->far_geom;
    if (!c->h_edge_count) SD_CUDA(c, cudaHostAlloc((void **)&c->h_edge_count, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    SD_CUDA(c, cudaMemcpyAsync(c->h_edge_count, c->edge_count.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(c->h_edge_count + 1, c->lev_info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long M = (long long)*c->h_edge_count;
    c->far_active = *reinterpret_cast<const int *>(c->h_edge_count + 1);  // usable hierarchy levels on this grid
    SD_CHECK(c, M >= 0 && M < 2147483647LL, SD_ERR_STATE, "window-edge list too long (%lld entries)", M);
    SD_TRY(sd_ensure(c, c->edge_keys, sizeof(unsigned long long) * (size_t)(M > 0 ? M : 1)));
    SD_TRY(sd_ensure(c, c->edge_off, sizeof(int) * (2 * c->D * SD_FAR_LEVELS + 1)));
    if (M > 0) {
        const int end_bit = 1 + fg.depth_bits + SD_FAR_LMIN_BITS + fg.pix_bits + fg.l_bits;
        const unsigned long long *in = c->edge_unsorted.as<unsigned long long>();
        unsigned long long *out = c->edge_keys.as<unsigned long long>();
        size_t bytes = 0;
        SD_CUDA(c, cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, (int)M, 0, end_bit, c->stream));
        SD_TRY(sd_ensure(c, c->edge_sort_tmp, bytes));
        SD_CUDA(c, cub::DeviceRadixSort::SortKeys(c->edge_sort_tmp.p, bytes, in, out, (int)M, 0, end_bit, c->stream));
        c->launches += 2 + (end_bit + 7) / 8;  // histogram + scan + one-sweep passes (library kernels, approximate)
    }
    fg.edge_keys = c->edge_keys.as<unsigned long long>();
    fg.edge_off = c->edge_off.as<int>();
    k_edge_offsets<<<(2 * c->D * SD_FAR_LEVELS + 1 + 127) / 128, 128, 0, c->stream>>>(fg, fg.edge_keys, c->edge_count.as<unsigned long long>(), c->D,
                                                                      c->edge_off.as<int>());
    SD_TRY(sd_launch_check(c, "k_edge_offsets"));
    const long long n_fc = (long long)c->D * SD_FAR_LEVELS * (fg.n_tiles[0] + 1);
    SD_TRY(sd_ensure(c, c->fc_tab, sizeof(int) * (size_t)n_fc));
    SD_TRY(sd_ensure(c, c->edge_tab, sizeof(int) * (size_t)(2 * n_fc)));
    fg.fc_tab = c->fc_tab.as<int>();
    fg.edge_tab = c->edge_tab.as<int>();
    k_range_tables<<<(unsigned)((3 * n_fc + 255) / 256), 256, 0, c->stream>>>(
        fg, c->D, c->L, c->line_idx.as<int>(), c->cls_list.as<int>(), c->cls_off.as<int>(), fg.edge_keys, c->fc_tab.as<int>(),
        c->edge_tab.as<int>());
    return sd_launch_check(c, "k_range_tables");
}
