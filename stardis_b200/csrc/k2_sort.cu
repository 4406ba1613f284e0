// k2_sort.cu -- ordering of the far-capable (line, depth) pairs by the pixel position of their window edges.
//
// Part of the K2 PREPARATION pass of the far-field scheme (not of the per-pixel hot loop): the pairs whose window
// starts (ends) strictly inside a given tile are found by two binary searches in these sorted arrays instead of by
// scanning whole class lists per tile.  Keys are 32-bit (depth << key_shift | edge pixel), values the line index; a stable
// LSD radix sort keeps equal keys in line order, so everything downstream stays deterministic.  This is the one place
// where a CUDA toolkit utility (CUB, header-only, compiled here for sm_100a) is used instead of a hand-written kernel.
#include <cub/device/device_radix_sort.cuh>

#include "sd_internal.h"

// which == 1: window ENDS   (unsorted keys staged in edge_keys[0])  -> edge_keys[1], edge_l[1]
// which == 0: window STARTS (unsorted keys in edge_tmp_keys)        -> edge_keys[0], edge_l[0]
int sd_sort_edges(sd_ctx *c, int which, int64_t n) {
    const unsigned *keys_in = which ? c->edge_keys[0].as<unsigned>() : c->edge_tmp_keys.as<unsigned>();
    unsigned *keys_out = c->edge_keys[which].as<unsigned>();
    const int *vals_in = c->edge_tmp_l.as<int>();
    int *vals_out = c->edge_l[which].as<int>();
    int depth_bits = 1;
    while ((1 << depth_bits) < c->D) depth_bits++;
    const int end_bit = c->far_geom.key_shift + depth_bits;
    size_t bytes = 0;
    SD_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, c->stream));
    SD_TRY(sd_ensure(c, c->edge_sort_tmp, bytes));
    SD_CUDA(c, cub::DeviceRadixSort::SortPairs(c->edge_sort_tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit,
                                               c->stream));
    c->launches += 6;  // histogram + one-sweep passes of the radix sort (approximate; library kernels)
    return SD_OK;
}
