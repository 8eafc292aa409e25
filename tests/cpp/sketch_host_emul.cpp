// TEST INFRASTRUCTURE ONLY.  Runs the sketch kernels of nanospring_b200/csrc/sketch_kernels.cuh on the
// host (cuda_host_shim.h; NSMH_HOST_EMUL replaces the bulk copy + mbarrier of the filter kernel by a
// memcpy and a warp barrier) in the order sketch_reads (sketch.cu) launches them, the CUB prefix sum
// replaced by a loop.  tests/test_sketch_emul.py compares the sketch matrix with the oracle; a logic
// check for the container without a GPU, never a product path.
#define NSMH_HOST_EMUL 1
#include "cuda_host_shim.h"

#include <cstdlib>
#include <cstring>

#include "../../nanospring_b200/csrc/sketch_kernels.cuh"
#include "../../nanospring_b200/csrc/sketch_tables.h"

using namespace nsmh;

static int sketch_emul_deferred = 0;
extern "C" {
void sketch_emul_set_deferred(int on) { sketch_emul_deferred = on; }


// W: packed stream followed by kPackPadWords zero words.  mode 0 = filter kernel + exact fix-up,
// 1 = brute force.  sk [n_reads][n] out; *fixups = entries the fix-up pass recomputed.  Returns 0,
// or -1 when the configuration does not fit the filter kernel (n > 255).
int sketch_emul_run(const uint32_t *W, const uint64_t *off, uint32_t n_reads, uint32_t k, uint32_t n,
                    const uint64_t *rnd, int mode, int lambda_log2, uint32_t tile_words, unsigned grid, uint64_t *sk,
                    unsigned long long *fixups) {
    *fixups = 0;
    if (n_reads == 0) return 0;
    if (mode != 1 && n > 255) return -1;
    const FilterTables ft = make_filter_tables(rnd, n, k);
    unsigned long long counters[8] = {0};
    SketchArgs a;
    a.off = off;
    a.W = W;
    a.sk = sk;
    a.rnd = rnd;
    a.ftab_first = ft.first.data();
    a.ftab_next = ft.next.data();
    a.ftab_hit3 = ft.hit3.data();
    a.counters = counters;
    a.n_reads = n_reads;
    a.k = k;
    a.n = n;
    a.lambda_log2 = lambda_log2;
    a.tile_words = (tile_words + 63) & ~63u;
    if (a.tile_words > 4096) a.tile_words = 4096;
    const uint64_t num_words = (off[n_reads] + 15) / 16;
    const size_t max_tiles = (size_t)((num_words + 4 * (uint64_t)n_reads) / a.tile_words) + n_reads + 1;
    std::vector<uint32_t> cnt(n_reads + 1, 0), ts(n_reads + 1, 0), tile_read(max_tiles + 2, 0xFFFFFFFFu);
    unsigned int queue[2] = {0, 0};
    a.tile_start = ts.data();
    a.tile_read = tile_read.data();
    a.tile_queue = queue;
    emu_launch(grid, 64, [&] { sketch_init_kernel(a, cnt.data()); });
    for (uint32_t i = 0; i < n_reads; ++i) ts[i + 1] = ts[i] + cnt[i];
    if (ts[n_reads] > max_tiles) return -2;                                  // the bound sketch_reads allocates by
    emu_launch(grid, 64, [&] { sketch_tile_map_kernel(ts.data(), n_reads, tile_read.data()); });
    if (mode != 1) {
        const FilterSmem L(n, a.tile_words);
        const size_t bytes = L.tab_bytes + L.warp_bytes;                     // one warp per block
        uint8_t *smem = static_cast<uint8_t *>(aligned_alloc(16, (bytes + 15) & ~(size_t)15));
        memset(smem, 0xA5, bytes);
        emu_launch(grid, 32, [&] { sketch_filter_kernel(a, smem); });
        free(smem);
        std::vector<uint32_t> miss((size_t)n_reads * n + 1, 0);
        emu_launch(grid, 64, [&] { sketch_missing_kernel(a, miss.data(), queue + 1); });
        if (sketch_emul_deferred) {
            // nsmh_sketch_build: the values go to a buffer of their own; table_insert_list_kernel stores them later
            std::vector<uint64_t> vals((size_t)n_reads * n + 1, 0x1111111111111111ULL);
            emu_launch(grid, 64, [&] { sketch_fixup_kernel<4>(a, miss.data(), queue + 1, vals.data()); });
            for (unsigned int e = 0; e < queue[1]; ++e) {
                if (a.sk[miss[e]] != ~0ULL) return 7;                       // the matrix must not have been touched
                a.sk[miss[e]] = vals[e];
            }
        } else {
            emu_launch(grid, 64, [&] { sketch_fixup_kernel<4>(a, miss.data(), queue + 1, nullptr); });
        }
        *fixups = counters[0];
    } else {
        emu_launch(grid, 64, [&] { sketch_brute_kernel<8>(a); });
    }
    return 0;
}

}  // extern "C"
