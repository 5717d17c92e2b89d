// qfb_remap.cu -- the exchange half of a qubit remap of a sharded state (SURVEY 8e / quantumflow has no
// distributed path): k rank bits are exchanged with the top-k local bits, i.e. block j of this rank's shard
// trades places with block `mine` of the peer whose selected rank bits equal j.
//
// One kernel per remap and rank, over peer memory (NVLink 5 / NVSwitch, the peers' shards are mapped through
// CUDA IPC by the host side): for every partner block the rank swaps ITS HALF of the pair's data in place --
// a[i] <-> b[i] with a in the local shard and b in the peer's shard, 16-byte accesses, both read and written once.
// The lower rank of a pair takes the first half of the block, the higher rank the second half, so every pair's
// traffic is split evenly and both directions of every link carry data at the same time (remote reads come
// back while remote writes go out). No staging buffer, no pack / unpack pass, no NCCL call: the bytes that
// cross NVLink are exactly the algorithmic ones, (1 - 2^-k) of the shard per direction.
// Ordering against the sweeps before and after is the host's business (a stream-ordered barrier on either side).
#include <algorithm>
#include "qfb_common.cuh"

namespace qfb {

constexpr int REMAP_MAX_PAIRS = 16;

struct RemapParams {
    c128 *local[REMAP_MAX_PAIRS];
    c128 *remote[REMAP_MAX_PAIRS];
    uint64_t n[REMAP_MAX_PAIRS];       // amplitudes this rank swaps for the pair
    uint64_t start[REMAP_MAX_PAIRS + 1];   // prefix sums of n in units of UNROLL * blockDim chunks
    int npairs;
};

constexpr int UNROLL = 4;

__global__ void __launch_bounds__(256) remap_swap_kernel(RemapParams p) {
    const uint64_t per_block = (uint64_t)UNROLL * blockDim.x;
    const uint64_t total_chunks = p.start[p.npairs];
    for (uint64_t chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int pair = 0;
        while (pair + 1 < p.npairs && chunk >= p.start[pair + 1]) ++pair;
        const uint64_t base = (chunk - p.start[pair]) * per_block + threadIdx.x;
        c128 *a = p.local[pair], *b = p.remote[pair];
        const uint64_t n = p.n[pair];
        c128 va[UNROLL], vb[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < n) {
                va[u] = ldg_stream(a + i);
                vb[u] = ldg_stream(b + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < n) {
                stg_stream(a + i, vb[u]);
                stg_stream(b + i, va[u]);
            }
        }
    }
}

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_remap_swap(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                   void *stream) {
    QFB_CHECK_ARG(npairs >= 0 && npairs <= REMAP_MAX_PAIRS, "qfb_remap_swap: npairs=%d out of range", npairs);
    if (npairs == 0) return QFB_OK;
    QFB_CHECK_ARG(local_blocks && remote_blocks && nelems, "qfb_remap_swap: null pointer");
    RemapParams p;
    p.npairs = npairs;
    const uint64_t per_block = (uint64_t)UNROLL * 256;
    p.start[0] = 0;
    for (int i = 0; i < npairs; ++i) {
        QFB_CHECK_ARG(local_blocks[i] && remote_blocks[i], "qfb_remap_swap: null block");
        QFB_CHECK_ARG(((uintptr_t)local_blocks[i] % 16) == 0 && ((uintptr_t)remote_blocks[i] % 16) == 0,
                      "qfb_remap_swap: blocks must be 16-byte aligned");
        p.local[i] = (c128 *)local_blocks[i];
        p.remote[i] = (c128 *)remote_blocks[i];
        p.n[i] = nelems[i];
        p.start[i + 1] = p.start[i] + (nelems[i] + per_block - 1) / per_block;
    }
    const uint64_t chunks = p.start[npairs];
    if (chunks == 0) return QFB_OK;
    // enough resident CTAs to keep ~2 MB in flight per direction (NVLink round trip of a few microseconds)
    const uint64_t cap = (uint64_t)sm_count_cached() * 8;
    remap_swap_kernel<<<(unsigned)std::min<uint64_t>(chunks, cap), 256, 0, (cudaStream_t)stream>>>(p);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

}  // extern "C"
