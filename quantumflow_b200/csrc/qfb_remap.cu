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
    // slices (qfb_remap_swap_slice): only the amplitudes whose offset inside the pair's run has the bits selpos[] equal
    // to selval are swapped; n then counts the amplitudes of the slice and their offsets are expanded on the fly
    int nsel;
    int selpos[4];          // ascending
    uint64_t selval;
};

__device__ __forceinline__ uint64_t slice_offset(const RemapParams &p, uint64_t i) {
    for (int t = 0; t < p.nsel; ++t) {
        const int b = p.selpos[t];
        i = ((i >> b) << (b + 1)) | (i & ((1ull << b) - 1));
    }
    return i | p.selval;
}

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
                const uint64_t o = slice_offset(p, i);
                va[u] = ldg_stream(a + o);
                vb[u] = ldg_stream(b + o);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x;
            if (i < n) {
                const uint64_t o = slice_offset(p, i);
                stg_stream(a + o, vb[u]);
                stg_stream(b + o, va[u]);
            }
        }
    }
}

}  // namespace qfb

using namespace qfb;

namespace qfb {

// Barrier across the GPUs of one box through peer memory: every rank writes `epoch` into its slot of every peer's flag
// array and waits until its own array holds `epoch` in every slot. One CTA of `world` threads; stream ordered, so the
// kernels queued before it on every rank's stream are complete (and their peer writes visible) when it returns.
// The wait is bounded (two minutes): on expiry *error is set and the kernel returns, so a lost peer cannot hang the GPU.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct PeerTable {
    uint32_t *p[32];
};

__global__ void peer_barrier_kernel(volatile uint32_t *mine, PeerTable peers, int world, int rank, uint32_t epoch,
                                    uint32_t *error) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    volatile uint32_t *dst = (volatile uint32_t *)peers.p[r];
    dst[rank] = epoch;
    __threadfence_system();
    const unsigned long long t0 = global_ns();
    while ((int32_t)(mine[r] - epoch) < 0) {
        __nanosleep(200);
        if (global_ns() - t0 > 120ull * 1000000000ull) {      // two minutes: ranks may arrive seconds apart (e.g. after
            *error = epoch;                                   // each has staged a 128 GiB shard over PCIe)
            return;
        }
    }
    __threadfence_system();
}

}  // namespace qfb

extern "C" {

static int remap_swap_impl(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                           int nsel, const int *selpos, uint64_t selval, int ctas_per_sm, void *stream,
                           const char *who) {
    QFB_CHECK_ARG(npairs >= 0 && npairs <= REMAP_MAX_PAIRS, "%s: npairs=%d out of range", who, npairs);
    if (npairs == 0) return QFB_OK;
    QFB_CHECK_ARG(local_blocks && remote_blocks && nelems, "%s: null pointer", who);
    QFB_CHECK_ARG(nsel >= 0 && nsel <= 4 && (nsel == 0 || selpos), "%s: 0..4 selector bits", who);
    RemapParams p;
    p.npairs = npairs;
    p.nsel = nsel;
    p.selval = selval;
    uint64_t selmask = 0;
    for (int t = 0; t < nsel; ++t) {
        QFB_CHECK_ARG(selpos[t] >= 0 && selpos[t] < 62 && (t == 0 || selpos[t] > selpos[t - 1]),
                      "%s: selector bits must ascend", who);
        p.selpos[t] = selpos[t];
        selmask |= 1ull << selpos[t];
    }
    QFB_CHECK_ARG((selval & ~selmask) == 0, "%s: selector value outside the selector bits", who);
    const uint64_t per_block = (uint64_t)UNROLL * 256;
    p.start[0] = 0;
    for (int i = 0; i < npairs; ++i) {
        QFB_CHECK_ARG(local_blocks[i] && remote_blocks[i], "%s: null block", who);
        QFB_CHECK_ARG(((uintptr_t)local_blocks[i] % 16) == 0 && ((uintptr_t)remote_blocks[i] % 16) == 0,
                      "%s: blocks must be 16-byte aligned", who);
        QFB_CHECK_ARG(nsel == 0 || (nelems[i] % (2ull << selpos[nsel - 1])) == 0,
                      "%s: run length must be a multiple of twice the top selector bit", who);
        p.local[i] = (c128 *)local_blocks[i];
        p.remote[i] = (c128 *)remote_blocks[i];
        p.n[i] = nelems[i] >> nsel;
        p.start[i + 1] = p.start[i] + (p.n[i] + per_block - 1) / per_block;
    }
    const uint64_t chunks = p.start[npairs];
    if (chunks == 0) return QFB_OK;
    // enough resident CTAs to keep ~2 MB in flight per direction (NVLink round trip of a few microseconds); a slice
    // that runs beside a sweep asks for fewer so that the sweep's CTAs keep most of every SM
    if (ctas_per_sm <= 0) ctas_per_sm = 8;
    const uint64_t cap = (uint64_t)sm_count_cached() * (uint64_t)std::min(ctas_per_sm, 8);
    remap_swap_kernel<<<(unsigned)std::min<uint64_t>(chunks, cap), 256, 0, (cudaStream_t)stream>>>(p);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_remap_swap(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                   void *stream) {
    return remap_swap_impl(npairs, local_blocks, remote_blocks, nelems, 0, nullptr, 0, 0, stream, "qfb_remap_swap");
}

int qfb_remap_swap_slice(int npairs, void *const *local_blocks, void *const *remote_blocks, const uint64_t *nelems,
                         int nsel, const int *selpos, uint64_t selval, int ctas_per_sm, void *stream) {
    return remap_swap_impl(npairs, local_blocks, remote_blocks, nelems, nsel, selpos, selval, ctas_per_sm, stream,
                           "qfb_remap_swap_slice");
}

int qfb_peer_barrier(void *flags_local, void *const *flags_of_ranks, int world, int rank, uint32_t epoch,
                     void *error_dev, void *stream) {
    QFB_CHECK_ARG(flags_local && flags_of_ranks && error_dev, "qfb_peer_barrier: null pointer");
    QFB_CHECK_ARG(world >= 1 && world <= 32 && rank >= 0 && rank < world, "qfb_peer_barrier: world=%d rank=%d", world,
                  rank);
    PeerTable t;
    for (int r = 0; r < 32; ++r) t.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        QFB_CHECK_ARG(flags_of_ranks[r], "qfb_peer_barrier: null flag array of rank %d", r);
        t.p[r] = (uint32_t *)flags_of_ranks[r];
    }
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((volatile uint32_t *)flags_local, t, world, rank, epoch,
                                                           (uint32_t *)error_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

}  // extern "C"
