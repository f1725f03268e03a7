// cov_peer.cu — the two exchange steps of the sharded trajectory objective as ONE small kernel each, over NVLink peer
// memory instead of a collective-library launch.
//
// The point-sharded objective needs, per step, the global per-pose normalisers (MAX over ranks of 2 W
// floats, minima negated) after pass A and the global accumulators (SUM over ranks of 22 W + 1 doubles) after pass B.
// Both vectors are tiny (2.5 KB / 56 KB at W = 320): the cost of a collective is its launch and synchronisation
// latency, not bandwidth.  Here every rank owns an exchange buffer that all peers can address (allocated by the host as
// symmetric memory; the library only sees the pointers): a block
//     1. draws the call's epoch from a local ticket counter (no host-side sequence number: the kernel is replayed
//        from CUDA graphs),
//     2. stores ITS slice of the local vector into slot [epoch parity][own rank] of EVERY peer's buffer (st.global over
//        NVLink), fences system-wide, and raises flag [epoch parity][slice][own rank] = epoch on every peer,
//     3. waits until its own buffer shows the epoch in the flags of all ranks for that slice,
//     4. reduces the world's slices in RANK ORDER (every rank computes bit-identical results) and writes them back in
//        place.
// Slices are independent, so there is no grid-wide barrier.  Slots alternate with the epoch's parity: a rank can run at
// most one call ahead of the slowest (it needs everybody's flags of call e + 1 before it can leave it, and those are
// raised only after their owners finished reading call e), so data of call e is never overwritten while in use.
// Ranks are separate processes on separate GPUs of one NVLink domain; every rank issues the same sequence of calls.
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kSlice = 1024;  // elements per block

struct PeerHeader {            // at the start of every rank's exchange region
    unsigned ticket;           // blocks that have entered, over all calls: epoch = ticket / nblocks + 1
    unsigned pad[63];
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// region layout (bytes): [PeerHeader 256][flags: 2 parities x nblocks x COV_MAX_PEERS u32][slots: 2 x world x n elements]
template <typename T>
struct PeerLayout {
    __host__ __device__ static size_t flags_off() { return sizeof(PeerHeader); }
    __host__ __device__ static size_t slots_off(int nblocks) {
        return (flags_off() + (size_t)2 * nblocks * COV_MAX_PEERS * sizeof(unsigned) + 255) & ~(size_t)255;
    }
    __host__ __device__ static size_t bytes(int64_t n, int world) {
        const int nblocks = (int)((n + kSlice - 1) / kSlice);
        return slots_off(nblocks) + (size_t)2 * world * (size_t)n * sizeof(T);
    }
};

// KIND: 0 = MAX (a NaN in any rank propagates, as the maxima of the single-GPU kernels), 1 = SUM,
//       2 = MIN over the first half of the vector, MAX over the second (the 2 W normalisers; a NaN minimum is dropped, as
//           the single-GPU kernels' integer-encoded minima do)
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
cov_peer_allreduce_kernel(T* __restrict__ data, int64_t n, cov_peers peers, size_t region_off) {
    __shared__ unsigned s_epoch;
    const int tid = threadIdx.x, nblocks = gridDim.x, world = peers.world, rank = peers.rank;
    char* mine = reinterpret_cast<char*>(peers.ptr[rank]) + region_off;
    if (tid == 0) s_epoch = atomicAdd(&reinterpret_cast<PeerHeader*>(mine)->ticket, 1u) / (unsigned)nblocks + 1u;
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int par = (int)(epoch & 1u);
    const int64_t i0 = (int64_t)blockIdx.x * kSlice;
    const int64_t i1 = i0 + kSlice < n ? i0 + kSlice : n;
    const size_t slots = PeerLayout<T>::slots_off(nblocks);
    // 2. push the slice to every peer's slot [par][rank]
    for (int r = 0; r < world; ++r) {
        T* dst = reinterpret_cast<T*>(reinterpret_cast<char*>(peers.ptr[r]) + region_off + slots) +
                 ((size_t)par * world + rank) * (size_t)n;
        for (int64_t i = i0 + tid; i < i1; i += blockDim.x) dst[i] = data[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(peers.ptr[tid]) + region_off +
                                                      PeerLayout<T>::flags_off()) +
                         ((size_t)par * nblocks + blockIdx.x) * COV_MAX_PEERS + rank;
        st_release_sys_u32(flag, epoch);
    }
    // 3. wait for the slice of every rank
    if (tid < world) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(mine + PeerLayout<T>::flags_off()) +
                               ((size_t)par * nblocks + blockIdx.x) * COV_MAX_PEERS + tid;
        while (ld_acquire_sys_u32(flag) != epoch) {
        }
    }
    __syncthreads();
    // 4. reduce in rank order, in place
    const T* slot0 = reinterpret_cast<const T*>(mine + slots) + (size_t)par * world * (size_t)n;
    for (int64_t i = i0 + tid; i < i1; i += blockDim.x) {
        T v = __ldcg(slot0 + i);
        for (int r = 1; r < world; ++r) {
            const T u = __ldcg(slot0 + (size_t)r * (size_t)n + i);
            if (KIND == 1) v += u;
            else if (KIND == 2 && i < n / 2) v = (u < v || v != v) ? u : v;
            else v = (u > v || u != u) ? u : v;
        }
        data[i] = v;
    }
}

}  // namespace

extern "C" size_t cov_peer_region_bytes(int kind, int64_t n, int world) {
    if (n < 1) n = 1;
    if (world < 1) world = 1;
    return kind == COV_PEER_SUM_F64 ? PeerLayout<double>::bytes(n, world) : PeerLayout<float>::bytes(n, world);
}

extern "C" int cov_peer_allreduce(int kind, void* data, int64_t n, const cov_peers* peers, size_t region_offset,
                                  void* stream) {
    if (!data || n <= 0 || !peers || peers->world < 1 || peers->world > COV_MAX_PEERS || peers->rank < 0 ||
        peers->rank >= peers->world || (kind != COV_PEER_MAX_F32 && kind != COV_PEER_SUM_F64 && kind != COV_PEER_MINMAX_F32)) {
        cov_set_error("cov_peer_allreduce: bad argument");
        return COV_ERR_ARG;
    }
    for (int r = 0; r < peers->world; ++r)
        if (!peers->ptr[r] || (((uintptr_t)peers->ptr[r] + region_offset) & 255)) {
            cov_set_error("cov_peer_allreduce: peer buffer %d missing or region not 256-byte aligned", r);
            return COV_ERR_ALIGN;
        }
    const int nblocks = (int)((n + kSlice - 1) / kSlice);
    cudaStream_t s = (cudaStream_t)stream;
    if (kind == COV_PEER_MAX_F32)
        cov_peer_allreduce_kernel<float, 0><<<nblocks, 256, 0, s>>>(reinterpret_cast<float*>(data), n, *peers, region_offset);
    else if (kind == COV_PEER_MINMAX_F32)
        cov_peer_allreduce_kernel<float, 2><<<nblocks, 256, 0, s>>>(reinterpret_cast<float*>(data), n, *peers, region_offset);
    else
        cov_peer_allreduce_kernel<double, 1><<<nblocks, 256, 0, s>>>(reinterpret_cast<double*>(data), n, *peers, region_offset);
    return cov_check_launch("cov_peer_allreduce");
}
