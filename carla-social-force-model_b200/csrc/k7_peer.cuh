// K7 -- peer-memory exchange over NVLink for the multi-GPU tick (one process per GPU, up to 8 ranks of one box).
//
// The row-partitioned tick has two exchange steps (multi-GPU): the integer reduce-scatter of the fixed-point
// pair-force accumulators and the all-gather of the staged rows.  With peer memory mapped (cudaIpc handles exchanged once
// at set-up) both are folded into the kernels on either side of them:
//   * k1_sym_finish PULLS the partial accumulators of its own rows from every rank's buffer (P2P loads) and adds them --
//     the reduce-scatter fused into the fixed-point -> float64 conversion;
//   * K3 PUSHES the rows it has just staged into every rank's gather buffer (P2P stores) -- the all-gather fused into
//     the integrate kernel.  The gather buffer is double-buffered so that a rank still repairing a poisoned row of tick k
//     (k1_sym_finish reads all staged rows) never sees tick k+1's rows arrive.
// What remains between the kernels is a flag barrier (k7_barrier): every rank stores its epoch into every peer's flag
// array (release, system scope) and waits until all peers' epochs have arrived (acquire, system scope).  The wait is
// bounded: after ~10 s of spinning (SFM_BARRIER_TIMEOUT_MS) the kernel records WHICH rank it gave up on instead of hanging
// the GPU, and later barriers return at once (sfm_peer_status reports the stalled ranks; the engine and bench.py turn it
// into an exception).  The exchange cannot be resumed after that: destroy the contexts and build new ones.
#pragma once

#include "sfm_common.cuh"

namespace sfm {

constexpr int MAX_PEERS = 8;

struct PeerPtrs {
    void* p[MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// flags[r] on every rank = the last epoch rank r has signalled.  One thread per peer.  error[0] collects the ranks a
// barrier gave up on (bit r), error[1] the epoch of the first such barrier; once set, later barriers return at once.
__global__ void k7_barrier(PeerPtrs flags, unsigned* own_flags, int world, int rank, unsigned epoch, unsigned* error,
                           long long timeout_cycles) {
    const int r = threadIdx.x;
    if (r >= world) return;
    if (*reinterpret_cast<volatile unsigned*>(error)) return;  // a peer was lost earlier: do not wait again
    __threadfence_system();                                   // everything this rank wrote before the barrier
    st_release_sys(reinterpret_cast<unsigned*>(flags.p[r]) + rank, epoch);
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(own_flags + r) - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) {                // default ~10 s at 1.9 GHz: peer r is gone
            if (atomicOr(error, 1u << r) == 0u) atomicCAS(error + 1, 0u, epoch);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace sfm
