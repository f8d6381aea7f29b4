// One-shot all-reduce of a small fp64 vector over NVLink peer memory, executed INSIDE the single-CTA field
// kernels (fused compute + collective): every rank pushes its partial vector into a slot of every peer's
// mailbox with plain stores over NVLink, publishes a sequence flag with release semantics, spins on the
// flags that its peers wrote into its own (local) mailbox, and sums the slots in rank order — so all ranks
// obtain bitwise-identical sums, with no NCCL launch and no extra kernel between deposit and solve.
//
// The payload is 18 doubles (VP) to ~50 doubles (LB/CLB): pure latency, which is why it is a handful of
// stores and a flag instead of a ring.  Mailboxes are double-buffered by the parity of the sequence number:
// a rank can only run one collective ahead of the slowest rank (it needs that rank's flag to finish), so two
// buffers are enough.  A bounded spin (about 20 s of SM clocks; VPM_P2P_TIMEOUT_MS) turns a dead peer into an
// error instead of a hung GPU: the mailbox's error word is set (the synchronous steppers check it and return
// VPM_ERR_COMM) and the reduced vector is poisoned with NaN so that nothing downstream can pass for a result.
#pragma once
#include <cuda_runtime.h>

namespace vpm {

constexpr int kP2PMaxRanks = 8;
constexpr int kP2PCap = 2048;  // doubles per slot

struct P2PMailbox {
    double data[2][kP2PMaxRanks][kP2PCap];
    unsigned long long flag[2][kP2PMaxRanks];
    unsigned long long error;
};

struct P2PDev {
    P2PMailbox* mbox[kP2PMaxRanks];  // mbox[rank] is local, the others are IPC-mapped peer memory
    int nranks, rank;
    unsigned long long seq;          // 0 = inactive
    long long timeout_cycles;        // spin bound of one collective (SM clocks)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of a single-CTA kernel.  buf: global memory, count <= kP2PCap doubles, already
// complete and visible to the CTA (caller has synchronised).  On return buf holds the sum over ranks.
__device__ __forceinline__ void p2p_allreduce(const P2PDev& c, double* buf, int count)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int par = (int)(c.seq & 1ull);
    __shared__ int s_dead;
    if (tid == 0) s_dead = 0;
    for (int r = 0; r < c.nranks; r++) {
        double* dst = c.mbox[r]->data[par][c.rank];
        for (int i = tid; i < count; i += nt) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < c.nranks) st_release_sys(&c.mbox[tid]->flag[par][c.rank], c.seq);
    if (tid < c.nranks) {
        const unsigned long long* f = &c.mbox[c.rank]->flag[par][tid];
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < c.seq) {
            if (clock64() - t0 > c.timeout_cycles) {  // a peer died: flag it, poison the sum below
                c.mbox[c.rank]->error = c.seq;
                s_dead = 1;
                break;
            }
        }
    }
    __syncthreads();
    const P2PMailbox* mine = c.mbox[c.rank];
    for (int i = tid; i < count; i += nt) {
        double s = 0.0;
        for (int r = 0; r < c.nranks; r++) s += __ldcv(&mine->data[par][r][i]);  // fixed rank order
        buf[i] = s_dead ? __longlong_as_double(0x7ff8000000000000ll) : s;
    }
    __syncthreads();
}
#endif

}  // namespace vpm
