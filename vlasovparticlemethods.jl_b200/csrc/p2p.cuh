// One-shot all-reduce of a small fp64 vector over NVLink peer memory, executed INSIDE the single-CTA field
// kernels (fused compute + collective): every rank pushes its partial vector into a slot of every peer's
// mailbox with plain stores over NVLink, spins on the slots that its peers wrote into its own (local) mailbox,
// and sums the slots in rank order -- so all ranks obtain bitwise-identical sums, with no NCCL launch and no
// extra kernel between deposit and solve.
//
// The payload is 18 doubles (VP) to ~50 doubles (LB/CLB): pure latency.  Round 1 sent data, a system-scope
// fence and then a separate sequence flag (two NVLink round trips and the fence's wait for every remote write).
// This version is flag-in-data ("LL" protocol): every double travels as two 64-bit words, each carrying 32 bits
// of the value and a 32-bit tag of the collective's sequence number.  A 64-bit store is single-copy atomic, so
// a receiver that sees the tag in both words has the value -- no fence, no flag, one one-way trip.
// Mailboxes are double-buffered by the parity of the sequence number: a rank can only run one collective ahead
// of the slowest rank (it needs that rank's words to finish), so two buffers are enough.  A bounded spin (about
// 20 s of SM clocks; VPM_P2P_TIMEOUT_MS) turns a dead peer into an error instead of a hung GPU: the mailbox's
// error word is set (the synchronous steppers check it and return VPM_ERR_COMM) and the reduced vector is
// poisoned with NaN so that nothing downstream can pass for a result.
#pragma once
#include <cuda_runtime.h>

namespace vpm {

constexpr int kP2PMaxRanks = 8;
constexpr int kP2PCap = 2048;  // doubles per slot

struct P2PMailbox {
    unsigned long long ll[2][kP2PMaxRanks][2 * kP2PCap];   // low 32 bits: half of a double; high 32 bits: tag
    unsigned long long error;
};

struct P2PDev {
    P2PMailbox* mbox[kP2PMaxRanks];  // mbox[rank] is local, the others are IPC-mapped peer memory
    int nranks, rank;
    unsigned long long seq;          // 0 = inactive
    long long timeout_cycles;        // spin bound of one collective (SM clocks)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of a single-CTA kernel.  buf: global memory, count <= kP2PCap doubles, already
// complete and visible to the CTA (caller has synchronised).  On return buf holds the sum over ranks.
__device__ __forceinline__ void p2p_allreduce(const P2PDev& c, double* buf, int count)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int par = (int)(c.seq & 1ull);
    const unsigned long long tag = ((c.seq % 0xFFFFFFFFull) + 1ull) << 32;   // never 0: a zeroed mailbox matches nothing
    __shared__ int s_dead;
    if (tid == 0) s_dead = 0;
    __syncthreads();
    for (int i = tid; i < count; i += nt) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[i]);
        const unsigned long long w0 = (bits & 0xFFFFFFFFull) | tag, w1 = (bits >> 32) | tag;
        for (int r = 0; r < c.nranks; r++) {
            unsigned long long* dst = &c.mbox[r]->ll[par][c.rank][2 * i];
            st_relaxed_sys(dst, w0);
            st_relaxed_sys(dst + 1, w1);
        }
    }
    const P2PMailbox* mine = c.mbox[c.rank];
    for (int i = tid; i < count; i += nt) {
        // all 2 * nranks words of this element are requested back to back (independent loads: one round trip to where
        // the peers' stores land), and only then inspected; a rank whose words have not arrived yet is polled afterwards.
        // (Checking rank by rank, as the first version did, chains nranks dependent round trips: ~6 us at eight ranks.)
        unsigned long long w0[kP2PMaxRanks], w1[kP2PMaxRanks];
#pragma unroll
        for (int r = 0; r < kP2PMaxRanks; r++)
            if (r < c.nranks) {
                const unsigned long long* src = &mine->ll[par][r][2 * i];
                w0[r] = ld_relaxed_sys(src);
                w1[r] = ld_relaxed_sys(src + 1);
            }
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < kP2PMaxRanks; r++)   // fixed rank order: identical bits on every rank
            if (r < c.nranks) {
                if ((w0[r] & 0xFFFFFFFF00000000ull) != tag || (w1[r] & 0xFFFFFFFF00000000ull) != tag) {
                    const unsigned long long* src = &mine->ll[par][r][2 * i];
                    const long long t0 = clock64();
                    do {
                        w0[r] = ld_relaxed_sys(src);
                        w1[r] = ld_relaxed_sys(src + 1);
                        if (clock64() - t0 > c.timeout_cycles) {  // a peer died: flag it, poison the sum below
                            c.mbox[c.rank]->error = c.seq;
                            s_dead = 1;
                            break;
                        }
                    } while ((w0[r] & 0xFFFFFFFF00000000ull) != tag || (w1[r] & 0xFFFFFFFF00000000ull) != tag);
                }
                s += __longlong_as_double((long long)((w1[r] << 32) | (w0[r] & 0xFFFFFFFFull)));
            }
        buf[i] = s;
    }
    __syncthreads();
    if (s_dead)
        for (int i = tid; i < count; i += nt) buf[i] = __longlong_as_double(0x7ff8000000000000ll);
    __syncthreads();
}
#endif

}  // namespace vpm
