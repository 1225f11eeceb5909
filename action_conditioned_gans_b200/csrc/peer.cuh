// Device side of the NVLink peer-memory exchange (protocol: peer.cu header).  Shared by the stand-alone exchange kernel
// and by the kernels whose LAST CTA exchanges its result itself (conv epilogue moments: no extra launch per SyncBN layer).
#pragma once
#include "common.cuh"

namespace acg {

struct PeerExchange {
    unsigned char* mbox[ACG_MAX_PEERS];     // every rank's mailbox segment as mapped HERE; [rank] is the own one
    unsigned long long* epoch;              // device counter of this slot on this rank (advanced by the exchange)
    long long slot_off;                     // byte offset of the slot inside every mailbox (256-byte aligned)
    long long timeout_ns;
    int rank, world, cap;                   // world == 0: no exchange
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// vec[0:n] <- sum over ranks, in rank order (identical bits on every rank).  Called by ALL threads of ONE CTA; vec must be
// complete and visible to this CTA.
//
// Low-latency protocol (one NVLink one-way latency per exchange, no fences, no separate flag): every fp64 value travels
// as ONE 16-byte cell {low word, tag, high word, tag} with tag = the exchange's epoch -- two 8-byte halves that each
// carry their own tag, so the 8-byte store atomicity of the fabric is all that is needed.  A receiver polls the cell in
// its OWN mailbox until both tags show the epoch it waits for.  Cells are double-buffered by epoch parity: a rank can be
// at most one exchange ahead of a peer on the same slot (it needs the peer's cells of the current epoch to get further),
// so the previous tag in a cell is always epoch - 2 and a cell is never overwritten while a peer still has to read it.
// Slot layout at slot_off: 128 B header (unused) | cells[2 parities][world][cap], 16 B each.
// AFTER_WAIT runs once the first value has arrived (the stand-alone kernel triggers its PDL dependents there).
template <typename AfterWait>
__device__ __forceinline__ void peer_exchange(const PeerExchange& x, double* vec, int n, int tid, int nthreads,
                                              AfterWait after_wait) {
    const unsigned long long epoch = *x.epoch + 1ull;
    const unsigned int tag = (unsigned int)epoch;
    const size_t cells_off = (size_t)x.slot_off + 128 + (size_t)(epoch & 1ull) * x.world * x.cap * 16;
    // push: one 16-byte store per value and peer (the own mailbox included)
    for (int i = tid; i < n; i += nthreads) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(__ldcg(vec + i));
        const uint4 cell = make_uint4((unsigned int)bits, tag, (unsigned int)(bits >> 32), tag);
        for (int p = 0; p < x.world; ++p) {
            uint4* dst = reinterpret_cast<uint4*>(x.mbox[p] + cells_off) + (size_t)x.rank * x.cap + i;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(cell.x), "r"(cell.y),
                         "r"(cell.z), "r"(cell.w) : "memory");
        }
    }
    // pull: poll the own mailbox, add in rank order
    const uint4* mine = reinterpret_cast<const uint4*>(x.mbox[x.rank] + cells_off);
    bool first = true;
    for (int i0 = 0; i0 < n; i0 += nthreads) {
        const int i = i0 + tid;
        double s = 0.0;
        if (i < n) {
            for (int r = 0; r < x.world; ++r) {
                const uint4* src = mine + (size_t)r * x.cap + i;
                uint4 c;
                unsigned int spins = 0;
                unsigned long long t0 = 0;
                for (;;) {
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(src) : "memory");
                    if (c.y == tag && c.w == tag) break;
                    if ((++spins & 1023u) == 0u) {
                        const unsigned long long now = global_ns();
                        if (t0 == 0) t0 = now;
                        if ((long long)(now - t0) > x.timeout_ns) {
                            printf("acg: peer exchange timed out (rank %d waits for rank %d, slot offset %lld, epoch %llu)\n",
                                   x.rank, r, x.slot_off, epoch);
                            __trap();
                        }
                    }
                }
                s += __longlong_as_double((long long)(((unsigned long long)c.z << 32) | c.x));
            }
        }
        if (first) {
            __syncthreads();
            after_wait();
            first = false;
        }
        if (i < n) vec[i] = s;
    }
    __syncthreads();
    if (tid == 0) *x.epoch = epoch;
}

// host: fills a PeerExchange from the C-ABI description; ACG_OK or an error code
int fill_peer_exchange(PeerExchange* x, const acg_peer_exchange* d, int n, const char* who);

}  // namespace acg
