// kernels.cuh -- hand-written fp64 sm_100a kernels of the HPR-LP engine.
//
// One streaming skeleton, `csr_stream_kernel`, carries every pass over a CSR matrix (the two
// fused HPR phases, the KKT residual passes, power iteration, Ruiz/Pock-Chambolle/Curtis-Reid
// row statistics).  It is an nnz-balanced ("merge-style") CSR-stream kernel at WARP granularity:
//
//   item  = a fixed chunk of kWarpChunk (256) consecutive nonzeros owned by ONE WARP, so the work
//           per warp is identical whatever the row-length distribution is (power-law rows
//           included; no row-id lists, no short/long buckets) and warps never wait on each other:
//           there is no CTA barrier on the path (one at kernel entry only), every warp of an SM is at a different point of
//           its item, which is what hides the HBM/L2 latency of the gathers (r1 v1 used CTA-wide
//           items + __syncthreads and was latency-bound: profiles/r1_v1_ncu_full_c2_details.txt).
//   phase 1  the warp streams its nonzeros with 128-bit/64-bit coalesced loads (double2 values,
//           int2 column indices, evict-first), gathers the dense vector through the read-only path
//           (L2 resident) and stores the products in its private slice of shared memory.
//   phase 2  G lanes per row (G chosen per matrix from the mean row length) sum the row's slice of
//           the product array; the row totals are handed to one lane per row, so that the fused
//           epilogue (projection, dual update, Halpern averaging, residual terms ...) runs
//           lane-parallel over consecutive rows with coalesced vector loads/stores.
//   rows cut by an item boundary: every item but the last one of the row publishes its partial sum as one
//           8-byte packet (an aligned 64-bit relaxed store: single-copy atomic, so the value IS the ready flag --
//           an all-ones NaN pattern that no arithmetic produces means "not published"); the warp whose item
//           holds the END of the row waits for the packets of the items to its left, adds them in item order,
//           runs the epilogue and re-arms the packets for the next launch.  A CTA takes the chunk it works on
//           from an atomic ticket (not from blockIdx), so the items to the left always belong to CTAs that
//           STARTED earlier: the decoupled look-back of single-pass scans, deadlock-free under any dispatch
//           order, preemption or residency.  No second launch, no fences, and the summation order is fixed by the
//           item order, so results are bitwise run-to-run reproducible.
//
// Replaces, on the iteration path, the reference's fused_update_* kernels
// (src/cuda_kernels/HPR_cuda_kernels.cu:297-427), its cuSPARSE SpMV + elementwise kernels
// (src/main_iterate.cu:422-481, HPR_cuda_kernels.cu:203-295), the residual kernels + cuBLAS
// reductions (src/main_iterate.cu:207-309, HPR_cuda_kernels.cu:160-189) and advance_halpern_factors
// (HPR_cuda_kernels.cu:192-200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hpr {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
#ifndef HPR_LANE_NNZ
#define HPR_LANE_NNZ 8
#endif
#ifndef HPR_ROUND_NNZ
#define HPR_ROUND_NNZ 8
#endif
#ifndef HPR_MIN_BLOCKS
#define HPR_MIN_BLOCKS 6
#endif
#ifndef HPR_PRE_OPERANDS
#define HPR_PRE_OPERANDS 1   // x-/y-phase: epilogue operands requested at the start of phase 2 (0: behind the row sums)
#endif
#ifndef HPR_META_LATE
#define HPR_META_LATE 1   // row metadata requested behind the first round of the nonzero stream (0: in front of it)
#endif
#ifndef HPR_TEX_GATHER
#define HPR_TEX_GATHER 1
#endif
#ifndef HPR_BULK_STREAM
#define HPR_BULK_STREAM 0   // 1: the col/val item is brought into shared memory by two cp.async.bulk (TMA engine) copies per warp
#endif
constexpr int kLaneNnz = HPR_LANE_NNZ;          // nonzeros per lane per item
constexpr int kRoundNnz = HPR_ROUND_NNZ;        // nonzeros per lane per load round (loads in flight)
constexpr int kWarpChunk = 32 * kLaneNnz;       // nonzeros per warp item
constexpr int kChunk = kWarps * kWarpChunk;     // nonzeros per CTA (array padding granularity)
constexpr int kMaxSlots = 8;                    // reduction slots per CTA
constexpr int kSeqPartials = 8;                 // split rows with more partials are summed by the whole warp

#ifndef HPR_PARTSLOT_DEFINED
#define HPR_PARTSLOT_DEFINED
typedef unsigned long long PartSlot;   // one published partial sum: the bits of the double, kPartEmpty = not published
#endif
constexpr unsigned long long kPartEmpty = ~0ULL;   // a NaN payload no fp64 instruction generates (canonical NaN = 0xfff8...)

template <typename RP>
struct CsrView {
    int rows;
    long long nnz;
    const RP *rowPtr;
    const int *col;       // padded to a multiple of kChunk (pad: col 0, value 0)
    const double *val;    // padded likewise
    const int *item_row;  // n_ctas*kWarps + 1 entries: first row finalised by warp item i
    int n_items;          // CTAs in the grid
    PartSlot *head_part;  // [items * 2] partial of the row entering the item from the left (and leaving it to the right)
    PartSlot *tail_part;  // [items * 2] partial of the row that starts in the item and leaves it to the right
    unsigned *ticket;             // chunk tickets handed out since the counter was last zeroed.  Every launch has exactly
                                  // n_items CTAs and each takes one, so chunk = ticket % n_items without any reset between
                                  // launches; the host zeroes the counter (stream-ordered, between launches) long before it
                                  // can wrap (launch_one, engine.cu).  32-bit on purpose: the 64-bit remainder is a ~100
                                  // instruction routine that one thread would execute while the CTA waits at the barrier.
    unsigned long long *issued_host;   // HOST bookkeeping for that (never dereferenced on the device)
    int chunk_offset;             // chunks are handed out starting at this one (cyclically); 0 except for the row-partitioned
                                  // push pass, where every rank starts at a different x-block so that no GPU is the target
                                  // of all peers at once.  Only the row cut at the wrap-around boundary then waits for an
                                  // item that starts LATER (the very last one): a bounded wait, not a deadlock.
    // Column-banded matrices (engine.cu, build_bands): a pass over the matrix is one launch per band; every band but the
    // last stores its row sums (+ those of the bands before it) in carry_out instead of running the epilogue, the last
    // band adds carry_in to its own row sums first.  Both null for an ordinary matrix.  Single-product ops only.
    const double *carry_in;
    double *carry_out;
};

// Publish / consume one partial sum.  The packet is ONE aligned 64-bit word: value and "ready" cannot be torn and
// no fence is needed (nothing else is communicated through memory).  The consumer re-arms the packet.
__device__ __forceinline__ unsigned long long part_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b == kPartEmpty ? 0xfff8000000000000ULL : b;   // (unreachable for computed values; keeps the protocol total)
}
__device__ __forceinline__ void part_publish(PartSlot *p, double v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(part_bits(v)) : "memory");
}
__device__ __forceinline__ double part_consume(PartSlot *p) {
    unsigned long long a;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
        if (a != kPartEmpty) break;
        __nanosleep(64);   // every poll is a request on the L1->crossbar port, the unit that bounds this kernel
        if (++spins > (1u << 28)) __trap();   // > 15 s: a protocol bug must kill the context, not hang the GPU
    }
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(kPartEmpty) : "memory");
    return __longlong_as_double((long long)a);
}
// The same hand-off between two warps of one CTA goes through shared memory (7 of 8 cut rows): no L2 round trip.
__device__ __forceinline__ void part_publish_cta(PartSlot *p, double v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(part_bits(v)) : "memory");
}
__device__ __forceinline__ double part_consume_cta(PartSlot *p) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(p);
    unsigned long long a;
    unsigned spins = 0;
    do {
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(a) : "r"(s) : "memory");
        if (++spins > (1u << 30)) __trap();
    } while (a == kPartEmpty);
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(s), "l"(kPartEmpty) : "memory");
    return __longlong_as_double((long long)a);
}

__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ int2 ld_stream(const int2 *p) { return __ldcs(p); }

// ---- bulk-async (TMA engine) copy of a contiguous global range into shared memory, completion on an mbarrier -----------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    // make the initialised barrier visible to the async proxy (the TMA engine completes transactions on it); CTA scope on
    // purpose: fence.mbarrier_init.release.cluster compiles to CCTL.IVALL, i.e. an L1 flush per warp item
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gmem_src),
                 "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}

// Block-wide deterministic sum of NS per-thread accumulators; thread 0 writes them to
// out[blockIdx.x * kMaxSlots + s].  The caller's final_reduce_kernel adds the per-CTA values
// in block order, so every reduction is run-to-run reproducible.
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NS], double *out, double *scratch /* >= 8*NS doubles */,
                                                   int block = -1) {
    if (block < 0) block = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double v = acc[s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) scratch[s * (kThreads / 32) + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += scratch[threadIdx.x * (kThreads / 32) + w];
        out[(size_t)block * kMaxSlots + threadIdx.x] = v;
    }
}

template <bool MAX>
__device__ __forceinline__ double combine(double a, double b) { return MAX ? fmax(a, b) : a + b; }

// ------------------------------------------------------------------------------------------------
// The streaming skeleton.  Op supplies:
//   static constexpr int NV      number of product streams (1 or 2)
//   static constexpr bool kMax   combine with fmax instead of +
//   void init()                  per-thread scalar loads
//   void elem(v, col, out[NV])   per-nonzero term
//   void row(r, acc[NV], p0, p1) fused epilogue of a complete row
//   void finish(scratch)         CTA-level reductions (may be empty)
// Dynamic shared memory: kWarps * NV * kWarpChunk doubles (product slices) + reduction scratch.
// ------------------------------------------------------------------------------------------------
template <class Op>
constexpr size_t stream_smem_bytes() {
    return sizeof(double) * ((size_t)kWarps * Op::NV * kWarpChunk + (size_t)kMaxSlots * kWarps) +
           (HPR_BULK_STREAM ? (size_t)kWarps * (sizeof(int) * kWarpChunk + 16) : 0);   // + column-index stage and one mbarrier per warp
}

template <class Op, int G, typename RP>
__global__ void __launch_bounds__(kThreads, HPR_MIN_BLOCKS) csr_stream_kernel(CsrView<RP> M, Op op) {
    constexpr int NV = Op::NV;
    constexpr bool MX = Op::kMax;
    constexpr int RPR = 32 / G;   // rows per reduce round
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *prod = smem + (size_t)warp * NV * kWarpChunk;            // [NV][kWarpChunk], private to this warp
    double *red_scratch = smem + (size_t)kWarps * NV * kWarpChunk;
    __shared__ double own_part[kWarps * 2];     // [warp][2]: this item's share of the row it finishes
    __shared__ PartSlot cta_part[kWarps * 4];   // [warp][head, tail][2]: partials handed to a warp of this CTA
    __shared__ unsigned chunk_s;
    // The chunk this CTA works on comes from a ticket: chunks are handed out in the order CTAs START, so every item to
    // the left of ours belongs to a CTA that is already running (or done) -- the look-back below cannot wait for a CTA
    // that was never scheduled, whatever order the hardware dispatches blockIdx in.
#ifdef HPR_STATIC_CHUNKS   // measurement only (tools/build_variants.sh): chunk = blockIdx, i.e. r1's dispatch-order assumption
    if (threadIdx.x == 0) chunk_s = blockIdx.x;
#else
    if (threadIdx.x == 0) chunk_s = (atomicAdd(M.ticket, 1u) % gridDim.x + (unsigned)M.chunk_offset) % gridDim.x;
#endif
    if (threadIdx.x < kWarps * 4) cta_part[threadIdx.x] = kPartEmpty;
    // Programmatic dependent launch (launch_one sets the attribute; without it both instructions do nothing): the next
    // kernel of the stream may fill the SM slots this grid's last wave leaves free.  Its CTAs draw their tickets and stop at
    // the wait until this grid has completed and its writes are visible; nothing above the wait depends on an earlier kernel.
    // The trigger comes after the barrier, i.e. after this CTA's ticket has RETURNED: a dependent launch on the same matrix
    // draws from the same counter, and a launch must own n_items consecutive tickets.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    op.init();         // its loads (sigma, Halpern counter) do not depend on the chunk: in flight under the ticket's round trip
    __syncthreads();   // the only CTA barrier: before any work, so no warp ever waits for a slower one
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int chunk = (int)chunk_s;

    auto complete_row = [&](int r, double (&t)[NV], long long q0, long long q1) {
        if (M.carry_in) t[0] += M.carry_in[r];
        if (M.carry_out) M.carry_out[r] = t[0];
        else op.row(r, t, q0, q1);
    };
    const int item = chunk * kWarps + warp;
    const int cta_item0 = chunk * kWarps;
    const long long cta_end = ((long long)chunk + 1) * kChunk;   // a row with p1 <= cta_end ends inside this CTA
    const long long s = (long long)item * kWarpChunk;
    const long long e = (s + kWarpChunk < M.nnz) ? s + kWarpChunk : (s < M.nnz ? M.nnz : s);

    // Row metadata of the first batch.  item_row -> rowPtr is a chain of two dependent loads: it is requested BEHIND the
    // first round of the nonzero stream (HPR_META_LATE), so that the warp does not sit on item_row's round trip with
    // nothing else in flight (ncu: 6 % of the x-phase's stall samples were on the address arithmetic that waits for it).
    int rA = 0, rB = 0, r_last = -1;
    long long p0 = 0, p1 = 0;
    auto load_row_meta = [&]() {
        rA = __ldg(M.item_row + item);
        rB = __ldg(M.item_row + item + 1);
        r_last = (rB < M.rows) ? rB : M.rows - 1;   // last row touched (rB included: it may start here)
        if (rA + lane <= r_last) {
            p0 = (long long)M.rowPtr[rA + lane];
            p1 = (long long)M.rowPtr[rA + lane + 1];
        }
    };
    if (!HPR_META_LATE || HPR_BULK_STREAM) load_row_meta();

    // ---- phase 1: stream nonzeros, gather, multiply (kRoundNnz loads in flight per lane) -----------
#if HPR_BULK_STREAM
    if constexpr (NV == 1) {
        // The warp's 3 KB of (col, val) arrive by two bulk-async copies issued by one lane (TMA engine, no LSU wavefronts for the
        // stream); the values land in the product slice itself and are overwritten by the products.
        int *col_stage = reinterpret_cast<int *>(red_scratch + kMaxSlots * kWarps) + warp * kWarpChunk;
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(reinterpret_cast<int *>(red_scratch + kMaxSlots * kWarps) + kWarps * kWarpChunk) + 2 * warp;
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(bar, kWarpChunk * 12);
            bulk_g2s(prod, M.val + s, kWarpChunk * 8, bar);
            bulk_g2s(col_stage, M.col + s, kWarpChunk * 4, bar);
        }
        mbar_wait(bar, 0);
        const int2 *c2 = reinterpret_cast<const int2 *>(col_stage);
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            const int t = u * 32 + lane;
            const int2 cc = c2[t];
            const double2 vv = *reinterpret_cast<const double2 *>(prod + 2 * t);
            double o0[NV], o1[NV];
            op.elem(vv.x, cc.x, o0);
            op.elem_b(vv.y, cc.y, o1);
            *reinterpret_cast<double2 *>(prod + 2 * t) = make_double2(o0[0], o1[0]);
        }
    } else
#endif
    {
        const double2 *v2 = reinterpret_cast<const double2 *>(M.val + s);
        const int2 *c2 = reinterpret_cast<const int2 *>(M.col + s);
#pragma unroll
        for (int rd = 0; rd < kLaneNnz / kRoundNnz; ++rd) {
            double2 vv[kRoundNnz / 2];
            int2 cc[kRoundNnz / 2];
#pragma unroll
            for (int u = 0; u < kRoundNnz / 2; ++u) {
                const int t = (rd * (kRoundNnz / 2) + u) * 32 + lane;
                cc[u] = ld_stream(c2 + t);
                vv[u] = ld_stream(v2 + t);
            }
            if (HPR_META_LATE && !HPR_BULK_STREAM && rd == 0) load_row_meta();
#pragma unroll
            for (int u = 0; u < kRoundNnz / 2; ++u) {
                double o0[NV], o1[NV];
                op.elem(vv[u].x, cc[u].x, o0);
                op.elem_b(vv[u].y, cc[u].y, o1);
                const int t = (rd * (kRoundNnz / 2) + u) * 32 + lane;
#pragma unroll
                for (int q = 0; q < NV; ++q)
                    *reinterpret_cast<double2 *>(prod + q * kWarpChunk + 2 * t) = make_double2(o0[q], o1[q]);
            }
        }
    }
    __syncwarp();

    // ---- phase 2: row sums (G lanes per row) handed to one lane per row, lane-parallel epilogue ------
    for (int base = rA; base <= r_last; base += 32) {   // warp-uniform
        const int r = base + lane;
        const bool valid = r <= r_last;
        if (base != rA) {
            p0 = 0; p1 = 0;
            if (valid) { p0 = (long long)M.rowPtr[r]; p1 = (long long)M.rowPtr[r + 1]; }
        }
        int lo = 0, hi = 0;
        if (valid) {
            const long long a = p0 > s ? p0 : s;
            const long long b = p1 < e ? p1 : e;
            if (b > a) { lo = (int)(a - s); hi = (int)(b - s); }
        }
        // rows this lane completes in this batch (neither entered from the left nor continuing to the right): their
        // epilogue operands are requested now, in flight under the row sums (+4-9 % when the gathers coalesce, +2 % on
        // configs[1], -0.7 % on configs[2]: profiles/r2_operand_prefetch_experiment.md)
        typename Op::Pre pre{};
        bool pre_ok = false;
        if constexpr (Op::kPre) {
            pre_ok = valid && !M.carry_out && !((r == rA) && (p0 < s)) && !(p1 > e);
            if (pre_ok) pre = op.pre(r);
        }
        const int nrows = min(32, r_last - base + 1);
        double tot[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) tot[q] = 0.0;   // |a| >= 0, so 0 is also the identity of fmax here
        if (G == 1) {
            for (int k = lo; k < hi; ++k) {
#pragma unroll
                for (int q = 0; q < NV; ++q) tot[q] = combine<MX>(tot[q], prod[q * kWarpChunk + k]);
            }
        } else {
            const int gl = lane & (G - 1), gid = lane / G;
            const int rounds = (nrows + RPR - 1) / RPR;
            for (int t = 0; t < rounds; ++t) {
                const int o = t * RPR + gid;   // row (offset in the batch) reduced by my group this round
                const int glo = __shfl_sync(0xffffffffu, lo, o);
                const int ghi = __shfl_sync(0xffffffffu, hi, o);
                double acc[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) acc[q] = 0.0;
                for (int k = glo + gl; k < ghi; k += G) {
#pragma unroll
                    for (int q = 0; q < NV; ++q) acc[q] = combine<MX>(acc[q], prod[q * kWarpChunk + k]);
                }
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
                    for (int q = 0; q < NV; ++q) acc[q] = combine<MX>(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], off));
                }
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const double v = __shfl_sync(0xffffffffu, acc[q], (lane % RPR) * G);
                    if (lane / RPR == t) tot[q] = v;   // lane L owns row L of the batch
                }
            }
        }

        // lane-parallel epilogue.  A row cut by an item boundary is finished by the item that holds its end (below);
        // the other items it spans only publish their partial sums.
        if (valid) {
            const bool head = (r == rA) && (p0 < s);   // row entered this item from the left (lane 0, first batch)
            const bool cont = (p1 > e);                // row continues to the right
            if (!head && !cont) {
                if constexpr (Op::kPre) {
                    if (pre_ok) {
                        if (M.carry_in) tot[0] += M.carry_in[r];
                        op.row_pre(r, tot, pre);
                    } else complete_row(r, tot, p0, p1);
                } else complete_row(r, tot, p0, p1);
            } else if (head && !cont) {                // finished below; park this item's share (no live registers)
#pragma unroll
                for (int q = 0; q < NV; ++q) own_part[warp * 2 + q] = tot[q];
            } else if (head || p0 < e) {               // (else: r == rB and it starts in a later item)
                if (p1 <= cta_end) {                   // finished by a later warp of this CTA
                    PartSlot *slot = cta_part + warp * 4 + (head ? 0 : 2);
#pragma unroll
                    for (int q = 0; q < NV; ++q) part_publish_cta(slot + q, tot[q]);
                } else {                               // finished by a later CTA
                    PartSlot *slot = (head ? M.head_part : M.tail_part) + (size_t)item * 2;
#pragma unroll
                    for (int q = 0; q < NV; ++q) part_publish(slot + q, tot[q]);
                }
            }
        }
    }

    // ---- the row that ends here after entering from the left: total = tail[ia] + head[ia+1] + ... + head[ib-1] + own,
    // in item order (kSeqPartials or more partials: stride-32 order + xor tree).  Done last, after this item has published
    // everything other warps may be waiting for.  The lanes fetch the partials in parallel (one wait, not one per partial).
    long long P0 = 0, P1 = 0;
    if (rA <= r_last) { P0 = (long long)M.rowPtr[rA]; P1 = (long long)M.rowPtr[rA + 1]; }   // warp-uniform reload
    if (rA <= r_last && P0 < s && P1 <= e) {
        __syncwarp();   // own_part written by lane 0 above
        const int ia = (int)(P0 / kWarpChunk), ib = item;
        auto fetch = [&](int j, bool tail, int q) -> double {   // partial of item j: same CTA -> shared, earlier CTA -> global
            if (j >= cta_item0) return part_consume_cta(cta_part + (j - cta_item0) * 4 + (tail ? 2 : 0) + q);
            return part_consume((tail ? M.tail_part : M.head_part) + (size_t)j * 2 + q);
        };
        double sum[NV];
        if (ib - ia < kSeqPartials) {
            double v[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) v[q] = (ia + lane < ib) ? fetch(ia + lane, lane == 0, q) : own_part[warp * 2 + q];
#pragma unroll
            for (int q = 0; q < NV; ++q) sum[q] = __shfl_sync(0xffffffffu, v[q], 0);
            for (int t = 1; t <= ib - ia; ++t) {
#pragma unroll
                for (int q = 0; q < NV; ++q) sum[q] = combine<MX>(sum[q], __shfl_sync(0xffffffffu, v[q], t));
            }
        } else {
#pragma unroll
            for (int q = 0; q < NV; ++q) sum[q] = 0.0;
            for (int j = ia + lane; j <= ib; j += 32) {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const double v = (j == ib) ? own_part[warp * 2 + q] : fetch(j, j == ia, q);
                    sum[q] = combine<MX>(sum[q], v);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int q = 0; q < NV; ++q) sum[q] = combine<MX>(sum[q], __shfl_xor_sync(0xffffffffu, sum[q], off));
            }
        }
        if (lane == 0) complete_row(rA, sum, P0, P1);
    }
    if (!M.carry_out) op.finish(red_scratch, chunk);   // (warp-uniform: kernel argument); partials indexed by chunk: deterministic
}

// item_row[i] = first row finalised by warp item i = first r with rowPtr[r+1] > i*kWarpChunk (item 0 also owns
// leading empty rows; items at or beyond the last real one get `rows`, so the last real item owns trailing
// empty rows and pure padding items own nothing).
template <typename RP>
__global__ void build_item_rows_kernel(const RP *rowPtr, int rows, long long nnz, int n_entries, int *item_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    if (i == 0) { item_row[0] = 0; return; }
    const long long target = (long long)i * kWarpChunk;
    if (target >= nnz) { item_row[i] = rows; return; }
    int lo = 0, hi = rows;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)rowPtr[mid + 1] > target) hi = mid; else lo = mid + 1;
    }
    item_row[i] = lo;
}


// ================================================================================================
// Ops
// ================================================================================================
// gather through the TEX pipe when the vector is bound as a linear texture (tex != 0), else ld.global.nc
__device__ __forceinline__ double gather_tex_or_ldg(cudaTextureObject_t tex, const double *vec, int col) {
    if (tex) {
        const int2 t = tex1Dfetch<int2>(tex, col);
        return __hiloint2double(t.y, t.x);
    }
    return __ldg(vec + col);
}

struct OpBase {
    static constexpr int NV = 1;
    static constexpr bool kMax = false;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void finish(double *, int) {}
    // Ops whose epilogue reads per-row operands may declare kPre: the kernel then requests them (pre) when phase 2 starts and
    // hands them to row_pre after the row sums, instead of row() loading them behind the sums.
    static constexpr bool kPre = false;
    struct Pre {};
    __device__ __forceinline__ Pre pre(int) const { return Pre{}; }
};

// x-phase (reference fused_update_x_z_rows_*_kernel, HPR_cuda_kernels.cu:297-361; check variant
// update_zx_check_kernel :203-226):  w = (A^T y)_j ; zt = x + sigma (w - c) ; x_bar = proj_[l,u] zt ;
// x_hat = 2 x_bar - x ; x <- f2 x_hat + f1 x0 ; check also stores x_bar, z_bar=(x_bar-zt)/sigma, x_bar-x_hat.
// Halpern counter: this kernel reads k from kx and mirrors it into ky for the y-phase.
// TEX: gathers through the TEX pipe (y bound as a linear texture); false: ld.global.nc (vectors beyond the texture size limit)
template <bool CHECK, bool TEX = true>
struct XPhaseOp : OpBase {
    const double *y;
    double *x, *x_hat;
    const double *c, *l, *u, *x0;
    double *x_bar, *z_bar, *x_tmp;
    const double *params;   // [sigma, lambda*sigma, 1/(lambda*sigma), 1/sigma]
    const int *kx;
    int *ky;
    double sigma, f1, f2;
    __device__ __forceinline__ void init() {
        sigma = params[0];
        const int k = *kx;
        f1 = 1.0 / (k + 2.0);
        f2 = 1.0 - f1;
        if (blockIdx.x == 0 && threadIdx.x == 0) *ky = k;
    }
    // Gathers are split between the two L1 pipes: HPR_TEX_GATHER == 0 all through the LSU (ld.global.nc),
    // 1 all through the TEX pipe (y bound as an int2 linear texture), 2 first nonzero of each pair LSU, second TEX.
    cudaTextureObject_t tex;
    __device__ __forceinline__ double g_lsu(int col) const { return __ldg(y + col); }
    __device__ __forceinline__ double g_tex(int col) const {
        const int2 t = tex1Dfetch<int2>(tex, col);
        return __hiloint2double(t.y, t.x);
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * ((TEX && HPR_TEX_GATHER == 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { o[0] = v * ((TEX && HPR_TEX_GATHER >= 1) ? g_tex(col) : g_lsu(col)); }
    static constexpr bool kPre = HPR_PRE_OPERANDS != 0;
    struct Pre { double xi, cj, lj, uj, x0j; };
    __device__ __forceinline__ Pre pre(int j) const { return Pre{x[j], c[j], l[j], u[j], x0[j]}; }
    __device__ __forceinline__ void row_pre(int j, const double (&acc)[1], const Pre &p) const { apply(j, acc[0], p.xi, p.cj, p.lj, p.uj, p.x0j); }
    __device__ __forceinline__ void row(int j, const double (&acc)[1], long long, long long) const { apply(j, acc[0], x[j], c[j], l[j], u[j], x0[j]); }
    __device__ __forceinline__ void apply(int j, double w, double xi, double cj, double lj, double uj, double x0j) const {
        const double zt = fma(sigma, w - cj, xi);
        const double xb = fmin(uj, fmax(lj, zt));
        const double xh = 2.0 * xb - xi;
        x[j] = fma(f2, xh, f1 * x0j);
        x_hat[j] = xh;
        if (CHECK) {
            x_bar[j] = xb;
            z_bar[j] = (xb - zt) / sigma;
            x_tmp[j] = xb - xh;
        }
    }
};

// y-phase (reference fused_update_y_rows_*_kernel :363-427; check variant update_y_check_kernel :249-272):
// v = (A x_hat)_i - lambda sigma y ; d = max(AL - v, min(AU - v, 0)) ; y_bar = d/(lambda sigma) ;
// y_hat = 2 y_bar - y ; y <- f2 y_hat + f1 y0 ; check also stores y_bar, y_obj = v + d, y_bar - y_hat.
// Advances the Halpern counter: kx <- ky + 1 (reference advance_halpern_factors_kernel :192-200).
template <bool CHECK, bool TEX = true>
struct YPhaseOp : OpBase {
    const double *x_hat;
    double *y;
    const double *AL, *AU, *y0;
    double *y_bar, *y_obj, *y_tmp;
    const double *params;
    const int *ky;
    int *kx;
    double lamsig, inv_lamsig, f1, f2;
    __device__ __forceinline__ void init() {
        lamsig = params[1];
        inv_lamsig = params[2];
        const int k = *ky;
        f1 = 1.0 / (k + 2.0);
        f2 = 1.0 - f1;
        if (blockIdx.x == 0 && threadIdx.x == 0) *kx = k + 1;
    }
    cudaTextureObject_t tex;   // x_hat as an int2 linear texture (see XPhaseOp)
    __device__ __forceinline__ double g_lsu(int col) const { return __ldg(x_hat + col); }
    __device__ __forceinline__ double g_tex(int col) const {
        const int2 t = tex1Dfetch<int2>(tex, col);
        return __hiloint2double(t.y, t.x);
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * ((TEX && HPR_TEX_GATHER == 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { o[0] = v * ((TEX && HPR_TEX_GATHER >= 1) ? g_tex(col) : g_lsu(col)); }
    static constexpr bool kPre = HPR_PRE_OPERANDS != 0;
    struct Pre { double yi, ALi, AUi, y0i; };
    __device__ __forceinline__ Pre pre(int i) const { return Pre{y[i], AL[i], AU[i], y0[i]}; }
    __device__ __forceinline__ void row_pre(int i, const double (&acc)[1], const Pre &p) const { apply(i, acc[0], p.yi, p.ALi, p.AUi, p.y0i); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) const { apply(i, acc[0], y[i], AL[i], AU[i], y0[i]); }
    __device__ __forceinline__ void apply(int i, double w, double yi, double ALi, double AUi, double y0i) const {
        const double v = fma(-lamsig, yi, w);
        const double d = fmax(ALi - v, fmin(AUi - v, 0.0));
        const double yb = inv_lamsig * d;
        const double yh = 2.0 * yb - yi;
        y[i] = fma(f2, yh, f1 * y0i);
        if (CHECK) {
            y_bar[i] = yb;
            y_obj[i] = v + d;
            y_tmp[i] = yb - yh;
        }
    }
};

// Dual residual pass over A^T (reference residual_compute_Rd_cusparse + queue_dot/nrm2,
// src/main_iterate.cu:218-226,237-258): slots 0 |Rd|^2, 1 <c,x_bar>, 2 <x_bar,z_bar>, 3 |x_tmp|^2 (gap),
// 4 |bound violation / col_norm|^2 (iteration 0, reference residual_compute_lu_kernel :174-180).
template <bool GAP, bool ITER0>
struct ResidualDualOp : OpBase {
    const double *y_bar, *c, *z_bar, *x_bar, *x_tmp, *col_norm, *l, *u;
    cudaTextureObject_t tex;   // y_bar as a texture (0: plain loads)
    double *partials;
    double t[5];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 5; ++s) t[s] = 0.0;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * gather_tex_or_ldg(tex, y_bar, col); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int j, const double (&acc)[1], long long, long long) {
        const double cj = c[j], zb = z_bar[j], xb = x_bar[j], cn = col_norm[j];
        const double rd = (cj - acc[0] - zb) * cn;
        t[0] += rd * rd;
        t[1] += cj * xb;
        t[2] += xb * zb;
        if (GAP) { const double dx = x_tmp[j]; t[3] += dx * dx; }
        if (ITER0) {
            const double lj = l[j], uj = u[j];
            const double viol = (xb < lj) ? (lj - xb) : ((xb > uj) ? (xb - uj) : 0.0);
            const double q = viol / cn;
            t[4] += q * q;
        }
    }
    __device__ __forceinline__ void finish(double *scratch, int block) { block_reduce_store<5>(t, partials, scratch, block); }
};

// Primal residual pass over A (reference residual_compute_Rp_cusparse, src/main_iterate.cu:207-215,
// plus the restart-gap SpMV/dots :245-254): slots 0 |Rp|^2, 1 <y_obj,y_bar>, 2 <A x_tmp, y_tmp>, 3 |y_tmp|^2.
template <bool GAP>
struct ResidualPrimalOp : OpBase {
    static constexpr int NV = GAP ? 2 : 1;
    const double *x_bar, *x_tmp, *AL, *AU, *row_norm, *y_obj, *y_bar, *y_tmp;
    cudaTextureObject_t tex;   // x_bar as a texture (0: plain loads)
    double *partials;
    double t[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 4; ++s) t[s] = 0.0;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[NV]) const {
        o[0] = v * gather_tex_or_ldg(tex, x_bar, col);
        if (GAP) o[NV - 1] = v * __ldg(x_tmp + col);
    }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[NV]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int i, const double (&acc)[NV], long long, long long) {
        const double ax = acc[0];
        const double rp = fmax(fmin(AU[i] - ax, 0.0), AL[i] - ax) * row_norm[i];
        t[0] += rp * rp;
        t[1] += y_obj[i] * y_bar[i];
        if (GAP) { const double dy = y_tmp[i]; t[2] += acc[NV - 1] * dy; t[3] += dy * dy; }
    }
    __device__ __forceinline__ void finish(double *scratch, int block) { block_reduce_store<4>(t, partials, scratch, block); }
};

// M-norm cross term after a restart iteration (reference compute_weighted_norm,
// src/main_iterate.cu:486-515): slots 0 <A dx, dy>, 1 |dy|^2.
struct WeightedNormOp : OpBase {
    const double *dx, *dy;
    cudaTextureObject_t tex;   // dx as a texture (0: plain loads)
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * gather_tex_or_ldg(tex, dx, col); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) {
        const double d = dy[i];
        t[0] += acc[0] * d;
        t[1] += d * d;
    }
    __device__ __forceinline__ void finish(double *scratch, int block) { block_reduce_store<2>(t, partials, scratch, block); }
};

// Plain SpMV out = M * g, with optional fused <out,out> and <q,out> (power iteration,
// reference src/power_iteration.cu:73-90).
template <bool DOTS, bool TEX = false>
struct SpmvOp : OpBase {
    const double *g;
    cudaTextureObject_t tex;   // g as an int2 linear texture when TEX (gathers through the TEX pipe, see XPhaseOp)
    double *out;
    const double *q;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const {
        if (TEX) {
            const int2 w = tex1Dfetch<int2>(tex, col);
            o[0] = v * __hiloint2double(w.y, w.x);
        } else {
            o[0] = v * __ldg(g + col);
        }
    }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) {
        out[i] = acc[0];
        if (DOTS) { t[0] += acc[0] * acc[0]; t[1] += q[i] * acc[0]; }
    }
    __device__ __forceinline__ void finish(double *scratch, int block) {
        if (DOTS) block_reduce_store<2>(t, partials, scratch, block);
    }
};

// Row-partitioned x-side pass with the reduce-scatter fused in (collective.h, PeerExchange): w_p = A_p^T y_p, and row j
// of the result is stored straight into the receive slot that the OWNER of column j keeps for this rank -- a peer store
// over NVLink for remote owners -- so the transfer runs under the pass instead of after it.
struct SpmvPushOp : OpBase {
    const double *g;
    cudaTextureObject_t tex;   // g as an int2 linear texture (0: plain loads)
    double *slot[16];          // slot[q][j] = where row j goes if rank q owns it (= recv buffer of q + (rank - q) * xblock)
    int xblock;
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * gather_tex_or_ldg(tex, g, col); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int j, const double (&acc)[1], long long, long long) const { slot[j / xblock][j] = acc[0]; }
};

// Row statistic for Ruiz (sqrt max|a|) and Pock-Chambolle (sqrt sum|a|) scaling
// (reference CSR_A_row_norm_kernel, HPR_cuda_kernels.cu:91-120); <1e-15 -> 1.  RAW: store max / sum only (the
// row-partitioned mode reduces the column statistic across GPUs before the sqrt + clamp).
template <bool MAX, bool RAW>
struct RowNormOp : OpBase {
    static constexpr bool kMax = MAX;
    double *out;
    __device__ __forceinline__ void elem(double v, int, double (&o)[1]) const { o[0] = fabs(v); }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) const {
        if (RAW) { out[i] = acc[0]; return; }
        double r = sqrt(acc[0]);
        if (r < 1e-15) r = 1.0;
        out[i] = r;
    }
};

// Curtis-Reid log-domain sweep (reference curtis_reid_log_update_kernel, src/scaling.cu:5-31).  RAW: store the
// sum and the entry count (row-partitioned mode: mean over all GPUs' entries).
template <bool RAW>
struct CurtisReidOp : OpBase {
    const double *other;
    double *out;
    double *cnt_out;
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const {
        o[0] = -log(fmax(fabs(v), 1e-300)) - __ldg(other + col);
    }
    __device__ __forceinline__ void elem_b(double v, int col, double (&o)[1]) const { elem(v, col, o); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long p0, long long p1) const {
        const long long cnt = p1 - p0;
        if (RAW) { out[i] = acc[0]; cnt_out[i] = (double)cnt; return; }
        out[i] = cnt > 0 ? acc[0] / (double)cnt : 0.0;
    }
};

// ------------------------------------------------------------------------------------------------
// Matrix value scaling, one pass: v <- (v op f_first) op f_second with two separately rounded
// operations in the reference's order (mul_CSR_A_row then mul_CSR_AT_row, HPR_cuda_kernels.cu:122-157;
// call order src/scaling.cu:72-76,136-141): for A   first = rowfac[row],  second = gathfac[col];
//                                           for A^T first = gathfac[col], second = rowfac[row].
// CTA-level items (kChunk nonzeros = kWarps warp items): row factors are expanded into shared memory by
// the row owners, then the nnz-parallel phase is fully coalesced.
// ------------------------------------------------------------------------------------------------
template <bool DIVIDE, bool ROW_FIRST, typename RP>
__global__ void __launch_bounds__(kThreads, 4)
scale_values_kernel(CsrView<RP> M, double *val_rw, const double *rowfac, const double *gathfac) {
    __shared__ double rf[kChunk];
    const long long s = (long long)blockIdx.x * kChunk;
    const long long e = (s + kChunk < M.nnz) ? s + kChunk : M.nnz;
    const int rA = M.item_row[blockIdx.x * kWarps];
    const int rB = M.item_row[(blockIdx.x + 1) * kWarps];
    constexpr int GG = 8;
    const int gl = threadIdx.x & (GG - 1);
    const int gid = threadIdx.x / GG;
    for (int r = rA + gid; r <= rB && r < M.rows; r += kThreads / GG) {
        const long long p0 = (long long)M.rowPtr[r], p1 = (long long)M.rowPtr[r + 1];
        const long long a = p0 > s ? p0 : s, b = p1 < e ? p1 : e;
        const double f = rowfac[r];
        for (long long k = a + gl; k < b; k += GG) rf[(int)(k - s)] = f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (int)(e - s); t += kThreads) {
        double v = val_rw[s + t];
        const double fr = rf[t];
        const double fg = __ldg(gathfac + M.col[s + t]);
        const double f1 = ROW_FIRST ? fr : fg;
        const double f2 = ROW_FIRST ? fg : fr;
        if (DIVIDE) { v = v / f1; v = v / f2; } else { v = v * f1; v = v * f2; }
        val_rw[s + t] = v;
    }
}

}  // namespace hpr
