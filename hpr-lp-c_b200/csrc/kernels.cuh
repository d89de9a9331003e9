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
//   layout  inside an item the col/val arrays are stored LANE-CONTIGUOUS (r2 v5): the 128-bit / 64-bit coalesced loads
//           of lane t (positions (u*32 + t)*2 + h, u = 0..3) deliver the 8 CONSECUTIVE nonzeros 8t .. 8t+7 of the item.
//           The permutation is applied in place on the device once per matrix (permute_items_kernel); everything that
//           indexes nonzeros logically goes through item_slot().
//   phase 1 the warp streams its nonzeros (double2 values, int2 column indices, evict-first), gathers the dense vector
//           through the TEX pipe (L2 resident) and multiplies: 8 products per lane, in registers.
//   phase 2 (r2 v5) every lane sums its own products row segment by row segment -- a per-lane byte of row-start flags
//           (precomputed from rowPtr, 1 bit per nonzero) says where a row begins; segments that span lanes are closed by
//           one segmented warp scan (5 shuffle steps).  Only ONE value per row segment goes through shared memory (the
//           segment total, picked up by the lane that runs the row's epilogue) instead of one store + one load per
//           nonzero as in r1: ~13 instead of 256 shared-memory stores per item on a 20-nonzeros-per-row matrix.  The
//           fused epilogue (projection, dual update, Halpern averaging, residual terms ...) runs lane-parallel over
//           consecutive rows with coalesced vector loads/stores, as before.
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
#ifndef HPR_TEX_GATHER
#define HPR_TEX_GATHER 1
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
    const int *col;       // padded to a multiple of kChunk (pad: col 0, value 0); lane-contiguous item order (item_slot)
    const double *val;    // padded likewise
    const int *item_row;  // n_ctas*kWarps + 1 entries: first row finalised by warp item i
    int n_items;          // CTAs in the grid
    const unsigned char *flags;   // [items * 32] row-start bits of lane t of item i at flags[i*32 + t] (build_row_flags_kernel)
    PartSlot *head_part;  // [items] partial of the row entering the item from the left (and leaving it to the right)
    PartSlot *tail_part;  // [items] partial of the row that starts in the item and leaves it to the right
    unsigned long long *ticket;   // chunk tickets handed out so far over ALL launches on this matrix (never reset: every
                                  // launch has exactly n_items CTAs and each takes one, so chunk = ticket % n_items)
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
// Lane-contiguous item layout: logical nonzero k (0..255) of an item is stored at slot item_slot(k), so that the
// coalesced vector loads of lane t (double2 / int2 at index u*32 + t, u = 0..3) return nonzeros 8t .. 8t+7.
// ------------------------------------------------------------------------------------------------
static_assert(kLaneNnz == 8, "item layout is written for 8 nonzeros per lane");
__host__ __device__ __forceinline__ int item_slot(int k) { return ((((k & 7) >> 1) * 32 + (k >> 3)) << 1) | (k & 1); }
__host__ __device__ __forceinline__ int item_logical(int p) { return (((p >> 1) & 31) << 3) | (((p >> 1) >> 5) << 1) | (p & 1); }
// position in the stored (permuted) arrays of logical nonzero q of the whole matrix
__host__ __device__ __forceinline__ long long stored_pos(long long q) { return (q & ~(long long)(kWarpChunk - 1)) + item_slot((int)(q & (kWarpChunk - 1))); }

// In-place switch of one matrix between the logical CSR order and the lane-contiguous item order; one warp per item.
template <bool TO_ITEM_ORDER>
__global__ void permute_items_kernel(int *col, double *val, long long n_witems) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_witems) return;
    int *c = col + item * kWarpChunk;
    double *v = val + item * kWarpChunk;
    int cc[kLaneNnz];
    double vv[kLaneNnz];
#pragma unroll
    for (int j = 0; j < kLaneNnz; ++j) {   // slot read by this lane -> slot written by this lane (a bijection on 0..255)
        const int k = lane * kLaneNnz + j;
        const int src = TO_ITEM_ORDER ? k : item_slot(k);
        cc[j] = c[src];
        vv[j] = v[src];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kLaneNnz; ++j) {
        const int k = lane * kLaneNnz + j;
        const int dst = TO_ITEM_ORDER ? item_slot(k) : k;
        c[dst] = cc[j];
        v[dst] = vv[j];
    }
}

// Row-start flags: bit (q & 7) of flags[q >> 3] is set iff a NONEMPTY row starts at logical nonzero q.  flags[item*32 + t]
// is therefore the byte of lane t of that item.  (flags zeroed before; words are updated with atomicOr.)
template <typename RP>
__global__ void build_row_flags_kernel(const RP *rowPtr, int rows, unsigned *flag_words) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long q = (long long)rowPtr[r];
    if ((long long)rowPtr[r + 1] <= q) return;
    const long long byte = q >> 3;
    atomicOr(flag_words + (byte >> 2), 1u << (((int)(byte & 3) << 3) + (int)(q & 7)));
}

// ------------------------------------------------------------------------------------------------
// The streaming skeleton.  Op supplies:
//   static constexpr bool kMax   combine with fmax instead of +
//   void init()                  per-thread scalar loads
//   double elem(v, col)          per-nonzero term (elem_b: the same for the second nonzero of a pair)
//   void row(r, acc, p0, p1)     fused epilogue of a complete row
//   void finish(scratch, block)  CTA-level reductions (may be empty)
// Dynamic shared memory: kWarps * kSegStride doubles (row-segment totals) + reduction scratch.
// ------------------------------------------------------------------------------------------------
constexpr int kSegStride = kWarpChunk + 8;   // up to kWarpChunk + 1 row segments per item (every nonzero its own row + the leading one)
template <class Op>
constexpr size_t stream_smem_bytes() {
    return sizeof(double) * ((size_t)kWarps * kSegStride + (size_t)kMaxSlots * kWarps);
}

template <class Op, typename RP>
__global__ void __launch_bounds__(kThreads, HPR_MIN_BLOCKS) csr_stream_kernel(CsrView<RP> M, Op op) {
    constexpr bool MX = Op::kMax;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *seg = smem + (size_t)warp * kSegStride;                  // row-segment totals of this warp's item
    double *red_scratch = smem + (size_t)kWarps * kSegStride;
    __shared__ double own_part[kWarps];         // this item's share of the row it finishes
    __shared__ PartSlot cta_part[kWarps * 2];   // [warp][head, tail]: partials handed to a later warp of this CTA
    __shared__ unsigned chunk_s;
    // The chunk this CTA works on comes from a ticket: chunks are handed out in the order CTAs START, so every item to
    // the left of ours belongs to a CTA that is already running (or done) -- the look-back below cannot wait for a CTA
    // that was never scheduled, whatever order the hardware dispatches blockIdx in.
    if (threadIdx.x == 0) chunk_s = (unsigned)(atomicAdd(M.ticket, 1ULL) % (unsigned long long)gridDim.x);
    if (threadIdx.x < kWarps * 2) cta_part[threadIdx.x] = kPartEmpty;
    __syncthreads();   // the only CTA barrier: before any work, so no warp ever waits for a slower one
    const int chunk = (int)chunk_s;

    op.init();
    auto complete_row = [&](int r, double t, long long q0, long long q1) {
        if (M.carry_in) t += M.carry_in[r];
        if (M.carry_out) M.carry_out[r] = t;
        else op.row(r, t, q0, q1);
    };
    const int item = chunk * kWarps + warp;
    const int cta_item0 = chunk * kWarps;
    const long long cta_end = ((long long)chunk + 1) * kChunk;   // a row with p1 <= cta_end ends inside this CTA
    const long long s = (long long)item * kWarpChunk;
    const long long e = (s + kWarpChunk < M.nnz) ? s + kWarpChunk : (s < M.nnz ? M.nnz : s);

    // row metadata of the first batch and the row-start flags: requested before the nonzeros so their latency overlaps phase 1
    const int rA = __ldg(M.item_row + item);
    const int rB = __ldg(M.item_row + item + 1);
    const int r_last = (rB < M.rows) ? rB : M.rows - 1;   // last row touched (rB included: it may start here)
    const unsigned f = __ldg(M.flags + (size_t)item * 32 + lane);   // bit j: a row starts at my nonzero j
    long long p0 = 0, p1 = 0;
    if (rA + lane <= r_last) {
        p0 = (long long)M.rowPtr[rA + lane];
        p1 = (long long)M.rowPtr[rA + lane + 1];
    }

    // ---- phase 1: stream nonzeros, gather, multiply: 8 consecutive nonzeros of the item per lane, all loads in flight ------
    double prod[kLaneNnz];
    {
        const double2 *v2 = reinterpret_cast<const double2 *>(M.val + s);
        const int2 *c2 = reinterpret_cast<const int2 *>(M.col + s);
        double2 vv[kLaneNnz / 2];
        int2 cc[kLaneNnz / 2];
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            cc[u] = ld_stream(c2 + u * 32 + lane);
            vv[u] = ld_stream(v2 + u * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            prod[2 * u] = op.elem(vv[u].x, cc[u].x);
            prod[2 * u + 1] = op.elem_b(vv[u].y, cc[u].y);
        }
    }
    if (e - s < kWarpChunk) {   // last item of the matrix (warp-uniform): padding slots contribute the identity
#pragma unroll
        for (int j = 0; j < kLaneNnz; ++j)
            if (s + lane * kLaneNnz + j >= e) prod[j] = 0.0;
    }

    // ---- phase 2a: row-segment sums.  Segment 0 = nonzeros before the first row start of the item (the tail of a row
    // that entered from the left, possibly empty); segment i >= 1 starts at the i-th row start.  A lane closes the
    // segments that end inside it on its own; the ones that run across lanes are closed by a segmented scan.
    int nf_before = 0;   // row starts in lower lanes
#pragma unroll
    for (int j = 0; j < kLaneNnz; ++j) nf_before += __popc(__ballot_sync(0xffffffffu, (f >> j) & 1u) & ((1u << lane) - 1u));
    const bool has = f != 0u;
    double acc = 0.0, head = 0.0;   // |a| >= 0, so 0 is also the identity of fmax here
    int seen = 0;
#pragma unroll
    for (int j = 0; j < kLaneNnz; ++j) {
        if ((f >> j) & 1u) {
            if (seen == 0) head = acc;                 // end of the segment that entered this lane from the left
            else seg[nf_before + seen] = acc;          // a segment that started at my previous row start: complete
            acc = 0.0;
            ++seen;
        }
        acc = combine<MX>(acc, prod[j]);
    }
    // inclusive segmented scan of the open right ends: S(t) = acc(t) [+ S(t-1) if lane t has no row start]
    const unsigned hasmask = __ballot_sync(0xffffffffu, has);
    double S = acc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, S, off);
        // lanes lane-off+1 .. lane must all be free of row starts for the two runs to belong to one segment
        const bool join = lane >= off && ((hasmask >> (lane - off + 1)) & ((off == 32 ? 0u : (1u << off)) - 1u)) == 0u;
        if (join) S = combine<MX>(up, S);
    }
    const double left = __shfl_up_sync(0xffffffffu, S, 1);
    if (has) seg[nf_before] = combine<MX>(lane > 0 ? left : 0.0, head);   // the segment that ends at my first row start
    if (lane == 31) seg[nf_before + seen] = S;                          // the last segment of the item
    __syncwarp();

    // ---- phase 2b: lane-parallel epilogue over the rows of the item; a row's total is its segment's ----------------------
    const int seg0 = (rA <= r_last && (long long)M.rowPtr[rA <= r_last ? rA : 0] < s) ? 0 : 1;   // does row rA enter from the left?
    int seen_rows = 0;   // nonempty rows of the item in earlier batches
    for (int base = rA; base <= r_last; base += 32) {   // warp-uniform
        const int r = base + lane;
        const bool valid = r <= r_last;
        if (base != rA) {
            p0 = 0; p1 = 0;
            if (valid) { p0 = (long long)M.rowPtr[r]; p1 = (long long)M.rowPtr[r + 1]; }
        }
        bool nonempty = false;
        if (valid) {
            const long long a = p0 > s ? p0 : s;
            const long long b = p1 < e ? p1 : e;
            nonempty = b > a;
        }
        const unsigned ne_mask = __ballot_sync(0xffffffffu, nonempty);
        const double tot = nonempty ? seg[seg0 + seen_rows + __popc(ne_mask & ((1u << lane) - 1u))] : 0.0;
        seen_rows += __popc(ne_mask);

        // A row cut by an item boundary is finished by the item that holds its end (below);
        // the other items it spans only publish their partial sums.
        if (valid) {
            const bool hd = (r == rA) && (p0 < s);     // row entered this item from the left (lane 0, first batch)
            const bool cont = (p1 > e);                // row continues to the right
            if (!hd && !cont) {
                complete_row(r, tot, p0, p1);
            } else if (hd && !cont) {                  // finished below; park this item's share (no live registers)
                own_part[warp] = tot;
            } else if (hd || p0 < e) {                 // (else: r == rB and it starts in a later item)
                if (p1 <= cta_end) {                   // finished by a later warp of this CTA
                    part_publish_cta(cta_part + warp * 2 + (hd ? 0 : 1), tot);
                } else {                               // finished by a later CTA
                    part_publish((hd ? M.head_part : M.tail_part) + (size_t)item, tot);
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
        auto fetch = [&](int j, bool tail) -> double {   // partial of item j: same CTA -> shared, earlier CTA -> global
            if (j >= cta_item0) return part_consume_cta(cta_part + (j - cta_item0) * 2 + (tail ? 1 : 0));
            return part_consume((tail ? M.tail_part : M.head_part) + (size_t)j);
        };
        double sum;
        if (ib - ia < kSeqPartials) {
            const double v = (ia + lane < ib) ? fetch(ia + lane, lane == 0) : own_part[warp];
            sum = __shfl_sync(0xffffffffu, v, 0);
            for (int t = 1; t <= ib - ia; ++t) sum = combine<MX>(sum, __shfl_sync(0xffffffffu, v, t));
        } else {
            sum = 0.0;
            for (int j = ia + lane; j <= ib; j += 32) {
                const double v = (j == ib) ? own_part[warp] : fetch(j, j == ia);
                sum = combine<MX>(sum, v);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum = combine<MX>(sum, __shfl_xor_sync(0xffffffffu, sum, off));
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
struct OpBase {
    static constexpr bool kMax = false;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void finish(double *, int) {}
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
    __device__ __forceinline__ double elem(double v, int col) const { return v * ((TEX && HPR_TEX_GATHER == 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return v * ((TEX && HPR_TEX_GATHER >= 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ void row(int j, double acc0, long long, long long) const {
        const double xi = x[j];
        const double zt = fma(sigma, acc0 - c[j], xi);
        const double xb = fmin(u[j], fmax(l[j], zt));
        const double xh = 2.0 * xb - xi;
        x[j] = fma(f2, xh, f1 * x0[j]);
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
    __device__ __forceinline__ double elem(double v, int col) const { return v * ((TEX && HPR_TEX_GATHER == 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return v * ((TEX && HPR_TEX_GATHER >= 1) ? g_tex(col) : g_lsu(col)); }
    __device__ __forceinline__ void row(int i, double acc0, long long, long long) const {
        const double yi = y[i];
        const double v = fma(-lamsig, yi, acc0);
        const double d = fmax(AL[i] - v, fmin(AU[i] - v, 0.0));
        const double yb = inv_lamsig * d;
        const double yh = 2.0 * yb - yi;
        y[i] = fma(f2, yh, f1 * y0[i]);
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
    double *partials;
    double t[5];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 5; ++s) t[s] = 0.0;
    }
    __device__ __forceinline__ double elem(double v, int col) const { return v * __ldg(y_bar + col); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int j, double acc0, long long, long long) {
        const double cj = c[j], zb = z_bar[j], xb = x_bar[j], cn = col_norm[j];
        const double rd = (cj - acc0 - zb) * cn;
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

// Primal residual pass over A (reference residual_compute_Rp_cusparse, src/main_iterate.cu:207-215):
// slots 0 |Rp|^2, 1 <y_obj,y_bar>.  (The restart-gap terms <A x_tmp, y_tmp>, |y_tmp|^2 of :245-254 are a second
// single-product pass, WeightedNormOp.)
struct ResidualPrimalOp : OpBase {
    const double *x_bar, *AL, *AU, *row_norm, *y_obj, *y_bar;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ double elem(double v, int col) const { return v * __ldg(x_bar + col); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int i, double ax, long long, long long) {
        const double rp = fmax(fmin(AU[i] - ax, 0.0), AL[i] - ax) * row_norm[i];
        t[0] += rp * rp;
        t[1] += y_obj[i] * y_bar[i];
    }
    __device__ __forceinline__ void finish(double *scratch, int block) { block_reduce_store<2>(t, partials, scratch, block); }
};

// M-norm cross term after a restart iteration (reference compute_weighted_norm,
// src/main_iterate.cu:486-515): slots 0 <A dx, dy>, 1 |dy|^2.
struct WeightedNormOp : OpBase {
    const double *dx, *dy;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ double elem(double v, int col) const { return v * __ldg(dx + col); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int i, double acc0, long long, long long) {
        const double d = dy[i];
        t[0] += acc0 * d;
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
    __device__ __forceinline__ double elem(double v, int col) const {
        if (TEX) {
            const int2 w = tex1Dfetch<int2>(tex, col);
            return v * __hiloint2double(w.y, w.x);
        }
        return v * __ldg(g + col);
    }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int i, double acc0, long long, long long) {
        out[i] = acc0;
        if (DOTS) { t[0] += acc0 * acc0; t[1] += q[i] * acc0; }
    }
    __device__ __forceinline__ void finish(double *scratch, int block) {
        if (DOTS) block_reduce_store<2>(t, partials, scratch, block);
    }
};

// Row statistic for Ruiz (sqrt max|a|) and Pock-Chambolle (sqrt sum|a|) scaling
// (reference CSR_A_row_norm_kernel, HPR_cuda_kernels.cu:91-120); <1e-15 -> 1.  RAW: store max / sum only (the
// row-partitioned mode reduces the column statistic across GPUs before the sqrt + clamp).
template <bool MAX, bool RAW>
struct RowNormOp : OpBase {
    static constexpr bool kMax = MAX;
    double *out;
    __device__ __forceinline__ double elem(double v, int) const { return fabs(v); }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int i, double acc0, long long, long long) const {
        if (RAW) { out[i] = acc0; return; }
        double r = sqrt(acc0);
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
    __device__ __forceinline__ double elem(double v, int col) const {
        return -log(fmax(fabs(v), 1e-300)) - __ldg(other + col);
    }
    __device__ __forceinline__ double elem_b(double v, int col) const { return elem(v, col); }
    __device__ __forceinline__ void row(int i, double acc0, long long p0, long long p1) const {
        const long long cnt = p1 - p0;
        if (RAW) { out[i] = acc0; cnt_out[i] = (double)cnt; return; }
        out[i] = cnt > 0 ? acc0 / (double)cnt : 0.0;
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
    for (int t = threadIdx.x; t < kChunk; t += kThreads) {   // t = stored slot; the row factor is indexed by the logical position
        const int lt = (t & ~(kWarpChunk - 1)) + item_logical(t & (kWarpChunk - 1));
        if (s + lt >= e) continue;
        double v = val_rw[s + t];
        const double fr = rf[lt];
        const double fg = __ldg(gathfac + M.col[s + t]);
        const double f1 = ROW_FIRST ? fr : fg;
        const double f2 = ROW_FIRST ? fg : fr;
        if (DIVIDE) { v = v / f1; v = v / f2; } else { v = v * f1; v = v * f2; }
        val_rw[s + t] = v;
    }
}

}  // namespace hpr
