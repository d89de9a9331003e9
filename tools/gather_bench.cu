// gather_bench.cu -- measured ceiling of the random 8-byte gather that bounds a CSR SpMV on a uniformly random
// matrix (BASELINE.json's synthetic LPs), on the same launch shape as csr_stream_kernel (256 threads, warp item =
// 256 nonzeros, 8 per lane, grid = nnz / 2048).  Not part of the product; its numbers are the "gather roofline"
// quoted in DESIGN.md section 4 and profiles/r1_gather_ceiling.json.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/gather_bench tools/gather_bench.cu
//   build/gather_bench [nnz]
//
// Modes (per nonzero):
//   idx        4 B index stream only (floor of the index stream)
//   idxval     4 B index + 8 B value stream, no gather (floor of the 12 B/nnz stream = HBM roofline of the pass)
//   lsu        index stream + ld.global.nc gather
//   lsu_na     index stream + ld.global.nc.L1::no_allocate gather
//   tex        index stream + tex1Dfetch<int2> gather
//   mix        index stream + alternating LSU / TEX gathers
//   half       index stream + LSU gathers issued 16 lanes at a time
//   hash       gathers only, indices computed in registers (no stream at all)
//   spmv_tex   index + value stream + TEX gather + FMA (the whole phase-1 of the product kernel, no reduction)
//   tmaT       index stream, T of the 8 gathers per lane through cp.async.bulk (16 B, TMA unit -> shared memory,
//              bypasses the L1 tag stage), the rest through TEX
//   dsmemC     index stream + gathers from DISTRIBUTED SHARED MEMORY: the vector (V <= C * slice) is spread over the shared
//              memory of a C-CTA thread-block cluster (C = 8 portable, 16 non-portable), one 1024-thread CTA per SM, every
//              gather an ld.shared::cluster to the CTA that holds the entry (only vectors <= ~1.6 MB fit)
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kThreads = 256, kLaneNnz = 8, kWarpChunk = 32 * kLaneNnz, kChunk = 8 * kWarpChunk;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void fill_kernel(int *col, double *val, long long nnz, int V) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t h = mix64((uint64_t)i);
        col[i] = (int)(h % (uint64_t)V);
        val[i] = 1.0 + (double)(h >> 40) * 1e-9;
    }
}
__global__ void fill_vec_kernel(double *g, int V) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) g[i] = 1.0 + 1e-6 * (i & 1023);
}

__device__ __forceinline__ double ld_nc(const double *p) { return __ldg(p); }
__device__ __forceinline__ double ld_nc_na(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_tex(cudaTextureObject_t t, int i) {
    const int2 w = tex1Dfetch<int2>(t, i);
    return __hiloint2double(w.y, w.x);
}

enum Mode { IDX, IDXVAL, LSU, LSU_NA, TEX, MIX, HALF, HASH, SPMV_TEX };

template <int MODE>
__global__ void __launch_bounds__(kThreads, 8)
gather_kernel(const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ g,
              cudaTextureObject_t tex, int V, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long s = ((long long)blockIdx.x * 8 + warp) * kWarpChunk;
    const int2 *c2 = reinterpret_cast<const int2 *>(col + s);
    const double2 *v2 = reinterpret_cast<const double2 *>(val + s);
    int2 cc[kLaneNnz / 2];
    double2 vv[kLaneNnz / 2];
    double acc = 0.0;
    if (MODE == HASH) {
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            const uint64_t h = mix64((uint64_t)(s + u * 64 + lane));
            cc[u].x = (int)((h & 0xffffffffu) % (unsigned)V);
            cc[u].y = (int)((h >> 32) % (unsigned)V);
        }
    } else {
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            cc[u] = __ldcs(c2 + u * 32 + lane);
            if (MODE == IDXVAL || MODE == SPMV_TEX) vv[u] = __ldcs(v2 + u * 32 + lane);
        }
    }
#pragma unroll
    for (int u = 0; u < kLaneNnz / 2; ++u) {
        if (MODE == IDX) acc += (double)(cc[u].x ^ cc[u].y);
        else if (MODE == IDXVAL) acc += vv[u].x * (double)cc[u].x + vv[u].y * (double)cc[u].y;
        else if (MODE == LSU || MODE == HASH) acc += ld_nc(g + cc[u].x) + ld_nc(g + cc[u].y);
        else if (MODE == LSU_NA) acc += ld_nc_na(g + cc[u].x) + ld_nc_na(g + cc[u].y);
        else if (MODE == TEX) acc += ld_tex(tex, cc[u].x) + ld_tex(tex, cc[u].y);
        else if (MODE == MIX) acc += ld_nc(g + cc[u].x) + ld_tex(tex, cc[u].y);
        else if (MODE == SPMV_TEX) acc = fma(vv[u].x, ld_tex(tex, cc[u].x), fma(vv[u].y, ld_tex(tex, cc[u].y), acc));
        else if (MODE == HALF) {
            double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
            if (lane < 16) a = ld_nc(g + cc[u].x);
            if (lane >= 16) b = ld_nc(g + cc[u].x);
            if (lane < 16) c = ld_nc(g + cc[u].y);
            if (lane >= 16) d = ld_nc(g + cc[u].y);
            acc += a + b + c + d;
        }
    }
    if (acc == 123.456) out[blockIdx.x * kThreads + threadIdx.x] = acc;   // never true: keeps the loads alive
}

// T of the 8 gathers per lane go through the TMA unit (cp.async.bulk, 16-byte aligned pair containing the element),
// landing in the warp's private shared slice and signalled on a per-warp mbarrier; the rest through TEX.
template <int T>
__global__ void __launch_bounds__(kThreads, 6)
gather_tma_kernel(const int *__restrict__ col, const double *__restrict__ g, cudaTextureObject_t tex, int V, double *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 *slice = reinterpret_cast<double2 *>(smem) + (size_t)warp * 32 * T;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + (size_t)8 * 32 * T * 16) + warp;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long s = ((long long)blockIdx.x * 8 + warp) * kWarpChunk;
    const int2 *c2 = reinterpret_cast<const int2 *>(col + s);
    int2 cc[kLaneNnz / 2];
#pragma unroll
    for (int u = 0; u < kLaneNnz / 2; ++u) cc[u] = __ldcs(c2 + u * 32 + lane);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(32 * T * 16) : "memory");
    __syncwarp();
    int cols[kLaneNnz];
#pragma unroll
    for (int u = 0; u < kLaneNnz / 2; ++u) { cols[2 * u] = cc[u].x; cols[2 * u + 1] = cc[u].y; }
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slice + k * 32 + lane);
        const double *src = g + (cols[k] & ~1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                     ::"r"(dst), "l"(src), "r"(mb) : "memory");
    }
    double acc = 0.0;
#pragma unroll
    for (int k = T; k < kLaneNnz; ++k) acc += ld_tex(tex, cols[k]);
    // wait for the bulk copies (phase 0)
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(mb) : "memory");
    }
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const double2 w = slice[k * 32 + lane];
        acc += (cols[k] & 1) ? w.y : w.x;
    }
    if (acc == 123.456) out[blockIdx.x * kThreads + threadIdx.x] = acc;
}

namespace cg = cooperative_groups;
// slice = 2^SHIFT doubles of the vector per CTA of the cluster
template <int SHIFT>
__global__ void __launch_bounds__(1024, 1)
gather_dsmem_kernel(const int *__restrict__ col, const double *__restrict__ g, int V, long long n_items, double *out) {
    extern __shared__ __align__(16) double vec[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank();
    constexpr int slice = 1 << SHIFT;
    for (int i = threadIdx.x; i < slice; i += blockDim.x) {
        const long long gi = (long long)r * slice + i;
        vec[i] = gi < V ? g[gi] : 0.0;
    }
    cluster.sync();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warps_total = (long long)gridDim.x * 32;
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * 32 + warp; item < n_items; item += warps_total) {
        const int2 *c2 = reinterpret_cast<const int2 *>(col + item * kWarpChunk);
        int2 cc[kLaneNnz / 2];
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) cc[u] = __ldcs(c2 + u * 32 + lane);
#pragma unroll
        for (int u = 0; u < kLaneNnz / 2; ++u) {
            const double *a = cluster.map_shared_rank(vec + (cc[u].x & (slice - 1)), (unsigned)(cc[u].x >> SHIFT));
            const double *b = cluster.map_shared_rank(vec + (cc[u].y & (slice - 1)), (unsigned)(cc[u].y >> SHIFT));
            acc += *a + *b;
        }
    }
    cluster.sync();   // no CTA may retire while a peer can still read its shared memory
    if (acc == 123.456) out[blockIdx.x * 1024 + threadIdx.x] = acc;
}

struct Result { std::string mode; int V; double ms, gps; };

int main(int argc, char **argv) {
    long long nnz = argc > 1 ? atoll(argv[1]) : 100000000LL;
    nnz = (nnz / kChunk) * kChunk;
    const int grid = (int)(nnz / kChunk);
    int *col; double *val, *out;
    CK(cudaMalloc(&col, nnz * sizeof(int)));
    CK(cudaMalloc(&val, nnz * sizeof(double)));
    CK(cudaMalloc(&out, (size_t)grid * kThreads * sizeof(double)));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"nnz\": %lld, \"grid\": %d, \"results\": [\n", prop.name, prop.multiProcessorCount, nnz, grid);
    const int Vs[] = {100000, 1000000, 2000000, 5000000, 50000000};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bool first = true;
    for (int V : Vs) {
        double *g; CK(cudaMalloc(&g, (size_t)V * sizeof(double)));
        fill_vec_kernel<<<1024, 256>>>(g, V);
        fill_kernel<<<4096, 256>>>(col, val, nnz, V);
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = g; rd.res.linear.desc = cudaCreateChannelDesc<int2>();
        rd.res.linear.sizeInBytes = (size_t)V * sizeof(double);
        cudaTextureDesc td; memset(&td, 0, sizeof(td)); td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        CK(cudaDeviceSynchronize());
        auto run = [&](const char *name, auto launch) {
            for (int i = 0; i < 3; ++i) launch();
            CK(cudaDeviceSynchronize());
            const int reps = 10;
            CK(cudaEventRecord(e0));
            for (int i = 0; i < reps; ++i) launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
            printf("%s {\"mode\": \"%s\", \"V\": %d, \"ms\": %.4f, \"gnnz_per_s\": %.2f, \"equiv_GBps_12B\": %.1f}", first ? " " : ",\n ", name, V, ms,
                   nnz / ms * 1e-6, 12.0 * nnz / ms * 1e-6);
            first = false;
            fflush(stdout);
        };
#define RUN(M) run(#M, [&] { gather_kernel<M><<<grid, kThreads>>>(col, val, g, tex, V, out); })
        if (V == Vs[0]) { RUN(IDX); RUN(IDXVAL); }
        RUN(LSU); RUN(LSU_NA); RUN(TEX); RUN(MIX); RUN(HALF); RUN(HASH); RUN(SPMV_TEX);
#define RUNT(T) do { const size_t sm = (size_t)8 * 32 * T * 16 + 64; \
        CK(cudaFuncSetAttribute(gather_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        run("tma" #T, [&] { gather_tma_kernel<T><<<grid, kThreads, sm>>>(col, g, tex, V, out); }); } while (0)
        RUNT(1); RUNT(2); RUNT(4); RUNT(8);
        if (V <= 131072) {   // the vector fits a cluster's distributed shared memory
            auto dsmem = [&](const char *name, auto kernel, int cluster_size, int shift) {
                const size_t sm = sizeof(double) << shift;
                CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                if (cluster_size > 8) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
                const int n_clusters = prop.multiProcessorCount / cluster_size;
                cfg.gridDim = dim3(n_clusters * cluster_size); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = sm;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                const long long n_items = nnz / kWarpChunk;
                run(name, [&] { CK(cudaLaunchKernelEx(&cfg, kernel, (const int *)col, (const double *)g, V, n_items, out)); });
            };
            dsmem("dsmem8", gather_dsmem_kernel<14>, 8, 14);     // 8 x 16384 doubles (128 KB per CTA)
            dsmem("dsmem16", gather_dsmem_kernel<13>, 16, 13);   // 16 x 8192 doubles (64 KB per CTA)
        }
        CK(cudaDestroyTextureObject(tex));
        CK(cudaFree(g));
    }
    printf("\n]}\n");
    return 0;
}
