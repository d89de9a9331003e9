// engine.h -- private state of the B200-native HPR-LP engine (not part of the ABI).
// Plays the role of the reference's HPRLP_workspace_gpu / LP_info_gpu / Scaling_info /
// HPRLP_restart / HPRLP_residuals (reference include/structs.h:127-277), minus every cuSPARSE /
// cuBLAS handle: all device work is done by the kernels in kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/structs.h"
#include "../../include/hprlp_b200.h"
#include "collective.h"
#include "nccl_shim.h"

namespace hpr {

#define HPR_CUDA_CHECK(call)                                                                  \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            throw std::runtime_error(std::string("CUDA error at ") + __FILE__ + ":" +         \
                                     std::to_string(__LINE__) + ": " + cudaGetErrorString(err__)); \
        }                                                                                     \
    } while (0)

#ifndef HPR_PARTSLOT_DEFINED
#define HPR_PARTSLOT_DEFINED
typedef unsigned long long PartSlot;   // one published partial sum (kernels.cuh): bits of the double, ~0 = not published
#endif

// Device CSR matrix + the item decomposition used by csr_stream_kernel.
struct DevCsr {
    int rows = 0, cols = 0;
    long long nnz = 0;
    int *rowPtr = nullptr;
    int *col = nullptr;     // padded to n_items * kChunk
    double *val = nullptr;  // padded likewise
    int *item_row = nullptr;
    int n_items = 0;
    PartSlot *head_part = nullptr, *tail_part = nullptr;   // partial sums of rows cut by item boundaries (all-ones = empty)
    unsigned *ticket = nullptr;             // chunk tickets handed out by csr_stream_kernel since the last zeroing (kernels.cuh)
    unsigned long long tickets_issued = 0;  // host count of the same: the counter is re-zeroed between launches before it can wrap
    int push_chunk0 = 0;    // first chunk of the row-partitioned push pass of this rank (kernels.cuh, CsrView::chunk_offset)
    int G = 1;              // lanes per row in phase 2, from the row-length statistics (pick_lanes)
    double mean_len = 0.0;
    double len_cv = 0.0;    // std/mean of the row lengths
    int max_len = 0;
    // Column bands (Engine::build_bands): when the gathered vector is larger than the L2 can hold, the passes over this
    // matrix run band by band over column slices that do fit, carrying the row sums in `carry` (rows doubles).
    std::vector<DevCsr> bands;
    double *carry = nullptr;
    void *band_store = nullptr;   // one allocation behind all band arrays
    void *band_rowptr_store = nullptr;
};

struct RestartState {   // reference HPRLP_restart, include/structs.h:215-228
    int restart_flag = 0;
    bool first_restart = true;
    double last_gap = std::numeric_limits<double>::infinity();
    double current_gap = std::numeric_limits<double>::infinity();
    double save_gap = std::numeric_limits<double>::infinity();
    double best_gap = std::numeric_limits<double>::infinity();
    double best_sigma = 1.0;
    int inner = 0, sufficient = 0, necessary = 0, long_ = 0, times = 0;
};

struct Residuals {      // reference HPRLP_residuals, include/structs.h:255-263
    double err_Rp = 0, err_Rd = 0, primal_obj = 0, dual_obj = 0, rel_gap = 0;
    double kkt = std::numeric_limits<double>::infinity();
};

// Optional hooks used by tests/bench through include/hprlp_b200.h.
struct SolveHooks {
    const double *power_z0 = nullptr;   // host, length m: overrides the cuRAND start vector
    int n_trace = 0;                    // snapshots of unscaled (x_bar,y_bar,z_bar) after the check
    const int *trace_iters = nullptr;   //   iteration with iter+1 == trace_iters[t]
    double *trace_x = nullptr, *trace_y = nullptr, *trace_z = nullptr;
    bool quiet = false;                 // suppress the stdout log
    // filled on return
    double lambda_max = 0, sigma = 0, setup_seconds = 0, scaling_seconds = 0, power_seconds = 0;
    double loop_device_ms = 0;          // CUDA-event time of the main loop on the engine stream
    int restarts = 0, power_iters = 0;
    long long kernel_launches = 0;
    double scal[6] = {0, 0, 0, 0, 0, 0};  // b_scale,c_scale,norm_b,norm_c,norm_b_org,norm_c_org
};

struct LoopState {
    int iter = 0;
    RestartState rs;
    Residuals res;
    HPRLP_results output;
    bool first_4 = true, first_6 = true, first_8 = true;
    double uploaded_sigma = 0, uploaded_lambda = 0;
    double t_start_alg = 0;
};

class Engine {
   public:
    Engine() = default;
    ~Engine();
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    // ---- setup (reference copy_lpinfo_to_device + allocate_memory, src/preprocess.cu:66-256) ----
    void upload(const LP_info_cpu *lp, int device);
    void prepare(int m, int n, long long nnz, int device);   // allocate; caller fills A + vectors, then finish_setup()
    void finish_setup(bool build_transpose);
    void spmv_A(const double *g, double *out);
    void spmv_AT(const double *g, double *out);
    // device-resident CSR (already on this GPU); arrays are copied into padded engine storage
    void upload_device(int m, int n, long long nnz, const int *d_rowPtr, const int *d_col, const double *d_val,
                       const double *d_AL, const double *d_AU, const double *d_l, const double *d_u,
                       const double *d_c, double obj_constant, int device);
    // ---- reference scaling(), src/scaling.cu:88-216 ----
    void scale(const HPRLP_parameters *p);
    // ---- reference power_method_cusparse(), src/power_iteration.cu:20-119 (returns lambda, un-multiplied) ----
    double power_iteration(int max_iter, double tol, const double *host_z0, int *iters_out);
    void power_start_vector(double *d_z);   // cuRAND XORWOW seed 1 N(0,1) + 1e-8 (odd m: 1e-8)
    // ---- reference HPRLP_main_solve loop, src/HPRLP.cu:150-311 ----
    HPRLP_results solve(const HPRLP_parameters *p, SolveHooks *hooks);
    void solve_begin(const HPRLP_parameters *p, SolveHooks *hooks);
    bool solve_advance(const HPRLP_parameters *p, SolveHooks *hooks, int pause_at);
    void fill_hooks(SolveHooks *hooks);
    LoopState loop;

    // building blocks (also driven step-wise by the extended API)
    void init_iterates();
    void upload_params();
    void reset_halpern_counter();
    void launch_iteration(bool check);
    void launch_x_phase(bool check);
    void launch_y_phase(bool check);
    void run_normal(int count);            // count normal iterations (CUDA-graph replay)
    void compute_residuals(int iter, bool compute_gap, Residuals *res, RestartState *rs);
    double weighted_norm_after_restart();
    void restart_and_sigma(RestartState *rs, const Residuals &res);
    void collect_solution(double *hx, double *hy, double *hz);
    double time_phase_ms(int which, int reps);   // CUDA-event timing of one fused kernel (bench roofline)

    int m = 0, n = 0;
    long long nnz = 0;
    int device = 0;
    // Row-block partition over several GPUs: this engine owns rows [row0, row0+m) of a global
    // m_global x n problem -- A_p, A_p^T, the y-side vectors of those rows -- and the x-block J_p = [xb0, xb1) of
    // columns: the x-update, the x-side residual terms and the movement norms run on J_p only.  Per iteration:
    // partial A_p^T y_p -> reduce-scatter -> x-update on J_p -> all-gather of x_hat -> fused y-phase on the local rows.
    // n-vectors are allocated with npad = nranks * xblock entries so both exchanges run in place.
    // coll == nullptr: single GPU.
    Collective *coll = nullptr;
    PeerExchange *px = nullptr;    // NVLink peer-memory exchange (collective.h); null: NCCL reduce-scatter + all-gather
    // reduce-scatter + x-update on the owned block + all-gather, by either transport (mode 0); modes 1-3: the other exchanges of the
    // partitioned mode on the peer-memory transport (fused_exchange_x_kernel, engine.cu)
    void exchange_x(bool check, int mode = 0, const double *src = nullptr);
    void partial_ATy_pass(const double *g = nullptr, cudaTextureObject_t tex = 0);   // w_p = A_p^T g_p (g = y by default) into wn, or (peer exchange) pushed into the owners' receive slots
    bool push_mode() const { return px != nullptr && AT.bands.empty(); }
    int nranks = 1, rank = 0, m_global = 0, row0 = 0;
    int xb0 = 0, xb1 = 0;          // owned columns
    size_t xblock = 0, npad = 0;   // exchange block (multiple of 64 entries), padded n-vector length
    bool dist() const { return coll != nullptr; }
    void set_partition(Collective *c, int m_global_, int row0_);   // before upload()/prepare()
    void allreduce(double *buf, size_t count, bool max_op = false);
    int mg() const { return dist() ? m_global : m; }   // length of the y returned by collect_solution
    DevCsr A, AT;
    double *AL = nullptr, *AU = nullptr, *c = nullptr, *l = nullptr, *u = nullptr;
    double *row_norm = nullptr, *col_norm = nullptr;
    double b_scale = 1, c_scale = 1, norm_b = 0, norm_c = 0, norm_b_org = 1, norm_c_org = 1;
    double obj_constant = 0;
    double sigma = 1, lambda_max = 1;

    double *x = nullptr, *x0 = nullptr, *x_hat = nullptr, *x_bar = nullptr, *z_bar = nullptr, *x_tmp = nullptr;
    double *y = nullptr, *y0 = nullptr, *y_bar = nullptr, *y_obj = nullptr, *y_tmp = nullptr;
    double *wn = nullptr, *wm = nullptr, *wm2 = nullptr;   // scratch: n, m, m

    double *d_params = nullptr;     // [sigma, lambda*sigma, 1/(lambda*sigma), 1/sigma]
    int *d_k = nullptr;             // [kx, ky] Halpern counters
    double *d_partials = nullptr;   // per-CTA reduction partials
    int partial_blocks = 0;
    double *d_scal = nullptr;       // 16 reduced scalars
    double *h_scal = nullptr;       // pinned mirror
    double *h_params = nullptr;     // pinned [4]
    cudaStream_t stream = nullptr;
    long long launches = 0;
    bool pooled_ = false;
    void *arena_ = nullptr;         // single device allocation all buffers are carved from
    double *zo_buf = nullptr;       // n doubles for the unscaled z on output
    cudaTextureObject_t tex_y = 0, tex_xhat = 0, tex_q = 0, tex_atq = 0;   // gathered vectors bound as int2 linear textures
    cudaTextureObject_t tex_ybar = 0, tex_xbar = 0, tex_xtmp = 0;          // ... of the residual / restart-gap passes

   private:
    void alloc_common();
    void finish_matrix(DevCsr &M);
    void build_bands(DevCsr &M);
    void fetch_scalars(int count);
    std::map<int, cudaGraphExec_t> graphs_;
};

// model layer helpers (api.cpp / mps_reader.cpp)
void free_lp_info_cpu(LP_info_cpu *lp);
bool build_model_from_mps(const char *path, LP_info_cpu *lp);
void csr_transpose_host(int rows, int cols, int nnz, const int *rp, const int *ci, const double *v,
                        int *trp, int *tci, double *tv);

// host_utils.cpp
void copy_mt(void *dst, const void *src, size_t bytes);
void csr_transpose_host_mt(int rows, int cols, int nnz, const int *rp, const int *ci, const double *v, int *trp, int *tci,
                           double *tv);

void band_count(int rows, const int *d_rowPtr, const int *d_col, int band_cols, int n_bands, int *band_rowPtr,
                long long *band_nnz, cudaStream_t st);
void band_fill(int rows, const int *d_rowPtr, const int *d_col, const double *d_val, int band_cols, int n_bands,
               const int *band_rowPtr, int *const *d_bcol, double *const *d_bval, cudaStream_t st);

void device_transpose_csr(int rows, int cols, int nnz, const int *d_rowPtr, const int *d_col, const double *d_val,
                          int *d_trp, int *d_tcol, double *d_tval, cudaStream_t st);

// row-partitioned mode (partitioned.cu)
std::vector<int> row_blocks_by_nnz(const int *rowPtr, int m, int P);
void upload_row_block(Engine &eng, const LP_info_cpu *model, int r0, int r1, int device);
Collective *open_nccl_rank(const char *uid128, int rank, int nranks, int device);
}  // namespace hpr
// one rank's endpoint of a process-per-GPU partitioned solve (include/hprlp_b200.h)
struct hprlp_b200_comm {
    std::unique_ptr<hpr::Collective> coll;
    int device = 0;
};
namespace hpr {
void fill_b200_info(const Engine &eng, const SolveHooks &h, hprlp_b200_info *info);

// device memory pool of the engines, staged copies of large pageable arrays (engine.cu)
void release_cached_device_memory();
void warm_device(int device);   // first-solve warm-up: context, modules, cuRAND (errors ignored)
void *pool_alloc_zeroed(size_t bytes, int device, cudaStream_t st);
void pool_free(void *p, cudaStream_t st);
void *pool_alloc_raw(size_t bytes, int device, cudaStream_t st);
void h2d_large(void *dst, const void *src, size_t bytes, cudaStream_t stream);   // src may be reused on return; dst is stream-ordered
void d2h_large(void *dst, const void *src, size_t bytes, cudaStream_t stream);   // returns when the copy is complete

// presolve bridge (presolve.cpp); returns false when unavailable / failed (caller solves the original model)
bool presolve_run(const LP_info_cpu *model, const HPRLP_parameters *param, LP_info_cpu *reduced, void **handle);
void presolve_postsolve(HPRLP_results *result, const LP_info_cpu *original, void *handle,
                        const HPRLP_parameters *param);
void presolve_free(void *handle);

}  // namespace hpr
