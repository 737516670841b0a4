// common.cuh -- shared declarations of the vbmc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/vbmc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "vbmc_b200 is written for sm_100a (B200) only"
#endif

namespace vbmc {

// ----------------------------------------------------------------------------- errors
void set_error(const std::string &msg);
#define VBMC_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::vbmc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +  \
                              __FILE__ + ":" + std::to_string(__LINE__) + ")");            \
            return VBMC_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
#define VBMC_REQUIRE(cond, code, msg)   \
    do {                                \
        if (!(cond)) {                  \
            ::vbmc::set_error(msg);     \
            return code;                \
        }                               \
    } while (0)
#define VBMC_TRY(expr)              \
    do {                            \
        int _rc = (expr);           \
        if (_rc != VBMC_OK) return _rc; \
    } while (0)

// ----------------------------------------------------------------------------- layout
// D is padded to a multiple of 4 (DP) so that per-dimension tables are float4/double2
// addressable and register arrays have a compile-time size.
static inline int pad_dim(int D) {
    const int opts[] = {4, 8, 12, 16, 20, 24, 28, 32};
    for (int o : opts)
        if (D <= o) return o;
    return -1;
}
constexpr int kMaxD = 32;

// processed hyper-parameter record per GP sample s (doubles), see ctx.cu:pack_hyp
//   [0..DP)        ell_d
//   [DP..2DP)      xm_d          (negquad mean)
//   [2DP..3DP)     1/omega_d^2   (negquad mean)
//   [3DP+0]        ln sf^2
//   [3DP+1]        sum_d ln ell_d
//   [3DP+2]        m0
//   [3DP+3]        sn2_eff
//   [3DP+4]        L_chol (0/1)
static inline int hyp_stride(int DP) { return 3 * DP + 8; }

// device parameter block (doubles): what the kernels read for one evaluation
struct ParamLayout {
    int D, DP, K;
    __host__ __device__ int mu() const { return 0; }                  // [K][D] component-major
    __host__ __device__ int sigma() const { return K * D; }           // [K]
    __host__ __device__ int lambd() const { return K * D + K; }       // [D]
    __host__ __device__ int w() const { return K * D + K + D; }       // [K]
    __host__ __device__ int eta() const { return K * D + 2 * K + D; } // [K]
    __host__ __device__ int lnsig_b() const { return K * D + 3 * K + D; }     // [K]
    __host__ __device__ int lnlam_b() const { return K * D + 4 * K + D; }     // [D]
    __host__ __device__ int eta_b() const { return K * D + 4 * K + 2 * D; }   // [K]
    __host__ __device__ int total() const { return K * D + 5 * K + 2 * D; }
};

// raw (pre-Jacobian) vector exchanged between ranks:  [H, G, flag, 0 | ent block | gp block]
// each block = [gmu (K*D) | gsig (K) | glam (D) | gw (K)]
struct RawLayout {
    int D, K;
    __host__ __device__ int block() const { return K * D + 2 * K + D; }
    __host__ __device__ int ent() const { return 4; }
    __host__ __device__ int gp() const { return 4 + block(); }
    __host__ __device__ int total() const { return 4 + 2 * block(); }
    __host__ __device__ int o_mu() const { return 0; }
    __host__ __device__ int o_sig() const { return K * D; }
    __host__ __device__ int o_lam() const { return K * D + K; }
    __host__ __device__ int o_w() const { return K * D + K + D; }
};

// out vector:  [F, G, H, varF, varG_ss, Lbound, Lpen, nonfinite | dF (Pfull) | dH (Pfull) | dG (Pfull)]
constexpr int kOutHead = 8;

// entmc per-CTA partial record (doubles): [hacc | A (DP) | Be (DP) | racc (K)]
static inline int entpart_stride(int DP, int K) { return 1 + 2 * DP + K; }

struct EvalFlags {
    int grad[4];       // which gradient groups are wanted
    int jacobian;      // apply log / softmax Jacobians
    int use_ent_mc;    // 1: entmc partials present, 0: entlb raw already in place
    int have_ent, have_gp;
    int use_bounds;
    int optimize[4];
    int avg;
    int parts;  // also write dH and dG (the out vector may live in pinned HOST memory: every byte crosses PCIe)
};

// ----------------------------------------------------------------------------- context
struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: gplj overlaps entmc
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t launches = 0;

    // GP pack
    bool has_gp = false, has_L = false;
    int gD = 0, gDP = 0, N = 0, S = 0, mean_kind = VBMC_MEAN_NEGQUAD;
    double *d_Xt = nullptr;    // [DP][N] transposed, padded rows are zero
    double *d_alpha = nullptr; // [S][N]
    double *d_hyp = nullptr;   // [S][hyp_stride]
    double *d_L = nullptr;     // [S][N][N]
    double *d_Linv = nullptr;  // gppred.cu: [S][N][NP] L^-1 (or L) + [S][DP][N] scaled inputs, built on first use
    double *d_xs = nullptr, *d_pred = nullptr;  // search points in, [f_mu | f_s2] or [y | dy] out
    size_t xs_cap = 0, pred_cap = 0;

    // bounds
    int n_bnd = 0;
    double *d_lb = nullptr, *d_ub = nullptr;
    size_t bnd_cap = 0;
    double tol_con = 0, w_thr = 0, w_pen = 0;

    // per-evaluation buffers (grown on demand)
    double *d_in = nullptr, *h_in = nullptr;
    size_t in_cap = 0;
    double *h_theta = nullptr;  // pinned: [theta (P) | template block (param_len) | key (2) | vp_out (2K + D)] (vbmc_negelcbo_theta)
    size_t theta_cap = 0;
    double *d_entpart = nullptr;
    size_t entpart_cap = 0;
    double *d_lamc = nullptr; // [S][K][D] lambda-gradient contributions of the log joint
    size_t lamc_cap = 0;
    double *d_gps = nullptr; // [S][1 + block] per-sample raw log-joint terms
    size_t gps_cap = 0;
    size_t finalize_smem_set = 0;
    // arguments of the last reduce stage (for the fused assemble+finalize launch)
    bool red_args_valid = false;
    int red_plan_slabs = 0, red_s_begin = 0, red_s_step = 1, red_S_glob = 1;
    double red_Ns_glob = 0, red_draws_local = 0;
    double *d_raw = nullptr;
    size_t raw_cap = 0;
    double *d_tctab = nullptr;  // per-component table images of the tensor-core entmc kernel
    size_t tctab_cap = 0;
    // Noise-tile images of the tensor-core entmc kernel: two buffers.  The tiles are unscaled standard normals (a
    // function of the Philox key only), so while the tail of evaluation n runs, a side-stream launch fills the OTHER
    // buffer with the draws of evaluation n + 1 (key offset + 1): device-resident loops (vbmc_adam_steps,
    // vbmc_negelcbo_enqueue, the split-phase multi-GPU step) never wait for the generator.
    double *d_tctiles[2] = {nullptr, nullptr};
    size_t tctiles_cap[2] = {0, 0};
    int noise_buf = 0;             // buffer the next main kernel reads
    bool noise_ready = false;      // ... already holds the draws tagged below (look-ahead launch of the previous evaluation)
    uint64_t noise_seed = 0, noise_offset = 0;   // Philox key the ready buffer was generated for
    uint64_t noise_sig[6] = {0, 0, 0, 0, 0, 0};  // shape / work split it was generated for
    bool noise_needs_wait = false; // generated by vbmc_noise_prefetch on the side stream: not yet ordered before the main stream
    uint64_t cur_seed = 0, cur_offset = 0;       // host mirror of the key currently behind the parameter block
    int64_t key_delta = 0;         // evaluation = key + key_delta (vbmc_negelcbo_enqueue: one fresh key per call)
    bool lookahead = false;        // the caller's next evaluation uses key + key_delta + 1
    bool noise_pending_join = false;
    cudaEvent_t ev_main = nullptr, ev_noise = nullptr;
    // root fork: an evaluation whose draws were NOT generated ahead runs its (theta-independent) tile generator on a
    // third stream forked BEFORE the parameter kernel, beside parameter block + tables; it reads the Philox key from
    // the pinned host block through its device alias (`key_host`), so it does not depend on the parameter kernel
    cudaStream_t stream3 = nullptr;
    cudaEvent_t ev_root = nullptr, ev_root_join = nullptr;
    bool root_forked = false;
    const double *key_host = nullptr;
    double *d_key = nullptr;  // device copy of the key for the root-forked generator (one PCIe read per evaluation, not one per CTA)
    double *d_csum = nullptr;  // [K][entpart_stride] per-component record sums (tail kernel scratch)
    unsigned *d_tailsync = nullptr;  // barrier words of the tail kernel (grid mode)
    size_t csum_cap = 0;
    // peer-memory all-reduce of the raw vector (vbmc_p2p_export / vbmc_p2p_open): exchange buffers of all ranks
    int p2p_world = 0, p2p_rank = 0, p2p_stride = 0;
    unsigned long long p2p_epoch = 0;
    double *p2p_local = nullptr;  // owned (cudaMalloc, exported through CUDA IPC)
    size_t p2p_bytes = 0;
    double *p2p_peer[VBMC_P2P_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool raw_pending_p2p = false;
    bool raw_pending = false;  // reduce stage deferred into the next finalize launch (single GPU, or W ranks with P2P)
    alignas(8) unsigned char raw_pending_blob[320];
    double *d_out = nullptr, *h_out = nullptr;
    size_t out_cap = 0;
    double *d_eps = nullptr;
    size_t eps_cap = 0;
    double *d_lbws = nullptr; // entlb workspace
    size_t lbws_cap = 0;
    double *d_var = nullptr;  // variance-path workspace
    size_t var_cap = 0;
    double *d_outs = nullptr; // per-sample finalize outputs (avg_flag == 0)
    size_t outs_cap = 0;
    double *d_bprm = nullptr, *d_bout = nullptr;  // batched sieve evaluation: parameter blocks in, [B][4] out
    size_t bprm_cap = 0, bout_cap = 0;

    // state of the evaluation uploaded by vbmc_negelcbo_upload
    bool staged = false;
    int D = 0, DP = 0, K = 0;
    vbmc_elcbo_in cur{};
    int ent_grid_slabs = 0;

    // fp32 entropy kernel selection (VBMC_ENTMC_VARIANT / VBMC_ENTMC_GUARD environment overrides)
    int entmc_variant = -1;  // auto (ENTMC_WARP for large draw counts, ENTMC_FAST otherwise)
    float entmc_guard = 1024.0f;  // worst-case per-term relative error of the expanded form ~ 8e-8 * guard

    // optional per-stage timeline (VBMC_STAGE_TIMING=1): events on the main stream
    bool stage_timing = false;
    cudaEvent_t sev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double host_us = 0;  // wall time of the last synchronous evaluation call

    // entmc kernel timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool time_entmc = false;
    double entmc_ms_sum = 0;
    int64_t entmc_ms_n = 0;
    cudaEvent_t ev2 = nullptr;      // recorded right before the dominant kernel (entmc_kernel_tc) when timing is on
    bool ev2_recorded = false;
    double entmc_main_ms_sum = 0;   // ... that kernel alone (without the table / generator launch in front of it)
};

static inline void stage_mark(Ctx *c, int i) {
    if (c->stage_timing) cudaEventRecord(c->sev[i], c->stream);
}
// piecewise-uniform split of the flattened (component, pair) space over CTAs (see EntmcPlan::n_big)
struct ChunkMap {
    long long big, small;
    int n_big;
    __host__ __device__ long long start(int c) const {
        return c < n_big ? (long long)c * big : (long long)n_big * big + (long long)(c - n_big) * small;
    }
    __host__ __device__ int cta_of(long long p) const {
        const long long B = (long long)n_big * big;
        return p < B ? (int)(p / big) : n_big + (int)((p - B) / small);
    }
};
int ensure(double **p, size_t *cap, size_t need);
int ensure_pinned(double **d, double **h, size_t *cap, size_t need);

// ----------------------------------------------------------------------------- launchers
// entmc.cu
enum { ENTMC_FAST = 0, ENTMC_DSPLIT = 1, ENTMC_PACKED = 2, ENTMC_SCALAR = 3, ENTMC_WARP = 4, ENTMC_TC = 5, ENTMC_SMALL = 6 };
struct EntmcPlan {
    int variant;
    int64_t chunk;    // ENTMC_WARP / ENTMC_TC: pairs of the flattened (component, pair) space per CTA
    // ENTMC_TC: the first n_big CTAs take `chunk` pairs, the others `chunk_small` (one tile less): with 10.56 tiles per
    // SM at C3, two resident CTAs of 6 + 5 tiles finish sooner than 6 + 6 next to SMs holding a single CTA
    int64_t chunk_small;
    int n_big;
    int grid, maxseg; // ENTMC_WARP: CTAs, records reserved per CTA
    int threads, slabs, pairs_per_thread;
    int64_t half;      // pairs per component handled by THIS rank
    int64_t pair0;     // first pair index (global) of this rank's range
    int64_t half_glob; // pairs per component over all ranks (Ns/2)
    size_t smem;
};
int entmc_plan(const Ctx *c, int D, int K, int64_t half_local, bool wgrad, int precision, EntmcPlan *plan);
int entmc_launch(Ctx *c, const double *d_params, int D, int K, const EntmcPlan &plan, bool anygrad,
                 bool wgrad, int precision, int rng_mode, const double *d_eps, uint64_t seed,
                 uint64_t offset, double *d_part);
int philox_normals_launch(Ctx *c, int D, int K, int64_t half, uint64_t seed, uint64_t offset, double *d_eps);
// entmc_tc.cu (tcgen05 / TMEM kernel)
int entmc_tc_prefetch(Ctx *c, ParamLayout lay, const EntmcPlan &plan, uint64_t seed, uint64_t offset, bool root = false);
bool entmc_tc_supported(int DP, int K);
int entmc_tc_plan(const Ctx *c, int D, int K, int64_t half_local, EntmcPlan *plan);
int entmc_tc_launch(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool philox,
                    const double *d_eps, double *d_part);

// gplj.cu
int gplj_launch(Ctx *c, const double *d_params, int K, int s_begin, int s_step, bool anygrad, cudaStream_t stream,
                double *Zout /* [S][K][N] or null */);
// gpvar.cu
size_t gpvar_workspace(int S, int K, int N);
double *gpvar_Z(Ctx *c);
double *gpvar_J(Ctx *c, int K);
double *gpvar_out(Ctx *c, int K);  // [varG, var_ss, varG_s (S)]
int gpvar_launch(Ctx *c, const double *d_params, int K, int avg);

// gppred.cu: GP predictive mean / variance at search points, variational-posterior density
int gppred_launch(Ctx *c, const double *d_Xs, int Nx, double *d_mu, double *d_s2);
int gppred_prepare(Ctx *c);  // builds d_Linv = [S][N][NP] L^-1 (Cholesky samples) or L (low-noise samples) once per GP
int gppred_np(int N);
int vp_pdf_launch(Ctx *c, const double *d_params, int D, int K, const double *d_Xs, int Nx, int log_flag, int grad_flag,
                  double *d_y, double *d_dy);

// entlb.cu
int entlb_launch(Ctx *c, const double *d_params, int D, int K, const int grad[4], double *d_raw_ent /*H at [0], block*/,
                 double *d_H);

// adam.cu: device-resident minimize_adam state (all pointers are device memory)
struct AdamDev {
    ParamLayout lay;
    int P, opt[4];
    double *theta, *m, *v;   // [P]
    const double *tmpl;      // [lay.total()] parameter block supplying the groups theta does not carry
    const double *lb, *ub;   // [P] or null
    double *xtab, *ytab;     // [max_iter][P], [max_iter]
    long long *iter;         // iterations done so far
    uint64_t seed, offset0;
    double master_min, master_max, master_decay;
};
int adam_prepare_launch(Ctx *c, const AdamDev &a, double *d_prm);
int adam_update_launch(Ctx *c, const AdamDev &a, const double *d_out);
int adam_update_prepare_launch(Ctx *c, const AdamDev &a, const double *d_out, double *d_prm);  // + next iteration's block
// theta (+ template + key, all in pinned host memory) -> parameter block; vp_out = [sigma | lambda | w] (pinned host)
int theta_prepare_launch(Ctx *c, const AdamDev &a, double *d_prm, double *vp_out, const uint64_t *key_src);

// sieve.cu: batched value-only negative ELCBO (entlb + log joint + bounds) of B candidates
int sieve_launch(Ctx *c, int B, int D, int K, const int optimize[4], bool use_bounds, const double *d_prm, double *d_out);

// finalize.cu
int stage_copy_launch(Ctx *c, double *d_dst, const double *h_pinned_src, int n, cudaStream_t st);
int reduce_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const EntmcPlan *plan,
                  int64_t Ns_glob, int s_begin, int s_step, int S_glob, double *d_raw, bool defer);
int finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const double *d_raw,
                    double *d_out);
int gps_finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, double *d_out_s);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum, result valid in every thread; scratch must hold 33 doubles.  Fixed order:
// butterfly inside each warp, then a butterfly over the (<= 32) warp sums => deterministic.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
        if (lane == 0) scratch[32] = r;
    }
    __syncthreads();
    return scratch[32];
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = scratch[0];
    for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
    return r;
}
#endif

}  // namespace vbmc
