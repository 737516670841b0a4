/* vbmc_b200.h -- C ABI of the B200-native PyVBMC ELBO inner loop.
 *
 * The reference (acerbilab/pyvbmc) has no FFI: its hot path is reached through
 * module-level Python functions.  Each entry point below replaces one of them
 * (citations are reference file:line) and is what a ctypes binding on the
 * reference side would call (INTEGRATION.md shows the stub):
 *
 *   vbmc_entmc        <- entmc_vbmc          pyvbmc/entropy/entmc_vbmc.py:6-134
 *   vbmc_entlb        <- entlb_vbmc          pyvbmc/entropy/entlb_vbmc.py:6-180
 *   vbmc_gp_pack      <- the gpyreg posterior fields read at
 *                                            pyvbmc/vbmc/variational_optimization.py:1311,1367-1398
 *   vbmc_gplogjoint   <- _gp_log_joint       pyvbmc/vbmc/variational_optimization.py:1238-1606
 *   vbmc_set_bounds   <- theta_bnd dict      pyvbmc/variational_posterior/variational_posterior.py:140-239
 *   vbmc_negelcbo     <- _neg_elcbo          pyvbmc/vbmc/variational_optimization.py:991-1235
 *                        (incl. _vp_bound_loss :503-606, _soft_bound_loss :609-657,
 *                         weight penalty :1212-1229)
 *
 * Conventions
 *   - plain pointers and sizes only; every array is host memory, fp64, C-contiguous,
 *     unless a parameter name ends in _dev (device pointer) -- those entry points are
 *     the split-phase / device-resident variants used for multi-GPU sharding and for
 *     kernel-only timing;
 *   - mu is COMPONENT-MAJOR: mu[k*D + d]  (== vp.mu.ravel(order="F"), the theta layout);
 *   - gradients come back in the reference's theta order [mu (D*K) | sigma (K) |
 *     lambda (D) | w (K)], only the groups whose grad_flags[i] != 0;
 *   - return value 0 = success, otherwise a VBMC_ERR_* code; vbmc_last_error() gives
 *     the message.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with VBMC_ERR_CUDA.
 *   - all work of one call is enqueued on the context's stream and the call returns
 *     after the results have landed in the host buffers (the *_async variants return
 *     after enqueueing).
 */
#ifndef VBMC_B200_H
#define VBMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBMC_B200_ABI_VERSION 1

enum {
    VBMC_OK = 0,
    VBMC_ERR_CUDA = 1,        /* CUDA runtime error / no device */
    VBMC_ERR_ARG = 2,         /* invalid argument */
    VBMC_ERR_UNSUPPORTED = 3, /* maps to NotImplementedError on the Python side */
    VBMC_ERR_STATE = 4        /* e.g. GP not packed */
};

/* GP mean function (variational_optimization.py:1328-1330,1383-1392) */
enum { VBMC_MEAN_ZERO = 0, VBMC_MEAN_CONST = 1, VBMC_MEAN_NEGQUAD = 2 };

/* arithmetic of the Monte-Carlo entropy kernel */
enum {
    VBMC_PREC_F32 = 0, /* fp32 compute, fp64 accumulation (default) */
    VBMC_PREC_F64 = 1  /* all fp64 */
};

/* source of the standard-normal draws of entmc */
enum {
    VBMC_RNG_EPS = 0,   /* caller supplies eps[K][Ns/2][D] (parity mode; entmc_vbmc.py:64-68) */
    VBMC_RNG_PHILOX = 1 /* device counter-based Philox4x32-10 + Box-Muller keyed (seed, j, i, d) */
};

typedef struct vbmc_ctx vbmc_ctx;

/* ---- lifetime ------------------------------------------------------------------ */
int vbmc_abi_version(void);
const char *vbmc_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing); never fails */
int vbmc_device_count(void);
int vbmc_ctx_create(int device, vbmc_ctx **out);
void vbmc_ctx_destroy(vbmc_ctx *ctx);
/* stream all kernels/copies of this context are enqueued on (a cudaStream_t) */
void *vbmc_ctx_stream(vbmc_ctx *ctx);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
int64_t vbmc_ctx_launch_count(vbmc_ctx *ctx);

/* ---- variational posterior parameters -------------------------------------------
 * All compute entry points take the mixture as it stands AFTER
 * VariationalPosterior.set_parameters (variational_posterior.py:680-759):            */
typedef struct vbmc_vp {
    int D, K;
    const double *mu;    /* [K*D] component-major */
    const double *sigma; /* [K] */
    const double *lambd; /* [D] */
    const double *w;     /* [K] */
    const double *eta;   /* [K] (softmax pre-activations; used by the w Jacobian) */
} vbmc_vp;

/* ---- entropy --------------------------------------------------------------------- */
/* entmc_vbmc(vp, Ns, grad_flags, jacobian_flag) -> (H, dH).  Ns = draws per component,
 * rounded up to even (entmc_vbmc.py:61).  rng_mode VBMC_RNG_EPS: eps = [K][Ns_even/2][D]
 * host doubles, mirrored antithetically on the device.  VBMC_RNG_PHILOX: eps ignored,
 * (seed, offset) key the draws.  dH has room for D*K+K+D+K doubles.                     */
int vbmc_entmc(vbmc_ctx *ctx, const vbmc_vp *vp, int64_t Ns, const int grad_flags[4],
               int jacobian_flag, int rng_mode, const double *eps, uint64_t seed,
               uint64_t offset, int precision, double *H, double *dH);

/* entlb_vbmc(vp, grad_flags, jacobian_flag) -> (H, dH) */
int vbmc_entlb(vbmc_ctx *ctx, const vbmc_vp *vp, const int grad_flags[4], int jacobian_flag,
               double *H, double *dH);

/* debugging / test hook: the eps[K][Ns_even/2][D] that VBMC_RNG_PHILOX uses */
int vbmc_philox_normals(vbmc_ctx *ctx, int D, int K, int64_t Ns, uint64_t seed,
                        uint64_t offset, double *eps_out);

/* ---- GP surrogate ---------------------------------------------------------------- */
/* Upload a trained GP: X[N][D]; hyp[S][H] in gpyreg order [ln ell (D), ln sf, noise
 * (noise_N), m0, xm (D), ln omega (D)] with cov_N = D+1; alpha[S][N]; L[S][N][N] (may be
 * NULL if the variance path is never used), L_chol[S], sn2_eff[S] = 1/sW[0]^2.           */
int vbmc_gp_pack(vbmc_ctx *ctx, int D, int N, int S, const double *X, const double *hyp,
                 int H, const double *alpha, const double *L, const int *L_chol,
                 const double *sn2_eff, int mean_kind, int cov_N, int noise_N);

/* _gp_log_joint(vp, gp, grad_flags, avg_flag, jacobian_flag, compute_var, separate_K).
 * Outputs (any may be NULL when not wanted):
 *   G      [1] if avg_flag and S>1 (or S==1), else [S]
 *   dG     [P] (averaged) or [P][S] (avg_flag == 0), P from grad_flags
 *   varG   [1] or [S] (compute_var != 0),  var_ss [1]
 *   I_sk   [S][K], J_sjk [S][K][K] (separate_K)                                          */
int vbmc_gplogjoint(vbmc_ctx *ctx, const vbmc_vp *vp, const int grad_flags[4], int avg_flag,
                    int jacobian_flag, int compute_var, int separate_K, double *G, double *dG,
                    double *varG, double *var_ss, double *I_sk, double *J_sjk);

/* ---- negative ELCBO -------------------------------------------------------------- */
/* Soft bounds (theta_bnd of VariationalPosterior.get_bounds).  n = length of lb/ub =
 * [mu (D*K) if optimize_mu | ln-scale (D*K) | eta (K) if optimize_weights].  Pass n = 0
 * to clear (theta_bnd=None).                                                             */
int vbmc_set_bounds(vbmc_ctx *ctx, int n, const double *lb, const double *ub, double tol_con,
                    double weight_threshold, double weight_penalty);

typedef struct vbmc_elcbo_in {
    vbmc_vp vp;          /* state after vp.set_parameters(theta) and the eta shift (:1080-1085) */
    int optimize[4];     /* vp.optimize_{mu,sigma,lambd,weights} */
    /* inputs of _vp_bound_loss taken from theta itself (:536-555); ignored without bounds */
    const double *ln_sigma_b; /* [K] theta's ln sigma block, or log(vp.sigma)  */
    const double *ln_lambd_b; /* [D] theta's ln lambda block, or log(vp.lambd) */
    const double *eta_b;      /* [K] theta[-K:] (unshifted), NULL if !optimize_weights */
    int64_t Ns;          /* draws per component; 0 -> entlb (:1163-1168) */
    int compute_grad;
    int compute_var;     /* 0/1 */
    int separate_K;
    int use_bounds;      /* theta_bnd is not None */
    int rng_mode;
    const double *eps;
    uint64_t seed, offset;
    int precision;
} vbmc_elcbo_in;

typedef struct vbmc_elcbo_out {
    double F, G, H, varF, varG_ss;
    double *dF;    /* [P] or NULL */
    double *dH;    /* [P] or NULL */
    double *I_sk;  /* [S][K] or NULL */
    double *J_sjk; /* [S][K][K] or NULL */
} vbmc_elcbo_out;

int vbmc_negelcbo(vbmc_ctx *ctx, const vbmc_elcbo_in *in, vbmc_elcbo_out *out);

/* Same evaluation (value + gradient, no variance path) with ONE packed input and ONE packed output --
 * the low-overhead form the Python shim uses inside the minimize_adam loop (:238-249).
 *   params = [mu (K*D, component-major) | sigma (K) | lambd (D) | w (K) | eta (K) |
 *             ln_sigma_b (K) | ln_lambd_b (D) | eta_b (K)]           (D*K + 5K + 2D doubles)
 *   out    = [F, G, H, 0, 0, L_bound, L_penalty, nonfinite | dF (P) | dH (P) if want_dH]       */
int vbmc_negelcbo_flat(vbmc_ctx *ctx, int D, int K, const double *params, const int optimize[4], int64_t Ns,
                       int compute_grad, int use_bounds, int rng_mode, const double *eps, uint64_t seed,
                       uint64_t offset, int precision, int want_dH, double *out);

/* Same evaluation with the RAW optimiser vector in: VariationalPosterior.set_parameters(theta) (variational_posterior.py:
 * 680-759: exp, lambda normalisation, softmax weights), the eta shift (variational_optimization.py:1082-1085) and the
 * bound-loss inputs (:536-555) run on the device, so the host does no O(P) arithmetic per call.
 *   theta  [P]  in/out: P = D*K, K, D, K summed over the optimised groups; on return its eta block is shifted so that
 *               max(eta) == 0 -- the reference's own in-place side effect on the caller's array
 *   tmpl        parameter block (vbmc_param_len doubles, layout as vbmc_negelcbo_flat) supplying the groups theta does
 *               not carry; may be NULL when all four groups are optimised
 *   out         [F, G, H, 0, 0, L_bound, L_penalty, nonfinite | dF (P)]
 *   vp_out      [sigma (K) | lambda (D) | w (K)] after normalisation: what set_parameters leaves in the posterior
 * Draws: VBMC_RNG_PHILOX keyed (seed, offset).                                                                    */
int vbmc_negelcbo_theta(vbmc_ctx *ctx, int D, int K, double *theta, const double *tmpl, const int optimize[4],
                        int64_t Ns, int compute_grad, int use_bounds, uint64_t seed, uint64_t offset, int precision,
                        double *out, double *vp_out);
/* Start generating the Monte-Carlo noise of the evaluation with Philox key (seed, offset) NOW, on a side stream: the
 * noise does not depend on theta, so the generator runs while the host is still preparing the call.  A following
 * vbmc_negelcbo_theta / vbmc_negelcbo_flat with the same (D, K, Ns, seed, offset) uses it; anything else ignores it.
 * A no-op for problem sizes that do not take the tensor-core entropy kernel.  OPTIONAL: without it the evaluation forks
 * the generator at the root of its own CUDA graph, beside the parameter kernel (the default of the Python shim).    */
int vbmc_noise_prefetch(vbmc_ctx *ctx, int D, int K, int64_t Ns, uint64_t seed, uint64_t offset);

/* ---- split-phase, device-resident variants ----------------------------------------
 * One evaluation = partials (this rank's shard of draws and hyper-samples, raw sums
 * BEFORE the sigma/lambda/softmax Jacobians, already scaled by the GLOBAL 1/Ns and 1/S)
 * -> [all-reduce SUM of raw_dev over ranks, e.g. NCCL] -> finalize (identical on every
 * rank).  raw_dev holds vbmc_raw_len(D, K) doubles.  Results stay on the device in
 * out_dev: [F, G, H, varF, varG_ss, 0, 0, 0, dF (P)...].  Nothing is synchronised.
 * Philox mode: the first vbmc_negelcbo_partials_async after an upload evaluates the uploaded key (seed, offset);
 * every further call WITHOUT a new upload is the next evaluation of a device-resident loop and uses
 * (seed, offset + 1), (seed, offset + 2), ... -- its noise tiles are generated beside the previous call's tail.    */
/* ---- batched sieve evaluation -------------------------------------------------------------
 * Replaces the loop of variational_optimization.py:775-787 (`_sieve`): for b < B
 *     F_b, _, G_b, H_b, _ = _neg_elcbo(theta_b, gp, vp_b, 0, Ns = 0, compute_grad = 0, compute_var = 0, theta_bnd)
 * (deterministic entropy bound entlb_vbmc, expected log joint, soft bounds, weight penalty; values only) in ONE
 * launch.  params: B parameter blocks of vbmc_param_len(D, K) doubles each, laid out
 *     [mu (K*D, component-major = theta order) | sigma (K) | lambda (D) | w (K) | eta (K) |
 *      ln sigma as in theta (K) | ln lambda as in theta (D) | eta as in theta (K)]      (the last three feed the bounds)
 * after VariationalPosterior.set_parameters' normalisation; bounds as set by vbmc_set_bounds.
 * out: [B][4] = F, G, H, L_bound + L_penalty.  Synchronous.                                                  */
size_t vbmc_param_len(int D, int K);
int vbmc_negelcbo_batch(vbmc_ctx *ctx, int B, int D, int K, const double *params, const int optimize[4],
                        int use_bounds, double *out);

/* ---- acquisition-function ingredients ------------------------------------------------------
 * Replaces, for the packed GP (vbmc_gp_pack WITH the factor L),
 *     f_mu, f_s2 = gp.predict(x_star = Xs, separate_samples = True)
 * (pyvbmc/acquisition_functions/abstract_acq_fcn.py:79; gpyreg's GP.predict: SE-ARD cross-covariance against the N
 * training points, predictive mean m(x*) + k* . alpha and latent variance sf2 - |L^-T (sW k*)|^2 (Cholesky branch)
 * or sf2 + k* . (L k*) (low-noise branch), clamped at 0, per hyper-sample).  Xs: [Nx][D] row-major host array;
 * f_mu, f_s2: [Nx][S] row-major host arrays.  fp64; synchronous.                                               */
int vbmc_gp_predict(vbmc_ctx *ctx, int Nx, const double *Xs, double *f_mu, double *f_s2);
/* bench aid: average device time (ms) of the prediction kernel on the points uploaded by the last vbmc_gp_predict */
int vbmc_gp_predict_device_ms(vbmc_ctx *ctx, int Nx, int reps, double *ms);
/* Replaces VariationalPosterior.pdf(x, orig_flag = False, log_flag, grad_flag) for df = inf
 * (pyvbmc/variational_posterior/variational_posterior.py:447-468,524-533; called by the acquisition functions,
 * acq_fcn_log.py:38-47): y [Nx] = pdf or log pdf (-inf where the pdf underflows to 0), dy [Nx][D] = gradient of the
 * pdf, or of the log pdf when log_flag (NULL unless grad_flag).  Transformed space; needs no packed GP.          */
int vbmc_vp_pdf(vbmc_ctx *ctx, const vbmc_vp *vp, int Nx, const double *Xs, int log_flag, int grad_flag, double *y,
                double *dy);

/* ---- device-resident Adam -----------------------------------------------------------------
 * minimize_adam(f, x0, lb, ub, tol_fun, max_iter, master_min, master_max, master_decay)
 * (pyvbmc/vbmc/minimize_adam.py:61-145) with f = the closure of variational_optimization.py:238-249,
 *     f(theta) = _neg_elcbo(theta, gp, vp, 0, Ns, compute_grad = True, compute_var = False, theta_bnd)[0:2].
 * theta, the moment estimates and the iterate table stay on the device; one iteration (theta -> parameters,
 * evaluation, Adam update, clamp) is a CUDA graph replayed without host synchronisation.  The early-stopping test
 * (:106-138) stays with the caller: run batches of 20 steps and inspect y / x.
 * Draws: VBMC_RNG_PHILOX keyed (seed, offset + iteration).  The fp32 -> fp64 overflow fallback of
 * vbmc_negelcbo_flat does not exist here: a non-finite y must make the caller fall back to the host loop.      */
typedef struct vbmc_adam_in {
    int D, K;
    const double *params;  /* parameter block (vbmc_param_len doubles, layout as vbmc_negelcbo_batch) of the CURRENT
                              vp: supplies the groups theta does not carry (optimize[i] == 0)                  */
    const double *theta0;  /* [P] start point, P = D*K, K, D, K summed over the optimised groups               */
    int optimize[4];
    int64_t Ns;            /* draws per component (> 0) */
    int use_bounds;        /* soft bounds as set by vbmc_set_bounds */
    uint64_t seed, offset;
    const double *lb, *ub; /* [P] hard box of minimize_adam, or NULL */
    int max_iter;
    double master_min, master_max, master_decay;
    int precision;
} vbmc_adam_in;
int vbmc_adam_init(vbmc_ctx *ctx, const vbmc_adam_in *in);
/* run the next n iterations; y[n] = objective values f(x) seen by them, x[n][P] = iterates AFTER each update
 * (row i = x_tab[:, i] of the reference).  Synchronous.                                                       */
int vbmc_adam_steps(vbmc_ctx *ctx, int n, double *y, double *x);

/* Split-phase form of vbmc_adam_steps, for a host loop that keeps one batch in flight while it evaluates the
 * early-stopping rule (minimize_adam.py:106-138) on the previous one: vbmc_adam_enqueue issues n iterations and returns
 * without synchronising; vbmc_adam_fetch waits for iterations [i0, i0 + n) ONLY (not for what was issued behind them)
 * and copies their objective values y[n] and iterates x[n][P] to the host. */
int vbmc_adam_enqueue(vbmc_ctx *ctx, int n);
int vbmc_adam_fetch(vbmc_ctx *ctx, int64_t i0, int n, double *y, double *x);

size_t vbmc_raw_len(int D, int K);
size_t vbmc_out_len(int D, int K);
int vbmc_negelcbo_upload(vbmc_ctx *ctx, const vbmc_elcbo_in *in);
/* world == 1: the raw phases are deferred and run fused with the following vbmc_negelcbo_finalize_async on the same
 * raw_dev (one cluster kernel for the whole tail), i.e. raw_dev is complete only after that call is enqueued.
 * world > 1: raw_dev holds this rank's partial raw vector when the call's work completes (all-reduce it, then finalize). */
int vbmc_negelcbo_partials_async(vbmc_ctx *ctx, int rank, int world, double *raw_dev);
int vbmc_negelcbo_finalize_async(vbmc_ctx *ctx, const double *raw_dev, double *out_dev);
int vbmc_stream_synchronize(vbmc_ctx *ctx);
/* Re-enqueue the evaluation staged last (by vbmc_negelcbo_flat or vbmc_negelcbo_upload) on the context stream:
 * the kernels of one evaluation with parameters and GP already resident in HBM, result left in the library's device
 * out buffer, no host synchronisation.  This is the device-resident step bench.py times with CUDA events (`value`):
 * what one `_neg_elcbo` call (variational_optimization.py:991-1235) costs the GPU, without the host round trip.
 * Every call is a NEW evaluation: Philox key (seed, offset + number of calls since staging), as in consecutive
 * iterations of the Adam loop.                                                                                   */
int vbmc_negelcbo_enqueue(vbmc_ctx *ctx);

/* ---- all-reduce of the raw vector over NVLink peer memory (optional; replaces the NCCL all-reduce between
 * vbmc_negelcbo_partials_async and vbmc_negelcbo_finalize_async).  One process per GPU on ONE node:
 *   1. every rank: vbmc_p2p_export(ctx, world, D, K, handle)  -- allocates this rank's exchange buffer (sized for
 *      raw vectors up to vbmc_raw_len(D, K)) and returns its 64-byte CUDA IPC handle;
 *   2. all-gather the handles by any host-side means (e.g. torch.distributed.all_gather_object);
 *   3. every rank: vbmc_p2p_open(ctx, rank, world, handles[world][64]).
 * From then on vbmc_negelcbo_partials_async(rank, world, raw) defers its reduce stage (as with world == 1) and
 * vbmc_negelcbo_finalize_async runs raw phases -> peer stores + flags -> fixed-order sum -> final phase in ONE
 * cluster kernel; raw holds the all-reduced vector afterwards.  Every rank must issue the same sequence of
 * evaluations.  A peer that does not answer within ~2 s poisons the result (F = NaN, out[7] = 2) instead of
 * hanging the GPU.  vbmc_p2p_close (or destroying the context) unmaps the peers.                              */
#define VBMC_P2P_MAX_WORLD 8
#define VBMC_P2P_HANDLE_BYTES 64
int vbmc_p2p_export(vbmc_ctx *ctx, int world, int D, int K, unsigned char *handle);
int vbmc_p2p_open(vbmc_ctx *ctx, int rank, int world, const unsigned char *handles);
/* vbmc_p2p_unmap only drops the mappings of the PEERS' buffers (the local buffer stays allocated): tear-down order
 * is every rank unmaps -> host barrier -> every rank closes, so that no buffer is freed while a peer still maps it. */
int vbmc_p2p_unmap(vbmc_ctx *ctx);
int vbmc_p2p_close(vbmc_ctx *ctx);
/* copy n doubles from device memory to the host through the context's pinned staging buffer,
 * ordered after everything enqueued on the context stream (synchronises)                     */
int vbmc_read_device(vbmc_ctx *ctx, const double *src_dev, size_t n, double *dst_host);
/* Kernel timing for bench.py's roofline: when enabled, every entmc launch is bracketed by
 * CUDA events on the context stream (and synchronised -- measurement mode only);
 * vbmc_entmc_kernel_ms returns the average device time (ms) since the last call.        */
int vbmc_set_kernel_timing(vbmc_ctx *ctx, int on);
int vbmc_entmc_kernel_ms(vbmc_ctx *ctx, double *avg_ms, int64_t *launches);
/* ... and of the dominant kernel alone (entmc_kernel_tc, without the table kernel launched in front of it); call it
 * BEFORE vbmc_entmc_kernel_ms, which resets the counters.  0 when another kernel variant ran.                     */
int vbmc_entmc_main_kernel_ms(vbmc_ctx *ctx, double *avg_ms);
/* which fp32 entmc kernel the last staged evaluation used: 0 expanded ("fast"), 1 dimension-split, 2 packed,
 * 3 scalar (also the all-fp64 kernel), 4 warp-autonomous, 5 tensor-core (tcgen05 / TMEM); -1 nothing planned yet */
int vbmc_entmc_variant_used(vbmc_ctx *ctx);
/* measurement only (VBMC_STAGE_TIMING=1 in the environment when the context is created): device time in
 * microseconds between the stage marks of the last synchronous evaluation --
 * us[0] H2D, [1] entmc (+ launch of gplj), [2] wait for the side stream, [3] reduce, [4] finalize,
 * [5] D2H, and us[6] = host wall time of the whole call                                          */
int vbmc_stage_times(vbmc_ctx *ctx, double *us);
/* measurement only: best-of-5 FMA issue peak of this device in TFLOP/s
 * (fp64: 0 -> FFMA, 1 -> DFMA, 2 -> packed FFMA2)                                         */
int vbmc_fma_peak(vbmc_ctx *ctx, int fp64, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* VBMC_B200_H */
