// capi.cu -- extern "C" entry points declared in include/vbmc_b200.h: context, GP pack and the
// orchestration of one evaluation  upload -> partials -> [all-reduce] -> finalize -> download.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace vbmc {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

// Bumped whenever ANY context's device buffer moves (never reset).  Every context remembers the epoch it last looked
// at (CtxEx::seen_epoch): a captured graph is only replayed if no buffer anywhere moved since it was captured.  The
// counter is process-wide and monotonic on purpose -- a consumed-once flag shared by all contexts let context B (or
// another thread) swallow the notification meant for context A, which then replayed freed pointers.  A reallocation
// in one context costs the others one re-capture; it can never be missed.
static std::atomic<uint64_t> g_realloc_epoch{1};

int ensure(double **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return VBMC_OK;
    g_realloc_epoch.fetch_add(1, std::memory_order_relaxed);
    if (*p) VBMC_CUDA_CHECK(cudaFree(*p));
    *p = nullptr;
    size_t n = need + need / 4 + 64;
    VBMC_CUDA_CHECK(cudaMalloc((void **)p, n * sizeof(double)));
    *cap = n;
    return VBMC_OK;
}

int ensure_pinned(double **d, double **h, size_t *cap, size_t need) {
    if (need <= *cap && *d && *h) return VBMC_OK;
    g_realloc_epoch.fetch_add(1, std::memory_order_relaxed);
    if (*d) VBMC_CUDA_CHECK(cudaFree(*d));
    if (*h) VBMC_CUDA_CHECK(cudaFreeHost(*h));
    *d = *h = nullptr;
    size_t n = need + need / 4 + 64;
    VBMC_CUDA_CHECK(cudaMalloc((void **)d, n * sizeof(double)));
    VBMC_CUDA_CHECK(cudaMallocHost((void **)h, n * sizeof(double)));
    *cap = n;
    return VBMC_OK;
}

int ensure_host_pinned(double **h, size_t *cap, size_t need) {
    if (need <= *cap && *h) return VBMC_OK;
    g_realloc_epoch.fetch_add(1, std::memory_order_relaxed);
    if (*h) VBMC_CUDA_CHECK(cudaFreeHost(*h));
    *h = nullptr;
    size_t n = need + need / 4 + 64;
    VBMC_CUDA_CHECK(cudaMallocHost((void **)h, n * sizeof(double)));
    *cap = n;
    return VBMC_OK;
}

namespace {

struct Bind {
    explicit Bind(Ctx *c) { cudaSetDevice(c->device); }
};

int check_vp(const vbmc_vp *vp, bool flat) {
    VBMC_REQUIRE(vp && (flat || (vp->mu && vp->sigma && vp->lambd && vp->w && vp->eta)), VBMC_ERR_ARG, "vp: null field");
    VBMC_REQUIRE(vp->D >= 1 && vp->K >= 1, VBMC_ERR_ARG, "vp: D and K must be >= 1");
    VBMC_REQUIRE(vp->D <= kMaxD, VBMC_ERR_UNSUPPORTED, "D > 32 is not supported by the CUDA path");
    return VBMC_OK;
}

int64_t even_ns(int64_t Ns) { return ((Ns + 1) / 2) * 2; }

int packed_len(int D, int K, const int g[4]) {
    return (g[0] ? D * K : 0) + (g[1] ? K : 0) + (g[2] ? D : 0) + (g[3] ? K : 0);
}

// internal description of one evaluation
struct Spec {
    // theta-in fast path (vbmc_negelcbo_theta): the raw optimiser vector; VariationalPosterior.set_parameters runs on
    // the device.  tmpl = parameter block supplying the groups theta does not carry (may be null when all are optimised)
    const double *theta = nullptr, *tmpl = nullptr;
    int P = 0;
    const double *flat = nullptr;  // whole parameter block in ParamLayout order (fast path), else vp + *_b
    vbmc_vp vp;
    int grad[4];
    int jacobian = 1;
    int optimize[4] = {1, 1, 1, 1};
    const double *ln_sigma_b = nullptr, *ln_lambd_b = nullptr, *eta_b = nullptr;
    int64_t Ns = 0;
    bool have_ent = true, have_gp = true;
    bool use_bounds = false;
    int rng_mode = VBMC_RNG_EPS;
    const double *eps = nullptr;
    uint64_t seed = 0, offset = 0;
    int precision = VBMC_PREC_F32;
    bool compute_var = false;
    int avg = 1;
    bool parts = true;  // dH / dG wanted besides dF
};

struct Staged {
    Spec s;
    int D = 0, DP = 0, K = 0;
    EvalFlags f{};
    EntmcPlan plan{};
    bool planned = false;
    EvalFlags f_partials{};
    int launches_per_eval = 0;
};

// per-context staged state (kept outside Ctx to keep common.cuh light)
// signature of the launch sequence of one flat evaluation (everything that is baked into a captured graph)
struct GraphKey {
    int D = 0, K = 0, flags = 0, n_bnd = 0, precision = 0, variant = 0, S = 0, N = 0;
    int64_t Ns = -1;
    uint64_t gen = 0;  // bumped by gp_pack / set_bounds / any reallocation
    bool operator==(const GraphKey &o) const {
        return D == o.D && K == o.K && flags == o.flags && n_bnd == o.n_bnd && precision == o.precision &&
               variant == o.variant && S == o.S && N == o.N && Ns == o.Ns && gen == o.gen;
    }
};

struct CtxEx {
    Ctx c;
    Staged st;
    std::vector<double> host_tmp;
    // CUDA-graph replay of the flat evaluation
    bool graphs_on = true;
    uint64_t gen = 1;
    uint64_t seen_epoch = 0;  // g_realloc_epoch when this context last looked (see there)
    GraphKey gkey, last_key;
    cudaGraphExec_t gexec = nullptr;
    cudaGraph_t graph = nullptr;
    int64_t graph_launches = 0;
    // device-resident Adam (vbmc_adam_init / vbmc_adam_steps)
    bool adam_ready = false, adam_eager_done = false;
    AdamDev adam{};
    double *d_adam = nullptr;  // one allocation: theta, m, v, tmpl, lb, ub, ytab, xtab, iter
    size_t adam_cap = 0;
    int adam_max_iter = 0, adam_launches_per_iter = 0;
    long long adam_done = 0;
    cudaGraphExec_t adam_gexec = nullptr;
    cudaGraph_t adam_graph = nullptr;
    uint64_t adam_gen = 0;
    int64_t partials_calls = 0;     // vbmc_negelcbo_partials_async calls since the last upload
    long long adam_issued = 0;      // Adam iterations issued (eager + captured) since vbmc_adam_init
    bool adam_params_ready = false; // d_in holds the parameter block of the NEXT Adam iteration (fused update kernel)
    static constexpr int kAdamTickets = 8;
    struct AdamTicket {
        cudaEvent_t ev = nullptr;
        long long i_end = 0;  // iterations [.., i_end) are complete once `ev` has fired
    } adam_tickets[kAdamTickets];
    int adam_ticket_next = 0;
    bool adam_graph_params_ready = false;
    int adam_graph_buf = 0;         // noise-tile buffer parity / look-ahead state the captured pair starts from
    bool adam_graph_ready = false;
};

CtxEx *ex(vbmc_ctx *p) { return reinterpret_cast<CtxEx *>(p); }

int stage(CtxEx *x, const Spec &s, bool single_gpu = false) {
    Ctx *c = &x->c;
    VBMC_TRY(check_vp(&s.vp, s.flat != nullptr || s.theta != nullptr));
    const int D = s.vp.D, K = s.vp.K, DP = pad_dim(D);
    if (s.have_gp) {
        VBMC_REQUIRE(c->has_gp, VBMC_ERR_STATE, "no GP packed (call vbmc_gp_pack first)");
        VBMC_REQUIRE(c->gD == D, VBMC_ERR_ARG, "vp.D does not match the packed GP");
    }
    ParamLayout lay{D, DP, K};
    RawLayout rl{D, K};
    VBMC_TRY(ensure_pinned(&c->d_in, &c->h_in, &c->in_cap, (size_t)lay.total() + 2));
    double *h = c->h_in;
    // Monte-Carlo draws not generated ahead for this key (single-GPU evaluation through run_flat / run_single): the tile
    // generator depends on the key only, so it is forked off HERE, in front of the parameter kernel, on a third stream
    // (launched below, once the key sits in the pinned block); entmc_tc.cu joins it in front of the main kernel and
    // partials() in any case.  It must be enqueued BEFORE the kernels that wait for the parameter kernel: a captured
    // graph submits its kernels in creation order, and a launch behind such a wait does not start before it either.
    c->root_forked = false;
    static const bool root_fork_on = getenv("VBMC_ROOT_FORK") ? atoi(getenv("VBMC_ROOT_FORK")) != 0 : true;
    EntmcPlan root_plan{};
    bool root_gen = false;
    if (root_fork_on && single_gpu && s.have_ent && s.Ns > 0 && s.rng_mode == VBMC_RNG_PHILOX && s.precision != VBMC_PREC_F64 &&
        !(c->noise_ready && c->noise_seed == s.seed && c->noise_offset == s.offset)) {
        const int64_t half = even_ns(s.Ns) / 2;
        if (entmc_plan(c, D, K, half, s.grad[3] != 0, s.precision, &root_plan) == VBMC_OK && root_plan.variant == ENTMC_TC) {
            root_plan.pair0 = 0, root_plan.half_glob = half;
            root_gen = true;
        }
    }
    if (root_gen) {
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_root, c->stream));
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream3, c->ev_root, 0));
        c->root_forked = true;
    }
    // (enqueued in FRONT of the parameter kernel.  Measured alternatives, profiles/r4_e2e_timeline.md: behind it, the
    // generator's branch starts ~4 us late and its single wave then keeps the table kernel waiting for SM resources;
    // whichever of the two root branches of the captured graph is created second starts 4-8 us after the first.)
    auto launch_root_gen = [&]() -> int {
        return root_gen ? entmc_tc_prefetch(c, lay, root_plan, s.seed, s.offset, true) : VBMC_OK;
    };
    if (s.theta) {
        // theta, template and key go to PINNED host memory; one small kernel reads them through their device aliases
        // and writes the parameter block (set_parameters + eta shift + bound inputs on the device)
        const size_t T = (size_t)lay.total();
        VBMC_TRY(ensure_host_pinned(&c->h_theta, &c->theta_cap, (size_t)s.P + T + 2 + (size_t)(2 * K + D)));
        double *th = c->h_theta, *tm = th + s.P, *key = tm + T, *vpo = key + 2;
        memcpy(th, s.theta, sizeof(double) * s.P);
        if (s.tmpl) memcpy(tm, s.tmpl, sizeof(double) * T);
        memcpy(key, &s.seed, sizeof(uint64_t));
        memcpy(key + 1, &s.offset, sizeof(uint64_t));
        c->key_host = key;
        VBMC_TRY(launch_root_gen());
        AdamDev a{};
        a.lay = lay, a.P = s.P;
        for (int i = 0; i < 4; ++i) a.opt[i] = s.optimize[i];
        a.theta = th, a.tmpl = tm;
        VBMC_TRY(theta_prepare_launch(c, a, c->d_in, vpo, reinterpret_cast<const uint64_t *>(key)));
    } else {
    if (s.flat) {
        memcpy(h, s.flat, sizeof(double) * lay.total());
    } else {
        memcpy(h + lay.mu(), s.vp.mu, sizeof(double) * K * D);
        memcpy(h + lay.sigma(), s.vp.sigma, sizeof(double) * K);
        memcpy(h + lay.lambd(), s.vp.lambd, sizeof(double) * D);
        memcpy(h + lay.w(), s.vp.w, sizeof(double) * K);
        memcpy(h + lay.eta(), s.vp.eta, sizeof(double) * K);
        for (int k = 0; k < K; ++k) h[lay.lnsig_b() + k] = s.ln_sigma_b ? s.ln_sigma_b[k] : log(s.vp.sigma[k]);
        for (int d = 0; d < D; ++d) h[lay.lnlam_b() + d] = s.ln_lambd_b ? s.ln_lambd_b[d] : log(s.vp.lambd[d]);
        for (int k = 0; k < K; ++k) h[lay.eta_b() + k] = s.eta_b ? s.eta_b[k] : s.vp.eta[k];
    }
    memcpy(h + lay.total(), &s.seed, sizeof(uint64_t));  // Philox key rides behind the parameter block
    memcpy(h + lay.total() + 1, &s.offset, sizeof(uint64_t));
    c->key_host = h + lay.total();
    VBMC_TRY(launch_root_gen());
    // parameter block -> HBM.  A one-CTA kernel that reads the pinned block through its device alias (UVA) instead
    // of a copy-engine memcpy: inside the captured graph a kernel node starts ~5 us sooner than a memcpy node, and
    // this copy heads the critical path of every evaluation.
    VBMC_TRY(stage_copy_launch(c, c->d_in, h, lay.total() + 2, c->stream));
    }

    if (s.use_bounds) {
        const int n_expect = (s.optimize[0] ? K * D : 0) + K * D + (s.optimize[3] ? K : 0);
        VBMC_REQUIRE(c->n_bnd == n_expect, VBMC_ERR_ARG, "soft bounds: lb/ub length does not match [mu|ln-scale|eta]");
    }
    VBMC_TRY(ensure(&c->d_raw, &c->raw_cap, (size_t)rl.total()));
    VBMC_TRY(ensure_pinned(&c->d_out, &c->h_out, &c->out_cap, vbmc_out_len(D, K)));
    if (s.have_gp) {
        VBMC_TRY(ensure(&c->d_lamc, &c->lamc_cap, (size_t)c->S * K * D));
        VBMC_TRY(ensure(&c->d_gps, &c->gps_cap, (size_t)c->S * (1 + rl.block())));
    }
    if (s.have_ent && s.Ns > 0 && s.rng_mode == VBMC_RNG_EPS) {
        VBMC_REQUIRE(s.eps != nullptr, VBMC_ERR_ARG, "entmc: eps required in VBMC_RNG_EPS mode");
        const size_t n = (size_t)K * (size_t)(even_ns(s.Ns) / 2) * (size_t)D;
        VBMC_TRY(ensure(&c->d_eps, &c->eps_cap, n));
        VBMC_CUDA_CHECK(cudaMemcpyAsync(c->d_eps, s.eps, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    Staged &st = x->st;
    st.s = s;
    st.D = D, st.DP = DP, st.K = K;
    EvalFlags &f = st.f;
    for (int i = 0; i < 4; ++i) f.grad[i] = s.grad[i], f.optimize[i] = s.optimize[i];
    f.jacobian = s.jacobian;
    f.use_ent_mc = s.Ns > 0;
    f.have_ent = s.have_ent, f.have_gp = s.have_gp;
    f.use_bounds = s.use_bounds;
    f.avg = 1;
    f.parts = s.parts;
    st.planned = false;
    c->staged = true;
    x->adam_params_ready = false;  // d_in no longer holds the block the Adam loop left for its next iteration
    // host mirror of the Philox key behind the parameter block (noise tiles generated ahead are matched against it)
    c->cur_seed = s.seed, c->cur_offset = s.offset;
    c->key_delta = 0;
    c->lookahead = false;
    return VBMC_OK;
}

// A vbmc_noise_prefetch launch runs on the side stream and is not yet ordered before the main stream.  Every entry
// point joins it HERE, outside any stream capture (a wait on an event recorded outside a capture cannot be captured),
// whether or not the prefetched tiles end up being used.
int settle_prefetch(Ctx *c) {
    if (c->noise_needs_wait) {
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_noise, 0));
        c->noise_needs_wait = false;
    }
    return VBMC_OK;
}

int partials(CtxEx *x, int rank, int world, double *raw_dev) {
    Ctx *c = &x->c;
    Staged &st = x->st;
    VBMC_REQUIRE(c->staged, VBMC_ERR_STATE, "nothing staged (call vbmc_negelcbo_upload first)");
    VBMC_REQUIRE(world >= 1 && rank >= 0 && rank < world, VBMC_ERR_ARG, "bad rank/world");
    const Spec &s = st.s;
    const int D = st.D, DP = st.DP, K = st.K;
    const bool anyg = s.grad[0] || s.grad[1] || s.grad[2] || s.grad[3];
    RawLayout rl{D, K};
    EvalFlags f = st.f;
    const EntmcPlan *planp = nullptr;
    int64_t Ns_glob = 0;
    // the fp64 log-joint kernel runs beside the fp32 entropy kernel on a side stream
    const bool fork = s.have_gp && s.have_ent && s.Ns > 0;
    if (s.have_gp) {
        if (fork) {
            VBMC_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
            VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        }
        double *Zout = nullptr;
        if (s.compute_var) {
            VBMC_REQUIRE(world == 1, VBMC_ERR_UNSUPPORTED, "the variance path is not sharded");
            VBMC_TRY(ensure(&c->d_var, &c->var_cap, gpvar_workspace(c->S, K, c->N)));
            Zout = gpvar_Z(c);
        }
        VBMC_TRY(gplj_launch(c, c->d_in, K, rank, world, anyg, fork ? c->stream2 : c->stream, Zout));
        if (fork) VBMC_CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
    }
    if (s.have_ent) {
        if (s.Ns > 0) {
            Ns_glob = even_ns(s.Ns);
            const int64_t half_glob = Ns_glob / 2;
            const int64_t p0 = half_glob * rank / world, p1 = half_glob * (rank + 1) / world;
            VBMC_TRY(entmc_plan(c, D, K, p1 - p0, s.grad[3] != 0, s.precision, &st.plan));
            st.plan.pair0 = p0;
            st.plan.half_glob = half_glob;
            const size_t n_rec = (st.plan.variant == ENTMC_WARP || st.plan.variant == ENTMC_TC) ? (size_t)st.plan.grid * st.plan.maxseg
                                                                : (size_t)K * st.plan.slabs;
            VBMC_TRY(ensure(&c->d_entpart, &c->entpart_cap, n_rec * entpart_stride(DP, K)));
            VBMC_TRY(entmc_launch(c, c->d_in, D, K, st.plan, anyg, s.grad[3] != 0, s.precision, s.rng_mode, c->d_eps,
                                  s.seed, s.offset, c->d_entpart));
            planp = &st.plan;
        } else if (rank == 0) {
            VBMC_TRY(entlb_launch(c, c->d_in, D, K, s.grad, raw_dev + rl.ent(), raw_dev));
        } else {
            f.have_ent = 0;
        }
    }
    if (c->root_forked) {  // forked in stage() and not used by the entropy kernel that ran: close the fork
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_root_join, c->stream3));
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_root_join, 0));
        c->root_forked = false;
    }
    stage_mark(c, 2);
    if (fork) VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    stage_mark(c, 3);
    st.f_partials = f;
    // single GPU: the raw phases are fused into the finalize launch (one cluster kernel for the whole tail); W ranks
    // with mapped peer buffers: the same, with the all-reduce over peer memory between the phases
    const bool p2p = world > 1 && c->p2p_world == world && c->p2p_rank == rank && rl.total() <= c->p2p_stride;
    VBMC_TRY(reduce_launch(c, c->d_in, D, K, f, planp, Ns_glob, rank, world, c->S, raw_dev, world == 1 || p2p));
    return VBMC_OK;
}

int finalize(CtxEx *x, const double *raw_dev, double *out_dev) {
    Ctx *c = &x->c;
    Staged &st = x->st;
    VBMC_REQUIRE(c->staged, VBMC_ERR_STATE, "nothing staged");
    VBMC_TRY(finalize_launch(c, c->d_in, st.D, st.K, st.f, raw_dev, out_dev));
    if (c->noise_pending_join) {  // side-stream generator of the next evaluation's draws (entmc_tc.cu): runs beside the tail
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_noise, 0));
        c->noise_pending_join = false;
    }
    // the variance path needs the per-sample G_s completed by the assemble phase
    if (st.s.compute_var) VBMC_TRY(gpvar_launch(c, c->d_in, st.K, st.s.avg));
    return VBMC_OK;
}

// run everything on one GPU and bring `n` leading doubles of out back to the host
int run_single(CtxEx *x, const Spec &s, size_t n_out) {
    Ctx *c = &x->c;
    VBMC_TRY(settle_prefetch(c));
    const auto t0 = std::chrono::steady_clock::now();
    stage_mark(c, 0);
    VBMC_TRY(stage(x, s, true));
    stage_mark(c, 1);  // after the H2D copies
    VBMC_TRY(partials(x, 0, 1, c->d_raw));
    stage_mark(c, 4);  // after the reduce stage (2 = entmc done, 3 = side stream joined)
    // zero-copy results: the final kernel writes straight into the pinned host buffer (UVA: pinned host
    // memory is device-addressable under the same pointer), so there is no D2H copy to wait for
    (void)n_out;
    VBMC_TRY(finalize(x, c->d_raw, c->h_out));
    stage_mark(c, 5);
    stage_mark(c, 6);
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    return VBMC_OK;
}

void drop_adam_graph(CtxEx *x) {
    if (x->adam_gexec) cudaGraphExecDestroy(x->adam_gexec);
    if (x->adam_graph) cudaGraphDestroy(x->adam_graph);
    x->adam_gexec = nullptr, x->adam_graph = nullptr;
}

// one Adam iteration on the context stream: theta -> parameter block, evaluation, update (no host sync)
int adam_iteration(CtxEx *x) {
    Ctx *c = &x->c;
    // the parameter block (and the key (seed, offset0 + iteration) behind it) of this iteration: left by the previous
    // iteration's fused update kernel, unless this is the first iteration or something else used the context in between
    if (!x->adam_params_ready) VBMC_TRY(adam_prepare_launch(c, x->adam, c->d_in));
    c->cur_seed = x->adam.seed, c->cur_offset = x->adam.offset0 + (uint64_t)x->adam_issued;
    x->adam_issued++;
    c->key_delta = 0;
    c->lookahead = true;  // the next iteration's draws are generated under this iteration's tail
    VBMC_TRY(partials(x, 0, 1, c->d_raw));
    VBMC_TRY(finalize(x, c->d_raw, c->d_out));
    VBMC_TRY(adam_update_prepare_launch(c, x->adam, c->d_out, c->d_in));
    x->adam_params_ready = true;
    return VBMC_OK;
}

void drop_graph(CtxEx *x) {
    if (x->gexec) cudaGraphExecDestroy(x->gexec);
    if (x->graph) cudaGraphDestroy(x->graph);
    x->gexec = nullptr, x->graph = nullptr;
    x->gkey = GraphKey{};
}

// Flat evaluation with CUDA-graph replay: the first call with a new launch signature runs eagerly (and
// sizes every buffer), the second one is captured (H2D copy, gplj on the side stream, entmc, raw, final),
// every later call only refreshes the pinned parameter block (theta + Philox key) and replays the graph:
// one driver call instead of ~12, and no CPU-side gaps between the dependent kernels.
int run_flat(CtxEx *x, const Spec &s, size_t n_out) {
    Ctx *c = &x->c;
    if (x->seen_epoch != g_realloc_epoch.load(std::memory_order_relaxed)) {  // some buffer moved since this context
        x->gen++;                                                           // last looked: captured pointers are suspect
        x->seen_epoch = g_realloc_epoch.load(std::memory_order_relaxed);
    }
    const bool eligible = x->graphs_on && !c->stage_timing && !c->time_entmc && (s.flat != nullptr || s.theta != nullptr) &&
                          !(s.Ns > 0 && s.rng_mode == VBMC_RNG_EPS) && !s.compute_var;
    if (!eligible) return run_single(x, s, n_out);
    GraphKey k;
    k.D = s.vp.D, k.K = s.vp.K, k.Ns = s.Ns, k.n_bnd = s.use_bounds ? c->n_bnd : 0, k.precision = s.precision;
    k.variant = c->entmc_variant, k.S = c->S, k.N = c->N, k.gen = x->gen;
    for (int i = 0; i < 4; ++i) k.flags |= (s.grad[i] ? 1 : 0) << i | (s.optimize[i] ? 1 : 0) << (4 + i);
    k.flags |= (s.use_bounds ? 1 : 0) << 8 | (s.have_gp ? 1 : 0) << 9 | (s.have_ent ? 1 : 0) << 10 | (s.parts ? 1 : 0) << 11;
    // noise tiles generated ahead by vbmc_noise_prefetch for exactly this key: the graph then holds no generator
    const bool pre = s.Ns > 0 && s.rng_mode == VBMC_RNG_PHILOX && c->noise_ready && c->noise_seed == s.seed &&
                     c->noise_offset == s.offset && c->noise_sig[0] == ((uint64_t)k.D << 32 | (uint64_t)k.K) &&
                     c->noise_sig[3] == (uint64_t)(even_ns(s.Ns) / 2);
    if (!pre) c->noise_ready = false;
    k.flags |= (s.theta ? 1 : 0) << 12 | (pre ? 1 : 0) << 13 | (c->noise_buf & 1) << 14 | (s.tmpl ? 1 : 0) << 15;
    VBMC_TRY(settle_prefetch(c));  // order the side-stream generator before everything this call enqueues
    if (x->gexec && k == x->gkey) {
        const ParamLayout lay{k.D, pad_dim(k.D), k.K};
        if (s.theta) {
            double *th = c->h_theta, *tm = th + s.P, *key = tm + lay.total();
            memcpy(th, s.theta, sizeof(double) * s.P);
            if (s.tmpl) memcpy(tm, s.tmpl, sizeof(double) * lay.total());
            memcpy(key, &s.seed, sizeof(uint64_t));
            memcpy(key + 1, &s.offset, sizeof(uint64_t));
        } else {
            memcpy(c->h_in, s.flat, sizeof(double) * lay.total());
            memcpy(c->h_in + lay.total(), &s.seed, sizeof(uint64_t));
            memcpy(c->h_in + lay.total() + 1, &s.offset, sizeof(uint64_t));
        }
        c->cur_seed = s.seed, c->cur_offset = s.offset;  // a new key behind the parameter block (see stage())
        c->key_delta = 0;
        c->lookahead = false;
        VBMC_CUDA_CHECK(cudaGraphLaunch(x->gexec, c->stream));
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        x->graph_launches++;
        c->launches += x->st.launches_per_eval;
        c->noise_ready = false;  // (a prefetched buffer is consumed by the replay)
        return VBMC_OK;
    }
    if (!(k == x->last_key)) {  // first sighting: eager run (allocations happen here)
        x->last_key = k;
        return run_single(x, s, n_out);
    }
    // second call with the same signature: capture
    drop_graph(x);
    const uint64_t epoch0 = g_realloc_epoch.load(std::memory_order_relaxed);
    const int64_t l0 = c->launches;
    VBMC_CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = stage(x, s, true);
    if (rc == VBMC_OK) rc = partials(x, 0, 1, c->d_raw);
    if (rc == VBMC_OK) rc = finalize(x, c->d_raw, c->h_out);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
    if (rc != VBMC_OK || ce != cudaSuccess || epoch0 != g_realloc_epoch.load(std::memory_order_relaxed)) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        x->gen++;  // whatever moved, start over with eager runs
        x->seen_epoch = g_realloc_epoch.load(std::memory_order_relaxed);
        x->last_key = GraphKey{};
        if (rc != VBMC_OK) return rc;
        return run_single(x, s, n_out);
    }
    x->graph = g;
    x->st.launches_per_eval = (int)(c->launches - l0);
    c->launches = l0;
    VBMC_CUDA_CHECK(cudaGraphInstantiate(&x->gexec, g, 0));
    x->gkey = k;
    VBMC_CUDA_CHECK(cudaGraphLaunch(x->gexec, c->stream));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    x->graph_launches++;
    c->launches += x->st.launches_per_eval;
    return VBMC_OK;
}

}  // namespace
}  // namespace vbmc

using namespace vbmc;

static void p2p_unmap(Ctx *c) {
    for (int q = 0; q < c->p2p_world; ++q)
        if (c->p2p_peer[q] && q != c->p2p_rank) cudaIpcCloseMemHandle(c->p2p_peer[q]);
    for (int q = 0; q < VBMC_P2P_MAX_WORLD; ++q) c->p2p_peer[q] = nullptr;
    c->p2p_world = 0;
}


extern "C" {

int vbmc_abi_version(void) { return VBMC_B200_ABI_VERSION; }
const char *vbmc_last_error(void) { return g_err.c_str(); }

int vbmc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vbmc_ctx_create(int device, vbmc_ctx **out) {
    VBMC_REQUIRE(out != nullptr, VBMC_ERR_ARG, "ctx_create: null out");
    *out = nullptr;
    int n = 0;
    VBMC_CUDA_CHECK(cudaGetDeviceCount(&n));
    VBMC_REQUIRE(n > 0, VBMC_ERR_CUDA, "no CUDA device visible (vbmc_b200 has no CPU fallback)");
    VBMC_REQUIRE(device >= 0 && device < n, VBMC_ERR_ARG, "ctx_create: bad device index");
    VBMC_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    VBMC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    VBMC_REQUIRE(prop.major >= 10, VBMC_ERR_CUDA, "vbmc_b200 kernels are built for sm_100a (B200) only");
    CtxEx *x = new CtxEx();
    x->c.device = device;
    x->c.sm_count = prop.multiProcessorCount;
    if (const char *gq = getenv("VBMC_GRAPH")) x->graphs_on = atoi(gq) != 0;
    if (const char *t = getenv("VBMC_STAGE_TIMING")) x->c.stage_timing = atoi(t) != 0;
    if (x->c.stage_timing)
        for (int i = 0; i < 8; ++i) VBMC_CUDA_CHECK(cudaEventCreate(&x->c.sev[i]));
    if (const char *v = getenv("VBMC_ENTMC_VARIANT")) x->c.entmc_variant = atoi(v);
    if (const char *g = getenv("VBMC_ENTMC_GUARD")) x->c.entmc_guard = (float)atof(g);
    // the main stream outranks the side stream: when the tail (a cluster of 8 x 1024 threads) and the look-ahead noise
    // generator (hundreds of CTAs) become runnable together behind the entropy kernel, the tail's cluster must get its
    // 8 SMs first -- otherwise it waits for the generator to drain and nothing overlaps
    int prio_least = 0, prio_greatest = 0;
    VBMC_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    VBMC_CUDA_CHECK(cudaStreamCreateWithPriority(&x->c.stream, cudaStreamNonBlocking, prio_greatest));
    VBMC_CUDA_CHECK(cudaStreamCreateWithPriority(&x->c.stream2, cudaStreamNonBlocking, prio_least));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_fork, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_join, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_main, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_noise, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaStreamCreateWithPriority(&x->c.stream3, cudaStreamNonBlocking, prio_least));
    VBMC_CUDA_CHECK(cudaMalloc((void **)&x->c.d_key, 2 * sizeof(double)));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_root, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&x->c.ev_root_join, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventCreate(&x->c.ev0));
    VBMC_CUDA_CHECK(cudaEventCreate(&x->c.ev1));
    VBMC_CUDA_CHECK(cudaEventCreate(&x->c.ev2));
    *out = reinterpret_cast<vbmc_ctx *>(x);
    return VBMC_OK;
}

void vbmc_ctx_destroy(vbmc_ctx *p) {
    if (!p) return;
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drop_graph(x);
    drop_adam_graph(x);
    p2p_unmap(c);
    if (c->p2p_local) cudaFree(c->p2p_local);
    if (x->d_adam) cudaFree(x->d_adam);
    for (auto &t : x->adam_tickets)
        if (t.ev) cudaEventDestroy(t.ev);
    if (c->d_tailsync) cudaFree(c->d_tailsync);
    if (c->d_key) cudaFree(c->d_key);
    double *dev[] = {c->d_Xt, c->d_alpha, c->d_hyp, c->d_L,  c->d_Linv, c->d_lb,   c->d_ub, c->d_in, c->d_lamc, c->d_outs,
                     c->d_entpart, c->d_gps, c->d_raw, c->d_csum, c->d_tctab, c->d_tctiles[0], c->d_tctiles[1], c->d_bprm, c->d_bout, c->d_out, c->d_eps, c->d_lbws, c->d_var, c->d_xs, c->d_pred};
    for (double *d : dev)
        if (d) cudaFree(d);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_theta) cudaFreeHost(c->h_theta);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_noise) cudaEventDestroy(c->ev_noise);
    if (c->ev_root) cudaEventDestroy(c->ev_root);
    if (c->ev_root_join) cudaEventDestroy(c->ev_root_join);
    if (c->stream3) cudaStreamDestroy(c->stream3);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete x;
}

void *vbmc_ctx_stream(vbmc_ctx *p) { return p ? (void *)ex(p)->c.stream : nullptr; }
int64_t vbmc_ctx_launch_count(vbmc_ctx *p) { return p ? ex(p)->c.launches : 0; }

int vbmc_gp_pack(vbmc_ctx *p, int D, int N, int S, const double *X, const double *hyp, int H, const double *alpha,
                 const double *L, const int *L_chol, const double *sn2_eff, int mean_kind, int cov_N, int noise_N) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "gp_pack: null ctx");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_REQUIRE(X && hyp && alpha, VBMC_ERR_ARG, "gp_pack: null array");
    VBMC_REQUIRE(D >= 1 && N >= 1 && S >= 1, VBMC_ERR_ARG, "gp_pack: bad sizes");
    VBMC_REQUIRE(D <= kMaxD, VBMC_ERR_UNSUPPORTED, "D > 32 is not supported by the CUDA path");
    VBMC_REQUIRE(cov_N == D + 1, VBMC_ERR_UNSUPPORTED, "gp_pack: only the SE-ARD covariance (D+1 hyper-parameters) is supported");
    VBMC_REQUIRE(mean_kind == VBMC_MEAN_ZERO || mean_kind == VBMC_MEAN_CONST || mean_kind == VBMC_MEAN_NEGQUAD,
                 VBMC_ERR_UNSUPPORTED, "gp_pack: unsupported mean function");
    const int base = cov_N + noise_N;
    const int need_H = base + (mean_kind == VBMC_MEAN_ZERO ? 0 : (mean_kind == VBMC_MEAN_CONST ? 1 : 1 + 2 * D));
    VBMC_REQUIRE(H >= need_H, VBMC_ERR_ARG, "gp_pack: hyp rows are too short for this mean function");
    const int DP = pad_dim(D), hs = hyp_stride(DP);

    cudaStreamSynchronize(c->stream);
    double *old[] = {c->d_Xt, c->d_alpha, c->d_hyp, c->d_L, c->d_Linv};
    for (double *d : old)
        if (d) cudaFree(d);
    c->d_Xt = c->d_alpha = c->d_hyp = c->d_L = c->d_Linv = nullptr;
    c->has_gp = false;

    std::vector<double> Xt((size_t)DP * N, 0.0), hp((size_t)S * hs, 0.0);
    for (int n = 0; n < N; ++n)
        for (int d = 0; d < D; ++d) Xt[(size_t)d * N + n] = X[(size_t)n * D + d];
    for (int s = 0; s < S; ++s) {
        const double *h = hyp + (size_t)s * H;
        double *o = hp.data() + (size_t)s * hs;
        double sum_lnell = 0.0;
        for (int d = 0; d < D; ++d) {
            o[d] = exp(h[d]);
            sum_lnell += h[d];
            o[DP + d] = 0.0;
            o[2 * DP + d] = 0.0;
        }
        for (int d = D; d < DP; ++d) o[d] = 1.0;
        o[3 * DP + 0] = 2.0 * h[D];
        o[3 * DP + 1] = sum_lnell;
        o[3 * DP + 2] = mean_kind == VBMC_MEAN_ZERO ? 0.0 : h[base];
        o[3 * DP + 3] = sn2_eff ? sn2_eff[s] : 1.0;
        o[3 * DP + 4] = L_chol ? (double)(L_chol[s] != 0) : 1.0;
        if (mean_kind == VBMC_MEAN_NEGQUAD)
            for (int d = 0; d < D; ++d) {
                o[DP + d] = h[base + 1 + d];
                const double om = exp(h[base + 1 + D + d]);
                o[2 * DP + d] = 1.0 / (om * om);
            }
    }
    VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_Xt, Xt.size() * sizeof(double)));
    VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_alpha, (size_t)S * N * sizeof(double)));
    VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_hyp, hp.size() * sizeof(double)));
    VBMC_CUDA_CHECK(cudaMemcpy(c->d_Xt, Xt.data(), Xt.size() * sizeof(double), cudaMemcpyHostToDevice));
    VBMC_CUDA_CHECK(cudaMemcpy(c->d_alpha, alpha, (size_t)S * N * sizeof(double), cudaMemcpyHostToDevice));
    VBMC_CUDA_CHECK(cudaMemcpy(c->d_hyp, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->has_L = false;
    if (L) {
        VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_L, (size_t)S * N * N * sizeof(double)));
        VBMC_CUDA_CHECK(cudaMemcpy(c->d_L, L, (size_t)S * N * N * sizeof(double), cudaMemcpyHostToDevice));
        c->has_L = true;
    }
    c->gD = D, c->gDP = DP, c->N = N, c->S = S, c->mean_kind = mean_kind;
    c->has_gp = true;
    c->staged = false;
    ex(p)->gen++;  // captured graphs reference the old GP buffers
    return VBMC_OK;
}

int vbmc_set_bounds(vbmc_ctx *p, int n, const double *lb, const double *ub, double tol_con, double weight_threshold,
                    double weight_penalty) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "set_bounds: null ctx");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_REQUIRE(n >= 0, VBMC_ERR_ARG, "set_bounds: negative length");
    c->n_bnd = 0;
    ex(p)->gen++;
    if (n == 0) return VBMC_OK;
    VBMC_REQUIRE(lb && ub, VBMC_ERR_ARG, "set_bounds: null array");
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if ((size_t)n > c->bnd_cap) {
        if (c->d_lb) cudaFree(c->d_lb);
        if (c->d_ub) cudaFree(c->d_ub);
        c->d_lb = c->d_ub = nullptr;
        VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_lb, (size_t)n * sizeof(double)));
        VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_ub, (size_t)n * sizeof(double)));
        c->bnd_cap = n;
    }
    VBMC_CUDA_CHECK(cudaMemcpy(c->d_lb, lb, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    VBMC_CUDA_CHECK(cudaMemcpy(c->d_ub, ub, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    c->n_bnd = n;
    c->tol_con = tol_con, c->w_thr = weight_threshold, c->w_pen = weight_penalty;
    ex(p)->gen++;  // bounds (pointers and scalars) are baked into captured graphs
    return VBMC_OK;
}

int vbmc_entmc(vbmc_ctx *p, const vbmc_vp *vp, int64_t Ns, const int grad_flags[4], int jacobian_flag, int rng_mode,
               const double *eps, uint64_t seed, uint64_t offset, int precision, double *H, double *dH) {
    VBMC_REQUIRE(p && vp && grad_flags && H, VBMC_ERR_ARG, "entmc: null argument");
    VBMC_REQUIRE(Ns > 0, VBMC_ERR_ARG, "entmc: Ns must be > 0");
    CtxEx *x = ex(p);
    Bind b(&x->c);
    Spec s;
    s.vp = *vp;
    for (int i = 0; i < 4; ++i) s.grad[i] = grad_flags[i] != 0;
    s.jacobian = jacobian_flag != 0;
    s.Ns = Ns;
    s.have_gp = false;
    s.rng_mode = rng_mode, s.eps = eps, s.seed = seed, s.offset = offset, s.precision = precision;
    const int P = packed_len(vp->D, vp->K, s.grad);
    const size_t Pfull = RawLayout{vp->D, vp->K}.block();
    VBMC_TRY(run_single(x, s, kOutHead + 2 * Pfull));
    const double *o = x->c.h_out;
    *H = o[2];
    if (dH)
        for (int i = 0; i < P; ++i) dH[i] = o[kOutHead + Pfull + i];
    return VBMC_OK;
}

int vbmc_entlb(vbmc_ctx *p, const vbmc_vp *vp, const int grad_flags[4], int jacobian_flag, double *H, double *dH) {
    VBMC_REQUIRE(p && vp && grad_flags && H, VBMC_ERR_ARG, "entlb: null argument");
    CtxEx *x = ex(p);
    Bind b(&x->c);
    Spec s;
    s.vp = *vp;
    for (int i = 0; i < 4; ++i) s.grad[i] = grad_flags[i] != 0;
    s.jacobian = jacobian_flag != 0;
    s.Ns = 0;
    s.have_gp = false;
    const int P = packed_len(vp->D, vp->K, s.grad);
    const size_t Pfull = RawLayout{vp->D, vp->K}.block();
    VBMC_TRY(run_single(x, s, kOutHead + 2 * Pfull));
    const double *o = x->c.h_out;
    *H = o[2];
    if (dH)
        for (int i = 0; i < P; ++i) dH[i] = o[kOutHead + Pfull + i];
    return VBMC_OK;
}

int vbmc_philox_normals(vbmc_ctx *p, int D, int K, int64_t Ns, uint64_t seed, uint64_t offset, double *eps_out) {
    VBMC_REQUIRE(p && eps_out, VBMC_ERR_ARG, "philox_normals: null argument");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    const int64_t half = even_ns(Ns) / 2;
    const size_t n = (size_t)K * half * D;
    VBMC_TRY(ensure(&c->d_eps, &c->eps_cap, n));
    VBMC_TRY(philox_normals_launch(c, D, K, half, seed, offset, c->d_eps));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(eps_out, c->d_eps, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return VBMC_OK;
}

int vbmc_gplogjoint(vbmc_ctx *p, const vbmc_vp *vp, const int grad_flags[4], int avg_flag, int jacobian_flag,
                    int compute_var, int separate_K, double *G, double *dG, double *varG, double *var_ss,
                    double *I_sk, double *J_sjk) {
    VBMC_REQUIRE(p && vp && grad_flags, VBMC_ERR_ARG, "gplogjoint: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    Spec s;
    s.vp = *vp;
    // the reference appends the sigma/lambda/w blocks only under jacobian_flag (:1529-1546)
    s.grad[0] = grad_flags[0] != 0;
    for (int i = 1; i < 4; ++i) s.grad[i] = (grad_flags[i] != 0) && (jacobian_flag != 0);
    s.jacobian = jacobian_flag != 0;
    s.have_ent = false;
    const bool anyg = grad_flags[0] || grad_flags[1] || grad_flags[2] || grad_flags[3];
    VBMC_REQUIRE(!(compute_var && anyg), VBMC_ERR_UNSUPPORTED,
                 "gradient of the log-joint variance is not available (reference raises at :1302-1307)");
    VBMC_REQUIRE(compute_var == 0 || compute_var == 1, VBMC_ERR_UNSUPPORTED,
                 "diagonal variance approximation is not implemented (reference raises at :1467-1471)");
    s.compute_var = compute_var != 0;
    s.avg = avg_flag != 0;
    const int D = vp->D, K = vp->K;
    const int P = packed_len(D, K, s.grad);
    const size_t Pfull = RawLayout{D, K}.block();
    VBMC_TRY(run_single(x, s, kOutHead + 3 * Pfull));
    const int S = c->S;
    const double *o = c->h_out;
    const bool per_s = !(avg_flag && S > 1) && S > 1;
    std::vector<double> gps;
    if (per_s || separate_K) {
        gps.resize((size_t)S * (1 + Pfull));
        VBMC_CUDA_CHECK(cudaMemcpyAsync(gps.data(), c->d_gps, gps.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    if (!per_s) {
        if (G) *G = o[1];
        if (dG)
            for (int i = 0; i < P; ++i) dG[i] = o[kOutHead + 2 * Pfull + i];
    } else {
        // per-sample Jacobians on the device, then [S][1+P] -> G[S], dG[P][S]
        x->st.f.avg = 0;
        VBMC_TRY(ensure(&c->d_outs, &c->outs_cap, (size_t)S * (1 + Pfull)));
        VBMC_TRY(gps_finalize_launch(c, c->d_in, D, K, x->st.f, c->d_outs));
        std::vector<double> hs((size_t)S * (1 + Pfull));
        VBMC_CUDA_CHECK(cudaMemcpyAsync(hs.data(), c->d_outs, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int si = 0; si < S; ++si) {
            if (G) G[si] = hs[(size_t)si * (1 + Pfull)];
            if (dG)
                for (int i = 0; i < P; ++i) dG[(size_t)i * S + si] = hs[(size_t)si * (1 + Pfull) + 1 + i];
        }
    }
    if (separate_K && I_sk) {
        RawLayout rl{D, K};
        for (int si = 0; si < S; ++si)
            for (int k = 0; k < K; ++k) I_sk[(size_t)si * K + k] = gps[(size_t)si * (1 + Pfull) + 1 + rl.o_w() + k];
    }
    if (compute_var) {
        std::vector<double> ov(2 + S);
        VBMC_CUDA_CHECK(cudaMemcpyAsync(ov.data(), gpvar_out(c, K), ov.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (separate_K && J_sjk)
            VBMC_CUDA_CHECK(cudaMemcpyAsync(J_sjk, gpvar_J(c, K), (size_t)S * K * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (var_ss) *var_ss = ov[1];
        if (varG) {
            if (per_s || S == 1)
                for (int si = 0; si < S; ++si) varG[si] = ov[2 + si];
            else
                *varG = ov[0];
        }
    }
    return VBMC_OK;
}

static int spec_from_in(const vbmc_elcbo_in *in, Spec *s) {
    VBMC_REQUIRE(in, VBMC_ERR_ARG, "negelcbo: null input");
    VBMC_REQUIRE(!(in->separate_K && in->compute_grad), VBMC_ERR_ARG,
                 "gradient and per-component results requested together (reference raises ValueError at :1114-1118)");
    s->vp = in->vp;
    for (int i = 0; i < 4; ++i) {
        s->optimize[i] = in->optimize[i] != 0;
        s->grad[i] = in->compute_grad ? s->optimize[i] : 0;
    }
    s->jacobian = 1;
    s->ln_sigma_b = in->ln_sigma_b, s->ln_lambd_b = in->ln_lambd_b, s->eta_b = in->eta_b;
    s->Ns = in->Ns;
    s->use_bounds = in->use_bounds != 0;
    s->rng_mode = in->rng_mode, s->eps = in->eps, s->seed = in->seed, s->offset = in->offset;
    s->precision = in->precision;
    return VBMC_OK;
}

int vbmc_negelcbo(vbmc_ctx *p, const vbmc_elcbo_in *in, vbmc_elcbo_out *out) {
    VBMC_REQUIRE(p && in && out, VBMC_ERR_ARG, "negelcbo: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    Spec s;
    VBMC_TRY(spec_from_in(in, &s));
    s.compute_var = in->compute_var != 0;
    VBMC_REQUIRE(!(s.compute_var && in->compute_grad), VBMC_ERR_UNSUPPORTED,
                 "gradient of the ELBO variance is not available (reference raises at :1066-1070 / :1302-1307)");
    const int D = in->vp.D, K = in->vp.K;
    const int P = packed_len(D, K, s.grad);
    const size_t Pfull = RawLayout{D, K}.block();
    VBMC_TRY(run_single(x, s, kOutHead + (in->compute_grad ? 2 * Pfull : 0)));
    const double *o = c->h_out;
    if (o[7] != 0.0 && s.precision == VBMC_PREC_F32 && s.Ns > 0) {
        // fp32 density ratios overflowed (or a weight underflowed): redo the entropy in fp64 on the GPU
        s.precision = VBMC_PREC_F64;
        VBMC_TRY(run_single(x, s, kOutHead + (in->compute_grad ? 2 * Pfull : 0)));
        o = c->h_out;
    }
    out->F = o[0], out->G = o[1], out->H = o[2], out->varF = 0.0, out->varG_ss = 0.0;
    // every device-to-host copy of the extra outputs is enqueued first, ONE synchronisation for all of them
    const bool want_I = in->separate_K && out->I_sk;
    std::vector<double> ov, gps;
    if (s.compute_var) {
        ov.resize(2 + c->S);
        VBMC_CUDA_CHECK(cudaMemcpyAsync(ov.data(), gpvar_out(c, K), ov.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (in->separate_K && out->J_sjk)
            VBMC_CUDA_CHECK(cudaMemcpyAsync(out->J_sjk, gpvar_J(c, K), (size_t)c->S * K * K * sizeof(double),
                                            cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_I) {
        gps.resize((size_t)c->S * (1 + Pfull));
        VBMC_CUDA_CHECK(cudaMemcpyAsync(gps.data(), c->d_gps, gps.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    if (s.compute_var || want_I) VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (s.compute_var) {
        out->varF = ov[0];     // varG (+ varH == 0, :1179-1181)
        out->varG_ss = ov[1];  // the 5th output of _gp_log_joint, i.e. var_ss (:1121,1586)
    }
    if (in->compute_grad) {
        if (out->dF)
            for (int i = 0; i < P; ++i) out->dF[i] = o[kOutHead + i];
        if (out->dH)
            for (int i = 0; i < P; ++i) out->dH[i] = o[kOutHead + Pfull + i];
    }
    if (want_I) {
        RawLayout rl{D, K};
        for (int si = 0; si < c->S; ++si)
            for (int k = 0; k < K; ++k) out->I_sk[(size_t)si * K + k] = gps[(size_t)si * (1 + Pfull) + 1 + rl.o_w() + k];
    }
    return VBMC_OK;
}

int vbmc_negelcbo_flat(vbmc_ctx *p, int D, int K, const double *params, const int optimize[4], int64_t Ns,
                       int compute_grad, int use_bounds, int rng_mode, const double *eps, uint64_t seed,
                       uint64_t offset, int precision, int want_dH, double *out) {
    VBMC_REQUIRE(p && params && optimize && out, VBMC_ERR_ARG, "negelcbo_flat: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    Spec s;
    s.flat = params;
    s.vp.D = D, s.vp.K = K;
    s.vp.mu = s.vp.sigma = s.vp.lambd = s.vp.w = s.vp.eta = nullptr;
    for (int i = 0; i < 4; ++i) {
        s.optimize[i] = optimize[i] != 0;
        s.grad[i] = compute_grad ? s.optimize[i] : 0;
    }
    s.jacobian = 1;
    s.Ns = Ns;
    s.use_bounds = use_bounds != 0;
    s.rng_mode = rng_mode, s.eps = eps, s.seed = seed, s.offset = offset, s.precision = precision;
    s.parts = want_dH != 0;
    const int P = packed_len(D, K, s.grad);
    const size_t Pfull = RawLayout{D, K}.block();
    const size_t n_dev = kOutHead + (compute_grad ? (want_dH ? 2 * Pfull : (size_t)P) : 0);
    VBMC_TRY(run_flat(x, s, n_dev));
    const double *o = c->h_out;
    if (o[7] != 0.0 && s.precision == VBMC_PREC_F32 && s.Ns > 0) {
        s.precision = VBMC_PREC_F64;  // fp32 density ratios overflowed: redo the entropy in fp64 on the GPU
        VBMC_TRY(run_single(x, s, n_dev));
        o = c->h_out;
    }
    memcpy(out, o, sizeof(double) * (kOutHead + (compute_grad ? P : 0)));
    if (compute_grad && want_dH) memcpy(out + kOutHead + P, o + kOutHead + Pfull, sizeof(double) * P);
    return VBMC_OK;
}

int vbmc_noise_prefetch(vbmc_ctx *p, int D, int K, int64_t Ns, uint64_t seed, uint64_t offset) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "noise_prefetch: null ctx");
    if (Ns <= 0 || D < 1 || K < 1 || D > kMaxD) return VBMC_OK;
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    const int64_t half = even_ns(Ns) / 2;
    EntmcPlan plan{};
    if (entmc_plan(c, D, K, half, true, VBMC_PREC_F32, &plan) != VBMC_OK || plan.variant != ENTMC_TC) return VBMC_OK;
    plan.pair0 = 0, plan.half_glob = half;
    return entmc_tc_prefetch(c, ParamLayout{D, pad_dim(D), K}, plan, seed, offset);
}

int vbmc_negelcbo_theta(vbmc_ctx *p, int D, int K, double *theta, const double *tmpl, const int optimize[4], int64_t Ns,
                        int compute_grad, int use_bounds, uint64_t seed, uint64_t offset, int precision, double *out,
                        double *vp_out) {
    VBMC_REQUIRE(p && theta && optimize && out && vp_out, VBMC_ERR_ARG, "negelcbo_theta: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    Spec s;
    s.theta = theta, s.tmpl = tmpl;
    s.vp.D = D, s.vp.K = K;
    s.vp.mu = s.vp.sigma = s.vp.lambd = s.vp.w = s.vp.eta = nullptr;
    for (int i = 0; i < 4; ++i) {
        s.optimize[i] = optimize[i] != 0;
        s.grad[i] = compute_grad ? s.optimize[i] : 0;
    }
    s.P = packed_len(D, K, s.optimize);
    VBMC_REQUIRE(s.P > 0, VBMC_ERR_ARG, "negelcbo_theta: no parameter group is optimised (theta is empty)");
    VBMC_REQUIRE(tmpl || (s.optimize[0] && s.optimize[1] && s.optimize[2] && s.optimize[3]), VBMC_ERR_ARG,
                 "negelcbo_theta: a template block is required for the groups theta does not carry");
    s.jacobian = 1;
    s.Ns = Ns;
    s.use_bounds = use_bounds != 0;
    s.rng_mode = VBMC_RNG_PHILOX, s.eps = nullptr, s.seed = seed, s.offset = offset, s.precision = precision;
    s.parts = false;
    const int Pg = packed_len(D, K, s.grad);
    const size_t n_dev = kOutHead + (size_t)Pg;
    VBMC_TRY(run_flat(x, s, n_dev));
    const double *o = c->h_out;
    if (o[7] != 0.0 && s.precision == VBMC_PREC_F32 && s.Ns > 0) {
        s.precision = VBMC_PREC_F64;  // fp32 density ratios overflowed: redo the entropy in fp64 on the GPU
        VBMC_TRY(run_single(x, s, n_dev));
        o = c->h_out;
    }
    memcpy(out, o, sizeof(double) * (kOutHead + Pg));
    const double *vpo = c->h_theta + s.P + ParamLayout{D, pad_dim(D), K}.total() + 2;
    memcpy(vp_out, vpo, sizeof(double) * (2 * K + D));
    // the reference shifts the eta block of the caller's theta in place (variational_optimization.py:1082-1085);
    // the prepare kernel did it on the pinned copy
    if (s.optimize[3]) memcpy(theta + s.P - K, c->h_theta + s.P - K, sizeof(double) * K);
    return VBMC_OK;
}

int vbmc_adam_init(vbmc_ctx *p, const vbmc_adam_in *in) {
    VBMC_REQUIRE(p && in && in->params && in->theta0, VBMC_ERR_ARG, "adam_init: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    const int D = in->D, K = in->K;
    VBMC_REQUIRE(in->max_iter >= 1, VBMC_ERR_ARG, "adam_init: max_iter must be >= 1");
    VBMC_REQUIRE(in->Ns > 0, VBMC_ERR_ARG, "adam_init: the stochastic optimiser needs Ns > 0 (reference: :173-176)");
    Spec s;
    s.flat = in->params;
    s.vp.D = D, s.vp.K = K;
    s.vp.mu = s.vp.sigma = s.vp.lambd = s.vp.w = s.vp.eta = nullptr;
    for (int i = 0; i < 4; ++i) s.optimize[i] = in->optimize[i] != 0, s.grad[i] = s.optimize[i];
    s.jacobian = 1;
    s.Ns = in->Ns;
    s.use_bounds = in->use_bounds != 0;
    s.rng_mode = VBMC_RNG_PHILOX, s.eps = nullptr, s.seed = in->seed, s.offset = in->offset, s.precision = in->precision;
    s.parts = false;
    const int P = packed_len(D, K, s.grad);
    VBMC_REQUIRE(P > 0, VBMC_ERR_ARG, "adam_init: nothing to optimise");
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    VBMC_TRY(stage(x, s));  // flags, plan inputs, parameter block (the template) on the device
    const ParamLayout lay{D, pad_dim(D), K};
    const size_t T = (size_t)lay.total();
    // theta, m, v, lb, ub (P each) | tmpl (T) | ytab (max_iter) | iter (1, as 8 bytes) | xtab (max_iter * P)
    const size_t need = 5 * (size_t)P + T + (size_t)in->max_iter + 1 + (size_t)in->max_iter * P;
    if (need > x->adam_cap) {
        if (x->d_adam) cudaFree(x->d_adam);
        x->d_adam = nullptr, x->adam_cap = 0;
        VBMC_CUDA_CHECK(cudaMalloc((void **)&x->d_adam, need * sizeof(double)));
        x->adam_cap = need;
    }
    drop_adam_graph(x);
    AdamDev &a = x->adam;
    a = AdamDev{};
    a.lay = lay, a.P = P;
    for (int i = 0; i < 4; ++i) a.opt[i] = s.optimize[i];
    double *q = x->d_adam;
    a.theta = q, q += P;
    a.m = q, q += P;
    a.v = q, q += P;
    double *d_lb = q;
    q += P;
    double *d_ub = q;
    q += P;
    double *d_tmpl = q;
    q += T;
    a.ytab = q, q += in->max_iter;
    a.iter = reinterpret_cast<long long *>(q), q += 1;
    a.xtab = q;
    VBMC_CUDA_CHECK(cudaMemsetAsync(a.m, 0, 2 * (size_t)P * sizeof(double), c->stream));  // m, v
    VBMC_CUDA_CHECK(cudaMemsetAsync(a.iter, 0, sizeof(long long), c->stream));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(a.theta, in->theta0, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(d_tmpl, in->params, T * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    a.tmpl = d_tmpl;
    a.lb = a.ub = nullptr;
    if (in->lb) {
        VBMC_CUDA_CHECK(cudaMemcpyAsync(d_lb, in->lb, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        a.lb = d_lb;
    }
    if (in->ub) {
        VBMC_CUDA_CHECK(cudaMemcpyAsync(d_ub, in->ub, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        a.ub = d_ub;
    }
    a.seed = in->seed, a.offset0 = in->offset;
    a.master_min = in->master_min, a.master_max = in->master_max, a.master_decay = in->master_decay;
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    x->adam_max_iter = in->max_iter;
    x->adam_done = 0;
    x->adam_issued = 0;
    x->adam_ticket_next = 0;
    for (auto &t : x->adam_tickets) t.i_end = 0;
    x->adam_params_ready = false;
    x->adam_ready = true, x->adam_eager_done = false;
    return VBMC_OK;
}

// enqueue n iterations on the context stream (no synchronisation)
static int adam_enqueue(CtxEx *x, int n) {
    Ctx *c = &x->c;
    VBMC_REQUIRE(x->adam_ready && c->staged, VBMC_ERR_STATE, "adam_steps: call vbmc_adam_init first");
    VBMC_REQUIRE(n >= 0 && x->adam_done + n <= x->adam_max_iter, VBMC_ERR_ARG, "adam_steps: more steps than max_iter");
    VBMC_TRY(settle_prefetch(c));
    const long long i0 = x->adam_done;
    // Iterations are captured and replayed in PAIRS: the tensor-core entropy kernel alternates between two noise-tile
    // buffers (the draws of iteration i + 1 are generated under the tail of iteration i), so one captured pair leaves
    // the buffers where it found them and can be replayed back to back.  Single iterations (the first one, which
    // sizes every buffer; an odd remainder) run eagerly; an eager iteration also flips the buffer parity back when a
    // remainder left it opposite to the one the pair was captured with.
    for (int it = 0; it < n;) {
        if (x->seen_epoch != g_realloc_epoch.load(std::memory_order_relaxed)) {  // a buffer moved: captured pointers are stale
            x->gen++;
            x->seen_epoch = g_realloc_epoch.load(std::memory_order_relaxed);
        }
        const bool pair_ok = x->graphs_on && x->adam_eager_done && n - it >= 2;
        const bool state_ok = c->noise_buf == x->adam_graph_buf && c->noise_ready == x->adam_graph_ready &&
                              x->adam_params_ready == x->adam_graph_params_ready &&
                              (!c->noise_ready || (c->noise_seed == x->adam.seed &&
                                                   c->noise_offset == x->adam.offset0 + (uint64_t)x->adam_issued));
        if (pair_ok && x->adam_gexec && x->adam_gen == x->gen && state_ok) {
            VBMC_CUDA_CHECK(cudaGraphLaunch(x->adam_gexec, c->stream));
            c->launches += x->adam_launches_per_iter;
            x->adam_issued += 2;  // what the two captured iterations did to the host-side bookkeeping
            if (c->noise_ready) c->noise_offset += 2;
            it += 2;
            continue;
        }
        if (!pair_ok || (x->adam_gexec && x->adam_gen == x->gen && !state_ok)) {
            VBMC_TRY(adam_iteration(x));
            x->adam_eager_done = true;
            it += 1;
            continue;
        }
        drop_adam_graph(x);
        const uint64_t epoch0 = g_realloc_epoch.load(std::memory_order_relaxed);
        const int64_t l0 = c->launches;
        x->adam_graph_buf = c->noise_buf, x->adam_graph_ready = c->noise_ready;
        x->adam_graph_params_ready = x->adam_params_ready;
        VBMC_CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = adam_iteration(x);
        if (rc == VBMC_OK) rc = adam_iteration(x);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
        if (rc != VBMC_OK || ce != cudaSuccess || epoch0 != g_realloc_epoch.load(std::memory_order_relaxed)) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            x->gen++;
            x->seen_epoch = g_realloc_epoch.load(std::memory_order_relaxed);
            x->adam_eager_done = false;  // start over with an eager iteration (nothing of the capture has run)
            c->noise_ready = false;
            c->noise_pending_join = false;
            x->adam_issued = x->adam_done + it;  // nothing of the capture has run
            x->adam_params_ready = x->adam_graph_params_ready;
            if (rc != VBMC_OK) return rc;
            if (ce != cudaSuccess) {
                set_error(std::string("adam_steps: stream capture failed: ") + cudaGetErrorString(ce));
                return VBMC_ERR_CUDA;
            }
            continue;
        }
        x->adam_graph = g;
        x->adam_launches_per_iter = (int)(c->launches - l0);
        c->launches = l0;
        VBMC_CUDA_CHECK(cudaGraphInstantiate(&x->adam_gexec, g, 0));
        x->adam_gen = x->gen;
        VBMC_CUDA_CHECK(cudaGraphLaunch(x->adam_gexec, c->stream));
        c->launches += x->adam_launches_per_iter;
        it += 2;
    }
    x->adam_done += n;
    (void)i0;
    return VBMC_OK;
}

int vbmc_adam_steps(vbmc_ctx *p, int n, double *y, double *xs) {
    VBMC_REQUIRE(p && y && xs, VBMC_ERR_ARG, "adam_steps: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    const long long i0 = x->adam_done;
    VBMC_TRY(adam_enqueue(x, n));
    const int P = x->adam.P;
    if (n > 0) {
        VBMC_CUDA_CHECK(cudaMemcpyAsync(y, x->adam.ytab + i0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VBMC_CUDA_CHECK(cudaMemcpyAsync(xs, x->adam.xtab + (size_t)i0 * P, (size_t)n * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return VBMC_OK;
}

// Split-phase form: vbmc_adam_enqueue(n) returns as soon as the n iterations are in the stream; vbmc_adam_fetch(i0, n)
// waits for iterations [i0, i0 + n) only (an event recorded behind the enqueue call that issued them) and copies their
// objective values and iterates on a second stream.  The host loop keeps ONE batch in flight while it digests the
// previous one (early-stopping fit, minimize_adam.py:106-138), so the device never waits for the host.
int vbmc_adam_enqueue(vbmc_ctx *p, int n) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "adam_enqueue: null ctx");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    VBMC_TRY(adam_enqueue(x, n));
    auto &t = x->adam_tickets[x->adam_ticket_next % CtxEx::kAdamTickets];
    if (!t.ev) VBMC_CUDA_CHECK(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
    VBMC_CUDA_CHECK(cudaEventRecord(t.ev, c->stream));
    t.i_end = x->adam_done;
    x->adam_ticket_next++;
    return VBMC_OK;
}

int vbmc_adam_fetch(vbmc_ctx *p, int64_t i0, int n, double *y, double *xs) {
    VBMC_REQUIRE(p && y && xs, VBMC_ERR_ARG, "adam_fetch: null argument");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    VBMC_REQUIRE(x->adam_ready, VBMC_ERR_STATE, "adam_fetch: call vbmc_adam_init first");
    VBMC_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= x->adam_done, VBMC_ERR_ARG, "adam_fetch: iterations not issued yet");
    if (n == 0) return VBMC_OK;
    // the oldest live ticket that covers the range
    cudaEvent_t ev = nullptr;
    const int first = x->adam_ticket_next > CtxEx::kAdamTickets ? x->adam_ticket_next - CtxEx::kAdamTickets : 0;
    for (int k = first; k < x->adam_ticket_next && !ev; ++k) {
        auto &t = x->adam_tickets[k % CtxEx::kAdamTickets];
        if (t.ev && t.i_end >= i0 + n) ev = t.ev;
    }
    const int P = x->adam.P;
    if (!ev) {  // issued through vbmc_adam_steps, or the ticket ring wrapped: everything in the stream
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream3, ev, 0));
    }
    cudaStream_t cs = ev ? c->stream3 : c->stream;
    VBMC_CUDA_CHECK(cudaMemcpyAsync(y, x->adam.ytab + i0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, cs));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(xs, x->adam.xtab + (size_t)i0 * P, (size_t)n * P * sizeof(double), cudaMemcpyDeviceToHost, cs));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(cs));
    return VBMC_OK;
}

size_t vbmc_param_len(int D, int K) { return (size_t)ParamLayout{D, pad_dim(D), K}.total(); }

int vbmc_negelcbo_batch(vbmc_ctx *p, int B, int D, int K, const double *params, const int optimize[4], int use_bounds,
                        double *out) {
    VBMC_REQUIRE(p && params && optimize && out, VBMC_ERR_ARG, "negelcbo_batch: null argument");
    VBMC_REQUIRE(B >= 0 && D >= 1 && K >= 1, VBMC_ERR_ARG, "negelcbo_batch: bad sizes");
    VBMC_REQUIRE(D <= kMaxD, VBMC_ERR_UNSUPPORTED, "D > 32 is not supported by the CUDA path");
    if (B == 0) return VBMC_OK;
    Ctx *c = &ex(p)->c;
    Bind b(c);
    const size_t n_in = (size_t)B * vbmc_param_len(D, K), n_out = (size_t)B * 4;
    VBMC_TRY(ensure(&c->d_bprm, &c->bprm_cap, n_in));
    VBMC_TRY(ensure(&c->d_bout, &c->bout_cap, n_out));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(c->d_bprm, params, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VBMC_TRY(sieve_launch(c, B, D, K, optimize, use_bounds != 0, c->d_bprm, c->d_bout));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(out, c->d_bout, n_out * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return VBMC_OK;
}

size_t vbmc_raw_len(int D, int K) { return (size_t)RawLayout{D, K}.total(); }
size_t vbmc_out_len(int D, int K) { return (size_t)kOutHead + 3 * (size_t)RawLayout{D, K}.block() + (size_t)K * D; }

int vbmc_negelcbo_upload(vbmc_ctx *p, const vbmc_elcbo_in *in) {
    VBMC_REQUIRE(p && in, VBMC_ERR_ARG, "negelcbo_upload: null argument");
    CtxEx *x = ex(p);
    Bind b(&x->c);
    Spec s;
    VBMC_TRY(spec_from_in(in, &s));
    VBMC_REQUIRE(!in->compute_var, VBMC_ERR_UNSUPPORTED, "split-phase evaluation does not cover the variance path");
    x->partials_calls = 0;
    return stage(x, s);
}

int vbmc_negelcbo_partials_async(vbmc_ctx *p, int rank, int world, double *raw_dev) {
    VBMC_REQUIRE(p && raw_dev, VBMC_ERR_ARG, "partials: null argument");
    CtxEx *x = ex(p);
    Bind b(&x->c);
    // first call after vbmc_negelcbo_upload: the uploaded key; every further call: the next key (offset + 1, + 2, ...),
    // like consecutive optimiser steps -- its draws are generated under the previous call's tail
    VBMC_TRY(settle_prefetch(&x->c));
    x->c.key_delta = x->partials_calls++;
    x->c.lookahead = true;
    return partials(x, rank, world, raw_dev);
}

int vbmc_negelcbo_finalize_async(vbmc_ctx *p, const double *raw_dev, double *out_dev) {
    VBMC_REQUIRE(p && raw_dev && out_dev, VBMC_ERR_ARG, "finalize: null argument");
    CtxEx *x = ex(p);
    Bind b(&x->c);
    return finalize(x, raw_dev, out_dev);
}

int vbmc_gp_predict(vbmc_ctx *p, int Nx, const double *Xs, double *f_mu, double *f_s2) {
    VBMC_REQUIRE(p && Xs && f_mu && f_s2, VBMC_ERR_ARG, "gp_predict: null argument");
    VBMC_REQUIRE(Nx >= 0, VBMC_ERR_ARG, "gp_predict: negative number of points");
    if (Nx == 0) return VBMC_OK;
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_REQUIRE(c->has_gp, VBMC_ERR_STATE, "gp_predict: no GP packed");
    const int D = c->gD, S = c->S;
    VBMC_TRY(ensure(&c->d_xs, &c->xs_cap, (size_t)Nx * D));
    VBMC_TRY(ensure(&c->d_pred, &c->pred_cap, (size_t)2 * Nx * S));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(c->d_xs, Xs, (size_t)Nx * D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VBMC_TRY(gppred_launch(c, c->d_xs, Nx, c->d_pred, c->d_pred + (size_t)Nx * S));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(f_mu, c->d_pred, (size_t)Nx * S * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(f_s2, c->d_pred + (size_t)Nx * S, (size_t)Nx * S * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return VBMC_OK;
}

int vbmc_gp_predict_device_ms(vbmc_ctx *p, int Nx, int reps, double *ms) {
    // development / bench aid: average duration of the prediction kernel on the points uploaded last (no copies)
    VBMC_REQUIRE(p && ms && reps >= 1, VBMC_ERR_ARG, "gp_predict_device_ms: bad argument");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_REQUIRE(c->d_xs && c->d_pred && (size_t)Nx * c->gD <= c->xs_cap, VBMC_ERR_STATE, "call vbmc_gp_predict first");
    VBMC_TRY(gppred_launch(c, c->d_xs, Nx, c->d_pred, c->d_pred + (size_t)Nx * c->S));
    VBMC_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) VBMC_TRY(gppred_launch(c, c->d_xs, Nx, c->d_pred, c->d_pred + (size_t)Nx * c->S));
    VBMC_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    VBMC_CUDA_CHECK(cudaEventSynchronize(c->ev1));
    float t = 0;
    VBMC_CUDA_CHECK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    *ms = (double)t / reps;
    return VBMC_OK;
}

int vbmc_vp_pdf(vbmc_ctx *p, const vbmc_vp *vp, int Nx, const double *Xs, int log_flag, int grad_flag, double *y,
                double *dy) {
    VBMC_REQUIRE(p && vp && Xs && y && (!grad_flag || dy), VBMC_ERR_ARG, "vp_pdf: null argument");
    VBMC_REQUIRE(Nx >= 0, VBMC_ERR_ARG, "vp_pdf: negative number of points");
    VBMC_TRY(check_vp(vp, false));
    if (Nx == 0) return VBMC_OK;
    Ctx *c = &ex(p)->c;
    Bind b(c);
    const int D = vp->D, K = vp->K;
    const ParamLayout lay{D, pad_dim(D), K};
    VBMC_TRY(ensure_pinned(&c->d_in, &c->h_in, &c->in_cap, (size_t)lay.total() + 2));
    double *h = c->h_in;
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the pinned block may still feed an earlier copy
    memcpy(h + lay.mu(), vp->mu, sizeof(double) * K * D);
    memcpy(h + lay.sigma(), vp->sigma, sizeof(double) * K);
    memcpy(h + lay.lambd(), vp->lambd, sizeof(double) * D);
    memcpy(h + lay.w(), vp->w, sizeof(double) * K);
    c->staged = false;  // the parameter block no longer belongs to a staged evaluation
    ex(p)->adam_params_ready = false;
    VBMC_CUDA_CHECK(cudaMemcpyAsync(c->d_in, h, sizeof(double) * (K * D + 2 * K + D), cudaMemcpyHostToDevice, c->stream));
    VBMC_TRY(ensure(&c->d_xs, &c->xs_cap, (size_t)Nx * D));
    VBMC_TRY(ensure(&c->d_pred, &c->pred_cap, (size_t)Nx * (1 + D)));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(c->d_xs, Xs, (size_t)Nx * D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VBMC_TRY(vp_pdf_launch(c, c->d_in, D, K, c->d_xs, Nx, log_flag != 0, grad_flag != 0, c->d_pred, c->d_pred + Nx));
    VBMC_CUDA_CHECK(cudaMemcpyAsync(y, c->d_pred, (size_t)Nx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (grad_flag)
        VBMC_CUDA_CHECK(cudaMemcpyAsync(dy, c->d_pred + Nx, (size_t)Nx * D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return VBMC_OK;
}

int vbmc_negelcbo_enqueue(vbmc_ctx *p) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "enqueue: null ctx");
    CtxEx *x = ex(p);
    Ctx *c = &x->c;
    Bind b(c);
    VBMC_REQUIRE(c->staged && c->d_raw && c->d_out, VBMC_ERR_STATE,
                 "enqueue: nothing staged (call vbmc_negelcbo_flat or vbmc_negelcbo_upload first)");
    VBMC_REQUIRE(!x->st.s.compute_var, VBMC_ERR_UNSUPPORTED, "enqueue does not cover the variance path");
    // every call is a NEW evaluation: Philox key (seed, offset + number of calls since staging), like consecutive
    // iterations of the Adam loop; its draws were generated under the previous call's tail
    VBMC_TRY(settle_prefetch(c));
    c->key_delta += 1;
    c->lookahead = true;
    VBMC_TRY(partials(x, 0, 1, c->d_raw));
    return finalize(x, c->d_raw, c->d_out);
}

int vbmc_p2p_export(vbmc_ctx *p, int world, int D, int K, unsigned char *handle) {
    VBMC_REQUIRE(p && handle, VBMC_ERR_ARG, "p2p_export: null argument");
    VBMC_REQUIRE(world >= 2 && world <= VBMC_P2P_MAX_WORLD, VBMC_ERR_UNSUPPORTED, "p2p_export: world must be 2..8");
    static_assert(sizeof(cudaIpcMemHandle_t) == VBMC_P2P_HANDLE_BYTES, "CUDA IPC handle size");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    p2p_unmap(c);
    if (c->p2p_local) cudaFree(c->p2p_local);
    c->p2p_local = nullptr;
    const int stride = (int)((vbmc_raw_len(D, K) + 15) / 16 * 16);
    const size_t bytes = ((size_t)2 * world * stride + (size_t)2 * world) * sizeof(double);
    VBMC_CUDA_CHECK(cudaMalloc((void **)&c->p2p_local, bytes));
    VBMC_CUDA_CHECK(cudaMemset(c->p2p_local, 0, bytes));  // flags = 0: no epoch published yet
    VBMC_CUDA_CHECK(cudaDeviceSynchronize());
    c->p2p_bytes = bytes, c->p2p_stride = stride, c->p2p_epoch = 0;
    cudaIpcMemHandle_t h;
    VBMC_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->p2p_local));
    memcpy(handle, &h, sizeof(h));
    return VBMC_OK;
}

int vbmc_p2p_open(vbmc_ctx *p, int rank, int world, const unsigned char *handles) {
    VBMC_REQUIRE(p && handles, VBMC_ERR_ARG, "p2p_open: null argument");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    VBMC_REQUIRE(c->p2p_local != nullptr, VBMC_ERR_STATE, "p2p_open: call vbmc_p2p_export first");
    VBMC_REQUIRE(world >= 2 && world <= VBMC_P2P_MAX_WORLD && rank >= 0 && rank < world, VBMC_ERR_ARG, "p2p_open: bad rank/world");
    p2p_unmap(c);
    c->p2p_rank = rank;
    for (int q = 0; q < world; ++q) {
        if (q == rank) {
            c->p2p_peer[q] = c->p2p_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * VBMC_P2P_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            c->p2p_world = q;  // unmap what was opened so far
            p2p_unmap(c);
            set_error(std::string("p2p_open: cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
            return VBMC_ERR_CUDA;
        }
        c->p2p_peer[q] = static_cast<double *>(ptr);
    }
    c->p2p_world = world;
    c->p2p_epoch = 0;
    return VBMC_OK;
}

int vbmc_p2p_unmap(vbmc_ctx *p) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "null ctx");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    cudaStreamSynchronize(c->stream);
    p2p_unmap(c);
    return VBMC_OK;
}

int vbmc_p2p_close(vbmc_ctx *p) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "null ctx");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    cudaStreamSynchronize(c->stream);
    p2p_unmap(c);
    if (c->p2p_local) cudaFree(c->p2p_local);
    c->p2p_local = nullptr;
    return VBMC_OK;
}

int vbmc_read_device(vbmc_ctx *p, const double *src_dev, size_t n, double *dst_host) {
    VBMC_REQUIRE(p && src_dev && dst_host, VBMC_ERR_ARG, "read_device: null argument");
    Ctx *c = &ex(p)->c;
    Bind b(c);
    if (!c->h_out) VBMC_TRY(ensure_pinned(&c->d_out, &c->h_out, &c->out_cap, 4096));
    for (size_t off = 0; off < n; off += c->out_cap) {  // staged through the pinned buffer in chunks
        const size_t m = n - off < c->out_cap ? n - off : c->out_cap;
        VBMC_CUDA_CHECK(cudaMemcpyAsync(c->h_out, src_dev + off, m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VBMC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        memcpy(dst_host + off, c->h_out, m * sizeof(double));
    }
    return VBMC_OK;
}

int vbmc_stage_times(vbmc_ctx *p, double *us) {
    VBMC_REQUIRE(p && us, VBMC_ERR_ARG, "stage_times: null argument");
    Ctx *c = &ex(p)->c;
    VBMC_REQUIRE(c->stage_timing, VBMC_ERR_STATE, "stage timing is off (set VBMC_STAGE_TIMING=1 before creating the context)");
    for (int i = 0; i < 6; ++i) {
        float ms = 0;
        VBMC_CUDA_CHECK(cudaEventElapsedTime(&ms, c->sev[i], c->sev[i + 1]));
        us[i] = 1e3 * ms;
    }
    us[6] = c->host_us;
    return VBMC_OK;
}

int vbmc_stream_synchronize(vbmc_ctx *p) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "null ctx");
    Bind b(&ex(p)->c);
    VBMC_CUDA_CHECK(cudaStreamSynchronize(ex(p)->c.stream));
    return VBMC_OK;
}

int vbmc_entmc_kernel_ms(vbmc_ctx *p, double *avg_ms, int64_t *launches) {
    VBMC_REQUIRE(p && avg_ms && launches, VBMC_ERR_ARG, "null argument");
    Ctx *c = &ex(p)->c;
    *launches = c->entmc_ms_n;
    *avg_ms = c->entmc_ms_n ? c->entmc_ms_sum / (double)c->entmc_ms_n : 0.0;
    c->entmc_ms_sum = 0, c->entmc_ms_n = 0, c->entmc_main_ms_sum = 0;
    return VBMC_OK;
}

int vbmc_entmc_main_kernel_ms(vbmc_ctx *p, double *avg_ms) {
    VBMC_REQUIRE(p && avg_ms, VBMC_ERR_ARG, "null argument");
    Ctx *c = &ex(p)->c;
    *avg_ms = c->entmc_ms_n ? c->entmc_main_ms_sum / (double)c->entmc_ms_n : 0.0;
    return VBMC_OK;
}

int vbmc_entmc_variant_used(vbmc_ctx *p) {
    if (!p) return -1;
    const Staged &st = ex(p)->st;
    return st.plan.threads > 0 ? st.plan.variant : -1;
}

int vbmc_set_kernel_timing(vbmc_ctx *p, int on) {
    VBMC_REQUIRE(p, VBMC_ERR_ARG, "null ctx");
    Ctx *c = &ex(p)->c;
    c->time_entmc = on != 0;
    c->entmc_ms_sum = 0, c->entmc_ms_n = 0, c->entmc_main_ms_sum = 0;
    return VBMC_OK;
}

}  // extern "C"
