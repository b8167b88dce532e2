// hvs_engine.cu -- the C ABI of include/hvs.h: engine lifetime, solve orchestration, statistics.
//
// One engine = one GPU = one stream.  solve():
//   H2D queries -> K1 slice search -> D2H slices -> host planner -> H2D work lists
//   -> K4 direct scan (sparse slices)  +  K2/K3 tile sweeps (shared slices) -> K5 finalize -> D2H ids.
// No CPU fallback anywhere: every distance and every selection runs in the CUDA kernels of this
// library; the host only plans.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "hvs_engine.h"

using namespace hvs;

static thread_local std::string g_create_err = "";

#define EFAIL(code, msg)            \
    do {                            \
        e->err = (msg);             \
        return (code);              \
    } while (0)

#define ECUDA(call)                                                                        \
    do {                                                                                   \
        cudaError_t _c = (call);                                                           \
        if (_c != cudaSuccess) {                                                           \
            if (e->err.empty()) e->err = std::string(#call) + ": " + cudaGetErrorString(_c);                    \
            return HVS_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

extern "C" uint32_t hvs_abi_version(void) { return HVS_ABI_VERSION; }

extern "C" const char *hvs_last_error(const hvs_engine *e)
{
    if (!e) return g_create_err.c_str();
    return e->err.c_str();
}

extern "C" int hvs_create(hvs_engine **out, const hvs_config *cfg)
{
    if (!out) { g_create_err = "hvs_create: out is NULL"; return HVS_ERR_INVALID; }
    *out = nullptr;
    hvs_config c{};
    c.device = -1;
    if (cfg) {
        if (cfg->struct_size < 16 || cfg->struct_size > sizeof(hvs_config)) {
            g_create_err = "hvs_create: bad hvs_config.struct_size";
            return HVS_ERR_INVALID;
        }
        std::memcpy(&c, cfg, cfg->struct_size);
    }
    if (c.mode > HVS_MODE_TENSOR) { g_create_err = "hvs_create: unknown mode"; return HVS_ERR_INVALID; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_err = std::string("hvs_create: no CUDA device (") + cudaGetErrorString(ce) +
                       "); this engine has no CPU fallback";
        cudaGetLastError();
        return HVS_ERR_NO_DEVICE;
    }
    int dev = c.device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) { g_create_err = "hvs_create: device ordinal out of range"; return HVS_ERR_INVALID; }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_err = "hvs_create: device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
        return HVS_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(dev) != cudaSuccess) { g_create_err = "hvs_create: cudaSetDevice failed"; return HVS_ERR_CUDA; }
    hvs_engine *e = new (std::nothrow) hvs_engine();
    if (!e) { g_create_err = "hvs_create: out of host memory"; return HVS_ERR_NOMEM; }
    e->device = dev;
    e->mode = c.mode;
    e->id_offset = c.id_offset;
    e->flags = c.flags;
    if (const char *v = getenv("HVS_MARGIN_AUDIT")) if (v[0] == '1') e->flags |= HVS_FLAG_MARGIN_AUDIT;
    e->sm_count = prop.multiProcessorCount;
    cudaError_t cc = cudaSuccess;
    auto keep = [&](cudaError_t r) { if (cc == cudaSuccess) cc = r; };
    if (c.stream || (c.flags & HVS_FLAG_USE_GIVEN_STREAM)) { e->stream = (cudaStream_t)c.stream; e->own_stream = false; }
    else {
        keep(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->own_stream = cc == cudaSuccess;
    }
    for (auto &ev : e->ev) keep(cudaEventCreate(&ev));
    for (auto &ev : e->evg) keep(cudaEventCreate(&ev));
    for (auto &ev : e->ev_sync) keep(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    keep(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
    keep(cudaStreamCreateWithFlags(&e->stream_up, cudaStreamNonBlocking));
    for (auto &ev : e->ev_up) keep(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (cc == cudaSuccess) cc = init_kernel_attributes();      // dynamic shared-memory opt-in of every kernel, for this device
    if (cc != cudaSuccess) {
        g_create_err = std::string("hvs_create: ") + cudaGetErrorString(cc);
        cudaGetLastError();
        hvs_destroy(e);
        return HVS_ERR_CUDA;
    }
    e->stats.struct_size = sizeof(hvs_stats);
    *out = e;
    return HVS_OK;
}

extern "C" void hvs_destroy(hvs_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->own_stream ? e->stream != nullptr : true) cudaStreamSynchronize(e->stream);
    Index &ix = e->index;
    for (int a = 0; a < 2; ++a) { ix.x[a].release(); ix.ids[a].release(); ix.xnorm[a].release(); ix.xb[a].release(); }
    ix.keys_t.release(); ix.keys_ct.release(); ix.tail.release(); ix.inv_t.release(); ix.outl[0].release(); ix.outl[1].release();
    DevBuf *bufs[] = {&e->d_queries, &e->d_out, &e->d_slices, &e->d_direct_q, &e->d_items, &e->d_item_q, &e->d_tile_q,
                      &e->d_qoff, &e->d_qlists, &e->d_cand, &e->d_cand_cnt, &e->d_scratch, &e->d_flags, &e->d_gthr, &e->d_pool, &e->d_gbest, &e->d_glock,
                      &e->d_work_counter, &e->d_rescore_ids, &e->d_rescore_out, &e->d_audit, &e->d_shard_q, &e->d_shard_sl, &e->d_shard_own, &e->d_split, &e->d_k1acc};
    for (DevBuf *b : bufs) b->release();
    e->h_slices.release(); e->h_flags.release(); e->h_stage_own.release(); e->h_stage.release(); e->h_ingest[0].release(); e->h_ingest[1].release();
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->evg) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->ev_sync) if (ev) cudaEventDestroy(ev);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->stream_up) cudaStreamDestroy(e->stream_up);
    for (auto &ev : e->ev_up) if (ev) cudaEventDestroy(ev);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int hvs_set_mode(hvs_engine *e, uint32_t mode)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (mode > HVS_MODE_TENSOR) EFAIL(HVS_ERR_INVALID, "hvs_set_mode: unknown mode");
    e->mode = mode;
    return HVS_OK;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return ms;
}

extern "C" int hvs_index_build_device(hvs_engine *e, const float *rows_dev, uint32_t n, float sample_proportion)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!rows_dev) EFAIL(HVS_ERR_INVALID, "hvs_index_build: rows is NULL");
    if (n < HVS_K)
        EFAIL(HVS_ERR_INVALID, "hvs_index_build: n < 100 (the reference's pad rule reads nodes[n-s], include/baseline.hpp:138-147)");
    if (!(sample_proportion >= 0.f)) EFAIL(HVS_ERR_INVALID, "hvs_index_build: sample_proportion must be >= 0");
    ECUDA(cudaSetDevice(e->device));
    cudaEventRecord(e->ev[0], e->stream);
    cudaError_t c = index_build_device(e, rows_dev, n, sample_proportion);
    if (c != cudaSuccess) { if (e->err.empty()) e->err = cudaGetErrorString(c); return HVS_ERR_CUDA; }
    cudaEventRecord(e->ev[1], e->stream);
    ECUDA(cudaStreamSynchronize(e->stream));
    e->stats.ms_index_build = ev_ms(e->ev[0], e->ev[1]);
    e->stats.n = e->index.n;
    e->stats.n_total = e->index.n_total;
    e->stats.n_outliers = e->index.n_outl[0];
    return HVS_OK;
}

// ---- solve ----------------------------------------------------------------------------------
// Everything after the slices are known on both sides (d_sl on the device, h_sl on the host): plan, sweeps, finalize.
// e->ev[2] must have been recorded on the stream before the slice search.
static int solve_core(hvs_engine *e, const float *q_dev, uint32_t m, const QSlice *d_sl, const QSlice *h_sl, bool partial,
                      uint32_t *out_ids, float *out_dist, uint32_t *out_count, bool small_launched);

static bool use_device_planner(const hvs_engine *e);
static int solve_core_dev(hvs_engine *e, const float *q_dev, uint32_t m, const QSlice *d_sl, bool partial, uint32_t *out_ids,
                          float *out_dist, uint32_t *out_count, bool small_launched);

// K4s (tiny slices, a warp per query) decides on the device which queries are its own, so it is launched right behind
// the slice search, before the host has seen a slice.  HVS_SMALL=0 switches it off (everything small takes the CTA scan).
static uint32_t small_max()
{
    static const uint32_t v = [] { const char *s = getenv("HVS_SMALL"); return (s && s[0] == '0') ? 0u : SMALL_MAX; }();
    return v;
}
static cudaError_t run_small(hvs_engine *e, const float *q_dev, uint32_t m, const QSlice *d_sl, bool partial, uint32_t *out_ids,
                             float *out_dist, uint32_t *out_count)
{
    cudaEventRecord(e->ev[5], e->stream);
    cudaError_t c = cudaSuccess;
    if (small_max()) {
        c = launch_small(e, q_dev, d_sl, nullptr, m, small_max(), partial, out_ids, out_dist, out_count);
        e->stats.launches++;
    }
    cudaEventRecord(e->ev[6], e->stream);
    return c;
}

static void reset_solve_stats(hvs_engine *e, uint32_t m)
{
    hvs_stats &st = e->stats;
    st.m = m;
    st.pairs = st.pairs_computed = st.rows_union = st.pairs_tile = st.pairs_direct = 0;
    st.n_direct = st.n_tile = st.n_items_ffma = st.n_items_tensor = st.n_fallback = st.launches = 0;
    st.ms_plan = st.ms_direct = st.ms_tile = st.ms_tile_ffma = st.ms_tile_tensor = st.ms_finalize = st.ms_solve_device = 0.f;
    st.margin_audit = 0.f;
}

static int solve_impl(hvs_engine *e, const float *q_dev, uint32_t m, bool partial, uint32_t *out_ids, float *out_dist,
                      uint32_t *out_count)
{
    reset_solve_stats(e, m);
    if (!m) return HVS_OK;
    cudaStream_t s = e->stream;
    ECUDA(e->d_slices.ensure((size_t)m * sizeof(QSlice)));
    ECUDA(e->h_slices.ensure((size_t)m * sizeof(QSlice)));
    QSlice *d_sl = e->d_slices.as<QSlice>();
    cudaEventRecord(e->ev[2], s);
    ECUDA(launch_plan_search(e, q_dev, m, d_sl));
    e->stats.launches++;
    ECUDA(run_small(e, q_dev, m, d_sl, partial, out_ids, out_dist, out_count));   // runs while the slices travel to the host and are planned
    if (use_device_planner(e)) return solve_core_dev(e, q_dev, m, d_sl, partial, out_ids, out_dist, out_count, true);
    ECUDA(cudaMemcpyAsync(e->h_slices.p, d_sl, (size_t)m * sizeof(QSlice), cudaMemcpyDeviceToHost, s));
    ECUDA(cudaStreamSynchronize(s));
    return solve_core(e, q_dev, m, d_sl, e->h_slices.as<QSlice>(), partial, out_ids, out_dist, out_count, true);
}

// ---- the solve with the planner on the device (default): no slice ever travels to the host ------------------------------
static bool use_device_planner(const hvs_engine *e)
{
    static const bool host_forced = [] { const char *v = getenv("HVS_PLAN"); return v && v[0] == 'h'; }();     // HVS_PLAN=host: the host planner (A/B, debugging)
    return !host_forced && e->index.n < (1u << 30);       // sort keys hold [class:2][arena:1][begin:nb][end:nb+1] in 64 bits
}

static int solve_core_dev(hvs_engine *e, const float *q_dev, uint32_t m, const QSlice *d_sl, bool partial, uint32_t *out_ids,
                          float *out_dist, uint32_t *out_count, bool small_launched)
{
    hvs_stats &st = e->stats;
    cudaStream_t s = e->stream;
    if (e->flags & HVS_FLAG_MARGIN_AUDIT) {
        ECUDA(e->d_audit.ensure(4));
        ECUDA(cudaMemsetAsync(e->d_audit.p, 0, 4, s));
    }
    if (!small_launched) ECUDA(run_small(e, q_dev, m, d_sl, partial, out_ids, out_dist, out_count));
    const bool tensor = e->index.xb[0].p != nullptr && e->index.xb[1].p != nullptr && (e->mode == HVS_MODE_AUTO || e->mode == HVS_MODE_TENSOR);
    PlanParams defaults;
    PlanCfg cfg{};
    cfg.need = tensor ? defaults.tensor_min_depth : (double)QT / defaults.direct_cost_ratio;
    cfg.min_tile_pairs = defaults.min_tile_pairs;
    static const long long min_pairs_env = [] { const char *v = getenv("HVS_MIN_TILE_PAIRS"); return v ? atoll(v) : -1ll; }();
    if (min_pairs_env >= 0) cfg.min_tile_pairs = (unsigned long long)min_pairs_env;
    cfg.small_max = small_max();
    cfg.min_tile_len = defaults.min_tile_len;
    cfg.tile_allowed = (e->mode != HVS_MODE_DIRECT && e->index.approx_ok) ? 1u : 0u;
    cfg.force_tile = e->mode == HVS_MODE_TENSOR ? 1u : 0u;
    cfg.bq = tensor ? (uint32_t)QT_TENSOR : (uint32_t)QT;
    static const uint32_t ips = [] { const char *v = getenv("HVS_ITEMS_PER_SM"); int k = v ? atoi(v) : 0; return (uint32_t)(k > 0 ? k : 16); }();
    cfg.items_per_sm = ips;
    cfg.sm_count = (uint32_t)e->sm_count;
    cfg.kind = tensor ? 1u : 0u;
    static const uint32_t ct_r = [] { const char *v = getenv("HVS_CT_R"); long k = v ? atol(v) : -1; return (uint32_t)(k >= 0 ? k : 0); }();
    cfg.ct_min_rows = tensor ? ct_r : 0u;
    // Item order.  2 (default): the items in which slices BEGIN -- cold thresholds, the slow items -- first, then the
    // items that only hold queries begun in earlier chunks, in chunk order, all in ONE launch.  Measured against plain
    // chunk order (0): headline K3 33.9-34.4 -> 32.1 ms, category-only queries 6.6 -> 5.5 ms, BASELINE configs[1]
    // 3.5 -> 3.0 ms, one rank's share of 2 / 4 / 8: 17.5 -> 15.8 / 9.6 -> 9.2 / 6.0 -> 6.1 ms.  1: the same order with a
    // launch boundary between the two sets (slower everywhere: the first launch's idle tail).
    static const uint32_t seed_env = [] { const char *v = getenv("HVS_SEED_PHASE"); return (uint32_t)((v && v[0] >= '0' && v[0] <= '2') ? v[0] - '0' : 2); }();
    cfg.seed_phase = seed_env;
    // A job that cannot reach the tiny-job bound whatever its slices are (m x n pairs at most) needs no plan at all:
    // K4s took the small slices, the CTA scan takes every other query, nothing is read back but the pair count.
    if (small_launched && !cfg.force_tile && (unsigned long long)m * e->index.n < cfg.min_tile_pairs) {
        cudaEventRecord(e->ev[3], s);
        ECUDA(launch_direct(e, q_dev, d_sl, nullptr, m, partial, out_ids, out_dist, out_count, small_max() ? small_max() : 0u));
        st.launches++;
        cudaEventRecord(e->ev[4], s);
        cudaEventRecord(e->ev[9], s);
        ECUDA(e->h_header.ensure(sizeof(PlanHeader)));
        ECUDA(cudaMemcpyAsync(e->h_header.p, e->d_k1acc.p, 16, cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        ECUDA(cudaGetLastError());
        const unsigned long long *acc = e->h_header.as<unsigned long long>();
        st.pairs = st.pairs_direct = st.pairs_computed = acc[0];
        st.n_direct = m;
        const float ms_small = ev_ms(e->ev[5], e->ev[6]);
        st.ms_plan = std::max(0.f, ev_ms(e->ev[2], e->ev[3]) - ms_small);
        st.ms_direct = ev_ms(e->ev[3], e->ev[4]) + ms_small;
        st.ms_solve_device = ev_ms(e->ev[2], e->ev[9]);
        return HVS_OK;
    }
    if (small_launched && small_max()) {
        // Every query small (selective predicates, BASELINE configs[4])?  K4s has then solved the whole batch and the
        // planner has nothing to plan: one 16-byte look at what the slice search counted saves its dozen launches.
        ECUDA(e->h_header.ensure(sizeof(PlanHeader)));
        ECUDA(cudaMemcpyAsync(e->h_header.p, e->d_k1acc.p, 16, cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        const unsigned long long *acc = e->h_header.as<unsigned long long>();
        if (acc[1] == (unsigned long long)m) {
            cudaEventRecord(e->ev[9], s);
            ECUDA(cudaStreamSynchronize(s));
            ECUDA(cudaGetLastError());
            st.pairs = st.pairs_direct = st.pairs_computed = acc[0];
            st.n_direct = m;
            st.ms_direct = ev_ms(e->ev[5], e->ev[6]);
            st.ms_plan = std::max(0.f, ev_ms(e->ev[2], e->ev[9]) - st.ms_direct);
            st.ms_solve_device = ev_ms(e->ev[2], e->ev[9]);
            return HVS_OK;
        }
    }
    PlanHeader h{};
    ECUDA(plan_dev_begin(e, d_sl, m, cfg, &h));                  // the one host round trip of the plan: a 112-byte header
    st.launches += e->pdev.launches;
    st.pairs = h.pairs;
    st.pairs_tile = h.pairs_tile;
    st.pairs_direct = h.pairs - h.pairs_tile;
    st.n_direct = h.n_direct + h.n_small;
    st.n_tile = h.n_tile;
    cudaEventRecord(e->ev[3], s);
    const uint32_t *sorted_q = e->pdev.vals.as<uint32_t>();       // tile queries [0, n_tile), CTA-scan queries [n_tile, n_tile + n_direct)
    if (h.n_direct) {
        ECUDA(launch_direct(e, q_dev, d_sl, sorted_q + h.n_tile, h.n_direct, partial, out_ids, out_dist, out_count));
        st.launches++;
    }
    cudaEventRecord(e->ev[4], s);
    const bool any_items = h.n_items > 0;
    if (any_items) {
        ECUDA(e->d_items.ensure((size_t)h.n_items * sizeof(TileItem)));
        ECUDA(e->d_item_q.ensure((size_t)h.incid * 4));
        ECUDA(e->d_qlists.ensure((size_t)h.incid * 4));
        ECUDA(e->d_cand.ensure((size_t)h.incid * KOUT * 8));
        ECUDA(e->d_cand_cnt.ensure((size_t)h.incid * 4));
        ECUDA(e->d_flags.ensure((size_t)m * 4));
        ECUDA(cudaMemsetAsync(e->d_flags.p, 0, (size_t)m * 4, s));
        ECUDA(e->d_gthr.ensure((size_t)m * 4));
        ECUDA(launch_fill_u32(e, e->d_gthr.as<uint32_t>(), 0xff800000u /* okey(+inf) */, m));
        if (tensor) ECUDA(tile_tensor_begin(e));
        ECUDA(plan_dev_fill(e, d_sl, h, cfg, e->d_item_q.as<uint32_t>(), e->d_items.as<TileItem>(), e->d_qlists.as<uint32_t>()));
        st.launches += 2;
        e->pool_slot = 0;
        cudaEventRecord(e->evg[0], s);
        // seed phase (if planned): the items in which slices begin first, everything else in a second launch behind them
        const uint32_t cut = (cfg.seed_phase == 1 && h.n_seed_items && h.n_seed_items < h.n_items) ? h.n_seed_items : 0u;
        for (int ph = 0; ph < 2; ++ph) {
            const uint32_t ib = ph == 0 ? 0u : cut, ic = ph == 0 ? (cut ? cut : h.n_items) : h.n_items - cut;
            if (ph == 1 && !cut) break;
            if (tensor)
                ECUDA(launch_tile_tensor(e, q_dev, d_sl, e->d_items.as<TileItem>(), ib, ic, e->d_item_q.as<uint32_t>(), e->d_cand.as<uint64_t>(),
                                         e->d_cand_cnt.as<uint32_t>(), e->d_gthr.as<uint32_t>(), e->d_flags.as<uint32_t>()));
            else
                ECUDA(launch_tile_ffma(e, q_dev, d_sl, e->d_items.as<TileItem>(), ib, ic, e->d_item_q.as<uint32_t>(), e->d_cand.as<uint64_t>(),
                                       e->d_cand_cnt.as<uint32_t>(), e->d_gthr.as<uint32_t>(), e->d_flags.as<uint32_t>(), 1.0f));
            if (ph == 0 && cut) { cudaEventRecord(e->evg[2], s); st.launches++; }
        }
        cudaEventRecord(e->evg[1], s);
        st.launches++;
        (tensor ? st.n_items_tensor : st.n_items_ffma) = h.n_items;
        cudaEventRecord(e->ev[7], s);
        ECUDA(launch_finalize(e, q_dev, d_sl, sorted_q, h.n_tile, e->pdev.qoff.as<uint32_t>(), e->d_qlists.as<uint32_t>(),
                              e->d_cand.as<uint64_t>(), e->d_cand_cnt.as<uint32_t>(), e->d_flags.as<uint32_t>(), partial, tensor, out_ids,
                              out_dist, out_count));
        st.launches++;
        cudaEventRecord(e->ev[8], s);
        // queries whose candidate buffers overflowed their margin guarantee are re-solved exactly by K4: the list is built on the device
        ECUDA(e->d_scratch.ensure((size_t)h.n_tile * 4));
        ECUDA(plan_dev_redo(e, e->d_flags.as<uint32_t>(), h.n_tile, e->d_scratch.as<uint32_t>()));
        st.launches++;
        ECUDA(cudaMemcpyAsync(e->h_header.p, e->pdev.header.p, sizeof(PlanHeader), cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        const PlanHeader h2 = *e->h_header.as<PlanHeader>();
        st.pairs_computed = h2.pairs_computed + (h.pairs - h.pairs_tile);
        if (h2.n_redo) {
            ECUDA(launch_direct(e, q_dev, d_sl, e->d_scratch.as<uint32_t>(), h2.n_redo, partial, out_ids, out_dist, out_count));
            st.launches++;
            st.n_fallback = h2.n_redo;
        }
    } else {
        st.pairs_computed = h.pairs;
    }
    cudaEventRecord(e->ev[9], s);
    if (e->flags & HVS_FLAG_MARGIN_AUDIT) ECUDA(cudaMemcpyAsync(&st.margin_audit, e->d_audit.p, 4, cudaMemcpyDeviceToHost, s));
    ECUDA(cudaStreamSynchronize(s));
    ECUDA(cudaGetLastError());
    const float ms_small = ev_ms(e->ev[5], e->ev[6]);
    st.ms_plan = std::max(0.f, ev_ms(e->ev[2], e->ev[3]) - ms_small);
    st.ms_direct = ev_ms(e->ev[3], e->ev[4]) + ms_small;
    if (any_items) {
        const float tile = ev_ms(e->evg[0], e->evg[1]);
        if (tensor) st.ms_tile_tensor = tile; else st.ms_tile_ffma = tile;
        st.ms_tile = tile;
        st.ms_finalize = ev_ms(e->ev[7], e->ev[8]);
    }
    st.ms_solve_device = ev_ms(e->ev[2], e->ev[9]);
    static const bool timeline = getenv("HVS_TIMELINE") != nullptr;
    if (timeline)
        fprintf(stderr, "timeline(dev planner): plan end %.2f  direct end %.2f  sweep [%.2f, %.2f]  finalize [%.2f, %.2f]  end %.2f  (R=%u, %u items of which %u seed, %u tile queries)\n",
                ev_ms(e->ev[2], e->ev[3]), ev_ms(e->ev[2], e->ev[4]), any_items ? ev_ms(e->ev[2], e->evg[0]) : 0.f, any_items ? ev_ms(e->ev[2], e->evg[1]) : 0.f,
                any_items ? ev_ms(e->ev[2], e->ev[7]) : 0.f, any_items ? ev_ms(e->ev[2], e->ev[8]) : 0.f, st.ms_solve_device, h.R, h.n_items, h.n_seed_items, h.n_tile);
    return HVS_OK;
}

static int solve_core(hvs_engine *e, const float *q_dev, uint32_t m, const QSlice *d_sl, const QSlice *h_sl, bool partial,
                      uint32_t *out_ids, float *out_dist, uint32_t *out_count, bool small_launched)
{
    hvs_stats &st = e->stats;
    cudaStream_t s = e->stream;
    if (!small_launched) ECUDA(run_small(e, q_dev, m, d_sl, partial, out_ids, out_dist, out_count));
    if (e->flags & HVS_FLAG_MARGIN_AUDIT) {
        ECUDA(e->d_audit.ensure(4));
        ECUDA(cudaMemsetAsync(e->d_audit.p, 0, 4, s));
    }

    PlanParams pp;
    pp.mode = e->mode;
    pp.tensor_available = e->index.xb[0].p != nullptr && e->index.xb[1].p != nullptr;
    pp.approx_ok = e->index.approx_ok;
    pp.small_max = small_max();
    static const long long min_pairs_env = [] { const char *v = getenv("HVS_MIN_TILE_PAIRS"); return v ? atoll(v) : -1ll; }();
    if (min_pairs_env >= 0) pp.min_tile_pairs = (uint64_t)min_pairs_env;   // tests set 0: tile kernels run on tiny inputs too
    Plan &P = e->plan;
    plan_begin(h_sl, m, pp, P);                      // classify; tile queries are cut into groups of chunk blocks
    st.pairs = P.pairs;
    st.pairs_tile = P.pairs_tile;
    st.pairs_direct = P.pairs - P.pairs_tile;
    st.n_direct = (uint32_t)P.direct_q.size() + P.n_small;

    // one pinned staging buffer for all uploads of this solve; every upload gets its own region
    const size_t n_groups = P.group_end.size();
    const size_t max_items = (size_t)(P.incid / (P.BQ ? P.BQ : 1)) + 2 * (P.R ? ((size_t)e->index.n / P.R + 2) : 0) + 64;
    ECUDA(e->h_stage.ensure(P.direct_q.size() * 4 + max_items * sizeof(TileItem) + (size_t)P.incid * 8 + (size_t)m * 12 + 1024));
    unsigned char *hs = e->h_stage.as<unsigned char>();
    size_t o = 0;
    auto up_on = [&](cudaStream_t st_, void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (!bytes) return cudaSuccess;
        std::memcpy(hs + o, src, bytes);
        cudaError_t c = cudaMemcpyAsync(dst, hs + o, bytes, cudaMemcpyHostToDevice, st_);
        o += (bytes + 15) & ~(size_t)15;
        return c;
    };
    auto up_at = [&](void *dst, const void *src, size_t bytes) -> cudaError_t { return up_on(s, dst, src, bytes); };
    ECUDA(e->d_direct_q.ensure(P.direct_q.size() * 4 + 16));
    ECUDA(up_at(e->d_direct_q.p, P.direct_q.data(), P.direct_q.size() * 4));
    cudaEventRecord(e->ev[3], s);

    // K4: direct scans
    if (!P.direct_q.empty()) {
        ECUDA(launch_direct(e, q_dev, d_sl, e->d_direct_q.as<uint32_t>(), (uint32_t)P.direct_q.size(), partial, out_ids, out_dist, out_count));
        st.launches++;
    }
    cudaEventRecord(e->ev[4], s);

    // K2 / K3: tile sweeps, group by group -- the GPU sweeps group g while the host plans group g + 1 -- then K5
    bool any_items = false, side_lane_used = false;
    bool ran[8] = {false, false, false, false, false, false, false, false};
    if (P.incid) {
        any_items = true;
        ECUDA(e->d_items.ensure(max_items * sizeof(TileItem)));
        ECUDA(e->d_item_q.ensure((size_t)P.incid * 4));
        ECUDA(e->d_cand.ensure((size_t)P.incid * KOUT * 8));
        ECUDA(e->d_cand_cnt.ensure((size_t)P.incid * 4));
        ECUDA(e->d_flags.ensure((size_t)m * 4));
        ECUDA(cudaMemsetAsync(e->d_flags.p, 0, (size_t)m * 4, s));
        ECUDA(e->d_gthr.ensure((size_t)m * 4));
        ECUDA(launch_fill_u32(e, e->d_gthr.as<uint32_t>(), 0xff800000u /* okey(+inf) */, m));
        st.launches++;
        if (P.tensor) ECUDA(tile_tensor_begin(e));
        TileItem *d_items = e->d_items.as<TileItem>();
        cudaEventRecord(e->ev_sync[0], s);                         // setup of this solve is enqueued up to here
        for (size_t g = 0; g < n_groups; ++g) {
            uint32_t ib = 0, ie = 0;
            const size_t iq0 = P.item_q.size();
            plan_group(h_sl, P, g, ib, ie);
            if (ie == ib) continue;
            if (ie > max_items) EFAIL(HVS_ERR_STATE, "planner produced more items than it announced");
            // The work lists go up on their own stream: queued on a launch lane they would wait for the sweep
            // running there, and the next group could not start before the previous one had drained.
            cudaStream_t su = e->stream_up && g < 8 ? e->stream_up : s;
            ECUDA(up_on(su, d_items + ib, P.items.data() + ib, (size_t)(ie - ib) * sizeof(TileItem)));
            ECUDA(up_on(su, e->d_item_q.as<uint32_t>() + iq0, P.item_q.data() + iq0, (P.item_q.size() - iq0) * 4));
            // odd groups run on a second stream with their own survivor pools, so that the tail of one launch
            // (CTAs running out of items) overlaps the head of the next
            cudaStream_t sg = (g & 1) && e->stream2 ? e->stream2 : s;
            if (su != s) {
                cudaEventRecord(e->ev_up[g], su);
                ECUDA(cudaStreamWaitEvent(sg, e->ev_up[g], 0));
            }
            if (sg != s) ECUDA(cudaStreamWaitEvent(sg, e->ev_sync[0], 0));   // the solve's setup (memsets, fills) on the main lane
            e->pool_slot = (uint32_t)(g & 1);
            cudaStream_t saved = e->stream;
            e->stream = sg;
            if (g < 8) cudaEventRecord(e->evg[2 * g], sg);
            cudaError_t lc;
            if (P.tensor)
                lc = launch_tile_tensor(e, q_dev, d_sl, d_items, ib, ie - ib, e->d_item_q.as<uint32_t>(), e->d_cand.as<uint64_t>(),
                                        e->d_cand_cnt.as<uint32_t>(), e->d_gthr.as<uint32_t>(), e->d_flags.as<uint32_t>());
            else
                lc = launch_tile_ffma(e, q_dev, d_sl, d_items, ib, ie - ib, e->d_item_q.as<uint32_t>(), e->d_cand.as<uint64_t>(),
                                      e->d_cand_cnt.as<uint32_t>(), e->d_gthr.as<uint32_t>(), e->d_flags.as<uint32_t>(), 1.0f);
            if (g < 8) { cudaEventRecord(e->evg[2 * g + 1], sg); ran[g] = true; }
            e->stream = saved;
            ECUDA(lc);
            if (sg != s) { cudaEventRecord(e->ev_sync[1], sg); side_lane_used = true; }
            st.launches++;
        }
        if (side_lane_used) ECUDA(cudaStreamWaitEvent(s, e->ev_sync[1], 0));   // K5 needs every group's lists
        plan_finish(h_sl, m, P);                     // candidate-list CSR, built while the last group is swept
        st.pairs_computed = P.pairs_computed;
        st.n_tile = (uint32_t)P.tile_q.size();
        st.n_items_ffma = P.n_ffma;
        st.n_items_tensor = P.n_tensor;
        ECUDA(e->d_tile_q.ensure(P.tile_q.size() * 4 + 16));
        ECUDA(e->d_qoff.ensure(P.q_list_off.size() * 4 + 16));
        ECUDA(e->d_qlists.ensure(P.q_lists.size() * 4 + 16));
        ECUDA(up_at(e->d_tile_q.p, P.tile_q.data(), P.tile_q.size() * 4));
        ECUDA(up_at(e->d_qoff.p, P.q_list_off.data(), P.q_list_off.size() * 4));
        ECUDA(up_at(e->d_qlists.p, P.q_lists.data(), P.q_lists.size() * 4));
        cudaEventRecord(e->ev[7], s);
        ECUDA(launch_finalize(e, q_dev, d_sl, e->d_tile_q.as<uint32_t>(), (uint32_t)P.tile_q.size(), e->d_qoff.as<uint32_t>(),
                              e->d_qlists.as<uint32_t>(), e->d_cand.as<uint64_t>(), e->d_cand_cnt.as<uint32_t>(),
                              e->d_flags.as<uint32_t>(), partial, P.n_tensor > 0, out_ids, out_dist, out_count));
        st.launches++;
        cudaEventRecord(e->ev[8], s);
        // queries whose candidate buffers overflowed their margin guarantee are re-solved exactly by K4
        ECUDA(e->h_flags.ensure((size_t)m * 4));
        ECUDA(cudaMemcpyAsync(e->h_flags.p, e->d_flags.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        const uint32_t *fl = e->h_flags.as<uint32_t>();
        std::vector<uint32_t> redo;
        for (uint32_t q : P.tile_q) if (fl[q]) redo.push_back(q);
        if (!redo.empty()) {
            ECUDA(e->d_scratch.ensure(redo.size() * 4));
            ECUDA(cudaMemcpyAsync(e->d_scratch.p, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, s));
            ECUDA(launch_direct(e, q_dev, d_sl, e->d_scratch.as<uint32_t>(), (uint32_t)redo.size(), partial, out_ids, out_dist, out_count));
            ECUDA(cudaStreamSynchronize(s));   // redo lives on the host stack
            st.launches++;
            st.n_fallback = (uint32_t)redo.size();
        }
    }
    cudaEventRecord(e->ev[9], s);
    if (e->flags & HVS_FLAG_MARGIN_AUDIT) ECUDA(cudaMemcpyAsync(&st.margin_audit, e->d_audit.p, 4, cudaMemcpyDeviceToHost, s));
    ECUDA(cudaStreamSynchronize(s));
    ECUDA(cudaGetLastError());
    const float ms_small = ev_ms(e->ev[5], e->ev[6]);
    st.ms_plan = std::max(0.f, ev_ms(e->ev[2], e->ev[3]) - ms_small);   // slice search + classification (K4s runs under it); group planning overlaps the sweeps
    st.ms_direct = ev_ms(e->ev[3], e->ev[4]) + ms_small;
    if (any_items) {
        // the groups' launches overlap (two lanes): the tile phase is the span from the first start to the last end
        float tile = 0.f;
        int g0 = -1;
        for (size_t g = 0; g < n_groups && g < 8; ++g)
            if (ran[g]) {
                if (g0 < 0) g0 = (int)g;
                tile = std::max(tile, ev_ms(e->evg[2 * g0], e->evg[2 * g + 1]));
            }
        if (P.tensor) st.ms_tile_tensor = tile; else st.ms_tile_ffma = tile;
        st.ms_tile = tile;
        st.ms_finalize = ev_ms(e->ev[7], e->ev[8]);
    }
    st.ms_solve_device = ev_ms(e->ev[2], e->ev[9]);
    static const bool timeline = getenv("HVS_TIMELINE") != nullptr;
    if (timeline) {                                  // developer aid: when each phase ran, ms since the start of the solve
        fprintf(stderr, "timeline: plan end %.2f  direct end %.2f ", ev_ms(e->ev[2], e->ev[3]), ev_ms(e->ev[2], e->ev[4]));
        for (size_t g = 0; g < n_groups && g < 8; ++g)
            if (ran[g]) fprintf(stderr, " group %zu [%.2f, %.2f]", g, ev_ms(e->ev[2], e->evg[2 * g]), ev_ms(e->ev[2], e->evg[2 * g + 1]));
        if (any_items) fprintf(stderr, "  finalize [%.2f, %.2f]", ev_ms(e->ev[2], e->ev[7]), ev_ms(e->ev[2], e->ev[8]));
        fprintf(stderr, "  end %.2f\n", st.ms_solve_device);
    }
    return HVS_OK;
}

static int check_solve_args(hvs_engine *e, const void *q, uint32_t m, const void *out)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!e->index.built) EFAIL(HVS_ERR_STATE, "hvs_solve: hvs_index_build has not succeeded on this engine");
    if (m && (!q || !out)) EFAIL(HVS_ERR_INVALID, "hvs_solve: NULL buffer");
    if (cudaSetDevice(e->device) != cudaSuccess) EFAIL(HVS_ERR_CUDA, "cudaSetDevice failed");
    return HVS_OK;
}

extern "C" int hvs_solve_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t *out_ids_dev)
{
    int rc = check_solve_args(e, queries_dev, m, out_ids_dev);
    if (rc) return rc;
    auto t0 = std::chrono::steady_clock::now();
    e->stats.ms_h2d = e->stats.ms_d2h = 0.f;
    rc = solve_impl(e, queries_dev, m, false, out_ids_dev, nullptr, nullptr);
    e->stats.ms_solve_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" int hvs_solve(hvs_engine *e, const float *queries_host, uint32_t m, uint32_t *out_ids_host)
{
    int rc = check_solve_args(e, queries_host, m, out_ids_host);
    if (rc) return rc;
    if (!m) return HVS_OK;
    auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = e->stream;
    ECUDA(e->d_queries.ensure((size_t)m * QROW * 4));
    ECUDA(e->d_out.ensure((size_t)m * K * 4));
    cudaEventRecord(e->ev[0], s);
    ECUDA(cudaMemcpyAsync(e->d_queries.p, queries_host, (size_t)m * QROW * 4, cudaMemcpyHostToDevice, s));
    cudaEventRecord(e->ev[1], s);
    rc = solve_impl(e, e->d_queries.as<float>(), m, false, e->d_out.as<uint32_t>(), nullptr, nullptr);
    if (rc) return rc;
    cudaEventRecord(e->ev[10], s);
    ECUDA(cudaMemcpyAsync(out_ids_host, e->d_out.p, (size_t)m * K * 4, cudaMemcpyDeviceToHost, s));
    cudaEventRecord(e->ev[11], s);
    ECUDA(cudaStreamSynchronize(s));
    e->stats.ms_h2d = ev_ms(e->ev[0], e->ev[1]);
    e->stats.ms_d2h = ev_ms(e->ev[10], e->ev[11]);
    e->stats.ms_solve_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return HVS_OK;
}

// Query-sharded solve: this rank's share of one batch (see shard_assign, hvs_plan.cu).
extern "C" int hvs_solve_shard_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t rank, uint32_t world,
                                      uint32_t *out_ids_dev, uint32_t *out_order_host, uint32_t *out_counts_host)
{
    int rc = check_solve_args(e, queries_dev, m, out_ids_dev);
    if (rc) return rc;
    if (!world || world > 255 || rank >= world) EFAIL(HVS_ERR_INVALID, "hvs_solve_shard_device: need rank < world <= 255");
    if (!out_counts_host) EFAIL(HVS_ERR_INVALID, "hvs_solve_shard_device: NULL buffer");
    auto t0 = std::chrono::steady_clock::now();
    e->stats.ms_h2d = e->stats.ms_d2h = 0.f;
    reset_solve_stats(e, 0);
    for (uint32_t r = 0; r < world; ++r) out_counts_host[r] = 0;
    if (!m) return HVS_OK;
    cudaStream_t s = e->stream;
    ECUDA(e->d_slices.ensure((size_t)m * sizeof(QSlice)));
    ECUDA(e->h_slices.ensure((size_t)m * sizeof(QSlice)));
    QSlice *d_sl_all = e->d_slices.as<QSlice>();
    cudaEventRecord(e->ev[2], s);
    ECUDA(launch_plan_search(e, queries_dev, m, d_sl_all));                 // every rank resolves ALL predicates (two binary searches each)
    const uint32_t stripes = shard_stripes(m, world);
    // the assignment runs on the device unless its 64-bit cost arithmetic could overflow (m x n beyond ~10^15) or the host planner is forced
    const bool on_device = use_device_planner(e) && (double)m * ((double)e->index.n + (double)shard_query_cost()) * world * stripes < 9.0e18;
    const uint32_t *own_dev = nullptr;
    uint32_t off = 0, m_own = 0;
    uint32_t sa_launches = 0;
    if (on_device) {
        uint32_t *order_dev = nullptr, *counts_dev = nullptr;
        ECUDA(shard_assign_dev(e, d_sl_all, m, world, stripes, &order_dev, &counts_dev));
        sa_launches = e->pdev.launches;
        ECUDA(e->h_stage_own.ensure((size_t)m * 4 + 256 * 4));
        uint32_t *h_order = e->h_stage_own.as<uint32_t>(), *h_counts = h_order + m;
        if (out_order_host) ECUDA(cudaMemcpyAsync(h_order, order_dev, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        ECUDA(cudaMemcpyAsync(h_counts, counts_dev, (size_t)world * 4, cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        if (out_order_host) std::memcpy(out_order_host, h_order, (size_t)m * 4);
        std::memcpy(out_counts_host, h_counts, (size_t)world * 4);
        for (uint32_t r = 0; r < rank; ++r) off += out_counts_host[r];
        m_own = out_counts_host[rank];
        own_dev = order_dev + off;
    } else {
        ECUDA(cudaMemcpyAsync(e->h_slices.p, d_sl_all, (size_t)m * sizeof(QSlice), cudaMemcpyDeviceToHost, s));
        ECUDA(cudaStreamSynchronize(s));
        e->h_shard_order.resize(m);
        if (!out_order_host) out_order_host = e->h_shard_order.data();
        shard_assign(e->h_slices.as<QSlice>(), m, world, out_order_host, out_counts_host);        // the same on every rank
        for (uint32_t r = 0; r < rank; ++r) off += out_counts_host[r];
        m_own = out_counts_host[rank];
        // keep the assignment on the device too (hvs_shard_scatter_device)
        ECUDA(e->pdev.sa_order.ensure((size_t)m * 4));
        ECUDA(e->pdev.sa_counts.ensure(256 * 4));
        ECUDA(cudaMemcpyAsync(e->pdev.sa_order.p, out_order_host, (size_t)m * 4, cudaMemcpyHostToDevice, s));
        ECUDA(cudaMemcpyAsync(e->pdev.sa_counts.p, out_counts_host, (size_t)world * 4, cudaMemcpyHostToDevice, s));
        ECUDA(cudaStreamSynchronize(s));            // caller-owned pageable arrays
        ECUDA(e->d_shard_own.ensure((size_t)m_own * 4 + 16));
        ECUDA(e->h_stage_own.ensure((size_t)m_own * 4 + 16));
        std::memcpy(e->h_stage_own.p, out_order_host + off, (size_t)m_own * 4);
        ECUDA(cudaMemcpyAsync(e->d_shard_own.p, e->h_stage_own.p, (size_t)m_own * 4, cudaMemcpyHostToDevice, s));
        own_dev = e->d_shard_own.as<uint32_t>();
    }
    e->shard_m = m;
    e->shard_world = world;
    reset_solve_stats(e, m_own);
    e->stats.launches = 1u + sa_launches;                   // the slice search over all m, the assignment's kernels
    if (m_own) {
        ECUDA(e->d_shard_q.ensure((size_t)m_own * QROW * 4));
        ECUDA(e->d_shard_sl.ensure((size_t)m_own * sizeof(QSlice)));
        ECUDA(launch_gather_queries(e, queries_dev, d_sl_all, own_dev, m_own, e->d_shard_q.as<float>(), e->d_shard_sl.as<QSlice>()));
        e->stats.launches++;
        if (use_device_planner(e))
            rc = solve_core_dev(e, e->d_shard_q.as<float>(), m_own, e->d_shard_sl.as<QSlice>(), false, out_ids_dev, nullptr, nullptr, false);
        else {
            const QSlice *h_all = e->h_slices.as<QSlice>();
            const uint32_t *own = out_order_host + off;
            e->h_shard_sl.resize(m_own);
            for (uint32_t i = 0; i < m_own; ++i) e->h_shard_sl[i] = h_all[own[i]];
            rc = solve_core(e, e->d_shard_q.as<float>(), m_own, e->d_shard_sl.as<QSlice>(), e->h_shard_sl.data(), false, out_ids_dev,
                            nullptr, nullptr, false);
        }
    }
    e->stats.ms_solve_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" int hvs_shard_scatter_device(hvs_engine *e, const uint32_t *gathered_dev, uint32_t cap, uint32_t *out_ids_dev)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!e->shard_m || !e->shard_world) EFAIL(HVS_ERR_STATE, "hvs_shard_scatter_device: no hvs_solve_shard_device before it");
    if (!gathered_dev || !out_ids_dev || !cap) EFAIL(HVS_ERR_INVALID, "hvs_shard_scatter_device: NULL buffer or cap == 0");
    ECUDA(cudaSetDevice(e->device));
    ECUDA(launch_shard_scatter(e, gathered_dev, cap, e->pdev.sa_order.as<uint32_t>(), e->pdev.sa_counts.as<uint32_t>(), e->shard_m,
                               e->shard_world, out_ids_dev));
    return HVS_OK;
}

extern "C" int hvs_solve_full(hvs_engine *e, const float *queries_host, uint32_t m, uint32_t *out_ids_host, float *out_dist_host)
{
    int rc = check_solve_args(e, queries_host, m, out_ids_host);
    if (rc) return rc;
    if (m && !out_dist_host) EFAIL(HVS_ERR_INVALID, "hvs_solve_full: NULL buffer");
    if (!m) return HVS_OK;
    auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = e->stream;
    ECUDA(e->d_queries.ensure((size_t)m * QROW * 4));
    ECUDA(e->d_out.ensure((size_t)m * K * 4));
    ECUDA(e->d_rescore_out.ensure((size_t)m * K * 4));
    cudaEventRecord(e->ev[0], s);
    ECUDA(cudaMemcpyAsync(e->d_queries.p, queries_host, (size_t)m * QROW * 4, cudaMemcpyHostToDevice, s));
    cudaEventRecord(e->ev[1], s);
    rc = solve_impl(e, e->d_queries.as<float>(), m, false, e->d_out.as<uint32_t>(), nullptr, nullptr);
    if (rc) return rc;
    ECUDA(launch_rescore(e, e->d_queries.as<float>(), m, e->d_out.as<uint32_t>(), e->d_rescore_out.as<float>(), nullptr));
    cudaEventRecord(e->ev[10], s);
    ECUDA(cudaMemcpyAsync(out_ids_host, e->d_out.p, (size_t)m * K * 4, cudaMemcpyDeviceToHost, s));
    ECUDA(cudaMemcpyAsync(out_dist_host, e->d_rescore_out.p, (size_t)m * K * 4, cudaMemcpyDeviceToHost, s));
    cudaEventRecord(e->ev[11], s);
    ECUDA(cudaStreamSynchronize(s));
    e->stats.ms_h2d = ev_ms(e->ev[0], e->ev[1]);
    e->stats.ms_d2h = ev_ms(e->ev[10], e->ev[11]);
    e->stats.ms_solve_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return HVS_OK;
}

extern "C" int hvs_solve_partial_device(hvs_engine *e, const float *queries_dev, uint32_t m, float *out_dist_dev,
                                        uint32_t *out_ids_dev, uint32_t *out_count_dev)
{
    int rc = check_solve_args(e, queries_dev, m, out_ids_dev);
    if (rc) return rc;
    if (m && (!out_dist_dev || !out_count_dev)) EFAIL(HVS_ERR_INVALID, "hvs_solve_partial_device: NULL buffer");
    auto t0 = std::chrono::steady_clock::now();
    rc = solve_impl(e, queries_dev, m, true, out_ids_dev, out_dist_dev, out_count_dev);
    e->stats.ms_solve_wall = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

extern "C" int hvs_merge_partials_device(hvs_engine *e, const float *queries_dev, uint32_t m, uint32_t g,
                                         const float *dist_dev, const uint32_t *ids_dev, const uint32_t *count_dev,
                                         const float *tail_rows_dev, uint32_t n_total, uint32_t *out_ids_dev)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!m) return HVS_OK;
    if (!queries_dev || !dist_dev || !ids_dev || !count_dev || !tail_rows_dev || !out_ids_dev || !g)
        EFAIL(HVS_ERR_INVALID, "hvs_merge_partials_device: NULL buffer or g == 0");
    if (n_total < HVS_K) EFAIL(HVS_ERR_INVALID, "hvs_merge_partials_device: n_total < 100");
    ECUDA(cudaSetDevice(e->device));
    ECUDA(launch_merge_partials(e, queries_dev, m, g, dist_dev, ids_dev, count_dev, tail_rows_dev, n_total, out_ids_dev));
    ECUDA(cudaStreamSynchronize(e->stream));
    return HVS_OK;
}

extern "C" int hvs_rescore(hvs_engine *e, const float *queries_host, uint32_t m, const uint32_t *ids_host, float *out_dist_host)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!e->index.built) EFAIL(HVS_ERR_STATE, "hvs_rescore: no index");
    if (!m) return HVS_OK;
    if (!queries_host || !ids_host || !out_dist_host) EFAIL(HVS_ERR_INVALID, "hvs_rescore: NULL buffer");
    ECUDA(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    ECUDA(e->d_queries.ensure((size_t)m * QROW * 4));
    ECUDA(e->d_rescore_ids.ensure((size_t)m * K * 4));
    ECUDA(e->d_rescore_out.ensure((size_t)m * K * 4));
    ECUDA(cudaMemcpyAsync(e->d_queries.p, queries_host, (size_t)m * QROW * 4, cudaMemcpyHostToDevice, s));
    ECUDA(cudaMemcpyAsync(e->d_rescore_ids.p, ids_host, (size_t)m * K * 4, cudaMemcpyHostToDevice, s));
    ECUDA(launch_rescore(e, e->d_queries.as<float>(), m, e->d_rescore_ids.as<uint32_t>(), e->d_rescore_out.as<float>(), nullptr));
    ECUDA(cudaMemcpyAsync(out_dist_host, e->d_rescore_out.p, (size_t)m * K * 4, cudaMemcpyDeviceToHost, s));
    ECUDA(cudaStreamSynchronize(s));
    return HVS_OK;
}

extern "C" int hvs_get_stats(const hvs_engine *e, hvs_stats *out)
{
    if (!e || !out) return HVS_ERR_INVALID;
    uint32_t sz = out->struct_size;
    if (sz < 8 || sz > sizeof(hvs_stats)) sz = sizeof(hvs_stats);
    hvs_stats tmp = e->stats;
    tmp.struct_size = sz;
    std::memcpy(out, &tmp, sz);
    return HVS_OK;
}

extern "C" int hvs_measure_ffma_peak(hvs_engine *e, uint32_t iters, float *out_tflops, float *out_sm_mhz)
{
    if (!e || !out_tflops) return HVS_ERR_INVALID;
    e->err.clear();
    ECUDA(cudaSetDevice(e->device));
    float mhz = 0.f;
    ECUDA(measure_ffma_peak(e, iters ? iters : 5, out_tflops, &mhz));
    if (out_sm_mhz) *out_sm_mhz = mhz;
    return HVS_OK;
}
