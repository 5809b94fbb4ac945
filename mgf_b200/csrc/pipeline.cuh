// pipeline.cuh -- World::step for callers that stream state to the host every step (rendering, logging,
// a host-side controller): mgfb_step_enqueue queues  [H2D v, omega] -> step -> [pack x, q, v, omega -> D2H]
// and returns at once; mgfb_step_wait blocks until the OLDEST queued step's outputs are in the caller's
// buffers.  Up to MGFB_PIPE_DEPTH steps may be in flight, so the PCIe transfers of step k (separate copy streams, both
// directions at once) overlap the kernels of step k+1.  Same kernels, same results as mgfb_step.  On a tiled
// world every rank enqueues and waits the same sequence (the tiles meet on the device, as with mgfb_step_n).
// Included at the end of capi.cu.
#pragma once

struct PipeSlot {
    Buf in, out;                       // device staging: v, omega in; x, q, v, omega out
    Counters* h_ctr = nullptr;         // pinned copy of the step's counters
    Counters* d_ctr = nullptr;         // device copy made by k_step_done; travels to h_ctr on the D2H stream
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // step begin / end, solve begin / end
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    // what was asked for, kept so that the step can be re-queued after a work list was regrown
    float dt = 0.0f; unsigned iters = 0, input_mode = 0; bool has_in = false;
    float *x_out = nullptr, *q_out = nullptr, *v_out = nullptr, *w_out = nullptr;
};

void pipe_destroy(mgfb_ctx* ctx) {
    if (!ctx->pipe) return;
    for (int k = 0; k < MGFB_PIPE_DEPTH; ++k) {
        PipeSlot& s = ctx->pipe[k];
        release(s.in); release(s.out);
        if (s.h_ctr) cudaFreeHost(s.h_ctr);
        if (s.d_ctr) cudaFree(s.d_ctr);
        for (auto& e : s.ev) if (e) cudaEventDestroy(e);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    delete[] ctx->pipe; ctx->pipe = nullptr; ctx->pipe_pending = nullptr;
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    ctx->s_h2d = ctx->s_d2h = nullptr;
}

namespace {
int32_t pipe_init(mgfb_ctx* ctx) {
    if (ctx->pipe) return MGFB_OK;
    ctx->pipe = new PipeSlot[MGFB_PIPE_DEPTH];
    CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < MGFB_PIPE_DEPTH; ++k) {
        PipeSlot& s = ctx->pipe[k];
        CU(cudaMallocHost(&s.h_ctr, sizeof(Counters)));
        CU(cudaMalloc(&s.d_ctr, sizeof(Counters)));
        for (auto& e : s.ev) CU(cudaEventCreate(&e));
        CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        // mgfb_step_wait sleeps on this one instead of spinning (unless MGFB_PIPE_SPIN=1): with one process per GPU a spinning waiter
        // per rank takes the cores the other ranks need to keep THEIR launch queues full, and tiles run in lock-step
        const char* spin = getenv("MGFB_PIPE_SPIN");
        CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming | ((spin && spin[0] == '1') ? 0u : (unsigned)cudaEventBlockingSync)));
    }
    return MGFB_OK;
}
}  // namespace

namespace {
// Host transfers of a queued step's results: after its own kernels (ev_done) and, when `after` is given, not before that
// event.  Nothing host-bound sits in the step's own stream (the one D2H copy engine is busy with the previous step's state
// for 0.1-0.4 ms, and a copy queued behind it there would hold the next step back by as much).
// On a TILED world the transfers of step k are held until step k+1 reaches its solver (`after` = that step's solve-begin
// event), and the inputs of step k+2 likewise: the front half of a step is a chain of peer hand-shakes (ghosts, link tables),
// each a system-scope fence + flag, and a fence issued while PCIe is saturated by a 5 MB transfer waits for it (measured on
// 8 GPUs sharing PCIe: +0.3 ms of device time per step with the transfers running freely, tools/e2e_diag.py); the solver's
// own peer hand-overs are fence-free (one 32-byte store each) and do not care.
bool pipe_defer(const mgfb_ctx* ctx) {
    static const int env = [] { const char* e = getenv("MGFB_PIPE_DEFER"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    return env < 0 ? ctx->tiled : env == 1;
}
int32_t pipe_issue_d2h(mgfb_ctx* ctx, PipeSlot& s, cudaEvent_t after) {
    const unsigned n = ctx->n;
    float* sx = s.out.as<float>(); float* sq = sx + (size_t)3 * n; float* sv = sq + (size_t)4 * n; float* sw = sv + (size_t)3 * n;
    CU(cudaStreamWaitEvent(ctx->s_d2h, s.ev_done, 0));
    if (after) CU(cudaStreamWaitEvent(ctx->s_d2h, after, 0));
    CU(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (s.x_out) CU(cudaMemcpyAsync(s.x_out, sx, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (s.q_out) CU(cudaMemcpyAsync(s.q_out, sq, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (s.v_out) CU(cudaMemcpyAsync(s.v_out, sv, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (s.w_out) CU(cudaMemcpyAsync(s.w_out, sw, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU(cudaEventRecord(s.ev_out, ctx->s_d2h));
    if (ctx->pipe_pending == &s) ctx->pipe_pending = nullptr;
    return MGFB_OK;
}
int32_t pipe_flush_pending(mgfb_ctx* ctx) { return ctx->pipe_pending ? pipe_issue_d2h(ctx, *ctx->pipe_pending, nullptr) : (int32_t)MGFB_OK; }
// Device side of one queued step: [inputs applied] -> step -> counters -> pack -> D2H.  `from_integrate` = false re-runs only the
// part after integration (the step overflowed a work list after integrating; the lists have been regrown since).
int32_t pipe_launch(mgfb_ctx* ctx, PipeSlot& s, bool from_integrate) {
    const unsigned n = ctx->n;
    if (s.has_in && from_integrate) {
        float* sv = s.in.as<float>(); float* sw = sv + (size_t)3 * n;
        CU(cudaStreamWaitEvent(ctx->stream, s.ev_in, 0));
        if (s.input_mode == MGFB_INPUT_ADD) k_set_velocity<true><<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, sv, sw, dctr(ctx));
        else k_set_velocity<false><<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, sv, sw, dctr(ctx));
        ctx->launches += 1;
    }
    ctx->cur_ev = s.ev; ctx->ctr_snap = s.d_ctr;
    CU(cudaEventRecord(s.ev[0], ctx->stream));
    int32_t st = enqueue_step(ctx, s.dt, s.iters, from_integrate, true);
    ctx->cur_ev = ctx->ev; ctx->ctr_snap = nullptr;
    TRY(st);
    CU(cudaEventRecord(s.ev[1], ctx->stream));
    float* sx = s.out.as<float>(); float* sq = sx + (size_t)3 * n; float* sv = sq + (size_t)4 * n; float* sw = sv + (size_t)3 * n;
    if (s.x_out || s.q_out || s.v_out || s.w_out) {
        k_pack_state<<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, s.x_out ? sx : nullptr,
                                                                                               s.q_out ? sq : nullptr, s.v_out ? sv : nullptr, s.w_out ? sw : nullptr);
        ctx->launches += 1;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(s.ev_done, ctx->stream));
    if (pipe_defer(ctx)) {
        // tiled world: the previous step's state goes to the host while THIS step's solver runs (see pipe_issue_d2h)
        if (ctx->pipe_pending && ctx->pipe_pending != &s) TRY(pipe_issue_d2h(ctx, *ctx->pipe_pending, s.ev[2]));
        ctx->pipe_pending = &s;
    } else TRY(pipe_issue_d2h(ctx, s, nullptr));
    return MGFB_OK;
}
int32_t pipe_ensure_lists(mgfb_ctx* ctx) {
    TRY(ensure_step_buffers(ctx, ctx->pipe_scale));
    TRY(ensure_grid(ctx, ctx->pipe_scale));
    TRY(ensure_rows(ctx, ctx->contact_cap, false, 4096));
    return MGFB_OK;
}
}  // namespace

extern "C" {
int32_t mgfb_step_enqueue(mgfb_ctx* ctx, float dt, uint32_t iters, uint32_t input_mode, const float* v_in, const float* omega_in, float* x_out,
                          float* q_out, float* v_out, float* omega_out) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (!(dt > 0.0f)) return fail(ctx, MGFB_ERR_INVALID_ARG, "dt must be > 0");
    if ((v_in == nullptr) != (omega_in == nullptr)) return fail(ctx, MGFB_ERR_INVALID_ARG, "v_in and omega_in go together");
    if (input_mode > MGFB_INPUT_ADD) return fail(ctx, MGFB_ERR_INVALID_ARG, "unknown input mode");
    if (ctx->n == 0) return fail(ctx, MGFB_ERR_STATE, "no bodies");
    if (ctx->pipe_inflight >= MGFB_PIPE_DEPTH) return fail(ctx, MGFB_ERR_STATE, "the pipeline is full: mgfb_step_wait first");
    CU(cudaSetDevice(ctx->device));
    TRY(pipe_init(ctx));
    // a work list cannot be regrown under a step that is running: sized generously here, and when one overflows all the same
    // mgfb_step_wait drains the pipeline, regrows and re-queues (a tile cannot: its neighbours hold pointers into the lists)
    TRY(pipe_ensure_lists(ctx));
    const unsigned n = ctx->n;
    PipeSlot& s = ctx->pipe[(ctx->pipe_head + ctx->pipe_inflight) % MGFB_PIPE_DEPTH];
    TRY(ensure(ctx, s.in, (size_t)n * 24)); TRY(ensure(ctx, s.out, (size_t)n * 52));
    s.dt = dt; s.iters = iters; s.input_mode = input_mode; s.has_in = v_in != nullptr;
    s.x_out = x_out; s.q_out = q_out; s.v_out = v_out; s.w_out = omega_out;
    if (v_in) {
        float* sv = s.in.as<float>(); float* sw = sv + (size_t)3 * n;
        if (pipe_defer(ctx) && ctx->pipe_pending) CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->pipe_pending->ev[2], 0));   // ride the youngest queued step's solver
        CU(cudaMemcpyAsync(sv, v_in, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU(cudaMemcpyAsync(sw, omega_in, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU(cudaEventRecord(s.ev_in, ctx->s_h2d));
    }
    TRY(pipe_launch(ctx, s, true));
    ctx->pipe_inflight++;
    return MGFB_OK;
}

int32_t mgfb_step_wait(mgfb_ctx* ctx, mgfb_step_stats* stats) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (!ctx->pipe || ctx->pipe_inflight == 0) return fail(ctx, MGFB_ERR_STATE, "no step in flight");
    CU(cudaSetDevice(ctx->device));
    PipeSlot& s = ctx->pipe[ctx->pipe_head % MGFB_PIPE_DEPTH];
    if (ctx->pipe_pending == &s) TRY(pipe_flush_pending(ctx));   // no younger step was queued behind it: nothing to wait for
    CU(cudaEventSynchronize(s.ev_out));
    auto drain = [&]() {   // after a sticky flag every queued kernel is a no-op: wait them out
        pipe_flush_pending(ctx);
        for (unsigned k = 1; k < ctx->pipe_inflight; ++k) cudaEventSynchronize(ctx->pipe[(ctx->pipe_head + k) % MGFB_PIPE_DEPTH].ev_out);
        cudaStreamSynchronize(ctx->stream);
    };
    auto drop_all = [&]() { ctx->pipe_head += ctx->pipe_inflight; ctx->pipe_inflight = 0; ctx->pipe_pending = nullptr; };
    unsigned overflowed = 0;
    if (ctx->tiled && !(s.h_ctr->nan_bounds | s.h_ctr->overflow) && s.h_ctr->comm_error) {   // same reports as mgfb_step_n
        const unsigned ce = s.h_ctr->comm_error;
        drain(); drop_all();
        if (ce & COMM_TILE_TOO_THIN)
            return fail(ctx, MGFB_ERR_TILE, "tile too thin: a body is a ghost on the left neighbour and touches a ghost from the right (or reaches two tiles away); use fewer, wider tiles");
        return fail(ctx, MGFB_ERR_TILE, "neighbour tile did not answer within the time limit");
    }
    if (s.h_ctr->nan_bounds) {
        drain(); drop_all(); clear_sticky(ctx);
        return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (NaN in body state; bounds.rs:125-127)");
    }
    if (s.h_ctr->overflow) {
        // A work list overflowed inside this step: it integrated, then every later kernel -- the rest of this step and ALL of
        // the younger queued steps, their velocity inputs included -- returned early on the sticky flag.  Nothing is lost:
        // regrow, re-run this step from after its integration, re-queue the younger ones whole (mgfb_step_n does the same).
        drain();
        if (ctx->tiled || (s.h_ctr->overflow & OVF_GROUPS)) {
            const bool groups = (s.h_ctr->overflow & OVF_GROUPS) != 0;
            drop_all(); clear_sticky(ctx);
            return fail(ctx, MGFB_ERR_CAPACITY, groups ? "more constraint groups than group capacity"
                                                       : "a work list overflowed in a pipelined tiled step: raise ghost_capacity (the lists hold 16 pairs / 16 contacts per body slot)");
        }
        for (int attempt = 0;; ++attempt) {
            overflowed |= s.h_ctr->overflow;
            if (attempt == 7) { drop_all(); clear_sticky(ctx); return fail(ctx, MGFB_ERR_CAPACITY, "work lists still overflow after growing 128x"); }
            ctx->pipe_scale *= 2;
            TRY(clear_sticky(ctx));
            TRY(pipe_ensure_lists(ctx));
            TRY(pipe_launch(ctx, s, false));
            TRY(pipe_flush_pending(ctx));
            CU(cudaEventSynchronize(s.ev_out));
            if (s.h_ctr->nan_bounds) { drop_all(); clear_sticky(ctx); return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)"); }
            if (!s.h_ctr->overflow) break;
        }
        for (unsigned k = 1; k < ctx->pipe_inflight; ++k) TRY(pipe_launch(ctx, ctx->pipe[(ctx->pipe_head + k) % MGFB_PIPE_DEPTH], true));
    }
    ctx->pipe_head++; ctx->pipe_inflight--;
    const Counters& h = *s.h_ctr;
    *ctx->h_ctr = h;
    ctx->last_constraints = h.contacts;
    ctx->have_step = ctx->pipe_inflight == 0;   // mgfb_step_constraints describes the LAST step the device ran
    if (stats) fill_step_stats(ctx, stats, s.iters, true, overflowed, &h, s.ev);
    return MGFB_OK;
}
}  // extern "C"
