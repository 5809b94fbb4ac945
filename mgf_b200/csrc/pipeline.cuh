// pipeline.cuh -- World::step for callers that stream state to the host every step (rendering, logging,
// a host-side controller): mgfb_step_enqueue queues  [H2D v, omega] -> step -> [pack x, q, v, omega -> D2H]
// and returns at once; mgfb_step_wait blocks until the OLDEST queued step's outputs are in the caller's
// buffers.  Two steps may be in flight, so the PCIe transfers of step k (separate copy streams, both
// directions at once) overlap the kernels of step k+1.  Same kernels, same results as mgfb_step.  On a tiled
// world every rank enqueues and waits the same sequence (the tiles meet on the device, as with mgfb_step_n).
// Included at the end of capi.cu.
#pragma once

struct PipeSlot {
    Buf in, out;                       // device staging: v, omega in; x, q, v, omega out
    Counters* h_ctr = nullptr;         // pinned copy of the step's counters
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // step begin / end, solve begin / end
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    unsigned iters = 0;
};

void pipe_destroy(mgfb_ctx* ctx) {
    if (!ctx->pipe) return;
    for (int k = 0; k < 2; ++k) {
        PipeSlot& s = ctx->pipe[k];
        release(s.in); release(s.out);
        if (s.h_ctr) cudaFreeHost(s.h_ctr);
        for (auto& e : s.ev) if (e) cudaEventDestroy(e);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    delete[] ctx->pipe; ctx->pipe = nullptr;
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    ctx->s_h2d = ctx->s_d2h = nullptr;
}

namespace {
int32_t pipe_init(mgfb_ctx* ctx) {
    if (ctx->pipe) return MGFB_OK;
    ctx->pipe = new PipeSlot[2];
    CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        PipeSlot& s = ctx->pipe[k];
        CU(cudaMallocHost(&s.h_ctr, sizeof(Counters)));
        for (auto& e : s.ev) CU(cudaEventCreate(&e));
        CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
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
    if (ctx->pipe_inflight >= 2) return fail(ctx, MGFB_ERR_STATE, "two steps are in flight: mgfb_step_wait first");
    CU(cudaSetDevice(ctx->device));
    TRY(pipe_init(ctx));
    // like a tile, a pipelined world cannot regrow a work list in the middle of a step it is not waiting for: sized once, generously
    TRY(ensure_step_buffers(ctx, 2));
    TRY(ensure_grid(ctx, 2));
    TRY(ensure_rows(ctx, ctx->contact_cap, false, 4096));
    const unsigned n = ctx->n;
    PipeSlot& s = ctx->pipe[(ctx->pipe_head + ctx->pipe_inflight) & 1u];
    TRY(ensure(ctx, s.in, (size_t)n * 24)); TRY(ensure(ctx, s.out, (size_t)n * 52));
    s.iters = iters;
    if (v_in) {
        float* sv = s.in.as<float>(); float* sw = sv + (size_t)3 * n;
        CU(cudaMemcpyAsync(sv, v_in, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU(cudaMemcpyAsync(sw, omega_in, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU(cudaEventRecord(s.ev_in, ctx->s_h2d));
        CU(cudaStreamWaitEvent(ctx->stream, s.ev_in, 0));
        if (input_mode == MGFB_INPUT_ADD) k_set_velocity<true><<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, sv, sw);
        else k_set_velocity<false><<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, sv, sw);
        ctx->launches += 1;
    }
    ctx->cur_ev = s.ev;
    CU(cudaEventRecord(s.ev[0], ctx->stream));
    int32_t st = enqueue_step(ctx, dt, iters, true, true);
    ctx->cur_ev = ctx->ev;
    TRY(st);
    CU(cudaEventRecord(s.ev[1], ctx->stream));
    CU(cudaMemcpyAsync(s.h_ctr, dctr(ctx), sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    float* sx = s.out.as<float>(); float* sq = sx + (size_t)3 * n; float* sv = sq + (size_t)4 * n; float* sw = sv + (size_t)3 * n;
    if (x_out || q_out || v_out || omega_out) {
        k_pack_state<<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), 0, n, x_out ? sx : nullptr,
                                                                                               q_out ? sq : nullptr, v_out ? sv : nullptr, omega_out ? sw : nullptr);
        ctx->launches += 1;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(s.ev_done, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->s_d2h, s.ev_done, 0));
    if (x_out) CU(cudaMemcpyAsync(x_out, sx, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (q_out) CU(cudaMemcpyAsync(q_out, sq, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (v_out) CU(cudaMemcpyAsync(v_out, sv, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (omega_out) CU(cudaMemcpyAsync(omega_out, sw, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU(cudaEventRecord(s.ev_out, ctx->s_d2h));
    ctx->pipe_inflight++;
    return MGFB_OK;
}

int32_t mgfb_step_wait(mgfb_ctx* ctx, mgfb_step_stats* stats) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (!ctx->pipe || ctx->pipe_inflight == 0) return fail(ctx, MGFB_ERR_STATE, "no step in flight");
    CU(cudaSetDevice(ctx->device));
    PipeSlot& s = ctx->pipe[ctx->pipe_head & 1u];
    CU(cudaEventSynchronize(s.ev_out));
    ctx->pipe_head++; ctx->pipe_inflight--;
    const Counters& h = *s.h_ctr;
    if (ctx->tiled && !(h.nan_bounds | h.overflow) && h.comm_error) {   // same reports as mgfb_step_n
        while (ctx->pipe_inflight) { cudaEventSynchronize(ctx->pipe[ctx->pipe_head & 1u].ev_out); ctx->pipe_head++; ctx->pipe_inflight--; }
        if (h.comm_error & COMM_TILE_TOO_THIN)
            return fail(ctx, MGFB_ERR_TILE, "tile too thin: a body is a ghost on the left neighbour and touches a ghost from the right (or reaches two tiles away); use fewer, wider tiles");
        return fail(ctx, MGFB_ERR_TILE, "neighbour tile did not answer within the time limit");
    }
    if (h.nan_bounds | h.overflow) {
        // every later kernel returned early on the sticky flag: drain what is queued, then report
        while (ctx->pipe_inflight) { cudaEventSynchronize(ctx->pipe[ctx->pipe_head & 1u].ev_out); ctx->pipe_head++; ctx->pipe_inflight--; }
        CU(cudaStreamSynchronize(ctx->stream));
        bool nan = h.nan_bounds != 0;
        clear_sticky(ctx);
        if (nan) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (NaN in body state; bounds.rs:125-127)");
        return fail(ctx, MGFB_ERR_CAPACITY, "a work list overflowed in a pipelined step (lists hold 16 pairs / 16 contacts per body): the step was "
                                            "integrated but not solved; call mgfb_step, which regrows the lists, to continue");
    }
    *ctx->h_ctr = h;
    ctx->last_constraints = h.contacts;
    ctx->have_step = ctx->pipe_inflight == 0;   // mgfb_step_constraints describes the LAST step the device ran
    if (stats) fill_step_stats(ctx, stats, s.iters, true, 0, &h, s.ev);
    return MGFB_OK;
}
}  // extern "C"
