// selftest.cuh -- the one hardware property the dataflow solver leans on, checked instead of assumed.
//
// k_solve_df hands a body's velocity from one constraint row to the next with ONE 32-byte store (st.relaxed.{gpu,sys}.v8.f32:
// v, tag | omega, tag) and the consumer accepts the record when both tags match -- no fence, no separate flag.  The PTX memory
// model does not promise that a vector store becomes visible as a unit; sm_100 issues it as one 32-byte-sector write
// (STG.E.ENL2.256), which is what makes payload and tag inseparable in practice.  This test hammers that: writer warps
// overwrite 32-byte records with all eight words = round number while reader warps on other SMs poll them; a read whose eight
// words differ is a TORN hand-over.  It runs at mgfb_ctx_create on local memory (gpu scope) and at mgfb_tile_connect on each
// neighbour's peer-mapped memory (sys scope, over NVLink).  Any torn read switches the context to MGFB_SCHEDULE_PHASES (grid
// barriers + fences, no such assumption) and mgfb_selftest_handover reports it.
// Included at the end of capi.cu.
#pragma once

namespace mgfb {
template <bool SYS>
__global__ void __launch_bounds__(32) k_inbox_selftest(Inbox* box, unsigned rounds, unsigned* torn, unsigned* observed) {
    const unsigned lane = threadIdx.x, pair = blockIdx.x >> 1;
    Inbox* p = box + pair * 32u + lane;
    if ((blockIdx.x & 1u) == 0u) {            // writer
        for (unsigned r = 1; r <= rounds; ++r) {
            const float f = __uint_as_float(r);
            st_inbox<SYS>(p, mk3(f, f, f), mk3(f, f, f), r);
        }
    } else {                                   // reader, on another SM
        unsigned last = 0, seen = 0, bad = 0;
        for (unsigned spin = 0; spin < (1u << 22) && last != rounds; ++spin) {
            Inbox r = ld_inbox<SYS>(p);
            const unsigned w0 = __float_as_uint(r.lo.x);
            const bool whole = __float_as_uint(r.lo.y) == w0 && __float_as_uint(r.lo.z) == w0 && __float_as_uint(r.lo.w) == w0 &&
                               __float_as_uint(r.hi.x) == w0 && __float_as_uint(r.hi.y) == w0 && __float_as_uint(r.hi.z) == w0 &&
                               __float_as_uint(r.hi.w) == w0;
            if (!whole) ++bad;
            else if (w0 != last) { ++seen; last = w0; }
        }
        if (bad) atomicAdd(torn, bad);
        atomicAdd(observed, seen);
    }
}
}  // namespace mgfb

namespace {
// `box`: SELFTEST_PAIRS * 32 zeroed Inbox records reachable from this context's device (local or peer-mapped).
int32_t run_handover_selftest(mgfb_ctx* ctx, Inbox* box, bool sys, unsigned rounds, unsigned* torn, unsigned* observed) {
    Buf res;
    TRY(ensure(ctx, res, 8, false, true));
    if (sys) k_inbox_selftest<true><<<2 * SELFTEST_PAIRS, 32, 0, ctx->stream>>>(box, rounds, res.as<unsigned>(), res.as<unsigned>() + 1);
    else k_inbox_selftest<false><<<2 * SELFTEST_PAIRS, 32, 0, ctx->stream>>>(box, rounds, res.as<unsigned>(), res.as<unsigned>() + 1);
    CU(cudaGetLastError());
    unsigned h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, res.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    release(res);
    ctx->launches += 1;
    if (torn) *torn = h[0];
    if (observed) *observed = h[1];
    ctx->handover_torn += h[0]; ctx->handover_observed += h[1];
    return MGFB_OK;
}
int32_t local_handover_selftest(mgfb_ctx* ctx, unsigned rounds, unsigned* torn, unsigned* observed) {
    Buf box;
    TRY(ensure(ctx, box, (size_t)SELFTEST_PAIRS * 32 * sizeof(Inbox), false, true));
    int32_t st = run_handover_selftest(ctx, box.as<Inbox>(), false, rounds, torn, observed);
    release(box);
    return st;
}
}  // namespace

extern "C" int32_t mgfb_selftest_handover(mgfb_ctx* ctx, uint32_t rounds, uint32_t* torn, uint32_t* observed) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    unsigned t = 0, o = 0;
    if (rounds) TRY(local_handover_selftest(ctx, rounds, &t, &o));
    if (torn) *torn = ctx->handover_torn;
    if (observed) *observed = ctx->handover_observed;
    return MGFB_OK;
}
