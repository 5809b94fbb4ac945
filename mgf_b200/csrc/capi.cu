// capi.cu -- context, device memory arenas and the extern "C" entry points of include/mgfb.h.
// Host code here only stages buffers and enqueues kernels; every arithmetic step of the hot
// path runs in the kernels of kernels.cuh.  There is no CPU fallback: without a CUDA device
// mgfb_ctx_create fails.
#include <unistd.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mgfb.h"
#include "kernels.cuh"
#include "reftree.cuh"

using namespace mgfb;

#define MGFB_PIPE_DEPTH 4   // steps mgfb_step_enqueue may have in flight

namespace {

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct TerrainData {
    bool present = false;
    Buf verts, faces, boxes, cell_count, cell_start, ent_id, ent_key, max_bits;
    unsigned nverts = 0, nfaces = 0, table = 0, ent_cap = 0;
    float inv_cell = 1.0f;
    float x[3] = {0, 0, 0};
};

}  // namespace

struct mgfb_ctx {
    mgfb_config cfg;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    unsigned n = 0, cap = 0;
    unsigned n_capsules = 0;
    // body SoA
    Buf x, q, vel, force, torque, imb, col, tight, fat, gid;
    // counters
    Buf ctr;
    Counters* h_ctr = nullptr;  // pinned
    // work lists
    Buf pair_list[4]; unsigned pair_cap = 0;
    Buf tpair_list[2]; unsigned tpair_cap = 0;
    Buf c_a, c_b, c_face, c_sub, c_la, c_lb, c_nt; unsigned contact_cap = 0;
    // ordering scratch
    Buf body_best, body_scratch /* mask[n] u64 + last[n] u32 */, group, group_count, group_start, phase_start, perm;
    unsigned group_cap = 0;
    // rows
    Buf r_ab, r_n, r_t0, r_t1, r_ra, r_rb, r_imp, r_xra, r_xrb, r_xtm, r_dep; unsigned row_cap = 0, xrow_cap = 0;
    // dataflow solver (k_solve_df): rows per body (CSR), successor links, per-row inboxes and inertia
    Buf body_deg, body_start, r_inc, r_next, r_in_a, r_in_b, r_ia;
    Buf c_key, c_csr, c_next, c_inbox;   // dataflow colouring (k_colour_df): keys, per-body constraint lists, chain links, mask inboxes
    unsigned df_epoch = 0;   // inbox tags of one solve are df_epoch + 1 .. df_epoch + iters + 1
    // body grid
    Buf cell_count, cell_start, bg_ent, scan_sums, scan_state; unsigned table = 0, ent_cap = 0;
    TerrainData terrain;
    Buf stage;          // packed state staging for get_state / set_velocity
    Buf convex_pool; unsigned convex_n = 0;   // vertices of the ConvexMeshes the GJK batch may name (mgfb_convex_vertices_set)
    unsigned long long launches = 0;   // kernels launched since the last mgfb_step_totals(reset)
    // user-path staging
    Buf u_a, u_b, u_sc, u_sf, u_n, u_t, u_nc, u_la, u_lb;
    // cooperative grid sizes
    int coop_order = 0, coop_solve = 0, coop_df = 0, coop_colour = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // coherent broadphase (bpcache.cuh): on unless MGFB_BROADPHASE=sweep
    bool bp_on = true, bp_forced = false, bp_invalidate = true;
    Buf bp_state, bp_s[2], bp_stale, bp_c0, bp_ovf, bp_ref_flag, bp_ref_list; unsigned bp_s_cap = 0, bp_slots = 0;
    Buf scan_step;                        // the two look-back scan states of one step (zeroed with the step's scratch)
    Counters* ctr_snap = nullptr;         // pipelined step being enqueued: where k_step_done leaves a copy of its counters
    cudaEvent_t* cur_ev = nullptr;        // the four timing events of the step being enqueued (ev, or a pipeline slot's)
    struct PipeSlot* pipe = nullptr;     // mgfb_step_enqueue / mgfb_step_wait (pipeline.cuh)
    struct PipeSlot* pipe_pending = nullptr;   // queued step whose host transfers have not been issued yet (pipeline.cuh)
    struct RefOrderState* reforder = nullptr;   // World::step in the reference's own constraint order (reforder.cuh)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t gjk_streams[8] = {};     // GJK/EPA batch: the per-shape-pair launches run side by side (an EPA run is one warp for up to ~0.1 s)
    cudaEvent_t gjk_ev[9] = {};
    cudaStream_t s_aux = nullptr;         // the terrain half of the broad/narrowphase runs beside the body half
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // coherent broadphase: its mutually exclusive paths run side by side (the one not chosen returns at once)
    cudaStream_t s_bp[2] = {nullptr, nullptr}; cudaEvent_t ev_bp[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned pipe_head = 0, pipe_inflight = 0, pipe_scale = 2;
    // last step
    unsigned last_constraints = 0;
    bool have_step = false;
    // spatial tiling (tile.cuh): this ctx owns one slab of a world spread over several GPUs
    bool tile_exported = false, tiled = false;
    unsigned ghost_cap = 0;
    Buf edge_idx, edge_mark, ridx, mbox;
    Buf edge_slot, tile_df;              // dataflow solve across tiles: [in_a | in_b | link_l | link_r] in ONE exported allocation
    unsigned tile_row_cap = 0;           // row capacity frozen at export (the neighbours index my inboxes)
    TileLink link{};
    unsigned long long tile_step = 0;
    std::vector<void*> ipc_opened;
    int max_ctas = 0;                    // cfg.reserved[0]: cap on cooperative grids (two tiles sharing one GPU in tests)
    unsigned long long tile_timeout_ns = 20000000000ULL;
    // mgfb_step_profile: events between the phases of one step (prof_ev[k] = start of phase k; [8] = end; [9],[10] = terrain side stream)
    // selftest.cuh: torn / whole 32-byte hand-overs seen so far (create-time local test + connect-time peer tests + on request)
    unsigned handover_torn = 0, handover_observed = 0;
    bool prof_on = false;
    cudaEvent_t prof_ev[11] = {};
};

// host state of World::step in reference order (reforder.cuh): the two trees of the demo world, replayed
struct RefOrderState {
    RefTree body_tree, mesh_tree;
    std::vector<int> leaf;            // leaf of every body in body_tree
    std::vector<Box> fat;             // the fat box each leaf was inserted with
    std::vector<Box> h_tight, h_fat;  // this step's boxes as the device computed them
    std::vector<int2> cand;
    bool mesh_built = false;
    Buf d_cand, d_la, d_lb, d_nt, d_cnt, d_offs;
};

namespace {

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                    \
            return MGFB_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)
#define TRY(call) do { int32_t _s = (call); if (_s != MGFB_OK) return _s; } while (0)

#define PROF(id) do { if (ctx->prof_on) CU(cudaEventRecord(ctx->prof_ev[id], ctx->stream)); } while (0)
int32_t fail(mgfb_ctx* ctx, int32_t code, const char* msg) { if (ctx) ctx->err = msg; return code; }

// (re)allocate to at least `bytes`, optionally keeping the old contents
int32_t ensure(mgfb_ctx* ctx, Buf& b, size_t bytes, bool keep = false, bool zero = false) {
    if (b.bytes >= bytes && b.p) return MGFB_OK;
    void* np = nullptr;
    CU(cudaMalloc(&np, bytes));
    if (zero) CU(cudaMemsetAsync(np, 0, bytes, ctx->stream));
    if (keep && b.p && b.bytes) CU(cudaMemcpyAsync(np, b.p, b.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (b.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(b.p)); }
    b.p = np; b.bytes = bytes;
    return MGFB_OK;
}
void release(Buf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }

unsigned next_pow2(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; }
int grid_for(const mgfb_ctx* ctx, size_t items) {
    size_t blocks = (items + MGFB_THREADS - 1) / MGFB_THREADS;
    size_t cap = (size_t)ctx->num_sms * 8;
    if (ctx->max_ctas) cap = std::min<size_t>(cap, (size_t)ctx->max_ctas * 4);
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

BodyArrays body_arrays(const mgfb_ctx* ctx) {
    BodyArrays B;
    B.x = ctx->x.as<float4>(); B.q = ctx->q.as<float4>(); B.vel = ctx->vel.as<BodyVel>();
    B.force = ctx->force.as<float4>(); B.torque = ctx->torque.as<float4>(); B.imb = ctx->imb.as<float4>();
    B.col = ctx->col.as<Collider>(); B.tight = ctx->tight.as<Box>(); B.fat = ctx->fat.as<Box>();
    B.gid = ctx->gid.as<unsigned>();
    B.ref_flag = nullptr; B.ref_list = nullptr; B.ref_cap = 0;
    if (ctx->bp_on && ctx->bp_ref_flag.p) { B.ref_flag = ctx->bp_ref_flag.as<unsigned char>(); B.ref_list = ctx->bp_ref_list.as<unsigned>(); B.ref_cap = BP_REF_CAP; }
    return B;
}
BpView bp_view(const mgfb_ctx* ctx) {
    BpView V;
    V.st = ctx->bp_state.as<BpState>();
    V.S[0] = ctx->bp_s[0].as<int2>(); V.S[1] = ctx->bp_s[1].as<int2>(); V.s_cap = ctx->bp_s_cap;
    V.stale = ctx->bp_stale.as<unsigned char>(); V.c0 = ctx->bp_c0.as<float4>(); V.ovf = ctx->bp_ovf.as<unsigned>(); V.ovf_cap = BP_OVF_CAP;
    V.ref_flag = ctx->bp_ref_flag.as<unsigned char>(); V.ref_list = ctx->bp_ref_list.as<unsigned>(); V.ref_cap = BP_REF_CAP;
    return V;
}
// bodies this ctx may hold in a step: its own plus the ghosts a tiled world sends it
unsigned body_slots(const mgfb_ctx* ctx) { return ctx->n + ctx->ghost_cap; }
int32_t grow_bodies(mgfb_ctx* ctx, unsigned need) {
    if (need <= ctx->cap) return MGFB_OK;
    if (ctx->tile_exported) return fail(ctx, MGFB_ERR_STATE, "body arrays are exported to neighbour tiles and cannot grow");
    unsigned nc = std::max(need, std::max(1024u, ctx->cap * 2));
    TRY(ensure(ctx, ctx->x, (size_t)nc * sizeof(float4), true));
    TRY(ensure(ctx, ctx->q, (size_t)nc * sizeof(float4), true));
    TRY(ensure(ctx, ctx->vel, (size_t)nc * sizeof(BodyVel), true));
    TRY(ensure(ctx, ctx->force, (size_t)nc * sizeof(float4), true));
    TRY(ensure(ctx, ctx->torque, (size_t)nc * sizeof(float4), true));
    TRY(ensure(ctx, ctx->imb, (size_t)nc * 3 * sizeof(float4), true));
    TRY(ensure(ctx, ctx->col, (size_t)nc * sizeof(Collider), true));
    TRY(ensure(ctx, ctx->tight, (size_t)nc * sizeof(Box), true));
    TRY(ensure(ctx, ctx->fat, (size_t)nc * sizeof(Box), true));
    TRY(ensure(ctx, ctx->gid, (size_t)nc * 4, true));
    TRY(ensure(ctx, ctx->body_best, (size_t)nc * 8, false, true));
    TRY(ensure(ctx, ctx->body_scratch, (size_t)nc * 12, false, true));
    TRY(ensure(ctx, ctx->body_deg, (size_t)nc * 4)); TRY(ensure(ctx, ctx->body_start, ((size_t)nc + 1) * 4));
    TRY(ensure(ctx, ctx->scan_sums, ((size_t)nc / SCAN_ITEMS + 2) * 4));
    ctx->cap = nc;
    return MGFB_OK;
}
// Work-list capacities scale with the body count; a step that overflows one regrows and reruns.
int32_t ensure_step_buffers(mgfb_ctx* ctx, unsigned scale) {
    unsigned n = std::max(body_slots(ctx), 1024u);
    unsigned pc = n * 8 * scale, tc = n * 8 * scale, cc = n * 8 * scale;
    if (pc > ctx->pair_cap) { for (auto& b : ctx->pair_list) TRY(ensure(ctx, b, (size_t)pc * sizeof(int2))); ctx->pair_cap = pc; }
    if (tc > ctx->tpair_cap) { for (auto& b : ctx->tpair_list) TRY(ensure(ctx, b, (size_t)tc * sizeof(int2))); ctx->tpair_cap = tc; }
    if (ctx->bp_on) {   // the cached superset S (two buffers), stale marks, overflow list, this step's replaced bodies
        unsigned sc = n * 24 * scale, slots = body_slots(ctx);   // fat-box overlaps per body: ~11 in a settled pile of equal spheres, more for mixed sizes
        if (sc > ctx->bp_s_cap) { for (auto& b : ctx->bp_s) TRY(ensure(ctx, b, (size_t)sc * sizeof(int2))); ctx->bp_s_cap = sc; ctx->bp_invalidate = true; }
        if (slots > ctx->bp_slots || !ctx->bp_state.p) {
            TRY(ensure(ctx, ctx->bp_stale, (size_t)slots + 4, false, true)); TRY(ensure(ctx, ctx->bp_c0, (size_t)slots * 16)); TRY(ensure(ctx, ctx->bp_ref_flag, (size_t)slots + 4, false, true));
            TRY(ensure(ctx, ctx->bp_ovf, (size_t)BP_OVF_CAP * 4)); TRY(ensure(ctx, ctx->bp_ref_list, (size_t)BP_REF_CAP * 4));
            TRY(ensure(ctx, ctx->bp_state, sizeof(BpState), false, true));
            ctx->bp_slots = slots; ctx->bp_invalidate = true;
        }
    }
    if (cc > ctx->contact_cap) {
        TRY(ensure(ctx, ctx->c_a, (size_t)cc * 4)); TRY(ensure(ctx, ctx->c_b, (size_t)cc * 4));
        TRY(ensure(ctx, ctx->c_face, (size_t)cc * 4)); TRY(ensure(ctx, ctx->c_sub, (size_t)cc * 4));
        TRY(ensure(ctx, ctx->c_la, (size_t)cc * 16)); TRY(ensure(ctx, ctx->c_lb, (size_t)cc * 16)); TRY(ensure(ctx, ctx->c_nt, (size_t)cc * 16));
        ctx->contact_cap = cc;
    }
    return MGFB_OK;
}
int32_t ensure_rows(mgfb_ctx* ctx, unsigned m, bool extras, unsigned groups) {
    if (m > ctx->row_cap) {
        unsigned rc = std::max(m, 1024u);
        TRY(ensure(ctx, ctx->r_ab, (size_t)rc * 8)); TRY(ensure(ctx, ctx->r_n, (size_t)rc * 16)); TRY(ensure(ctx, ctx->r_t0, (size_t)rc * 16));
        TRY(ensure(ctx, ctx->r_t1, (size_t)rc * 16)); TRY(ensure(ctx, ctx->r_ra, (size_t)rc * 16)); TRY(ensure(ctx, ctx->r_rb, (size_t)rc * 16));
        TRY(ensure(ctx, ctx->r_imp, (size_t)rc * 4)); TRY(ensure(ctx, ctx->group, (size_t)rc * 4)); TRY(ensure(ctx, ctx->perm, (size_t)rc * 4));
        TRY(ensure(ctx, ctx->r_dep, (size_t)rc * 4)); TRY(ensure(ctx, ctx->r_inc, (size_t)rc * 8)); TRY(ensure(ctx, ctx->r_next, (size_t)rc * 8));
        TRY(ensure(ctx, ctx->r_ia, (size_t)rc * 80));
        TRY(ensure(ctx, ctx->c_key, (size_t)rc * 8)); TRY(ensure(ctx, ctx->c_csr, (size_t)rc * 8)); TRY(ensure(ctx, ctx->c_next, (size_t)rc * 8));
        TRY(ensure(ctx, ctx->c_inbox, (size_t)rc * 16));
        // inbox tags must never match by accident: zeroed when (re)allocated, epochs only grow
        TRY(ensure(ctx, ctx->r_in_a, (size_t)rc * sizeof(Inbox), false, true)); TRY(ensure(ctx, ctx->r_in_b, (size_t)rc * sizeof(Inbox), false, true));
        ctx->row_cap = rc;
    }
    if (extras && m > ctx->xrow_cap) {
        unsigned rc = std::max(m, 1024u);
        TRY(ensure(ctx, ctx->r_xra, (size_t)rc * 48)); TRY(ensure(ctx, ctx->r_xrb, (size_t)rc * 48)); TRY(ensure(ctx, ctx->r_xtm, (size_t)rc * 48));
        ctx->xrow_cap = rc;
    }
    if (groups > ctx->group_cap) {
        unsigned gc = std::max(groups, 4096u);
        TRY(ensure(ctx, ctx->group_count, (size_t)gc * 4)); TRY(ensure(ctx, ctx->group_start, ((size_t)gc + 1) * 4));
        TRY(ensure(ctx, ctx->phase_start, ((size_t)gc + 1) * 4));
        ctx->group_cap = gc;
    }
    return MGFB_OK;
}
int32_t ensure_grid(mgfb_ctx* ctx, unsigned scale) {
    unsigned table = next_pow2(std::max(4096u, body_slots(ctx) * 4));
    unsigned ecap = std::max(body_slots(ctx), 1024u) * 12 * scale;
    if (table > ctx->table) {
        TRY(ensure(ctx, ctx->cell_count, (size_t)table * 4)); TRY(ensure(ctx, ctx->cell_start, ((size_t)table + 1) * 4));
        TRY(ensure(ctx, ctx->scan_sums, ((size_t)table / SCAN_ITEMS + 2) * 4));
        ctx->table = table;
    }
    (void)ecap;
    if (ctx->cap > ctx->ent_cap) { TRY(ensure(ctx, ctx->bg_ent, (size_t)ctx->cap * 32)); ctx->ent_cap = ctx->cap; }
    {   // look-back states of the step's two scans (capi.cu step_scan_words): sized here, outside the step
        size_t w = (((size_t)ctx->table + SCAN_ITEMS - 1) / SCAN_ITEMS + 2) * 2 + (((size_t)body_slots(ctx) + SCAN_ITEMS - 1) / SCAN_ITEMS + 2) * 2;
        TRY(ensure(ctx, ctx->scan_step, w * 4));
    }
    return MGFB_OK;
}
// exclusive scan of `in[0..n)` into out[0..n], out[n] = total (also stored at *total_dev if given)
// single-pass scan (decoupled look-back) for the step path: one launch + one small memset
// `step_state`: a look-back state already zeroed by the step's k_zero_ranges (no memset in the step's stream)
int32_t scan_u32_lb(mgfb_ctx* ctx, const unsigned* in, unsigned* out, unsigned n, unsigned* total_dev, unsigned long long* step_state = nullptr) {
    unsigned nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    if (!step_state) {
        TRY(ensure(ctx, ctx->scan_state, ((size_t)nb + 2) * 8));
        CU(cudaMemsetAsync(ctx->scan_state.p, 0, ((size_t)nb + 1) * 8, ctx->stream));
        step_state = ctx->scan_state.as<unsigned long long>();
    }
    k_scan_lookback<<<nb, 256, 0, ctx->stream>>>(in, n, out, step_state, total_dev);
    CU(cudaGetLastError());
    return MGFB_OK;
}
// the two look-back states of one step (cell scan, body-degree scan), zeroed together with the rest of the step's scratch
size_t step_scan_words(const mgfb_ctx* ctx, unsigned which) {
    size_t n = which == 0 ? ctx->table : body_slots(ctx);
    return ((n + SCAN_ITEMS - 1) / SCAN_ITEMS + 2) * 2;   // 8-byte entries, in 4-byte words
}
unsigned long long* step_scan_state(const mgfb_ctx* ctx, unsigned which) {
    return ctx->scan_step.as<unsigned long long>() + (which == 0 ? 0 : step_scan_words(ctx, 0) / 2);
}
int32_t zero_ranges(mgfb_ctx* ctx, const ZeroRanges& Z) {
    size_t total = 0;
    for (unsigned r = 0; r < Z.n; ++r) total += Z.words[r];
    k_zero_ranges<<<grid_for(ctx, total / 4 + 1), MGFB_THREADS, 0, ctx->stream>>>(Z);
    CU(cudaGetLastError());
    ctx->launches += 1;
    return MGFB_OK;
}
int32_t scan_u32(mgfb_ctx* ctx, const unsigned* in, unsigned* out, unsigned n, unsigned* sums, unsigned* total_dev) {
    unsigned nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    k_scan_reduce<<<nb, 256, 0, ctx->stream>>>(in, n, sums);
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(sums, nb, total_dev);
    k_scan_final<<<nb, 256, 0, ctx->stream>>>(in, n, sums, out, total_dev);
    CU(cudaGetLastError());
    return MGFB_OK;
}

BodyGrid body_grid(const mgfb_ctx* ctx) {
    BodyGrid G;
    G.cell_count = ctx->cell_count.as<unsigned>(); G.cell_start = ctx->cell_start.as<unsigned>();
    G.ent = ctx->bg_ent.as<float4>(); G.table_mask = ctx->table - 1;
    return G;
}
TerrainView terrain_view(const mgfb_ctx* ctx) {
    TerrainView T;
    const TerrainData& t = ctx->terrain;
    T.verts = t.verts.as<float4>(); T.faces = t.faces.as<uint4>(); T.boxes = t.boxes.as<Box>();
    T.G.cell_count = t.cell_count.as<unsigned>(); T.G.cell_start = t.cell_start.as<unsigned>();
    T.G.ent_id = t.ent_id.as<unsigned>(); T.G.ent_key = t.ent_key.as<unsigned long long>();
    T.G.table_mask = t.table - 1; T.G.ent_cap = t.ent_cap;
    T.inv_cell = t.inv_cell; T.nfaces = t.nfaces;
    T.x = make_float4(t.x[0], t.x[1], t.x[2], 0.0f);
    return T;
}
ContactList contact_list(const mgfb_ctx* ctx) {
    ContactList L;
    L.a = ctx->c_a.as<int>(); L.b = ctx->c_b.as<int>(); L.face = ctx->c_face.as<uint32_t>(); L.sub = ctx->c_sub.as<uint32_t>();
    L.la = ctx->c_la.as<float4>(); L.lb = ctx->c_lb.as<float4>(); L.nt = ctx->c_nt.as<float4>();
    return L;
}
ConstraintRows rows_view(const mgfb_ctx* ctx) {
    ConstraintRows R;
    R.ab = ctx->r_ab.as<int2>(); R.n = ctx->r_n.as<float4>(); R.t0 = ctx->r_t0.as<float4>(); R.t1 = ctx->r_t1.as<float4>();
    R.ra = ctx->r_ra.as<float4>(); R.rb = ctx->r_rb.as<float4>(); R.impulse = ctx->r_imp.as<float>();
    R.xra = ctx->r_xra.as<float4>(); R.xrb = ctx->r_xrb.as<float4>(); R.xtm = ctx->r_xtm.as<float4>();
    return R;
}
BodyInfoView body_info(const mgfb_ctx* ctx) {
    BodyInfoView B;
    B.x = ctx->x.as<float4>(); B.vel = ctx->vel.as<BodyVel>(); B.col = ctx->col.as<Collider>();
    B.force = ctx->force.as<float4>(); B.torque = ctx->torque.as<float4>();
    return B;
}
OrderView order_view(const mgfb_ctx* ctx, const int* a, const int* b, const uint32_t* face, const uint32_t* sub) {
    OrderView O;
    O.a = a; O.b = b; O.face = face; O.sub = sub;
    O.body_best = ctx->body_best.as<unsigned long long>();
    O.body_mask = ctx->body_scratch.as<unsigned long long>();
    O.body_last = reinterpret_cast<unsigned*>(ctx->body_scratch.as<unsigned long long>() + ctx->cap);
    O.group = ctx->group.as<int>(); O.group_count = ctx->group_count.as<unsigned>(); O.gcap = ctx->group_cap;
    O.gid = ctx->gid.as<unsigned>(); O.n_own = ctx->tiled ? ctx->n : 0xffffffffu;
    O.x = face ? ctx->x.as<float4>() : nullptr; O.col = face ? ctx->col.as<Collider>() : nullptr;   // step path only: caller-built manifolds carry no geometry
    return O;
}
Counters* dctr(const mgfb_ctx* ctx) { return ctx->ctr.as<Counters>(); }

template <class K>
int coop_blocks(const mgfb_ctx* ctx, K kernel, int threads, int max_per_sm) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    per_sm = std::max(1, std::min(per_sm, max_per_sm));
    int blocks = per_sm * ctx->num_sms;
    return ctx->max_ctas ? std::min(blocks, ctx->max_ctas) : blocks;
}

// order -> scan -> scatter -> build -> solve, for `m` constraints counted on the device
// (m_ptr) or known on the host (m_host).
int32_t enqueue_order_and_solve(mgfb_ctx* ctx, const OrderView& O, const ManifoldInput& M, const unsigned* m_ptr, unsigned m_host,
                                unsigned m_bound, bool as_given, float dt, unsigned iters, bool time_solve, bool step_scratch_zeroed = false,
                                bool count_fused = false) {
    Counters* c = dctr(ctx);
    ConstraintRows R = rows_view(ctx);
    BodyInfoView BI = body_info(ctx);
    unsigned* gcount = ctx->group_count.as<unsigned>();
    unsigned* gstart = ctx->group_start.as<unsigned>();
    unsigned* pstart = ctx->phase_start.as<unsigned>();
    unsigned* perm = ctx->perm.as<unsigned>();
    const bool tiled = ctx->tiled && !M.user;
    BodyVel* vel = ctx->vel.as<BodyVel>();
    unsigned gcap = ctx->group_cap;
    const unsigned nb = M.user ? ctx->n : body_slots(ctx);
    int g = grid_for(ctx, m_bound);
    const bool colour_df = !as_given && ctx->cfg.solver_schedule != MGFB_SCHEDULE_PHASES_JP;
    PROF(MGFB_PHASE_COLOURING);
    if (colour_df) {
        // constraints per body -> CSR -> chains sorted by key -> colours travel down the chains (no grid barrier)
        ColourView V{};
        V.key = ctx->c_key.as<unsigned long long>(); V.deg = ctx->body_deg.as<unsigned>(); V.body_start = ctx->body_start.as<unsigned>();
        V.csr = ctx->c_csr.as<unsigned>(); V.next = ctx->c_next.as<unsigned>(); V.inbox = ctx->c_inbox.as<unsigned long long>(); V.cap = ctx->row_cap;
        if (!step_scratch_zeroed) CU(cudaMemsetAsync(V.deg, 0, (size_t)nb * 4, ctx->stream));
        if (!count_fused) k_inc_count<<<g, MGFB_THREADS, 0, ctx->stream>>>(O, V, m_ptr, m_host, c);   // (the step's narrowphase did it while emitting: EmitHook)
        TRY(scan_u32_lb(ctx, V.deg, ctx->body_start.as<unsigned>(), nb, &c->df_links, step_scratch_zeroed ? step_scan_state(ctx, 1) : nullptr));
        k_inc_fill<<<g, MGFB_THREADS, 0, ctx->stream>>>(O, V, m_ptr, m_host, c);
        k_inc_sort<<<grid_for(ctx, nb), MGFB_THREADS, 0, ctx->stream>>>(O, V, nb, c);
        CU(cudaGetLastError());
        OrderView Ov = O; unsigned mh = m_host; const unsigned* mp = m_ptr;
        void* args[] = {&Ov, &V, &mp, &mh, &c};
        CU(cudaLaunchCooperativeKernel((void*)k_colour_df, dim3(ctx->coop_colour), dim3(MGFB_THREADS), args, 0, ctx->stream));
        ctx->launches += count_fused ? 4 : 5;
    }
    {
        OrderView Ov = O; bool ag = as_given; unsigned mh = m_host; const unsigned* mp = m_ptr;
        unsigned fallback_only = colour_df ? 1u : 0u; unsigned nbo = nb;
        void* args[] = {&Ov, &mp, &mh, &ag, &c, &fallback_only, &nbo};
        CU(cudaLaunchCooperativeKernel((void*)k_order, dim3(ctx->coop_order), dim3(MGFB_THREADS), args, 0, ctx->stream));
    }
    k_group_scan<<<1, 1024, 0, ctx->stream>>>(gcount, gstart, pstart, c, gcap, tiled ? (unsigned)TILE_INTERIOR_COLOURS : 0xffffffffu);
    k_scatter_rows<<<g, MGFB_THREADS, 0, ctx->stream>>>(O.group, gcount, gstart, perm, m_ptr, m_host, c);
    // Schedule of the solve: dataflow (body version counters, no grid barrier) for coloured single-GPU
    // solves; grid-barrier phases for as-given (level) order, tiled worlds, > 64 colours, or on request.
    // (a tiled world derives its inbox tags from the common step number: 2048 tags per step)
    const bool dataflow = colour_df && ctx->cfg.solver_schedule == MGFB_SCHEDULE_DATAFLOW && (!tiled || (m_bound <= ctx->tile_row_cap && iters < 2000u));
    DfArrays D{};
    if (dataflow) {   // rows per body = number of colours at the body; CSR offsets (already made by the colouring) for the successor links
        D.in_a = ctx->r_in_a.as<Inbox>(); D.in_b = ctx->r_in_b.as<Inbox>(); D.ia = ctx->r_ia.as<float4>(); D.row_cap = ctx->row_cap;
        if (tiled) { D.in_a = ctx->tile_df.as<Inbox>(); D.in_b = D.in_a + ctx->tile_row_cap; }   // the inboxes the neighbours push into
        D.next = ctx->r_next.as<unsigned>(); D.dep = ctx->r_dep.as<unsigned>(); D.body_start = ctx->body_start.as<unsigned>();
        D.inc = ctx->r_inc.as<unsigned>();
    }
    PROF(MGFB_PHASE_BUILD_ROWS);
    k_build_rows<<<g, MGFB_THREADS, 0, ctx->stream>>>(M, BI, perm, R, m_ptr, m_host, dt, ctx->cfg.baumgarte, ctx->cfg.penetration_slop, c,
                                                      tiled ? ctx->edge_mark.as<unsigned char>() : nullptr, tiled ? ctx->n : 0xffffffffu,
                                                      O.group, O.body_mask, D);
    unsigned epoch = ctx->df_epoch;
    TileLink TL = ctx->link;
    if (dataflow && !tiled) {
        if (ctx->df_epoch > 0x7fffffffu - 2u * (iters + 2u)) {   // tag space exhausted (once per ~10^8 solves): start over from clean inboxes
            CU(cudaMemsetAsync(ctx->r_in_a.p, 0, ctx->r_in_a.bytes, ctx->stream)); CU(cudaMemsetAsync(ctx->r_in_b.p, 0, ctx->r_in_b.bytes, ctx->stream));
            ctx->df_epoch = epoch = 0;
        }
        ctx->df_epoch += iters + 2u;
        k_df_init<false><<<g, MGFB_THREADS, 0, ctx->stream>>>(vel, nb, R.ab, D, epoch, m_ptr, m_host, c, TL);
        ctx->launches += 1;
    }
    if (dataflow && tiled) {
        // every tile derives the same tags from the common step number (upper half of the tag space; the inboxes
        // were cleared at step start whenever the 2^20-step cycle restarts)
        epoch = 0x80000000u | (unsigned)((ctx->tile_step & 0xfffffULL) << 11);
        k_tile_links_send<<<1, 1024, 0, ctx->stream>>>(TL, D, R.ab, c);
        k_df_init<true><<<g, MGFB_THREADS, 0, ctx->stream>>>(vel, nb, R.ab, D, epoch, m_ptr, m_host, c, TL);   // waits for both neighbours' link tables itself
        ctx->launches += 2;
    }
    CU(cudaGetLastError());
    if (time_solve) CU(cudaEventRecord(ctx->cur_ev[2], ctx->stream));
    PROF(MGFB_PHASE_SOLVE);
    if (dataflow) {
        const unsigned* ps = pstart; unsigned it = iters;
        void* args[] = {&R, &D, &vel, &ps, &it, &epoch, &c, &TL};
        void* fn = tiled ? (void*)k_solve_df<true> : (void*)k_solve_df<false>;
        // rows of the previous solve decide the block size (the count of THIS step is still on the device)
        const unsigned threads = ctx->last_constraints > MGFB_DF_LARGE_ROWS ? MGFB_DF_THREADS_LARGE : MGFB_DF_THREADS_SMALL;
        CU(cudaLaunchCooperativeKernel(fn, dim3(ctx->coop_df), dim3(threads), args, 0, ctx->stream));
        ctx->launches += 1;
        if (tiled) {   // my ghosts' final velocities are in their owner's records; wait for the same from the left tile
            k_tile_solve_done<<<1, 1, 0, ctx->stream>>>(TL, c);
            if (TL.has_left) k_tile_wait<<<1, 1, 0, ctx->stream>>>(&TL.mine->done_from_left.flag, TL.step, TL.timeout_ns, c);
            ctx->launches += 1 + (TL.has_left ? 1 : 0);
        }
    }
    if (!(dataflow && tiled))
    {
        const unsigned* ps = pstart; unsigned it = iters;
        TileLink T = ctx->link;
        unsigned only_beyond = dataflow ? (unsigned)MGFB_DF_MAX_PHASES : 0u;   // after k_solve_df: only if it declined (> 64 colours)
        void* args[] = {&R, &vel, &ps, &it, &c, &T, &only_beyond};
        void* fn = tiled ? (void*)k_solve<true> : (void*)k_solve<false>;
        CU(cudaLaunchCooperativeKernel(fn, dim3(ctx->coop_solve), dim3(MGFB_SOLVE_THREADS), args, 0, ctx->stream));
    }
    if (time_solve) CU(cudaEventRecord(ctx->cur_ev[3], ctx->stream));
    PROF(MGFB_PHASE_COUNT);
    ctx->launches += 4 + ((dataflow && tiled) ? 0 : 1);   // k_order, k_group_scan, k_scatter_rows, k_build_rows (+ k_solve)
    return MGFB_OK;
}

// One World::step.  `from_integrate` = false re-runs only the part after integration (used when
// a work list overflowed and was regrown).
int32_t enqueue_step(mgfb_ctx* ctx, float dt, unsigned iters, bool from_integrate, bool timed) {
    Counters* c = dctr(ctx);
    unsigned n = ctx->n;
    BodyArrays B = body_arrays(ctx);
    const bool tiled = ctx->tiled;
    const unsigned slots = body_slots(ctx);   // upper bound of n_total (device-resident: own + this step's ghosts)
    {   // the step's scratch, zeroed by ONE kernel (not memsets: see k_zero_ranges)
        static_assert(offsetof(Counters, overflow) % 4 == 0, "counters are 32-bit words");
        TRY(ensure(ctx, ctx->scan_step, (step_scan_words(ctx, 0) + step_scan_words(ctx, 1)) * 4));
        ZeroRanges Z{};
        auto add = [&](void* p, size_t bytes) { Z.p[Z.n] = reinterpret_cast<unsigned*>(p); Z.words[Z.n] = (unsigned)((bytes + 3) / 4); Z.n++; };
        add(c, offsetof(Counters, overflow));
        add(ctx->body_scratch.p, (size_t)ctx->cap * 12);
        add(ctx->group_count.p, (size_t)ctx->group_cap * 4);
        add(ctx->cell_count.p, (size_t)ctx->table * 4);
        add(ctx->scan_step.p, (step_scan_words(ctx, 0) + step_scan_words(ctx, 1)) * 4);
        add(ctx->body_deg.p, (size_t)slots * 4);
        if (tiled) add(ctx->edge_mark.p, ctx->n);
        if (ctx->bp_on) {
            add(ctx->bp_ref_flag.p, slots);
            if (ctx->bp_invalidate) { add(&ctx->bp_state.as<BpState>()->valid, 4); ctx->bp_invalidate = false; }   // rebuild from scratch this step
        }
        TRY(zero_ranges(ctx, Z));
    }
    int gb = grid_for(ctx, n);
    PROF(MGFB_PHASE_INTEGRATE);
    if (!tiled) {
        if (from_integrate) k_integrate<true, true, true><<<gb, MGFB_THREADS, 0, ctx->stream>>>(B, n, dt, ctx->cfg.fat_margin, c);
        else k_integrate<false, false, true><<<gb, MGFB_THREADS, 0, ctx->stream>>>(B, n, dt, ctx->cfg.fat_margin, c);
    } else {
        // tile.cuh: my right extent -> right neighbour; my edge bodies -> left neighbour's ghost slots
        ctx->tile_step++;
        ctx->link.step = ctx->tile_step;
        const TileLink& T = ctx->link;
        if ((ctx->tile_step & 0xfffffULL) == 0) CU(cudaMemsetAsync(ctx->tile_df.p, 0, (size_t)ctx->tile_row_cap * 2 * sizeof(Inbox), ctx->stream));
        k_integrate<true, true, true, true><<<gb, MGFB_THREADS, 0, ctx->stream>>>(B, n, dt, ctx->cfg.fat_margin, c);
        k_tile_publish<<<1, 1, 0, ctx->stream>>>(T, c);
        if (T.has_left) k_ghost_send<<<gb, 256, 0, ctx->stream>>>(B, T, c);   // waits for the left extent itself
        k_ghost_recv<<<grid_for(ctx, std::max(ctx->ghost_cap, 1u)), 256, 0, ctx->stream>>>(B, T, c);   // waits for the right neighbour's ghosts itself
        ctx->launches += 2 + (T.has_left ? 1 : 0);
    }
    int gs = grid_for(ctx, slots);
    if (ctx->terrain.present) CU(cudaEventRecord(ctx->ev_fork, ctx->stream));   // tight boxes are final: the terrain half may start
    // broadphase over the stored fat boxes
    PROF(MGFB_PHASE_BODY_GRID);
    BodyGrid G = body_grid(ctx);
    PairLists PL; for (int k = 0; k < 4; ++k) PL.p[k] = ctx->pair_list[k].as<int2>();
    if (ctx->bp_on) {
        // coherent broadphase (bpcache.cuh): the device picks this step's path; the kernels of the other one return at once
        BpView V = bp_view(ctx);
        const int gn = grid_for(ctx, slots);
        k_bp_decide<<<1, 1, 0, ctx->stream>>>(V, c, n, (tiled && ctx->tile_step % BP_TILED_PERIOD == 0) ? 1u : 0u);
        // The paths exclude each other at run time (the kernels of the one not chosen return at once), so they are queued
        // side by side: [A] rebuild: grid + one sweep, on s_bp[0]; [B] coherent: mark + filter, on the step's stream;
        // [C] the queries, on s_bp[1], after the marks and -- for a rebuild's ghosts -- after the grid.
        cudaStream_t sA = ctx->s_bp[0], sC = ctx->s_bp[1];
        CU(cudaEventRecord(ctx->ev_bp[0], ctx->stream));
        CU(cudaStreamWaitEvent(sA, ctx->ev_bp[0], 0));
        // -- [A] rebuild: grid over the own bodies' fat boxes; one sweep writes S and this step's pair lists
        k_bp_grid<false><<<gn, MGFB_THREADS, 0, sA>>>(B.fat, B.col, B.gid, n, G, V, c);
        {
            unsigned nbk = (ctx->table + SCAN_ITEMS - 1) / SCAN_ITEMS;
            k_scan_lookback<<<nbk, 256, 0, sA>>>(G.cell_count, ctx->table, G.cell_start, step_scan_state(ctx, 0), &c->grid_entries, &V.st->mode, (unsigned)BP_COHERENT);
        }
        k_bp_grid<true><<<gn, MGFB_THREADS, 0, sA>>>(B.fat, B.col, B.gid, n, G, V, c);
        {
            int gw = std::max(1, std::min((int)((slots + BP_WARPS * BP_PER - 1) / (BP_WARPS * BP_PER)), ctx->num_sms * 8));
            if (ctx->max_ctas) gw = std::min(gw, ctx->max_ctas * 4);
            k_body_pairs_warp<true><<<gw, MGFB_THREADS, 0, sA>>>(B.tight, B.col, B.gid, tiled ? n : 0xffffffffu, G, PL, ctx->pair_cap, c, V, n);
        }
        CU(cudaEventRecord(ctx->ev_bp[1], sA));
        PROF(MGFB_PHASE_PAIR_SWEEP);
        // -- [B] coherent: S -> this step's pair lists
        k_bp_mark<<<grid_for(ctx, BP_REF_CAP), MGFB_THREADS, 0, ctx->stream>>>(B.fat, V, c);
        CU(cudaEventRecord(ctx->ev_bp[2], ctx->stream));
        k_bp_filter<<<grid_for(ctx, (size_t)ctx->bp_s_cap / BP_FILTER_ITEMS + 1), MGFB_THREADS, 0, ctx->stream>>>(B.tight, B.fat, V, PL, ctx->pair_cap, c);
        // -- [C] replaced bodies (and ghosts) queried against the cached grid and the overflow list
        if (tiled) CU(cudaStreamWaitEvent(sC, ctx->ev_bp[1], 0));   // (only a rebuild's ghost queries need the new grid; untiled worlds have no ghosts)
        CU(cudaStreamWaitEvent(sC, ctx->ev_bp[2], 0));
        k_bp_query<<<grid_for(ctx, (size_t)(BP_REF_CAP / 4 + ctx->ghost_cap) * 32), MGFB_THREADS, 0, sC>>>(B.tight, B.fat, B.col, B.gid, n, G, V, PL, ctx->pair_cap, c);
        k_bp_query_ovf<<<ctx->num_sms * 4, MGFB_THREADS, 0, sC>>>(B.tight, B.fat, B.col, B.gid, n, V, PL, ctx->pair_cap, c);
        CU(cudaEventRecord(ctx->ev_bp[3], sC));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_bp[1], 0));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_bp[3], 0));
        k_bp_finish<<<1, 1, 0, ctx->stream>>>(V, c);
        ctx->launches += 6;   // (10 launches where the sweep path has 4)
    } else {
    k_bgrid_insert<false><<<gs, MGFB_THREADS, 0, ctx->stream>>>(B.fat, B.col, B.gid, G, c);
    TRY(scan_u32_lb(ctx, G.cell_count, G.cell_start, ctx->table, &c->grid_entries, step_scan_state(ctx, 0)));
    k_bgrid_insert<true><<<gs, MGFB_THREADS, 0, ctx->stream>>>(B.fat, B.col, B.gid, G, c);
    PROF(MGFB_PHASE_PAIR_SWEEP);
    {
        int gw = std::max(1, std::min((int)((slots + BP_WARPS * BP_PER - 1) / (BP_WARPS * BP_PER)), ctx->num_sms * 8));
        if (ctx->max_ctas) gw = std::min(gw, ctx->max_ctas * 4);
        k_body_pairs_warp<false><<<gw, MGFB_THREADS, 0, ctx->stream>>>(B.tight, B.col, B.gid, tiled ? n : 0xffffffffu, G, PL, ctx->pair_cap, c, BpView{}, 0u);
    }
    }
    ContactList L = contact_list(ctx);
    // the chain colouring's per-constraint setup rides on the contact emission (one launch and one pass less)
    EmitHook H{};
    const bool count_fused = ctx->cfg.solver_schedule != MGFB_SCHEDULE_PHASES_JP;
    if (count_fused) {
        H.gid = B.gid; H.x = B.x; H.col = B.col;
        H.key = ctx->c_key.as<unsigned long long>(); H.deg = ctx->body_deg.as<unsigned>(); H.inbox = ctx->c_inbox.as<unsigned long long>();
        H.next = ctx->c_next.as<unsigned>(); H.group = ctx->group.as<int>(); H.cap = ctx->row_cap;
    }
    TerrainView T{};
    bool caps = ctx->n_capsules > 0, sph = ctx->n_capsules < ctx->n;
    int gp = grid_for(ctx, (size_t)slots * 4);
    if (ctx->terrain.present) {
        // body x terrain (mesh query + its narrowphase) shares nothing with body x body but the contact list's cursor:
        // it runs on a side stream beside the body grid / pair sweep / body narrowphase and joins before the colouring
        T = terrain_view(ctx);
        PairLists TL; for (int k = 0; k < 4; ++k) TL.p[k] = k < 2 ? ctx->tpair_list[k].as<int2>() : nullptr;
        CU(cudaStreamWaitEvent(ctx->s_aux, ctx->ev_fork, 0));
        if (ctx->prof_on) CU(cudaEventRecord(ctx->prof_ev[9], ctx->s_aux));
        k_terrain_pairs<<<gb, MGFB_THREADS, 0, ctx->s_aux>>>(B.tight, B.col, n, T, TL, ctx->tpair_cap, c);
        if (sph) k_narrow_terrain<0><<<gp, MGFB_THREADS, 0, ctx->s_aux>>>(B.col, ctx->tpair_list[0].as<int2>(), T, L, ctx->contact_cap, c, H);
        if (caps) k_narrow_terrain<1><<<gp, MGFB_THREADS, 0, ctx->s_aux>>>(B.col, ctx->tpair_list[1].as<int2>(), T, L, ctx->contact_cap, c, H);
        if (ctx->prof_on) CU(cudaEventRecord(ctx->prof_ev[10], ctx->s_aux));
        CU(cudaEventRecord(ctx->ev_join, ctx->s_aux));
    }
    // narrowphase, one specialisation per shape pair
    PROF(MGFB_PHASE_NARROW_BODIES);
    if (sph) k_narrow_bodies<0, 0><<<gp, MGFB_THREADS, 0, ctx->stream>>>(B.col, PL.p[0], L, ctx->contact_cap, c, H);
    if (sph && caps) {
        k_narrow_bodies<0, 1><<<gp, MGFB_THREADS, 0, ctx->stream>>>(B.col, PL.p[1], L, ctx->contact_cap, c, H);
        k_narrow_bodies<1, 0><<<gp, MGFB_THREADS, 0, ctx->stream>>>(B.col, PL.p[2], L, ctx->contact_cap, c, H);
    }
    if (caps) k_narrow_bodies<1, 1><<<gp, MGFB_THREADS, 0, ctx->stream>>>(B.col, PL.p[3], L, ctx->contact_cap, c, H);
    if (ctx->terrain.present) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    CU(cudaGetLastError());
    // constraints: colour, build rows in solve order, solve
    OrderView O = order_view(ctx, L.a, L.b, L.face, L.sub);
    ManifoldInput M{};
    M.a = L.a; M.b = L.b; M.la = L.la; M.lb = L.lb; M.nt = L.nt; M.user = false;
    M.terrain_center = make_float4(ctx->terrain.x[0], ctx->terrain.x[1], ctx->terrain.x[2], 0.0f);
    TRY(enqueue_order_and_solve(ctx, O, M, &c->contacts, 0, ctx->contact_cap, false, dt, iters, timed, true, count_fused));
    k_step_done<<<1, 64, 0, ctx->stream>>>(c, ctx->ctr_snap);
    CU(cudaGetLastError());
    // k_integrate, 2x k_grid_insert, scan, k_body_pairs, k_step_done (+ terrain, narrowphase)
    ctx->launches += 6 + (ctx->terrain.present ? 1 : 0) + (sph ? 1 : 0) + (sph && caps ? 2 : 0) + (caps ? 1 : 0) +
                     (ctx->terrain.present ? (sph ? 1 : 0) + (caps ? 1 : 0) : 0);
    return MGFB_OK;
}

int32_t read_counters(mgfb_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->h_ctr, dctr(ctx), sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}
int32_t clear_sticky(mgfb_ctx* ctx) {
    ctx->bp_invalidate = true;   // whatever tripped the flag, the cached broadphase starts over (bpcache.cuh)
    CU(cudaMemsetAsync(reinterpret_cast<char*>(dctr(ctx)) + offsetof(Counters, overflow), 0, 8, ctx->stream));
    return MGFB_OK;
}

Collider shape_to_collider(const mgfb_shape& s, float half_h) {
    Collider k;
    if (s.kind == MGFB_SPHERE) {
        k.p0 = make_float4(s.p[0], s.p[1], s.p[2], s.p[3]);
        k.p1 = make_float4(0, 0, 0, ibits(0));
    } else {
        k.p0 = make_float4(s.p[0], s.p[1], s.p[2], s.p[6]);
        k.p1 = make_float4(s.p[3], s.p[4], s.p[5], ibits(1));
    }
    k.v = make_float4(s.v[0], s.v[1], s.v[2], half_h);
    return k;
}

// physics.rs:30-46
M3 sphere_tensor(V3 c, float r, float m) {
    float i = 0.4f * m * r * r;
    M3 I = m_diag(i, i, i);
    M3 outer = mkm(c * c.x, c * c.y, c * c.z);
    return madd(I, smul(m, msub(mscale(m_ident(), dot3(c, c)), outer)));
}
// physics.rs:48-84
M3 capsule_tensor(V3 a, V3 d, float r, float m) {
    float h = len(d);
    float mh = m * 2.0f * r / (4.0f * r + 3.0f * h);
    float mc = m * h / (4.0f / 3.0f * r + h);
    float ic_x = 1.0f / 12.0f * mc * (3.0f * r * r + h * h);
    float ic_y = 0.5f * mc * r * r;
    float ic_z = ic_x;
    float is_x = mh * (3.0f * r + 2.0f * h) / 4.0f * h;
    float is_y = 4.0f / 5.0f * mh * r * r;
    float is_z = is_x;
    float i_x = ic_x + is_x, i_y = ic_y + is_y, i_z = ic_z + is_z;
    M3 rot = m_from_q(q_from_arc(mk3(0.0f, 1.0f, 0.0f) * h, d));
    M3 I = mmul(mmul(rot, m_diag(i_x, i_y, i_z)), mtrans(rot));
    V3 disp = a + d * 0.5f;
    M3 outer = mkm(disp * disp.x, disp * disp.y, disp * disp.z);
    return madd(I, smul(m, msub(mscale(m_ident(), dot3(disp, disp)), outer)));
}

}  // namespace

void pipe_destroy(mgfb_ctx* ctx);   // pipeline.cuh
namespace { void reforder_free(mgfb_ctx* ctx); int32_t step_reference_order(mgfb_ctx* ctx, float dt, unsigned iters, mgfb_step_stats* stats); }   // reforder.cuh
namespace { int32_t local_handover_selftest(mgfb_ctx* ctx, unsigned rounds, unsigned* torn, unsigned* observed);
            int32_t run_handover_selftest(mgfb_ctx* ctx, Inbox* box, bool sys, unsigned rounds, unsigned* torn, unsigned* observed); }   // selftest.cuh

// ============================================================================ C ABI
extern "C" {

int32_t mgfb_abi_version(void) { return MGFB_ABI_VERSION; }

void mgfb_config_default(mgfb_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->penetration_slop = 0.05f;
    cfg->baumgarte = 0.2f;
    cfg->persistent_threshold_sq = 0.5f;
    cfg->fat_margin = 0.25f;
}

static std::string g_create_err;
const char* mgfb_last_error(const mgfb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int32_t mgfb_ctx_create(const mgfb_config* cfg, mgfb_ctx** out) {
    if (!out) return MGFB_ERR_INVALID_ARG;
    *out = nullptr;
    mgfb_ctx* ctx = new mgfb_ctx();
    if (cfg) ctx->cfg = *cfg; else mgfb_config_default(&ctx->cfg);
    // "sweep": grid + sweep every step; "cache": the coherent broadphase even on tiled worlds (default: untiled worlds only)
    { const char* e = getenv("MGFB_BROADPHASE"); ctx->bp_on = !(e && std::strcmp(e, "sweep") == 0); ctx->bp_forced = e && std::strcmp(e, "cache") == 0; }
    ctx->device = ctx->cfg.device;
    auto bail = [&](cudaError_t e, const char* what) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(e);
        delete ctx;
        return (int32_t)MGFB_ERR_CUDA;
    };
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return bail(e == cudaSuccess ? cudaErrorNoDevice : e, "no CUDA device (mgfb has no CPU fallback)");
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->num_sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) return bail(cudaErrorNotSupported, "device lacks cooperative launch");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (auto& ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    ctx->cur_ev = ctx->ev;
    if ((e = cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto& st : ctx->s_bp) if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (auto& ev : ctx->ev_bp) if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMallocHost(&ctx->h_ctr, sizeof(Counters))) != cudaSuccess) return bail(e, "cudaMallocHost");
    if ((e = cudaMalloc(&ctx->ctr.p, sizeof(Counters))) != cudaSuccess) return bail(e, "cudaMalloc");
    ctx->ctr.bytes = sizeof(Counters);
    cudaMemsetAsync(ctx->ctr.p, 0, sizeof(Counters), ctx->stream);
    {
        // Load every step kernel NOW.  With CUDA's lazy module loading the first launch of a kernel
        // may have to wait for the device to go idle -- which never happens while a neighbour tile's
        // kernel is spinning on a flag this very launch would set.
        const void* fns[] = {(const void*)k_integrate<true, true, true, false>, (const void*)k_integrate<true, true, true, true>,
                             (const void*)k_integrate<false, false, true, false>, (const void*)k_tile_publish, (const void*)k_tile_wait,
                             (const void*)k_ghost_send, (const void*)k_ghost_recv, (const void*)k_bgrid_insert<false>,
                             (const void*)k_bgrid_insert<true>, (const void*)k_scan_reduce, (const void*)k_scan_sums, (const void*)k_scan_final, (const void*)k_scan_lookback,
                             (const void*)k_body_pairs_warp<false>, (const void*)k_terrain_pairs, (const void*)k_narrow_bodies<0, 0>,
                             (const void*)k_narrow_bodies<0, 1>, (const void*)k_narrow_bodies<1, 0>, (const void*)k_narrow_bodies<1, 1>,
                             (const void*)k_narrow_terrain<0>, (const void*)k_narrow_terrain<1>, (const void*)k_order, (const void*)k_group_scan,
                             (const void*)k_scatter_rows, (const void*)k_build_rows, (const void*)k_solve<false>, (const void*)k_solve<true>,
                             (const void*)k_solve_df<false>, (const void*)k_solve_df<true>, (const void*)k_df_init<false>, (const void*)k_df_init<true>,
                             (const void*)k_tile_links_send, (const void*)k_tile_solve_done, (const void*)k_inc_count, (const void*)k_inc_fill, (const void*)k_inc_sort, (const void*)k_colour_df,
                             (const void*)k_zero_ranges, (const void*)k_bp_decide, (const void*)k_bp_finish, (const void*)k_bp_grid<false>, (const void*)k_bp_grid<true>,
                             (const void*)k_body_pairs_warp<true>, (const void*)k_bp_mark, (const void*)k_bp_filter, (const void*)k_bp_query, (const void*)k_bp_query_ovf,
                             (const void*)k_step_done, (const void*)k_pack_state, (const void*)k_set_velocity<false>, (const void*)k_set_velocity<true>, (const void*)k_set_state};
        cudaFuncAttributes fa;
        for (const void* f : fns) if ((e = cudaFuncGetAttributes(&fa, f)) != cudaSuccess) return bail(e, "cudaFuncGetAttributes");
    }
    ctx->max_ctas = (int)ctx->cfg.max_cooperative_ctas;
    if (ctx->cfg.tile_timeout_ms) ctx->tile_timeout_ns = (unsigned long long)ctx->cfg.tile_timeout_ms * 1000000ULL;
    ctx->coop_order = coop_blocks(ctx, k_order, MGFB_THREADS, 2);
    ctx->coop_colour = coop_blocks(ctx, k_colour_df, MGFB_THREADS, 8);
    ctx->coop_solve = std::min(coop_blocks(ctx, k_solve<false>, MGFB_SOLVE_THREADS, 1), coop_blocks(ctx, k_solve<true>, MGFB_SOLVE_THREADS, 1));
    ctx->coop_df = std::min(coop_blocks(ctx, k_solve_df<false>, MGFB_DF_THREADS_LARGE, 1), coop_blocks(ctx, k_solve_df<true>, MGFB_DF_THREADS_LARGE, 1));
    int32_t s = grow_bodies(ctx, std::max(ctx->cfg.initial_body_capacity, 1024u));
    if (s == MGFB_OK) s = ensure_rows(ctx, 1024, false, 4096);
    // the dataflow solver's hand-over is one 32-byte store: check that this device delivers it whole (selftest.cuh)
    if (s == MGFB_OK && ctx->cfg.solver_schedule == MGFB_SCHEDULE_DATAFLOW) {
        s = local_handover_selftest(ctx, 512, nullptr, nullptr);
        if (s == MGFB_OK && ctx->handover_torn) ctx->cfg.solver_schedule = MGFB_SCHEDULE_PHASES;
    }
    if (s != MGFB_OK) { g_create_err = ctx->err; delete ctx; return s; }
    *out = ctx;
    return MGFB_OK;
}

void mgfb_ctx_destroy(mgfb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
#ifdef MGFB_DF_PROFILE
    {
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_df_prof, sizeof(h));
        double v = (double)std::max(1ULL, h[4]);
        fprintf(stderr, "[df profile] warps=%llu visits=%llu per visit (cycles): fetch %.0f inbox-poll %.0f compute+publish %.0f ; polls %.2f\n",
                h[5], h[4], h[0] / v, h[1] / v, h[2] / v, h[3] / v);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_df_prof, z, sizeof(z));
    }
#endif
    Buf* all[] = {&ctx->x, &ctx->q, &ctx->vel, &ctx->force, &ctx->torque, &ctx->imb, &ctx->col, &ctx->tight, &ctx->fat, &ctx->ctr,
                  &ctx->pair_list[0], &ctx->pair_list[1], &ctx->pair_list[2], &ctx->pair_list[3], &ctx->tpair_list[0], &ctx->tpair_list[1],
                  &ctx->c_a, &ctx->c_b, &ctx->c_face, &ctx->c_sub, &ctx->c_la, &ctx->c_lb, &ctx->c_nt, &ctx->body_best, &ctx->body_scratch,
                  &ctx->group, &ctx->group_count, &ctx->group_start, &ctx->perm, &ctx->r_ab, &ctx->r_n, &ctx->r_t0, &ctx->r_t1, &ctx->r_ra,
                  &ctx->r_rb, &ctx->r_imp, &ctx->r_xra, &ctx->r_xrb, &ctx->r_xtm, &ctx->r_dep, &ctx->body_deg, &ctx->body_start, &ctx->r_inc, &ctx->r_next, &ctx->r_in_a, &ctx->r_in_b, &ctx->r_ia, &ctx->c_key, &ctx->c_csr, &ctx->c_next, &ctx->c_inbox, &ctx->cell_count, &ctx->cell_start, &ctx->bg_ent,
                  &ctx->scan_sums, &ctx->scan_state, &ctx->scan_step, &ctx->bp_state, &ctx->bp_s[0], &ctx->bp_s[1], &ctx->bp_stale, &ctx->bp_c0, &ctx->bp_ovf, &ctx->bp_ref_flag, &ctx->bp_ref_list, &ctx->u_a, &ctx->u_b, &ctx->u_sc, &ctx->u_sf, &ctx->u_n, &ctx->u_t, &ctx->u_nc, &ctx->u_la,
                  &ctx->u_lb, &ctx->stage, &ctx->convex_pool, &ctx->terrain.verts, &ctx->terrain.faces, &ctx->terrain.boxes, &ctx->terrain.cell_count,
                  &ctx->terrain.cell_start, &ctx->terrain.ent_id, &ctx->terrain.ent_key, &ctx->terrain.max_bits,
                  &ctx->gid, &ctx->phase_start, &ctx->edge_idx, &ctx->edge_mark, &ctx->ridx, &ctx->mbox, &ctx->edge_slot, &ctx->tile_df};
    pipe_destroy(ctx);
    reforder_free(ctx);
    for (void* ptr : ctx->ipc_opened) cudaIpcCloseMemHandle(ptr);
    for (Buf* b : all) release(*b);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& e : ctx->prof_ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (auto& st : ctx->s_bp) if (st) cudaStreamDestroy(st);
    for (auto& ev : ctx->ev_bp) if (ev) cudaEventDestroy(ev);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    for (auto& st : ctx->gjk_streams) if (st) cudaStreamDestroy(st);
    for (auto& ev : ctx->gjk_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int32_t mgfb_synchronize(mgfb_ctx* ctx) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_bodies_count(const mgfb_ctx* ctx, uint32_t* n) {
    if (!ctx || !n) return MGFB_ERR_INVALID_ARG;
    *n = ctx->n;
    return MGFB_OK;
}

int32_t mgfb_bodies_add(mgfb_ctx* ctx, uint32_t n, const mgfb_shape* shapes, const float* mass, const float* restitution,
                        const float* friction, const float* world_force, uint32_t* first_id) {
    if (ctx) ctx->bp_invalidate = true;   // the stored fat boxes / the body set change under the cached broadphase
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (n == 0) { if (first_id) *first_id = ctx->n; return MGFB_OK; }
    if (!shapes || !mass || !restitution || !friction || !world_force) return fail(ctx, MGFB_ERR_INVALID_ARG, "null input array");
    if ((unsigned long long)ctx->n + n >= (1ULL << 29)) return fail(ctx, MGFB_ERR_CAPACITY, "a context holds fewer than 2^29 bodies (bits of the pair words)");
    CU(cudaSetDevice(ctx->device));
    std::vector<float4> hx(n), hq(n), hforce(n), htorque(n), himb((size_t)n * 3);
    std::vector<BodyVel> hvel(n);
    std::vector<Collider> hcol(n);
    std::vector<Box> htight(n), hfat(n);
    std::vector<unsigned> hgid(n);
    if (ctx->tile_exported) return fail(ctx, MGFB_ERR_STATE, "bodies cannot be added after mgfb_tile_export");
    unsigned ncaps = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const mgfb_shape& s = shapes[i];
        V3 x; Q4 q; float half_h = 0.0f; M3 tensor;
        float m = mass[i];
        if (s.kind == MGFB_SPHERE) {  // compound.rs:44-45
            float r = s.p[3];
            if (!(r > 0.0f)) return fail(ctx, MGFB_ERR_INVALID_ARG, "sphere radius must be > 0 (geom.rs:300)");
            x = mk3(s.p[0], s.p[1], s.p[2]); q = qident();
            V3 c0 = x + (-x);  // collider - x (physics.rs:212)
            tensor = sphere_tensor(c0, r, m);
        } else if (s.kind == MGFB_CAPSULE) {  // compound.rs:46-50
            float r = s.p[6];
            if (!(r > 0.0f)) return fail(ctx, MGFB_ERR_INVALID_ARG, "capsule radius must be > 0 (geom.rs:328)");
            V3 a = mk3(s.p[0], s.p[1], s.p[2]), d = mk3(s.p[3], s.p[4], s.p[5]);
            float h = len(d);
            q = q_from_arc(mk3(0.0f, 1.0f, 0.0f) * h, d);
            x = a + d * 0.5f; half_h = h * 0.5f;
            tensor = capsule_tensor(a + (-x), d, r, m);
            ncaps++;
        } else {
            return fail(ctx, MGFB_ERR_INVALID_ARG, "RigidBodyVec holds Component::{Sphere,Capsule} only (physics.rs:153-154)");
        }
        M3 inv;
        if (!minv(tensor, &inv)) return fail(ctx, MGFB_ERR_SINGULAR_INERTIA, "inertia tensor is singular (physics.rs:212 unwrap)");
        hx[i] = v4(x, 0.0f);
        hq[i] = make_float4(q.s, q.v.x, q.v.y, q.v.z);
        BodyVel bv; bv.a = make_float4(0, 0, 0, 0); bv.b = make_float4(0, 0, 1.0f / m, 0);
        bv.c = make_float4(0, 0, 0, 0); bv.d = make_float4(0, 0, 0, 0);
        vel_set_inertia(bv, inv);
        hvel[i] = bv;
        V3 wf = mk3(world_force[3 * i], world_force[3 * i + 1], world_force[3 * i + 2]);
        hforce[i] = v4(wf * m, restitution[i]);          // physics.rs:207
        htorque[i] = make_float4(0, 0, 0, friction[i]);
        himb[3 * i] = v4(inv.c0, 0); himb[3 * i + 1] = v4(inv.c1, 0); himb[3 * i + 2] = v4(inv.c2, 0);
        mgfb_shape s0 = s; s0.v[0] = s0.v[1] = s0.v[2] = 0.0f;  // Moving::sweep(collider, 0)
        hcol[i] = shape_to_collider(s0, half_h);
        V3 tc, tr;
        if (!swept_bounds(hcol[i], &tc, &tr)) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125)");
        htight[i].c = v4(tc, 0); htight[i].r = v4(tr, 0);
        float mg = ctx->cfg.fat_margin;
        hfat[i].c = v4(tc, 0); hfat[i].r = v4(tr + mk3(mg, mg, mg), 0);  // world.rs:180-181
        hgid[i] = ctx->n + i;
    }
    unsigned first = ctx->n;
    TRY(grow_bodies(ctx, first + n));
    auto up = [&](Buf& b, const void* src, size_t elem, size_t per = 1) {
        return cudaMemcpyAsync(reinterpret_cast<char*>(b.p) + (size_t)first * elem * per, src, (size_t)n * elem * per, cudaMemcpyHostToDevice, ctx->stream);
    };
    CU(up(ctx->x, hx.data(), sizeof(float4))); CU(up(ctx->q, hq.data(), sizeof(float4))); CU(up(ctx->vel, hvel.data(), sizeof(BodyVel)));
    CU(up(ctx->force, hforce.data(), sizeof(float4))); CU(up(ctx->torque, htorque.data(), sizeof(float4)));
    CU(up(ctx->imb, himb.data(), sizeof(float4), 3)); CU(up(ctx->col, hcol.data(), sizeof(Collider)));
    CU(up(ctx->tight, htight.data(), sizeof(Box))); CU(up(ctx->fat, hfat.data(), sizeof(Box)));
    CU(up(ctx->gid, hgid.data(), 4));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n += n; ctx->n_capsules += ncaps;
    if (first_id) *first_id = first;
    // reference-order replay (reforder.cuh): World::add_body inserts the new leaves into the body BVH as it is (world.rs:178-184)
    if (ctx->reforder) {
        RefOrderState* S = ctx->reforder;
        for (uint32_t i = 0; i < n; ++i) {
            S->fat.push_back(hfat[i]);
            S->leaf.push_back(S->body_tree.insert(f4v(hfat[i].c), f4v(hfat[i].r), (int)(first + i)));
            if (S->leaf.back() < 0) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)");
        }
    }
    return MGFB_OK;
}

int32_t mgfb_bodies_get_state(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* x, float* q, float* v, float* omega) {
    if (!ctx || (uint64_t)first + n > ctx->n) return fail(ctx, MGFB_ERR_INVALID_ARG, "body range out of bounds");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->stage, (size_t)n * 13 * 4));
    float* sx = ctx->stage.as<float>(); float* sq = sx + (size_t)3 * n; float* sv = sq + (size_t)4 * n; float* sw = sv + (size_t)3 * n;
    k_pack_state<<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), first, n, x ? sx : nullptr,
                                                                                           q ? sq : nullptr, v ? sv : nullptr, omega ? sw : nullptr);
    CU(cudaGetLastError());
    ctx->launches += 1;
    if (x) CU(cudaMemcpyAsync(x, sx, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (q) CU(cudaMemcpyAsync(q, sq, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (v) CU(cudaMemcpyAsync(v, sv, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (omega) CU(cudaMemcpyAsync(omega, sw, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_bodies_set_velocity(mgfb_ctx* ctx, uint32_t first, uint32_t n, const float* v, const float* omega) {
    if (!ctx || (uint64_t)first + n > ctx->n || !v || !omega) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->stage, (size_t)n * 13 * 4));
    float* sv = ctx->stage.as<float>(); float* sw = sv + (size_t)3 * n;
    CU(cudaMemcpyAsync(sv, v, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(sw, omega, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_set_velocity<false><<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), first, n, sv, sw, nullptr);
    CU(cudaGetLastError());
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_bodies_get_colliders(mgfb_ctx* ctx, uint32_t first, uint32_t n, mgfb_shape* out) {
    if (!ctx || (uint64_t)first + n > ctx->n || !out) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    std::vector<Collider> hc(n);
    CU(cudaMemcpyAsync(hc.data(), ctx->col.as<Collider>() + first, (size_t)n * sizeof(Collider), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n; ++i) {
        mgfb_shape s; std::memset(&s, 0, sizeof(s));
        const Collider& k = hc[i];
        if (col_kind(k) == 0) { s.kind = MGFB_SPHERE; s.p[0] = k.p0.x; s.p[1] = k.p0.y; s.p[2] = k.p0.z; s.p[3] = k.p0.w; }
        else { s.kind = MGFB_CAPSULE; s.p[0] = k.p0.x; s.p[1] = k.p0.y; s.p[2] = k.p0.z; s.p[3] = k.p1.x; s.p[4] = k.p1.y; s.p[5] = k.p1.z; s.p[6] = k.p0.w; }
        s.v[0] = k.v.x; s.v[1] = k.v.y; s.v[2] = k.v.z;
        out[i] = s;
    }
    return MGFB_OK;
}

int32_t mgfb_bodies_get_inv_moment(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* out) {
    if (!ctx || (uint64_t)first + n > ctx->n || !out) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    std::vector<BodyVel> bv(n);
    CU(cudaMemcpyAsync(bv.data(), ctx->vel.as<BodyVel>() + first, (size_t)n * sizeof(BodyVel), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n; ++i) {
        M3 I = vel_inertia(bv[i]);
        float* o = out + 9 * i;
        o[0] = I.c0.x; o[1] = I.c0.y; o[2] = I.c0.z; o[3] = I.c1.x; o[4] = I.c1.y; o[5] = I.c1.z; o[6] = I.c2.x; o[7] = I.c2.y; o[8] = I.c2.z;
    }
    return MGFB_OK;
}

int32_t mgfb_bodies_get_fat_bounds(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* boxes) {
    if (!ctx || (uint64_t)first + n > ctx->n || !boxes) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    std::vector<Box> hb(n);
    CU(cudaMemcpyAsync(hb.data(), ctx->fat.as<Box>() + first, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n; ++i) {
        float* o = boxes + 6 * i;
        o[0] = hb[i].c.x; o[1] = hb[i].c.y; o[2] = hb[i].c.z; o[3] = hb[i].r.x; o[4] = hb[i].r.y; o[5] = hb[i].r.z;
    }
    return MGFB_OK;
}

int32_t mgfb_bodies_set_state(mgfb_ctx* ctx, uint32_t first, uint32_t n, const float* x, const float* q, const float* v, const float* omega,
                              const mgfb_shape* colliders, const float* fat_boxes) {
    if (ctx) ctx->bp_invalidate = true;   // the stored fat boxes / the body set change under the cached broadphase
    if (!ctx || (uint64_t)first + n > ctx->n) return fail(ctx, MGFB_ERR_INVALID_ARG, "body range out of bounds");
    if (n == 0) return MGFB_OK;
    if (ctx->pipe_inflight) return fail(ctx, MGFB_ERR_STATE, "steps are in flight: mgfb_step_wait first");
    CU(cudaSetDevice(ctx->device));
    std::vector<float> hc;
    if (colliders) {
        std::vector<Collider> cur(n);
        CU(cudaMemcpyAsync(cur.data(), ctx->col.as<Collider>() + first, (size_t)n * sizeof(Collider), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        hc.resize((size_t)n * 9);
        for (uint32_t i = 0; i < n; ++i) {
            const mgfb_shape& s = colliders[i];
            if ((int)s.kind != col_kind(cur[i])) return fail(ctx, MGFB_ERR_INVALID_ARG, "a collider cannot change its Component kind (physics.rs:153 constructor)");
            const float r_new = s.kind == MGFB_SPHERE ? s.p[3] : s.p[6];
            if (r_new != cur[i].p0.w) return fail(ctx, MGFB_ERR_INVALID_ARG, "a collider cannot change its radius (the inertia tensor was built for it, physics.rs:212)");
            float* c = hc.data() + 9 * (size_t)i;
            c[0] = s.p[0]; c[1] = s.p[1]; c[2] = s.p[2];
            c[3] = s.kind == MGFB_CAPSULE ? s.p[3] : 0.0f; c[4] = s.kind == MGFB_CAPSULE ? s.p[4] : 0.0f; c[5] = s.kind == MGFB_CAPSULE ? s.p[5] : 0.0f;
            c[6] = s.v[0]; c[7] = s.v[1]; c[8] = s.v[2];
        }
    }
    if (fat_boxes) for (uint32_t i = 0; i < n; ++i) for (int k = 3; k < 6; ++k)
        if (!(fat_boxes[6 * (size_t)i + k] >= 0.0f)) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB half extent must be >= 0 (bounds.rs:125)");
    TRY(ensure(ctx, ctx->stage, (size_t)n * 28 * 4));
    float* sx = ctx->stage.as<float>(); float* sq = sx + (size_t)3 * n; float* sv = sq + (size_t)4 * n; float* sw = sv + (size_t)3 * n;
    float* sc = sw + (size_t)3 * n; float* sf = sc + (size_t)9 * n;
    auto up = [&](float* dst, const float* src, size_t per) { return src ? cudaMemcpyAsync(dst, src, (size_t)n * per * 4, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess; };
    CU(up(sx, x, 3)); CU(up(sq, q, 4)); CU(up(sv, v, 3)); CU(up(sw, omega, 3)); CU(up(sc, colliders ? hc.data() : nullptr, 9)); CU(up(sf, fat_boxes, 6));
    k_set_state<<<(n + MGFB_THREADS - 1) / MGFB_THREADS, MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), first, n, x ? sx : nullptr, q ? sq : nullptr,
                                                                                          v ? sv : nullptr, omega ? sw : nullptr, colliders ? sc : nullptr,
                                                                                          fat_boxes ? sf : nullptr);
    CU(cudaGetLastError());
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    if (fat_boxes && ctx->reforder) ctx->reforder->leaf.clear();   // the stored boxes changed: the body tree is rebuilt from them, in body order
    return MGFB_OK;
}

int32_t mgfb_integrate(mgfb_ctx* ctx, float dt) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (ctx->n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    k_integrate<false, true, false><<<grid_for(ctx, ctx->n), MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), ctx->n, dt, ctx->cfg.fat_margin, dctr(ctx));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_complete_motion(mgfb_ctx* ctx) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (ctx->n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    k_integrate<true, false, false><<<grid_for(ctx, ctx->n), MGFB_THREADS, 0, ctx->stream>>>(body_arrays(ctx), ctx->n, 0.0f, ctx->cfg.fat_margin, dctr(ctx));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_terrain_set(mgfb_ctx* ctx, const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces, const float x[3]) {
    if (!ctx || !verts || !faces || !x) return fail(ctx, MGFB_ERR_INVALID_ARG, "null terrain arrays");
    if (nfaces >= (1u << 30)) return fail(ctx, MGFB_ERR_INVALID_ARG, "too many faces");
    for (uint32_t i = 0; i < nfaces * 3; ++i)
        if (faces[i] >= nverts) return fail(ctx, MGFB_ERR_INVALID_ARG, "face index out of range (Rust bounds panic, mesh.rs:65-67)");
    CU(cudaSetDevice(ctx->device));
    TerrainData& t = ctx->terrain;
    t.present = false;
    t.nverts = nverts; t.nfaces = nfaces;
    t.x[0] = x[0]; t.x[1] = x[1]; t.x[2] = x[2];
    if (nfaces == 0) return MGFB_OK;
    std::vector<float4> hv(nverts); std::vector<uint4> hf(nfaces);
    for (uint32_t i = 0; i < nverts; ++i) hv[i] = make_float4(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], 0);
    for (uint32_t i = 0; i < nfaces; ++i) hf[i] = make_uint4(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2], 0);
    TRY(ensure(ctx, t.verts, (size_t)nverts * 16)); TRY(ensure(ctx, t.faces, (size_t)nfaces * 16)); TRY(ensure(ctx, t.boxes, (size_t)nfaces * sizeof(Box)));
    TRY(ensure(ctx, t.max_bits, 16));
    CU(cudaMemcpyAsync(t.verts.p, hv.data(), (size_t)nverts * 16, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(t.faces.p, hf.data(), (size_t)nfaces * 16, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(t.max_bits.p, 0, 16, ctx->stream));
    k_face_boxes<<<(nfaces + 255) / 256, 256, 0, ctx->stream>>>(t.verts.as<float4>(), t.faces.as<uint4>(), nfaces, t.boxes.as<Box>(), t.max_bits.as<unsigned>());
    CU(cudaGetLastError());
    unsigned bits = 0;
    CU(cudaMemcpyAsync(&bits, t.max_bits.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float cell; std::memcpy(&cell, &bits, 4);
    if (!(cell > 0.0f) || !std::isfinite(cell)) cell = 1.0f;
    t.inv_cell = 1.0f / cell;
    t.table = next_pow2(std::max(1024u, nfaces * 4));
    TRY(ensure(ctx, t.cell_count, (size_t)t.table * 4)); TRY(ensure(ctx, t.cell_start, ((size_t)t.table + 1) * 4));
    Buf sums; TRY(ensure(ctx, sums, ((size_t)t.table / SCAN_ITEMS + 2) * 4));
    // entries: count first (the count pass needs no entry storage), then size exactly
    GridView G; G.cell_count = t.cell_count.as<unsigned>(); G.cell_start = t.cell_start.as<unsigned>();
    G.ent_id = nullptr; G.ent_key = nullptr; G.table_mask = t.table - 1; G.ent_cap = 0;
    CU(cudaMemsetAsync(t.cell_count.p, 0, (size_t)t.table * 4, ctx->stream));
    int gb = grid_for(ctx, nfaces);
    k_grid_insert<false><<<gb, MGFB_THREADS, 0, ctx->stream>>>(t.boxes.as<Box>(), nfaces, G, dctr(ctx), t.inv_cell);
    TRY(scan_u32(ctx, G.cell_count, G.cell_start, t.table, sums.as<unsigned>(), t.max_bits.as<unsigned>() + 1));
    unsigned total = 0;
    CU(cudaMemcpyAsync(&total, t.max_bits.as<unsigned>() + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    t.ent_cap = std::max(total, 1u);
    TRY(ensure(ctx, t.ent_id, (size_t)t.ent_cap * 4)); TRY(ensure(ctx, t.ent_key, (size_t)t.ent_cap * 8));
    G.ent_id = t.ent_id.as<unsigned>(); G.ent_key = t.ent_key.as<unsigned long long>(); G.ent_cap = t.ent_cap;
    k_grid_insert<true><<<gb, MGFB_THREADS, 0, ctx->stream>>>(t.boxes.as<Box>(), nfaces, G, dctr(ctx), t.inv_cell);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    release(sums);
    t.present = true;
    if (ctx->reforder) ctx->reforder->mesh_built = false;
    return MGFB_OK;
}

static void fill_step_stats(mgfb_ctx* ctx, mgfb_step_stats* st, unsigned iters, bool timed, unsigned overflowed, const Counters* hc = nullptr,
                            cudaEvent_t* evs = nullptr) {
    const Counters& h = hc ? *hc : *ctx->h_ctr;
    if (!evs) evs = ctx->ev;
    std::memset(st, 0, sizeof(*st));
    st->bodies = ctx->n;
    st->candidate_pairs = h.pairs[0] + h.pairs[1] + h.pairs[2] + h.pairs[3];
    st->terrain_candidates = h.tpairs[0] + h.tpairs[1];
    st->constraints = h.contacts;
    st->terrain_constraints = h.tcontacts;
    st->groups = h.ngroups;
    st->iterations = iters;
    st->fat_refreshes = h.fat_refreshes;
    st->overflow = overflowed;
    st->colouring_rounds = h.rounds;
    st->ghosts = h.n_total - ctx->n;
    st->boundary_constraints = h.contacts - h.n_int_rows;
    st->phases = h.n_phases;
    st->broadphase_path = h.bp_path;
    if (timed) {
        cudaEventElapsedTime(&st->step_ms, evs[0], evs[1]);
        cudaEventElapsedTime(&st->solve_ms, evs[2], evs[3]);
    }
}

int32_t mgfb_step_n(mgfb_ctx* ctx, float dt, uint32_t iters, uint32_t nsteps, mgfb_step_stats* stats) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (!(dt > 0.0f)) return fail(ctx, MGFB_ERR_INVALID_ARG, "dt must be > 0");
    CU(cudaSetDevice(ctx->device));
    if (ctx->n == 0 || nsteps == 0) { if (stats) std::memset(stats, 0, sizeof(*stats)); return MGFB_OK; }
    if (ctx->cfg.step_order == MGFB_STEP_ORDER_REFERENCE) {
        if (ctx->pipe_inflight) return fail(ctx, MGFB_ERR_STATE, "steps are in flight: mgfb_step_wait first");
        for (uint32_t s = 0; s < nsteps; ++s) TRY(step_reference_order(ctx, dt, iters, stats));
        return MGFB_OK;
    }
    unsigned scale = ctx->tile_exported ? 2 : 1, overflowed = 0;   // a tile cannot regrow mid-run: sized once, generously
    TRY(ensure_step_buffers(ctx, scale));
    TRY(ensure_grid(ctx, scale));
    TRY(ensure_rows(ctx, ctx->contact_cap, false, 4096));
    TRY(read_counters(ctx));
    unsigned done0 = ctx->h_ctr->steps_done, target = done0 + nsteps;
    bool resume_after_integrate = false;
    for (int attempt = 0; attempt < 8; ++attempt) {
        unsigned todo = target - ctx->h_ctr->steps_done;
        for (unsigned s = 0; s < todo; ++s) {
            bool last = (s + 1 == todo);
            if (last) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
            TRY(enqueue_step(ctx, dt, iters, !(s == 0 && resume_after_integrate), last));
            if (last) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        }
        TRY(read_counters(ctx));
        if (ctx->h_ctr->nan_bounds) { clear_sticky(ctx); return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (NaN in body state; bounds.rs:125-127)"); }
        if (ctx->tiled) {
            // a tile cannot rerun part of a step on its own: its neighbours have moved on
            // (an overflow makes this tile's kernels return early, so its neighbours -- and then it -- time out: report the cause)
            if (ctx->h_ctr->overflow & OVF_GHOSTS) return fail(ctx, MGFB_ERR_CAPACITY, "more ghost bodies than the neighbour's ghost capacity (mgfb_tile_export)");
            if (ctx->h_ctr->overflow) {
                ctx->err = "a work list overflowed in a tiled step (bits " + std::to_string(ctx->h_ctr->overflow) + "): lists hold 16 pairs / 16 contacts per body slot; raise ghost_capacity";
                return MGFB_ERR_CAPACITY;
            }
            if (ctx->h_ctr->comm_error & COMM_TILE_TOO_THIN)
                return fail(ctx, MGFB_ERR_TILE, "tile too thin: a body is a ghost on the left neighbour and touches a ghost from the right (or reaches two tiles away); use fewer, wider tiles");
            if (ctx->h_ctr->comm_error) return fail(ctx, MGFB_ERR_TILE, "neighbour tile did not answer within the time limit");
            break;
        }
        if (!ctx->h_ctr->overflow) break;
        // a work list overflowed inside step (steps_done): that step already integrated.
        overflowed |= ctx->h_ctr->overflow;
        if (ctx->h_ctr->overflow & OVF_GROUPS) { clear_sticky(ctx); return fail(ctx, MGFB_ERR_CAPACITY, "more constraint groups than group capacity"); }
        scale *= 2;
        TRY(clear_sticky(ctx));
        TRY(ensure_step_buffers(ctx, scale));
        TRY(ensure_grid(ctx, scale));
        TRY(ensure_rows(ctx, ctx->contact_cap, false, 4096));
        resume_after_integrate = true;
        if (attempt == 7) return fail(ctx, MGFB_ERR_CAPACITY, "work lists still overflow after growing 128x");
    }
    ctx->last_constraints = ctx->h_ctr->contacts;
    ctx->have_step = true;
    if (stats) fill_step_stats(ctx, stats, iters, true, overflowed);
    return MGFB_OK;
}

int32_t mgfb_step(mgfb_ctx* ctx, float dt, uint32_t iters, mgfb_step_stats* stats) { return mgfb_step_n(ctx, dt, iters, 1, stats); }

int32_t mgfb_step_profile(mgfb_ctx* ctx, float dt, uint32_t iters, mgfb_step_stats* stats, mgfb_phase_profile* out) {
    if (!ctx || !out) return MGFB_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    float* phase_ms = out->phase_ms;
    if (ctx->pipe_inflight) return fail(ctx, MGFB_ERR_STATE, "steps are in flight: mgfb_step_wait first");
    CU(cudaSetDevice(ctx->device));
    for (auto& e : ctx->prof_ev) if (!e) CU(cudaEventCreate(&e));
    ctx->prof_on = true;
    int32_t st = mgfb_step_n(ctx, dt, iters, 1, stats);
    ctx->prof_on = false;
    TRY(st);
    for (int k = 0; k < MGFB_PHASE_COUNT; ++k) phase_ms[k] = 0.0f;
    if (ctx->n == 0) return MGFB_OK;
    // phase k runs from mark k to the next mark on the main stream (the terrain half has its own pair on the side stream)
    const int order[] = {MGFB_PHASE_INTEGRATE, MGFB_PHASE_BODY_GRID, MGFB_PHASE_PAIR_SWEEP, MGFB_PHASE_NARROW_BODIES, MGFB_PHASE_COLOURING,
                         MGFB_PHASE_BUILD_ROWS, MGFB_PHASE_SOLVE, MGFB_PHASE_COUNT};
    for (int k = 0; k + 1 < 8; ++k) CU(cudaEventElapsedTime(&phase_ms[order[k]], ctx->prof_ev[order[k]], ctx->prof_ev[order[k + 1]]));
    if (ctx->terrain.present) CU(cudaEventElapsedTime(&phase_ms[MGFB_PHASE_TERRAIN], ctx->prof_ev[9], ctx->prof_ev[10]));
    const Counters& h = *ctx->h_ctr;
    for (int k = 0; k < 4; ++k) out->pairs[k] = h.pairs[k];
    out->terrain_pairs[0] = h.tpairs[0]; out->terrain_pairs[1] = h.tpairs[1];
    out->terrain_contacts = h.tcontacts; out->body_contacts = h.contacts - h.tcontacts;
    return MGFB_OK;
}

int32_t mgfb_step_constraints(mgfb_ctx* ctx, uint32_t capacity, uint32_t* body_a, int32_t* body_b, uint32_t* face, uint32_t* sub,
                              uint32_t* colour, uint32_t* count) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (!ctx->have_step) return fail(ctx, MGFB_ERR_STATE, "no step has been run");
    CU(cudaSetDevice(ctx->device));
    unsigned m = ctx->last_constraints;
    if (count) *count = m;
    if (capacity < m) return fail(ctx, MGFB_ERR_CAPACITY, "output arrays too small");
    if (m == 0) return MGFB_OK;
    std::vector<unsigned> perm(m), hface(m), hsub(m); std::vector<int> ha(m), hb(m), hg(m);
    std::vector<unsigned> hgid;
    if (ctx->tiled) {   // a tile reports GLOBAL body ids
        hgid.resize(body_slots(ctx));
        CU(cudaMemcpyAsync(hgid.data(), ctx->gid.p, (size_t)hgid.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    auto dl = [&](void* dst, const Buf& b) { return cudaMemcpyAsync(dst, b.p, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream); };
    CU(dl(perm.data(), ctx->perm)); CU(dl(ha.data(), ctx->c_a)); CU(dl(hb.data(), ctx->c_b)); CU(dl(hface.data(), ctx->c_face));
    CU(dl(hsub.data(), ctx->c_sub)); CU(dl(hg.data(), ctx->group));
    CU(cudaStreamSynchronize(ctx->stream));
    for (unsigned r = 0; r < m; ++r) {
        unsigned k = perm[r];
        if (body_a) body_a[r] = ctx->tiled ? hgid[ha[k]] : (uint32_t)ha[k];
        if (body_b) body_b[r] = (ctx->tiled && hb[k] >= 0) ? (int32_t)hgid[hb[k]] : hb[k];
        if (face) face[r] = hface[k];
        if (sub) sub[r] = hsub[k];
        if (colour) colour[r] = (uint32_t)hg[k];
    }
    return MGFB_OK;
}

int32_t mgfb_solver_solve(mgfb_ctx* ctx, const mgfb_manifolds* m, float dt, uint32_t iters, uint32_t order, uint32_t* perm_out,
                          float* normal_impulse_out, mgfb_solve_stats* stats) {
    if (!ctx || !m) return MGFB_ERR_INVALID_ARG;
    if (order > MGFB_ORDER_COLOURED) return fail(ctx, MGFB_ERR_INVALID_ARG, "unknown solve order");
    if (!(dt > 0.0f)) return fail(ctx, MGFB_ERR_INVALID_ARG, "dt must be > 0");
    unsigned n = m->n;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n == 0) return MGFB_OK;
    if (!m->obj_a || !m->obj_b || !m->normal || !m->tangent || !m->ncontacts || !m->local_a || !m->local_b)
        return fail(ctx, MGFB_ERR_INVALID_ARG, "null manifold array");
    CU(cudaSetDevice(ctx->device));
    unsigned total_contacts = 0;
    std::vector<float> zc, zf;
    for (unsigned k = 0; k < n; ++k) {
        int a = m->obj_a[k], b = m->obj_b[k];
        if (a >= (int)ctx->n || b >= (int)ctx->n) return fail(ctx, MGFB_ERR_INVALID_ARG, "constraint refers to a body that does not exist");
        if (a >= 0 && a == b) return fail(ctx, MGFB_ERR_INVALID_ARG, "constraint joins a body to itself");
        if ((a < 0 || b < 0) && (!m->static_center || !m->static_friction)) return fail(ctx, MGFB_ERR_INVALID_ARG, "static body without static_center/static_friction");
        if (m->ncontacts[k] < 1 || m->ncontacts[k] > 4) return fail(ctx, MGFB_ERR_INVALID_ARG, "ncontacts must be 1..4");
        total_contacts += m->ncontacts[k];
    }
    const float* sc = m->static_center; const float* sf = m->static_friction;
    if (!sc) { zc.assign((size_t)n * 3, 0.0f); sc = zc.data(); }
    if (!sf) { zf.assign(n, 0.0f); sf = zf.data(); }
    bool as_given = order == MGFB_ORDER_AS_GIVEN;
    TRY(ensure_rows(ctx, n, true, as_given ? n + 1 : 4096));
    auto up = [&](Buf& b, const void* src, size_t bytes) -> int32_t {
        TRY(ensure(ctx, b, bytes));
        CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return MGFB_OK;
    };
    TRY(up(ctx->u_a, m->obj_a, (size_t)n * 4)); TRY(up(ctx->u_b, m->obj_b, (size_t)n * 4));
    TRY(up(ctx->u_sc, sc, (size_t)n * 12)); TRY(up(ctx->u_sf, sf, (size_t)n * 4));
    TRY(up(ctx->u_n, m->normal, (size_t)n * 12)); TRY(up(ctx->u_t, m->tangent, (size_t)n * 24));
    TRY(up(ctx->u_nc, m->ncontacts, (size_t)n * 4));
    TRY(up(ctx->u_la, m->local_a, (size_t)n * 48)); TRY(up(ctx->u_lb, m->local_b, (size_t)n * 48));
    Counters* c = dctr(ctx);
    CU(cudaMemsetAsync(c, 0, offsetof(Counters, overflow), ctx->stream));
    CU(cudaMemsetAsync(ctx->body_scratch.p, 0, (size_t)ctx->cap * 12, ctx->stream));
    CU(cudaMemsetAsync(ctx->group_count.p, 0, (size_t)ctx->group_cap * 4, ctx->stream));
    // a static endpoint may sit in slot a: the ordering only cares about dynamic endpoints,
    // so present (dynamic, other) to it with the dynamic one first.
    std::vector<int> oa(n), ob(n);
    for (unsigned k = 0; k < n; ++k) {
        int a = m->obj_a[k], b = m->obj_b[k];
        if (a < 0) { oa[k] = b; ob[k] = -1; } else { oa[k] = a; ob[k] = b; }
        if (oa[k] < 0) return fail(ctx, MGFB_ERR_INVALID_ARG, "constraint between two static bodies");
    }
    Buf d_oa, d_ob;
    TRY(up(d_oa, oa.data(), (size_t)n * 4)); TRY(up(d_ob, ob.data(), (size_t)n * 4));
    OrderView O = order_view(ctx, d_oa.as<int>(), d_ob.as<int>(), nullptr, nullptr);
    O.n_own = 0xffffffffu;   // caller-built manifolds never involve ghosts
    ManifoldInput M{};
    M.user = true; M.a = ctx->u_a.as<int>(); M.b = ctx->u_b.as<int>();
    M.normal = ctx->u_n.as<float>(); M.tangent = ctx->u_t.as<float>(); M.ncontacts = ctx->u_nc.as<uint32_t>();
    M.ula = ctx->u_la.as<float>(); M.ulb = ctx->u_lb.as<float>(); M.static_center = ctx->u_sc.as<float>(); M.static_friction = ctx->u_sf.as<float>();
    int32_t s = enqueue_order_and_solve(ctx, O, M, nullptr, n, n, as_given, dt, iters, true);
    if (s == MGFB_OK) s = read_counters(ctx);
    release(d_oa); release(d_ob);
    TRY(s);
    if (ctx->h_ctr->overflow) { clear_sticky(ctx); return fail(ctx, MGFB_ERR_CAPACITY, "group capacity exceeded"); }
    std::vector<unsigned> perm(n);
    CU(cudaMemcpyAsync(perm.data(), ctx->perm.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<float> imp(n); std::vector<float4> xtm;
    CU(cudaMemcpyAsync(imp.data(), ctx->r_imp.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (normal_impulse_out && total_contacts > n) {
        xtm.resize((size_t)n * 3);
        CU(cudaMemcpyAsync(xtm.data(), ctx->r_xtm.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (perm_out) std::memcpy(perm_out, perm.data(), (size_t)n * 4);
    if (normal_impulse_out) {
        std::memset(normal_impulse_out, 0, (size_t)n * 16);
        for (unsigned r = 0; r < n; ++r) {
            unsigned k = perm[r];
            normal_impulse_out[4 * k] = imp[r];
            for (unsigned cidx = 1; cidx < m->ncontacts[k]; ++cidx) normal_impulse_out[4 * k + cidx] = xtm[(size_t)r * 3 + cidx - 1].z;
        }
    }
    if (stats) {
        stats->constraints = n; stats->contacts = total_contacts; stats->groups = ctx->h_ctr->ngroups; stats->iterations = iters;
        cudaEventElapsedTime(&stats->solve_ms, ctx->ev[2], ctx->ev[3]);
    }
    return MGFB_OK;
}

int32_t mgfb_step_totals(mgfb_ctx* ctx, uint64_t* steps, uint64_t* constraints, uint64_t* candidate_pairs, uint64_t* groups,
                         uint64_t* kernel_launches, int32_t reset) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    TRY(read_counters(ctx));
    if (steps) *steps = ctx->h_ctr->acc_steps;
    if (constraints) *constraints = ctx->h_ctr->acc_constraints;
    if (candidate_pairs) *candidate_pairs = ctx->h_ctr->acc_pairs;
    if (groups) *groups = ctx->h_ctr->acc_groups;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (reset) {
        CU(cudaMemsetAsync(reinterpret_cast<char*>(dctr(ctx)) + offsetof(Counters, acc_constraints), 0, 32, ctx->stream));
        ctx->launches = 0;
    }
    return MGFB_OK;
}

// ---------------------------------------------------------------- spatial tiling (tile.cuh)
namespace {
struct TileDescRaw {           // what one tile tells the others (fits mgfb_tile_desc)
    uint64_t magic;
    int64_t pid;
    int32_t device; uint32_t n_own, ghost_cap, row_cap;
    void* ptr[11];             // x vel force torque col tight fat gid ridx mbox tile_df (valid inside the exporting process)
    cudaIpcMemHandle_t ipc[11];
};
static_assert(sizeof(TileDescRaw) <= sizeof(mgfb_tile_desc), "mgfb_tile_desc too small");
const uint64_t TILE_MAGIC = 0x6d6766625f74696cULL;
}  // namespace

int32_t mgfb_bodies_set_gid(mgfb_ctx* ctx, uint32_t first, uint32_t n, const uint32_t* gids) {
    if (!ctx || (uint64_t)first + n > ctx->n || !gids) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return MGFB_OK;
    for (uint32_t i = 0; i < n; ++i) if (gids[i] >= (1u << 30)) return fail(ctx, MGFB_ERR_INVALID_ARG, "global ids must be < 2^30");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->gid.as<unsigned>() + first, gids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MGFB_OK;
}

int32_t mgfb_tile_export(mgfb_ctx* ctx, uint32_t ghost_capacity, mgfb_tile_desc* out) {
    // Tiled worlds keep the plain grid + sweep: measured on 2 GPUs, the coherent broadphase saves ~30 us in the broadphase but the
    // dataflow solver -- whose chains cross the tiles in lock-step -- runs ~50 us longer behind it (tools/e2e_diag.py phases).
    if (ctx && !ctx->bp_forced) ctx->bp_on = false;
    if (ctx) ctx->bp_invalidate = true;   // the stored fat boxes / the body set change under the cached broadphase
    if (!ctx || !out) return MGFB_ERR_INVALID_ARG;
    if (ctx->tile_exported) return fail(ctx, MGFB_ERR_STATE, "tile already exported");
    CU(cudaSetDevice(ctx->device));
    ghost_capacity = std::max(ghost_capacity, 1u);
    TRY(grow_bodies(ctx, ctx->n + ghost_capacity));
    ctx->ghost_cap = ghost_capacity;
    TRY(ensure(ctx, ctx->edge_idx, (size_t)std::max(ctx->n, 1u) * 4));
    TRY(ensure(ctx, ctx->edge_mark, (size_t)ctx->n + 4, false, true));   // (zeroed in whole words every step)
    TRY(ensure(ctx, ctx->ridx, (size_t)ghost_capacity * 4, false, true));
    TRY(ensure(ctx, ctx->mbox, sizeof(TileMailbox), false, true));
    // every per-step buffer at its final size now: nothing may be reallocated while neighbours hold pointers
    TRY(ensure_step_buffers(ctx, 2));
    TRY(ensure_grid(ctx, 2));
    TRY(ensure_rows(ctx, ctx->contact_cap, false, 4096));
    ctx->tile_row_cap = ctx->row_cap;
    TRY(ensure(ctx, ctx->edge_slot, (size_t)std::max(ctx->n, 1u) * 4, false, true));
    // [in_a | in_b | link_l | link_r | pad to 32 B | 2 x hand-over self-test scratch (written by the left / right neighbour)]
    TRY(ensure(ctx, ctx->tile_df, (size_t)ctx->tile_row_cap * 2 * sizeof(Inbox) + (((size_t)ghost_capacity * 8 + 31) & ~(size_t)31) + 2 * (size_t)SELFTEST_PAIRS * 32 * sizeof(Inbox), false, true));
    CU(cudaStreamSynchronize(ctx->stream));
    TileDescRaw d; std::memset(&d, 0, sizeof(d));
    d.magic = TILE_MAGIC; d.pid = (int64_t)getpid(); d.device = ctx->device; d.n_own = ctx->n; d.ghost_cap = ghost_capacity; d.row_cap = ctx->tile_row_cap;
    void* ptrs[11] = {ctx->x.p, ctx->vel.p, ctx->force.p, ctx->torque.p, ctx->col.p, ctx->tight.p, ctx->fat.p, ctx->gid.p, ctx->ridx.p, ctx->mbox.p,
                      ctx->tile_df.p};
    for (int k = 0; k < 11; ++k) { d.ptr[k] = ptrs[k]; CU(cudaIpcGetMemHandle(&d.ipc[k], ptrs[k])); }
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out, &d, sizeof(d));
    ctx->tile_exported = true;
    return MGFB_OK;
}

static int32_t tile_open_peer(mgfb_ctx* ctx, const mgfb_tile_desc* desc, TilePeer* P) {
    TileDescRaw d; std::memcpy(&d, desc, sizeof(d));
    if (d.magic != TILE_MAGIC) return fail(ctx, MGFB_ERR_INVALID_ARG, "not a tile descriptor");
    void* p[11];
    if (d.pid == (int64_t)getpid()) {   // a context of this very process: its pointers are ours already
        if (d.device != ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return MGFB_ERR_CUDA; }
        }
        for (int k = 0; k < 11; ++k) p[k] = d.ptr[k];
    } else {                            // another process: map its allocations over NVLink
        for (int k = 0; k < 11; ++k) {
            CU(cudaIpcOpenMemHandle(&p[k], d.ipc[k], cudaIpcMemLazyEnablePeerAccess));
            ctx->ipc_opened.push_back(p[k]);
        }
    }
    P->x = (float4*)p[0]; P->vel = (BodyVel*)p[1]; P->force = (float4*)p[2]; P->torque = (float4*)p[3]; P->col = (Collider*)p[4];
    P->tight = (Box*)p[5]; P->fat = (Box*)p[6]; P->gid = (unsigned*)p[7]; P->ridx = (unsigned*)p[8]; P->mbox = (TileMailbox*)p[9];
    P->n_own = d.n_own; P->ghost_cap = d.ghost_cap;
    // the link tables of the dataflow solve are indexed by a NEIGHBOUR's ghost slots: one common capacity, or a nearly full
    // ghost set would run past them
    if (d.ghost_cap != ctx->ghost_cap) return fail(ctx, MGFB_ERR_INVALID_ARG, "every tile must export the same ghost_capacity");
    P->in_a = (Inbox*)p[10]; P->in_b = P->in_a + d.row_cap;
    P->link_l = reinterpret_cast<unsigned*>(P->in_b + d.row_cap); P->link_r = P->link_l + d.ghost_cap;
    P->selftest = reinterpret_cast<Inbox*>(reinterpret_cast<char*>(P->link_l) + (((size_t)d.ghost_cap * 8 + 31) & ~(size_t)31));
    return MGFB_OK;
}

int32_t mgfb_tile_connect(mgfb_ctx* ctx, uint32_t rank, uint32_t nranks, const mgfb_tile_desc* descs) {
    if (ctx) ctx->bp_invalidate = true;   // the stored fat boxes / the body set change under the cached broadphase
    if (!ctx || !descs || nranks == 0 || rank >= nranks) return fail(ctx, MGFB_ERR_INVALID_ARG, "bad arguments");
    if (!ctx->tile_exported) return fail(ctx, MGFB_ERR_STATE, "mgfb_tile_export first");
    if (ctx->tiled) return fail(ctx, MGFB_ERR_STATE, "tile already connected");
    CU(cudaSetDevice(ctx->device));
    TileLink T{};
    T.has_left = rank > 0; T.has_right = rank + 1 < nranks;
    if (T.has_left) TRY(tile_open_peer(ctx, &descs[rank - 1], &T.left));
    if (T.has_right) TRY(tile_open_peer(ctx, &descs[rank + 1], &T.right));
    T.mine = ctx->mbox.as<TileMailbox>();
    T.edge_idx = ctx->edge_idx.as<unsigned>(); T.edge_mark = ctx->edge_mark.as<unsigned char>(); T.ridx = ctx->ridx.as<unsigned>();
    T.edge_slot = ctx->edge_slot.as<unsigned>();
    T.link_l = reinterpret_cast<unsigned*>(ctx->tile_df.as<Inbox>() + 2 * (size_t)ctx->tile_row_cap); T.link_r = T.link_l + ctx->ghost_cap;
    T.n_own = ctx->n; T.ghost_cap = ctx->ghost_cap; T.step = 0; T.timeout_ns = ctx->tile_timeout_ns;
    ctx->link = T;
    ctx->tile_step = 0;
    // hand-overs cross the tile boundary as 32-byte peer stores: check each neighbour link delivers them whole (selftest.cuh);
    // every tile must end up with the same schedule, so the verdict is also part of what the caller's next collective carries
    if (ctx->cfg.solver_schedule == MGFB_SCHEDULE_DATAFLOW) {
        if (T.has_left) TRY(run_handover_selftest(ctx, T.left.selftest + SELFTEST_PAIRS * 32, true, 256, nullptr, nullptr));   // its "from the right" half
        if (T.has_right) TRY(run_handover_selftest(ctx, T.right.selftest, true, 256, nullptr, nullptr));                        // its "from the left" half
        if (ctx->handover_torn) return fail(ctx, MGFB_ERR_TILE, "a 32-byte peer store was observed torn on this interconnect: create every tile with MGFB_SCHEDULE_PHASES");
    }
    ctx->tiled = true;
    return MGFB_OK;
}

int32_t mgfb_device_view_get(mgfb_ctx* ctx, mgfb_device_view* out) {
    if (!ctx || !out) return MGFB_ERR_INVALID_ARG;
    out->x = ctx->x.p; out->q = ctx->q.p; out->vel = ctx->vel.p; out->collider = ctx->col.p; out->n = ctx->n; out->stream = ctx->stream;
    return MGFB_OK;
}

}  // extern "C"

#include "batch.cuh"
#include "gjk.cuh"
#include "bvh.cuh"
#include "pipeline.cuh"
#include "selftest.cuh"
#include "compound.cuh"
#include "reforder.cuh"
