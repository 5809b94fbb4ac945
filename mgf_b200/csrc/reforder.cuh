// reforder.cuh -- World::step (mgf_demo/world.rs:227-294) in the REFERENCE's own constraint order (mgfb_config.step_order =
// MGFB_STEP_ORDER_REFERENCE).  Gauss-Seidel results depend on the order constraints are added to the Solver; the demo world
// adds them body by body, for body i first its terrain contacts in the order the mesh BVH's query calls back, then its body
// pairs (j < i) in the order the body BVH's query calls back (world.rs:240-291).  Both trees are grown incrementally (insert,
// and remove + insert whenever a body leaves its fat box, world.rs:235-238), so the order is a function of the trees' whole
// history.  Here that history is replayed on the host with RefTree (reftree.cuh = bvh.rs), interleaved per body exactly like
// the reference (body i's leaf is refreshed right before body i queries; later bodies still hang where they were), and the
// resulting ORDERED candidate list goes to the device: ordered narrowphase -> contacts compacted in order -> level-scheduled
// solve of the list as given (MGFB_ORDER_AS_GIVEN: bit-identical to the sequential sweep).  The state after every step is then
// bit-identical to the reference's own World::step -- no order to export or replay.  Integration, swept / fat boxes, every
// contact and every impulse are still computed by the device kernels; the host only walks the trees.  Meant for worlds of up
// to ~10^5 bodies (the walk is sequential, like the reference's).  Included at the end of capi.cu.
#pragma once

namespace mgfb {
// One candidate of the ordered list: body i against body j (>= 0) or terrain face -1 - j.  Up to two contacts each, written to
// fixed slots; cnt[c] says how many.
__global__ void __launch_bounds__(MGFB_THREADS) k_narrow_ordered(const Collider* __restrict__ col, const int2* __restrict__ cand, unsigned ncand, TerrainView T,
                                                                float4* la, float4* lb, float4* nt, unsigned* cnt) {
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const int2 ij = cand[c];
    const Collider A = col[ij.x];
    Hit hs[2]; V3 pa[2], pb[2]; int nh = 0;
    if (ij.y >= 0) {
        const Collider Bc = col[ij.y];
        const int ki = col_kind(A), kj = col_kind(Bc);
        bool ok;
        if (ki == 0 && kj == 0) ok = body_pair_contact<0, 0>(A, Bc, &hs[0], &pa[0], &pb[0]);
        else if (ki == 0) ok = body_pair_contact<0, 1>(A, Bc, &hs[0], &pa[0], &pb[0]);
        else if (kj == 0) ok = body_pair_contact<1, 0>(A, Bc, &hs[0], &pa[0], &pb[0]);
        else ok = body_pair_contact<1, 1>(A, Bc, &hs[0], &pa[0], &pb[0]);
        nh = ok ? 1 : 0;
        if (ok) hs[0].n = (zero3() + hs[0].n) / 1.0f;   // ContactPruner with one contact -> Manifold::from(pruner) (manifold.rs:135-140)
    } else {
        const V3 mx = f4v(T.x);
        const uint4 fc = T.faces[-1 - ij.y];
        Tri tri; tri.a = f4v(T.verts[fc.x]) + mx; tri.b = f4v(T.verts[fc.y]) + mx; tri.c = f4v(T.verts[fc.z]) + mx;
        nh = col_kind(A) == 0 ? body_tri_contacts<0>(A, tri, mx, hs, pa, pb) : body_tri_contacts<1>(A, tri, mx, hs, pa, pb);
        if (nh > 2) nh = 2;
    }
    cnt[c] = (unsigned)nh;
    for (int k = 0; k < nh; ++k) {
        la[2 * c + k] = v4(pa[k], hs[k].n.x); lb[2 * c + k] = v4(pb[k], hs[k].n.y); nt[2 * c + k] = make_float4(hs[k].n.z, hs[k].t, 0.0f, 0.0f);
    }
}
__global__ void __launch_bounds__(MGFB_THREADS) k_compact_ordered(const int2* __restrict__ cand, unsigned ncand, const unsigned* __restrict__ offs, const float4* __restrict__ la,
                                                                 const float4* __restrict__ lb, const float4* __restrict__ nt, ContactList L, Counters* ctr) {
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const unsigned o = offs[c], n = offs[c + 1] - o;
    const int2 ij = cand[c];
    for (unsigned k = 0; k < n; ++k) {
        L.a[o + k] = ij.x; L.b[o + k] = ij.y >= 0 ? ij.y : -1; L.face[o + k] = ij.y >= 0 ? 0u : (unsigned)(-1 - ij.y); L.sub[o + k] = ij.y >= 0 ? 0u : k;
        L.la[o + k] = la[2 * c + k]; L.lb[o + k] = lb[2 * c + k]; L.nt[o + k] = nt[2 * c + k];
    }
    if (c == ncand - 1) { ctr->contacts = offs[ncand]; }
    if (ij.y < 0 && n) atomicAdd(&ctr->tcontacts, n);
}
}  // namespace mgfb

namespace {
void reforder_free(mgfb_ctx* ctx) {
    if (!ctx->reforder) return;
    RefOrderState* S = ctx->reforder;
    release(S->d_cand); release(S->d_la); release(S->d_lb); release(S->d_nt); release(S->d_cnt); release(S->d_offs);
    delete S; ctx->reforder = nullptr;
}
// (re)build the body tree: BVH::new, then the inserts add_body makes, in body order (world.rs:178-184)
int32_t reforder_rebuild_bodies(mgfb_ctx* ctx) {
    RefOrderState* S = ctx->reforder;
    const unsigned n = ctx->n;
    S->h_fat.resize(n);
    if (n) { CU(cudaMemcpyAsync(S->h_fat.data(), ctx->fat.p, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, ctx->stream)); CU(cudaStreamSynchronize(ctx->stream)); }
    S->body_tree.clear(); S->leaf.assign(n, -1); S->fat = S->h_fat;
    for (unsigned i = 0; i < n; ++i) {
        S->leaf[i] = S->body_tree.insert(f4v(S->fat[i].c), f4v(S->fat[i].r), (int)i);
        if (S->leaf[i] < 0) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)");
    }
    return MGFB_OK;
}
// Mesh::push_face inserts bounds(triangle) face by face (mesh.rs:60-71, bounds.rs:137-152)
int32_t reforder_build_mesh(mgfb_ctx* ctx) {
    RefOrderState* S = ctx->reforder;
    S->mesh_tree.clear(); S->mesh_built = true;
    const TerrainData& t = ctx->terrain;
    if (!t.present) return MGFB_OK;
    std::vector<Box> fb(t.nfaces);
    CU(cudaMemcpyAsync(fb.data(), t.boxes.p, (size_t)t.nfaces * sizeof(Box), cudaMemcpyDeviceToHost, ctx->stream));   // k_face_boxes: the same expression
    CU(cudaStreamSynchronize(ctx->stream));
    for (unsigned f = 0; f < t.nfaces; ++f)
        if (S->mesh_tree.insert(f4v(fb[f].c), f4v(fb[f].r), (int)f) < 0) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)");
    return MGFB_OK;
}

int32_t step_reference_order(mgfb_ctx* ctx, float dt, unsigned iters, mgfb_step_stats* stats) {
    if (ctx->tiled || ctx->tile_exported) return fail(ctx, MGFB_ERR_STATE, "reference order is for a world on one GPU");
    if (!ctx->reforder) { ctx->reforder = new RefOrderState(); TRY(reforder_rebuild_bodies(ctx)); }
    RefOrderState* S = ctx->reforder;
    if (S->leaf.size() != ctx->n) TRY(reforder_rebuild_bodies(ctx));   // (bodies are normally inserted as they are added, see mgfb_bodies_add)
    if (!S->mesh_built) TRY(reforder_build_mesh(ctx));
    const unsigned n = ctx->n;
    Counters* c = dctr(ctx);
    BodyArrays B = body_arrays(ctx);
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaMemsetAsync(c, 0, offsetof(Counters, overflow), ctx->stream));
    CU(cudaMemsetAsync(ctx->body_scratch.p, 0, (size_t)ctx->cap * 12, ctx->stream));
    k_integrate<true, true, true><<<grid_for(ctx, n), MGFB_THREADS, 0, ctx->stream>>>(B, n, dt, ctx->cfg.fat_margin, c);
    S->h_tight.resize(n); S->h_fat.resize(n);
    CU(cudaMemcpyAsync(S->h_tight.data(), ctx->tight.p, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(S->h_fat.data(), ctx->fat.p, (size_t)n * sizeof(Box), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(read_counters(ctx));
    if (ctx->h_ctr->nan_bounds) { clear_sticky(ctx); return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (NaN in body state; bounds.rs:125-127)"); }
    const unsigned refreshes = ctx->h_ctr->fat_refreshes;
    // ---- the reference's loop over bodies (world.rs:233-291), trees only
    S->cand.clear();
    unsigned n_terrain = 0, n_body = 0;
    const V3 mx = mk3(ctx->terrain.x[0], ctx->terrain.x[1], ctx->terrain.x[2]);
    for (unsigned i = 0; i < n; ++i) {
        const Box& tb = S->h_tight[i]; const Box& fb = S->h_fat[i];
        if (std::memcmp(&fb, &S->fat[i], sizeof(Box)) != 0) {          // the device replaced the stored box: !contains (world.rs:235-238)
            if (!S->body_tree.remove(S->leaf[i])) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)");
            S->leaf[i] = S->body_tree.insert(f4v(fb.c), f4v(fb.r), (int)i);
            if (S->leaf[i] < 0) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)");
            S->fat[i] = fb;
        }
        const V3 tc = f4v(tb.c), tr = f4v(tb.r);
        if (ctx->terrain.present) S->mesh_tree.query(tc + (-mx), tr, [&](int f) { S->cand.push_back(make_int2((int)i, -1 - f)); ++n_terrain; });   // mesh.rs:121
        if (i == 0) continue;                                                                                                                    // world.rs:256
        S->body_tree.query(tc, tr, [&](int j) { if ((unsigned)j < i) { S->cand.push_back(make_int2((int)i, j)); ++n_body; } });               // world.rs:261-268
    }
    const unsigned ncand = (unsigned)S->cand.size();
    // ---- ordered narrowphase, contacts compacted in candidate order
    const unsigned mcap = std::max(2u * ncand, 1024u);
    if (mcap > ctx->contact_cap) {
        TRY(ensure(ctx, ctx->c_a, (size_t)mcap * 4)); TRY(ensure(ctx, ctx->c_b, (size_t)mcap * 4)); TRY(ensure(ctx, ctx->c_face, (size_t)mcap * 4));
        TRY(ensure(ctx, ctx->c_sub, (size_t)mcap * 4)); TRY(ensure(ctx, ctx->c_la, (size_t)mcap * 16)); TRY(ensure(ctx, ctx->c_lb, (size_t)mcap * 16));
        TRY(ensure(ctx, ctx->c_nt, (size_t)mcap * 16));
        ctx->contact_cap = mcap;
    }
    TRY(ensure_rows(ctx, mcap, false, mcap + 1));
    ContactList L = contact_list(ctx);
    if (ncand) {
        TRY(ensure(ctx, S->d_cand, (size_t)ncand * 8)); TRY(ensure(ctx, S->d_la, (size_t)ncand * 32)); TRY(ensure(ctx, S->d_lb, (size_t)ncand * 32));
        TRY(ensure(ctx, S->d_nt, (size_t)ncand * 32)); TRY(ensure(ctx, S->d_cnt, (size_t)ncand * 4)); TRY(ensure(ctx, S->d_offs, ((size_t)ncand + 1) * 4));
        CU(cudaMemcpyAsync(S->d_cand.p, S->cand.data(), (size_t)ncand * 8, cudaMemcpyHostToDevice, ctx->stream));
        const unsigned nb = (ncand + MGFB_THREADS - 1) / MGFB_THREADS;
        TerrainView T{}; if (ctx->terrain.present) T = terrain_view(ctx);
        k_narrow_ordered<<<nb, MGFB_THREADS, 0, ctx->stream>>>(ctx->col.as<Collider>(), S->d_cand.as<int2>(), ncand, T, S->d_la.as<float4>(), S->d_lb.as<float4>(),
                                                               S->d_nt.as<float4>(), S->d_cnt.as<unsigned>());
        TRY(scan_u32_lb(ctx, S->d_cnt.as<unsigned>(), S->d_offs.as<unsigned>(), ncand, nullptr));
        k_compact_ordered<<<nb, MGFB_THREADS, 0, ctx->stream>>>(S->d_cand.as<int2>(), ncand, S->d_offs.as<unsigned>(), S->d_la.as<float4>(), S->d_lb.as<float4>(),
                                                                S->d_nt.as<float4>(), L, c);
        CU(cudaGetLastError());
        ctx->launches += 4;
    }
    CU(cudaMemsetAsync(ctx->group_count.p, 0, (size_t)ctx->group_cap * 4, ctx->stream));
    OrderView O = order_view(ctx, L.a, L.b, L.face, L.sub);
    ManifoldInput M{};
    M.a = L.a; M.b = L.b; M.la = L.la; M.lb = L.lb; M.nt = L.nt; M.user = false;
    M.terrain_center = make_float4(ctx->terrain.x[0], ctx->terrain.x[1], ctx->terrain.x[2], 0.0f);
    TRY(enqueue_order_and_solve(ctx, O, M, &c->contacts, 0, mcap, true, dt, iters, true));
    k_step_done<<<1, 64, 0, ctx->stream>>>(c, nullptr);
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->launches += 3;
    TRY(read_counters(ctx));
    if (ctx->h_ctr->overflow) { clear_sticky(ctx); return fail(ctx, MGFB_ERR_CAPACITY, "more constraint groups than group capacity"); }
    ctx->last_constraints = ctx->h_ctr->contacts;
    ctx->have_step = true;
    if (stats) {
        fill_step_stats(ctx, stats, iters, true, 0);
        stats->candidate_pairs = n_body; stats->terrain_candidates = n_terrain; stats->fat_refreshes = refreshes;
    }
    return MGFB_OK;
}
}  // namespace
