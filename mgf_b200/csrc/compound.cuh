// compound.cuh -- Compound (src/compound.rs:232-352): an aggregate of Components (spheres / capsules in the compound's own
// frame) with a displacement and a rotation, as batched device queries:
//   Contacts<RHS> for Compound (compound.rs:339-357)    mgfb_compound_contacts_batch       RHS = Moving<Sphere|Capsule|Triangle|Rectangle>
//   Intersects<Compound> for P  (compound.rs:314-337)    mgfb_compound_intersections_batch  P = Ray | Segment
//   Shape::closest_point        (compound.rs:299-311)    mgfb_compound_closest_points
//   BoundedBy<AABB>, <Sphere>   (compound.rs:277-288)    mgfb_compound_bounds
// The reference keeps the components in a BVH<AABB, Component> grown one insert at a time; which leaves a query visits, and in
// which ORDER its callback fires (the last contact is what `last_contact` returns), depends on that tree.  A compound holds
// a handful of components and is built once, so the tree is grown on the host exactly the way bvh.rs grows it (RefTree,
// reftree.cuh) and the device
// threads walk it with the reference's own stack discipline (push left, push right, pop: right child first, bvh.rs:283-310).
// One thread per query; every contact / intersection primitive is the narrowphase's (narrow.cuh).
// Included at the end of capi.cu.
#pragma once

namespace mgfb {

struct CompoundView {
    const CompNode* nodes; const mgfb_shape* comps;   // the RefTree's nodes; comps in insertion order (Compound.shapes)
    int root; unsigned ncomp;
    float4 disp; float4 rot;   // rot = (s, x, y, z)
};
// geom.rs:940-986: the AABB of the eight rotated corners, folded p1.min(p2.min(...p8))
HD void box_rotate(V3 c, V3 r, Q4 rot, V3* oc, V3* orr) {
    V3 vx = qrot(rot, mk3(r.x, 0.0f, 0.0f)), vy = qrot(rot, mk3(0.0f, r.y, 0.0f)), vz = qrot(rot, mk3(0.0f, 0.0f, r.z));
    V3 p[8] = {c + (vx + vy + vz), c + (vx + vy - vz), c + (vx - vy + vz), c + (vx - vy - vz),
               c + (-vx + vy + vz), c + (-vx + vy - vz), c + (-vx - vy + vz), c + (-vx - vy - vz)};
    V3 lo = p[7], hi = p[7];
    for (int i = 6; i >= 0; --i) {
        lo = mk3(fminf(p[i].x, lo.x), fminf(p[i].y, lo.y), fminf(p[i].z, lo.z));
        hi = mk3(fmaxf(p[i].x, hi.x), fmaxf(p[i].y, hi.y), fmaxf(p[i].z, hi.z));
    }
    *orr = (hi - lo) / 2.0f; *oc = (hi + lo) / 2.0f;
}
HD V3 comp_center(const mgfb_shape& s) {   // Shape::center (geom.rs:747, :787)
    return s.kind == MGFB_SPHERE ? mk3(s.p[0], s.p[1], s.p[2]) : mk3(s.p[0], s.p[1], s.p[2]) + mk3(s.p[3], s.p[4], s.p[5]) * 0.5f;
}
HD void comp_translate(mgfb_shape& s, V3 v) { s.p[0] += v.x; s.p[1] += v.y; s.p[2] += v.z; }   // AddAssign (compound.rs:98-105)
HD void comp_rotate(mgfb_shape& s, Q4 r) {            // Volumetric::rotate (compound.rs:56-63, geom.rs:999-1014)
    if (s.kind != MGFB_CAPSULE) return;
    V3 c = comp_center(s), a = mk3(s.p[0], s.p[1], s.p[2]), d = mk3(s.p[3], s.p[4], s.p[5]);
    V3 na = c + qrot(r, a - c), nd = qrot(r, d);
    s.p[0] = na.x; s.p[1] = na.y; s.p[2] = na.z; s.p[3] = nd.x; s.p[4] = nd.y; s.p[5] = nd.z;
}
HD void comp_rotate_about(mgfb_shape& s, Q4 r, V3 p) {   // geom.rs:933-937
    V3 c = comp_center(s);
    comp_translate(s, (p + qrot(r, c - p)) - c);          // set_pos (geom.rs:459-462)
    comp_rotate(s, r);
}
HD V3 sphere_closest(V3 c, float r, V3 to) { V3 d = to - c; return c + d * (len2(d) / (r * r)); }   // geom.rs:751-755 (sic)
HD V3 comp_closest(const mgfb_shape& s, V3 to) {                                                   // compound.rs:124-129
    if (s.kind == MGFB_SPHERE) return sphere_closest(mk3(s.p[0], s.p[1], s.p[2]), s.p[3], to);
    V3 a = mk3(s.p[0], s.p[1], s.p[2]);
    return sphere_closest(seg_closest(a, a + mk3(s.p[3], s.p[4], s.p[5]), to), s.p[6], to);          // geom.rs:791-796
}

// `rhs.contacts(&shape, cb)` for rhs = Moving<Recv>: collision.rs:1368-1383 (the argument moves by -v, the hit is carried back
// by v*t) over Recv.contacts(&Moving<Component>) (compound.rs:163-177).  Returns up to 2 hits.
__device__ __forceinline__ Hits moving_vs_component(const mgfb_shape& rhs, const mgfb_shape& comp) {
    const V3 v = mk3(rhs.v[0], rhs.v[1], rhs.v[2]), nv = -v;
    Hits hs; hs.n = 0;
    const bool cs = comp.kind == MGFB_SPHERE;
    Sph csph; csph.c = mk3(comp.p[0], comp.p[1], comp.p[2]); csph.r = comp.p[3];
    Cap ccap; ccap.a = csph.c; ccap.d = mk3(comp.p[3], comp.p[4], comp.p[5]); ccap.r = comp.p[6];
    Hit h;
    switch (rhs.kind) {
        case MGFB_SPHERE: { Sph s = sh_sphere(rhs); if (cs ? sphere_msphere(s, csph, nv, &h) : sphere_mcapsule(s, ccap, nv, &h)) push(hs, h); break; }
        case MGFB_CAPSULE: { Cap c = sh_capsule(rhs); if (cs ? capsule_msphere(c, csph, nv, &h) : capsule_mcapsule(c, ccap, nv, &h)) push(hs, h); break; }
        case MGFB_TRIANGLE: { Tri t = sh_tri(rhs); if (cs) { if (poly_msphere(t, csph, nv, &h)) push(hs, h); } else hs = poly_mcapsule(t, ccap, nv); break; }
        default: { Rct t = sh_rect(rhs); if (cs) { if (poly_msphere(t, csph, nv, &h)) push(hs, h); } else hs = poly_mcapsule(t, ccap, nv); break; }
    }
    for (int k = 0; k < hs.n && k < 2; ++k) { V3 d = v * hs.h[k].t; hs.h[k].a = hs.h[k].a + d; hs.h[k].b = hs.h[k].b + d; }
    return hs;
}
// BoundedBy<AABB> for Moving<T> (bounds.rs:60-68) over the shapes' own boxes (bounds.rs:137-188)
__device__ __forceinline__ void moving_shape_bounds(const mgfb_shape& s, V3* oc, V3* orr) {
    V3 c, r;
    if (s.kind == MGFB_SPHERE) { c = mk3(s.p[0], s.p[1], s.p[2]); r = mk3(s.p[3], s.p[3], s.p[3]); }
    else if (s.kind == MGFB_CAPSULE) { V3 d = mk3(s.p[3], s.p[4], s.p[5]); float rr = s.p[6] + len(d) * 0.5f; c = mk3(s.p[0], s.p[1], s.p[2]) + d * 0.5f; r = mk3(rr, rr, rr); }
    else if (s.kind == MGFB_TRIANGLE) {
        V3 a = mk3(s.p[0], s.p[1], s.p[2]), b = mk3(s.p[3], s.p[4], s.p[5]), cc = mk3(s.p[6], s.p[7], s.p[8]);
        c = (a + b + cc) / 3.0f;
        r = mk3(fmaxf(fabsf(a.x - c.x), fmaxf(fabsf(b.x - c.x), fabsf(cc.x - c.x))), fmaxf(fabsf(a.y - c.y), fmaxf(fabsf(b.y - c.y), fabsf(cc.y - c.y))),
                fmaxf(fabsf(a.z - c.z), fmaxf(fabsf(b.z - c.z), fabsf(cc.z - c.z))));
    } else {
        Rct q = sh_rect(s);
        V3 p1 = q.c + q.u0 * q.e0, p2 = q.c + q.u1 * q.e1;
        c = q.c;
        r = mk3(fmaxf(fabsf(p1.x - c.x), fabsf(p2.x - c.x)), fmaxf(fabsf(p1.y - c.y), fabsf(p2.y - c.y)), fmaxf(fabsf(p1.z - c.z), fabsf(p2.z - c.z)));
    }
    box_combine(c, r, c + mk3(s.v[0], s.v[1], s.v[2]), r, oc, orr);
}

#define COMP_STACK 64   // bvh.rs:287 SmallVec<[usize; 64]>
// compound.rs:339-357, one thread per rhs
__global__ void __launch_bounds__(128) k_compound_contacts(CompoundView C, const mgfb_shape* __restrict__ rhs, unsigned n, unsigned slots, mgfb_contact* out,
                                                           unsigned* counts) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mgfb_shape R = rhs[i];
    const Q4 rot = mkq(C.rot.x, mk3(C.rot.y, C.rot.z, C.rot.w)), conj = qconj(rot);
    const V3 disp = f4v(C.disp);
    V3 bc, br; moving_shape_bounds(R, &bc, &br);
    box_rotate(bc, br, conj, &bc, &br);
    const V3 bounds_disp = qrot(conj, bc + (-disp)) + disp;
    bc = bc + (bounds_disp - bc);                                   // set_pos
    unsigned cnt = 0;
    int stack[COMP_STACK]; int sp = 0;
    if (C.ncomp) stack[sp++] = C.root;
    while (sp) {
        const CompNode nd = C.nodes[stack[--sp]];
        if (!box_overlaps(bc, br, f4v(nd.c), f4v(nd.r))) continue;  // collision.rs:22-29 (argument first, bvh.rs:296)
        if (nd.left >= 0) { if (sp + 2 <= COMP_STACK) { stack[sp++] = nd.left; stack[sp++] = nd.right; } continue; }
        mgfb_shape shape = C.comps[nd.right];
        comp_rotate_about(shape, rot, zero3());
        comp_translate(shape, disp);
        Hits hs = moving_vs_component(R, shape);
        for (int k = 0; k < hs.n && k < 2; ++k) {
            if (cnt < slots) {
                Hit h = flip(hs.h[k]);                               // callback(-c)
                mgfb_contact o;
                o.a[0] = h.a.x; o.a[1] = h.a.y; o.a[2] = h.a.z; o.b[0] = h.b.x; o.b[1] = h.b.y; o.b[2] = h.b.z;
                o.n[0] = h.n.x; o.n[1] = h.n.y; o.n[2] = h.n.z; o.t = h.t;
                out[(size_t)i * slots + cnt] = o;
            }
            ++cnt;
        }
    }
    counts[i] = cnt;
}
// Particle.intersection(&Component) on a transformed component: the narrowphase's ray casts (collision.rs:249-359)
__device__ __forceinline__ bool particle_vs_component(V3 p, V3 d, float DT, const mgfb_shape& s, float* t, V3* ip) {
    if (s.kind == MGFB_SPHERE) return ray_sphere(p, d, mk3(s.p[0], s.p[1], s.p[2]), s.p[3], t, ip) && !(*t > DT);
    return ray_capsule(p, d, mk3(s.p[0], s.p[1], s.p[2]), mk3(s.p[3], s.p[4], s.p[5]), s.p[6], t, ip) && !(*t > DT);
}
// compound.rs:314-337, one thread per particle
__global__ void __launch_bounds__(128) k_compound_intersections(CompoundView C, unsigned particle_kind, const float* __restrict__ particles, unsigned n,
                                                                mgfb_intersection* out, unsigned* hit) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = particles + 6 * (size_t)i;
    const bool seg = particle_kind == MGFB_SEGMENT;
    V3 pos = mk3(q[0], q[1], q[2]), dir = mk3(q[3], q[4], q[5]);
    if (seg) dir = dir - pos;                                        // geom.rs:848-850
    const float DT = seg ? 1.0f : __builtin_huge_valf();
    const Q4 rot = mkq(C.rot.x, mk3(C.rot.y, C.rot.z, C.rot.w)), conj = qconj(rot);
    const V3 disp = f4v(C.disp);
    const V3 rp = qrot(conj, pos + (-disp)) + disp, rd = qrot(conj, dir);   // the Ray the tree is traced with (its own DT is infinite)
    bool have = false; float best_t = 0.0f; V3 best_p = zero3();
    int stack[COMP_STACK]; int sp = 0;
    if (C.ncomp) stack[sp++] = C.root;
    while (sp) {
        const CompNode nd = C.nodes[stack[--sp]];
        float bt; V3 bp;
        if (!ray_aabb(rp, rd, f4v(nd.c), f4v(nd.r), __builtin_huge_valf(), &bt, &bp)) continue;   // bvh.rs:345-369, collision.rs:202-236
        if (nd.left >= 0) { if (sp + 2 <= COMP_STACK) { stack[sp++] = nd.left; stack[sp++] = nd.right; } continue; }
        if (bt > DT) continue;
        mgfb_shape shape = C.comps[nd.right];
        comp_rotate(shape, rot);                                     // (sic: rotate about its own centre, then + disp)
        comp_translate(shape, disp);
        float t; V3 ip;
        if (particle_vs_component(pos, dir, DT, shape, &t, &ip)) {
            if (have && t > best_t) continue;
            have = true; best_t = t; best_p = ip;
        }
    }
    hit[i] = have ? 1u : 0u;
    mgfb_intersection o; o.p[0] = have ? best_p.x : 0.0f; o.p[1] = have ? best_p.y : 0.0f; o.p[2] = have ? best_p.z : 0.0f; o.t = have ? best_t : 0.0f;
    out[i] = o;
}
// compound.rs:299-311: the components as stored (the reference does not apply disp / rot here), insertion order, strict <
__global__ void __launch_bounds__(128) k_compound_closest(CompoundView C, const float* __restrict__ to, unsigned n, float* out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = mk3(to[3 * i], to[3 * i + 1], to[3 * i + 2]);
    V3 best = zero3(); float best_d = __builtin_huge_valf();
    for (unsigned k = 0; k < C.ncomp; ++k) {
        V3 np = comp_closest(C.comps[k], p);
        float nd = len2(p - np);
        if (nd < best_d) { best = np; best_d = nd; }
    }
    out[3 * i] = best.x; out[3 * i + 1] = best.y; out[3 * i + 2] = best.z;
}
}  // namespace mgfb

struct mgfb_compound {
    mgfb_ctx* ctx = nullptr;
    RefTree tree;                       // host copy of the tree (bounds() reads the root)
    std::vector<mgfb_shape> comps;
    Buf d_nodes, d_comps;
    float disp[3] = {0, 0, 0}, rot[4] = {1, 0, 0, 0};
};

namespace {
CompoundView compound_view(const mgfb_compound* c) {
    CompoundView V;
    V.nodes = c->d_nodes.as<CompNode>(); V.comps = c->d_comps.as<mgfb_shape>(); V.root = c->tree.root; V.ncomp = (unsigned)c->comps.size();
    V.disp = make_float4(c->disp[0], c->disp[1], c->disp[2], 0.0f); V.rot = make_float4(c->rot[0], c->rot[1], c->rot[2], c->rot[3]);
    return V;
}
}  // namespace

extern "C" {
int32_t mgfb_compound_create(mgfb_ctx* ctx, const mgfb_shape* components, uint32_t n, mgfb_compound** out) {
    if (!ctx || !out || (n && !components)) return fail(ctx, MGFB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n > 4096) return fail(ctx, MGFB_ERR_INVALID_ARG, "a compound holds at most 4096 components");
    CU(cudaSetDevice(ctx->device));
    mgfb_compound* c = new mgfb_compound(); c->ctx = ctx;
    for (uint32_t i = 0; i < n; ++i) {
        mgfb_shape s = components[i]; s.v[0] = s.v[1] = s.v[2] = 0.0f;
        V3 bc, br;
        if (s.kind == MGFB_SPHERE && s.p[3] > 0.0f) { bc = mk3(s.p[0], s.p[1], s.p[2]); br = mk3(s.p[3], s.p[3], s.p[3]); }            // bounds.rs:170-177
        else if (s.kind == MGFB_CAPSULE && s.p[6] > 0.0f) { V3 d = mk3(s.p[3], s.p[4], s.p[5]); float rr = s.p[6] + len(d) * 0.5f;    // bounds.rs:179-188
                                                            bc = mk3(s.p[0], s.p[1], s.p[2]) + d * 0.5f; br = mk3(rr, rr, rr); }
        else { delete c; return fail(ctx, MGFB_ERR_INVALID_ARG, "a Component is a Sphere or a Capsule with radius > 0 (compound.rs:33-37, geom.rs:300,328)"); }
        c->comps.push_back(s);
        if (c->tree.insert(bc, br, (int)i) < 0) { delete c; return fail(ctx, MGFB_ERR_NAN_BOUNDS, "AABB::combine: r >= 0 violated (bounds.rs:125-127)"); }
    }
    int32_t st = MGFB_OK;
    if (n) {
        st = ensure(ctx, c->d_nodes, c->tree.N.size() * sizeof(CompNode));
        if (st == MGFB_OK) st = ensure(ctx, c->d_comps, c->comps.size() * sizeof(mgfb_shape));
        if (st == MGFB_OK && (cudaMemcpyAsync(c->d_nodes.p, c->tree.N.data(), c->tree.N.size() * sizeof(CompNode), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                              cudaMemcpyAsync(c->d_comps.p, c->comps.data(), c->comps.size() * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                              cudaStreamSynchronize(ctx->stream) != cudaSuccess)) st = fail(ctx, MGFB_ERR_CUDA, "uploading the compound failed");
    }
    if (st != MGFB_OK) { release(c->d_nodes); release(c->d_comps); delete c; return st; }
    *out = c;
    return MGFB_OK;
}
void mgfb_compound_destroy(mgfb_compound* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    release(c->d_nodes); release(c->d_comps);
    delete c;
}
int32_t mgfb_compound_set_transform(mgfb_compound* c, const float disp[3], const float rot[4]) {
    if (!c || !disp || !rot) return MGFB_ERR_INVALID_ARG;
    std::memcpy(c->disp, disp, 12); std::memcpy(c->rot, rot, 16);
    return MGFB_OK;
}
int32_t mgfb_compound_bounds(const mgfb_compound* c, float aabb[6], float sphere[4]) {
    if (!c) return MGFB_ERR_INVALID_ARG;
    if (c->comps.empty()) return fail(c->ctx, MGFB_ERR_STATE, "BVH is empty, there is no root node (bvh.rs:263)");
    const CompNode& rt = c->tree.N[c->tree.root];
    const V3 disp = mk3(c->disp[0], c->disp[1], c->disp[2]);
    if (aabb) {   // compound.rs:277-281: root.rotate(rot) + disp
        V3 oc, orr; box_rotate(f4v(rt.c), f4v(rt.r), mkq(c->rot[0], mk3(c->rot[1], c->rot[2], c->rot[3])), &oc, &orr);
        oc = oc + disp;
        aabb[0] = oc.x; aabb[1] = oc.y; aabb[2] = oc.z; aabb[3] = orr.x; aabb[4] = orr.y; aabb[5] = orr.z;
    }
    if (sphere) {   // compound.rs:283-288 over bounds.rs:291-298
        V3 sc = f4v(rt.c) + disp;
        sphere[0] = sc.x; sphere[1] = sc.y; sphere[2] = sc.z; sphere[3] = len(f4v(rt.r));
    }
    return MGFB_OK;
}
int32_t mgfb_compound_closest_points(mgfb_compound* c, const float* to, uint32_t n, float* out) {
    if (!c || (n && (!to || !out))) return MGFB_ERR_INVALID_ARG;
    if (n == 0) return MGFB_OK;
    mgfb_ctx* ctx = c->ctx;
    CU(cudaSetDevice(ctx->device));
    Buf dt, dout; int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        TRY(ensure(ctx, dt, (size_t)n * 12)); TRY(ensure(ctx, dout, (size_t)n * 12));
        CU(cudaMemcpyAsync(dt.p, to, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        k_compound_closest<<<(n + 127) / 128, 128, 0, ctx->stream>>>(compound_view(c), dt.as<float>(), n, dout.as<float>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 1;
        return MGFB_OK;
    };
    s = body(); release(dt); release(dout);
    return s;
}
int32_t mgfb_compound_intersections_batch(mgfb_compound* c, uint32_t particle_kind, const float* particles, uint32_t n, mgfb_intersection* out, uint32_t* hit) {
    if (!c || (n && (!particles || !out || !hit))) return MGFB_ERR_INVALID_ARG;
    mgfb_ctx* ctx = c->ctx;
    if (particle_kind > MGFB_SEGMENT) return fail(ctx, MGFB_ERR_INVALID_ARG, "particle kind must be MGFB_RAY or MGFB_SEGMENT");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    Buf dp, dout, dh; int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        TRY(ensure(ctx, dp, (size_t)n * 24)); TRY(ensure(ctx, dout, (size_t)n * sizeof(mgfb_intersection))); TRY(ensure(ctx, dh, (size_t)n * 4));
        CU(cudaMemcpyAsync(dp.p, particles, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
        k_compound_intersections<<<(n + 127) / 128, 128, 0, ctx->stream>>>(compound_view(c), particle_kind, dp.as<float>(), n, dout.as<mgfb_intersection>(), dh.as<unsigned>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(mgfb_intersection), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(hit, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 1;
        return MGFB_OK;
    };
    s = body(); release(dp); release(dout); release(dh);
    return s;
}
int32_t mgfb_compound_contacts_batch(mgfb_compound* c, const mgfb_shape* rhs, uint32_t n, uint32_t slots, mgfb_contact* out, uint32_t* counts) {
    if (!c || (n && (!rhs || !out || !counts)) || slots == 0) return MGFB_ERR_INVALID_ARG;
    mgfb_ctx* ctx = c->ctx;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t k = rhs[i].kind;
        if (k != MGFB_SPHERE && k != MGFB_CAPSULE && k != MGFB_TRIANGLE && k != MGFB_RECTANGLE)
            return fail(ctx, MGFB_ERR_INVALID_ARG, "RHS must be a Moving<Sphere|Capsule|Triangle|Rectangle> (Contacts<Component> + BoundedBy<AABB>, compound.rs:339-342)");
        if ((k == MGFB_SPHERE && !(rhs[i].p[3] > 0.0f)) || (k == MGFB_CAPSULE && !(rhs[i].p[6] > 0.0f))) return fail(ctx, MGFB_ERR_INVALID_ARG, "radius must be > 0 (geom.rs:300,328)");
    }
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    Buf dr, dout, dc; int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        TRY(ensure(ctx, dr, (size_t)n * sizeof(mgfb_shape))); TRY(ensure(ctx, dout, (size_t)n * slots * sizeof(mgfb_contact))); TRY(ensure(ctx, dc, (size_t)n * 4));
        CU(cudaMemcpyAsync(dr.p, rhs, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(dout.p, 0, (size_t)n * slots * sizeof(mgfb_contact), ctx->stream));
        k_compound_contacts<<<(n + 127) / 128, 128, 0, ctx->stream>>>(compound_view(c), dr.as<mgfb_shape>(), n, slots, dout.as<mgfb_contact>(), dc.as<unsigned>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t)n * slots * sizeof(mgfb_contact), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(counts, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 1;
        return MGFB_OK;
    };
    s = body(); release(dr); release(dout); release(dc);
    return s;
}
}  // extern "C"
