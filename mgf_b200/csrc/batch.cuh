// batch.cuh -- mgfb_contacts_batch: the narrowphase specialisations exposed one pair kind at a
// time, so that every `Contacts` impl can be checked against the reference's own unit-test
// vectors on the device.  Included at the end of capi.cu.
#pragma once

namespace {

__device__ __forceinline__ Sph sh_sphere(const mgfb_shape& s) { Sph r; r.c = mk3(s.p[0], s.p[1], s.p[2]); r.r = s.p[3]; return r; }
__device__ __forceinline__ Cap sh_capsule(const mgfb_shape& s) { Cap c; c.a = mk3(s.p[0], s.p[1], s.p[2]); c.d = mk3(s.p[3], s.p[4], s.p[5]); c.r = s.p[6]; return c; }
__device__ __forceinline__ Tri sh_tri(const mgfb_shape& s) { Tri t; t.a = mk3(s.p[0], s.p[1], s.p[2]); t.b = mk3(s.p[3], s.p[4], s.p[5]); t.c = mk3(s.p[6], s.p[7], s.p[8]); return t; }
__device__ __forceinline__ Rct sh_rect(const mgfb_shape& s) {
    Rct r; r.c = mk3(s.p[0], s.p[1], s.p[2]); r.u0 = mk3(s.p[3], s.p[4], s.p[5]); r.u1 = mk3(s.p[6], s.p[7], s.p[8]); r.e0 = s.p[9]; r.e1 = s.p[10];
    return r;
}
__device__ __forceinline__ Pln sh_plane(const mgfb_shape& s) { Pln p; p.n = mk3(s.p[0], s.p[1], s.p[2]); p.d = s.p[3]; return p; }
__device__ __forceinline__ V3 sh_vel(const mgfb_shape& s) { return mk3(s.v[0], s.v[1], s.v[2]); }
__device__ __forceinline__ Collider sh_collider(const mgfb_shape& s) {
    Collider k;
    if (s.kind == MGFB_SPHERE) { k.p0 = make_float4(s.p[0], s.p[1], s.p[2], s.p[3]); k.p1 = make_float4(0, 0, 0, ibits(0)); }
    else { k.p0 = make_float4(s.p[0], s.p[1], s.p[2], s.p[6]); k.p1 = make_float4(s.p[3], s.p[4], s.p[5], ibits(1)); }
    k.v = make_float4(s.v[0], s.v[1], s.v[2], 0.0f);
    return k;
}
__device__ __forceinline__ void put_contact(mgfb_contact* o, const Hit& h) {
    o->a[0] = h.a.x; o->a[1] = h.a.y; o->a[2] = h.a.z; o->b[0] = h.b.x; o->b[1] = h.b.y; o->b[2] = h.b.z;
    o->n[0] = h.n.x; o->n[1] = h.n.y; o->n[2] = h.n.z; o->t = h.t;
}
__device__ __forceinline__ void put_local(mgfb_local_contact* o, V3 la, V3 lb, const Hit& h) {
    o->local_a[0] = la.x; o->local_a[1] = la.y; o->local_a[2] = la.z; o->local_b[0] = lb.x; o->local_b[1] = lb.y; o->local_b[2] = lb.z;
    put_contact(&o->global, h);
}

template <int KIND>
__global__ void __launch_bounds__(128) k_contacts_batch(const mgfb_shape* __restrict__ recv, const mgfb_shape* __restrict__ arg, unsigned n,
                                                        mgfb_contact* out, mgfb_local_contact* out_local, unsigned* counts) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mgfb_shape R = recv[i], A = arg[i];
    Hits hs; hs.n = 0;
    V3 la[2], lb[2];
    Hit h;
    V3 v = sh_vel(A);
    if (KIND == MGFB_SPHERE_X_MSPHERE) { if (sphere_msphere(sh_sphere(R), sh_sphere(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_CAPSULE_X_MSPHERE) { if (capsule_msphere(sh_capsule(R), sh_sphere(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_SPHERE_X_MCAPSULE) { if (sphere_mcapsule(sh_sphere(R), sh_capsule(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_CAPSULE_X_MCAPSULE) { if (capsule_mcapsule(sh_capsule(R), sh_capsule(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_PLANE_X_MSPHERE) { if (plane_msphere(sh_plane(R), sh_sphere(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_PLANE_X_MCAPSULE) { if (plane_mcapsule(sh_plane(R), sh_capsule(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_TRI_X_MSPHERE) { if (poly_msphere(sh_tri(R), sh_sphere(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_TRI_X_MCAPSULE) { hs = poly_mcapsule(sh_tri(R), sh_capsule(A), v); }
    else if (KIND == MGFB_RECT_X_MSPHERE) { if (poly_msphere(sh_rect(R), sh_sphere(A), v, &h)) push(hs, h); }
    else if (KIND == MGFB_RECT_X_MCAPSULE) { hs = poly_mcapsule(sh_rect(R), sh_capsule(A), v); }
    else if (KIND == MGFB_MCOMP_X_MCOMP) {
        Collider ca = sh_collider(R), cb = sh_collider(A);
        bool ok;
        int ki = R.kind == MGFB_CAPSULE, kj = A.kind == MGFB_CAPSULE;
        if (!ki && !kj) ok = body_pair_contact<0, 0>(ca, cb, &h, &la[0], &lb[0]);
        else if (ki && !kj) ok = body_pair_contact<1, 0>(ca, cb, &h, &la[0], &lb[0]);
        else if (!ki && kj) ok = body_pair_contact<0, 1>(ca, cb, &h, &la[0], &lb[0]);
        else ok = body_pair_contact<1, 1>(ca, cb, &h, &la[0], &lb[0]);
        if (ok) push(hs, h);
    } else if (KIND == MGFB_MCOMP_X_TRI) {
        Collider ca = sh_collider(R);
        Hit o2[2];
        int nh = R.kind == MGFB_CAPSULE ? body_tri_contacts<1>(ca, sh_tri(A), v, o2, la, lb) : body_tri_contacts<0>(ca, sh_tri(A), v, o2, la, lb);
        hs.n = nh; for (int k = 0; k < nh; ++k) hs.h[k] = o2[k];
    }
    counts[i] = (unsigned)hs.n;
    for (int k = 0; k < hs.n && k < 2; ++k) {
        put_contact(&out[2 * i + k], hs.h[k]);
        if (out_local && (KIND == MGFB_MCOMP_X_MCOMP || KIND == MGFB_MCOMP_X_TRI)) put_local(&out_local[2 * i + k], la[k], lb[k], hs.h[k]);
    }
}

template <int KIND>
void launch_batch(mgfb_ctx* ctx, const mgfb_shape* r, const mgfb_shape* a, unsigned n, mgfb_contact* o, mgfb_local_contact* ol, unsigned* c) {
    k_contacts_batch<KIND><<<(n + 127) / 128, 128, 0, ctx->stream>>>(r, a, n, o, ol, c);
}

}  // namespace

extern "C" int32_t mgfb_contacts_batch(mgfb_ctx* ctx, uint32_t pair_kind, const mgfb_shape* recv, const mgfb_shape* arg, uint32_t n,
                                       mgfb_contact* out, mgfb_local_contact* out_local, uint32_t* counts) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (pair_kind >= MGFB_PAIR_KIND_COUNT) return fail(ctx, MGFB_ERR_INVALID_ARG, "unknown pair kind");
    if (n == 0) return MGFB_OK;
    if (!recv || !arg || !out || !counts) return fail(ctx, MGFB_ERR_INVALID_ARG, "null array");
    static const int want_recv[MGFB_PAIR_KIND_COUNT] = {MGFB_SPHERE, MGFB_CAPSULE, MGFB_SPHERE, MGFB_CAPSULE, MGFB_PLANE, MGFB_PLANE,
                                                       MGFB_TRIANGLE, MGFB_TRIANGLE, MGFB_RECTANGLE, MGFB_RECTANGLE, -1, -1};
    static const int want_arg[MGFB_PAIR_KIND_COUNT] = {MGFB_SPHERE, MGFB_SPHERE, MGFB_CAPSULE, MGFB_CAPSULE, MGFB_SPHERE, MGFB_CAPSULE,
                                                      MGFB_SPHERE, MGFB_CAPSULE, MGFB_SPHERE, MGFB_CAPSULE, -1, MGFB_TRIANGLE};
    for (uint32_t i = 0; i < n; ++i) {
        int rk = (int)recv[i].kind, ak = (int)arg[i].kind;
        bool ok = (want_recv[pair_kind] < 0 ? (rk == MGFB_SPHERE || rk == MGFB_CAPSULE) : rk == want_recv[pair_kind]) &&
                  (want_arg[pair_kind] < 0 ? (ak == MGFB_SPHERE || ak == MGFB_CAPSULE) : ak == want_arg[pair_kind]);
        if (!ok) return fail(ctx, MGFB_ERR_INVALID_ARG, "shape kind does not match pair_kind (batches are homogeneous)");
    }
    CU(cudaSetDevice(ctx->device));
    Buf dr, da, dout, dl, dc;
    int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        TRY(ensure(ctx, dr, (size_t)n * sizeof(mgfb_shape))); TRY(ensure(ctx, da, (size_t)n * sizeof(mgfb_shape)));
        TRY(ensure(ctx, dout, (size_t)n * 2 * sizeof(mgfb_contact))); TRY(ensure(ctx, dc, (size_t)n * 4));
        if (out_local) TRY(ensure(ctx, dl, (size_t)n * 2 * sizeof(mgfb_local_contact)));
        CU(cudaMemcpyAsync(dr.p, recv, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(da.p, arg, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(dout.p, 0, (size_t)n * 2 * sizeof(mgfb_contact), ctx->stream));
        if (out_local) CU(cudaMemsetAsync(dl.p, 0, (size_t)n * 2 * sizeof(mgfb_local_contact), ctx->stream));
        const mgfb_shape* r = dr.as<mgfb_shape>(); const mgfb_shape* a = da.as<mgfb_shape>();
        mgfb_contact* o = dout.as<mgfb_contact>(); mgfb_local_contact* ol = out_local ? dl.as<mgfb_local_contact>() : nullptr;
        unsigned* c = dc.as<unsigned>();
        switch (pair_kind) {
            case 0: launch_batch<0>(ctx, r, a, n, o, ol, c); break;
            case 1: launch_batch<1>(ctx, r, a, n, o, ol, c); break;
            case 2: launch_batch<2>(ctx, r, a, n, o, ol, c); break;
            case 3: launch_batch<3>(ctx, r, a, n, o, ol, c); break;
            case 4: launch_batch<4>(ctx, r, a, n, o, ol, c); break;
            case 5: launch_batch<5>(ctx, r, a, n, o, ol, c); break;
            case 6: launch_batch<6>(ctx, r, a, n, o, ol, c); break;
            case 7: launch_batch<7>(ctx, r, a, n, o, ol, c); break;
            case 8: launch_batch<8>(ctx, r, a, n, o, ol, c); break;
            case 9: launch_batch<9>(ctx, r, a, n, o, ol, c); break;
            case 10: launch_batch<10>(ctx, r, a, n, o, ol, c); break;
            default: launch_batch<11>(ctx, r, a, n, o, ol, c); break;
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t)n * 2 * sizeof(mgfb_contact), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_local) CU(cudaMemcpyAsync(out_local, dl.p, (size_t)n * 2 * sizeof(mgfb_local_contact), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(counts, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return MGFB_OK;
    };
    s = body();
    release(dr); release(da); release(dout); release(dl); release(dc);
    return s;
}

// ---------------------------------------------------------------- ray casts (Intersects<RHS> for Particle)
namespace {
// collision.rs:163-373.  One thread per (particle, shape) query; Ray: DT = inf, Segment: DT = 1 and dir = b - a.
__global__ void __launch_bounds__(128) k_intersections_batch(unsigned particle_kind, const float* __restrict__ particles,
                                                             const mgfb_shape* __restrict__ shapes, unsigned n, mgfb_intersection* out, unsigned* hit) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = particles + 6 * (size_t)i;
    const bool seg = particle_kind == MGFB_SEGMENT;
    V3 p = mk3(q[0], q[1], q[2]), d = mk3(q[3], q[4], q[5]);
    if (seg) d = d - p;   // geom.rs:848-850
    const float DT = seg ? 1.0f : __builtin_huge_valf();
    mgfb_shape S = shapes[i];
    float t = 0.0f; V3 ip = zero3(); bool ok = false;
    switch (S.kind) {
        case MGFB_PLANE: ok = ray_plane(p, d, sh_plane(S), DT, &t, &ip); break;
        case MGFB_TRIANGLE: { Tri tr = sh_tri(S); ok = ray_plane(p, d, tr.plane(), DT, &t, &ip) && tr.contains(ip); break; }      // collision.rs:186-200
        case MGFB_RECTANGLE: { Rct rc = sh_rect(S); ok = ray_plane(p, d, rc.plane(), DT, &t, &ip) && rc.contains(ip); break; }
        case MGFB_AABB: ok = ray_aabb(p, d, mk3(S.p[0], S.p[1], S.p[2]), mk3(S.p[3], S.p[4], S.p[5]), DT, &t, &ip); break;
        case MGFB_OBB: {   // collision.rs:238-247: rotate the particle around the centre (geom.rs:829-836, :853-861), then the AABB test
            V3 c = mk3(S.p[0], S.p[1], S.p[2]);
            Q4 rot; rot.s = S.p[6]; rot.v = mk3(S.p[7], S.p[8], S.p[9]);
            V3 rp = qrot(rot, p - c) + c, rd = qrot(rot, d);
            if (seg) { V3 b = rp + rd; rd = b - rp; }
            ok = ray_aabb(rp, rd, c, mk3(S.p[3], S.p[4], S.p[5]), DT, &t, &ip);
            break;
        }
        case MGFB_SPHERE:
            if (S.v[0] != 0.0f || S.v[1] != 0.0f || S.v[2] != 0.0f)   // Moving<Sphere>: the capsule its sweep covers (collision.rs:361-373)
                ok = ray_capsule(p, d, mk3(S.p[0], S.p[1], S.p[2]), sh_vel(S), S.p[3], &t, &ip) && !(t > DT);
            else ok = ray_sphere(p, d, mk3(S.p[0], S.p[1], S.p[2]), S.p[3], &t, &ip) && !(t > DT);
            break;
        case MGFB_CAPSULE: { Cap c = sh_capsule(S); ok = ray_capsule(p, d, c.a, c.d, c.r, &t, &ip) && !(t > DT); break; }
        default: break;
    }
    hit[i] = ok ? 1u : 0u;
    mgfb_intersection o;
    o.p[0] = ok ? ip.x : 0.0f; o.p[1] = ok ? ip.y : 0.0f; o.p[2] = ok ? ip.z : 0.0f; o.t = ok ? t : 0.0f;
    out[i] = o;
}
}  // namespace

extern "C" int32_t mgfb_intersections_batch(mgfb_ctx* ctx, uint32_t particle_kind, const float* particles, const mgfb_shape* shapes, uint32_t n,
                                            mgfb_intersection* out, uint32_t* hit) {
    if (!ctx || (n && (!particles || !shapes || !out || !hit))) return fail(ctx, MGFB_ERR_INVALID_ARG, "null argument");
    if (particle_kind > MGFB_SEGMENT) return fail(ctx, MGFB_ERR_INVALID_ARG, "particle kind must be MGFB_RAY or MGFB_SEGMENT");
    if (n == 0) return MGFB_OK;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t k = shapes[i].kind;
        if (k > MGFB_OBB) return fail(ctx, MGFB_ERR_INVALID_ARG, "unknown shape kind");
        if ((k == MGFB_SPHERE && !(shapes[i].p[3] > 0.0f)) || (k == MGFB_CAPSULE && !(shapes[i].p[6] > 0.0f)))
            return fail(ctx, MGFB_ERR_INVALID_ARG, "radius must be > 0 (geom.rs:300,328)");
    }
    CU(cudaSetDevice(ctx->device));
    Buf dp, ds, dout, dh;
    int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        TRY(ensure(ctx, dp, (size_t)n * 24)); TRY(ensure(ctx, ds, (size_t)n * sizeof(mgfb_shape)));
        TRY(ensure(ctx, dout, (size_t)n * sizeof(mgfb_intersection))); TRY(ensure(ctx, dh, (size_t)n * 4));
        CU(cudaMemcpyAsync(dp.p, particles, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ds.p, shapes, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream));
        k_intersections_batch<<<(n + 127) / 128, 128, 0, ctx->stream>>>(particle_kind, dp.as<float>(), ds.as<mgfb_shape>(), n,
                                                                        dout.as<mgfb_intersection>(), dh.as<unsigned>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(mgfb_intersection), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(hit, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 1;
        return MGFB_OK;
    };
    s = body();
    release(dp); release(ds); release(dout); release(dh);
    return s;
}

// ---------------------------------------------------------------- ContactPruner + Manifold::from (manifold.rs:42-148)
namespace {
#define PRUNER_KEEP 8   // contacts a device pruner holds (the reference's SmallVec<[LocalContact; 4]> spills to the heap; a Manifold row holds 4)
struct PrunedContact { V3 la, lb, ga, gb, n; float t; };
__device__ __forceinline__ PrunedContact load_lc(const mgfb_local_contact& c) {
    PrunedContact p;
    p.la = mk3(c.local_a[0], c.local_a[1], c.local_a[2]); p.lb = mk3(c.local_b[0], c.local_b[1], c.local_b[2]);
    p.ga = mk3(c.global.a[0], c.global.a[1], c.global.a[2]); p.gb = mk3(c.global.b[0], c.global.b[1], c.global.b[2]);
    p.n = mk3(c.global.n[0], c.global.n[1], c.global.n[2]); p.t = c.global.t;
    return p;
}
// One thread per group: ContactPruner::new(), push(contact) for the group's contacts in order (manifold.rs:72-102), then
// Manifold::from(pruner) (manifold.rs:131-148): time, the arithmetic mean of the kept normals, compute_basis of it, the local points.
__global__ void __launch_bounds__(128) k_manifolds_prune(const mgfb_local_contact* __restrict__ contacts, const unsigned* __restrict__ offsets, unsigned ngroups,
                                                         float threshold_sq, float* time, float* normal, float* tangent, unsigned* ncontacts, float* local_a,
                                                         float* local_b) {
    unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    PrunedContact keep[PRUNER_KEEP]; unsigned n = 0;
    float min_t = __builtin_huge_valf();
    for (unsigned k = offsets[g]; k < offsets[g + 1]; ++k) {
        const PrunedContact c = load_lc(contacts[k]);
        if (c.t < min_t - MGFB_EPS) { n = 1; keep[0] = c; min_t = c.t; continue; }     // an earlier collision replaces everything
        if (c.t > min_t + MGFB_EPS) continue;
        bool merged = false;
        for (unsigned j = 0; j < n && j < PRUNER_KEEP && !merged; ++j) {
            if (len2(c.ga - keep[j].ga) <= threshold_sq || len2(c.gb - keep[j].gb) <= threshold_sq) {
                // keep the one whose points are further from the objects' centres
                if (len2(keep[j].la) + len2(keep[j].lb) < len2(c.la) + len2(c.lb)) keep[j] = c;
                merged = true;
            }
        }
        if (merged) continue;
        if (n < PRUNER_KEEP) keep[n] = c;
        ++n;
    }
    V3 sum = zero3();
    for (unsigned j = 0; j < n && j < PRUNER_KEEP; ++j) sum = sum + keep[j].n;
    const V3 avg = sum / (float)n;                                                        // 0 / 0 = NaN for an empty pruner, like the reference
    V3 t0, t1; basis_from_normal(avg, &t0, &t1);
    time[g] = min_t; ncontacts[g] = n;
    normal[3 * g] = avg.x; normal[3 * g + 1] = avg.y; normal[3 * g + 2] = avg.z;
    tangent[6 * g] = t0.x; tangent[6 * g + 1] = t0.y; tangent[6 * g + 2] = t0.z; tangent[6 * g + 3] = t1.x; tangent[6 * g + 4] = t1.y; tangent[6 * g + 5] = t1.z;
    for (unsigned j = 0; j < 4; ++j) {
        const bool have = j < n;
        const V3 a = have ? keep[j].la : zero3(), b = have ? keep[j].lb : zero3();
        local_a[12 * g + 3 * j] = a.x; local_a[12 * g + 3 * j + 1] = a.y; local_a[12 * g + 3 * j + 2] = a.z;
        local_b[12 * g + 3 * j] = b.x; local_b[12 * g + 3 * j + 1] = b.y; local_b[12 * g + 3 * j + 2] = b.z;
    }
}
}  // namespace

extern "C" int32_t mgfb_manifolds_prune(mgfb_ctx* ctx, const mgfb_local_contact* contacts, const uint32_t* offsets, uint32_t ngroups, float* time, float* normal,
                                        float* tangent, uint32_t* ncontacts, float* local_a, float* local_b) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (ngroups == 0) return MGFB_OK;
    if (!offsets || !normal || !tangent || !ncontacts || !local_a || !local_b) return fail(ctx, MGFB_ERR_INVALID_ARG, "null array");
    for (uint32_t g = 0; g < ngroups; ++g) if (offsets[g + 1] < offsets[g]) return fail(ctx, MGFB_ERR_INVALID_ARG, "offsets must not decrease");
    const uint32_t total = offsets[ngroups];
    if (offsets[0] != 0 || (total && !contacts)) return fail(ctx, MGFB_ERR_INVALID_ARG, "offsets[0] must be 0 and contacts given");
    CU(cudaSetDevice(ctx->device));
    Buf dc, dof, dout; int32_t s = MGFB_OK;
    auto body = [&]() -> int32_t {
        const size_t per = 4 + 12 + 24 + 4 + 48 + 48;   // bytes of output per group
        TRY(ensure(ctx, dc, std::max<size_t>((size_t)total * sizeof(mgfb_local_contact), 16))); TRY(ensure(ctx, dof, ((size_t)ngroups + 1) * 4));
        TRY(ensure(ctx, dout, (size_t)ngroups * per));
        if (total) CU(cudaMemcpyAsync(dc.p, contacts, (size_t)total * sizeof(mgfb_local_contact), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dof.p, offsets, ((size_t)ngroups + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        float* d_time = dout.as<float>(); float* d_n = d_time + ngroups; float* d_t = d_n + 3 * (size_t)ngroups;
        unsigned* d_nc = reinterpret_cast<unsigned*>(d_t + 6 * (size_t)ngroups); float* d_la = reinterpret_cast<float*>(d_nc + ngroups); float* d_lb = d_la + 12 * (size_t)ngroups;
        k_manifolds_prune<<<(ngroups + 127) / 128, 128, 0, ctx->stream>>>(dc.as<mgfb_local_contact>(), dof.as<unsigned>(), ngroups, ctx->cfg.persistent_threshold_sq,
                                                                         d_time, d_n, d_t, d_nc, d_la, d_lb);
        CU(cudaGetLastError());
        ctx->launches += 1;
        auto down = [&](void* dst, const void* src, size_t bytes) { return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess; };
        CU(down(time, d_time, (size_t)ngroups * 4)); CU(down(normal, d_n, (size_t)ngroups * 12)); CU(down(tangent, d_t, (size_t)ngroups * 24));
        CU(down(ncontacts, d_nc, (size_t)ngroups * 4)); CU(down(local_a, d_la, (size_t)ngroups * 48)); CU(down(local_b, d_lb, (size_t)ngroups * 48));
        CU(cudaStreamSynchronize(ctx->stream));
        return MGFB_OK;
    };
    s = body(); release(dc); release(dof); release(dout);
    return s;
}
