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
