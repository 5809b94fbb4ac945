// gjk.cuh -- the reference's DISCRETE narrowphase on the device: GJK distance
// (src/simplex.rs:172-415 Simplex::closest_point_to_origin and its Vertex/Edge/Face/Volume state
// objects), EPA penetration (simplex.rs:417-553 compute_contact), the Minkowski difference and
// the Convex::support functions (geom.rs:1027-1133), behind the generic impls
// `Contacts for Convex x Convex` (collision.rs:497-519) and `Penetrates::separation`
// (collision.rs:404-425).  Never reached from World::step (every RigidBodyVec collider is a
// Moving<Component>): it is its own batch API.  Included at the end of capi.cu.
//
// Mapping: one thread per pair; pairs are binned by (kind_a, kind_b) on the host so each launch
// is one template instantiation with the two support functions inlined (no dispatch
// divergence).  The EPA polytope lives in per-thread local memory in INDEXED form -- support
// points once (36 B each), faces and horizon edges as byte indices -- and the face slab follows
// mgf's Pool free-list discipline (pool.rs:81-113), because slot order decides ties in the
// closest-face search and the order new faces are created in.
//
// The reference keeps the horizon in a HashMap whose iteration order is randomised per process
// (simplex.rs:523): it is not bit-deterministic on ties itself.  Like the oracle we iterate in
// insertion order, one of the orders the reference can take.
#pragma once

namespace {

struct SP { V3 p, a, b; };   // geom.rs:1077 SupportPoint

template <int K> struct GShape;
template <> struct GShape<MGFB_SPHERE> {
    V3 c; float r;
    __device__ explicit GShape(const mgfb_shape& s) : c(mk3(s.p[0], s.p[1], s.p[2])), r(s.p[3]) {}
    __device__ V3 support(V3 d) const { return c + d * r; }   // geom.rs:1050-1054
};
template <> struct GShape<MGFB_CAPSULE> {
    V3 a, d; float r;
    __device__ explicit GShape(const mgfb_shape& s) : a(mk3(s.p[0], s.p[1], s.p[2])), d(mk3(s.p[3], s.p[4], s.p[5])), r(s.p[6]) {}
    __device__ V3 support(V3 dir) const {   // geom.rs:1056-1072
        V3 c = a + d * 0.5f;
        V3 u = unit(d);
        float ud = dot3(u, dir);
        V3 w = dir - u * ud;
        V3 cap = ((len(d) * 0.5f + r) * u) * sgn(ud);
        if (w.x == 0.0f && w.y == 0.0f && w.z == 0.0f) return c + cap;
        return c + cap + unit(w) * r;
    }
};
template <> struct GShape<MGFB_AABB> {
    V3 c, r;
    __device__ explicit GShape(const mgfb_shape& s) : c(mk3(s.p[0], s.p[1], s.p[2])), r(mk3(s.p[3], s.p[4], s.p[5])) {}
    __device__ V3 support(V3 d) const { return mk3(sgn(d.x) * r.x, sgn(d.y) * r.y, sgn(d.z) * r.z) + c; }   // geom.rs:1027-1035
};
template <> struct GShape<MGFB_OBB> {
    V3 c, r; Q4 q;
    __device__ explicit GShape(const mgfb_shape& s)
        : c(mk3(s.p[0], s.p[1], s.p[2])), r(mk3(s.p[3], s.p[4], s.p[5])), q(mkq(s.p[6], mk3(s.p[7], s.p[8], s.p[9]))) {}
    __device__ V3 support(V3 d0) const {   // geom.rs:1037-1048
        V3 d = qrot(qinv(q), d0);
        return qrot(q, mk3(sgn(d.x) * r.x, sgn(d.y) * r.y, sgn(d.z) * r.z)) + c;
    }
};

template <> struct GShape<MGFB_CONVEX_MESH> {   // mesh.rs:141: a slice of the context's vertex pool (the host wrote its device address into p[2..3])
    const float* v; unsigned n;
    __device__ explicit GShape(const mgfb_shape& s) : n((unsigned)s.p[1]) {
        unsigned long long bits = (unsigned long long)__float_as_uint(s.p[2]) | ((unsigned long long)__float_as_uint(s.p[3]) << 32);
        v = reinterpret_cast<const float*>(bits);
    }
    __device__ V3 support(V3 d) const {   // mesh.rs:223-236: strict >, the first of equally good vertices wins
        V3 best = mk3(v[0], v[1], v[2]);
        float best_norm = dot3(d, best);
        for (unsigned i = 1; i < n; ++i) {
            V3 p = mk3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
            float norm = dot3(d, p);
            if (norm > best_norm) { best = p; best_norm = norm; }
        }
        return best;
    }
};

// geom.rs:1099-1133 MinkowskiDiff::support_pt
template <class A, class B>
__device__ __forceinline__ SP mink_support(const A& sa, const B& sb, V3 axis) {
    SP s; s.a = sa.support(axis); s.b = sb.support(-axis); s.p = s.a - s.b;
    return s;
}

enum { GJK_VERTEX = 1, GJK_EDGE = 2, GJK_FACE = 3, GJK_VOLUME = 4 };

// FaceSimplex::min_norm (simplex.rs:274-340): closest point of triangle simp[0..3) to the origin;
// reorders simp and says which state comes next.
__device__ __forceinline__ V3 gjk_face_min_norm(SP simp[4], int* next) {
    V3 a = simp[0].p, b = simp[1].p, c = simp[2].p;
    V3 ab = b - a, ac = c - a, ap = -a;
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { *next = GJK_EDGE; return simp[0].p; }
    V3 bp = -b;
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { simp[0] = simp[1]; *next = GJK_EDGE; return simp[1].p; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float v = d1 / (d1 - d3);
        *next = GJK_FACE; return simp[0].p + ab * v;
    }
    V3 cp = -c;
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { simp[0] = simp[2]; *next = GJK_EDGE; return simp[2].p; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        simp[1] = simp[2];
        *next = GJK_FACE; return simp[0].p + ac * w;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        simp[0] = simp[2];
        *next = GJK_FACE; return simp[1].p + (simp[2].p - simp[1].p) * w;
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    *next = GJK_VOLUME; return simp[0].p + ab * v + ac * w;
}
__device__ __forceinline__ bool gjk_origin_outside(V3 a, V3 b, V3 c, V3 d) {   // simplex.rs:342-349
    V3 n = cross3(b - a, c - a);
    return dot3(-a, n) * dot3(d - a, n) < 0.0f;
}
// VolumeSimplex::min_norm helper (simplex.rs:351-415): one face of the tetrahedron
__device__ __forceinline__ void gjk_volume_face(bool outside, const SP& n0, const SP& n1, const SP& n2, const SP& n3, bool update_best,
                                                SP simp[4], V3* closest, float* best, int* next_state) {
    if (!outside) return;
    SP ns[4] = {n0, n1, n2, n3};
    int st;
    V3 p = gjk_face_min_norm(ns, &st);
    float nd = len2(p);
    if (nd < *best) {
        *closest = p;
        if (update_best) *best = nd;
        *next_state = st;
        for (int i = 0; i < 4; ++i) simp[i] = ns[i];
    }
}
__device__ __forceinline__ V3 gjk_min_norm(SP simp[4], int state, int* next) {
    if (state == GJK_VERTEX) { *next = GJK_EDGE; return simp[0].p; }   // simplex.rs:224-238
    if (state == GJK_EDGE) {                                           // simplex.rs:240-266
        V3 ab = simp[1].p - simp[0].p;
        float t = dot3(ab, -simp[0].p);
        if (t <= 0.0f) { *next = GJK_EDGE; return simp[0].p; }
        float denom = dot3(ab, ab);
        if (t >= denom) { simp[0] = simp[1]; *next = GJK_EDGE; return simp[1].p; }
        *next = GJK_FACE; return simp[0].p + ab * (t / denom);
    }
    if (state == GJK_FACE) return gjk_face_min_norm(simp, next);
    V3 closest = zero3();
    float best = __int_as_float(0x7f800000);
    int next_state = GJK_VERTEX;
    SP a = simp[0], b = simp[1], c = simp[2], d = simp[3];
    gjk_volume_face(gjk_origin_outside(a.p, b.p, c.p, d.p), a, b, c, d, true, simp, &closest, &best, &next_state);
    gjk_volume_face(gjk_origin_outside(a.p, c.p, d.p, b.p), a, c, d, b, true, simp, &closest, &best, &next_state);
    gjk_volume_face(gjk_origin_outside(a.p, d.p, b.p, c.p), a, d, b, c, true, simp, &closest, &best, &next_state);
    gjk_volume_face(gjk_origin_outside(b.p, d.p, c.p, a.p), b, d, c, a, false, simp, &closest, &best, &next_state);   // sic: best not updated
    *next = next_state;
    return closest;
}
#define GJK_MAX_STEPS 256    // the reference loops until convergence; a NaN input must not hang the GPU
// simplex.rs:172-200.  Returns false if the step cap was hit.
template <class A, class B>
__device__ bool gjk_closest_point_to_origin(const A& sa, const B& sb, SP simp[4], int* state, V3* out) {
    V3 prev = zero3();
    for (int step = 0; step < GJK_MAX_STEPS; ++step) {
        int next;
        V3 mn = gjk_min_norm(simp, *state, &next);
        if (len2(mn) < 1.0e-6f) {   // COLLISION_EPSILON: the origin is inside; grow to a tetrahedron for EPA
            for (int i = *state; i < 4; ++i) {
                V3 m2 = -mk3(prev.z, prev.x, prev.y);
                V3 dir = -unit(m2);
                simp[i] = mink_support(sa, sb, dir);
                prev = dir;
            }
            *state = GJK_VOLUME;
            *out = zero3();
            return true;
        }
        SP sp = mink_support(sa, sb, -unit(mn));
        prev = mn;
        if (len2(mn) >= len2(sp.p)) { *out = mn; return true; }
        *state = next;
        simp[next - 1] = sp;   // add_point
    }
    *out = zero3();
    return false;
}

// ---- EPA (simplex.rs:456-553): one WARP per overlapping pair, polytope staged in shared memory.
// The reference's polytope is not kept convex (horizon edges are matched by exact f32 bit pattern,
// simplex.rs:426-451, so near-duplicate support points leave interior faces behind) and grows to
// thousands of faces within its 101 iterations on round shapes; every iteration scans all faces
// twice.  So: the 32 lanes share those scans, and everything whose ORDER the result depends on
// (Pool slot reuse, horizon insertion order, first-minimum tie break) is done in the reference's
// sequential order through ballots.
#define EPA_MAX_ITERS 100
#define EPA_MAXV (4 + EPA_MAX_ITERS + 1)
#define EPA_MAXF 6144
#define EPA_MAXE 3072
struct EpaShared {
    SP v[EPA_MAXV];                         // support points, referenced by byte index
    unsigned short nextf[EPA_MAXF];          // Pool free list (pool.rs:28-35)
    unsigned char fa[EPA_MAXF], fb[EPA_MAXF], fc[EPA_MAXF];
    unsigned char tag[EPA_MAXF];             // 0 free-list end, 1 free-list ptr, 2 occupied, 3 occupied + sees the new point
    unsigned char ea[EPA_MAXE], eb[EPA_MAXE], elive[EPA_MAXE];   // horizon "HashMap" in insertion order
};
struct EpaState { int nv, nslots, ne, free_head; bool has_free, overflow; };   // warp-uniform registers

__device__ __forceinline__ void epa_push_face(EpaShared& P, EpaState& S, int a, int b, int c) {   // pool.rs:81-96 (lane 0)
    int slot;
    if (S.has_free) {
        slot = S.free_head;
        if (P.tag[slot] == 0) S.has_free = false; else S.free_head = P.nextf[slot];
    } else {
        if (S.nslots >= EPA_MAXF) { S.overflow = true; return; }
        slot = S.nslots++;
    }
    P.tag[slot] = 2; P.fa[slot] = (unsigned char)a; P.fb[slot] = (unsigned char)b; P.fc[slot] = (unsigned char)c;
}
__device__ __forceinline__ void epa_remove_face(EpaShared& P, EpaState& S, int slot) {            // pool.rs:100-113 (lane 0)
    if (S.has_free) { P.tag[slot] = 1; P.nextf[slot] = (unsigned short)S.free_head; } else P.tag[slot] = 0;
    S.has_free = true; S.free_head = slot;
}
__device__ __forceinline__ bool same_bits(V3 a, V3 b) {   // the HashMap key is the f32 bit pattern (simplex.rs:426-431)
    return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) && __float_as_uint(a.z) == __float_as_uint(b.z);
}
// simplex.rs:423-451 for one directed edge (a, b), the whole warp: the FIRST live entry holding the
// reversed edge is removed; else the first live entry with the same key is overwritten in place
// (HashMap::insert); else the edge is appended.
__device__ __forceinline__ void epa_add_edge(EpaShared& P, EpaState& S, int a, int b, unsigned lane) {
    const V3 pa = P.v[a].p, pb = P.v[b].p;
    int same = -1;
    for (int e0 = 0; e0 < S.ne; e0 += 32) {
        int e = e0 + (int)lane;
        bool rev = false, fwd = false;
        if (e < S.ne && P.elive[e]) {
            V3 qa = P.v[P.ea[e]].p, qb = P.v[P.eb[e]].p;
            rev = same_bits(qa, pb) && same_bits(qb, pa);
            fwd = same_bits(qa, pa) && same_bits(qb, pb);
        }
        unsigned mr = __ballot_sync(0xffffffffu, rev), mf = __ballot_sync(0xffffffffu, fwd);
        if (mr) { if (lane == 0) P.elive[e0 + __ffs(mr) - 1] = 0; __syncwarp(); return; }
        if (mf && same < 0) same = e0 + __ffs(mf) - 1;
    }
    if (same >= 0) { if (lane == 0) { P.ea[same] = (unsigned char)a; P.eb[same] = (unsigned char)b; } __syncwarp(); return; }
    if (S.ne >= EPA_MAXE) { S.overflow = true; return; }
    if (lane == 0) { P.ea[S.ne] = (unsigned char)a; P.eb[S.ne] = (unsigned char)b; P.elive[S.ne] = 1; }
    S.ne++;
    __syncwarp();
}
__device__ __forceinline__ V3 epa_face_normal(const EpaShared& P, int f) {   // geom.rs:149
    V3 a = P.v[P.fa[f]].p, b = P.v[P.fb[f]].p, c = P.v[P.fc[f]].p;
    return unit(cross3(b - a, c - a));
}
// Returns 1 on success, -2 when a fixed capacity was exceeded, -4 where the reference would panic.
template <class A, class B>
__device__ int epa_compute_contact(const A& sa, const B& sb, const SP* simp, EpaShared& P, unsigned lane, Hit* out, int* iterations) {
    EpaState S; S.nv = 4; S.nslots = 0; S.ne = 0; S.free_head = 0; S.has_free = false; S.overflow = false;
    if (lane < 4) P.v[lane] = simp[lane];
    if (lane == 0) { epa_push_face(P, S, 0, 1, 2); epa_push_face(P, S, 0, 2, 3); epa_push_face(P, S, 0, 3, 1); epa_push_face(P, S, 1, 3, 2); }
    S.nslots = 4;   // (lane 0 counted them; keep the state warp-uniform)
    __syncwarp();
    for (int iter = 0; iter <= EPA_MAX_ITERS; ++iter) {
        // closest face: first minimum in slot order, strict ">" (simplex.rs:473-481)
        float best = __int_as_float(0x7f800000); int besti = 0x7fffffff; V3 bestn = zero3();
        for (int f = (int)lane; f < S.nslots; f += 32) {
            if (P.tag[f] != 2) continue;
            V3 n = epa_face_normal(P, f);
            float dist = fabsf(dot3(n, P.v[P.fa[f]].p));
            if (best > dist) { best = dist; besti = f; bestn = n; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            float ox = __shfl_xor_sync(0xffffffffu, bestn.x, o), oy = __shfl_xor_sync(0xffffffffu, bestn.y, o), oz = __shfl_xor_sync(0xffffffffu, bestn.z, o);
            if (oi != 0x7fffffff && (besti == 0x7fffffff || best > od || (best == od && oi < besti))) { best = od; besti = oi; bestn = mk3(ox, oy, oz); }
        }
        const float closest_dist = best; const V3 closest_n = bestn;
        const int closest_i = besti == 0x7fffffff ? 0 : besti;
        if (P.tag[closest_i] != 2) return -4;   // `tris[closest_i]` on a free slot: the reference panics here (pool.rs:111)
        SP ca = P.v[P.fa[closest_i]], cb = P.v[P.fb[closest_i]], cc = P.v[P.fc[closest_i]];
        SP sup = mink_support(sa, sb, closest_n);
        float v = dot3(closest_n, sup.p) - closest_dist;
        if (v < 1.0e-6f || iter == EPA_MAX_ITERS) {
            // barycentric coordinates of the projected origin (geom.rs:154-167), contact on shape A
            V3 p = closest_dist * closest_n;
            V3 v0 = cb.p - ca.p, v1 = cc.p - ca.p, v2 = p - ca.p;
            float d0 = dot3(v0, v0), d1 = dot3(v0, v1), d2 = dot3(v1, v1), d3 = dot3(v2, v0), d4 = dot3(v2, v1);
            float denom = d0 * d2 - d1 * d1;
            float bv = (d2 * d3 - d1 * d4) / denom;
            float bw = (d0 * d4 - d1 * d3) / denom;
            float u = bv, vv = bw, w = 1.0f - bv - bw;
            V3 a = u * ca.a + vv * cb.a + w * cc.a;
            out->a = a; out->b = a - closest_dist * closest_n; out->n = closest_n; out->t = 0.0f;
            *iterations = iter;
            return 1;
        }
        if (S.nv >= EPA_MAXV) return -2;
        const int isup = S.nv++;
        if (lane == 0) P.v[isup] = sup;
        __syncwarp();
        // faces that see the new point, in slot order: their edges enter the horizon map (simplex.rs:506-516)
        for (int f0 = 0; f0 < S.nslots; f0 += 32) {
            int f = f0 + (int)lane;
            bool sees = false;
            if (f < S.nslots && P.tag[f] == 2) {
                V3 n = epa_face_normal(P, f);
                sees = dot3(n, sup.p - P.v[P.fa[f]].p) > 0.0f;
            }
            unsigned m = __ballot_sync(0xffffffffu, sees);
            if (sees) P.tag[f] = 3;
            __syncwarp();
            while (m) {
                int g = f0 + __ffs(m) - 1; m &= m - 1;
                int ga = P.fa[g], gb = P.fb[g], gc = P.fc[g];
                epa_add_edge(P, S, ga, gb, lane); epa_add_edge(P, S, gb, gc, lane); epa_add_edge(P, S, gc, ga, lane);
            }
            if (S.overflow) return -2;
        }
        // remove them in ascending slot order (simplex.rs:518-520), then one new face per live horizon edge
        for (int f0 = 0; f0 < S.nslots; f0 += 32) {
            int f = f0 + (int)lane;
            unsigned m = __ballot_sync(0xffffffffu, f < S.nslots && P.tag[f] == 3);
            if (lane == 0) while (m) { epa_remove_face(P, S, f0 + __ffs(m) - 1); m &= m - 1; }
            S.has_free = __shfl_sync(0xffffffffu, (int)S.has_free, 0) != 0; S.free_head = __shfl_sync(0xffffffffu, S.free_head, 0);
        }
        __syncwarp();
        for (int e0 = 0; e0 < S.ne; e0 += 32) {
            int e = e0 + (int)lane;
            unsigned m = __ballot_sync(0xffffffffu, e < S.ne && P.elive[e]);
            if (lane == 0) while (m) { int k = e0 + __ffs(m) - 1; m &= m - 1; epa_push_face(P, S, isup, P.ea[k], P.eb[k]); }
            S.has_free = __shfl_sync(0xffffffffu, (int)S.has_free, 0) != 0; S.free_head = __shfl_sync(0xffffffffu, S.free_head, 0);
            S.nslots = __shfl_sync(0xffffffffu, S.nslots, 0); S.overflow = __shfl_sync(0xffffffffu, (int)S.overflow, 0) != 0;
        }
        __syncwarp();
        S.ne = 0;
        if (S.overflow) return -2;
    }
    return -2;   // unreachable
}

// status: 0 no contact / None, 1 contact / Some, 2 polytope capacity exceeded, 3 GJK step cap hit, 4 reference panics
struct EpaWork { unsigned pair; SP simp[4]; };
// Pass 1, one thread per pair: GJK.  Separated pairs are finished here; overlapping pairs go to the EPA work list.
template <int KA, int KB, bool SEPARATION>
__global__ void __launch_bounds__(128) k_gjk(const mgfb_shape* __restrict__ a, const mgfb_shape* __restrict__ b, const unsigned* __restrict__ index,
                                             unsigned n, float* sep, unsigned* status, unsigned* epa_iters, EpaWork* work, unsigned* work_count) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned i = index[t];
    GShape<KA> sa(a[i]); GShape<KB> sb(b[i]);
    // seeds: +-y for contacts (collision.rs:503), +-x for separation (collision.rs:410)
    V3 d = SEPARATION ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);
    SP simp[4];
    simp[0] = mink_support(sa, sb, d); simp[1] = mink_support(sa, sb, -d);
    simp[2].p = simp[2].a = simp[2].b = zero3(); simp[3] = simp[2];
    int state = GJK_EDGE;
    V3 md;
    unsigned st = 0;
    if (!gjk_closest_point_to_origin(sa, sb, simp, &state, &md)) st = 3;
    else {
        float mag2 = len2(md);
        if (SEPARATION) {
            if (!(mag2 < 1.0e-6f)) { st = 1; sep[i] = sqrtf(mag2); }   // collision.rs:417-421
        } else if (!(mag2 > 1.0e-6f)) {                                  // collision.rs:512-517 -> EPA
            unsigned k = atomicAdd(work_count, 1u);
            work[k].pair = i;
            for (int j = 0; j < 4; ++j) work[k].simp[j] = simp[j];
        }
    }
    status[i] = st;
    if (epa_iters) epa_iters[i] = 0;
}
// Pass 2, one warp (= one CTA) per overlapping pair, persistent over the work list.
template <int KA, int KB>
__global__ void __launch_bounds__(32) k_epa(const mgfb_shape* __restrict__ a, const mgfb_shape* __restrict__ b, const EpaWork* __restrict__ work,
                                            const unsigned* __restrict__ work_count, unsigned* work_next, mgfb_contact* out, unsigned* status, unsigned* epa_iters) {
    extern __shared__ __align__(16) unsigned char epa_smem[];
    EpaShared& P = *reinterpret_cast<EpaShared*>(epa_smem);
    const unsigned lane = threadIdx.x, nw = *work_count;
    // items are claimed from a shared cursor: one EPA run takes 1..101 iterations over a polytope of 4..4000 faces, so a
    // fixed stride would leave most warps idle behind the few long ones
    for (;;) {
        unsigned k = 0;
        if (lane == 0) k = atomicAdd(work_next, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= nw) break;
        const unsigned i = work[k].pair;
        GShape<KA> sa(a[i]); GShape<KB> sb(b[i]);
        Hit h; int it = 0;
        __syncwarp();
        int rc = epa_compute_contact(sa, sb, work[k].simp, P, lane, &h, &it);
        if (lane == 0) {
            if (rc == 1) {
                mgfb_contact* o = out + i;
                o->a[0] = h.a.x; o->a[1] = h.a.y; o->a[2] = h.a.z; o->b[0] = h.b.x; o->b[1] = h.b.y; o->b[2] = h.b.z;
                o->n[0] = h.n.x; o->n[1] = h.n.y; o->n[2] = h.n.z; o->t = h.t;
                status[i] = 1;
                if (epa_iters) epa_iters[i] = (unsigned)it;
            } else status[i] = (unsigned)(-rc);
        }
    }
}

template <int KA, int KB>
void launch_gjk(bool separation, const mgfb_shape* a, const mgfb_shape* b, const unsigned* index, unsigned n, mgfb_contact* out, float* sep,
                unsigned* status, unsigned* epa_iters, EpaWork* work, unsigned* work_count, unsigned* work_next, int epa_grid, cudaStream_t s) {
    unsigned blocks = (n + 127) / 128;
    if (separation) { k_gjk<KA, KB, true><<<blocks, 128, 0, s>>>(a, b, index, n, sep, status, epa_iters, work, work_count); return; }
    k_gjk<KA, KB, false><<<blocks, 128, 0, s>>>(a, b, index, n, sep, status, epa_iters, work, work_count);
    cudaFuncSetAttribute(k_epa<KA, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EpaShared));
    k_epa<KA, KB><<<epa_grid, 32, sizeof(EpaShared), s>>>(a, b, work, work_count, work_next, out, status, epa_iters);
}
typedef void (*gjk_launcher)(bool, const mgfb_shape*, const mgfb_shape*, const unsigned*, unsigned, mgfb_contact*, float*, unsigned*, unsigned*, EpaWork*,
                             unsigned*, unsigned*, int, cudaStream_t);
#define GJK_NK 5
const int GJK_KINDS[GJK_NK] = {MGFB_SPHERE, MGFB_CAPSULE, MGFB_AABB, MGFB_OBB, MGFB_CONVEX_MESH};
const gjk_launcher GJK_TABLE[GJK_NK][GJK_NK] = {
    {launch_gjk<MGFB_SPHERE, MGFB_SPHERE>, launch_gjk<MGFB_SPHERE, MGFB_CAPSULE>, launch_gjk<MGFB_SPHERE, MGFB_AABB>, launch_gjk<MGFB_SPHERE, MGFB_OBB>, launch_gjk<MGFB_SPHERE, MGFB_CONVEX_MESH>},
    {launch_gjk<MGFB_CAPSULE, MGFB_SPHERE>, launch_gjk<MGFB_CAPSULE, MGFB_CAPSULE>, launch_gjk<MGFB_CAPSULE, MGFB_AABB>, launch_gjk<MGFB_CAPSULE, MGFB_OBB>, launch_gjk<MGFB_CAPSULE, MGFB_CONVEX_MESH>},
    {launch_gjk<MGFB_AABB, MGFB_SPHERE>, launch_gjk<MGFB_AABB, MGFB_CAPSULE>, launch_gjk<MGFB_AABB, MGFB_AABB>, launch_gjk<MGFB_AABB, MGFB_OBB>, launch_gjk<MGFB_AABB, MGFB_CONVEX_MESH>},
    {launch_gjk<MGFB_OBB, MGFB_SPHERE>, launch_gjk<MGFB_OBB, MGFB_CAPSULE>, launch_gjk<MGFB_OBB, MGFB_AABB>, launch_gjk<MGFB_OBB, MGFB_OBB>, launch_gjk<MGFB_OBB, MGFB_CONVEX_MESH>},
    {launch_gjk<MGFB_CONVEX_MESH, MGFB_SPHERE>, launch_gjk<MGFB_CONVEX_MESH, MGFB_CAPSULE>, launch_gjk<MGFB_CONVEX_MESH, MGFB_AABB>, launch_gjk<MGFB_CONVEX_MESH, MGFB_OBB>, launch_gjk<MGFB_CONVEX_MESH, MGFB_CONVEX_MESH>}};

int32_t gjk_batch_impl(mgfb_ctx* ctx, bool separation, const mgfb_shape* a, const mgfb_shape* b, uint32_t n, mgfb_contact* out, float* sep,
                       uint32_t* status, uint32_t* epa_iters) {
    if (!ctx) return MGFB_ERR_INVALID_ARG;
    if (n == 0) return MGFB_OK;
    if (!a || !b || !status || (separation ? !sep : !out)) return fail(ctx, MGFB_ERR_INVALID_ARG, "null array");
    auto slot = [](uint32_t kind) { for (int k = 0; k < GJK_NK; ++k) if ((uint32_t)GJK_KINDS[k] == kind) return k; return -1; };
    // bin by (kind_a, kind_b): one divergence-free launch per shape pair
    const int NB = GJK_NK * GJK_NK;
    std::vector<unsigned> index(n); unsigned count[GJK_NK * GJK_NK + 1] = {0};
    std::vector<mgfb_shape> ha, hb;   // copies with the vertex pool's device address written into the CONVEX_MESH shapes
    auto prepare = [&](const mgfb_shape* src, std::vector<mgfb_shape>& dst) -> bool {
        bool any = false;
        for (uint32_t i = 0; i < n; ++i) any = any || src[i].kind == MGFB_CONVEX_MESH;
        if (!any) return true;
        dst.assign(src, src + n);
        for (uint32_t i = 0; i < n; ++i) {
            mgfb_shape& s = dst[i];
            if (s.kind != MGFB_CONVEX_MESH) continue;
            const float first = s.p[0], cnt = s.p[1];
            if (!(first >= 0.0f) || !(cnt >= 1.0f) || first != floorf(first) || cnt != floorf(cnt) || (double)first + (double)cnt > (double)ctx->convex_n) return false;
            unsigned long long bits = (unsigned long long)(ctx->convex_pool.as<float>() + 3 * (size_t)first);
            unsigned lo = (unsigned)bits, hi = (unsigned)(bits >> 32);
            std::memcpy(&s.p[2], &lo, 4); std::memcpy(&s.p[3], &hi, 4);
        }
        return true;
    };
    if (!prepare(a, ha) || !prepare(b, hb))
        return fail(ctx, MGFB_ERR_INVALID_ARG, "a CONVEX_MESH names vertices outside the pool of mgfb_convex_vertices_set (verts[0] on an empty mesh panics, mesh.rs:225)");
    if (!ha.empty()) a = ha.data();
    if (!hb.empty()) b = hb.data();
    for (uint32_t i = 0; i < n; ++i) {
        int ka = slot(a[i].kind), kb = slot(b[i].kind);
        if (ka < 0 || kb < 0) return fail(ctx, MGFB_ERR_INVALID_ARG, "GJK shapes must be Sphere, Capsule, AABB, OBB or ConvexMesh (the Convex implementors, geom.rs:1027-1072, mesh.rs:223)");
        count[ka * GJK_NK + kb + 1]++;
    }
    for (int k = 0; k < NB; ++k) count[k + 1] += count[k];
    { unsigned cur[GJK_NK * GJK_NK]; for (int k = 0; k < NB; ++k) cur[k] = count[k];
      for (uint32_t i = 0; i < n; ++i) index[cur[slot(a[i].kind) * GJK_NK + slot(b[i].kind)]++] = i; }
    CU(cudaSetDevice(ctx->device));
    Buf da, db, di, dout, dsep, dst, dit, dwork, dcount;
    int32_t rc = MGFB_OK;
    auto done = [&](int32_t code) { release(da); release(db); release(di); release(dout); release(dsep); release(dst); release(dit); release(dwork); release(dcount); return code; };
    if ((rc = ensure(ctx, da, (size_t)n * sizeof(mgfb_shape))) || (rc = ensure(ctx, db, (size_t)n * sizeof(mgfb_shape))) || (rc = ensure(ctx, di, (size_t)n * 4)) ||
        (rc = ensure(ctx, dout, (size_t)n * sizeof(mgfb_contact))) || (rc = ensure(ctx, dsep, (size_t)n * 4)) || (rc = ensure(ctx, dst, (size_t)n * 4)) ||
        (rc = ensure(ctx, dit, (size_t)n * 4)) || (rc = ensure(ctx, dwork, separation ? sizeof(EpaWork) : (size_t)n * sizeof(EpaWork))) ||
        (rc = ensure(ctx, dcount, 64 * 4)))
        return done(rc);
    cudaMemcpyAsync(da.p, a, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(db.p, b, (size_t)n * sizeof(mgfb_shape), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(di.p, index.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(dout.p, 0, (size_t)n * sizeof(mgfb_contact), ctx->stream);
    cudaMemsetAsync(dsep.p, 0, (size_t)n * 4, ctx->stream);
    cudaMemsetAsync(dcount.p, 0, 64 * 4, ctx->stream);   // [0, 32): items per bin, [32, 64): the bins' work cursors
    int epa_grid = ctx->num_sms * 4;   // 4 x 52 KB polytopes per SM
    // One EPA run is ONE warp for up to ~10^8 cycles (100 iterations over a polytope of thousands of faces), so a launch lasts
    // as long as its slowest pair whatever its size: the 25 shape-pair launches run side by side on 8 streams.
    const int NS = 8;
    if (!ctx->gjk_streams[0]) {
        for (int k = 0; k < NS; ++k) if (cudaStreamCreateWithFlags(&ctx->gjk_streams[k], cudaStreamNonBlocking) != cudaSuccess) return done(fail(ctx, MGFB_ERR_CUDA, "cudaStreamCreate"));
        for (int k = 0; k <= NS; ++k) if (cudaEventCreateWithFlags(&ctx->gjk_ev[k], cudaEventDisableTiming) != cudaSuccess) return done(fail(ctx, MGFB_ERR_CUDA, "cudaEventCreate"));
    }
    cudaEventRecord(ctx->gjk_ev[NS], ctx->stream);   // inputs uploaded, outputs cleared
    for (int k = 0; k < NS; ++k) cudaStreamWaitEvent(ctx->gjk_streams[k], ctx->gjk_ev[NS], 0);
    int used = 0;
    for (int k = 0; k < NB; ++k) {
        unsigned m = count[k + 1] - count[k];
        if (!m) continue;
        cudaStream_t bin_stream = ctx->gjk_streams[used++ % NS];
        // each bin's EPA work list is a slice of dwork starting at the bin's first pair
        GJK_TABLE[k / GJK_NK][k % GJK_NK](separation, da.as<mgfb_shape>(), db.as<mgfb_shape>(), di.as<unsigned>() + count[k], m, dout.as<mgfb_contact>(),
                                dsep.as<float>(), dst.as<unsigned>(), dit.as<unsigned>(), dwork.as<EpaWork>() + (separation ? 0 : count[k]),
                                dcount.as<unsigned>() + k, dcount.as<unsigned>() + 32 + k, (int)std::min<unsigned>((unsigned)epa_grid, m), bin_stream);
        ctx->launches += separation ? 1 : 2;
    }
    for (int k = 0; k < NS; ++k) { cudaEventRecord(ctx->gjk_ev[k], ctx->gjk_streams[k]); cudaStreamWaitEvent(ctx->stream, ctx->gjk_ev[k], 0); }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        if (out) cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(mgfb_contact), cudaMemcpyDeviceToHost, ctx->stream);
        if (sep) cudaMemcpyAsync(sep, dsep.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(status, dst.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (epa_iters) cudaMemcpyAsync(epa_iters, dit.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) { ctx->err = std::string("gjk batch: ") + cudaGetErrorString(e); return done(MGFB_ERR_CUDA); }
    return done(MGFB_OK);
}

}  // namespace

extern "C" {
int32_t mgfb_convex_vertices_set(mgfb_ctx* ctx, const float* verts, uint32_t n) {
    if (!ctx || (n && !verts)) return fail(ctx, MGFB_ERR_INVALID_ARG, "null vertex array");
    if (n >= (1u << 24)) return fail(ctx, MGFB_ERR_INVALID_ARG, "the vertex pool holds fewer than 2^24 vertices");
    CU(cudaSetDevice(ctx->device));
    ctx->convex_n = 0;
    if (n) {
        TRY(ensure(ctx, ctx->convex_pool, (size_t)n * 12));
        CU(cudaMemcpyAsync(ctx->convex_pool.p, verts, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->convex_n = n;
    return MGFB_OK;
}
int32_t mgfb_gjk_batch(mgfb_ctx* ctx, const mgfb_shape* a, const mgfb_shape* b, uint32_t n, mgfb_contact* out, uint32_t* status, uint32_t* epa_iters) {
    return gjk_batch_impl(ctx, false, a, b, n, out, nullptr, status, epa_iters);
}
int32_t mgfb_separation_batch(mgfb_ctx* ctx, const mgfb_shape* a, const mgfb_shape* b, uint32_t n, float* separation, uint32_t* status) {
    return gjk_batch_impl(ctx, true, a, b, n, nullptr, separation, status, nullptr);
}
}
