// kernels.cuh -- the device kernels of the step path (see DESIGN.md section 4 for the list,
// the HBM bytes each moves and the roofline that bounds it).  All kernels are grid-stride
// over counts that live in device memory, so a whole step is enqueued without a single
// host round trip.
#pragma once
#include "tile.cuh"

namespace mgfb {

#define MGFB_THREADS 256

// ---------------------------------------------------------------- grid barrier
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// All CTAs of a cooperative launch.  `bar` is zero at kernel entry; phase counts barriers.
// One release-add + acquire-poll per CTA: the CTA barrier in front makes every thread's earlier
// writes happen-before thread 0's release, the one behind orders the acquire before every
// thread's later reads (PTX memory model: causality order is transitive across bar.sync).
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& phase) {
    __syncthreads();
    phase++;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned target = phase * gridDim.x;
        while (ld_acquire_u32(bar) < target) { }
    }
    __syncthreads();
}

// ---------------------------------------------------------------- integrate (+ AABB)
// physics.rs:262-269 complete_motion, :222-253 integrate, bounds.rs:60-68 swept AABB,
// world.rs:235-238 fat-box refresh -- one pass, 272 algorithmic bytes per body.
HD void collider_bounds(const Collider& k, V3* c, V3* r) {
    if (col_kind(k) == 0) {  // bounds.rs:170-177
        *c = f4v(k.p0); *r = mk3(k.p0.w, k.p0.w, k.p0.w);
    } else {                  // bounds.rs:179-188
        V3 d = f4v(k.p1);
        float rr = k.p0.w + len(d) * 0.5f;
        *c = f4v(k.p0) + d * 0.5f; *r = mk3(rr, rr, rr);
    }
}
// bounds.rs:60-68 + :113-130; returns false when the reference's assert!(r >= 0) would fire.
HD bool swept_bounds(const Collider& k, V3* oc, V3* orr) {
    V3 sc, sr; collider_bounds(k, &sc, &sr);
    V3 ec = sc + f4v(k.v);
    V3 lo = mk3(fminf(sc.x - sr.x, ec.x - sr.x), fminf(sc.y - sr.y, ec.y - sr.y), fminf(sc.z - sr.z, ec.z - sr.z));
    V3 hi = mk3(fmaxf(sc.x + sr.x, ec.x + sr.x), fmaxf(sc.y + sr.y, ec.y + sr.y), fmaxf(sc.z + sr.z, ec.z + sr.z));
    V3 r = (hi - lo) / 2.0f;
    *oc = (hi + lo) / 2.0f; *orr = r;
    return (r.x >= 0.0f) && (r.y >= 0.0f) && (r.z >= 0.0f);
}
HD bool box_contains_pt(V3 c, V3 r, V3 p) {  // collision.rs:114-120
    return fabsf(c.x - p.x) <= r.x && fabsf(c.y - p.y) <= r.y && fabsf(c.z - p.z) <= r.z;
}
HD bool box_overlaps(V3 ac, V3 ar, V3 bc, V3 br) {  // collision.rs:22-29 (closed)
    return fabsf(ac.x - bc.x) <= (ar.x + br.x) && fabsf(ac.y - bc.y) <= (ar.y + br.y) && fabsf(ac.z - bc.z) <= (ar.z + br.z);
}

template <bool COMPLETE, bool INTEGRATE, bool BOUNDS, bool TILED_XR = false>
__global__ void __launch_bounds__(MGFB_THREADS) k_integrate(BodyArrays B, unsigned n, float dt, float margin, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;  // poisoned: the host will resume from steps_done
    float my_fat = 0.0f, my_tight = 0.0f, my_xr = -3.0e38f;
    unsigned refreshed = 0;
    if (BOUNDS && blockIdx.x == 0 && threadIdx.x == 0) ctr->n_total = n;   // a tiled world adds its ghosts in k_ghost_recv
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Collider k = B.col[i];
        V3 xi = f4v(B.x[i]);
        if (COMPLETE) {  // x += collider.delta()
            xi = xi + f4v(k.v);
            B.x[i] = v4(xi, 0.0f);
        }
        if (INTEGRATE) {
            float4 q4 = B.q[i];
            Q4 qi = mkq(q4.x, mk3(q4.y, q4.z, q4.w));
            BodyVel bv = B.vel[i];
            V3 v = mk3(bv.a.x, bv.a.y, bv.a.z), w = mk3(bv.a.w, bv.b.x, bv.b.y);
            float inv_mass = bv.b.z;
            // q <- normalize(q + (Quat(0, w*dt) * 0.5) * q)                       physics.rs:226-227
            qi = qunit(qadd(qi, qmul(qscale(mkq(0.0f, w * dt), 0.5f), qi)));
            // I^-1 <- R * I^-1_body * R^T                                        physics.rs:231-232
            M3 R = m_from_q(qi);
            M3 Ib = mkm(f4v(B.imb[3 * i]), f4v(B.imb[3 * i + 1]), f4v(B.imb[3 * i + 2]));
            M3 I = mmul(mmul(R, Ib), mtrans(R));
            float4 f4 = B.force[i], t4 = B.torque[i];
            v = v + f4v(f4) * inv_mass * dt;                                    // physics.rs:236
            w = w + mmulv(I, f4v(t4)) * dt;                                     // physics.rs:240
            bv.a = make_float4(v.x, v.y, v.z, w.x);
            bv.b.x = w.y; bv.b.y = w.z;
            vel_set_inertia(bv, I);
            B.vel[i] = bv;
            B.q[i] = make_float4(qi.s, qi.v.x, qi.v.y, qi.v.z);
            // collider <- sweep(construct(x, q), v*dt)                          physics.rs:243-250
            if (col_kind(k) == 0) {
                k.p0 = v4(xi, k.p0.w);
            } else {
                V3 d = qrot(qi, mk3(0.0f, 1.0f, 0.0f) * k.v.w);                 // compound.rs:222-225
                k.p0 = v4(xi + (-d), k.p0.w);
                V3 d2 = d * 2.0f;
                k.p1 = make_float4(d2.x, d2.y, d2.z, k.p1.w);
            }
            V3 delta = v * dt;
            k.v = make_float4(delta.x, delta.y, delta.z, k.v.w);
            B.col[i] = k;
        }
        if (BOUNDS) {
            V3 tc, tr;
            if (!swept_bounds(k, &tc, &tr)) atomicOr(&ctr->nan_bounds, 1u);
            Box tb; tb.c = v4(tc, 0.0f); tb.r = v4(tr, 0.0f);
            B.tight[i] = tb;
            Box fb = B.fat[i];
            V3 fc = f4v(fb.c), fr = f4v(fb.r);
            // world.rs:235: if !stored.contains(&bounds)  (collision.rs:129-135)
            if (!(box_contains_pt(fc, fr, tc + tr) && box_contains_pt(fc, fr, tc + (-tr)))) {
                fr = tr + mk3(margin, margin, margin);   // bounds.rs:91-98
                fb.c = v4(tc, 0.0f); fb.r = v4(fr, 0.0f);
                B.fat[i] = fb;
                refreshed++;
                if (B.ref_flag) {
                    B.ref_flag[i] = 1;
                    unsigned k = atomicAdd(&ctr->n_ref, 1u);
                    if (k < B.ref_cap) B.ref_list[k] = i;
                }
            }
            my_fat = fmaxf(my_fat, fmaxf(fr.x, fmaxf(fr.y, fr.z)));
            my_tight = fmaxf(my_tight, fmaxf(tr.x, fmaxf(tr.y, tr.z)));
            my_xr = fmaxf(my_xr, fb.c.x + fr.x);
        }
    }
    if (BOUNDS) {
        for (int o = 16; o > 0; o >>= 1) {
            my_fat = fmaxf(my_fat, __shfl_xor_sync(0xffffffffu, my_fat, o));
            my_tight = fmaxf(my_tight, __shfl_xor_sync(0xffffffffu, my_tight, o));
            my_xr = fmaxf(my_xr, __shfl_xor_sync(0xffffffffu, my_xr, o));
            refreshed += __shfl_xor_sync(0xffffffffu, refreshed, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMax(&ctr->max_fat_bits, __float_as_uint(my_fat));
            atomicMax(&ctr->max_tight_bits, __float_as_uint(my_tight));
            if (TILED_XR) atomicMax(&ctr->xr_bits, ordered_bits(my_xr));
            if (refreshed) atomicAdd(&ctr->fat_refreshes, refreshed);
        }
    }
}

// ---------------------------------------------------------------- uniform hashed grid
struct GridView {
    unsigned* cell_count;   // [T]   (consumed by the fill)
    unsigned* cell_start;   // [T+1] exclusive scan of counts
    unsigned* ent_id;       // [E]
    unsigned long long* ent_key;  // [E] packed cell coordinates
    unsigned table_mask;    // T - 1
    unsigned ent_cap;
};
__device__ __forceinline__ int cell_coord(float x, float inv) {
    float f = floorf(x * inv);
    f = fminf(fmaxf(f, -1048576.0f), 1048575.0f);
    return (int)f;
}
__device__ __forceinline__ unsigned long long cell_key(int x, int y, int z) {
    return ((unsigned long long)(x & 0x1FFFFF) << 42) | ((unsigned long long)(y & 0x1FFFFF) << 21) | (unsigned long long)(z & 0x1FFFFF);
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser (bijective)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned cell_hash(unsigned long long key, unsigned mask) { return (unsigned)(mix64(key) >> 20) & mask; }
// Body grid: locality-preserving hash.  A 4x4x4 BLOCK of cells is hashed as one unit onto 64 consecutive buckets, so
// neighbouring cells sit in neighbouring buckets (same cache lines of cell_start, neighbouring runs of `ent`) and the
// grid's entry order is spatially coherent: the pair sweep walks the bodies in THAT order, which makes its gathers, the
// pair lists, the contact list and hence the rows of one colour spatially coherent too, whatever the body numbering.
// (table size is a power of two >= 4096)
__device__ __forceinline__ unsigned bcell_hash(int x, int y, int z, unsigned mask) {
    unsigned blk = (unsigned)(mix64(cell_key(x >> 2, y >> 2, z >> 2)) >> 20) & (mask >> 6);
    return (blk << 6) | (unsigned)((x & 3) | ((y & 3) << 2) | ((z & 3) << 4));
}
struct CellRange { int lo[3], hi[3]; };
// Cells covered by a stored box.  `slop` widens the range of INSERTED boxes by a few ulps so
// that a closed, rounded overlap test that passes always finds a shared cell.
__device__ __forceinline__ CellRange cell_range(V3 c, V3 r, float inv, bool slop) {
    CellRange cr;
    float cs[3] = {c.x, c.y, c.z}, rs[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = cs[a] - rs[a], hi = cs[a] + rs[a];
        if (slop) { float e = (fabsf(cs[a]) + rs[a]) * 2e-6f + 1e-30f; lo -= e; hi += e; }
        cr.lo[a] = cell_coord(lo, inv); cr.hi[a] = cell_coord(hi, inv);
    }
    return cr;
}
// Body grid: every body is binned ONCE, by the centre of its stored fat box, into cells of
// edge s >= max tight half extent + max fat half extent.  Then |c_i - c_j| <= r_i + R_j (the
// closed overlap test) implies the two centres are at most one cell apart on every axis, so a
// query looks at the 27 cells around the tight box's centre and no pair can be seen twice.
__device__ __forceinline__ float grid_inv_cell(const Counters* ctr) {
    float s = (__uint_as_float(ctr->max_fat_bits) + __uint_as_float(ctr->max_tight_bits)) * 1.001f;
    if (!(s > 0.0f) || !(s < 3.0e38f)) s = 1.0f;
    return 1.0f / s;
}
struct BodyGrid {
    unsigned* cell_count;   // [T]
    unsigned* cell_start;   // [T+1]
    float4* ent;            // [2n]: (fat.c.xyz, body index | kind << 31), (fat.r.xyz, global id)
    unsigned table_mask;
};
template <bool FILL>
__global__ void __launch_bounds__(MGFB_THREADS) k_bgrid_insert(const Box* __restrict__ fat, const Collider* __restrict__ col, const unsigned* __restrict__ gid,
                                                               BodyGrid G, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    float inv = grid_inv_cell(ctr);
    const unsigned n = ctr->n_total;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        Box b = fat[j];
        unsigned h = bcell_hash(cell_coord(b.c.x, inv), cell_coord(b.c.y, inv), cell_coord(b.c.z, inv), G.table_mask);
        if (!FILL) atomicAdd(&G.cell_count[h], 1u);
        else {
            unsigned pos = G.cell_start[h] + (atomicSub(&G.cell_count[h], 1u) - 1u);
            G.ent[2 * pos] = make_float4(b.c.x, b.c.y, b.c.z, __uint_as_float(j | ((unsigned)col_kind(col[j]) << 31)));
            G.ent[2 * pos + 1] = make_float4(b.r.x, b.r.y, b.r.z, __uint_as_float(gid[j]));
        }
    }
}
// Body-pair sweep, one warp per body i: lanes 0..26 take the 27 neighbouring cells.  Hits are
// staged in shared memory and appended with ONE global atomic per block per batch of 8 bodies
// (a single global cursor would otherwise serialise ~4 atomics per body at one L2 address).
// P = {(i, j): j < i, tight_i overlaps stored fat_j}  (world.rs:261-268), "<" on GLOBAL ids.  In
// a tiled world bodies >= n_own are ghosts: they take part as i or as j, never both.
#define BP_WARPS (MGFB_THREADS / 32)
#define BP_BUF 128
#define BP_PER 4     // bodies per warp between two CTA-wide reservations (3 CTA barriers per 32 bodies instead of per 8)
// buf entries: j | in_P << 29 | kind << 30 (bodies < 2^29), sub[e] = which of the warp's BP_PER bodies (body ids[sub]).
// base[k], k < 4: where this warp's kind-k entries start in list k; base[4]: where its entries start in the cached superset
// S (BUILD: every staged entry goes there, those with in_P set also to their kind's list).
template <bool CACHED>
__device__ __forceinline__ void bp_flush_warp(unsigned* buf, const unsigned char* sub, unsigned cnt, const unsigned* ids, const unsigned base[5], PairLists lists,
                                              unsigned cap, int2* s_list, unsigned s_cap, Counters* ctr, unsigned lane) {
    unsigned run[5] = {0, 0, 0, 0, 0};
    for (unsigned s0 = 0; s0 < cnt; s0 += 32) {
        const bool have = s0 + lane < cnt;
        unsigned e = have ? buf[s0 + lane] : 0u;
        const unsigned i = ids[have ? (unsigned)sub[s0 + lane] : 0u];
        const int kind_all = have ? (int)(e >> 30) : -1;
        const int kind = (have && ((e >> 29) & 1u)) ? kind_all : -1;
        const unsigned j = e & 0x1fffffffu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned m = __ballot_sync(0xffffffffu, kind == k);
            if (kind == k) {
                unsigned pos = base[k] + run[k] + __popc(m & ((1u << lane) - 1u));
                if (pos < cap) lists.p[k][pos] = make_int2((int)i, (int)j);
                else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);
            }
            run[k] += __popc(m);
        }
        if (CACHED && s_list) {
            unsigned m = __ballot_sync(0xffffffffu, have);
            if (have) {
                unsigned pos = base[4] + run[4] + __popc(m & ((1u << lane) - 1u));
                const unsigned ki = (unsigned)kind_all >> 1, kj = (unsigned)kind_all & 1u;
                if (pos < s_cap) s_list[pos] = make_int2((int)(i | (ki << 31)), (int)(j | (kj << 31)));
                else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);
            }
            run[4] += __popc(m);
        }
    }
}
}  // namespace mgfb
#include "bpcache.cuh"
namespace mgfb {
// CACHED = launched by the coherent broadphase (bpcache.cuh), whose k_bp_decide picked this step's path:
//   BP_REBUILD  the grid holds the own bodies binned by fat centre with cell edge s0; the sweep runs around the FAT centre and
//               stages every pair whose fat boxes overlap (-> S), marking those that also pass this step's exact test (-> lists);
//   BP_SWEEP    the plain sweep (tight box against stored fat boxes, ghosts in the grid), nothing kept;
//   BP_COHERENT nothing to do.
// !CACHED: the plain sweep.
template <bool CACHED>
__global__ void __launch_bounds__(MGFB_THREADS) k_body_pairs_warp(const Box* __restrict__ tight, const Collider* __restrict__ col, const unsigned* __restrict__ gid,
                                                                 unsigned n_own, BodyGrid G, PairLists lists, unsigned cap, Counters* ctr, BpView V, unsigned n_build) {
    __shared__ unsigned s_buf[BP_WARPS][BP_BUF];
    __shared__ unsigned char s_sub[BP_WARPS][BP_BUF];
    __shared__ unsigned s_cnt[BP_WARPS][5];     // per warp, per kind (+ S)
    __shared__ unsigned s_base[BP_WARPS][5];
    __shared__ unsigned s_ids[BP_WARPS][BP_PER];   // the warp's bodies of this batch
    if (ctr->overflow | ctr->nan_bounds) return;
    if (CACHED && V.st->mode == BP_COHERENT) return;
    const bool BUILD = CACHED && V.st->mode == BP_REBUILD;
    const float inv = CACHED ? 1.0f / V.st->s0 : grid_inv_cell(ctr);
    const unsigned n = BUILD ? n_build : ctr->n_total;
    int2* s_list = BUILD ? V.S[V.st->cur] : nullptr;
    unsigned* s_count = BUILD ? &V.st->s_count[V.st->cur] : nullptr;
    const unsigned s_cap = BUILD ? V.s_cap : 0u;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (unsigned batch = blockIdx.x * (BP_WARPS * BP_PER); batch < n; batch += gridDim.x * (BP_WARPS * BP_PER)) {   // uniform per block
        unsigned cnt = 0, kcnt[5] = {0, 0, 0, 0, 0};
        const unsigned* ids = s_ids[w];
        for (unsigned sub = 0; sub < BP_PER; ++sub) {
        // bodies are taken in GRID order (position in `ent` = bucket order: spatially coherent), not in index order
        const unsigned pos = batch + sub * BP_WARPS + w;
        if (pos < n) {   // (world.rs:256 skips body 0: it has no j < i)
            const float4 me0 = G.ent[2 * pos];
            const unsigned iw = __float_as_uint(me0.w);
            const unsigned i = iw & 0x7fffffffu;
            if (lane == 0) s_ids[w][sub] = i;
            __syncwarp();
            const unsigned gi = gid[i];
            const bool ghost_i = i >= n_own;
            Box tb = tight[i];
            V3 tc = f4v(tb.c), tr = f4v(tb.r);
            V3 fc = mk3(me0.x, me0.y, me0.z), fr = zero3();
            if (BUILD) { const float4 me1 = G.ent[2 * pos + 1]; fr = mk3(me1.x, me1.y, me1.z); }
            int ki = (int)(iw >> 31);
            const V3 qc = BUILD ? fc : tc;   // the query's centre
            int cx = cell_coord(qc.x, inv), cy = cell_coord(qc.y, inv), cz = cell_coord(qc.z, inv);
            unsigned e = 0, e1 = 0;
            {
                // lanes 0..26: the 27 cells around the centre.  Two of them may hash to the same bucket: only the
                // lowest such lane walks it.  Bodies of other cells in a bucket simply fail the exact overlap test
                // (an overlapping pair is never more than one cell apart), so no per-entry cell check is needed.
                unsigned h = 0xffffffffu - lane;   // lanes 27..31: unique dummies
                if (lane < 27) {
                    int x = cx + (int)(lane % 3) - 1, y = cy + (int)((lane / 3) % 3) - 1, z = cz + (int)(lane / 9) - 1;
                    h = bcell_hash(x, y, z, G.table_mask);
                }
                unsigned same = __match_any_sync(0xffffffffu, h);
                if (lane < 27 && (unsigned)__ffs((int)same) - 1u == lane) { e = G.cell_start[h]; e1 = G.cell_start[h + 1]; }
                e1 = min(e1, n); e = min(e, e1);   // (a grid that was not built must not turn into an endless walk)
            }
            // Flatten the (at most 27) bucket ranges: lane l owns [e, e1); an inclusive warp scan of the lengths
            // gives every candidate a global number t, and lane t%32 tests candidate t -- all 32 lanes busy
            // whatever the spread of bucket sizes (the longest bucket no longer sets the trip count).
            const unsigned len_l = e1 - e;
            unsigned incl = len_l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += y; }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            for (unsigned t0 = 0; t0 < total; t0 += 32) {
                const unsigned t = t0 + lane;
                // owner of candidate t = first lane whose inclusive prefix exceeds t (binary search over the 32 prefixes)
                unsigned lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    unsigned probe = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1));
                    if (probe <= t) lo += step;
                }
                const unsigned src_incl = __shfl_sync(0xffffffffu, incl, (int)lo);
                const unsigned src_len = __shfl_sync(0xffffffffu, len_l, (int)lo);
                const unsigned src_e = __shfl_sync(0xffffffffu, e, (int)lo);
                int kind = -1; unsigned j = 0; bool in_p = false;   // kind >= 0: staged
                if (t < total) {
                    const unsigned ee = src_e + (t - (src_incl - src_len));
                    float4 a = G.ent[2 * ee], b = G.ent[2 * ee + 1];
                    unsigned jw = __float_as_uint(a.w);
                    j = jw & 0x7fffffffu;
                    if (__float_as_uint(b.w) < gi && !(ghost_i && j >= n_own)) {
                        const V3 jc = mk3(a.x, a.y, a.z), jr = mk3(b.x, b.y, b.z);
                        in_p = box_overlaps(tc, tr, jc, jr);
                        if (in_p || (BUILD && box_overlaps_slop(fc, fr, jc, jr))) kind = ki * 2 + (int)(jw >> 31);
                    }
                }
                unsigned m = __ballot_sync(0xffffffffu, kind >= 0);
                if (!m) continue;
                unsigned nh = __popc(m);
                if (cnt + nh > BP_BUF) {   // staging full (a very dense blob): spill with a per-warp reservation
                    unsigned base[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        unsigned b0 = 0;
                        if (lane == 0 && kcnt[k]) b0 = atomicAdd(k < 4 ? &ctr->pairs[k] : s_count, kcnt[k]);
                        base[k] = __shfl_sync(0xffffffffu, b0, 0);
                        kcnt[k] = 0;
                    }
                    __syncwarp();
                    bp_flush_warp<CACHED>(s_buf[w], s_sub[w], cnt, ids, base, lists, cap, s_list, s_cap, ctr, lane);
                    __syncwarp();
                    cnt = 0;
                }
                if (kind >= 0) {
                    unsigned slot = cnt + __popc(m & ((1u << lane) - 1u));
                    s_buf[w][slot] = j | (in_p ? (1u << 29) : 0u) | ((unsigned)kind << 30); s_sub[w][slot] = (unsigned char)sub;
                }
                cnt += nh;
                if (BUILD) kcnt[4] += nh;
                // hits per kind: body i has ONE kind, so only kinds 2*ki and 2*ki+1 can occur
                unsigned np = __popc(__ballot_sync(0xffffffffu, in_p));
                unsigned n1 = __popc(__ballot_sync(0xffffffffu, in_p && kind == ki * 2 + 1));
#pragma unroll
                for (int k = 0; k < 4; ++k) kcnt[k] += k == ki * 2 + 1 ? n1 : (k == ki * 2 ? np - n1 : 0u);
            }
        }
        }
        if (lane < 5) s_cnt[w][lane] = kcnt[lane];
        __syncthreads();
        if (threadIdx.x < (BUILD ? 5 : 4)) {   // one reservation per kind (and for S) for the whole block
            unsigned tot = 0;
            for (int ww = 0; ww < BP_WARPS; ++ww) { s_base[ww][threadIdx.x] = tot; tot += s_cnt[ww][threadIdx.x]; }
            unsigned b0 = tot ? atomicAdd(threadIdx.x < 4 ? &ctr->pairs[threadIdx.x] : s_count, tot) : 0u;
            for (int ww = 0; ww < BP_WARPS; ++ww) s_base[ww][threadIdx.x] += b0;
        }
        __syncthreads();
        if (cnt) {
            unsigned base[5] = {s_base[w][0], s_base[w][1], s_base[w][2], s_base[w][3], BUILD ? s_base[w][4] : 0u};
            bp_flush_warp<CACHED>(s_buf[w], s_sub[w], cnt, ids, base, lists, cap, s_list, s_cap, ctr, lane);
        }
        __syncthreads();
    }
}

// pass 1: count entries per hashed cell; pass 2 (FILL): write them.
template <bool FILL>
__global__ void __launch_bounds__(MGFB_THREADS) k_grid_insert(const Box* __restrict__ boxes, unsigned n, GridView G, Counters* ctr, float fixed_inv) {
    if (ctr->overflow | ctr->nan_bounds) return;
    float inv = fixed_inv > 0.0f ? fixed_inv : grid_inv_cell(ctr);
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        Box b = boxes[j];
        CellRange cr = cell_range(f4v(b.c), f4v(b.r), inv, true);
        for (int x = cr.lo[0]; x <= cr.hi[0]; ++x)
            for (int y = cr.lo[1]; y <= cr.hi[1]; ++y)
                for (int z = cr.lo[2]; z <= cr.hi[2]; ++z) {
                    unsigned long long key = cell_key(x, y, z);
                    unsigned h = cell_hash(key, G.table_mask);
                    if (!FILL) {
                        atomicAdd(&G.cell_count[h], 1u);
                    } else {
                        unsigned pos = G.cell_start[h] + (atomicSub(&G.cell_count[h], 1u) - 1u);
                        if (pos < G.ent_cap) { G.ent_id[pos] = j; G.ent_key[pos] = key; }
                        else atomicOr(&ctr->overflow, (unsigned)OVF_GRID);
                    }
                }
    }
}

// Terrain candidates: Mesh::contacts' BVH query (mesh.rs:121) as a grid lookup over the
// triangles' AABBs (bounds.rs:137-152, precomputed).
struct TerrainView {
    const float4* verts;   // xyz
    const uint4* faces;    // a, b, c, unused
    const Box* boxes;      // per face, mesh-local
    GridView G;
    float inv_cell;
    unsigned nfaces;
    float4 x;              // Mesh.x
};
__global__ void __launch_bounds__(MGFB_THREADS) k_terrain_pairs(const Box* __restrict__ tight, const Collider* __restrict__ col, unsigned n,
                                                               TerrainView T, PairLists lists, unsigned cap, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Box tb = tight[i];
        V3 tc = f4v(tb.c) + (-f4v(T.x));  // rhs.bounds() - self.x  (mesh.rs:121)
        V3 tr = f4v(tb.r);
        CellRange qi = cell_range(tc, tr, T.inv_cell, false);
        int ki = col_kind(col[i]);
        for (int x = qi.lo[0]; x <= qi.hi[0]; ++x)
            for (int y = qi.lo[1]; y <= qi.hi[1]; ++y)
                for (int z = qi.lo[2]; z <= qi.hi[2]; ++z) {
                    unsigned long long key = cell_key(x, y, z);
                    unsigned h = cell_hash(key, T.G.table_mask);
                    unsigned e0 = T.G.cell_start[h], e1 = T.G.cell_start[h + 1];
                    for (unsigned e = e0; e < e1; ++e) {
                        if (T.G.ent_key[e] != key) continue;
                        unsigned f = T.G.ent_id[e];
                        Box fb = T.boxes[f];
                        V3 fc = f4v(fb.c), fr = f4v(fb.r);
                        if (!box_overlaps(tc, tr, fc, fr)) continue;
                        CellRange qj = cell_range(fc, fr, T.inv_cell, true);
                        if (x != max(qi.lo[0], qj.lo[0]) || y != max(qi.lo[1], qj.lo[1]) || z != max(qi.lo[2], qj.lo[2])) continue;
                        unsigned pos = atomicAdd(&ctr->tpairs[ki], 1u);
                        if (pos < cap) lists.p[ki][pos] = make_int2((int)i, (int)f);
                        else atomicOr(&ctr->overflow, (unsigned)OVF_TPAIRS);
                    }
                }
    }
}
// per-face AABB, bounds.rs:137-152 (centroid-centred box, not the tight one)
__global__ void k_face_boxes(const float4* verts, const uint4* faces, unsigned nfaces, Box* out, unsigned* max_bits) {
    unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    uint4 fc = faces[f];
    V3 a = f4v(verts[fc.x]), b = f4v(verts[fc.y]), c = f4v(verts[fc.z]);
    V3 ctr = (a + b + c) / 3.0f;
    float d0 = fmaxf(fabsf(a.x - ctr.x), fmaxf(fabsf(b.x - ctr.x), fabsf(c.x - ctr.x)));
    float d1 = fmaxf(fabsf(a.y - ctr.y), fmaxf(fabsf(b.y - ctr.y), fabsf(c.y - ctr.y)));
    float d2 = fmaxf(fabsf(a.z - ctr.z), fmaxf(fabsf(b.z - ctr.z), fabsf(c.z - ctr.z)));
    Box bx; bx.c = v4(ctr, 0.0f); bx.r = make_float4(d0, d1, d2, 0.0f);
    out[f] = bx;
    atomicMax(max_bits, __float_as_uint(2.0f * fmaxf(d0, fmaxf(d1, d2))));
}

// ---------------------------------------------------------------- narrowphase
// What the chain colouring needs of a fresh constraint (key, body degrees, cleared links: the loop body of k_inc_count), done
// by the thread that emits the contact: one pass over the contact list and one launch less per step.  NULL key: not fused.
struct EmitHook {
    const unsigned* gid; const float4* x; const Collider* col;   // for the key (chain_key)
    unsigned long long* key; unsigned* deg; unsigned long long* inbox; unsigned* next; int* group; unsigned cap;
};
__device__ __forceinline__ void emit_hook(const EmitHook& H, unsigned k, int a, int b, unsigned face, unsigned sub);
__device__ __forceinline__ void emit_contact(const ContactList& L, unsigned cap, Counters* ctr, int a, int b, unsigned face,
                                             unsigned sub, V3 la, V3 lb, V3 n, float t, const EmitHook& H) {
    unsigned k = atomicAdd(&ctr->contacts, 1u);
    if (k >= cap) { atomicOr(&ctr->overflow, (unsigned)OVF_CONTACTS); return; }
    L.a[k] = a; L.b[k] = b; L.face[k] = face; L.sub[k] = sub;
    L.la[k] = v4(la, n.x); L.lb[k] = v4(lb, n.y); L.nt[k] = make_float4(n.z, t, 0.0f, 0.0f);
    if (H.key) emit_hook(H, k, a, b, face, sub);
}

// Body i (receiver, kind KI) vs body j (argument, kind KJ): compound.rs:192-207 resolved per
// SURVEY Appendix B to ga.contacts(&Moving(gb, vb - va)) + frame shift (collision.rs:1387).
template <int KI, int KJ>
__device__ __forceinline__ bool body_pair_contact(const Collider& A, const Collider& Bc, Hit* h, V3* la, V3* lb) {
    V3 va = f4v(A.v), vb = f4v(Bc.v);
    V3 vrel = vb - va;
    bool ok;
    if (KI == 0 && KJ == 0) ok = sphere_msphere(col_sphere(A), col_sphere(Bc), vrel, h);
    else if (KI == 1 && KJ == 0) ok = capsule_msphere(col_capsule(A), col_sphere(Bc), vrel, h);
    else if (KI == 0 && KJ == 1) ok = sphere_mcapsule(col_sphere(A), col_capsule(Bc), vrel, h);
    else ok = capsule_mcapsule(col_capsule(A), col_capsule(Bc), vrel, h);
    if (!ok) return false;
    h->a = h->a + va * h->t;
    h->b = h->b + va * h->t;
    *la = h->a + (-(col_center(A) + va * h->t));
    *lb = h->b + (-(col_center(Bc) + vb * h->t));
    return true;
}
template <int KI, int KJ>
__global__ void __launch_bounds__(MGFB_THREADS) k_narrow_bodies(const Collider* __restrict__ col, const int2* __restrict__ pairs,
                                                               ContactList L, unsigned cap, Counters* ctr, EmitHook H) {
    if (ctr->overflow | ctr->nan_bounds) return;
    unsigned np = ctr->pairs[KI * 2 + KJ];
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        int2 ij = pairs[p];
        Collider A = col[ij.x], Bc = col[ij.y];
        Hit h; V3 la, lb;
        if (!body_pair_contact<KI, KJ>(A, Bc, &h, &la, &lb)) continue;
        // ContactPruner with one contact -> Manifold::from(pruner): normal = (0 + n) / 1  (manifold.rs:135-140)
        V3 n = (zero3() + h.n) / 1.0f;
        emit_contact(L, cap, ctr, ij.x, ij.y, 0u, 0u, la, lb, n, h.t, H);
    }
}
// Body i vs terrain face f: mesh.rs:119-137 + collision.rs:1490-1506 + compound.rs:179-190.
// Leaf hit has a on the triangle, b on the body, n = triangle normal; the delivered
// LocalContact has local_a = body point - body centre(t), local_b = triangle point - mesh.x,
// normal = -n_tri.
template <int KI>
__device__ __forceinline__ int body_tri_contacts(const Collider& A, const Tri& tri, V3 mesh_x, Hit out[2], V3 la[2], V3 lb[2]) {
    Hits hs; hs.n = 0;
    V3 v = f4v(A.v);
    if (KI == 0) { Hit h; if (poly_msphere(tri, col_sphere(A), v, &h)) push(hs, h); }
    else hs = poly_mcapsule(tri, col_capsule(A), v);
    for (int k = 0; k < hs.n; ++k) {
        Hit c = hs.h[k];
        V3 a_c = col_center(A) + v * c.t;
        la[k] = c.b + (-a_c);
        lb[k] = c.a + (-mesh_x);
        out[k] = flip(c);
    }
    return hs.n;
}
template <int KI>
__global__ void __launch_bounds__(MGFB_THREADS) k_narrow_terrain(const Collider* __restrict__ col, const int2* __restrict__ pairs,
                                                                TerrainView T, ContactList L, unsigned cap, Counters* ctr, EmitHook H) {
    if (ctr->overflow | ctr->nan_bounds) return;
    unsigned np = ctr->tpairs[KI];
    V3 mx = f4v(T.x);
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        int2 bf = pairs[p];
        Collider A = col[bf.x];
        uint4 fc = T.faces[bf.y];
        Tri tri; tri.a = f4v(T.verts[fc.x]) + mx; tri.b = f4v(T.verts[fc.y]) + mx; tri.c = f4v(T.verts[fc.z]) + mx;
        Hit hs[2]; V3 la[2], lb[2];
        int nh = body_tri_contacts<KI>(A, tri, mx, hs, la, lb);
        for (int k = 0; k < nh; ++k) {
            emit_contact(L, cap, ctr, bf.x, -1, (unsigned)bf.y, (unsigned)k, la[k], lb[k], hs[k].n, hs[k].t, H);
            atomicAdd(&ctr->tcontacts, 1u);
        }
    }
}

// ---------------------------------------------------------------- constraint ordering
// Jones-Plassmann greedy edge colouring (MGFB_ORDER_COLOURED) or order-preserving level
// scheduling (MGFB_ORDER_AS_GIVEN) of the constraint graph, in one cooperative kernel.
struct OrderView {
    const int* a; const int* b;        // constraint endpoints (b < 0: static)
    const uint32_t* face; const uint32_t* sub;  // identity for the priority hash (may be NULL)
    unsigned long long* body_best;     // [nbodies]
    unsigned long long* body_mask;     // [nbodies] colours used (coloured mode)
    unsigned* body_last;               // [nbodies] last level / highest overflow colour
    int* group;                        // [m] out: colour / level, -1 = unassigned
    unsigned* group_count;             // [gcap]
    unsigned gcap;
    const unsigned* gid;               // global body ids: priorities do not depend on the tiling
    const float4* x; const Collider* col;   // body positions / colliders for the geometric priority classes (NULL: hash only)
    unsigned n_own;                    // tiled world: bodies >= n_own are ghosts; 0xffffffff otherwise
};
#define TILE_INTERIOR_COLOURS 32
__device__ __forceinline__ unsigned long long order_key(const OrderView& O, unsigned k, bool as_given) {
    if (as_given) return ~(unsigned long long)k;             // earlier in the list = higher priority
    if (!O.face) { unsigned long long key = mix64((unsigned long long)k + 1ULL); return key ? key : 1ULL; }
    int a = O.a[k], b = O.b[k];
    unsigned lo = b >= 0 ? O.gid[b] : (0x80000000u | ((O.face ? O.face[k] : 0u) << 1) | (O.sub ? O.sub[k] : 0u));
    unsigned long long key = mix64(((unsigned long long)O.gid[a] << 32) | lo);
    return key ? key : 1ULL;
}
__global__ void __launch_bounds__(MGFB_THREADS) k_order(OrderView O, const unsigned* m_ptr, unsigned m_host, bool as_given, Counters* ctr,
                                                       unsigned fallback_only, unsigned nbodies) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (fallback_only && !ctr->colour_fallback) return;   // k_colour_df coloured this step
    const unsigned m = m_ptr ? *m_ptr : m_host;
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    unsigned phase = 0;
    if (fallback_only) {   // k_colour_df ran out of its 63 colours: start over from clean scratch
        for (unsigned i = tid; i < nbodies; i += nth) { O.body_best[i] = 0ULL; O.body_mask[i] = 0ULL; O.body_last[i] = 0u; }
        for (unsigned g = tid; g < O.gcap; g += nth) O.group_count[g] = 0u;
    }
    for (unsigned k = tid; k < m; k += nth) __stcg(&O.group[k], -1);
    if (tid == 0) { ctr->remaining = m; ctr->ngroups = 0; ctr->rounds = 0; }
    grid_barrier(&ctr->bar, phase);
    for (unsigned round = 0; round < 100000u; ++round) {
        if (__ldcg(&ctr->remaining) == 0) break;
        // propose
        for (unsigned k = tid; k < m; k += nth) {
            if (__ldcg(&O.group[k]) >= 0) continue;
            unsigned long long key = order_key(O, k, as_given);
            atomicMax(&O.body_best[O.a[k]], key);
            int b = O.b[k];
            if (b >= 0) atomicMax(&O.body_best[b], key);
        }
        grid_barrier(&ctr->bar, phase);
        // commit
        unsigned done = 0;
        for (unsigned k = tid; k < m; k += nth) {
            if (__ldcg(&O.group[k]) >= 0) continue;
            unsigned long long key = order_key(O, k, as_given);
            int a = O.a[k], b = O.b[k];
            if (__ldcg(&O.body_best[a]) != key) continue;
            if (b >= 0 && __ldcg(&O.body_best[b]) != key) continue;
            unsigned g;
            if (as_given) {
                unsigned la = __ldcg(&O.body_last[a]);
                unsigned lb = b >= 0 ? __ldcg(&O.body_last[b]) : 0u;
                g = max(la, lb);             // level index = (1 + max) - 1
                __stcg(&O.body_last[a], g + 1);
                if (b >= 0) __stcg(&O.body_last[b], g + 1);
            } else {
                unsigned long long ma = __ldcg(&O.body_mask[a]);
                unsigned long long mb = b >= 0 ? __ldcg(&O.body_mask[b]) : 0ULL;
                unsigned long long fr = ~(ma | mb);
                if (O.n_own != 0xffffffffu) {
                    // tiled world: interior constraints take colours 0..31, boundary ones (a ghost
                    // endpoint) 32..63, so every iteration runs interior | exchange | boundary
                    bool boundary = (unsigned)a >= O.n_own || (b >= 0 && (unsigned)b >= O.n_own);
                    fr &= boundary ? 0xFFFFFFFF00000000ULL : 0x00000000FFFFFFFFULL;
                    if (!fr) { atomicOr(&ctr->overflow, (unsigned)OVF_GROUPS); fr = boundary ? (1ULL << 63) : (1ULL << 31); }
                }
                if (fr) {
                    g = (unsigned)__ffsll((long long)fr) - 1u;
                    __stcg(&O.body_mask[a], ma | (1ULL << g));
                    if (b >= 0) __stcg(&O.body_mask[b], mb | (1ULL << g));
                } else {  // more than 64 colours around one body: stack further colours above
                    unsigned la = max(__ldcg(&O.body_last[a]), 63u);
                    unsigned lb = b >= 0 ? max(__ldcg(&O.body_last[b]), 63u) : 63u;
                    g = max(la, lb) + 1u;
                    __stcg(&O.body_last[a], g);
                    if (b >= 0) __stcg(&O.body_last[b], g);
                }
            }
            __stcg(&O.group[k], (int)g);
            __stcg(&O.body_best[a], 0ULL);
            if (b >= 0) __stcg(&O.body_best[b], 0ULL);
            if (g < O.gcap) atomicAdd(&O.group_count[g], 1u);
            else atomicOr(&ctr->overflow, (unsigned)OVF_GROUPS);
            atomicMax(&ctr->ngroups, g + 1u);
            done++;
        }
        for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
        if ((threadIdx.x & 31) == 0 && done) atomicSub(&ctr->remaining, done);
        if (tid == 0) ctr->rounds = round + 1;
        grid_barrier(&ctr->bar, phase);
    }
}
// ---------------------------------------------------------------- constraint colouring, dataflow
// The same greedy edge colouring as k_order's coloured mode -- constraints in descending key
// order, each taking the lowest colour free at both its bodies -- without grid barriers.  The
// constraints of a body, sorted by key, form a chain; the body's colour mask travels down the
// chain: a constraint that has received the masks of BOTH its bodies picks its colour and
// pushes mask | colour to its successor on each chain (one 8-byte word: valid bit + 63 colour
// bits, so flag and payload are one atomic store).  Steps: count constraints per body -> scan
// -> fill the per-body lists (CSR) -> sort each short list by key and link it -> colour.
// Priority of a constraint in the chain colouring: a geometric CLASS in the top bits, the identity hash below.
// Greedy colouring takes constraints class by class: body-body contacts by the dominant axis of the centre
// offset and the parity of round(x_lower / diameter_lower) along it (the lower body's place in a row of touching
// bodies), terrain contacts last.  In a stacked or
// settled pile the contacts of one class are (nearly) disjoint -- along a row of touching bodies they alternate
// parity -- so a class costs about one colour: 7 colours instead of 10 at C2 (the maximum degree is 7), and the
// chains the colours travel down are 7-10 links deep instead of ~26.  On irregular piles it is neutral (+-1 colour).
// Depends only on the two bodies' state, which ghosts copy exactly: the same key on every tile.
__device__ __forceinline__ unsigned long long chain_key_geo(unsigned long long h, int a, int b, const float4* x, const Collider* col) {
    unsigned cls = 6u;
    if (b >= 0) {
        float4 xa = x[a], xb = x[b];
        float dx = xa.x - xb.x, dy = xa.y - xb.y, dz = xa.z - xb.z;
        float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
        unsigned axis = (ay > ax) ? ((az > ay) ? 2u : 1u) : ((az > ax) ? 2u : 0u);
        // parity of the LOWER body's place in a row of touching bodies along that axis: round(x_lo / diameter_lo)
        float ca = axis == 0u ? xa.x : (axis == 1u ? xa.y : xa.z), cb = axis == 0u ? xb.x : (axis == 1u ? xb.y : xb.z);
        const bool lo_a = ca < cb;
        float q = floorf((lo_a ? ca : cb) / (2.0f * (lo_a ? col[a].p0.w : col[b].p0.w)) + 0.5f);
        unsigned par = (fabsf(q) < 1.0e9f) ? ((unsigned)(long long)q & 1u) : 0u;
        cls = axis * 2u + par;
    }
    return ((unsigned long long)(7u - cls) << 61) | (h >> 3);
}
__device__ __forceinline__ unsigned long long chain_key(const OrderView& O, unsigned k) {
    unsigned long long h = order_key(O, k, false);
    if (!O.x) return h;
    return chain_key_geo(h, O.a[k], O.b[k], O.x, O.col);
}
// the step path's constraint, initialised by the thread that emitted its contact (same values as k_inc_count computes)
__device__ __forceinline__ void emit_hook(const EmitHook& H, unsigned k, int a, int b, unsigned face, unsigned sub) {
    unsigned lo = b >= 0 ? H.gid[b] : (0x80000000u | (face << 1) | sub);
    unsigned long long h = mix64(((unsigned long long)H.gid[a] << 32) | lo);
    if (!h) h = 1ULL;
    H.key[k] = chain_key_geo(h, a, b, H.x, H.col);
    atomicAdd(&H.deg[a], 1u);
    if (b >= 0) atomicAdd(&H.deg[b], 1u);
    H.inbox[k] = 0ULL; H.inbox[H.cap + k] = 0ULL;
    H.next[k] = 0xffffffffu; H.next[H.cap + k] = 0xffffffffu;
    H.group[k] = -1;
}
struct ColourView {
    unsigned long long* key;     // [m] priority (bijective hash of the constraint's identity)
    unsigned* deg;               // [nbodies] constraints per body, consumed by the fill
    const unsigned* body_start;  // [nbodies + 1]
    unsigned* csr;               // [sum deg] constraints of every body, sorted by key descending
    unsigned* next;              // [2][cap] successor on the chain of body a / body b: k << 1 | side, or 0xffffffff
    unsigned long long* inbox;   // [2][cap] mask so far at body a / body b, bit 63 = valid
    unsigned cap;
};
#define COLOUR_VALID (1ULL << 63)
__global__ void __launch_bounds__(MGFB_THREADS) k_inc_count(OrderView O, ColourView V, const unsigned* m_ptr, unsigned m_host, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    const unsigned m = m_ptr ? *m_ptr : m_host;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        V.key[k] = chain_key(O, k);
        atomicAdd(&V.deg[O.a[k]], 1u);
        int b = O.b[k];
        if (b >= 0) atomicAdd(&V.deg[b], 1u);
        V.inbox[k] = 0ULL; V.inbox[V.cap + k] = 0ULL;
        V.next[k] = 0xffffffffu; V.next[V.cap + k] = 0xffffffffu;
        O.group[k] = -1;
    }
}
__global__ void __launch_bounds__(MGFB_THREADS) k_inc_fill(OrderView O, ColourView V, const unsigned* m_ptr, unsigned m_host, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    const unsigned m = m_ptr ? *m_ptr : m_host;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        int a = O.a[k], b = O.b[k];
        V.csr[V.body_start[a] + (atomicSub(&V.deg[a], 1u) - 1u)] = k;
        if (b >= 0) V.csr[V.body_start[b] + (atomicSub(&V.deg[b], 1u) - 1u)] = k;
    }
}
// One thread per body: sort its constraints by key (descending), link the chain, seed the head's inbox.
__global__ void __launch_bounds__(MGFB_THREADS) k_inc_sort(OrderView O, ColourView V, unsigned nbodies, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nbodies; i += gridDim.x * blockDim.x) {
        const unsigned s0 = V.body_start[i], d = V.body_start[i + 1] - s0;
        if (d == 0) continue;
        unsigned* L = V.csr + s0;
        if (d <= 8) {   // the common case: rank by counting, all in registers (keys are unique and non-zero)
            unsigned kk[8]; unsigned long long ky[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { kk[j] = 0u; ky[j] = 0ULL; if ((unsigned)j < d) { kk[j] = L[j]; ky[j] = V.key[kk[j]]; } }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((unsigned)j < d) {
                    unsigned rank = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) rank += (ky[i] > ky[j] || (ky[i] == ky[j] && kk[i] > kk[j])) ? 1u : 0u;   // (61 hash bits can tie: break by index)
                    L[rank] = kk[j];
                }
            }
        } else if (d <= 16) {   // a settled pile: up to 12-13 contacts per sphere.  Same rank counting, still all in registers
            unsigned kk[16]; unsigned long long ky[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { kk[j] = 0u; ky[j] = 0ULL; if ((unsigned)j < d) { kk[j] = L[j]; ky[j] = V.key[kk[j]]; } }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if ((unsigned)j < d) {
                    unsigned rank = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) rank += (ky[i] > ky[j] || (ky[i] == ky[j] && kk[i] > kk[j])) ? 1u : 0u;
                    L[rank] = kk[j];
                }
            }
        } else {   // rare (a body touching > 16 others): in place
            for (unsigned j = 1; j < d; ++j) {
                unsigned k = L[j]; unsigned long long y = V.key[k]; unsigned p = j;
                while (p > 0 && (V.key[L[p - 1]] < y || (V.key[L[p - 1]] == y && L[p - 1] < k))) { L[p] = L[p - 1]; --p; }
                L[p] = k;
            }
        }
        unsigned k = L[0]; unsigned side = O.a[k] == (int)i ? 0u : 1u;
        V.inbox[side * V.cap + k] = COLOUR_VALID;
        for (unsigned j = 0; j + 1 < d; ++j) {
            unsigned kn = L[j + 1]; unsigned sn = O.a[kn] == (int)i ? 0u : 1u;
            V.next[side * V.cap + k] = (kn << 1) | sn;
            k = kn; side = sn;
        }
    }
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Persistent, all threads co-resident (cooperative launch).  A thread owns constraints tid, tid+nth, ...
// and keeps sweeping the ones not yet coloured; it never blocks on one, so there is no ordering
// requirement: the uncoloured constraint with the largest key always has both masks.
#ifndef COLOUR_CACHED
#define COLOUR_CACHED 4   // constraints per thread whose endpoints and links live in registers (4 x 4 CTAs/SM measured best)
#endif
#ifndef COLOUR_CTAS
#define COLOUR_CTAS 4     // resident CTAs per SM the register budget is sized for
#endif
__device__ __forceinline__ unsigned colour_one(const OrderView& O, const ColourView& V, unsigned k, int a, int b, unsigned na, unsigned nb,
                                               unsigned long long ma, unsigned long long mb, Counters* ctr) {
    unsigned long long fr = ~(ma | mb);   // bit 63 is never free
    if (O.n_own != 0xffffffffu) {
        // tiled world: interior constraints take colours 0..31, boundary ones (a ghost endpoint) 32..62
        bool boundary = (unsigned)a >= O.n_own || (b >= 0 && (unsigned)b >= O.n_own);
        fr &= boundary ? 0x7FFFFFFF00000000ULL : 0x00000000FFFFFFFFULL;
    }
    unsigned g;
    if (fr) g = (unsigned)__ffsll((long long)fr) - 1u;
    else { atomicOr(&ctr->colour_fallback, 1u); g = 62u; }   // out of colours: k_order redoes the whole colouring
    const unsigned long long bit = 1ULL << g;
    if (na != 0xffffffffu) st_relaxed_u64(V.inbox + (na & 1u) * V.cap + (na >> 1), ma | bit);
    else O.body_mask[a] = (ma | bit) & ~COLOUR_VALID;
    if (b >= 0) {
        if (nb != 0xffffffffu) st_relaxed_u64(V.inbox + (nb & 1u) * V.cap + (nb >> 1), mb | bit);
        else O.body_mask[b] = (mb | bit) & ~COLOUR_VALID;
    }
    __stcg(&O.group[k], (int)g);
    return g;
}
// Persistent, all threads co-resident (cooperative launch).  A thread owns constraints tid, tid+nth, ...
// and keeps sweeping the ones not yet coloured; it never blocks on one, so there is no ordering
// requirement: the uncoloured constraint with the largest key always has both masks.  One sweep
// over the (register-cached) first COLOUR_CACHED constraints is a single batch of independent
// 8-byte L2 loads, so a colour travels one link of a chain per L2 round trip.
__global__ void __launch_bounds__(MGFB_THREADS, COLOUR_CTAS) k_colour_df(OrderView O, ColourView V, const unsigned* m_ptr, unsigned m_host, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    const unsigned m = m_ptr ? *m_ptr : m_host;
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const unsigned wm = __ballot_sync(0xffffffffu, tid < m);   // the lanes of this warp that own constraints
    if (tid >= m) return;
    const unsigned mine = (m - tid + nth - 1) / nth;
    int ca[COLOUR_CACHED], cb[COLOUR_CACHED]; unsigned cna[COLOUR_CACHED], cnb[COLOUR_CACHED];
    unsigned pending = 0;   // bit j: cached constraint j not yet coloured
#pragma unroll
    for (int j = 0; j < COLOUR_CACHED; ++j) {
        unsigned k = tid + (unsigned)j * nth;
        ca[j] = 0; cb[j] = -1; cna[j] = cnb[j] = 0xffffffffu;
        if (k < m) { ca[j] = O.a[k]; cb[j] = O.b[k]; cna[j] = V.next[k]; cnb[j] = V.next[V.cap + k]; pending |= 1u << j; }
    }
    unsigned left = mine, sweeps = 0, gmax = 0;
    unsigned hist_g[COLOUR_CACHED];
    while (left) {
        unsigned long long ma[COLOUR_CACHED], mb[COLOUR_CACHED];
#pragma unroll
        for (int j = 0; j < COLOUR_CACHED; ++j) {
            unsigned k = tid + (unsigned)j * nth;
            ma[j] = mb[j] = 0ULL;
            if ((pending >> j) & 1u) {
                ma[j] = ld_relaxed_u64(V.inbox + k);
                mb[j] = cb[j] >= 0 ? ld_relaxed_u64(V.inbox + V.cap + k) : COLOUR_VALID;
            }
        }
#pragma unroll
        for (int j = 0; j < COLOUR_CACHED; ++j) {
            if (((pending >> j) & 1u) && (ma[j] & mb[j] & COLOUR_VALID)) {
                unsigned k = tid + (unsigned)j * nth;
                unsigned g = colour_one(O, V, k, ca[j], cb[j], cna[j], cnb[j], ma[j], mb[j], ctr);
                hist_g[j] = g; gmax = max(gmax, g + 1u);
                pending &= ~(1u << j); --left;
            }
        }
        // beyond the cached ones (more than COLOUR_CACHED x resident threads constraints): same sweep from global memory
        for (unsigned k = tid + COLOUR_CACHED * nth; k < m; k += nth) {
            if (__ldcg(&O.group[k]) >= 0) continue;
            const int a = O.a[k], b = O.b[k];
            unsigned long long xa = ld_relaxed_u64(V.inbox + k);
            unsigned long long xb = b >= 0 ? ld_relaxed_u64(V.inbox + V.cap + k) : COLOUR_VALID;
            if (!(xa & xb & COLOUR_VALID)) continue;
            unsigned g = colour_one(O, V, k, a, b, V.next[k], V.next[V.cap + k], xa, xb, ctr);
            atomicAdd(&O.group_count[g], 1u); gmax = max(gmax, g + 1u);
            --left;
        }
        ++sweeps;
    }
    // colour histogram: cached constraints are counted here, warp-aggregated per colour
    __syncwarp(wm);
#pragma unroll
    for (int j = 0; j < COLOUR_CACHED; ++j) {
        bool have = tid + (unsigned)j * nth < m;
        unsigned act = __ballot_sync(wm, have);
        if (have) {
            unsigned peers = __match_any_sync(act, hist_g[j]);
            if ((threadIdx.x & 31u) == (unsigned)__ffs((int)peers) - 1u) atomicAdd(&O.group_count[hist_g[j]], (unsigned)__popc(peers));
        }
    }
    gmax = __reduce_max_sync(wm, gmax); sweeps = __reduce_max_sync(wm, sweeps);
    if ((threadIdx.x & 31u) == (unsigned)__ffs((int)wm) - 1u) { atomicMax(&ctr->ngroups, gmax); atomicMax(&ctr->rounds, sweeps); }
}

// Exclusive scan of group_count[0..ngroups) into group_start (single block, chunked), then the
// compact list of NON-EMPTY groups: phase p of the solver covers rows [phase_start[p],
// phase_start[p+1]).  `split` = first boundary colour of a tiled world (else >= ngroups).
__device__ __forceinline__ unsigned block_excl_scan_1024(unsigned v, unsigned* warp_sums, unsigned* total) {
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned w = warp_sums[threadIdx.x], ws = w;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, ws, o); if (threadIdx.x >= o) ws += y; }
        warp_sums[threadIdx.x] = ws - w;
        if (threadIdx.x == 31) *total = ws;
    }
    __syncthreads();
    unsigned excl = warp_sums[threadIdx.x >> 5] + x - v;
    __syncthreads();
    return excl;
}
__global__ void __launch_bounds__(1024) k_group_scan(const unsigned* count, unsigned* start, unsigned* phase_start, Counters* ctr, unsigned gcap,
                                                     unsigned split) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned tot;
    unsigned n = min(ctr->ngroups, gcap);
    if (threadIdx.x == 0) ctr->bar = 0;   // re-arm the grid barrier for the solve kernel
    unsigned carry = 0, pcarry = 0, pint = 0, int_rows = 0;   // every thread keeps the running totals
    for (unsigned base = 0; base < n; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < n ? count[i] : 0u;
        unsigned f = v > 0u ? 1u : 0u;
        unsigned excl = carry + block_excl_scan_1024(v, warp_sums, &tot);
        carry += tot;
        unsigned pos = pcarry + block_excl_scan_1024(f, warp_sums, &tot);
        pcarry += tot;
        if (i < n) start[i] = excl;
        if (f) phase_start[pos] = excl;
        // interior = groups below `split`
        (void)block_excl_scan_1024((f && i < split) ? 1u : 0u, warp_sums, &tot);
        pint += tot;
        (void)block_excl_scan_1024(i < split ? v : 0u, warp_sums, &tot);
        int_rows += tot;
    }
    if (threadIdx.x == 0) {
        start[n] = carry;
        phase_start[pcarry] = carry;
        ctr->n_phases = pcarry; ctr->n_int_phases = pint; ctr->n_int_rows = int_rows;
    }
}
// row = group_start[g] + (slot within the group); perm[row] = k.  Slots are claimed per CTA: a shared-memory
// histogram of the tile's colours, one global atomic per colour per tile (a handful of colours hold all
// constraints, so per-constraint global atomics would serialise on ~10 addresses).
__global__ void __launch_bounds__(MGFB_THREADS) k_scatter_rows(const int* __restrict__ group, unsigned* group_count, const unsigned* __restrict__ group_start,
                                                              unsigned* perm, const unsigned* m_ptr, unsigned m_host, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    __shared__ unsigned s_cnt[64], s_base[64];
    const unsigned m = m_ptr ? *m_ptr : m_host;
    for (unsigned base = blockIdx.x * blockDim.x; base < m; base += gridDim.x * blockDim.x) {
        if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
        unsigned k = base + threadIdx.x;
        int g = k < m ? group[k] : -1;
        unsigned local = 0;
        if (g >= 0 && g < 64) local = atomicAdd(&s_cnt[g], 1u);
        __syncthreads();
        if (threadIdx.x < 64 && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicSub(&group_count[threadIdx.x], s_cnt[threadIdx.x]) - s_cnt[threadIdx.x];
        __syncthreads();
        if (g >= 64) perm[group_start[g] + (atomicSub(&group_count[g], 1u) - 1u)] = k;   // colours stacked beyond the mask (k_order)
        else if (g >= 0) perm[group_start[g] + s_base[g] + local] = k;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- ContactConstraint::new
// solver.rs:101-191.  Input manifolds come either from the step's ContactList (one contact,
// tangents from compute_basis) or from user arrays (mgfb_solver_solve).
// dataflow solver schedule (k_solve_df, below): per-row inboxes, row-resident inertia, successor links
struct DfArrays {
    Inbox* in_a; Inbox* in_b;   // [row] inbox of the row's body a / body b
    float4* ia;                 // [5][row_cap]: I_a (9 floats, columns), inv_mass_a, I_b (9), inv_mass_b
    unsigned row_cap;
    unsigned* next;             // [2][row_cap]: successor of the row on body a / b: row << 2 | wraps << 1 | side (0 = its a, 1 = its b)
    unsigned* dep;              // [row] seq_a | deg_a << 8 | seq_b << 16 | deg_b << 24 (build-time scratch)
    const unsigned* body_start; // [n+1] CSR offsets: sum of rows per body
    unsigned* inc;              // [sum deg] rows of every body in solve order
};
struct ManifoldInput {
    const int* a; const int* b;
    const float4* la; const float4* lb; const float4* nt;   // step path (ContactList)
    // user path
    const float* normal; const float* tangent; const uint32_t* ncontacts;
    const float* ula; const float* ulb; const float* static_center; const float* static_friction;
    float4 terrain_center;   // step path: Static{center: terrain.center(), friction: 0}
    bool user;
};
struct BodyInfoView {
    const float4* x; const BodyVel* vel; const Collider* col; const float4* force; const float4* torque;
};
HD void basis_from_normal(V3 n, V3* t0, V3* t1) {  // geom.rs:1138-1145
    V3 b = fabsf(n.x) >= 0.57735f ? mk3(n.y, -n.x, 0.0f) : mk3(0.0f, n.z, -n.y);
    b = unit(b);
    *t0 = b; *t1 = cross3(n, b);
}
struct BodyState { V3 v, w, x; float inv_mass, rest, fric; M3 I; };
__device__ __forceinline__ BodyState load_body_info(const BodyInfoView& B, int i) {
    BodyState s;
    BodyVel r = B.vel[i];
    s.v = mk3(r.a.x, r.a.y, r.a.z); s.w = mk3(r.a.w, r.b.x, r.b.y); s.inv_mass = r.b.z; s.I = vel_inertia(r);
    s.x = f4v(B.x[i]) + f4v(B.col[i].v);   // physics.rs:282  x + collider.delta()
    s.rest = B.force[i].w; s.fric = B.torque[i].w;
    return s;
}
__device__ __forceinline__ BodyState static_body_info(V3 center, float fric) {  // physics.rs:289-302
    BodyState s; s.v = zero3(); s.w = zero3(); s.x = center; s.inv_mass = 0.0f; s.rest = 0.0f; s.fric = fric; s.I = m_zero();
    return s;
}
__device__ __forceinline__ void contact_state(const BodyState& A, const BodyState& Bs, V3 ra, V3 rb, V3 n, V3 t0, V3 t1, float dt,
                                              float baumgarte, float slop, float restitution, float* bias, float* nmass, float* tm0, float* tm1) {
    V3 ca = ra + A.x, cb = rb + Bs.x;
    V3 ra_cn = cross3(ra, n), rb_cn = cross3(rb, n);
    float pen = dot3(cb - ca, n);
    V3 dv = Bs.v + cross3(Bs.w, rb) - A.v - cross3(A.w, ra);
    float rel_v = dot3(dv, n);
    *bias = -baumgarte / dt * (pen > 0.0f ? 0.0f : pen + slop) + (rel_v < -1.0f ? -restitution * rel_v : 0.0f);
    *nmass = 1.0f / (A.inv_mass + dot3(ra_cn, mmulv(A.I, ra_cn)) + Bs.inv_mass + dot3(rb_cn, mmulv(Bs.I, rb_cn)));
    V3 ra_ct = cross3(ra, t0), rb_ct = cross3(rb, t0);
    *tm0 = 1.0f / (A.inv_mass + dot3(ra_ct, mmulv(A.I, ra_ct)) + Bs.inv_mass + dot3(rb_ct, mmulv(Bs.I, rb_ct)));
    ra_ct = cross3(ra, t1); rb_ct = cross3(rb, t1);
    *tm1 = 1.0f / (A.inv_mass + dot3(ra_ct, mmulv(A.I, ra_ct)) + Bs.inv_mass + dot3(rb_ct, mmulv(Bs.I, rb_ct)));
}
__global__ void __launch_bounds__(MGFB_THREADS) k_build_rows(ManifoldInput M, BodyInfoView B, const unsigned* __restrict__ perm, ConstraintRows R,
                                                            const unsigned* m_ptr, unsigned m_host, float dt, float baumgarte, float slop, Counters* ctr,
                                                            const unsigned char* __restrict__ edge_mark, unsigned n_own,
                                                            const int* __restrict__ group, const unsigned long long* __restrict__ body_mask, DfArrays D) {
    if (ctr->overflow | ctr->nan_bounds) return;
    const unsigned m = m_ptr ? *m_ptr : m_host;
    for (unsigned row = blockIdx.x * blockDim.x + threadIdx.x; row < m; row += gridDim.x * blockDim.x) {
        unsigned k = perm[row];
        int a = M.a[k], b = M.b[k];
        if (D.dep) {
            // dataflow schedule (k_solve_df): this row is the seq-th of deg rows at each of its bodies, in colour order
            int g = group[k];
            unsigned d = 0u;
            if (g >= 0 && g < 64) {
                unsigned long long below = (1ULL << g) - 1ULL;
                // inc = the rows of every body in solve order (CSR over body_start): k_df_init links each row to its successors
                if (a >= 0) { unsigned long long ma = body_mask[a]; unsigned sq = (unsigned)__popcll(ma & below); d |= sq | ((unsigned)__popcll(ma) << 8); D.inc[D.body_start[a] + sq] = row; }
                if (b >= 0) { unsigned long long mb = body_mask[b]; unsigned sq = (unsigned)__popcll(mb & below); d |= (sq << 16) | ((unsigned)__popcll(mb) << 24); D.inc[D.body_start[b] + sq] = row; }
            }
            D.dep[row] = d;
        }
        if (edge_mark && b >= 0 && (((unsigned)a >= n_own) != ((unsigned)b >= n_own))) {
            // boundary constraint: its owned body is updated here while the left neighbour may update
            // the bodies it holds as ghosts -- the two sets must not meet (tile too thin otherwise)
            unsigned own = (unsigned)a >= n_own ? (unsigned)b : (unsigned)a;
            if (edge_mark[own]) atomicOr(&ctr->comm_error, (unsigned)COMM_TILE_TOO_THIN);
        }
        V3 n, t0, t1; unsigned nc = 1;
        BodyState A, Bs;
        if (!M.user) {
            float4 la4 = M.la[k], lb4 = M.lb[k], nt4 = M.nt[k];
            n = mk3(la4.w, lb4.w, nt4.x);
            basis_from_normal(n, &t0, &t1);
            A = load_body_info(B, a);
            Bs = b >= 0 ? load_body_info(B, b) : static_body_info(f4v(M.terrain_center), 0.0f);
        } else {
            n = mk3(M.normal[3 * k], M.normal[3 * k + 1], M.normal[3 * k + 2]);
            t0 = mk3(M.tangent[6 * k], M.tangent[6 * k + 1], M.tangent[6 * k + 2]);
            t1 = mk3(M.tangent[6 * k + 3], M.tangent[6 * k + 4], M.tangent[6 * k + 5]);
            nc = M.ncontacts[k];
            V3 sc = mk3(M.static_center[3 * k], M.static_center[3 * k + 1], M.static_center[3 * k + 2]);
            A = a >= 0 ? load_body_info(B, a) : static_body_info(sc, M.static_friction[k]);
            Bs = b >= 0 ? load_body_info(B, b) : static_body_info(sc, M.static_friction[k]);
        }
        if (D.dep) {   // immutable during the solve: the dataflow solver streams them with the row instead of gathering BodyVel
            const unsigned rc = D.row_cap;
            D.ia[row] = make_float4(A.I.c0.x, A.I.c0.y, A.I.c0.z, A.I.c1.x);
            D.ia[rc + row] = make_float4(A.I.c1.y, A.I.c1.z, A.I.c2.x, A.I.c2.y);
            D.ia[2 * rc + row] = make_float4(A.I.c2.z, A.inv_mass, Bs.I.c0.x, Bs.I.c0.y);
            D.ia[3 * rc + row] = make_float4(Bs.I.c0.z, Bs.I.c1.x, Bs.I.c1.y, Bs.I.c1.z);
            D.ia[4 * rc + row] = make_float4(Bs.I.c2.x, Bs.I.c2.y, Bs.I.c2.z, Bs.inv_mass);
        }
        float restitution = fmaxf(A.rest, Bs.rest);   // solver.rs:125
        for (unsigned c = 0; c < nc; ++c) {
            V3 ra, rb;
            if (!M.user) { ra = f4v(M.la[k]); rb = f4v(M.lb[k]); }
            else {
                ra = mk3(M.ula[12 * k + 3 * c], M.ula[12 * k + 3 * c + 1], M.ula[12 * k + 3 * c + 2]);
                rb = mk3(M.ulb[12 * k + 3 * c], M.ulb[12 * k + 3 * c + 1], M.ulb[12 * k + 3 * c + 2]);
            }
            float bias, nmass, tm0, tm1;
            contact_state(A, Bs, ra, rb, n, t0, t1, dt, baumgarte, slop, restitution, &bias, &nmass, &tm0, &tm1);
            if (c == 0) {
                R.ab[row] = make_int2(a, b);
                R.n[row] = v4(n, bias);
                R.t0[row] = v4(t0, tm0);
                R.t1[row] = v4(t1, tm1);
                R.ra[row] = v4(ra, nmass);
                R.rb[row] = v4(rb, ibits((int)nc));
                R.impulse[row] = 0.0f;
            } else {
                unsigned e = row * 3 + (c - 1);
                R.xra[e] = v4(ra, nmass);
                R.xrb[e] = v4(rb, bias);
                R.xtm[e] = make_float4(tm0, tm1, 0.0f, 0.0f);
            }
        }
    }
}

// ---------------------------------------------------------------- Solver::solve
// solver.rs:72-78 + :203-252.  Persistent cooperative kernel: for each iteration, for each
// group (colour / level) in order, every row of the group in parallel, then a grid barrier.
// Rows of a group share no dynamic body, so the result equals the reference's sequential
// sweep over the rows in row order, bit for bit.
__device__ __forceinline__ void load_vel(const BodyVel* p, V3* v, V3* w, float* im, M3* I) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3);
    *v = mk3(a.x, a.y, a.z); *w = mk3(a.w, b.x, b.y); *im = b.z;
    *I = mkm(mk3(b.w, c.x, c.y), mk3(c.z, c.w, d.x), mk3(d.y, d.z, d.w));
}
__device__ __forceinline__ void store_vel(BodyVel* p, V3 v, V3 w, float im, const M3& I) {
    float4* q = reinterpret_cast<float4*>(p);
    __stcg(q, make_float4(v.x, v.y, v.z, w.x));
    __stcg(q + 1, make_float4(w.y, w.z, im, I.c0.x));
}
__device__ __forceinline__ void apply_impulse(V3 J, V3 ra, V3 rb, float ima, float imb, const M3& IA, const M3& IB, V3& va, V3& oa, V3& vb, V3& ob) {
    va = va - J * ima;                       // solver.rs:228-231 / :244-247
    oa = oa - mmulv(IA, cross3(ra, J));
    vb = vb + J * imb;
    ob = ob + mmulv(IB, cross3(rb, J));
}
// Immutable part of a row (everything but the bodies' velocities and the accumulated impulse).
struct RowData { int2 ab; float4 n, t0, t1, ra, rb; };
__device__ __forceinline__ RowData load_row(const ConstraintRows& R, unsigned row) {
    RowData d;
    d.ab = R.ab[row]; d.n = R.n[row]; d.t0 = R.t0[row]; d.t1 = R.t1[row]; d.ra = R.ra[row]; d.rb = R.rb[row];
    return d;
}
__device__ __forceinline__ void solve_row(const ConstraintRows& R, BodyVel* vel, unsigned row, const RowData& d) {
    int2 ab = d.ab;
    V3 va = zero3(), oa = zero3(), vb = zero3(), ob = zero3();
    float ima = 0.0f, imb = 0.0f; M3 IA = m_zero(), IB = m_zero();
    if (ab.x >= 0) load_vel(vel + ab.x, &va, &oa, &ima, &IA);
    if (ab.y >= 0) load_vel(vel + ab.y, &vb, &ob, &imb, &IB);
    float4 n4 = d.n, t04 = d.t0, t14 = d.t1, ra4 = d.ra, rb4 = d.rb;
    V3 n = f4v(n4), t0 = f4v(t04), t1 = f4v(t14);
    int nc = (int)fbits(rb4.w);
    for (int c = 0; c < nc; ++c) {
        V3 ra, rb; float bias, nmass, tm0, tm1, imp;
        if (c == 0) { ra = f4v(ra4); rb = f4v(rb4); bias = n4.w; nmass = ra4.w; tm0 = t04.w; tm1 = t14.w; imp = __ldcg(&R.impulse[row]); }
        else {
            unsigned e = row * 3 + (c - 1);
            float4 xa = R.xra[e], xb = R.xrb[e], xt = __ldcg(&R.xtm[e]);
            ra = f4v(xa); rb = f4v(xb); nmass = xa.w; bias = xb.w; tm0 = xt.x; tm1 = xt.y; imp = xt.z;
        }
        // friction: both tangents use the SAME dv, impulses are applied unclamped (solver.rs:217-232)
        V3 dv = vb + cross3(ob, rb) - va - cross3(oa, ra);
        float l0 = -dot3(dv, t0) * tm0;
        apply_impulse(t0 * l0, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
        float l1 = -dot3(dv, t1) * tm1;
        apply_impulse(t1 * l1, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
        // normal (solver.rs:234-247)
        V3 dv2 = vb + cross3(ob, rb) - va - cross3(oa, ra);
        float vn = dot3(dv2, n);
        float lambda = nmass * (-vn + bias);
        float prev = imp;
        imp = fmaxf(prev + lambda, 0.0f);
        lambda = imp - prev;
        apply_impulse(n * lambda, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
        if (c == 0) __stcg(&R.impulse[row], imp);
        else { unsigned e = row * 3 + (c - 1); __stcg(&R.xtm[e], make_float4(tm0, tm1, imp, 0.0f)); }
    }
    if (ab.x >= 0) store_vel(vel + ab.x, va, oa, ima, IA);
    if (ab.y >= 0) store_vel(vel + ab.y, vb, ob, imb, IB);
}
#define MGFB_SOLVE_THREADS 512
// TILED: every iteration runs  interior phases | edge velocities to the left neighbour's ghost
// slots | boundary phases | ghost velocities back to the right neighbour  (see tile.cuh).
template <bool TILED>
__global__ void __launch_bounds__(MGFB_SOLVE_THREADS, 1) k_solve(ConstraintRows R, BodyVel* vel, const unsigned* __restrict__ phase_start,
                                                                unsigned iters, Counters* ctr, TileLink T, unsigned only_beyond) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (only_beyond && ctr->ngroups <= only_beyond) return;   // k_solve_df (dataflow schedule) took this step
    const unsigned P = ctr->n_phases;
    const unsigned Pint = TILED ? ctr->n_int_phases : P;
    if (!TILED && P == 0) return;
    // Rows are dealt to warps round-robin ACROSS the SMs (warp w of CTA b is global warp
    // w*gridDim + b), so a group with fewer rows than threads still spreads over all 148 SMs
    // instead of saturating the issue slots of the first few.
    const unsigned tid = (((threadIdx.x >> 5) * gridDim.x + blockIdx.x) << 5) | (threadIdx.x & 31u), nth = gridDim.x * blockDim.x;
    unsigned phase = 0;
    unsigned n_edge = 0, n_ghost = 0;
    if (TILED) {
        n_edge = min(ctr->n_edge, T.left.ghost_cap);
        n_ghost = ctr->n_total - T.n_own;
        // the right neighbour may start overwriting my ghost velocities: the rows are built
        if (T.has_right && blockIdx.x == 0 && threadIdx.x == 0) st_release_sys_u64(&T.right.mbox->vel_from_left.flag, tile_seq(T.step, 1));
    }
    // The first row this thread owns in the next phase is fetched BEFORE the grid barrier: rows
    // are immutable during the solve, so only the body gather sits behind the barrier.
    unsigned r0 = 0, r1 = 0;
    RowData pre; bool have = false;
    if (P) { r0 = phase_start[0]; r1 = phase_start[1]; if (r0 + tid < r1) { pre = load_row(R, r0 + tid); have = true; } }
    for (unsigned it = 0; it < iters; ++it) {
        // my edge bodies carry the left neighbour's boundary results of the previous iteration
        if (TILED && T.has_left && it > 0) tile_wait_cta(&T.mine->vel_from_left.flag, tile_seq(T.step, 1 + it), T.timeout_ns, ctr);
        for (unsigned p = 0; p < P; ++p) {
            if (TILED && p == Pint) {
                if (T.has_left) {
                    if (it == 0) tile_wait_cta(&T.mine->vel_from_left.flag, tile_seq(T.step, 1), T.timeout_ns, ctr);
                    tile_push_edges(T, vel, n_edge, tid, nth);
                    grid_barrier(&ctr->bar, phase);
                    if (blockIdx.x == 0 && threadIdx.x == 0) st_release_sys_u64(&T.left.mbox->vel_from_right.flag, tile_seq(T.step, 1 + it));
                }
                if (T.has_right) tile_wait_cta(&T.mine->vel_from_right.flag, tile_seq(T.step, 1 + it), T.timeout_ns, ctr);
            }
            unsigned row = r0 + tid;
            if (have) solve_row(R, vel, row, pre);
            for (row += nth; row < r1; row += nth) { RowData d = load_row(R, row); solve_row(R, vel, row, d); }
            unsigned pn = p + 1 == P ? 0 : p + 1;
            bool more = (p + 1 < P) || (it + 1 < iters);
            r0 = phase_start[pn]; r1 = phase_start[pn + 1];
            have = false;
            if (more && r0 + tid < r1) { pre = load_row(R, r0 + tid); have = true; }
            grid_barrier(&ctr->bar, phase);
        }
        if (TILED) {
            if (Pint == P && T.has_left) {   // no boundary phase on this rank: the exchange still happens
                if (it == 0) tile_wait_cta(&T.mine->vel_from_left.flag, tile_seq(T.step, 1), T.timeout_ns, ctr);
                tile_push_edges(T, vel, n_edge, tid, nth);
                grid_barrier(&ctr->bar, phase);
                if (blockIdx.x == 0 && threadIdx.x == 0) st_release_sys_u64(&T.left.mbox->vel_from_right.flag, tile_seq(T.step, 1 + it));
            }
            if (T.has_right) {
                if (Pint == P) tile_wait_cta(&T.mine->vel_from_right.flag, tile_seq(T.step, 1 + it), T.timeout_ns, ctr);
                tile_return_ghosts(T, vel, n_ghost, tid, nth);
                grid_barrier(&ctr->bar, phase);
                if (blockIdx.x == 0 && threadIdx.x == 0) st_release_sys_u64(&T.right.mbox->vel_from_left.flag, tile_seq(T.step, 2 + it));
            }
        }
    }
    if (TILED && T.has_left) tile_wait_cta(&T.mine->vel_from_left.flag, tile_seq(T.step, 1 + iters), T.timeout_ns, ctr);
}
// ---------------------------------------------------------------- Solver::solve, dataflow schedule
// Same arithmetic and the same row order as k_solve, but NO grid barriers and no gathers.  The
// rows of one body form a chain in solve order (its colours ascending; the last row is followed
// by the first row of the next iteration).  A row does not fetch its bodies' velocities: its
// predecessor on each chain PUSHES them into the row's own inbox (one 32-byte record per side:
// v, omega and a tag = epoch + iteration, written with ONE 32-byte store, so tag and payload
// cannot be seen apart and no fence is needed).  A warp owns 32 consecutive rows of one colour
// (never dependent on each other), polls their inboxes with two fully coalesced 1 KB loads until
// every tag shows the current iteration, updates, and pushes the results to the successors'
// inboxes.  World inverse inertia and inverse mass are immutable during the solve and are copied
// into the row when it is built, so everything a row reads is a coalesced stream, loaded at the
// start of the visit together with the first look at the inboxes (no register prefetch: measured
// faster, the freed registers buy more resident warps).  The sequential sweep over the rows in row order is the unique execution these
// waits allow: the result is bit-identical to k_solve and to the reference's loop over that
// order.  Progress: a warp visits its warp-rows in increasing (iteration, row) order and all warps
// are co-resident (cooperative launch), so the smallest unfinished warp-row always has its
// inboxes filled.  Iterations pipeline: a body's rows of iteration it+1 start as soon as ITS
// OWN iteration `it` is complete.
template <bool SYS>
__device__ __forceinline__ Inbox ld_inbox(const Inbox* p) {
    Inbox r;
    if (SYS) asm volatile("ld.relaxed.sys.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p) : "memory");
    return r;
}
template <bool SYS>
__device__ __forceinline__ void st_inbox(Inbox* p, V3 v, V3 w, unsigned tag) {
    float ft = __uint_as_float(tag);
    if (SYS) asm volatile("st.relaxed.sys.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(ft), "f"(w.x), "f"(w.y), "f"(w.z), "f"(ft) : "memory");
    else asm volatile("st.relaxed.gpu.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(ft), "f"(w.x), "f"(w.y), "f"(w.z), "f"(ft) : "memory");
}
// Successor word of a row on one of its bodies: row << 4 | where << 2 | wraps << 1 | side.
//   side   0 = the body is the successor's body a, 1 = its body b
//   wraps  the successor runs in the NEXT iteration (this row is the last of the body's chain)
//   where  0 = this tile, 1 = the left neighbour's rows, 2 = the right neighbour's rows (tiled world)
#define DF_NONE 0xffffffffu
#define DF_LEFT 1u
#define DF_RIGHT 2u
// Tiled world: for every ghost, tell its owner (right) the ghost's first boundary row here; for every edge
// body, tell the left tile the body's first interior row here.  One CTA; flags after a system fence.
__global__ void __launch_bounds__(1024) k_tile_links_send(TileLink T, DfArrays D, const int2* __restrict__ ab, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds | ctr->comm_error) return;
    auto first_row = [&](unsigned body) {
        unsigned s0 = D.body_start[body], s1 = D.body_start[body + 1];
        if (s0 == s1) return DF_NONE;
        unsigned fr = D.inc[s0];
        return (fr << 1) | (ab[fr].x == (int)body ? 0u : 1u);
    };
    if (T.has_right) {
        const unsigned n_ghost = ctr->n_total - T.n_own;
        for (unsigned k = threadIdx.x; k < n_ghost; k += blockDim.x) T.right.link_l[k] = first_row(T.n_own + k);
    }
    if (T.has_left) {
        const unsigned n_edge = min(ctr->n_edge, T.left.ghost_cap);
        for (unsigned k = threadIdx.x; k < n_edge; k += blockDim.x) T.left.link_r[k] = first_row(T.edge_idx[k]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (T.has_right) st_release_sys_u64(&T.right.mbox->links_from_left.flag, T.step);
        if (T.has_left) st_release_sys_u64(&T.left.mbox->links_from_right.flag, T.step);
    }
}
__global__ void k_tile_solve_done(TileLink T, Counters* ctr) {
    if (!T.has_right) return;
    __threadfence_system();
    st_release_sys_u64(&T.right.mbox->done_from_left.flag, T.step);
}
// Rows: successor links.  Bodies: v, omega into the inbox of the body's FIRST row, tagged for iteration 0.
template <bool TILED>
__global__ void __launch_bounds__(MGFB_THREADS) k_df_init(const BodyVel* __restrict__ vel, unsigned n, const int2* __restrict__ ab, DfArrays D,
                                                         unsigned epoch, const unsigned* m_ptr, unsigned m_host, Counters* ctr, TileLink T) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (ctr->ngroups > 64u) return;   // k_solve takes this step
    if (TILED) {
        n = ctr->n_total;
        // the neighbours' link tables (k_tile_links_send over there): every CTA waits for the flags itself
        if (T.has_left) tile_wait_cta(&T.mine->links_from_left.flag, T.step, T.timeout_ns, ctr);
        if (T.has_right) tile_wait_cta(&T.mine->links_from_right.flag, T.step, T.timeout_ns, ctr);
        if (*reinterpret_cast<volatile unsigned*>(&ctr->comm_error) & COMM_TIMEOUT) return;
    }
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (unsigned i = tid; i < n; i += nth) {
        unsigned s0 = D.body_start[i], s1 = D.body_start[i + 1];
        if (s0 == s1) continue;
        // a ghost whose owner has rows of its own is seeded there: its chain reaches this tile through the boundary
        if (TILED && i >= T.n_own && T.link_r[i - T.n_own] != DF_NONE) continue;
        unsigned fr = D.inc[s0];
        const float4* q = reinterpret_cast<const float4*>(vel + i);
        float4 a = q[0], b = q[1];
        Inbox* dst = (ab[fr].x == (int)i) ? D.in_a + fr : D.in_b + fr;
        st_inbox<false>(dst, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), epoch + 1u);
    }
    const unsigned m = m_ptr ? *m_ptr : m_host;
    for (unsigned row = tid; row < m; row += nth) {
        int2 p = ab[row]; unsigned d = D.dep[row];
        unsigned seq[2] = {d & 255u, (d >> 16) & 255u}, deg[2] = {(d >> 8) & 255u, d >> 24};
        int body[2] = {p.x, p.y};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unsigned nx = DF_NONE;
            const int pb = body[s];
            if (pb >= 0) {
                const bool wrap = seq[s] + 1u == deg[s];
                unsigned remote = DF_NONE;
                if (TILED && wrap) {
                    if ((unsigned)pb >= T.n_own) {   // ghost: after its last boundary row here, back to its owner's first row, next iteration
                        unsigned l = T.link_r[(unsigned)pb - T.n_own];
                        if (l != DF_NONE) remote = ((l >> 1) << 4) | (DF_RIGHT << 2) | 2u | (l & 1u);
                    } else if (T.has_left && T.edge_mark[pb]) {   // edge body: after its last row here, on to its boundary rows on the left tile, same iteration
                        unsigned l = T.link_l[T.edge_slot[pb]];
                        if (l != DF_NONE) remote = ((l >> 1) << 4) | (DF_LEFT << 2) | (l & 1u);
                    }
                }
                if (remote != DF_NONE) nx = remote;
                else {
                    unsigned nr = D.inc[D.body_start[pb] + (wrap ? 0u : seq[s] + 1u)];
                    nx = (nr << 4) | (wrap ? 2u : 0u) | (ab[nr].x == pb ? 0u : 1u);
                }
            }
            D.next[s * D.row_cap + row] = nx;
        }
    }
}
#ifdef MGFB_DF_PROFILE
__device__ unsigned long long g_df_prof[8];   // cycles: fetch, inbox poll, compute+publish ; counts: polls, visits, warps
#define DF_T(x) long long x = clock64()
#define DF_ACC(i, v) prof[i] += (unsigned long long)(v)
#else
#define DF_T(x)
#define DF_ACC(i, v)
#endif
struct DfRow { RowData d; float4 i0, i1, i2, i3, i4; unsigned na, nb, row; float imp; bool valid; };
#define MGFB_DF_MAX_PHASES 64
// Block size is chosen at launch: 10 warps/SM measured best while the rows are L2-resident (100 k .. 300 k bodies;
// 256: +20 %, 384: +6 % solve time), 16 warps/SM once they stream from HBM (> ~1 M rows).
#ifndef MGFB_DF_THREADS_SMALL
#define MGFB_DF_THREADS_SMALL 320
#endif
#define MGFB_DF_THREADS_LARGE 512
#define MGFB_DF_LARGE_ROWS 1000000u
template <bool TILED>
__global__ void __launch_bounds__(MGFB_DF_THREADS_LARGE, 1) k_solve_df(ConstraintRows R, DfArrays D, BodyVel* vel, const unsigned* __restrict__ phase_start,
                                                                unsigned iters, unsigned epoch, Counters* ctr, TileLink T) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (TILED && ctr->comm_error) return;
    const unsigned P = ctr->n_phases;
    if (P == 0 || ctr->ngroups > MGFB_DF_MAX_PHASES || iters == 0) return;   // > 64 colours: k_solve takes the step
    __shared__ unsigned s_row0[MGFB_DF_MAX_PHASES + 1], s_wr0[MGFB_DF_MAX_PHASES + 1];
    if (threadIdx.x == 0) {
        unsigned w = 0;
        for (unsigned p = 0; p < P; ++p) { unsigned r0 = phase_start[p], r1 = phase_start[p + 1]; s_row0[p] = r0; s_wr0[p] = w; w += (r1 - r0 + 31u) >> 5; }
        s_row0[P] = phase_start[P]; s_wr0[P] = w;
    }
    __syncthreads();
    const unsigned nwr = s_wr0[P];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nW = gridDim.x * (blockDim.x >> 5);
    if (gw >= nwr) return;
    const unsigned rc = D.row_cap;
    auto fetch = [&](unsigned wr, unsigned& p) {
        DfRow f;
        while (wr >= s_wr0[p + 1]) ++p;
        unsigned row = s_row0[p] + ((wr - s_wr0[p]) << 5) + lane;
        f.row = row;
        f.valid = row < s_row0[p + 1];
        if (f.valid) {
            f.d.ab = make_int2(0, 0);   // not streamed: the links say which sides exist, the body index is only needed for the final store
            f.d.n = R.n[row]; f.d.t0 = R.t0[row]; f.d.t1 = R.t1[row]; f.d.ra = R.ra[row]; f.d.rb = R.rb[row];
            f.i0 = D.ia[row]; f.i1 = D.ia[rc + row]; f.i2 = D.ia[2 * rc + row]; f.i3 = D.ia[3 * rc + row]; f.i4 = D.ia[4 * rc + row];
            f.na = D.next[row]; f.nb = D.next[rc + row];
            f.imp = __ldcg(&R.impulse[row]);
        } else {
            f.d.ab = make_int2(-1, -1); f.na = f.nb = DF_NONE; f.imp = 0.0f;
            f.d.n = f.d.t0 = f.d.t1 = f.d.ra = f.d.rb = f.i0 = f.i1 = f.i2 = f.i3 = f.i4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        return f;
    };
    // where a row's result goes: the successor's inbox (here or on a neighbour tile), or -- after the body's
    // last row of the last iteration -- the body's record (on its owner, if the body is a ghost)
    auto publish = [&](unsigned nx, int body, V3 v, V3 w, float im, const M3& I, unsigned tag, bool last_it) {
        const bool wrap = (nx & 2u) != 0u;
        if (wrap && last_it) {
            BodyVel* dst = vel + body;
            if (TILED && (unsigned)body >= T.n_own) dst = T.right.vel + T.ridx[(unsigned)body - T.n_own];
            store_vel(dst, v, w, im, I);
            return;
        }
        const unsigned where = TILED ? (nx >> 2) & 3u : 0u, side = nx & 1u, nr = nx >> 4;
        Inbox* base = side ? D.in_b : D.in_a;
        if (TILED && where == DF_LEFT) base = side ? T.left.in_b : T.left.in_a;
        if (TILED && where == DF_RIGHT) base = side ? T.right.in_b : T.right.in_a;
        st_inbox<TILED>(base + nr, v, w, tag + (wrap ? 1u : 0u));
    };
    unsigned p_next = 0, wr = gw, it = 0;
#ifdef MGFB_DF_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    float carried_imp = 0.0f; bool have_carried = false;
    for (;;) {
        DF_T(t0);
        // no register prefetch: the row is loaded at the start of its visit, together with the first look at its
        // inboxes (both are L2 round trips, and most visits wait for an input anyway) -- 48 registers less per
        // thread, so more warps per SM hide each other's latency
        DfRow cur = fetch(wr, p_next);
        if (have_carried) cur.imp = carried_imp;
        const unsigned row = cur.row;
        unsigned wr_n = wr + nW, it_n = it;
        if (wr_n >= nwr) { wr_n = gw; it_n = it + 1; p_next = 0; }
        const bool more = it_n < iters;
        const bool same_rows = (wr_n == wr);   // this warp owns one warp-row: its impulse is carried in registers
        const bool needA = cur.valid && cur.na != DF_NONE, needB = cur.valid && cur.nb != DF_NONE;
        const unsigned tag = epoch + it + 1u;
        Inbox sa, sb;
        sa.lo = sa.hi = sb.lo = sb.hi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        bool okA = !needA, okB = !needB;
        DF_T(t1); DF_ACC(0, t1 - t0);
        unsigned long long t_wait0 = 0; unsigned spins = 0;
        for (;;) {
            if (!okA) { sa = ld_inbox<TILED>(D.in_a + row); okA = __float_as_uint(sa.lo.w) == tag && __float_as_uint(sa.hi.w) == tag; }
            if (!okB) { sb = ld_inbox<TILED>(D.in_b + row); okB = __float_as_uint(sb.lo.w) == tag && __float_as_uint(sb.hi.w) == tag; }
            DF_ACC(3, 1);
            if (__all_sync(0xffffffffu, okA && okB)) break;
            if (TILED && (++spins & 1023u) == 0u) {   // a dead or diverged neighbour must never hang the GPU
                bool dead = false;
                if (lane == 0) {
                    unsigned long long now = globaltimer_ns();
                    if (t_wait0 == 0) t_wait0 = now;
                    if (*reinterpret_cast<volatile unsigned*>(&ctr->comm_error) & COMM_TIMEOUT) dead = true;
                    else if (now - t_wait0 > T.timeout_ns) { atomicOr(&ctr->comm_error, (unsigned)COMM_TIMEOUT); dead = true; }
                }
                if (__shfl_sync(0xffffffffu, (int)dead, 0)) return;
            }
        }
        DF_T(t2); DF_ACC(1, t2 - t1);
        if (cur.valid) {
            V3 va = mk3(sa.lo.x, sa.lo.y, sa.lo.z), oa = mk3(sa.hi.x, sa.hi.y, sa.hi.z);
            V3 vb = mk3(sb.lo.x, sb.lo.y, sb.lo.z), ob = mk3(sb.hi.x, sb.hi.y, sb.hi.z);
            const M3 IA = mkm(mk3(cur.i0.x, cur.i0.y, cur.i0.z), mk3(cur.i0.w, cur.i1.x, cur.i1.y), mk3(cur.i1.z, cur.i1.w, cur.i2.x));
            const float ima = cur.i2.y;
            const M3 IB = mkm(mk3(cur.i2.z, cur.i2.w, cur.i3.x), mk3(cur.i3.y, cur.i3.z, cur.i3.w), mk3(cur.i4.x, cur.i4.y, cur.i4.z));
            const float imb = cur.i4.w;
            float4 n4 = cur.d.n, t04 = cur.d.t0, t14 = cur.d.t1, ra4 = cur.d.ra, rb4 = cur.d.rb;
            V3 n = f4v(n4), t0 = f4v(t04), t1 = f4v(t14);
            int nc = (int)fbits(rb4.w);
            for (int c = 0; c < nc; ++c) {
                V3 ra, rb; float bias, nmass, tm0, tm1, imp;
                if (c == 0) { ra = f4v(ra4); rb = f4v(rb4); bias = n4.w; nmass = ra4.w; tm0 = t04.w; tm1 = t14.w; imp = cur.imp; }
                else {
                    unsigned e = row * 3 + (c - 1);
                    float4 xa = R.xra[e], xb = R.xrb[e], xt = __ldcg(&R.xtm[e]);
                    ra = f4v(xa); rb = f4v(xb); nmass = xa.w; bias = xb.w; tm0 = xt.x; tm1 = xt.y; imp = xt.z;
                }
                V3 dv = vb + cross3(ob, rb) - va - cross3(oa, ra);        // solver.rs:217-232, same stale dv for both tangents
                float l0 = -dot3(dv, t0) * tm0;
                apply_impulse(t0 * l0, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                float l1 = -dot3(dv, t1) * tm1;
                apply_impulse(t1 * l1, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                V3 dv2 = vb + cross3(ob, rb) - va - cross3(oa, ra);       // solver.rs:234-247
                float vn = dot3(dv2, n);
                float lambda = nmass * (-vn + bias);
                float prev = imp;
                imp = fmaxf(prev + lambda, 0.0f);
                lambda = imp - prev;
                apply_impulse(n * lambda, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                if (c == 0) {
                    if (same_rows) { carried_imp = imp; have_carried = true; }
                    if (!same_rows || it + 1 == iters) __stcg(&R.impulse[row], imp);
                }
                else { unsigned e = row * 3 + (c - 1); __stcg(&R.xtm[e], make_float4(tm0, tm1, imp, 0.0f)); }
            }
            const bool last_it = it + 1 == iters;
            int a = 0, b = 0;
            if (last_it) { int2 ab = R.ab[row]; a = ab.x; b = ab.y; }
            if (needA) publish(cur.na, a, va, oa, ima, IA, tag, last_it);
            if (needB) publish(cur.nb, b, vb, ob, imb, IB, tag, last_it);
        }
        __syncwarp();
        DF_T(t3); DF_ACC(2, t3 - t2); DF_ACC(4, 1);
        if (!more) break;
        wr = wr_n; it = it_n;
    }
#ifdef MGFB_DF_PROFILE
    if (lane == 0) for (int i = 0; i < 5; ++i) atomicAdd(&g_df_prof[i], prof[i]);
    if (lane == 0) atomicAdd(&g_df_prof[5], 1ULL);
#endif
}

// `snap` (pipelined steps): a copy of the step's counters in a buffer of its own, so that the host copy can ride the
// D2H stream behind the step instead of sitting in the step's stream (where it would queue behind the previous step's
// state transfer on the one D2H copy engine and hold the next step back).
__global__ void __launch_bounds__(64) k_step_done(Counters* ctr, Counters* snap) {
    if (threadIdx.x == 0 && !(ctr->overflow | ctr->nan_bounds)) {
        ctr->steps_done++;
        ctr->acc_steps++;
        ctr->acc_constraints += ctr->contacts;
        ctr->acc_pairs += (unsigned long long)ctr->pairs[0] + ctr->pairs[1] + ctr->pairs[2] + ctr->pairs[3] + ctr->tpairs[0] + ctr->tpairs[1];
        ctr->acc_groups += ctr->ngroups;
    }
    __syncthreads();
    if (snap) {
        const unsigned* src = reinterpret_cast<const unsigned*>(ctr);
        unsigned* dst = reinterpret_cast<unsigned*>(snap);
        for (unsigned k = threadIdx.x; k < sizeof(Counters) / 4; k += blockDim.x) dst[k] = src[k];
    }
}
// state marshalling for the host API: packed f32 arrays <-> SoA records
__global__ void __launch_bounds__(MGFB_THREADS) k_pack_state(BodyArrays B, unsigned first, unsigned n, float* x, float* q, float* v, float* w) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned b = first + i;
    if (x) { float4 p = B.x[b]; x[3 * i] = p.x; x[3 * i + 1] = p.y; x[3 * i + 2] = p.z; }
    if (q) { float4 p = B.q[b]; q[4 * i] = p.x; q[4 * i + 1] = p.y; q[4 * i + 2] = p.z; q[4 * i + 3] = p.w; }
    if (v || w) {
        BodyVel r = B.vel[b];
        if (v) { v[3 * i] = r.a.x; v[3 * i + 1] = r.a.y; v[3 * i + 2] = r.a.z; }
        if (w) { w[3 * i] = r.a.w; w[3 * i + 1] = r.b.x; w[3 * i + 2] = r.b.y; }
    }
}
// ADD: v += v_in, omega += omega_in (what `bodies.v[i] += dv` between steps does on the CPU); else overwrite (ConstrainedSet::set)
template <bool ADD = false>
__global__ void __launch_bounds__(MGFB_THREADS) k_set_velocity(BodyArrays B, unsigned first, unsigned n, const float* v, const float* w, const Counters* ctr) {
    if (ctr && (ctr->overflow | ctr->nan_bounds)) return;   // a queued step of a poisoned pipeline: it will be re-queued whole
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BodyVel* r = B.vel + first + i;
    if (ADD) {
        float4 a = r->a;
        r->a = make_float4(a.x + v[3 * i], a.y + v[3 * i + 1], a.z + v[3 * i + 2], a.w + w[3 * i]);
        r->b.x = r->b.x + w[3 * i + 1]; r->b.y = r->b.y + w[3 * i + 2];
    } else {
        r->a = make_float4(v[3 * i], v[3 * i + 1], v[3 * i + 2], w[3 * i]);
        r->b.x = w[3 * i + 1]; r->b.y = w[3 * i + 2];
    }
}

// Restores saved state (mgfb_bodies_set_state): the pub fields x, q, collider (physics.rs:142-154), v, omega and the stored
// fat box of the world's body BVH (world.rs:180).  The world inverse inertia follows q like in integrate (physics.rs:231-232).
// Staged arrays are packed f32; any may be NULL (field kept).  col: 9 floats per body = p0.xyz|r is kept| d.xyz | delta.xyz:
//   [0..3) centre (sphere) / a (capsule), [3..6) d (capsule), [6..9) Moving.1
__global__ void __launch_bounds__(MGFB_THREADS) k_set_state(BodyArrays B, unsigned first, unsigned n, const float* x, const float* q, const float* v,
                                                           const float* w, const float* col, const float* fat) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned b = first + i;
    if (x) B.x[b] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.0f);
    BodyVel r = B.vel[b];
    if (q) {
        B.q[b] = make_float4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        Q4 qi = mkq(q[4 * i], mk3(q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]));
        M3 R = m_from_q(qi);
        M3 Ib = mkm(f4v(B.imb[3 * b]), f4v(B.imb[3 * b + 1]), f4v(B.imb[3 * b + 2]));
        vel_set_inertia(r, mmul(mmul(R, Ib), mtrans(R)));
    }
    if (v) { r.a.x = v[3 * i]; r.a.y = v[3 * i + 1]; r.a.z = v[3 * i + 2]; }
    if (w) { r.a.w = w[3 * i]; r.b.x = w[3 * i + 1]; r.b.y = w[3 * i + 2]; }
    if (q || v || w) B.vel[b] = r;
    if (col) {
        Collider k = B.col[b];
        const float* c = col + 9 * i;
        k.p0 = make_float4(c[0], c[1], c[2], k.p0.w);
        if (col_kind(k) != 0) k.p1 = make_float4(c[3], c[4], c[5], k.p1.w);
        k.v = make_float4(c[6], c[7], c[8], k.v.w);
        B.col[b] = k;
    }
    if (fat) {
        Box f; f.c = make_float4(fat[6 * i], fat[6 * i + 1], fat[6 * i + 2], 0.0f); f.r = make_float4(fat[6 * i + 3], fat[6 * i + 4], fat[6 * i + 5], 0.0f);
        B.fat[b] = f;
    }
}

// ---------------------------------------------------------------- single-pass exclusive scan (u32)
// Zero-fill of up to 12 ranges in one launch: the step's scratch (counters, cell counts, group counts, per-body masks,
// scan states ...).  A kernel, not cudaMemsetAsync: memsets ride a copy engine, and with the pipelined step API that
// engine is busy for 0.1-0.4 ms per step moving the previous step's state to the host -- every memset in the step's
// stream then waits for it (measured: +0.33 ms of device time per step on 8 GPUs sharing PCIe, tools/e2e_diag.py).
struct ZeroRanges { unsigned* p[12]; unsigned words[12]; unsigned n; };
__global__ void __launch_bounds__(MGFB_THREADS) k_zero_ranges(ZeroRanges Z) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (unsigned r = 0; r < Z.n; ++r) {
        unsigned* p = Z.p[r];
        const unsigned w = Z.words[r], w4 = w >> 2;
        uint4* p4 = reinterpret_cast<uint4*>(p);   // cudaMalloc'ed: 256-byte aligned
        for (unsigned k = tid; k < w4; k += nth) p4[k] = make_uint4(0u, 0u, 0u, 0u);
        for (unsigned k = (w4 << 2) + tid; k < w; k += nth) p[k] = 0u;
    }
}
// Decoupled look-back: tiles of SCAN_ITEMS items take their number from an atomic ticket (so every
// predecessor of a running tile is running or done), publish their aggregate, then their inclusive
// prefix, in one 64-bit word (flag << 62 | value).  `state` = [ticket, pad, status[ntiles]] zeroed before
// the launch.  out[n] = total (also *total_dev when given).
#define SCAN_ITEMS 2048
__global__ void __launch_bounds__(256) k_scan_lookback(const unsigned* __restrict__ in, unsigned n, unsigned* out, unsigned long long* state, unsigned* total_dev,
                                                       const unsigned* skip_if = nullptr, unsigned skip_value = 0u) {
    __shared__ unsigned s_tile, s_excl, ws[8];
    if (skip_if && *skip_if == skip_value) return;   // (a coherent broadphase step keeps its grid: bpcache.cuh)
    unsigned long long* status = state + 1;
    if (threadIdx.x == 0) s_tile = atomicAdd(reinterpret_cast<unsigned*>(state), 1u);
    __syncthreads();
    const unsigned tile = s_tile, base = tile * SCAN_ITEMS + threadIdx.x * 8u;
    unsigned v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = base + k < n ? in[base + k] : 0u; sum += v[k]; }
    unsigned x = sum;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < (int)(threadIdx.x >> 5)) woff += ws[w]; agg += ws[w]; }
    if (threadIdx.x < 32) {
        // look-back by a whole warp: 32 predecessors per L2 round trip (a single thread walking back one tile at a time
        // made the scan a chain of ~40 dependent round trips at 256 tiles)
        const unsigned lane = threadIdx.x;
        unsigned excl = 0;
        if (tile == 0) { if (lane == 0) st_relaxed_u64(status, (2ULL << 62) | agg); }
        else {
            if (lane == 0) st_relaxed_u64(status + tile, (1ULL << 62) | agg);
            for (int p = (int)tile - 1;; p -= 32) {
                const int idx = p - (int)lane;
                unsigned long long st = 2ULL << 62;   // before tile 0: an inclusive prefix of 0
                if (idx >= 0) do { st = ld_relaxed_u64(status + idx); } while ((st >> 62) == 0ULL);
                const unsigned incl = __ballot_sync(0xffffffffu, (st >> 62) == 2ULL);
                const unsigned upto = incl ? (unsigned)__ffs((int)incl) - 1u : 31u;   // nearest predecessor that knows its prefix
                excl += __reduce_add_sync(0xffffffffu, lane <= upto ? (unsigned)st : 0u);
                if (incl) break;
            }
            if (lane == 0) st_relaxed_u64(status + tile, (2ULL << 62) | (unsigned long long)(excl + agg));
        }
        if (lane == 0) {
            s_excl = excl;
            if ((unsigned long long)(tile + 1) * SCAN_ITEMS >= n) { out[n] = excl + agg; if (total_dev) *total_dev = excl + agg; }
        }
    }
    __syncthreads();
    unsigned run = s_excl + woff + x - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

// ---------------------------------------------------------------- multi-block exclusive scan (u32), three passes (terrain build)
__global__ void __launch_bounds__(256) k_scan_reduce(const unsigned* __restrict__ in, unsigned n, unsigned* block_sums) {
    __shared__ unsigned ws[8];
    unsigned base = blockIdx.x * SCAN_ITEMS, s = 0;
    for (unsigned i = threadIdx.x; i < SCAN_ITEMS; i += 256) { unsigned idx = base + i; if (idx < n) s += in[idx]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < 8; ++w) t += ws[w]; block_sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_sums(unsigned* block_sums, unsigned nblocks, unsigned* total) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < nblocks; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < nblocks ? block_sums[i] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = warp_sums[threadIdx.x], wsum = w;
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wsum, o); if (threadIdx.x >= o) wsum += y; }
            warp_sums[threadIdx.x] = wsum - w;
        }
        __syncthreads();
        unsigned excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_final(const unsigned* __restrict__ in, unsigned n, const unsigned* __restrict__ block_sums, unsigned* out, const unsigned* total) {
    __shared__ unsigned ws[8];
    __shared__ unsigned carry;
    unsigned base = blockIdx.x * SCAN_ITEMS;
    if (threadIdx.x == 0) carry = block_sums[blockIdx.x];
    __syncthreads();
    for (unsigned c = 0; c < SCAN_ITEMS; c += 256) {
        unsigned idx = base + c + threadIdx.x;
        unsigned v = idx < n ? in[idx] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += ws[w];
        unsigned excl = carry + woff + x - v;
        if (idx < n) out[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 255) carry = excl + v;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = *total;
}

}  // namespace mgfb
