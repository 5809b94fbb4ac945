// bpcache.cuh -- temporally coherent body broadphase (included by kernels.cuh after the grid / sweep definitions).
//
// world.rs:261-268 asks, every step, for P = {(i, j): j < i, tight_i overlaps stored fat_j}.  The stored fat boxes only change
// when a body leaves its own (world.rs:235-238): ~1 % of the bodies per step in a settled pile.  So the library keeps
//     S = {(i, j): j < i, fat_i overlaps fat_j}        (a superset of P: tight_i lies inside fat_i; a few ulps of slop on top)
// from step to step, together with the hashed grid it was found with, and a step only
//   * streams S once: pairs with an endpoint whose fat box was replaced THIS step are dropped, the others are kept and tested
//     exactly (tight_i vs fat_j, collision.rs:22-29) -> the step's pair lists,
//   * queries the replaced bodies (and, in a tiled world, this step's ghosts) against the cached grid and appends what it
//     finds to S and to the pair lists.  A grid entry of a body replaced since the grid was built still LOCATES the body (the
//     cells are `drift` wider than two fat half extents, and a body that has drifted further from where it was binned moves
//     to a short overflow list that every query scans); its current box is read from the body arrays.
// The grid and S are rebuilt from scratch (one sweep, k_body_pairs_warp<true>) when the overflow list is full, when too many
// bodies moved, when a fat box outgrew the grid's cell size, or when the host invalidated the cache (bodies added, state
// restored, lists regrown).
// Both paths are launched every step and the device decides (k_bp_decide): no host round trip.  P is the same SET either way;
// its order in the lists never mattered (the narrowphase emits contacts through an atomic cursor).
#pragma once

namespace mgfb {

enum { BP_REBUILD = 0, BP_COHERENT = 1, BP_SWEEP = 2 };   // BP_SWEEP: too much moves for a cache to pay: the plain grid + sweep of this step, nothing kept
struct BpState {            // device-resident, lives across steps
    unsigned mode;          // this step's path (k_bp_decide)
    unsigned valid;         // S and the grid describe the stored fat boxes of the previous step
    unsigned cur;           // which of the two S buffers is current
    unsigned s_count[2];
    unsigned n_ovf;         // bodies whose grid entry is stale (fat box replaced since the grid was built)
    unsigned n_grid;        // bodies in the cached grid
    float s0;               // its cell edge: >= 2 x the largest fat half extent + drift, so overlapping fat boxes are <= 1 cell apart
    float drift;            // how far a body may move from where it was binned before it goes to the overflow list
    unsigned rebuilds, coherent_steps;
    unsigned pad[5];
};
struct BpView {
    BpState* st;
    int2* S[2]; unsigned s_cap;      // (i | kind_i << 31, j | kind_j << 31)
    unsigned char* stale;            // [bodies] 0: the grid entry holds the body's box; 1: replaced since, entry still locates it; 2: in the overflow list
    float4* c0;                      // [bodies] fat centre at the time the grid was built
    unsigned* ovf; unsigned ovf_cap;
    unsigned char* ref_flag;         // [bodies] fat box replaced this step (k_integrate; zeroed with the step's scratch)
    unsigned* ref_list; unsigned ref_cap;
};
#define BP_OVF_CAP 4096u   // hard room: every body replaced in one step may turn out to be a runaway
#define BP_OVF_SOFT 512u   // beyond this many runaways the linear scan costs more than a rebuild saves
#define BP_REF_CAP 16384u

// closed overlap with a few ulps of slop: membership in the cached SUPERSET (never decides a pair by itself)
__device__ __forceinline__ bool box_overlaps_slop(V3 ac, V3 ar, V3 bc, V3 br) {
    float ex = (fabsf(ac.x) + fabsf(bc.x) + ar.x + br.x) * 4e-6f + 1e-30f;
    float ey = (fabsf(ac.y) + fabsf(bc.y) + ar.y + br.y) * 4e-6f + 1e-30f;
    float ez = (fabsf(ac.z) + fabsf(bc.z) + ar.z + br.z) * 4e-6f + 1e-30f;
    return fabsf(ac.x - bc.x) <= (ar.x + br.x) + ex && fabsf(ac.y - bc.y) <= (ar.y + br.y) + ey && fabsf(ac.z - bc.z) <= (ar.z + br.z) + ez;
}

// `force_rebuild`: tiled worlds rebuild on a common schedule (every BP_TILED_PERIOD steps, by the tiles' shared step number):
// tiles run in lock-step, so a tile that rebuilds alone makes its neighbours wait inside their solvers for as long as the
// rebuild takes -- measured on 2 GPUs: +50 us per step with every tile on its own schedule, more than the cache saves.
#define BP_TILED_PERIOD 8u
__global__ void k_bp_decide(BpView V, Counters* ctr, unsigned n_own, unsigned force_rebuild) {
    if (ctr->overflow | ctr->nan_bounds) return;
    BpState& s = *V.st;
    const float mf = __uint_as_float(ctr->max_fat_bits);   // own bodies (k_integrate) and this step's ghosts (k_ghost_recv)
    const bool sane = mf > 0.0f && mf < 1.0e18f;
    const unsigned nr = ctr->n_ref;
    const bool calm = sane && nr <= V.ref_cap && nr * 8u <= n_own + 64u && nr + BP_OVF_SOFT <= V.ovf_cap;
    const bool ok = calm && !force_rebuild && s.valid && s.n_grid == n_own && s.n_ovf <= BP_OVF_SOFT && s.n_ovf + nr <= V.ovf_cap &&
                    2.0f * mf * 1.001f + s.drift <= s.s0 && s.s_count[s.cur] <= V.s_cap;
    ctr->bp_path = ok ? 1u : (calm ? 2u : 3u);
    if (ok) { s.mode = BP_COHERENT; s.coherent_steps++; s.s_count[s.cur ^ 1u] = 0u; }
    else if (calm) {
        s.mode = BP_REBUILD; s.rebuilds++;
        s.valid = 0u; s.n_ovf = 0u; s.n_grid = n_own; s.s_count[s.cur] = 0u;
        s.drift = 0.5f * mf;
        s.s0 = 2.0f * mf * 1.03f + s.drift;
    } else {
        s.mode = BP_SWEEP; s.valid = 0u;
        float e = (mf + __uint_as_float(ctr->max_tight_bits)) * 1.001f;   // the plain sweep's cell edge (grid_inv_cell)
        s.s0 = (e > 0.0f && e < 3.0e38f) ? e : 1.0f;
    }
}
__global__ void k_bp_finish(BpView V, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) { V.st->valid = 0u; return; }   // the host regrows and re-runs: from scratch
    BpState& s = *V.st;
    if (s.mode == BP_COHERENT) s.cur ^= 1u;
    s.valid = s.mode == BP_SWEEP ? 0u : 1u;
}
// REBUILD: the hashed grid over the OWN bodies' fat boxes, binned by fat centre, cell edge s0.  SWEEP: the same over own bodies
// and ghosts with the plain sweep's cell edge.
template <bool FILL>
__global__ void __launch_bounds__(MGFB_THREADS) k_bp_grid(const Box* __restrict__ fat, const Collider* __restrict__ col, const unsigned* __restrict__ gid,
                                                          unsigned n_own, BodyGrid G, BpView V, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    const unsigned mode = V.st->mode;
    if (mode == BP_COHERENT) return;
    const float inv = 1.0f / V.st->s0;
    const unsigned n = mode == BP_SWEEP ? ctr->n_total : n_own;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        Box b = fat[j];
        unsigned h = bcell_hash(cell_coord(b.c.x, inv), cell_coord(b.c.y, inv), cell_coord(b.c.z, inv), G.table_mask);
        if (!FILL) { atomicAdd(&G.cell_count[h], 1u); V.stale[j] = 0; V.c0[j] = b.c; }
        else {
            unsigned pos = G.cell_start[h] + (atomicSub(&G.cell_count[h], 1u) - 1u);
            G.ent[2 * pos] = make_float4(b.c.x, b.c.y, b.c.z, __uint_as_float(j | ((unsigned)col_kind(col[j]) << 31)));
            G.ent[2 * pos + 1] = make_float4(b.r.x, b.r.y, b.r.z, __uint_as_float(gid[j]));
        }
    }
}
// The (at most 27) buckets around a cell, flattened over the warp: f(ent index, active) is called by ALL lanes the same
// number of times; `active` says whether this lane holds a candidate.
// `slab` = 0..2: only the nine cells with z = cz + slab - 1 (a query split over three warps), -1: all 27.
template <class F>
__device__ __forceinline__ void bp_for_neighbours(const BodyGrid& G, int cx, int cy, int cz, unsigned lane, int slab, F f) {
    unsigned h = 0xffffffffu - lane, e = 0, e1 = 0;
    const unsigned ncell = slab < 0 ? 27u : 9u;
    if (lane < ncell) {
        int x = cx + (int)(lane % 3) - 1, y = cy + (int)((lane / 3) % 3) - 1, z = cz + (slab < 0 ? (int)(lane / 9) : slab) - 1;
        h = bcell_hash(x, y, z, G.table_mask);
    }
    unsigned same = __match_any_sync(0xffffffffu, h);   // two cells in one bucket: the lowest lane walks it
    if (lane < ncell && (unsigned)__ffs((int)same) - 1u == lane) { e = G.cell_start[h]; e1 = G.cell_start[h + 1]; }
    if (e1 < e) e1 = e;
    const unsigned len_l = e1 - e;
    unsigned incl = len_l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += y; }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    for (unsigned t0 = 0; t0 < total; t0 += 32) {
        const unsigned t = t0 + lane;
        unsigned lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            unsigned probe = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1));
            if (probe <= t) lo += step;
        }
        const unsigned src_incl = __shfl_sync(0xffffffffu, incl, (int)lo);
        const unsigned src_len = __shfl_sync(0xffffffffu, len_l, (int)lo);
        const unsigned src_e = __shfl_sync(0xffffffffu, e, (int)lo);
        f(src_e + (t - (src_incl - src_len)), t < total);
    }
}
// append `hit` lanes' pairs to a list through one atomic per warp
__device__ __forceinline__ void bp_warp_append(int2* list, unsigned* count, unsigned cap, bool hit, int2 pr, unsigned lane, Counters* ctr) {
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) {
        unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (pos < cap) list[pos] = pr; else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);
    }
}
// COHERENT: a body whose fat box was replaced this step keeps its grid entry as a locator while it stays within `drift` of where
// it was binned; beyond that it moves to the overflow list
__global__ void __launch_bounds__(MGFB_THREADS) k_bp_mark(const Box* __restrict__ fat, BpView V, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    BpState& s = *V.st;
    if (s.mode != BP_COHERENT) return;
    const unsigned nr = ctr->n_ref;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nr; k += gridDim.x * blockDim.x) {
        const unsigned r = V.ref_list[k];
        if (V.stale[r] == 2) continue;
        const float4 c = fat[r].c, c0 = V.c0[r];
        const float d = fmaxf(fabsf(c.x - c0.x), fmaxf(fabsf(c.y - c0.y), fabsf(c.z - c0.z)));
        if (d <= s.drift * 0.999f) V.stale[r] = 1;
        else {
            V.stale[r] = 2;
            unsigned p = atomicAdd(&s.n_ovf, 1u);
            if (p < V.ovf_cap) V.ovf[p] = r;
            else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);   // more runaways in one step than k_bp_decide left room for: the host re-runs the step from scratch
        }
    }
}
// COHERENT: one pass over S.  Pairs with an endpoint replaced this step are dropped, the rest moves to the other buffer; every
// kept pair is tested exactly and goes to its kind's list.  Slots are claimed per CTA (5 atomics per 1024 pairs).
#define BP_FILTER_ITEMS 4
__global__ void __launch_bounds__(MGFB_THREADS) k_bp_filter(const Box* __restrict__ tight, const Box* __restrict__ fat, BpView V, PairLists lists, unsigned cap,
                                                            Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    BpState& s = *V.st;
    if (s.mode != BP_COHERENT) return;   // (a rebuild's sweep wrote this step's pair lists itself)
    const bool coherent = true;
    const unsigned cur = s.cur, ns = min(s.s_count[cur], V.s_cap);
    const int2* in = V.S[cur]; int2* keep_out = V.S[cur ^ 1u];
    __shared__ unsigned s_cnt[BP_WARPS][5], s_base[BP_WARPS][5];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned per_block = MGFB_THREADS * BP_FILTER_ITEMS;
    for (unsigned base = blockIdx.x * per_block; base < ns; base += gridDim.x * per_block) {   // uniform per block
        int2 pr[BP_FILTER_ITEMS]; unsigned rank_keep[BP_FILTER_ITEMS], rank_hit[BP_FILTER_ITEMS]; int kind[BP_FILTER_ITEMS]; bool keep[BP_FILTER_ITEMS];
        unsigned run[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < BP_FILTER_ITEMS; ++r) {
            const unsigned e = base + (unsigned)r * MGFB_THREADS + threadIdx.x;
            keep[r] = false; kind[r] = -1; pr[r] = make_int2(0, 0);
            if (e < ns) {
                pr[r] = in[e];
                const unsigned i = (unsigned)pr[r].x & 0x7fffffffu, j = (unsigned)pr[r].y & 0x7fffffffu;
                keep[r] = !(coherent && (V.ref_flag[i] | V.ref_flag[j]));
                if (keep[r]) {
                    Box tb = tight[i], fb = fat[j];
                    if (box_overlaps(f4v(tb.c), f4v(tb.r), f4v(fb.c), f4v(fb.r))) kind[r] = (int)(((unsigned)pr[r].x >> 31) * 2u + ((unsigned)pr[r].y >> 31));
                }
            }
            const unsigned mk = __ballot_sync(0xffffffffu, keep[r] && coherent);
            rank_keep[r] = run[4] + (unsigned)__popc(mk & ((1u << lane) - 1u)); run[4] += (unsigned)__popc(mk);
            rank_hit[r] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned m = __ballot_sync(0xffffffffu, kind[r] == k);
                if (kind[r] == k) rank_hit[r] = run[k] + (unsigned)__popc(m & ((1u << lane) - 1u));
                run[k] += (unsigned)__popc(m);
            }
        }
        if (lane < 5) s_cnt[w][lane] = run[lane];
        __syncthreads();
        if (threadIdx.x < 5) {
            unsigned tot = 0;
            for (int ww = 0; ww < BP_WARPS; ++ww) { s_base[ww][threadIdx.x] = tot; tot += s_cnt[ww][threadIdx.x]; }
            unsigned b0 = 0;
            if (tot) b0 = atomicAdd(threadIdx.x < 4 ? &ctr->pairs[threadIdx.x] : &s.s_count[cur ^ 1u], tot);
            for (int ww = 0; ww < BP_WARPS; ++ww) s_base[ww][threadIdx.x] += b0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < BP_FILTER_ITEMS; ++r) {
            if (keep[r] && coherent) {
                unsigned pos = s_base[w][4] + rank_keep[r];
                if (pos < V.s_cap) keep_out[pos] = pr[r]; else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);
            }
            if (kind[r] >= 0) {
                unsigned pos = s_base[w][kind[r]] + rank_hit[r];
                if (pos < cap) lists.p[kind[r]][pos] = make_int2(pr[r].x & 0x7fffffff, pr[r].y & 0x7fffffff);
                else atomicOr(&ctr->overflow, (unsigned)OVF_PAIRS);
            }
        }
        __syncthreads();
    }
}
// Queries: the bodies replaced this step (COHERENT only; their pairs also join S) and, in a tiled world, this step's ghosts
// (COHERENT and REBUILD; straight to the pair lists: a ghost is a different body next step).  Candidates: the cached grid
// (k_bp_query, one warp per query; entries of replaced bodies locate them, every candidate's current boxes come from the
// body arrays in ONE round of gathers) and the overflow list (k_bp_query_ovf, one warp per (query, 32 entries)).
struct BpQuery {
    unsigned q, gq, kq; bool ghost; V3 fc, fr, tc, tr;
};
__device__ __forceinline__ BpQuery bp_load_query(unsigned qk, unsigned nr, unsigned n_own, const BpView& V, const Box* tight, const Box* fat, const Collider* col,
                                                 const unsigned* gid) {
    BpQuery Q;
    Q.ghost = qk >= nr;
    Q.q = Q.ghost ? n_own + (qk - nr) : V.ref_list[qk];
    const Box fq = fat[Q.q], tq = tight[Q.q];
    Q.fc = f4v(fq.c); Q.fr = f4v(fq.r); Q.tc = f4v(tq.c); Q.tr = f4v(tq.r);
    Q.gq = gid[Q.q]; Q.kq = (unsigned)col_kind(col[Q.q]);
    return Q;
}
// candidate j (an own body): does (q, j) belong to S, which way round, and is it in this step's P
__device__ __forceinline__ void bp_consider(const BpQuery& Q, bool from_grid, bool active, unsigned j, unsigned kj, unsigned gj, const Box* tight, const Box* fat, const BpView& V,
                                            int2* s_out, unsigned* s_cnt, const PairLists& lists, unsigned cap, unsigned lane, Counters* ctr) {
    bool in_s = false, in_p = false; int2 pr = make_int2(0, 0); unsigned kind = 0;
    if (active && j != Q.q) {
        // one round of gathers, whichever way the pair turns out
        const Box fj = fat[j], tj = tight[j];
        const unsigned rj = V.ref_flag[j], stj = V.stale[j];
        // (a runaway is the overflow list's business, not the grid's; two bodies replaced in the same step find each other
        // twice: the one with the higher id reports the pair)
        if (!(from_grid && stj == 2u) && !(!Q.ghost && rj && !(Q.gq > gj))) {
            if (gj < Q.gq) {   // q is the pair's i
                pr = make_int2((int)(Q.q | (Q.kq << 31)), (int)(j | (kj << 31))); kind = Q.kq * 2u + kj;
                in_p = box_overlaps(Q.tc, Q.tr, f4v(fj.c), f4v(fj.r));
            } else {
                pr = make_int2((int)(j | (kj << 31)), (int)(Q.q | (Q.kq << 31))); kind = kj * 2u + Q.kq;
                in_p = box_overlaps(f4v(tj.c), f4v(tj.r), Q.fc, Q.fr);
            }
            in_s = in_p || box_overlaps_slop(Q.fc, Q.fr, f4v(fj.c), f4v(fj.r));
        }
    }
    if (!Q.ghost) bp_warp_append(s_out, s_cnt, V.s_cap, in_s, pr, lane, ctr);
    const int2 plain = make_int2(pr.x & 0x7fffffff, pr.y & 0x7fffffff);
#pragma unroll
    for (unsigned k = 0; k < 4; ++k) bp_warp_append(lists.p[k], &ctr->pairs[k], cap, in_p && kind == k, plain, lane, ctr);
}
__global__ void __launch_bounds__(MGFB_THREADS) k_bp_query(const Box* __restrict__ tight, const Box* __restrict__ fat, const Collider* __restrict__ col,
                                                           const unsigned* __restrict__ gid, unsigned n_own, BodyGrid G, BpView V, PairLists lists, unsigned cap,
                                                           Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    BpState& s = *V.st;
    if (s.mode == BP_SWEEP) return;   // (the plain sweep saw the ghosts in its grid)
    const unsigned nr = s.mode == BP_COHERENT ? ctr->n_ref : 0u, ng = ctr->n_total - n_own;
    const float inv = 1.0f / s.s0;
    const unsigned lane = threadIdx.x & 31u, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int2* s_out = V.S[s.cur ^ 1u]; unsigned* s_cnt = &s.s_count[s.cur ^ 1u];
    for (unsigned it = gw; it < (nr + ng) * 3u; it += nw) {   // a query is three work items, one per z slab of its 27 cells
        const unsigned qk = it / 3u; const int slab = (int)(it % 3u);
        const BpQuery Q = bp_load_query(qk, nr, n_own, V, tight, fat, col, gid);
        // the query looks around its own CURRENT fat centre: every body whose fat box overlaps was binned within one cell of
        // it, because the cells are `drift` wider than two fat half extents
        bp_for_neighbours(G, cell_coord(Q.fc.x, inv), cell_coord(Q.fc.y, inv), cell_coord(Q.fc.z, inv), lane, slab, [&](unsigned ee, bool active) {
            unsigned j = 0, kj = 0, gj = 0;
            if (active) {
                const unsigned jw = __float_as_uint(G.ent[2 * ee].w);
                j = jw & 0x7fffffffu; kj = jw >> 31; gj = __float_as_uint(G.ent[2 * ee + 1].w);
            }
            bp_consider(Q, true, active, j, kj, gj, tight, fat, V, s_out, s_cnt, lists, cap, lane, ctr);
        });
    }
}
__global__ void __launch_bounds__(MGFB_THREADS) k_bp_query_ovf(const Box* __restrict__ tight, const Box* __restrict__ fat, const Collider* __restrict__ col,
                                                               const unsigned* __restrict__ gid, unsigned n_own, BpView V, PairLists lists, unsigned cap, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    BpState& s = *V.st;
    if (s.mode == BP_SWEEP) return;
    const unsigned nr = s.mode == BP_COHERENT ? ctr->n_ref : 0u, ng = ctr->n_total - n_own, novf = min(s.n_ovf, V.ovf_cap);
    const unsigned chunks = (novf + 31u) >> 5;
    if (chunks == 0u) return;
    const unsigned lane = threadIdx.x & 31u, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int2* s_out = V.S[s.cur ^ 1u]; unsigned* s_cnt = &s.s_count[s.cur ^ 1u];
    const unsigned long long items = (unsigned long long)(nr + ng) * chunks;
    for (unsigned long long it = gw; it < items; it += nw) {
        const unsigned qk = (unsigned)(it / chunks), t = (unsigned)(it % chunks) * 32u + lane;
        const BpQuery Q = bp_load_query(qk, nr, n_own, V, tight, fat, col, gid);
        const bool active = t < novf;
        unsigned j = 0, kj = 0, gj = 0;
        if (active) { j = V.ovf[t]; kj = (unsigned)col_kind(col[j]); gj = gid[j]; }
        bp_consider(Q, false, active, j, kj, gj, tight, fat, V, s_out, s_cnt, lists, cap, lane, ctr);
    }
}

}  // namespace mgfb
