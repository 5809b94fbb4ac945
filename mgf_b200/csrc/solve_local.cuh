// solve_local.cuh -- Solver::solve (solver.rs:72-78), dataflow schedule with SM-LOCAL hand-overs.
//
// k_solve_df (kernels.cuh) deals the constraint rows to warps round-robin over the whole GPU, so every hand-over of a body's
// velocity from one row to the next goes through L2: a 32-byte store, then the consumer's poll.  Measured (profiles/r02_*): the
// solve is bound by that chain, ~220 dependent hand-overs (colours x iterations) at ~3300 cycles each, of which the arithmetic
// is ~1200 -- the rest is L2 latency UNDER THE LOAD of 1500 polling warps (an idle L2 hand-over costs ~550 cycles).
//
// Here the rows are dealt by PLACE: bodies are sorted along a Morton curve and cut into one chunk per SM (k_part_*, refreshed
// every few steps -- any assignment is correct, a stale one only loses locality); a constraint row lives on the SM that is home
// to its first body.  Rows are laid out (SM, colour)-major, each CTA walks its own rows in (iteration, colour) order, and the
// inboxes of a CTA's rows are in its SHARED MEMORY: when the next row of a body's chain lives on the same SM -- the common case
// -- the hand-over is two 16-byte shared-memory stores and the consumer polls shared memory (~100 cycles, no L2 traffic).  Only
// chains that leave the SM use the global inboxes of k_solve_df.  Same rows, same per-body order (colours ascending), same
// arithmetic: the result is bit-identical to k_solve_df / k_solve and to the oracle replaying the exported colour-major order.
//
// Progress: order the warp-rows by (iteration, colour, CTA, index).  A row only waits for rows of a smaller (iteration,
// colour); every warp takes its warp-rows in increasing order and all CTAs are co-resident (cooperative launch), so the
// smallest unfinished warp-row always has its inputs.
#pragma once
#include <cub/device/device_radix_sort.cuh>

namespace mgfb {

#define LOC_COLOURS 64            // == MGFB_DF_MAX_PHASES: more colours than that and k_solve takes the step
#define LOC_SEGS LOC_COLOURS       // segments per CTA: one per colour
#ifndef LOC_THREADS
#define LOC_THREADS 384           // 12 warps per SM
#endif
#define LOC_CAP 3072u             // rows per CTA whose inboxes (2 x 32 B) and accumulated impulse (4 B) fit in shared memory (204 KB); rows beyond use the global arrays
#define LOC_SMEM_BYTES (LOC_CAP * 68u)
#define LOC_WHERE 3u              // `where` code of a successor link (kernels.cuh): the successor's inbox is in this CTA's shared memory

// ---------------------------------------------------------------- partition: body -> home CTA
__device__ __forceinline__ unsigned spread10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void __launch_bounds__(MGFB_THREADS) k_part_bbox(const float4* __restrict__ x, unsigned n, unsigned* bb /* [6] ordered bits: min xyz, max xyz */) {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = x[i];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bb[a], ordered_bits(lo[a])); atomicMax(&bb[3 + a], ordered_bits(hi[a])); }
    }
}
__global__ void __launch_bounds__(MGFB_THREADS) k_part_keys(const float4* __restrict__ x, unsigned n, const unsigned* __restrict__ bb, unsigned* key, unsigned* val) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = x[i];
    float c[3] = {p.x, p.y, p.z};
    unsigned q[3];
    // one cubic lattice for the three axes (cells of the same size), anchored at the box's lower corner
    float ext = 1e-20f;
    for (int a = 0; a < 3; ++a) ext = fmaxf(ext, ordered_float(bb[3 + a]) - ordered_float(bb[a]));
    for (int a = 0; a < 3; ++a) {
        float t = (c[a] - ordered_float(bb[a])) / ext * 1023.0f;
        q[a] = (t >= 0.0f && t < 1024.0f) ? (unsigned)t : (t >= 1024.0f ? 1023u : 0u);   // NaN -> 0
    }
    key[i] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    val[i] = i;
}
// sorted position -> chunk; chunks hold equal WEIGHT, weight of a body = 1 + the rows it was home to last time (0 at first)
__global__ void __launch_bounds__(MGFB_THREADS) k_part_weights(const unsigned* __restrict__ sorted_body, const unsigned* __restrict__ owned, unsigned n, unsigned* w) {
    unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) w[s] = 1u + (owned ? owned[sorted_body[s]] : 0u);
}
__global__ void __launch_bounds__(MGFB_THREADS) k_part_assign(const unsigned* __restrict__ sorted_body, const unsigned* __restrict__ wprefix /* [n+1] exclusive */, unsigned n,
                                                             unsigned G, unsigned short* home) {
    unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned long long total = wprefix[n];
    unsigned c = (unsigned)(((unsigned long long)wprefix[s] * G) / (total ? total : 1ULL));
    home[sorted_body[s]] = (unsigned short)min(c, G - 1u);
}

// ---------------------------------------------------------------- rows laid out (CTA, colour)-major
struct LocalView {
    const unsigned short* home;   // [nbodies]
    const unsigned short* cta;    // [m] CTA of every constraint's row (k_inc_fill)
    unsigned* hist;               // [G * LOC_SEGS] rows per (CTA, colour), counted by k_colour_df; consumed by the scan
    unsigned* seg_start;          // [G * LOC_SEGS + 1] exclusive scan of hist: first row of every segment
    unsigned* slot;               // [m] rank of a constraint inside its segment (k_colour_df)
    unsigned* flags;              // [row] LOCF_* (which inputs arrive through shared memory)
    unsigned* owned;              // [nbodies] rows per home body of this step (weights of the next partition)
    unsigned G;
};
enum { LOCF_A = 1, LOCF_B = 2, LOCF_FIRST_A = 4, LOCF_FIRST_B = 8 };
// (the rows per (CTA, colour) segment are counted by k_colour_df as it colours: ColourView::seg_hist)
__global__ void __launch_bounds__(MGFB_THREADS) k_loc_scatter(const int* __restrict__ ca, const int* __restrict__ cb, const int* __restrict__ group, LocalView V,
                                                             unsigned* perm, const unsigned* m_ptr, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (ctr->ngroups > LOC_COLOURS || ctr->colour_fallback) return;
    const unsigned m = *m_ptr;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
        const int a = ca[k], b = cb[k];
        const unsigned c = V.cta[k];
        perm[V.seg_start[c * LOC_SEGS + (unsigned)group[k]] + V.slot[k]] = k;
        atomicAdd(&V.owned[(b < 0 || c == V.home[a]) ? a : b], 1u);   // weights of the next partition
    }
}
// which CTA holds row r: the last c with seg_start[c * LOC_SEGS] <= r
__device__ __forceinline__ unsigned loc_cta_of(const unsigned* __restrict__ seg_start, unsigned G, unsigned r) {
    unsigned lo = 0, hi = G;   // invariant: start(lo) <= r < start(hi)
    while (hi - lo > 1) { unsigned mid = (lo + hi) >> 1; if (seg_start[mid * LOC_SEGS] <= r) lo = mid; else hi = mid; }
    return lo;
}
// Successor links like k_df_init, plus: a hand-over whose two rows live on one CTA (and fit its shared memory) is marked LOCAL.
__global__ void __launch_bounds__(MGFB_THREADS) k_loc_init(const BodyVel* __restrict__ vel, unsigned n, const int2* __restrict__ ab, DfArrays D, LocalView V,
                                                          unsigned epoch, const unsigned* m_ptr, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (ctr->ngroups > LOC_COLOURS || ctr->colour_fallback) return;   // k_solve takes this step
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const unsigned m = *m_ptr;
    // bodies: v, omega into the GLOBAL inbox of the body's first row, tagged for iteration 0 (the row reads it there at it == 0)
    for (unsigned i = tid; i < n; i += nth) {
        unsigned s0 = D.body_start[i], s1 = D.body_start[i + 1];
        if (s0 == s1) continue;
        unsigned fr = D.inc[s0];
        const float4* q = reinterpret_cast<const float4*>(vel + i);
        float4 a = q[0], b = q[1];
        const bool side_b = ab[fr].x != (int)i;
        st_inbox<false>((side_b ? D.in_b : D.in_a) + fr, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), epoch + 1u);
        atomicOr(&V.flags[fr], side_b ? (unsigned)LOCF_FIRST_B : (unsigned)LOCF_FIRST_A);
    }
    unsigned loc = 0, glob = 0;
    for (unsigned row = tid; row < m; row += nth) {
        int2 p = ab[row]; unsigned d = D.dep[row];
        unsigned seq[2] = {d & 255u, (d >> 16) & 255u}, deg[2] = {(d >> 8) & 255u, d >> 24};
        int body[2] = {p.x, p.y};
        const unsigned my_cta = loc_cta_of(V.seg_start, V.G, row);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unsigned nx = DF_NONE;
            const int pb = body[s];
            if (pb >= 0) {
                const bool wrap = seq[s] + 1u == deg[s];
                const unsigned nr = D.inc[D.body_start[pb] + (wrap ? 0u : seq[s] + 1u)];
                const unsigned nside = ab[nr].x == pb ? 0u : 1u;
                const unsigned base = V.seg_start[my_cta * LOC_SEGS], end = V.seg_start[(my_cta + 1) * LOC_SEGS];
                if (nr >= base && nr < end && nr - base < LOC_CAP) {
                    nx = ((nr - base) << 4) | (LOC_WHERE << 2) | (wrap ? 2u : 0u) | nside;
                    atomicOr(&V.flags[nr], nside ? (unsigned)LOCF_B : (unsigned)LOCF_A);
                    ++loc;
                } else {
                    nx = (nr << 4) | (wrap ? 2u : 0u) | nside;
                    ++glob;
                }
            }
            D.next[s * D.row_cap + row] = nx;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { loc += __shfl_xor_sync(0xffffffffu, loc, o); glob += __shfl_xor_sync(0xffffffffu, glob, o); }
    if ((threadIdx.x & 31) == 0 && (loc | glob)) { atomicAdd(&ctr->loc_edges, loc); atomicAdd(&ctr->glob_edges, glob); }
}

// ---------------------------------------------------------------- the solver
__device__ __forceinline__ void lds_inbox(unsigned saddr, float4* lo, float4* hi) {
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lo->x), "=f"(lo->y), "=f"(lo->z), "=f"(lo->w) : "r"(saddr) : "memory");
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(hi->x), "=f"(hi->y), "=f"(hi->z), "=f"(hi->w) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void sts_inbox(unsigned saddr, V3 v, V3 w, unsigned tag) {
    const float ft = __uint_as_float(tag);
    asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(ft) : "memory");
    asm volatile("st.volatile.shared.v4.f32 [%0+16], {%1,%2,%3,%4};" ::"r"(saddr), "f"(w.x), "f"(w.y), "f"(w.z), "f"(ft) : "memory");
}
#ifdef MGFB_DF_PROFILE
__device__ unsigned long long g_loc_prof[8];   // visits, polls, cycles: fetch, poll, compute+publish ; warps
#endif
struct LocRow { float4 n4, t04, t14, ra4, rb4, i0, i1, i2, i3, i4; unsigned na, nb, fl, row; bool valid; };
__global__ void __launch_bounds__(LOC_THREADS, 1) k_solve_loc(ConstraintRows R, DfArrays D, LocalView V, BodyVel* vel, unsigned iters, unsigned epoch, Counters* ctr) {
    extern __shared__ __align__(32) unsigned char loc_smem[];   // Inbox[LOC_CAP][2]: side a, side b of every local row; then float[LOC_CAP]: accumulated impulses
    if (ctr->overflow | ctr->nan_bounds) return;
    if (ctr->ngroups > LOC_COLOURS || ctr->colour_fallback || ctr->n_phases == 0 || iters == 0) return;   // > 63 colours: k_solve takes the step
    __shared__ unsigned s_row0[LOC_SEGS + 1], s_wr0[LOC_SEGS + 1];
    const unsigned* seg = V.seg_start + blockIdx.x * LOC_SEGS;
    if (threadIdx.x == 0) {
        unsigned w = 0;
        for (unsigned p = 0; p < LOC_SEGS; ++p) { unsigned r0 = seg[p], r1 = seg[p + 1]; s_row0[p] = r0; s_wr0[p] = w; w += (r1 - r0 + 31u) >> 5; }
        s_row0[LOC_SEGS] = seg[LOC_SEGS]; s_wr0[LOC_SEGS] = w;
    }
    {   // tags start at 0 = "nothing here"; impulses start at 0 (k_build_rows zeroes the global accumulator too)
        const unsigned rows_here = min(seg[LOC_SEGS] - seg[0], LOC_CAP);
        float4* z = reinterpret_cast<float4*>(loc_smem);
        for (unsigned i = threadIdx.x; i < rows_here * 4u; i += blockDim.x) z[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float* zi = reinterpret_cast<float*>(loc_smem + (size_t)LOC_CAP * 64u);
        for (unsigned i = threadIdx.x; i < rows_here; i += blockDim.x) zi[i] = 0.0f;
    }
    __syncthreads();
    const unsigned nwr = s_wr0[LOC_SEGS], base = s_row0[0];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nW = blockDim.x >> 5;
    if (warp >= nwr) return;
    const unsigned rc = D.row_cap;
    const unsigned sbox = (unsigned)__cvta_generic_to_shared(loc_smem);
    float* s_imp = reinterpret_cast<float*>(loc_smem + (size_t)LOC_CAP * 64u);
#ifdef MGFB_DF_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    // The immutable part of the NEXT visit's rows is loaded into a second register set before this visit starts to wait, so a
    // visit whose inputs are already there costs no L2 round trip at all.
    unsigned pf_wr = warp, pf_p = 0;
    auto fetch = [&]() {
        LocRow f;
        while (pf_wr >= s_wr0[pf_p + 1]) ++pf_p;
        f.row = s_row0[pf_p] + ((pf_wr - s_wr0[pf_p]) << 5) + lane;
        f.valid = f.row < s_row0[pf_p + 1];
        if (f.valid) {
            const unsigned row = f.row;
            f.n4 = R.n[row]; f.t04 = R.t0[row]; f.t14 = R.t1[row]; f.ra4 = R.ra[row]; f.rb4 = R.rb[row];
            f.i0 = D.ia[row]; f.i1 = D.ia[rc + row]; f.i2 = D.ia[2 * rc + row]; f.i3 = D.ia[3 * rc + row]; f.i4 = D.ia[4 * rc + row];
            f.na = D.next[row]; f.nb = D.next[rc + row]; f.fl = V.flags[row];
        } else {
            f.n4 = f.t04 = f.t14 = f.ra4 = f.rb4 = f.i0 = f.i1 = f.i2 = f.i3 = f.i4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            f.na = f.nb = DF_NONE; f.fl = 0;
        }
        pf_wr += nW;
        if (pf_wr >= nwr) { pf_wr = warp; pf_p = 0; }   // next iteration: the same rows again
        return f;
    };
    LocRow nxt = fetch();
    for (unsigned it = 0; it < iters; ++it) {
        const bool last_it = it + 1 == iters;
        const unsigned gtag = epoch + it + 1u, ltag = it + 1u;
        for (unsigned wr = warp; wr < nwr; wr += nW) {
            DF_T(t0);
            const LocRow cur = nxt;
            if (!(last_it && wr + nW >= nwr)) nxt = fetch();
            const unsigned row = cur.row;
            const bool valid = cur.valid;
            const unsigned na = cur.na, nb = cur.nb, fl = cur.fl;
            const bool needA = na != DF_NONE, needB = nb != DF_NONE;
            // where each input arrives: this CTA's shared memory, or the global inbox (another CTA's row, or the seed of iteration 0)
            const bool locA = (fl & LOCF_A) && !(it == 0 && (fl & LOCF_FIRST_A)), locB = (fl & LOCF_B) && !(it == 0 && (fl & LOCF_FIRST_B));
            const unsigned lidx = row - base;
            const unsigned my_box = sbox + lidx * 64u;
            Inbox sa, sb;
            sa.lo = sa.hi = sb.lo = sb.hi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            bool okA = !needA, okB = !needB;
            DF_T(t1); DF_ACC(2, t1 - t0);
            for (;;) {
                if (!okA) {
                    if (locA) { lds_inbox(my_box, &sa.lo, &sa.hi); okA = __float_as_uint(sa.lo.w) == ltag && __float_as_uint(sa.hi.w) == ltag; }
                    else { sa = ld_inbox<false>(D.in_a + row); okA = __float_as_uint(sa.lo.w) == gtag && __float_as_uint(sa.hi.w) == gtag; }
                }
                if (!okB) {
                    if (locB) { lds_inbox(my_box + 32u, &sb.lo, &sb.hi); okB = __float_as_uint(sb.lo.w) == ltag && __float_as_uint(sb.hi.w) == ltag; }
                    else { sb = ld_inbox<false>(D.in_b + row); okB = __float_as_uint(sb.lo.w) == gtag && __float_as_uint(sb.hi.w) == gtag; }
                }
                DF_ACC(1, 1);
#ifdef MGFB_DF_NOWAIT
                break;   // timing experiment only (wrong results): the cost of the kernel without any waiting
#endif
                if (__all_sync(0xffffffffu, okA && okB)) break;
            }
            DF_T(t2); DF_ACC(3, t2 - t1);
            if (valid) {
                V3 va = mk3(sa.lo.x, sa.lo.y, sa.lo.z), oa = mk3(sa.hi.x, sa.hi.y, sa.hi.z);
                V3 vb = mk3(sb.lo.x, sb.lo.y, sb.lo.z), ob = mk3(sb.hi.x, sb.hi.y, sb.hi.z);
                const float4 i0 = cur.i0, i1 = cur.i1, i2 = cur.i2, i3 = cur.i3, i4 = cur.i4;
                const M3 IA = mkm(mk3(i0.x, i0.y, i0.z), mk3(i0.w, i1.x, i1.y), mk3(i1.z, i1.w, i2.x));
                const float ima = i2.y;
                const M3 IB = mkm(mk3(i2.z, i2.w, i3.x), mk3(i3.y, i3.z, i3.w), mk3(i4.x, i4.y, i4.z));
                const float imb = i4.w;
                const float4 n4 = cur.n4, t04 = cur.t04, t14 = cur.t14, ra4 = cur.ra4, rb4 = cur.rb4;
                V3 n = f4v(n4), t0v = f4v(t04), t1v = f4v(t14);
                const int nc = (int)fbits(rb4.w);
                const bool imp_local = lidx < LOC_CAP;
                for (int c = 0; c < nc; ++c) {
                    V3 ra, rb; float bias, nmass, tm0, tm1, acc;
                    if (c == 0) {
                        ra = f4v(ra4); rb = f4v(rb4); bias = n4.w; nmass = ra4.w; tm0 = t04.w; tm1 = t14.w;
                        acc = imp_local ? s_imp[lidx] : (it ? __ldcg(&R.impulse[row]) : 0.0f);
                    } else {
                        unsigned e = row * 3 + (c - 1);
                        float4 xa = R.xra[e], xb = R.xrb[e], xt = __ldcg(&R.xtm[e]);
                        ra = f4v(xa); rb = f4v(xb); nmass = xa.w; bias = xb.w; tm0 = xt.x; tm1 = xt.y; acc = xt.z;
                    }
                    V3 dv = vb + cross3(ob, rb) - va - cross3(oa, ra);        // solver.rs:217-232, same stale dv for both tangents
                    float l0 = -dot3(dv, t0v) * tm0;
                    apply_impulse(t0v * l0, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                    float l1 = -dot3(dv, t1v) * tm1;
                    apply_impulse(t1v * l1, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                    V3 dv2 = vb + cross3(ob, rb) - va - cross3(oa, ra);       // solver.rs:234-247
                    float vn = dot3(dv2, n);
                    float lambda = nmass * (-vn + bias);
                    float prev = acc;
                    acc = fmaxf(prev + lambda, 0.0f);
                    lambda = acc - prev;
                    apply_impulse(n * lambda, ra, rb, ima, imb, IA, IB, va, oa, vb, ob);
                    if (c == 0) {
                        if (imp_local && !last_it) s_imp[lidx] = acc;
                        else __stcg(&R.impulse[row], acc);
                    }
                    else { unsigned e = row * 3 + (c - 1); __stcg(&R.xtm[e], make_float4(tm0, tm1, acc, 0.0f)); }
                }
                int a = 0, b = 0;
                if (last_it) { int2 ab = R.ab[row]; a = ab.x; b = ab.y; }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const unsigned nx = s ? nb : na;
                    if (nx == DF_NONE) continue;
                    const V3 v = s ? vb : va, w = s ? ob : oa;
                    const bool wrap = (nx & 2u) != 0u;
                    if (wrap && last_it) { store_vel(vel + (s ? b : a), v, w, s ? imb : ima, s ? IB : IA); continue; }
                    const unsigned nside = nx & 1u, nr = nx >> 4;
                    if (((nx >> 2) & 3u) == LOC_WHERE) sts_inbox(sbox + nr * 64u + nside * 32u, v, w, ltag + (wrap ? 1u : 0u));
                    else st_inbox<false>((nside ? D.in_b : D.in_a) + nr, v, w, gtag + (wrap ? 1u : 0u));
                }
            }
            __syncwarp();
            DF_T(t3); DF_ACC(4, t3 - t2); DF_ACC(0, 1);
        }
    }
#ifdef MGFB_DF_PROFILE
    if (lane == 0) { for (int i = 0; i < 5; ++i) atomicAdd(&g_loc_prof[i], prof[i]); atomicAdd(&g_loc_prof[5], 1ULL); }
#endif
}

}  // namespace mgfb
