// bvh.cuh -- mgf's BVH<AABB, V> public API (src/bvh.rs:86-369) on the device: insert / remove / get /
// query / raytrace, batched.  SURVEY.md section 8(f) rank 1: the callers OUTSIDE World::step (ray
// picking, Compound, user code); the step path itself sweeps a hashed grid (kernels.cuh) and never
// builds a tree.  Included at the end of capi.cu.
//
// The reference grows its tree one insert at a time (bvh.rs:125-260: best-sibling descent, rotations)
// -- inherently sequential.  Here the tree is REBUILT on the device whenever the leaf set changed:
// Morton codes of the leaf centres -> stable radix sort (CUB, part of the CUDA toolkit) -> an
// implicit complete binary tree over the sorted leaves (node k has children 2k, 2k+1), internal
// boxes by exact min/max, bottom-up.  What callers observe is preserved, not the tree shape:
//   * query(arg)    = every leaf whose box overlaps arg by AABB::overlaps (collision.rs:22-29, closed)
//   * raytrace(arg) = every leaf whose box the particle intersects (collision.rs:202-236), with the Intersection
// as SETS.  The reference can additionally PRUNE a leaf that merely touches (its parents are rounded
// unions, bounds.rs:113-130); internal nodes here are tested conservatively so no leaf passing the
// exact test is ever missed.  Per-query results come in ascending Morton order of the leaves
// (deterministic), not in the reference's DFS order of its own history-dependent tree.  Indices returned
// by insert are stable leaf handles (freed slots are reused last-freed-first like pool.rs:60-98), not the
// reference's pool slots (which also number its internal nodes).
#pragma once
#include <cub/device/device_radix_sort.cuh>

struct mgfb_bvh {
    mgfb_ctx* ctx = nullptr;
    std::vector<unsigned char> alive;     // per slot
    std::vector<unsigned> free_list;      // LIFO
    std::vector<uint32_t> h_value;
    std::vector<float> h_box;             // 6 per slot (c, r)
    unsigned n_alive = 0, cap = 0;        // cap = device capacity in slots
    bool dirty = true;
    Buf d_c, d_r, d_value;                // float4 centre / half extents (w of r: 1 = alive), values
    Buf d_key, d_key2, d_slot, d_slot2, d_tmp, d_lo, d_hi, d_bounds;
    unsigned leaves_pow2 = 0;             // L: the implicit tree has nodes 1 .. 2L-1, leaves L .. 2L-1
    Buf q_in, q_cnt, q_off, q_val, q_hit;
};

namespace {

struct BvhView {
    const float4* c; const float4* r; const unsigned* value;
    const unsigned* slot;      // sorted leaf order -> slot
    const float4* lo; const float4* hi;   // [2L] node boxes (min / max corners)
    unsigned L, n_alive;
};
__device__ __forceinline__ unsigned morton_expand(unsigned v) {   // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void __launch_bounds__(256) k_bvh_bounds(const float4* __restrict__ c, const float4* __restrict__ r, unsigned cap, unsigned* bounds /* 6 ordered-float words */) {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        float4 rr = r[i];
        if (rr.w != 1.0f) continue;
        float4 cc = c[i];
        lo[0] = fminf(lo[0], cc.x); lo[1] = fminf(lo[1], cc.y); lo[2] = fminf(lo[2], cc.z);
        hi[0] = fmaxf(hi[0], cc.x); hi[1] = fmaxf(hi[1], cc.y); hi[2] = fmaxf(hi[2], cc.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], mgfb::ordered_bits(lo[k])); atomicMax(&bounds[3 + k], mgfb::ordered_bits(hi[k])); }
    }
}
__global__ void __launch_bounds__(256) k_bvh_keys(const float4* __restrict__ c, const float4* __restrict__ r, unsigned cap, const unsigned* __restrict__ bounds,
                                                  unsigned* key, unsigned* slot) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    slot[i] = i;
    if (r[i].w != 1.0f) { key[i] = 0xffffffffu; return; }   // dead slots sort last
    float4 cc = c[i];
    float p[3] = {cc.x, cc.y, cc.z}; unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo = mgfb::ordered_float(bounds[k]), hi = mgfb::ordered_float(bounds[3 + k]);
        float span = hi - lo;
        float t = span > 0.0f ? (p[k] - lo) / span : 0.0f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        q[k] = min(1023u, (unsigned)(t * 1023.0f));
    }
    key[i] = (morton_expand(q[0]) << 2) | (morton_expand(q[1]) << 1) | morton_expand(q[2]);
}
// leaves of the implicit tree: node L + j = sorted leaf j (or an empty box beyond the alive leaves)
__global__ void __launch_bounds__(256) k_bvh_leaves(const float4* __restrict__ c, const float4* __restrict__ r, const unsigned* __restrict__ slot, unsigned n_alive,
                                                    unsigned L, float4* lo, float4* hi) {
    unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    float4 l = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.0f), h = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.0f);
    if (j < n_alive) {
        float4 cc = c[slot[j]], rr = r[slot[j]];
        l = make_float4(cc.x - rr.x, cc.y - rr.y, cc.z - rr.z, 0.0f); h = make_float4(cc.x + rr.x, cc.y + rr.y, cc.z + rr.z, 0.0f);
    }
    lo[L + j] = l; hi[L + j] = h;
}
__global__ void __launch_bounds__(256) k_bvh_level(unsigned first, unsigned count, float4* lo, float4* hi) {   // nodes first .. first+count-1
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    unsigned node = first + k;
    float4 a = lo[2 * node], b = lo[2 * node + 1], c = hi[2 * node], d = hi[2 * node + 1];
    lo[node] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.0f);
    hi[node] = make_float4(fmaxf(c.x, d.x), fmaxf(c.y, d.y), fmaxf(c.z, d.z), 0.0f);
}
// Conservative node tests: a leaf that passes the exact reference test can never be pruned above it.
__device__ __forceinline__ bool node_overlaps(V3 qc, V3 qr, float4 lo, float4 hi) {
    float qs[3] = {qc.x, qc.y, qc.z}, rs[3] = {qr.x, qr.y, qr.z}, ls[3] = {lo.x, lo.y, lo.z}, hs[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float slack = (fabsf(qs[k]) + fabsf(rs[k]) + fmaxf(fabsf(ls[k]), fabsf(hs[k]))) * 1e-6f + 1e-30f;
        if (!(qs[k] - rs[k] - slack <= hs[k] && qs[k] + rs[k] + slack >= ls[k])) return false;
    }
    return true;
}
// MODE 0: BVH::query with an AABB (bvh.rs:283-310); MODE 1: BVH::raytrace with a Ray / Segment (bvh.rs:340-369).
// pass 0 counts, pass 1 writes at off[q].
template <int MODE, bool WRITE>
__global__ void __launch_bounds__(128) k_bvh_traverse(BvhView B, const float* __restrict__ in, unsigned nq, unsigned particle_kind, unsigned* cnt,
                                                      const unsigned* __restrict__ off, unsigned* out_val, mgfb_intersection* out_hit, unsigned capacity) {
    unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float* a = in + 6 * (size_t)q;
    V3 p = mk3(a[0], a[1], a[2]), d = mk3(a[3], a[4], a[5]);   // MODE 0: centre, half extents
    float DT = __builtin_huge_valf();
    if (MODE == 1 && particle_kind == MGFB_SEGMENT) { d = d - p; DT = 1.0f; }
    unsigned n = 0, w = WRITE ? off[q] : 0u;
    if (B.n_alive) {
        unsigned stack[64]; int sp = 0;
        stack[sp++] = 1u;
        while (sp) {
            unsigned node = stack[--sp];
            float4 lo = B.lo[node], hi = B.hi[node];
            if (!(lo.x <= hi.x)) continue;   // empty subtree
            if (node >= B.L) {               // leaf: the reference's exact test
                unsigned s = B.slot[node - B.L];
                float4 c4 = B.c[s], r4 = B.r[s];
                V3 lc = mk3(c4.x, c4.y, c4.z), lr = mk3(r4.x, r4.y, r4.z);
                bool hit; float t = 0.0f; V3 ip = zero3();
                if (MODE == 0) hit = mgfb::box_overlaps(p, d, lc, lr);
                else hit = ray_aabb(p, d, lc, lr, DT, &t, &ip);
                if (hit) {
                    if (WRITE && w < capacity) {
                        out_val[w] = B.value[s];
                        if (MODE == 1) { mgfb_intersection o; o.p[0] = ip.x; o.p[1] = ip.y; o.p[2] = ip.z; o.t = t; out_hit[w] = o; }
                    }
                    ++w; ++n;
                }
                continue;
            }
            bool go;
            if (MODE == 0) go = node_overlaps(p, d, lo, hi);
            else {   // the node's box, slightly inflated, through the same slab test
                V3 nc = mk3((lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f);
                V3 nr = mk3((hi.x - lo.x) * 0.5f, (hi.y - lo.y) * 0.5f, (hi.z - lo.z) * 0.5f);
                float s0 = (fabsf(nc.x) + nr.x) * 4e-6f + 1e-30f, s1 = (fabsf(nc.y) + nr.y) * 4e-6f + 1e-30f, s2 = (fabsf(nc.z) + nr.z) * 4e-6f + 1e-30f;
                float t; V3 ip;
                go = ray_aabb(p, d, nc, mk3(nr.x + s0, nr.y + s1, nr.z + s2), __builtin_huge_valf(), &t, &ip);
            }
            if (go && sp <= 62) { stack[sp++] = 2 * node + 1; stack[sp++] = 2 * node; }   // left first -> ascending Morton order
        }
    }
    if (!WRITE) cnt[q] = n;
}

int32_t bvh_grow(mgfb_bvh* b, unsigned need) {
    mgfb_ctx* ctx = b->ctx;
    if (need <= b->cap) return MGFB_OK;
    unsigned nc = std::max(need, std::max(1024u, b->cap * 2));
    TRY(ensure(ctx, b->d_c, (size_t)nc * 16, true)); TRY(ensure(ctx, b->d_r, (size_t)nc * 16, true)); TRY(ensure(ctx, b->d_value, (size_t)nc * 4, true));
    // new slots are dead until written (w of r != 1)
    CU(cudaMemsetAsync(b->d_r.as<float4>() + b->cap, 0, (size_t)(nc - b->cap) * 16, ctx->stream));
    b->cap = nc;
    return MGFB_OK;
}
int32_t bvh_rebuild(mgfb_bvh* b) {
    mgfb_ctx* ctx = b->ctx;
    if (!b->dirty) return MGFB_OK;
    b->dirty = false;
    if (b->n_alive == 0) { b->leaves_pow2 = 0; return MGFB_OK; }
    unsigned cap = b->cap;
    TRY(ensure(ctx, b->d_key, (size_t)cap * 4)); TRY(ensure(ctx, b->d_key2, (size_t)cap * 4));
    TRY(ensure(ctx, b->d_slot, (size_t)cap * 4)); TRY(ensure(ctx, b->d_slot2, (size_t)cap * 4)); TRY(ensure(ctx, b->d_bounds, 32));
    unsigned hb[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CU(cudaMemcpyAsync(b->d_bounds.p, hb, sizeof(hb), cudaMemcpyHostToDevice, ctx->stream));
    int g = std::max(1, std::min((int)((cap + 255) / 256), ctx->num_sms * 8));
    k_bvh_bounds<<<g, 256, 0, ctx->stream>>>(b->d_c.as<float4>(), b->d_r.as<float4>(), cap, b->d_bounds.as<unsigned>());
    k_bvh_keys<<<(cap + 255) / 256, 256, 0, ctx->stream>>>(b->d_c.as<float4>(), b->d_r.as<float4>(), cap, b->d_bounds.as<unsigned>(),
                                                           b->d_key.as<unsigned>(), b->d_slot.as<unsigned>());
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, b->d_key.as<unsigned>(), b->d_key2.as<unsigned>(), b->d_slot.as<unsigned>(),
                                       b->d_slot2.as<unsigned>(), (int)cap, 0, 32, ctx->stream));
    TRY(ensure(ctx, b->d_tmp, std::max<size_t>(tmp_bytes, 16)));
    CU(cub::DeviceRadixSort::SortPairs(b->d_tmp.p, tmp_bytes, b->d_key.as<unsigned>(), b->d_key2.as<unsigned>(), b->d_slot.as<unsigned>(),
                                       b->d_slot2.as<unsigned>(), (int)cap, 0, 32, ctx->stream));
    unsigned L = next_pow2(b->n_alive);
    b->leaves_pow2 = L;
    TRY(ensure(ctx, b->d_lo, (size_t)2 * L * 16)); TRY(ensure(ctx, b->d_hi, (size_t)2 * L * 16));
    k_bvh_leaves<<<(L + 255) / 256, 256, 0, ctx->stream>>>(b->d_c.as<float4>(), b->d_r.as<float4>(), b->d_slot2.as<unsigned>(), b->n_alive, L,
                                                           b->d_lo.as<float4>(), b->d_hi.as<float4>());
    for (unsigned first = L / 2; first >= 1; first /= 2)
        k_bvh_level<<<(first + 255) / 256, 256, 0, ctx->stream>>>(first, first, b->d_lo.as<float4>(), b->d_hi.as<float4>());
    CU(cudaGetLastError());
    ctx->launches += 4 + 2;
    return MGFB_OK;
}
BvhView bvh_view(const mgfb_bvh* b) {
    BvhView V;
    V.c = b->d_c.as<float4>(); V.r = b->d_r.as<float4>(); V.value = b->d_value.as<unsigned>(); V.slot = b->d_slot2.as<unsigned>();
    V.lo = b->d_lo.as<float4>(); V.hi = b->d_hi.as<float4>(); V.L = b->leaves_pow2; V.n_alive = b->n_alive;
    return V;
}
template <int MODE>
int32_t bvh_batch(mgfb_bvh* b, unsigned particle_kind, const float* in, uint32_t nq, uint32_t* offsets, uint32_t* values, mgfb_intersection* hits,
                  uint32_t capacity, uint32_t* total) {
    mgfb_ctx* ctx = b->ctx;
    CU(cudaSetDevice(ctx->device));
    TRY(bvh_rebuild(b));
    if (total) *total = 0;
    if (nq == 0) { if (offsets) offsets[0] = 0; return MGFB_OK; }
    TRY(ensure(ctx, b->q_in, (size_t)nq * 24)); TRY(ensure(ctx, b->q_cnt, ((size_t)nq + 1) * 4)); TRY(ensure(ctx, b->q_off, ((size_t)nq + 1) * 4));
    CU(cudaMemcpyAsync(b->q_in.p, in, (size_t)nq * 24, cudaMemcpyHostToDevice, ctx->stream));
    BvhView V = bvh_view(b);
    unsigned gq = (nq + 127) / 128;
    k_bvh_traverse<MODE, false><<<gq, 128, 0, ctx->stream>>>(V, b->q_in.as<float>(), nq, particle_kind, b->q_cnt.as<unsigned>(), nullptr, nullptr, nullptr, 0);
    TRY(scan_u32_lb(ctx, b->q_cnt.as<unsigned>(), b->q_off.as<unsigned>(), nq, nullptr));
    std::vector<uint32_t> hoff((size_t)nq + 1);
    CU(cudaMemcpyAsync(hoff.data(), b->q_off.p, ((size_t)nq + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uint32_t tot = hoff[nq];
    if (total) *total = tot;
    if (offsets) std::memcpy(offsets, hoff.data(), ((size_t)nq + 1) * 4);
    ctx->launches += 2;
    if (tot > capacity) return fail(ctx, MGFB_ERR_CAPACITY, "output arrays too small (see *total)");
    if (tot == 0) return MGFB_OK;
    TRY(ensure(ctx, b->q_val, (size_t)tot * 4));
    if (MODE == 1) TRY(ensure(ctx, b->q_hit, (size_t)tot * sizeof(mgfb_intersection)));
    k_bvh_traverse<MODE, true><<<gq, 128, 0, ctx->stream>>>(V, b->q_in.as<float>(), nq, particle_kind, nullptr, b->q_off.as<unsigned>(), b->q_val.as<unsigned>(),
                                                           MODE == 1 ? b->q_hit.as<mgfb_intersection>() : nullptr, tot);
    CU(cudaGetLastError());
    if (values) CU(cudaMemcpyAsync(values, b->q_val.p, (size_t)tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (MODE == 1 && hits) CU(cudaMemcpyAsync(hits, b->q_hit.p, (size_t)tot * sizeof(mgfb_intersection), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 1;
    return MGFB_OK;
}
}  // namespace

extern "C" {
int32_t mgfb_bvh_create(mgfb_ctx* ctx, mgfb_bvh** out) {
    if (!ctx || !out) return MGFB_ERR_INVALID_ARG;
    mgfb_bvh* b = new mgfb_bvh(); b->ctx = ctx; *out = b;
    return MGFB_OK;
}
void mgfb_bvh_destroy(mgfb_bvh* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    Buf* all[] = {&b->d_c, &b->d_r, &b->d_value, &b->d_key, &b->d_key2, &b->d_slot, &b->d_slot2, &b->d_tmp, &b->d_lo, &b->d_hi, &b->d_bounds,
                  &b->q_in, &b->q_cnt, &b->q_off, &b->q_val, &b->q_hit};
    for (Buf* x : all) release(*x);
    delete b;
}
int32_t mgfb_bvh_insert(mgfb_bvh* b, const float* boxes, const uint32_t* values, uint32_t n, uint32_t* indices) {
    if (!b || (n && (!boxes || !values || !indices))) return MGFB_ERR_INVALID_ARG;
    mgfb_ctx* ctx = b->ctx;
    for (uint32_t i = 0; i < n; ++i)   // AABB::combine asserts r >= 0 on the way up the reference's tree (bounds.rs:125-127)
        for (int k = 3; k < 6; ++k) if (!(boxes[6 * i + k] >= 0.0f)) return fail(ctx, MGFB_ERR_NAN_BOUNDS, "half extents must be >= 0 and not NaN");
    if (n == 0) return MGFB_OK;
    CU(cudaSetDevice(ctx->device));
    std::vector<float4> hc(n), hr(n);
    for (uint32_t i = 0; i < n; ++i) {
        unsigned s;
        if (!b->free_list.empty()) { s = b->free_list.back(); b->free_list.pop_back(); }
        else { s = (unsigned)b->alive.size(); b->alive.push_back(0); b->h_value.push_back(0); b->h_box.resize(b->h_box.size() + 6); }
        b->alive[s] = 1; b->h_value[s] = values[i];
        std::memcpy(&b->h_box[6 * (size_t)s], boxes + 6 * (size_t)i, 24);
        indices[i] = s;
    }
    TRY(bvh_grow(b, (unsigned)b->alive.size()));
    // contiguous runs of new slots are the common case; upload slot by slot otherwise (host staging kept simple)
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = boxes + 6 * (size_t)i;
        hc[i] = make_float4(q[0], q[1], q[2], 0.0f); hr[i] = make_float4(q[3], q[4], q[5], 1.0f);
    }
    uint32_t i = 0;
    while (i < n) {
        uint32_t j = i + 1;
        while (j < n && indices[j] == indices[j - 1] + 1) ++j;
        CU(cudaMemcpyAsync(b->d_c.as<float4>() + indices[i], hc.data() + i, (size_t)(j - i) * 16, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(b->d_r.as<float4>() + indices[i], hr.data() + i, (size_t)(j - i) * 16, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(b->d_value.as<unsigned>() + indices[i], values + i, (size_t)(j - i) * 4, cudaMemcpyHostToDevice, ctx->stream));
        i = j;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    b->n_alive += n; b->dirty = true;
    return MGFB_OK;
}
int32_t mgfb_bvh_remove(mgfb_bvh* b, const uint32_t* indices, uint32_t n) {
    if (!b || (n && !indices)) return MGFB_ERR_INVALID_ARG;
    mgfb_ctx* ctx = b->ctx;
    // validate EVERYTHING before touching anything (pool.rs:100-113 panics on a free or out-of-range slot; a slot named twice
    // would be removed twice): alive[] doubles as the mark, 2 = named in this call
    int32_t bad = MGFB_OK;
    uint32_t marked = 0;
    for (; marked < n; ++marked) {
        const uint32_t s = indices[marked];
        if (s >= b->alive.size() || !b->alive[s]) { bad = fail(ctx, MGFB_ERR_INVALID_ARG, "no leaf at that index"); break; }
        if (b->alive[s] == 2) { bad = fail(ctx, MGFB_ERR_INVALID_ARG, "leaf removed twice in one call"); break; }
        b->alive[s] = 2;
    }
    if (bad != MGFB_OK) { for (uint32_t i = 0; i < marked; ++i) b->alive[indices[i]] = 1; return bad; }   // nothing was changed
    CU(cudaSetDevice(ctx->device));
    const float4 dead = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (uint32_t i = 0; i < n; ++i) {
        unsigned s = indices[i];
        b->alive[s] = 0; b->free_list.push_back(s);
        CU(cudaMemcpyAsync(b->d_r.as<float4>() + s, &dead, 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    b->n_alive -= n; b->dirty = true;
    return MGFB_OK;
}
int32_t mgfb_bvh_get(const mgfb_bvh* b, uint32_t index, float* box, uint32_t* value) {
    if (!b) return MGFB_ERR_INVALID_ARG;
    if (index >= b->alive.size() || !b->alive[index]) return fail(b->ctx, MGFB_ERR_INVALID_ARG, "no leaf at that index");
    if (box) std::memcpy(box, &b->h_box[6 * (size_t)index], 24);
    if (value) *value = b->h_value[index];
    return MGFB_OK;
}
int32_t mgfb_bvh_len(const mgfb_bvh* b, uint32_t* n) { if (!b || !n) return MGFB_ERR_INVALID_ARG; *n = b->n_alive; return MGFB_OK; }
int32_t mgfb_bvh_query_batch(mgfb_bvh* b, const float* boxes, uint32_t nq, uint32_t* offsets, uint32_t* values, uint32_t capacity, uint32_t* total) {
    if (!b || (nq && !boxes)) return MGFB_ERR_INVALID_ARG;
    return bvh_batch<0>(b, 0, boxes, nq, offsets, values, nullptr, capacity, total);
}
int32_t mgfb_bvh_raytrace_batch(mgfb_bvh* b, uint32_t particle_kind, const float* particles, uint32_t nq, uint32_t* offsets, uint32_t* values,
                                mgfb_intersection* hits, uint32_t capacity, uint32_t* total) {
    if (!b || (nq && !particles)) return MGFB_ERR_INVALID_ARG;
    if (particle_kind > MGFB_SEGMENT) return fail(b->ctx, MGFB_ERR_INVALID_ARG, "particle kind must be MGFB_RAY or MGFB_SEGMENT");
    return bvh_batch<1>(b, particle_kind, particles, nq, offsets, values, hits, capacity, total);
}
}  // extern "C"
