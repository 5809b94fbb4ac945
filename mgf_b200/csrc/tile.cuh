// tile.cuh -- one world tiled across GPUs (no counterpart in mgf, which is
// single-threaded).  One process (or context) per GPU owns a slab of bodies; neighbours are
// reached through PEER-MAPPED device memory over NVLink (cudaIpc handles, or plain pointers when
// two contexts share a process), never through the host:
//
//   step start   rank r tells r+1 how far right its bodies reach (one float + flag);
//                r+1 writes every owned body whose stored fat box reaches that far straight
//                into r's body arrays behind r's own bodies ("ghosts") and raises a flag.
//   narrowphase  r finds own-own and own-ghost pairs (ghost-ghost pairs belong to r+1).
//   solve        every iteration: interior colours | r+1 pushes the velocities of its edge
//                bodies into r's ghost slots | r solves the boundary colours | r writes the
//                ghost velocities back into r+1's records.  All four inside the one persistent
//                solver kernel, flags with system-scope release/acquire.
//
// The executed order is a valid sequential Gauss-Seidel order (per iteration: all interior
// constraints rank by rank, then all boundary constraints rank by rank), so the oracle can
// replay it and the tiled world is checked bit for bit like the single-GPU one.
#pragma once
#include "layout.cuh"

namespace mgfb {

#define SELFTEST_PAIRS 16u   // writer/reader CTA pairs of the hand-over self-test (selftest.cuh)

// Lives in each rank's own memory; only the neighbours write to it (each field has one writer).
struct __align__(32) TileSlot { unsigned long long flag; unsigned u; float f; unsigned pad[4]; };
struct TileMailbox {
    TileSlot xr;              // from the LEFT neighbour: flag = step, f = its right extent max(fat.c.x + fat.r.x)
    TileSlot ghosts;          // from the RIGHT neighbour: flag = step, u = number of ghosts written
    TileSlot vel_from_right;  // from the RIGHT neighbour: flag = V(step, 1 + it): edge velocities of iteration it are in my ghost slots
    TileSlot vel_from_left;   // from the LEFT neighbour: flag = V(step, 1): it may be sent velocities; V(step, 2 + it): my edge bodies hold its results
    // dataflow solve (k_solve_df<true>): the chains of edge bodies run across the tile boundary
    TileSlot links_from_left;   // flag = step: link_l[] holds, per edge slot, the first boundary row of that body on the left tile
    TileSlot links_from_right;  // flag = step: link_r[] holds, per ghost slot, the first interior row of that body on its owner
    TileSlot done_from_left;    // flag = step: the left tile's solver has finished (the final velocities of my edge bodies are in my records)
};
// A neighbour's arrays as mapped into this process.
struct TilePeer {
    float4* x; BodyVel* vel; float4* force; float4* torque; Collider* col; Box* tight; Box* fat; unsigned* gid;
    unsigned* ridx;           // [ghost_cap] for each of ITS ghosts: the index of that body on the owner (me)
    TileMailbox* mbox;
    unsigned n_own, ghost_cap;
    // dataflow solve: the neighbour's row inboxes and the link tables it reads
    Inbox* in_a; Inbox* in_b;
    unsigned* link_l;         // [ghost_cap] written by ITS left neighbour (me, if I am that)
    unsigned* link_r;         // [ghost_cap] written by ITS right neighbour
    Inbox* selftest;          // 2 x SELFTEST_PAIRS x 32 scratch records: hand-over self-test of the link (selftest.cuh)
};
struct TileLink {
    TilePeer left, right;
    TileMailbox* mine;
    unsigned* edge_idx;       // [n_edge] my bodies that are ghosts on the left neighbour, in ghost-slot order
    unsigned char* edge_mark; // [n_own] 1 = sent left this step
    unsigned* ridx;           // [ghost_cap] owner index of each of my ghosts (written by the right neighbour)
    unsigned* edge_slot;      // [n_own] ghost slot on the left neighbour of each of my edge bodies (valid where edge_mark)
    unsigned* link_l;         // [ghost_cap] mine, written by the left neighbour
    unsigned* link_r;         // [ghost_cap] mine, written by the right neighbour
    unsigned n_own, ghost_cap;
    unsigned long long step;  // 1, 2, 3, ... the same on every rank
    unsigned long long timeout_ns;
    int has_left, has_right;
};
__host__ __device__ __forceinline__ unsigned long long tile_seq(unsigned long long step, unsigned k) { return (step << 16) | k; }

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// One thread waits until *flag >= target.  Gives up (sticky comm_error) after timeout_ns so a
// dead or diverged neighbour can never hang the GPU.
__device__ __forceinline__ void tile_spin(const unsigned long long* flag, unsigned long long target, unsigned long long timeout_ns, Counters* ctr) {
    if (ld_acquire_sys_u64(flag) >= target) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys_u64(flag) < target) {
        if ((++spins & 255u) == 0u) {
            if (*reinterpret_cast<volatile unsigned*>(&ctr->comm_error) & COMM_TIMEOUT) return;
            if (globaltimer_ns() - t0 > timeout_ns) { atomicOr(&ctr->comm_error, (unsigned)COMM_TIMEOUT); return; }
        }
    }
}
// Every CTA of a persistent kernel waits (thread 0 polls, the CTA barrier publishes).
__device__ __forceinline__ void tile_wait_cta(const unsigned long long* flag, unsigned long long target, unsigned long long timeout_ns, Counters* ctr) {
    if (threadIdx.x == 0) tile_spin(flag, target, timeout_ns, ctr);
    __syncthreads();
}

// order-preserving map float -> unsigned (for atomicMax on possibly negative coordinates)
HD unsigned ordered_bits(float f) {
#ifdef __CUDA_ARCH__
    unsigned u = __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; unsigned u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_float(unsigned u) {
    if (u == 0u) return -3.0e38f;   // nothing recorded
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- step start: extents right, ghosts left
__global__ void k_tile_publish(TileLink T, Counters* ctr) {
    if (!T.has_right) return;
    T.right.mbox->xr.f = ordered_float(ctr->xr_bits);
    __threadfence_system();
    st_release_sys_u64(&T.right.mbox->xr.flag, T.step);
}
__global__ void k_tile_wait(const unsigned long long* flag, unsigned long long target, unsigned long long timeout_ns, Counters* ctr) {
    tile_spin(flag, target, timeout_ns, ctr);
}
// Runs after k_tile_wait(mine->xr): every owned body whose stored fat box reaches the left
// neighbour's right extent is written into the neighbour's arrays at n_own_left + k.
__global__ void __launch_bounds__(256) k_ghost_send(BodyArrays B, TileLink T, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds | ctr->comm_error) return;
    // every CTA waits for the left neighbour's extent itself (one launch less than a wait kernel in front)
    tile_wait_cta(&T.mine->xr.flag, T.step, T.timeout_ns, ctr);
    if (*reinterpret_cast<volatile unsigned*>(&ctr->comm_error)) return;
    const float xr = T.mine->xr.f;
    const TilePeer& P = T.left;
    bool wrote = false;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < T.n_own; i += gridDim.x * blockDim.x) {
        Box fb = B.fat[i];
        float lo = fb.c.x - fb.r.x;
        float eps = (fabsf(fb.c.x) + fb.r.x) * 4e-6f + 1e-30f;   // never miss a closed, rounded overlap (collision.rs:22-29)
        if (!(lo - eps <= xr)) continue;
        unsigned k = atomicAdd(&ctr->n_edge, 1u);
        if (k >= P.ghost_cap) { atomicOr(&ctr->overflow, (unsigned)OVF_GHOSTS); continue; }
        T.edge_idx[k] = i;
        T.edge_mark[i] = 1;
        T.edge_slot[i] = k;
        unsigned d = P.n_own + k;
        P.x[d] = B.x[i];
        const float4* sv = reinterpret_cast<const float4*>(B.vel + i);
        float4* dv = reinterpret_cast<float4*>(P.vel + d);
        dv[0] = sv[0]; dv[1] = sv[1]; dv[2] = sv[2]; dv[3] = sv[3];
        P.force[d] = B.force[i]; P.torque[d] = B.torque[i];
        P.col[d] = B.col[i]; P.tight[d] = B.tight[i]; P.fat[d] = fb;
        P.gid[d] = B.gid[i];
        P.ridx[k] = i;
        wrote = true;
    }
    // last block done -> publish the count.  Only warps that wrote to the neighbour pay for a system-scope fence (a few
    // hundred of 100 k bodies are edge bodies); the chain writer fence -> CTA barrier -> blocks_done -> last block's fence +
    // release orders every ghost record before the flag (fences are cumulative).
    if (__any_sync(0xffffffffu, wrote)) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned done = atomicAdd(&ctr->blocks_done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            unsigned n = min(*reinterpret_cast<volatile unsigned*>(&ctr->n_edge), P.ghost_cap);
            P.mbox->ghosts.u = n;
            __threadfence_system();
            st_release_sys_u64(&P.mbox->ghosts.flag, T.step);
        }
    }
}
// Runs after k_tile_wait(mine->ghosts) (or directly when there is no right neighbour): fixes
// n_total and folds the ghosts' extents into the grid cell size.
__global__ void __launch_bounds__(256) k_ghost_recv(BodyArrays B, TileLink T, Counters* ctr) {
    if (ctr->overflow | ctr->nan_bounds) return;
    if (T.has_right) tile_wait_cta(&T.mine->ghosts.flag, T.step, T.timeout_ns, ctr);   // the right neighbour's ghosts are in
    unsigned ng = T.has_right ? min(T.mine->ghosts.u, T.ghost_cap) : 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->n_total = T.n_own + ng;
    float my_fat = 0.0f, my_tight = 0.0f;
    const float xl = T.has_left ? T.mine->xr.f : -3.0e38f;
    bool thin = false;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < ng; k += gridDim.x * blockDim.x) {
        Box fb = B.fat[T.n_own + k], tb = B.tight[T.n_own + k];
        my_fat = fmaxf(my_fat, fmaxf(fb.r.x, fmaxf(fb.r.y, fb.r.z)));
        my_tight = fmaxf(my_tight, fmaxf(tb.r.x, fmaxf(tb.r.y, tb.r.z)));
        // a ghost that also reaches my LEFT neighbour's bodies would need a pair two tiles apart
        float lo = fb.c.x - fb.r.x, eps = (fabsf(fb.c.x) + fb.r.x) * 4e-6f + 1e-30f;
        if (lo - eps <= xl) thin = true;
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_fat = fmaxf(my_fat, __shfl_xor_sync(0xffffffffu, my_fat, o));
        my_tight = fmaxf(my_tight, __shfl_xor_sync(0xffffffffu, my_tight, o));
    }
    if ((threadIdx.x & 31) == 0 && my_fat > 0.0f) {
        atomicMax(&ctr->max_fat_bits, __float_as_uint(my_fat));
        atomicMax(&ctr->max_tight_bits, __float_as_uint(my_tight));
    }
    if (thin) atomicOr(&ctr->comm_error, (unsigned)COMM_TILE_TOO_THIN);
}

// ---- inside the persistent solver: velocities across the boundary (32 B per body)
__device__ __forceinline__ void copy_vw(BodyVel* dst, const BodyVel* src) {
    const float4* s = reinterpret_cast<const float4*>(src);
    float4* d = reinterpret_cast<float4*>(dst);
    float4 a = __ldcg(s), b = __ldcg(s + 1);
    __stcg(d, a); __stcg(d + 1, b);   // b carries inv_mass and I.c0.x too: identical on both sides
}
// my edge bodies -> the left neighbour's ghost slots
__device__ __forceinline__ void tile_push_edges(const TileLink& T, const BodyVel* vel, unsigned n_edge, unsigned tid, unsigned nth) {
    for (unsigned k = tid; k < n_edge; k += nth) copy_vw(T.left.vel + T.left.n_own + k, vel + T.edge_idx[k]);
    __threadfence_system();
}
// my ghost slots -> the right neighbour's own records
__device__ __forceinline__ void tile_return_ghosts(const TileLink& T, const BodyVel* vel, unsigned n_ghost, unsigned tid, unsigned nth) {
    for (unsigned k = tid; k < n_ghost; k += nth) copy_vw(T.right.vel + T.ridx[k], vel + T.n_own + k);
    __threadfence_system();
}

}  // namespace mgfb
