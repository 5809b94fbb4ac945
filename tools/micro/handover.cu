// Micro-benchmark (development): latency of one dataflow hand-over, L2-mediated vs DSMEM, and cluster co-residency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handover handover.cu && ./handover
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct __align__(32) Box { float4 lo, hi; };
__device__ __forceinline__ void st_box(Box* p, unsigned tag) {
    float f = __uint_as_float(tag);
    asm volatile("st.relaxed.gpu.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(f) : "memory");
}
__device__ __forceinline__ unsigned ld_box(const Box* p) {
    float a, b, c, d, e, f, g, h;
    asm volatile("ld.relaxed.gpu.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(p) : "memory");
    return (__float_as_uint(d) == __float_as_uint(h)) ? __float_as_uint(d) : 0xffffffffu;
}
// CTA 0 and CTA `other` bounce a record through global memory (L2): round trip = 2 hand-overs.  Warp-wide (32 lanes, 32 records).
__global__ void k_l2_pingpong(Box* a, Box* b, unsigned rounds, unsigned other, long long* cycles) {
    if (blockIdx.x != 0 && blockIdx.x != other) return;
    const unsigned lane = threadIdx.x;
    long long t0 = clock64();
    if (blockIdx.x == 0) {
        for (unsigned r = 1; r <= rounds; ++r) {
            st_box(a + lane, r);
            while (!__all_sync(0xffffffffu, ld_box(b + lane) == r)) { }
        }
        if (lane == 0) *cycles = clock64() - t0;
    } else {
        for (unsigned r = 1; r <= rounds; ++r) {
            while (!__all_sync(0xffffffffu, ld_box(a + lane) == r)) { }
            st_box(b + lane, r);
        }
    }
}
// the same bounce between two CTAs of one cluster through distributed shared memory
__device__ __forceinline__ unsigned mapa(unsigned saddr, unsigned rank) {
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_dsmem(unsigned addr, unsigned tag) {
    float f = __uint_as_float(tag);
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "f"(f) : "memory");
    asm volatile("st.shared::cluster.v4.f32 [%0+16], {%1,%1,%1,%1};" ::"r"(addr), "f"(f) : "memory");
}
__global__ void k_dsmem_pingpong(unsigned rounds, unsigned peer_rank, long long* cycles) {
    __shared__ __align__(32) Box box[32];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned lane = threadIdx.x, rank = cl.block_rank();
    box[lane].lo = box[lane].hi = make_float4(0, 0, 0, 0);
    cl.sync();
    if (rank == 0 || rank == peer_rank) {
        const unsigned my = (unsigned)__cvta_generic_to_shared(&box[lane]);
        const unsigned remote = mapa(my, rank == 0 ? peer_rank : 0u);
        long long t0 = clock64();
        for (unsigned r = 1; r <= rounds; ++r) {
            if (rank == 0) st_dsmem(remote, r);
            for (;;) {
                float lw, hw, d0, d1, d2;
                asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(lw) : "r"(my) : "memory");
                asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(hw) : "r"(my) : "memory");
                bool ok = __float_as_uint(lw) == r && __float_as_uint(hw) == r;
                if (__all_sync(0xffffffffu, ok)) break;
            }
            if (rank != 0) st_dsmem(remote, r);
        }
        // rank 0 waits for the echo of round r before sending r+1: handled by polling its own box for tag r above? (rank 0 polls after sending)
        if (rank == 0 && lane == 0 && blockIdx.x < gridDim.x) cycles[blockIdx.x / cl.num_blocks()] = clock64() - t0;
    }
    cl.sync();
}
__global__ void k_dummy(int* x) { extern __shared__ int s[]; if (x) x[0] = s[0]; }

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    Box *a, *b; long long* cyc; cudaMalloc(&a, 32 * sizeof(Box)); cudaMalloc(&b, 32 * sizeof(Box)); cudaMallocManaged(&cyc, 64 * 8);
    const unsigned rounds = 2000;
    for (unsigned other : {1u, 2u, 37u, 74u, 147u}) {
        cudaMemset(a, 0, 32 * sizeof(Box)); cudaMemset(b, 0, 32 * sizeof(Box));
        void* args[] = {&a, &b, (void*)&rounds, &other, &cyc};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_l2_pingpong, dim3(148), dim3(32), args, 0, 0);
        cudaDeviceSynchronize();
        printf("L2 ping-pong CTA 0 <-> CTA %3u: %s  %.0f cycles per hand-over (store -> poll sees it -> store)\n", other, cudaGetErrorString(e), (double)cyc[0] / rounds / 2);
    }
    for (int cs : {2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 4); cfg.blockDim = dim3(32);
        cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cs > 8) cudaFuncSetAttribute(k_dsmem_pingpong, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (unsigned peer : {1u, (unsigned)cs - 1}) {
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_dsmem_pingpong, rounds, peer, cyc);
            cudaError_t e2 = cudaDeviceSynchronize();
            printf("DSMEM ping-pong cluster %2d rank 0 <-> %2u: %s/%s  %.0f cycles per hand-over\n", cs, peer, cudaGetErrorString(e), cudaGetErrorString(e2), (double)cyc[0] / rounds / 2);
        }
    }
    // co-residency of big-smem CTAs in clusters, and cooperative + cluster launch
    for (int cs : {1, 2, 4, 8, 16}) {
        int smem = 200 * 1024;
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (cs > 8) cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 148); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[2]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_dummy, &cfg);
        printf("cluster size %2d, 200 KB smem, 384 threads: max active clusters %d (%d CTAs) [%s]", cs, ncl, ncl * cs, cudaGetErrorString(e));
        if (ncl > 0) {
            cfg.gridDim = dim3(ncl * cs); cfg.numAttrs = 2;
            int* null = nullptr;
            e = cudaLaunchKernelEx(&cfg, k_dummy, null);
            cudaError_t e2 = cudaDeviceSynchronize();
            printf("  cooperative+cluster launch of %d CTAs: %s/%s", ncl * cs, cudaGetErrorString(e), cudaGetErrorString(e2));
        }
        printf("\n");
        cudaGetLastError();
    }
    return 0;
}
