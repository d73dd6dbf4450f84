// dsmem_bench.cu -- what does one dependent exchange between the CTAs of a thread-block cluster cost on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu && ./dsmem_bench
// Every CTA (256 threads) sends S bytes to its partner (rank ^ 1) and waits for the partner's S bytes, ITER times in a
// dependent chain (the pattern of a state split over a cluster: the next operation needs the partner's data).
// Variants: 0 = st.shared::cluster (16 B per thread and store) + barrier.cluster arrive/wait
//           1 = st.async ... mbarrier::complete_tx (16 B per store), consumer waits on its own mbarrier
//           2 = local STS + one cp.async.bulk shared::cta -> shared::cluster per warp + mbarrier complete_tx
//           3 = local STS + __syncthreads only (same bytes inside one CTA: the baseline an un-split CTA pays)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <int VARIANT, int per_thread>
__global__ void __launch_bounds__(256) bench(int iters, double* sink, long long* cycles) {
    extern __shared__ __align__(128) unsigned char raw[];
    // two receive buffers (alternating), one send staging buffer, two mbarriers
    const int tid = threadIdx.x;
    const int S = per_thread * 256 * 16;
    double2* recv0 = reinterpret_cast<double2*>(raw);
    double2* recv1 = reinterpret_cast<double2*>(raw + S);
    double2* stage = reinterpret_cast<double2*>(raw + 2 * S);
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 3 * S);
    const uint32_t me = cluster_rank(), peer = me ^ 1;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (VARIANT != 3) { cluster_arrive(); cluster_wait(); }
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = make_double2(tid + e, me);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double2* recv = (it & 1) ? recv1 : recv0;
        const uint32_t bar = smem_u32(&bars[it & 1]);
        const uint32_t parity = (it >> 1) & 1;
        if (VARIANT == 0) {
            const uint32_t dst = mapa(smem_u32(recv), peer);
            _Pragma("unroll") for (int e = 0; e < per_thread; ++e)
                asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(dst + (e * 256 + tid) * 16), "d"(v[e & 7].x), "d"(v[e & 7].y) : "memory");
            cluster_arrive();
            cluster_wait();
        } else if (VARIANT == 1) {
            if (tid == 0) mbar_expect_tx(bar, S);
            const uint32_t dst = mapa(smem_u32(recv), peer);
            const uint32_t rbar = mapa(bar, peer);
            _Pragma("unroll") for (int e = 0; e < per_thread; ++e)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(dst + (e * 256 + tid) * 16), "d"(v[e & 7].x), "d"(v[e & 7].y), "r"(rbar) : "memory");
            mbar_wait(bar, parity);
        } else if (VARIANT == 2) {
            if (tid == 0) mbar_expect_tx(bar, S);
            // warp w stages its slice [e][w*32 + lane] contiguously per warp: per_thread * 512 bytes per warp
            const int w = tid >> 5, lane = tid & 31;
            double2* mine = stage + (size_t)w * per_thread * 32;
            _Pragma("unroll") for (int e = 0; e < per_thread; ++e) mine[e * 32 + lane] = v[e & 7];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t dst = mapa(smem_u32(recv + (size_t)w * per_thread * 32), peer);
                const uint32_t rbar = mapa(bar, peer);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(smem_u32(mine)), "r"(per_thread * 512), "r"(rbar) : "memory");
            }
            mbar_wait(bar, parity);
        } else {
            _Pragma("unroll") for (int e = 0; e < per_thread; ++e) recv[e * 256 + tid] = v[e & 7];
            __syncthreads();
        }
        // consume: partner index differs (as a real exchange would read)
        _Pragma("unroll") for (int e = 0; e < per_thread; ++e) {
            const double2 p = recv[e * 256 + (tid ^ 128)];
            v[e & 7].x = fma(0.5, p.y, v[e & 7].x);
            v[e & 7].y = fma(-0.5, p.x, v[e & 7].y);
        }
        if (VARIANT == 2) __syncthreads();        // the staging buffer is rewritten next iteration while the bulk copy may still read it: it completed (peer's mbarrier) -- conservative
    }
    long long t1 = clock64();
    if (VARIANT != 3) { cluster_arrive(); cluster_wait(); }
    double acc = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += v[e].x + v[e].y;
    sink[blockIdx.x * 256 + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VARIANT, int per_thread>
void run(int cluster, int iters, double* sink, long long* cycles) {
    const int S = per_thread * 256 * 16;
    const size_t smem = 3 * (size_t)S + 64;
    cudaFuncSetAttribute(bench<VARIANT, per_thread>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = VARIANT == 3 ? 1 : cluster;
    attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench<VARIANT, per_thread>, iters, sink, cycles);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[8];
    cudaMemcpy(h, cycles, sizeof(long long) * cluster, cudaMemcpyDeviceToHost);
    printf("variant %d cluster %d bytes/CTA/exchange %6d : %8.1f cycles per exchange (%.1f B/cycle/CTA)\n", VARIANT, cluster, S,
           (double)h[0] / iters, (double)S * iters / h[0]);
}

int main() {
    double* sink;
    long long* cycles;
    cudaMalloc(&sink, 8 * 256 * sizeof(double));
    cudaMalloc(&cycles, 8 * sizeof(long long));
    const int iters = 2000;
#define SWEEP(PT) run<3, PT>(1, iters, sink, cycles); run<0, PT>(2, iters, sink, cycles); run<1, PT>(2, iters, sink, cycles); \
    run<2, PT>(2, iters, sink, cycles); run<0, PT>(4, iters, sink, cycles); run<1, PT>(4, iters, sink, cycles);
    SWEEP(1) SWEEP(2) SWEEP(4) SWEEP(6) SWEEP(8) SWEEP(12)
    return 0;
}
