// qsx_gemm.cuh -- read-out product out[m][n] = sum_k A[m][k] B[n][k] (complex128; include/qsx.h qsx_readout_gemm).
//
// Two forms of the 64 x 64 CTA tile with a two-stage shared-memory pipeline (the first form, readout_gemm_kernel in
// qsx_engine.cu, waits for every k-tile: ncu long_scoreboard 2.97):
//   simt::kernel  FP64 CUDA cores, 256 threads x (4 x 4 complex) register tile, k-tiles of 16 brought in by cp.async
//                 (16 B per element, zero-filled outside the operands) while the previous tile is multiplied;
//   dmma::kernel  FP64 tensor cores, mma.sync.m8n8k4.f64: 8 warps x (32 x 16) warp tile = 4 x 2 MMA tiles, four real
//                 MMAs per complex tile (Cr += Ar Br - Ai Bi, Ci += Ar Bi + Ai Br); operands split into real and
//                 imaginary planes in shared memory (row pitch 20 doubles: the 8 x 4 fragment loads of a warp hit 32
//                 different banks), next k-tile prefetched into registers.
// Both keep the split-K / batch conventions of the first form (blockIdx.y = batch * m_tiles + m-tile, blockIdx.z =
// k-range, partial sums at C + (z * n_batch + batch) * M * N).  qsx_readout_gemm_batched picks one (QSX_GEMM=old|simt|dmma
// overrides); profiles/r2_readout_gemm.md holds the A/B.
#pragma once

namespace gemm {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

namespace simt {

constexpr int LD = TM + 1;                                                  // complex elements per k-row of a stage
constexpr size_t SMEM = (size_t)2 * 2 * TK * LD * sizeof(double2);          // two stages x (A, B)

__global__ void __launch_bounds__(256, 2) kernel(int M, int N, int K, int k_chunk, int m_tiles, int n_batch,
                                                 const double2* __restrict__ A, const double2* __restrict__ B,
                                                 double2* __restrict__ C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* As = reinterpret_cast<double2*>(smem_raw);                     // [2][TK][LD]
    double2* Bs = As + 2 * TK * LD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int batch = blockIdx.y / m_tiles;
    const int m0 = (blockIdx.y % m_tiles) * TM, n0 = blockIdx.x * TN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    A += (size_t)batch * M * K;
    B += (size_t)batch * N * K;
    C += ((size_t)blockIdx.z * n_batch + batch) * M * N;
    double2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);

    // element q of a 64 x 16 tile: row q / 16, k = q % 16 (consecutive threads read consecutive k of one row)
    auto load_tile = [&](int stage, int k0) {
        double2* as = As + stage * TK * LD;
        double2* bs = Bs + stage * TK * LD;
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            const int q = tid + rep * 256, r = q >> 4, c = q & 15;
            const bool ka = k0 + c < k_end;
            const bool va = ka && m0 + r < M, vb = ka && n0 + r < N;
            cp_async16(&as[c * LD + r], va ? &A[(size_t)(m0 + r) * K + k0 + c] : A, va);
            cp_async16(&bs[c * LD + r], vb ? &B[(size_t)(n0 + r) * K + k0 + c] : B, vb);
        }
        cp_async_commit();
    };

    const int n_tiles = (k_end - k_begin + TK - 1) / TK;
    if (n_tiles > 0) load_tile(0, k_begin);
    for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) {
            load_tile((t + 1) & 1, k_begin + (t + 1) * TK);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double2* as = As + (t & 1) * TK * LD;
        const double2* bs = Bs + (t & 1) * TK * LD;
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double2 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = as[kk * LD + ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[kk * LD + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j].x = fma(a[i].x, b[j].x, fma(-a[i].y, b[j].y, acc[i][j].x));
                    acc[i][j].y = fma(a[i].x, b[j].y, fma(a[i].y, b[j].x, acc[i][j].y));
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < M && n < N) C[(size_t)m * N + n] = acc[i][j];
        }
}

}  // namespace simt

namespace dmma {

constexpr int LD = TK + 4;                                                  // doubles per row of a plane (bank spread)
constexpr int PLANE = TM * LD;
constexpr size_t SMEM = (size_t)2 * 4 * PLANE * sizeof(double);             // two stages x (Ar, Ai, Br, Bi)

__device__ __forceinline__ void mma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) kernel(int M, int N, int K, int k_chunk, int m_tiles, int n_batch,
                                                 const double2* __restrict__ A, const double2* __restrict__ B,
                                                 double2* __restrict__ C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* planes = reinterpret_cast<double*>(smem_raw);                   // [2][4][TM][LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                                // 2 x 4 warps: warp tile 32 (m) x 16 (n)
    const int grp = lane >> 2, tig = lane & 3;
    const int batch = blockIdx.y / m_tiles;
    const int m0 = (blockIdx.y % m_tiles) * TM, n0 = blockIdx.x * TN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    A += (size_t)batch * M * K;
    B += (size_t)batch * N * K;
    C += ((size_t)blockIdx.z * n_batch + batch) * M * N;
    double cr[4][2][2], ci[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

    double2 ra[4], rb[4];                                                   // the next k-tile, on its way to shared memory
    const double2 zero = make_double2(0.0, 0.0);
    auto fetch = [&](int k0) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            const int q = tid + rep * 256, r = q >> 4, c = q & 15;
            const bool ka = k0 + c < k_end;
            ra[rep] = (ka && m0 + r < M) ? A[(size_t)(m0 + r) * K + k0 + c] : zero;
            rb[rep] = (ka && n0 + r < N) ? B[(size_t)(n0 + r) * K + k0 + c] : zero;
        }
    };
    auto stash = [&](int stage) {
        double* p = planes + stage * 4 * PLANE;
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            const int q = tid + rep * 256, r = q >> 4, c = q & 15;
            p[r * LD + c] = ra[rep].x;
            p[PLANE + r * LD + c] = ra[rep].y;
            p[2 * PLANE + r * LD + c] = rb[rep].x;
            p[3 * PLANE + r * LD + c] = rb[rep].y;
        }
    };

    const int n_tiles = (k_end - k_begin + TK - 1) / TK;
    if (n_tiles > 0) {
        fetch(k_begin);
        stash(0);
    }
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) fetch(k_begin + (t + 1) * TK);                 // global loads in flight during the MMAs
        const double* p = planes + (t & 1) * 4 * PLANE;
        const double* ar = p + (wm * 32 + grp) * LD + tig;
        const double* br = p + 2 * PLANE + (wn * 16 + grp) * LD + tig;
#pragma unroll
        for (int k4 = 0; k4 < TK; k4 += 4) {
            double fa_r[4], fa_i[4], fa_n[4], fb_r[2], fb_i[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fa_r[i] = ar[i * 8 * LD + k4];
                fa_i[i] = ar[PLANE + i * 8 * LD + k4];
                fa_n[i] = -fa_i[i];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                fb_r[j] = br[j * 8 * LD + k4];
                fb_i[j] = br[PLANE + j * 8 * LD + k4];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_8x8x4(cr[i][j][0], cr[i][j][1], fa_r[i], fb_r[j]);
                    mma_8x8x4(ci[i][j][0], ci[i][j][1], fa_r[i], fb_i[j]);
                    mma_8x8x4(cr[i][j][0], cr[i][j][1], fa_n[i], fb_i[j]);
                    mma_8x8x4(ci[i][j][0], ci[i][j][1], fa_i[i], fb_r[j]);
                }
        }
        if (t + 1 < n_tiles) stash((t + 1) & 1);                           // the other stage: nobody reads it now
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + wm * 32 + i * 8 + grp;
            const int n = n0 + wn * 16 + j * 8 + tig * 2;
            if (m < M) {
                if (n < N) C[(size_t)m * N + n] = make_double2(cr[i][j][0], ci[i][j][0]);
                if (n + 1 < N) C[(size_t)m * N + n + 1] = make_double2(cr[i][j][1], ci[i][j][1]);
            }
        }
}

}  // namespace dmma

}  // namespace gemm
