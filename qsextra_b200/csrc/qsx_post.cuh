// Post-processing of a simulated response function on the device (reference: qsextra/spectroscopy/postprocessing.py).
//   frame_kernel : rotating frame, damping and zero padding in one pass (postprocessing.py:9-37 and the np.pad calls
//                  of :57 / :77) -- the padded array is what the FFT (cuFFT, called by the host) consumes;
//   shift_kernel : fftshift / ifftshift of both axes and the prefactor (postprocessing.py:61, :84-85) in one pass.
// Both are pure HBM streams of 16-byte elements; rows of the output are contiguous in t3 so accesses coalesce.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace post {

struct FrameParams {
    int n1, n2, n3, t2_index, pad, linear;
    double rf, damping;
    const double* t1;
    const double* t3;
    const double2* signal;
    double2* out;
};

__global__ void frame_kernel(FrameParams p) {
    const long long m3 = p.linear ? 1 : (long long)p.n3 * p.pad;
    const long long total = (long long)p.n1 * p.pad * m3;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / m3), j = (int)(e % m3);
        double2 v = make_double2(0.0, 0.0);
        if (i < p.n1 && (p.linear || j < p.n3)) {
            double2 r;
            double phase, decay;
            if (p.linear) {                                     // R(T) e^{+i w T} e^{-g T}
                r = p.signal[i];
                phase = p.rf * p.t1[i];
                decay = exp(-p.damping * p.t1[i]);
            } else {                                            // R(T1,T2,T3) e^{-i w (T1-T3)} e^{-g (T1+T3)}
                r = p.signal[((long long)i * p.n2 + p.t2_index) * p.n3 + j];
                phase = -(p.rf * (p.t1[i] - p.t3[j]));
                decay = exp(-p.damping * (p.t1[i] + p.t3[j]));
            }
            double s, c;
            sincos(phase, &s, &c);
            v.x = (r.x * c - r.y * s) * decay;
            v.y = (r.x * s + r.y * c) * decay;
        }
        p.out[e] = v;
    }
}

// out[i][j] = scale * in[(i + s1) % m1][(j + s3) % m3]
__global__ void shift_kernel(const double2* __restrict__ in, double2* __restrict__ out, int m1, int m3, int s1, int s3,
                             double scale) {
    const long long total = (long long)m1 * m3;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / m3), j = (int)(e % m3);
        int si = i + s1, sj = j + s3;
        if (si >= m1) si -= m1;
        if (sj >= m3) sj -= m3;
        const double2 v = in[(long long)si * m3 + sj];
        out[e] = make_double2(v.x * scale, v.y * scale);
    }
}

}  // namespace post
