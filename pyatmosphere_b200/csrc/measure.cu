// Fused reductions over the output field (replaces /root/reference/pyatmosphere/measures.py:1-38,
// pupils.py:8-13 and the closed form of simulations/beam.py:26-33), intensity, pupil masking, the device-side
// PDT histogram (simulations/pdt.py:30-31) and the phase -> turns conversion used when a caller supplies a
// ready-made screen.
#include "common.cuh"
#include "internal_measure.h"

namespace pa {

// ---- fused moments + aperture transmittances --------------------------------------------------------------
// One warp per row at a time; lanes stride over x.  Row sums are formed in float32 per lane (64 elements at
// N = 2048) and folded into float64 accumulators per row, so the result is independent of the launch shape up
// to float64 rounding.  Raw sums per CTA go to `partials`, k_measure_finish reduces them in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256) k_measure_partial(MeasureLaunch a) {
    using C = cplx<T>;
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const C* field = (const C*)a.field + (size_t)b * a.n * a.n;
    double acc[kRawMoments + kMaxPupils];
#pragma unroll
    for (int q = 0; q < kRawMoments + kMaxPupils; ++q) acc[q] = 0.0;
    float pr2[kMaxPupils], psx[kMaxPupils], psy[kMaxPupils];
#pragma unroll
    for (int p = 0; p < kMaxPupils; ++p) {
        const bool on = p < a.npupil;
        const float* pp = a.pupils + ((size_t)(a.pupils_per_field ? b : 0) * a.npupil + (on ? p : 0)) * 3;
        pr2[p] = on ? pp[0] : -1.0f;      // radius^2 rounded to float32 by the host, as numpy's weak scalar is
        psx[p] = on ? pp[1] : 0.0f;
        psy[p] = on ? pp[2] : 0.0f;
    }
    for (int row = blockIdx.x * warps + warp; row < a.n; row += gridDim.x * warps) {
        const C* r = field + (size_t)row * a.n;
        const float yv = a.y[row];
        T s0 = 0, s1 = 0, s2 = 0;     // row sums in the field's own precision
        T sp[kMaxPupils];
        float dy2[kMaxPupils];
#pragma unroll
        for (int p = 0; p < kMaxPupils; ++p) {
            sp[p] = 0;
            const float dy = __fadd_rn(yv, psy[p]);          // (y + shift_y), pupils.py:10
            dy2[p] = __fmul_rn(dy, dy);
        }
        for (int j = lane; j < a.n; j += 32) {
            const C u = r[j];
            const float xv = a.x[j];
            const T in = u.x * u.x + u.y * u.y;
            s0 += in;
            s1 += in * (T)xv;
            s2 += in * ((T)xv * (T)xv);
#pragma unroll
            for (int p = 0; p < kMaxPupils; ++p) {
                const float dx = __fsub_rn(xv, psx[p]);
                const bool inside = __fadd_rn(__fmul_rn(dx, dx), dy2[p]) <= pr2[p];
                sp[p] += inside ? in : (T)0;
            }
        }
        const double y = (double)yv;
        acc[0] += (double)s0;
        acc[1] += (double)s1;
        acc[2] += y * (double)s0;
        acc[3] += (double)s2;
        acc[4] += y * (double)s1;
        acc[5] += y * y * (double)s0;
#pragma unroll
        for (int p = 0; p < kMaxPupils; ++p) acc[kRawMoments + p] += (double)sp[p];
    }
    __shared__ double red[8][kRawMoments + kMaxPupils];
#pragma unroll
    for (int q = 0; q < kRawMoments + kMaxPupils; ++q) {
        double v = acc[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < kRawMoments + kMaxPupils) {
        double v = 0.0;
        for (int w = 0; w < warps; ++w) v += red[w][threadIdx.x];
        a.partials[((size_t)b * gridDim.x + blockIdx.x) * (kRawMoments + kMaxPupils) + threadIdx.x] = v;
    }
}

// out[b] = {eta, mean_x, mean_y, mean_x2, mean_xy, mean_y2, mean_x2_r, 0, eta_pupil[0..npupil)}
__global__ void k_measure_finish(MeasureLaunch a, int nparts) {
    const int b = blockIdx.x;
    __shared__ double tot[kRawMoments + kMaxPupils];
    if (threadIdx.x < kRawMoments + kMaxPupils) {
        double v = 0.0;
        for (int i = 0; i < nparts; ++i) v += a.partials[((size_t)b * nparts + i) * (kRawMoments + kMaxPupils) + threadIdx.x];
        tot[threadIdx.x] = v * a.delta2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* o = a.out + (size_t)b * a.out_stride;
        const double eta = tot[0], mx = tot[1], my = -tot[2], mx2 = tot[3], mxy = -tot[4], my2 = tot[5];
        o[0] = eta; o[1] = mx; o[2] = my; o[3] = mx2; o[4] = mxy; o[5] = my2;
        const double r0 = sqrt(mx * mx + my * my);
        const double c = mx / r0, s = my / r0;     // r0 == 0 gives NaN exactly like the reference's 0/0
        o[6] = c * c * mx2 + 2.0 * c * s * mxy + s * s * my2;
        o[7] = 0.0;
        for (int p = 0; p < a.npupil; ++p) o[kMeasureHead + p] = tot[kRawMoments + p];
    }
}

int launch_measure(int prec, const MeasureLaunch& a, cudaStream_t st) {
    const int nparts = a.nparts;
    dim3 g(nparts, a.batch);
    if (prec == 0) k_measure_partial<float><<<g, 256, 0, st>>>(a);
    else k_measure_partial<double><<<g, 256, 0, st>>>(a);
    k_measure_finish<<<a.batch, 32, 0, st>>>(a, nparts);
    return (int)cudaGetLastError();
}

// out[b] from the per-row sums of the fused final row pass: rowsums[b*n + i] = {S0, S1, S2, P0..P3}(row i)
__global__ void __launch_bounds__(256) k_measure_finish_rows(const double* rowsums, const float* y, int n, double delta2, int npupil,
                                                             double* out, int out_stride) {
    const int b = blockIdx.x;
    double acc[kRawMoments + kFusedPupils];
#pragma unroll
    for (int q = 0; q < kRawMoments + kFusedPupils; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* r = rowsums + ((size_t)b * n + i) * kRowSums;
        const double yv = (double)y[i];
        acc[0] += r[0];
        acc[1] += r[1];
        acc[2] += yv * r[0];
        acc[3] += r[2];
        acc[4] += yv * r[1];
        acc[5] += yv * yv * r[0];
#pragma unroll
        for (int p = 0; p < kFusedPupils; ++p) acc[kRawMoments + p] += r[3 + p];
    }
    __shared__ double red[8][kRawMoments + kFusedPupils];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < kRawMoments + kFusedPupils; ++q) {
        double v = acc[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[kRawMoments + kFusedPupils];
        for (int q = 0; q < kRawMoments + kFusedPupils; ++q) {
            double v = 0.0;
            for (int w = 0; w < 8; ++w) v += red[w][q];
            tot[q] = v * delta2;
        }
        double* o = out + (size_t)b * out_stride;
        const double eta = tot[0], mx = tot[1], my = -tot[2], mx2 = tot[3], mxy = -tot[4], my2 = tot[5];
        o[0] = eta; o[1] = mx; o[2] = my; o[3] = mx2; o[4] = mxy; o[5] = my2;
        const double r0 = sqrt(mx * mx + my * my);
        const double c = mx / r0, s = my / r0;
        o[6] = c * c * mx2 + 2.0 * c * s * mxy + s * s * my2;
        o[7] = 0.0;
        for (int p = 0; p < npupil; ++p) o[kMeasureHead + p] = tot[kRawMoments + p];
    }
}

int launch_measure_rows(const double* rowsums, const float* y, int n, int batch, double delta2, int npupil, double* out, int out_stride,
                        cudaStream_t st) {
    k_measure_finish_rows<<<batch, 256, 0, st>>>(rowsums, y, n, delta2, npupil, out, out_stride);
    return (int)cudaGetLastError();
}

// ---- element-wise helpers ----------------------------------------------------------------------------------
template <typename T> __global__ void k_intensity(const cplx<T>* u, T* out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const cplx<T> v = u[i];
        out[i] = v.x * v.x + v.y * v.y;
    }
}
int launch_intensity(int prec, const void* u, void* out, size_t count, cudaStream_t st) {
    const int blocks = (int)((count + 255) / 256 < 148 * 16 ? (count + 255) / 256 : 148 * 16);
    if (prec == 0) k_intensity<float><<<blocks, 256, 0, st>>>((const float2*)u, (float*)out, count);
    else k_intensity<double><<<blocks, 256, 0, st>>>((const double2*)u, (double*)out, count);
    return (int)cudaGetLastError();
}

// out = in or conj(in)   (last step of pa_fft2c)
template <typename T> __global__ void k_copy_conj(const cplx<T>* in, cplx<T>* out, size_t count, bool conj) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        cplx<T> v = in[i];
        if (conj) v.y = -v.y;
        out[i] = v;
    }
}
int launch_copy_conj(int prec, const void* in, void* out, size_t count, bool conj, cudaStream_t st) {
    const int blocks = (int)((count + 255) / 256 < 148 * 16 ? (count + 255) / 256 : 148 * 16);
    if (prec == 0) k_copy_conj<float><<<blocks, 256, 0, st>>>((const float2*)in, (float2*)out, count, conj);
    else k_copy_conj<double><<<blocks, 256, 0, st>>>((const double2*)in, (double2*)out, count, conj);
    return (int)cudaGetLastError();
}

// theory/sources.py:16-18  amp * exp(-(aw + i ac) r2), evaluated in float64 and rounded once
template <typename T> __global__ void k_gaussian_amplitude(const T* r2, cplx<T>* out, size_t count, double amp, double aw, double ac) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const double q = (double)r2[i];
        const double mag = amp * exp(-aw * q);
        double sn, cs;
        sincos(ac * q, &sn, &cs);
        out[i] = mkc<T>((T)(mag * cs), (T)(-mag * sn));
    }
}
int launch_gaussian_amplitude(int prec, const void* r2, void* out, size_t count, double amp, double aw, double ac, cudaStream_t st) {
    const int blocks = (int)((count + 255) / 256 < 148 * 16 ? (count + 255) / 256 : 148 * 16);
    if (prec == 0) k_gaussian_amplitude<float><<<blocks, 256, 0, st>>>((const float*)r2, (float2*)out, count, amp, aw, ac);
    else k_gaussian_amplitude<double><<<blocks, 256, 0, st>>>((const double*)r2, (double2*)out, count, amp, aw, ac);
    return (int)cudaGetLastError();
}

// out = in * [(x - sx)^2 + (y + sy)^2 <= r^2]   (pupils.py:8-13), float32 compare like the reference
template <typename T>
__global__ void k_pupil(const cplx<T>* in, cplx<T>* out, const float* x, const float* y, int n, int batch, float r2, float sx, float sy) {
    const size_t count = (size_t)batch * n * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % n), row = (int)((i / n) % n);
        const float dx = __fsub_rn(x[j], sx), dy = __fadd_rn(y[row], sy);
        const bool inside = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= r2;
        cplx<T> v = in[i];
        if (!inside) { v.x = 0; v.y = 0; }
        out[i] = v;
    }
}
int launch_pupil(int prec, const void* in, void* out, const float* x, const float* y, int n, int batch, float r2, float sx, float sy, cudaStream_t st) {
    const int blocks = 148 * 8;
    if (prec == 0) k_pupil<float><<<blocks, 256, 0, st>>>((const float2*)in, (float2*)out, x, y, n, batch, r2, sx, sy);
    else k_pupil<double><<<blocks, 256, 0, st>>>((const double2*)in, (double2*)out, x, y, n, batch, r2, sx, sy);
    return (int)cudaGetLastError();
}

// phase (radians, float32 or float64) -> turns in [-0.5, 0.5], reduced in float64
template <typename TIN, typename TOUT> __global__ void k_phase_to_turns(const TIN* phi, TOUT* turns, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const double t = (double)phi[i] * 0.15915494309189533576888376;
        turns[i] = (TOUT)(t - rint(t));
    }
}
int launch_phase_to_turns(const void* phi, int phi_f64, void* turns, int turns_f64, size_t count, cudaStream_t st) {
    const int blocks = 148 * 8;
    if (!phi_f64 && !turns_f64) k_phase_to_turns<float, float><<<blocks, 256, 0, st>>>((const float*)phi, (float*)turns, count);
    else if (!phi_f64 && turns_f64) k_phase_to_turns<float, double><<<blocks, 256, 0, st>>>((const float*)phi, (double*)turns, count);
    else if (phi_f64 && !turns_f64) k_phase_to_turns<double, float><<<blocks, 256, 0, st>>>((const double*)phi, (float*)turns, count);
    else k_phase_to_turns<double, double><<<blocks, 256, 0, st>>>((const double*)phi, (double*)turns, count);
    return (int)cudaGetLastError();
}

// ---- histogram with numpy.histogram semantics --------------------------------------------------------------
// edges: nbins+1 doubles (np.linspace on the host).  Values outside [edges[0], edges[nbins]] are dropped, the
// last bin is closed on the right, and the float index is corrected against the actual edges exactly as
// numpy/lib/_histograms_impl.py does for uniform bins.
__global__ void k_histogram(const double* values, size_t stride, size_t count, const double* edges, int nbins, unsigned long long* counts) {
    const double first = edges[0], last = edges[nbins];
    const double norm = (double)nbins / (last - first);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const double v = values[i * stride];
        if (!(v >= first && v <= last)) continue;
        int idx = (int)((v - first) * norm);
        if (idx == nbins) idx -= 1;
        if (v < edges[idx]) idx -= 1;
        else if (v >= edges[idx + 1] && idx != nbins - 1) idx += 1;
        atomicAdd(counts + idx, 1ull);
    }
}
int launch_histogram(const double* values, size_t stride, size_t count, const double* edges, int nbins, unsigned long long* counts, cudaStream_t st) {
    int blocks = (int)((count + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    k_histogram<<<blocks, 256, 0, st>>>(values, stride, count, edges, nbins, counts);
    return (int)cudaGetLastError();
}

}  // namespace pa
