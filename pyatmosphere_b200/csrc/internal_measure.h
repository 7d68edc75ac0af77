// Internal launch descriptors of the reduction / element-wise kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pa {

constexpr int kRawMoments = 6;    // sum I, I x, I y, I x^2, I x y, I y^2
constexpr int kMaxPupils = 8;     // apertures per launch
constexpr int kMeasureHead = 8;   // eta, mean_x, mean_y, mean_x2, mean_xy, mean_y2, mean_x2_r, (pad)
constexpr int kFusedPupils = 4;   // apertures the final row pass can reduce on the fly
constexpr int kRowSums = 3 + kFusedPupils;   // per row: sum I, sum I x, sum I x^2, sum I [inside aperture p]

struct MeasureLaunch {
    const void* field;      // [batch][n][n] complex
    int n, batch;
    const float* x;
    const float* y;
    double delta2;
    const float* pupils;    // [npupil][3] or [batch][npupil][3]: radius^2 (float32), shift_x, shift_y
    int npupil;
    int pupils_per_field;   // 1: every field has its own pupil table (tracked apertures)
    double* partials;       // workspace [batch][nparts][kRawMoments + kMaxPupils]
    int nparts;
    double* out;            // [batch][out_stride]
    int out_stride;         // >= kMeasureHead + npupil
};

int launch_measure(int prec, const MeasureLaunch& a, cudaStream_t st);
// second half of the fused path: fold the per-row sums written by the final row pass (k_rows<..., MEAS>)
int launch_measure_rows(const double* rowsums, const float* y, int n, int batch, double delta2, int npupil, double* out, int out_stride,
                        cudaStream_t st);
int launch_intensity(int prec, const void* u, void* out, size_t count, cudaStream_t st);
int launch_copy_conj(int prec, const void* in, void* out, size_t count, bool conj, cudaStream_t st);
int launch_gaussian_amplitude(int prec, const void* r2, void* out, size_t count, double amp, double aw, double ac, cudaStream_t st);
int launch_pupil(int prec, const void* in, void* out, const float* x, const float* y, int n, int batch, float r2, float sx, float sy, cudaStream_t st);
int launch_phase_to_turns(const void* phi, int phi_f64, void* turns, int turns_f64, size_t count, cudaStream_t st);
int launch_histogram(const double* values, size_t stride, size_t count, const double* edges, int nbins, unsigned long long* counts, cudaStream_t st);

}  // namespace pa
