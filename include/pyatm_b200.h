/* pyatm_b200.h -- C ABI of libpyatm_b200.so: hand-written sm_100a kernels for the Monte-Carlo split-step
 * beam-propagation hot path of KlenM/pyAtmosphere.
 *
 * This is the drop-in boundary: the reference reaches its array backend through
 * pyatmosphere/gpu.py:9-20 (`get_xp()` -> cupy, `get_array()`); every function below replaces the cupy calls
 * one reference function makes (file:line given per entry, relative to /root/reference/pyatmosphere/).
 * The Python host (pyatmosphere_b200/_native.py) binds exactly these symbols with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; pa_last_error() gives the text (thread local).
 *  - no exceptions, no Python or torch types: plain pointers, sizes and doubles only.
 *  - pointers named *_dev are device pointers owned by the caller (torch tensors on the Python side),
 *    pointers named *_host are host pointers.  The library never frees caller memory.
 *  - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it unless stated otherwise.
 *  - fields are [batch][N][N] row-major (row index <-> y, column index <-> x, grids.py:57-69), complex64
 *    (precision PA_C64) or complex128 (PA_C128), interleaved re/im.
 *  - a pa_ctx belongs to one device and one thread at a time (the reference is single-threaded too).
 */
#ifndef PYATM_B200_H
#define PYATM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_VERSION 100
#define PA_C64 0
#define PA_C128 1

/* screen synthesis methods */
#define PA_SCREEN_EXACT 0   /* float64 CUDA-core contraction (+ float64 polynomial for the low rings) */
#define PA_SCREEN_TC 1      /* tcgen05 split-fp16 tensor-core contraction (+ float64 polynomial) */

#define PA_MEASURE_HEAD 8   /* eta, mean_x, mean_y, mean_x2, mean_xy, mean_y2, mean_x2_r, 0 */
#define PA_MAX_PUPILS 8     /* apertures evaluated per pa_measure call */

#if defined(__GNUC__)
#define PA_API __attribute__((visibility("default")))
#else
#define PA_API
#endif

typedef struct pa_ctx pa_ctx;

PA_API int pa_version(void);
PA_API const char* pa_last_error(void);
PA_API int pa_device_count(int* count);

/* ---- context ------------------------------------------------------------------------------------------
 * replaces gpu.py:9-14 (backend selection) and the per-call rebuilding of grids (grids.py:63-85): twiddle
 * tables, permuted transfer-function tables and workspace live here. n: power of two in [64, 8192]. */
PA_API int pa_ctx_create(pa_ctx** ctx, int device, int n, int precision);
PA_API int pa_ctx_destroy(pa_ctx* ctx);
/* grids.py:63-72 RectGrid.get_x/get_y: the float32 axes exactly as numpy builds them, uploaded once. */
PA_API int pa_ctx_set_axes(pa_ctx* ctx, const float* x_host, const float* y_host, double delta);
/* frequency (numpy fft order index q in [0,N)) held at storage position p of a permuted spectrum */
PA_API int pa_ctx_permutation(pa_ctx* ctx, int* perm_host);
/* launch geometry of the two FFT passes: {rows threads, rows per CTA, rows smem, cols threads, cols per CTA, cols smem} */
PA_API int pa_ctx_fft_geometry(pa_ctx* ctx, int* six_ints_host);

/* ---- building blocks (one per reference function) --------------------------------------------------------*/
/* sources.py:21-23 GaussianSource.output -> theory/sources.py:16-18 GaussianBeam.amplitude */
PA_API int pa_source_gaussian(pa_ctx* ctx, void* field_dev, int batch, double w0, double wvl, double F0, void* stream);

/* pathes.py:27-40 VacuumPath.lossless_output -> theory/vacuum.py:5-7 -> utils.py:42-50 (fft2/ifft2).
 * In place, natural order in and out; length <= 0 leaves the field untouched. */
PA_API int pa_vacuum_leg(pa_ctx* ctx, void* field_dev, int batch, double length, double wvl, void* stream);

/* utils.py:42-44 fft2 (forward != 0) and utils.py:47-50 ifft2 (forward == 0): the centred two-dimensional transform pair,
 * index N/2 = origin / zero frequency on both sides, natural order in and out.
 *   forward:  out[p][q] = scale * sum_ij in[i][j] exp(-2 pi i ((i-N/2)(p-N/2) + (j-N/2)(q-N/2)) / N)      (fft2: scale = delta^2)
 *   inverse:  the same with exp(+...)                                                          (ifft2: scale = delta_f^2)
 * in_dev, out_dev: [batch][N][N] complex (context precision); in_dev == out_dev is allowed. */
PA_API int pa_fft2c(pa_ctx* ctx, const void* in_dev, void* out_dev, int batch, int forward, double scale, void* stream);

/* theory/sources.py:16-18 GaussianBeam.amplitude on caller-supplied squared radii:
 * out[i] = sqrt(2/pi)/w0 * exp(-(1/w0^2 + i k/(2 F0)) r2[i]);  r2_dev: float32 (PA_C64 ctx) or float64 (PA_C128). */
PA_API int pa_gaussian_amplitude(pa_ctx* ctx, const void* r2_dev, void* out_dev, size_t count, double w0, double wvl, double F0,
                          void* stream);

/* phase_screens.py:108-136 SSPhaseScreen.generate_phase_screen (real part, :25-28).
 * fx, fy: [nscreens][m] float32; coef: [nscreens][m] complex64 (interleaved); rings sorted by radius.
 * m_split / degree: rings below m_split are summed as a float64 polynomial of that total degree
 * (m_split = 0, degree = -1: every ring goes through the contraction).
 * turns_dev: [nscreens][N][N] phase/2pi reduced to [-0.5,0.5], float32 (PA_C64 ctx) or float64 (PA_C128);
 * phi_dev (optional): full phase in radians, float32 or float64 (phi_f64).
 * coef_bound: upper bound of |coef_m| over the rings m >= m_split (fp16 range control of the tensor-core
 * method; ignored by PA_SCREEN_EXACT; <= 0 means "at most 32"). */
PA_API int pa_screen_ss(pa_ctx* ctx, const float* fx_dev, const float* fy_dev, const float* coef_dev, int m, int m_split,
                 int degree, double shift_x, double shift_y, int nscreens, void* turns_dev, void* phi_dev,
                 int phi_f64, int method, double coef_bound, void* stream);

/* phase_screens.py:37-67 FFTPhaseScreen.generate_phase_screen:  ifft2(cn, 1) [utils.py:47-50] + subharmonic terms
 * [:55-65] - mean [:67].
 * spectrum_dev: [nscreens][N][N] complex (context precision), the coefficients cn of :43-47 in the reference's
 * centred frequency order (index N/2 = zero frequency, row <-> first index of cn).
 * terms_host: [nscreens][nterms][4] doubles {fx, fy, Re c, Im c}: the screen gains  c * exp(2 pi i (fx x_j + fy y_i));
 * nterms may be 0.  Outputs (either may be NULL): the complex screen and/or its real part, [nscreens][N][N], context
 * precision.  Not in place: spectrum_dev is only read. */
PA_API int pa_screen_fft(pa_ctx* ctx, const void* spectrum_dev, int nscreens, const double* terms_host, int nterms,
                  void* out_complex_dev, void* out_real_dev, void* stream);

/* pathes.py:72-73  u <- scale * exp(-i phi) * u  with phi given in turns (see pa_phase_to_turns) */
PA_API int pa_apply_screen(pa_ctx* ctx, void* field_dev, int batch, const void* turns_dev, double scale, void* stream);
PA_API int pa_phase_to_turns(pa_ctx* ctx, const void* phi_dev, int phi_f64, void* turns_dev, size_t count, void* stream);

/* measures.py:1-4 I = |u|^2 (float32 / float64 out) */
PA_API int pa_intensity(pa_ctx* ctx, const void* field_dev, void* out_dev, int batch, void* stream);
/* pupils.py:8-13 CirclePupil.output: out = in * [(x-sx)^2 + (y+sy)^2 <= r^2] */
PA_API int pa_pupil_apply(pa_ctx* ctx, const void* in_dev, void* out_dev, int batch, double radius, double shift_x,
                   double shift_y, void* stream);
/* measures.py:7-38 (eta, mean_x, mean_y, mean_x2, mean_xy, mean_y2), simulations/beam.py:26-33 (mean_x2_r)
 * and simulations/pdt.py:21-26 + measures.py:7-8 (eta behind each aperture) in one sweep.
 * pupils_dev: [npupil][3] (or [batch][npupil][3] when pupils_per_field) float32 {radius^2, shift_x, shift_y};
 * out_dev: [batch][out_stride] float64, PA_MEASURE_HEAD values then one eta per aperture. */
PA_API int pa_measure(pa_ctx* ctx, const void* field_dev, int batch, const float* pupils_dev, int npupil,
               int pupils_per_field, double* out_dev, int out_stride, void* stream);
/* simulations/pdt.py:30-31: numpy.histogram(values, bins, range) semantics; edges_dev has nbins+1 doubles */
PA_API int pa_histogram(pa_ctx* ctx, const double* values_dev, size_t stride, size_t count, const double* edges_dev,
                 int nbins, unsigned long long* counts_dev, void* stream);

/* grids.py:98-107 + phase_screens.py:98-103 drawn on the device with Philox4x32-10 keyed by
 * (seed; realization, screen, ring): production mode, independent of the number of GPUs. */
PA_API int pa_rng_spectrum(pa_ctx* ctx, unsigned long long seed, unsigned long long realization0, int batch,
                    int screen0, int nscreens, int m, const float* ring_edges_dev, const float* ring_psd_dev,
                    float* fx_dev, float* fy_dev, float* coef_dev, void* stream);

/* ---- fused path -------------------------------------------------------------------------------------------
 * pathes.py:61-75 PhaseScreensPath.generator drained by lossless_output (:53-59), started from
 * GaussianSource.output and followed (through_output) by AbstractPath.output (:23-24). */
typedef struct pa_path {
    int n_screens;
    const double* leg_lengths_host;   /* n_screens + 1 entries: leg in front of each screen, then the closing leg */
    const double* screen_scale_host;  /* n_screens entries: amplitude factor applied together with screen i */
    double final_scale;               /* amplitude factor after the closing leg */
    double wvl, w0, F0;               /* source */
    int m, m_split, degree;           /* screens, see pa_screen_ss */
    double shift_x, shift_y;
    int screen_method;
    double coef_bound;                /* see pa_screen_ss */
    int from_field;                   /* 0: start from the Gaussian source (generated inside the first pass);
                                         1: field_dev already holds the input field in natural order */
} pa_path;

/* pathes.py:61-75 (generator) drained by :53-59 (lossless_output), see pa_path above.
 * fx/fy/coef: [n_screens][batch][m] (one contiguous slab per path position); field_dev: [batch][N][N] result,
 * natural order */
PA_API int pa_propagate(pa_ctx* ctx, const pa_path* path, void* field_dev, int batch, const float* fx_dev,
                 const float* fy_dev, const float* coef_dev, void* stream);

/* One Monte-Carlo batch end to end (simulations/simulation.py:89-114 for BeamResult + PDTResult measures):
 * coefficients come either from host buffers (coef_host != NULL: fx_host, fy_host, coef_host
 * [n_screens][batch][m], copied host->device inside the call) or from the device RNG (coef_host == NULL, seed /
 * realization0 / ring tables used).  The per-realization table [batch][out_stride] is copied to out_host and
 * the call returns after the stream has drained.  The field buffer is ctx workspace.  `batch` may be any size: the
 * library works through it in chunks (32 realizations at 2048^2) with one host->device copy of the coefficients up front
 * and one synchronisation at the end. */
PA_API int pa_simulate_batch(pa_ctx* ctx, const pa_path* path, int batch, const float* fx_host, const float* fy_host,
                      const float* coef_host, unsigned long long seed, unsigned long long realization0,
                      const float* ring_edges_dev, const float* ring_psd_dev, const float* pupils_host,
                      int npupil, double* out_host, int out_stride, void* stream);
/* The same batch, enqueued only: the call returns as soon as the copies and kernels are on `stream` (simulations/
 * simulation.py:89-114, as pa_simulate_batch).  The host buffers (pinned memory for the copies to be asynchronous) must stay
 * untouched, and out_host unread, until pa_stream_synchronize (or any synchronisation of `stream`) returns; a host that
 * enqueues batch i+1 before it waits for batch i keeps the GPU busy across the call boundary. */
PA_API int pa_simulate_batch_async(pa_ctx* ctx, const pa_path* path, int batch, const float* fx_host, const float* fy_host,
                            const float* coef_host, unsigned long long seed, unsigned long long realization0,
                            const float* ring_edges_dev, const float* ring_psd_dev, const float* pupils_host,
                            int npupil, double* out_host, int out_stride, void* stream);
PA_API int pa_stream_synchronize(pa_ctx* ctx, void* stream);
/* same, everything stays on the device and nothing synchronises (throughput loop); table_dev [batch][out_stride] */
PA_API int pa_simulate_batch_device(pa_ctx* ctx, const pa_path* path, int batch, unsigned long long seed,
                             unsigned long long realization0, const float* ring_edges_dev,
                             const float* ring_psd_dev, const float* pupils_dev, int npupil, double* table_dev,
                             int out_stride, void* stream);

/* One pass of the split-step FFT pipeline on a field that is in row-spectrum form, for profiling and for the
 * roofline measurement of bench.py: kind 0 = column pass (FFT_y, x H, IFFT_y) of a leg of `length`;
 * kind 1 = fused row pass (IFFT_x, x exp(-2 pi i turns), FFT_x), turns_dev may be NULL. */
PA_API int pa_fft_pass(pa_ctx* ctx, void* field_dev, int batch, int kind, const void* turns_dev, double length,
                       double wvl, void* stream);

/* ---- the one collective of the path ----------------------------------------------------------------------------
 * Independent realizations are sharded over the GPUs; what the reference computes from one process's
 * samples -- the 200-bin transmittance histograms (simulations/pdt.py:30-31) and the sums behind sigma_BW / sigma_LT / W_ST
 * (simulations/beam.py:35-71) -- is summed over the ranks here, on the device, over NCCL (NVLink / NVSwitch).  NCCL is bound
 * with dlopen("libnccl.so.2") at the first call (override: PYATM_NCCL_LIB); hosts that never create a pa_comm never need it.
 * The 128-byte id is created on one rank and handed to the others by the host (any side channel: file, MPI, sockets,
 * torch.distributed's object broadcast); id_out / id point to PA_COMM_ID_BYTES bytes. */
#define PA_COMM_ID_BYTES 128
typedef struct pa_comm pa_comm;
PA_API int pa_comm_unique_id(unsigned char* id_out);
PA_API int pa_comm_create(pa_comm** comm, int device, int rank, int world, const unsigned char* id);
PA_API int pa_comm_destroy(pa_comm* comm);
/* Sum over the ranks of what simulations/pdt.py:30-31 (histogram counts) and simulations/beam.py:35-71 (n, sum v, sum v^2)
 * accumulate per process.  In place, both buffers on the device of the communicator: hist_dev [nbins] uint64 counts
 * (pa_histogram), sums_dev [nsums] float64; either count may be 0.  Asynchronous on `stream`. */
PA_API int pa_stats_allreduce(pa_comm* comm, unsigned long long* hist_dev, size_t nbins, double* sums_dev, size_t nsums,
                              void* stream);

/* number of kernel launches issued by this library since the counter was last reset (bench.py gpu_launches) */
PA_API unsigned long long pa_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* PYATM_B200_H */
