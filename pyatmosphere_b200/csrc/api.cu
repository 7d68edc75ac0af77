// C ABI of libpyatm_b200.so (see include/pyatm_b200.h): context, tables, and the orchestration of the
// split-step passes.  Host code only; kernels live in fft_n*.cu, screen*.cu, measure.cu, rng.cu.
#include "../../include/pyatm_b200.h"

#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <complex>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"
#include "internal.h"
#include "internal_fftscreen.h"
#include "internal_measure.h"
#include "internal_rng.h"
#include "internal_screen.h"

namespace pa {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
static inline void note(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

struct TensorMapEntry {
    const void* field;
    int batch;
    CUtensorMap map;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

struct SepTable {          // Gaussian source carried through the first leg analytically (separable)
    double length, wvl, w0, F0;
    void* dev;             // cplx<T>[n]: p(x) with u1(y,x) = scale * p(y) p(x)
    double scale_re, scale_im;
};

struct HTable {
    double length, wvl;
    void* dev;            // cplx<T>[n] permuted transfer-function factor
    void* dev_y;          // the same factor in the composite order of the split column pass (null without it)
    double alpha_re, alpha_im;   // e^{ikL} / n^2
};

}  // namespace pa

using namespace pa;

struct pa_ctx {
    int device = 0, n = 0, prec = 0, e = 0;
    double delta = 0.0;
    bool axes_set = false;
    std::vector<float> hx, hy;         // host copies of the axes
    float* x = nullptr;                // device axes
    float* y = nullptr;
    void* tw = nullptr;                // twiddles
    std::vector<int> perm;             // frequency index held at storage position p
    // split column pass (fft_split.cuh): outer radix (0 = off), inner-plan twiddles, outer twiddles, and the
    // frequency held at row p of a column spectrum
    int split = 0;
    void* tw_sub = nullptr;
    void* otw = nullptr;
    std::vector<int> perm_y;
    std::vector<HTable> htabs;
    std::vector<SepTable> seps;
    bool sep_first_leg = true;   // complex64 only: start from the analytic first leg (exact identity, float64 1-D transform)
    // workspace (grown on demand, never shrunk)
    void* turns = nullptr; size_t turns_bytes = 0;
    double* P = nullptr; double* Q = nullptr; size_t pq_bytes = 0;
    double* polyc = nullptr; size_t polyc_bytes = 0;
    double* partials = nullptr; size_t partials_bytes = 0;
    void* field = nullptr; size_t field_bytes = 0;
    float* spec = nullptr; size_t spec_bytes = 0;        // fx | fy | coef staging of one chunk of pa_simulate_batch
    float* spec_all = nullptr; size_t spec_all_bytes = 0;   // host coefficients of a whole multi-chunk batch
    float* pupils = nullptr; size_t pupils_bytes = 0;
    double* table = nullptr; size_t table_bytes = 0;
    std::vector<TensorMapEntry> tmaps;   // column-pass tensor maps, keyed by (field pointer, batch)
    bool use_tma = true;        // column pass: TMA-fed persistent kernel
    bool rows_tma = false;      // row pass through the TMA-fed kernel: 8192^2 complex64 and complex128 at every size (fft_tma.cuh: RowSlots)
    double* rowsums = nullptr; size_t rowsums_bytes = 0;   // per-row sums of the fused final pass
    void* tcws = nullptr; size_t tcws_bytes = 0;          // fp16 operand blocks of the tensor-core screen path
    int* tc_err = nullptr;
    cudaStream_t aux = nullptr;                           // second stream: operand generation next to the polynomial kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* perm_dev = nullptr;                              // device copy of `perm` (FFT phase screens)
    void* fftws = nullptr; size_t fftws_bytes = 0;        // subharmonic tables and mean partials of the FFT phase screens
    int num_sms = 148;
    size_t csize() const { return prec == 0 ? 8 : 16; }
    size_t rsize() const { return prec == 0 ? 4 : 8; }
};

static int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return PA_OK;
    if (*p) PA_CUDA(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    PA_CUDA(cudaMalloc(p, need));
    *have = need;
    return PA_OK;
}

static int check_launch(int rc, const char* what) {
    if (rc == 0) return PA_OK;
    if (rc == -1) {
        set_error("%s: unsupported grid size", what);
        return PA_ERR_ARG;
    }
    set_error("%s: %s", what, cudaGetErrorString((cudaError_t)rc));
    return PA_ERR_CUDA;
}

// ---- tables -------------------------------------------------------------------------------------------------
template <typename T> static int build_twiddles(int n, int e, void** out) {
    const int L = plan_len(n, e);
    const int total = plan_tw_size(n, e);
    std::vector<cplx<T>> tw((size_t)(total > 0 ? total : 1));
    for (int s = 0; s + 1 < L; ++s) {
        const int R = plan_radix(n, e, s), sigma = plan_sigma(n, e, s), off = plan_tw_off(n, e, s);
        for (int j = 1; j < R; ++j)
            for (int b = 0; b < sigma; ++b) {
                const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j * (long double)b /
                                        (long double)(sigma * R);
                tw[(size_t)off + (size_t)(j - 1) * sigma + b] = mkc<T>((T)cosl(ang), (T)sinl(ang));
            }
    }
    PA_CUDA(cudaMalloc(out, tw.size() * sizeof(cplx<T>)));
    PA_CUDA(cudaMemcpy(*out, tw.data(), tw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    return PA_OK;
}
// outer-stage twiddles of the split column transform: otw[t * R0 + j] = exp(-2 pi i t j / n), t < n / R0
template <typename T> static int build_outer_twiddles(int n, int r0, void** out) {
    const int m = n / r0;
    std::vector<cplx<T>> tw((size_t)n);
    for (int t = 0; t < m; ++t)
        for (int j = 0; j < r0; ++j) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)((long long)t * j) / (long double)n;
            tw[(size_t)t * r0 + j] = mkc<T>((T)cosl(ang), (T)sinl(ang));
        }
    PA_CUDA(cudaMalloc(out, tw.size() * sizeof(cplx<T>)));
    PA_CUDA(cudaMemcpy(*out, tw.data(), tw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    return PA_OK;
}

// perm[q] = frequency (numpy fft index) stored at index q of a spectrum.  Register idx of thread t holds, after
// the last forward stage, position p = RL * plan_last_butterfly(t, g) + j (idx = g RL + j), whose frequency is the
// mixed-radix digit reversal of p; it is stored at q = t + idx * (n/e)  (fft_core.cuh: io_pos).
static std::vector<int> storage_perm(int n, int e) {
    const int L = plan_len(n, e), tpf = n / e;
    const int RL = plan_radix(n, e, L - 1);
    std::vector<int> perm((size_t)n, -1);
    for (int t = 0; t < tpf; ++t)
        for (int idx = 0; idx < e; ++idx) {
            const int g = idx / RL, j = idx % RL;
            const int p = plan_last_butterfly(n, e, t, g) * RL + j;
            int k = 0, mult = 1;
            for (int s = 0; s < L; ++s) {
                const int R = plan_radix(n, e, s), sigma = plan_sigma(n, e, s);
                k += ((p / sigma) % R) * mult;
                mult *= R;
            }
            perm[t + idx * tpf] = k;
        }
    return perm;
}
// Column spectra: the same order as rows, or - with the split pass - block j of n/R0 rows holds the frequencies
// j + R0 * q, q in the storage order of the (n/R0)-point plan.
static void build_perm(pa_ctx* c) {
    c->perm = storage_perm(c->n, c->e);
    c->perm_y = c->perm;
    if (c->split) {
        const int r0 = c->split, m = c->n / r0;
        const std::vector<int> inner = storage_perm(m, c->e);
        for (int j = 0; j < r0; ++j)
            for (int q = 0; q < m; ++q) c->perm_y[(size_t)j * m + q] = j + r0 * inner[q];
    }
}

// Transfer function of one leg, separable and in permuted order:
//   H[ky][kx] = e^{ikL} h[ky] h[kx],  h[q] = exp(-i (pi L)(2 pi / k) f_q^2),  f_q = fl32(q~) * (1/(N delta))
// evaluated with the reference's operation order (theory/vacuum.py:7, grids.py:63-69,82-85).
static int get_htable(pa_ctx* c, double length, double wvl, const HTable** out) {
    for (const auto& h : c->htabs)
        if (h.length == length && h.wvl == wvl) {
            *out = &h;
            return PA_OK;
        }
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called before propagating");
    const int n = c->n;
    const double k = 2 * M_PI / wvl;
    const double df = 1 / ((double)n * c->delta);
    const double coef = (M_PI * length) * (2 * M_PI / k);
    HTable h;
    h.length = length;
    h.wvl = wvl;
    h.dev = nullptr;
    h.dev_y = nullptr;
    const double ang = k * length;
    const double inv_n2 = 1.0 / ((double)n * (double)n);
    h.alpha_re = cos(ang) * inv_n2;
    h.alpha_im = sin(ang) * inv_n2;
    for (int pass = 0; pass < (c->split ? 2 : 1); ++pass) {
        const std::vector<int>& perm = pass == 0 ? c->perm : c->perm_y;
        void** dst = pass == 0 ? &h.dev : &h.dev_y;
        std::vector<double> re(n), im(n);
        for (int p = 0; p < n; ++p) {
            const int q = perm[p];
            const int qs = q < n / 2 ? q : q - n;
            const double f = (double)(float)qs * df;
            const double ph = -(coef * (f * f));
            re[p] = cos(ph);
            im[p] = sin(ph);
        }
        PA_CUDA(cudaMalloc(dst, (size_t)n * c->csize()));
        if (c->prec == 0) {
            std::vector<float2> t(n);
            for (int p = 0; p < n; ++p) t[p] = make_float2((float)re[p], (float)im[p]);
            PA_CUDA(cudaMemcpy(*dst, t.data(), (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
        } else {
            std::vector<double2> t(n);
            for (int p = 0; p < n; ++p) t[p] = make_double2(re[p], im[p]);
            PA_CUDA(cudaMemcpy(*dst, t.data(), (size_t)n * sizeof(double2), cudaMemcpyHostToDevice));
        }
    }
    c->htabs.push_back(h);
    *out = &c->htabs.back();
    return PA_OK;
}

// 2-D tensor map of a field for the column pass: [batch*n rows][2n reals], box = (2*TC reals) x (256 rows).
static int get_tensor_map(pa_ctx* c, void* field, int batch, const CUtensorMap** out) {
    for (const auto& e : c->tmaps)
        if (e.field == field && e.batch == batch) {
            *out = &e.map;
            return PA_OK;
        }
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        PA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        PA_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        g_encode = (EncodeTiledFn)fn;
    }
    const int n = c->n;
    const size_t csz = c->csize();
    const int tc = c->split ? fft_split_cols_per_tile(c->prec, n) : fft_tma_cols_per_tile(c->prec, n);
    const int rows_per_tile = c->split ? n / c->split : n;
    const int boxr = rows_per_tile < 256 ? rows_per_tile : 256;
    TensorMapEntry e;
    e.field = field;
    e.batch = batch;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * n, (cuuint64_t)batch * n};
    const cuuint64_t strides[1] = {(cuuint64_t)n * csz};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * tc), (cuuint32_t)boxr};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(&e.map, c->prec == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, field, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (c->tmaps.size() >= 64) c->tmaps.erase(c->tmaps.begin());
    c->tmaps.push_back(e);
    *out = &c->tmaps.back().map;
    return PA_OK;
}

// In-place radix-2 FFT in double precision on the host (tables only; n is a power of two).
static void host_fft(std::vector<std::complex<double>>& a, bool inverse) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2 * M_PI / (double)len * (inverse ? 1 : -1);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> w(cos(ang * (double)k), sin(ang * (double)k));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
    if (inverse)
        for (auto& z : a) z /= (double)n;
}

// The Gaussian source is separable, u0(y,x) = amp g(y) g(x) with g(s) = exp(-(1/w0^2 + i k/(2 F0)) s^2), and so is
// the transfer function, hence the field after the first vacuum leg is  amp e^{ikL} p(y) p(x)  with the 1-D
// propagation p = IFFT(h * FFT(g)).  p is computed once per (L, source) in float64 on the host; the first row
// pass then starts from the outer product instead of running a source pass, a column pass and an inverse row
// transform.  (length <= 0: p = g.)  The only deviation from the reference's arithmetic is that rho^2 is not
// rounded to float32 (relative 6e-8 in the exponent); complex128 contexts keep the literal path.
static int get_sep_table(pa_ctx* c, double length, double wvl, double w0, double F0, const SepTable** out) {
    for (const auto& t : c->seps)
        if (t.length == length && t.wvl == wvl && t.w0 == w0 && ((t.F0 == F0) || (isinf(t.F0) && isinf(F0)))) {
            *out = &t;
            return PA_OK;
        }
    const int n = c->n;
    const double k = 2 * M_PI / wvl;
    const double aw = 1 / (w0 * w0), ac = isinf(F0) ? 0.0 : 2 * M_PI / wvl / 2 / F0;
    std::vector<std::complex<double>> g(n);
    for (int i = 0; i < n; ++i) {
        const double s2 = (double)c->hx[i] * (double)c->hx[i];
        g[i] = std::exp(std::complex<double>(-aw * s2, -ac * s2));
    }
    SepTable t;
    t.length = length; t.wvl = wvl; t.w0 = w0; t.F0 = F0; t.dev = nullptr;
    const double amp = sqrt(2 / M_PI) / w0;
    t.scale_re = amp;
    t.scale_im = 0.0;
    if (length > 0) {
        host_fft(g, false);
        const double df = 1 / ((double)n * c->delta);
        const double coef = (M_PI * length) * (2 * M_PI / k);
        for (int q = 0; q < n; ++q) {
            const int qs = q < n / 2 ? q : q - n;
            const double f = (double)(float)qs * df;
            const double ph = -(coef * (f * f));
            g[q] *= std::complex<double>(cos(ph), sin(ph));
        }
        host_fft(g, true);
        const double ang = k * length;
        t.scale_re = amp * cos(ang);
        t.scale_im = amp * sin(ang);
    }
    PA_CUDA(cudaMalloc(&t.dev, (size_t)n * c->csize()));
    if (c->prec == 0) {
        std::vector<float2> h(n);
        for (int i = 0; i < n; ++i) h[i] = make_float2((float)g[i].real(), (float)g[i].imag());
        PA_CUDA(cudaMemcpy(t.dev, h.data(), (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
    } else {
        std::vector<double2> h(n);
        for (int i = 0; i < n; ++i) h[i] = make_double2(g[i].real(), g[i].imag());
        PA_CUDA(cudaMemcpy(t.dev, h.data(), (size_t)n * sizeof(double2), cudaMemcpyHostToDevice));
    }
    c->seps.push_back(t);
    *out = &c->seps.back();
    return PA_OK;
}

// request to reduce the output field inside the final row pass (see fft_passes.cuh, MEAS)
struct FusedMeasure {
    double* rowsums;
    const float* pupils_dev;
    int npupil;
    int store;       // 0: the field itself is not written back
    bool done;       // set by the propagator when the final pass did reduce
};

// ---- pass helpers -----------------------------------------------------------------------------------------
static int rows(pa_ctx* c, void* field, int batch, bool in_perm, bool out_perm, bool src, const void* turns, double scale,
                double amp, double aw, double ac, cudaStream_t st, const SepTable* sep = nullptr, const struct FusedMeasure* fm = nullptr) {
    RowLaunch r;
    r.field = field;
    r.tw = c->tw;
    r.turns = turns;
    r.scale = scale;
    r.rows_total = batch * c->n;
    r.in_perm = in_perm;
    r.out_perm = out_perm;
    r.src = src;
    r.x = c->x;
    r.y = c->y;
    r.amp = amp;
    r.aw = aw;
    r.ac = ac;
    r.rowsums = fm ? fm->rowsums : nullptr;
    r.pupils = fm ? fm->pupils_dev : nullptr;
    r.npupil = fm ? fm->npupil : 0;
    r.store = fm ? fm->store : 1;
    r.sep = sep ? sep->dev : nullptr;
    r.sep_re = sep ? sep->scale_re : 0.0;
    r.sep_im = sep ? sep->scale_im : 0.0;
    r.use_tma = c->rows_tma;
    r.num_sms = c->num_sms;
    note(1);
    return check_launch(launch_rows(c->prec, c->n, r, st), "row pass");
}
static int cols(pa_ctx* c, void* field, int batch, double length, double wvl, cudaStream_t st) {
    const HTable* h = nullptr;
    int rc = get_htable(c, length, wvl, &h);
    if (rc) return rc;
    ColLaunch cl;
    cl.field = field;
    cl.tw = c->tw;
    cl.hp = h->dev;
    cl.alpha_re = h->alpha_re;
    cl.alpha_im = h->alpha_im;
    cl.batch = batch;
    cl.tmap = nullptr;
    cl.num_sms = c->num_sms;
    if (c->split || (c->use_tma && fft_tma_supported(c->prec, c->n))) {
        const CUtensorMap* tm = nullptr;
        rc = get_tensor_map(c, field, batch, &tm);
        if (rc) return rc;
        cl.tmap = tm;
    }
    if (c->split) {
        cl.hpy = h->dev_y;
        cl.tw_sub = c->tw_sub;
        cl.otw = c->otw;
        note(2);
    }
    note(1);
    return check_launch(launch_cols(c->prec, c->n, cl, st), "column pass");
}

static void source_params(double w0, double wvl, double F0, double* amp, double* aw, double* ac) {
    *amp = sqrt(2 / M_PI) / w0;
    *aw = 1 / (w0 * w0);
    *ac = isinf(F0) ? 0.0 : 2 * M_PI / wvl / 2 / F0;
}

static int screens(pa_ctx* c, const float* fx, const float* fy, const float* coef, int m, int m_split, int degree,
                   double shift_x, double shift_y, int nscreens, void* turns, void* phi, int phi_f64, int method,
                   double coef_bound, cudaStream_t st, int tc_phase = 2, int tc_first = 0, int tc_total = -1) {
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called before generating screens");
    PA_REQUIRE(m > 0 && m_split >= 0 && m_split <= m, "bad m / m_split (%d, %d)", m, m_split);
    PA_REQUIRE(degree >= -1 && degree <= kMaxPolyDegree, "polynomial degree %d outside [-1, %d]", degree, kMaxPolyDegree);
    PA_REQUIRE(m_split == 0 || degree >= 0, "m_split > 0 needs degree >= 0");
    PA_REQUIRE(method == PA_SCREEN_EXACT || method == PA_SCREEN_TC, "unknown screen method %d", method);
    PA_REQUIRE(method == PA_SCREEN_EXACT || (c->prec == PA_C64 && c->n % 256 == 0),
               "the tensor-core screen method needs a complex64 context and a grid size that is a multiple of 256");
    const int n = c->n;
    const int k2 = 2 * (m - m_split);
    if (method == PA_SCREEN_EXACT)
        PA_REQUIRE(c->pq_bytes >= (size_t)nscreens * (k2 > 0 ? k2 : 1) * n * sizeof(double), "screen workspace not reserved");
    ScreenLaunch a;
    a.n = n;
    a.m = m;
    a.m_split = m_split;
    a.degree = m_split > 0 ? degree : -1;
    a.nscreens = nscreens;
    a.x = c->x;
    a.y = c->y;
    a.shift_x = (float)shift_x;
    a.shift_y = (float)shift_y;
    double x0 = 0, y0 = 0;
    for (int i = 0; i < n; ++i) {
        x0 = fmax(x0, fabs((double)(c->hx[i] + a.shift_x)));
        y0 = fmax(y0, fabs((double)(c->hy[i] + a.shift_y)));
    }
    if (x0 == 0) x0 = 1;
    if (y0 == 0) y0 = 1;
    a.x0 = x0;
    a.y0 = y0;
    a.inv_x0 = 1 / x0;
    a.inv_y0 = 1 / y0;
    a.x_first = (double)c->hx[0] + (double)a.shift_x;
    a.dxu = ((double)c->hx[n - 1] - (double)c->hx[0]) / (double)(n - 1);
    a.y_first = (double)c->hy[0] + (double)a.shift_y;
    a.dyu = ((double)c->hy[n - 1] - (double)c->hy[0]) / (double)(n - 1);
    a.fx = fx;
    a.fy = fy;
    a.coef = (const float2*)coef;
    a.P = c->P;
    a.Q = c->Q;
    a.polyc = c->polyc;
    a.turns = turns;
    a.turns_f64 = c->prec == 1;
    a.phi = phi;
    a.phi_f64 = phi_f64;
    a.p_scale = 1024.0;
    if (method == PA_SCREEN_TC) {
        // keep |P| * p_scale below 2^15 (fp16 max 65504): p_scale = 2^(15 - ceil(log2 bound)), clipped to [1, 1024]
        const double bound = coef_bound > 0 ? coef_bound : 32.0;
        int e = 15 - (int)ceil(log2(bound));
        e = e < 0 ? 0 : (e > 10 ? 10 : e);
        a.p_scale = ldexp(1.0, e);
        const size_t need = screen_tc_workspace(n, m, m_split, tc_total > 0 ? tc_total : nscreens);
        PA_REQUIRE(c->tcws_bytes >= need, "tensor-core screen workspace not reserved");
        static const bool fork = !(getenv("PYATM_TC_FORK") && atoi(getenv("PYATM_TC_FORK")) == 0);
        cudaStream_t fstream = nullptr;
        if (tc_phase != 1) {            // preparation phase
            if (fork && a.degree >= 0) {
                // fork: the operand generation (HBM-store-bound) runs on the context's second stream while the polynomial
                // coefficients and node tables (float64, latency-bound small grids) run on `st`; joined before the contraction
                if (!c->aux) {
                    PA_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
                    PA_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
                    PA_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
                }
                PA_CUDA(cudaEventRecord(c->ev_fork, st));
                PA_CUDA(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
                fstream = c->aux;
            }
            int rc = launch_screen_poly(a, st);      // polynomial coefficients
            if (rc) return check_launch(rc, "screen polynomial");
        }
        static const int swap = getenv("PYATM_TC_SWAP") ? atoi(getenv("PYATM_TC_SWAP")) : 0;
        note(tc_phase == 1 ? 1 : (tc_phase == 0 ? 5 : 6));      // preparation: polynomial coefficients, operands, three node kernels
        return check_launch(launch_screen_tc(a, c->tcws, c->tc_err, c->num_sms, swap, st, tc_phase, tc_first, tc_total, fstream, c->ev_join),
                            "tensor-core screen synthesis");
    }
    note(3);
    return check_launch(launch_screen_exact(a, st), "screen synthesis");
}

static int ensure_screen_ws(pa_ctx* c, int nscreens, int m, int m_split, int degree, int method) {
    if (method == PA_SCREEN_TC) {
        int rc = grow(&c->tcws, &c->tcws_bytes, screen_tc_workspace(c->n, m, m_split, nscreens));
        if (rc) return rc;
        if (!c->tc_err) {
            PA_CUDA(cudaMalloc((void**)&c->tc_err, sizeof(int)));
            PA_CUDA(cudaMemset(c->tc_err, 0, sizeof(int)));
        }
        const size_t pneed = (size_t)nscreens * (degree + 2) * (degree + 2) * sizeof(double);
        return grow((void**)&c->polyc, &c->polyc_bytes, pneed);
    }
    const int k2 = 2 * (m - m_split);
    const size_t need = (size_t)nscreens * (k2 > 0 ? k2 : 1) * c->n * sizeof(double);
    if (c->pq_bytes < need) {
        if (c->P) PA_CUDA(cudaFree(c->P));
        if (c->Q) PA_CUDA(cudaFree(c->Q));
        c->P = c->Q = nullptr;
        c->pq_bytes = 0;
        PA_CUDA(cudaMalloc((void**)&c->P, need));
        PA_CUDA(cudaMalloc((void**)&c->Q, need));
        c->pq_bytes = need;
    }
    const size_t pneed = (size_t)nscreens * (degree + 2) * (degree + 2) * sizeof(double);
    return grow((void**)&c->polyc, &c->polyc_bytes, pneed);
}

// ---- exported functions ----------------------------------------------------------------------------------
extern "C" {

int pa_version(void) { return PA_VERSION; }
const char* pa_last_error(void) { return last_error(); }
int pa_device_count(int* count) {
    PA_CUDA(cudaGetDeviceCount(count));
    return PA_OK;
}
unsigned long long pa_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int pa_ctx_create(pa_ctx** out, int device, int n, int precision) {
    PA_REQUIRE(out != nullptr, "null ctx pointer");
    PA_REQUIRE(precision == PA_C64 || precision == PA_C128, "precision must be PA_C64 or PA_C128");
    PA_REQUIRE(fft_size_supported(precision, n), "grid size %d unsupported: power of two in [64, 8192] required", n);
    PA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PA_CUDA(cudaGetDeviceProperties(&prop, device));
    PA_REQUIRE(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    pa_ctx* c = new pa_ctx();
    c->device = device;
    c->n = n;
    c->prec = precision;
    c->e = elems_per_thread(precision);
    const bool direct = getenv("PYATM_FFT_DIRECT") && atoi(getenv("PYATM_FFT_DIRECT")) != 0;
    c->split = direct ? 0 : fft_split_radix(precision, n);
    build_perm(c);
    int rc = precision == 0 ? build_twiddles<float>(n, c->e, &c->tw) : build_twiddles<double>(n, c->e, &c->tw);
    if (!rc && c->split) {
        rc = precision == 0 ? build_twiddles<float>(n / c->split, c->e, &c->tw_sub) : build_twiddles<double>(n / c->split, c->e, &c->tw_sub);
        if (!rc) rc = precision == 0 ? build_outer_twiddles<float>(n, c->split, &c->otw) : build_outer_twiddles<double>(n, c->split, &c->otw);
    }
    if (rc) {
        pa_ctx_destroy(c);
        return rc;
    }
    rc = screen_init_constants();
    if (rc) {
        set_error("constant upload failed: %s", cudaGetErrorString((cudaError_t)rc));
        delete c;
        return PA_ERR_CUDA;
    }
    c->htabs.reserve(256);
    c->seps.reserve(64);
    c->num_sms = prop.multiProcessorCount;
    c->use_tma = !direct;
    // row pass: the direct-access kernel everywhere except 8192^2 complex64, where one row is a whole 64 KiB tile and the
    // TMA-fed ring is a little ahead (473 against 484 us; at 2048^2 / 4096^2 it loses: 145 / 234 against 132 / 150)
    c->rows_tma = getenv("PYATM_FFT_ROWS_TMA") ? atoi(getenv("PYATM_FFT_ROWS_TMA")) != 0 : ((n == 8192 && precision == PA_C64) || precision == PA_C128);
    c->sep_first_leg = !(getenv("PYATM_NO_ANALYTIC_LEG") && atoi(getenv("PYATM_NO_ANALYTIC_LEG")) != 0);
    *out = c;
    return PA_OK;
}

int pa_ctx_destroy(pa_ctx* c) {
    if (!c) return PA_OK;
    cudaSetDevice(c->device);
    void* ptrs[] = {c->x, c->y, c->tw, c->tw_sub, c->otw, c->turns, c->P, c->Q, c->polyc, c->partials, c->field, c->spec, c->spec_all, c->pupils, c->table, c->tcws, c->tc_err, c->rowsums, c->perm_dev, c->fftws};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& h : c->htabs) {
        if (h.dev) cudaFree(h.dev);
        if (h.dev_y) cudaFree(h.dev_y);
    }
    for (auto& t : c->seps)
        if (t.dev) cudaFree(t.dev);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
    return PA_OK;
}

int pa_ctx_set_axes(pa_ctx* c, const float* x_host, const float* y_host, double delta) {
    PA_REQUIRE(c && x_host && y_host && delta > 0, "bad arguments to pa_ctx_set_axes");
    PA_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->n * sizeof(float);
    if (!c->x) PA_CUDA(cudaMalloc((void**)&c->x, bytes));
    if (!c->y) PA_CUDA(cudaMalloc((void**)&c->y, bytes));
    PA_CUDA(cudaMemcpy(c->x, x_host, bytes, cudaMemcpyHostToDevice));
    PA_CUDA(cudaMemcpy(c->y, y_host, bytes, cudaMemcpyHostToDevice));
    c->hx.assign(x_host, x_host + c->n);
    c->hy.assign(y_host, y_host + c->n);
    if (c->delta != delta) {
        for (auto& h : c->htabs) {
            if (h.dev) cudaFree(h.dev);
            if (h.dev_y) cudaFree(h.dev_y);
        }
        c->htabs.clear();
    }
    for (auto& t : c->seps)
        if (t.dev) cudaFree(t.dev);
    c->seps.clear();
    c->delta = delta;
    c->axes_set = true;
    return PA_OK;
}

int pa_ctx_permutation(pa_ctx* c, int* perm_host) {
    PA_REQUIRE(c && perm_host, "bad arguments");
    memcpy(perm_host, c->perm.data(), (size_t)c->n * sizeof(int));
    return PA_OK;
}

int pa_ctx_fft_geometry(pa_ctx* c, int* g) {
    PA_REQUIRE(c && g, "bad arguments");
    fft_geometry(c->prec, c->n, &g[0], &g[1], &g[2], &g[3], &g[4], &g[5]);
    return PA_OK;
}

int pa_source_gaussian(pa_ctx* c, void* field, int batch, double w0, double wvl, double F0, void* stream) {
    PA_REQUIRE(c && field && batch > 0 && w0 > 0 && wvl > 0, "bad arguments to pa_source_gaussian");
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called first");
    double amp, aw, ac;
    source_params(w0, wvl, F0, &amp, &aw, &ac);
    return rows(c, field, batch, false, false, true, nullptr, 1.0, amp, aw, ac, (cudaStream_t)stream);
}

int pa_vacuum_leg(pa_ctx* c, void* field, int batch, double length, double wvl, void* stream) {
    PA_REQUIRE(c && field && batch > 0 && wvl > 0, "bad arguments to pa_vacuum_leg");
    if (!(length > 0)) return PA_OK;   // pathes.py:30,39-40
    cudaStream_t st = (cudaStream_t)stream;
    int rc = rows(c, field, batch, false, true, false, nullptr, 1.0, 0, 0, 0, st);
    if (rc) return rc;
    rc = cols(c, field, batch, length, wvl, st);
    if (rc) return rc;
    return rows(c, field, batch, true, false, false, nullptr, 1.0, 0, 0, 0, st);
}

int pa_screen_ss(pa_ctx* c, const float* fx, const float* fy, const float* coef, int m, int m_split, int degree,
                 double shift_x, double shift_y, int nscreens, void* turns, void* phi, int phi_f64, int method, double coef_bound,
                 void* stream) {
    PA_REQUIRE(c && fx && fy && coef && nscreens > 0, "bad arguments to pa_screen_ss");
    PA_REQUIRE(turns || phi, "pa_screen_ss needs at least one output");
    int rc = ensure_screen_ws(c, nscreens, m, m_split, degree, method);
    if (rc) return rc;
    return screens(c, fx, fy, coef, m, m_split, degree, shift_x, shift_y, nscreens, turns, phi, phi_f64, method, coef_bound,
                   (cudaStream_t)stream);
}

int pa_screen_fft(pa_ctx* c, const void* spectrum, int nscreens, const double* terms_host, int nterms, void* out_complex,
                  void* out_real, void* stream) {
    PA_REQUIRE(c && spectrum && nscreens > 0 && nterms >= 0 && (nterms == 0 || terms_host), "bad arguments to pa_screen_fft");
    PA_REQUIRE(out_complex || out_real, "pa_screen_fft needs at least one output");
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called before pa_screen_fft");
    PA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = c->n;
    const size_t plane = (size_t)n * n;
    int rc = grow(&c->field, &c->field_bytes, (size_t)nscreens * plane * c->csize());
    if (rc) return rc;
    if (!c->perm_dev) {
        PA_CUDA(cudaMalloc((void**)&c->perm_dev, (size_t)n * sizeof(int)));
        PA_CUDA(cudaMemcpy(c->perm_dev, c->perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    // group the terms of every screen by their x-frequency (see screen_fft.cu): stable order of first appearance
    std::vector<double> sorted((size_t)nscreens * nterms * 4);
    std::vector<std::vector<int>> offs(nscreens);
    int ngroups = 0;
    for (int b = 0; b < nscreens; ++b) {
        const double* tb = terms_host + (size_t)b * nterms * 4;
        std::vector<double> keys;
        std::vector<std::vector<int>> members;
        for (int t = 0; t < nterms; ++t) {
            size_t g = 0;
            while (g < keys.size() && keys[g] != tb[t * 4]) ++g;
            if (g == keys.size()) {
                keys.push_back(tb[t * 4]);
                members.emplace_back();
            }
            members[g].push_back(t);
        }
        int at = 0;
        offs[b].push_back(0);
        for (const auto& m : members) {
            for (int t : m) memcpy(&sorted[((size_t)b * nterms + at++) * 4], tb + t * 4, 4 * sizeof(double));
            offs[b].push_back(at);
        }
        ngroups = std::max(ngroups, (int)members.size());
    }
    std::vector<int> goff((size_t)nscreens * (ngroups + 1));
    for (int b = 0; b < nscreens; ++b)
        for (int g = 0; g <= ngroups; ++g) goff[(size_t)b * (ngroups + 1) + g] = offs[b][std::min<size_t>(g, offs[b].size() - 1)];
    // workspace layout (doubles): terms | ex | gy | partials | rowsum | group offsets (ints)
    const size_t per_row = (size_t)(n + 255) / 256;
    const size_t n_terms = (size_t)nscreens * nterms * 4, n_tab = (size_t)nscreens * ngroups * n * 2;
    const size_t n_part = (size_t)nscreens * n * per_row * 2, n_rows = (size_t)nscreens * n * 2;
    rc = grow(&c->fftws, &c->fftws_bytes, (n_terms + 2 * n_tab + n_part + n_rows + 8) * sizeof(double) + goff.size() * sizeof(int));
    if (rc) return rc;
    double* w = (double*)c->fftws;
    FftScreenLaunch a;
    a.n = n;
    a.nscreens = nscreens;
    a.spectrum = spectrum;
    a.ws = c->field;
    a.perm = c->perm_dev;
    a.terms = w;
    a.nterms = nterms;
    a.ngroups = ngroups;
    a.x = c->x;
    a.y = c->y;
    a.ex = (double2*)(w + ((n_terms + 1) & ~(size_t)1));
    a.ey = a.ex + n_tab / 2;
    a.partials = a.ey + n_tab / 2;
    a.rowsum = a.partials + n_part / 2;
    int* goff_dev = (int*)(a.rowsum + n_rows / 2);
    a.goff = goff_dev;
    a.out_complex = out_complex;
    a.out_real = out_real;
    if (nterms > 0) {
        PA_CUDA(cudaStreamSynchronize(st));      // an earlier call on this stream may still be reading the term tables
        PA_CUDA(cudaMemcpy(goff_dev, goff.data(), goff.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (nterms > 0) PA_CUDA(cudaMemcpy(w, sorted.data(), n_terms * sizeof(double), cudaMemcpyHostToDevice));
    note(1);
    rc = check_launch(launch_fftscreen_gather(c->prec, a, st), "FFT screen gather");
    if (rc) return rc;
    ColLaunch cl;                       // IFFT_y of every column (direct kernel, storage order in, natural out)
    cl.field = c->field;
    cl.tw = c->tw;
    cl.hp = nullptr;
    cl.alpha_re = 1.0;
    cl.alpha_im = 0.0;
    cl.batch = nscreens;
    cl.tmap = nullptr;
    cl.num_sms = c->num_sms;
    cl.inv_only = true;
    note(1);
    rc = check_launch(launch_cols(c->prec, n, cl, st), "FFT screen column pass");
    if (rc) return rc;
    rc = rows(c, c->field, nscreens, true, false, false, nullptr, 1.0, 0, 0, 0, st);     // IFFT_x of every row
    if (rc) return rc;
    note(fftscreen_finish_launches(a));
    return check_launch(launch_fftscreen_finish(c->prec, a, st), "FFT screen finish");
}

// utils.py:42-50: the centred transform pair.  Both directions run the inverse half of the split-step passes (as
// pa_screen_fft does): S(v)[i][j] = sum_pq v[p][q] e^{+2 pi i ((i-c)(p-c) + (j-c)(q-c)) / N}, c = N/2, is the unnormalised
// centred inverse, and the forward transform is conj(S(conj(u))).
int pa_fft2c(pa_ctx* c, const void* in, void* out, int batch, int forward, double scale, void* stream) {
    PA_REQUIRE(c && in && out && batch > 0, "bad arguments to pa_fft2c");
    PA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = c->n;
    const size_t plane = (size_t)n * n;
    int rc = grow(&c->field, &c->field_bytes, (size_t)batch * plane * c->csize());
    if (rc) return rc;
    if (!c->perm_dev) {
        PA_CUDA(cudaMalloc((void**)&c->perm_dev, (size_t)n * sizeof(int)));
        PA_CUDA(cudaMemcpy(c->perm_dev, c->perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    note(1);
    rc = check_launch(launch_fft2c_gather(c->prec, in, c->field, c->perm_dev, n, batch, forward != 0, st), "fft2c gather");
    if (rc) return rc;
    ColLaunch cl;
    cl.field = c->field;
    cl.tw = c->tw;
    cl.hp = nullptr;
    cl.alpha_re = scale;
    cl.alpha_im = 0.0;
    cl.batch = batch;
    cl.tmap = nullptr;
    cl.num_sms = c->num_sms;
    cl.inv_only = true;
    note(1);
    rc = check_launch(launch_cols(c->prec, n, cl, st), "fft2c column pass");
    if (rc) return rc;
    rc = rows(c, c->field, batch, true, false, false, nullptr, 1.0, 0, 0, 0, st);
    if (rc) return rc;
    note(1);
    return check_launch(launch_copy_conj(c->prec, c->field, out, (size_t)batch * plane, forward != 0, st), "fft2c finish");
}

// theory/sources.py:16-18 GaussianBeam.amplitude on caller-supplied squared radii
int pa_gaussian_amplitude(pa_ctx* c, const void* r2, void* out, size_t count, double w0, double wvl, double F0, void* stream) {
    PA_REQUIRE(c && r2 && out && count > 0 && w0 > 0 && wvl > 0, "bad arguments to pa_gaussian_amplitude");
    PA_CUDA(cudaSetDevice(c->device));
    double amp, aw, ac;
    source_params(w0, wvl, F0, &amp, &aw, &ac);
    note(1);
    return check_launch(launch_gaussian_amplitude(c->prec, r2, out, count, amp, aw, ac, (cudaStream_t)stream), "gaussian amplitude");
}

int pa_apply_screen(pa_ctx* c, void* field, int batch, const void* turns, double scale, void* stream) {
    PA_REQUIRE(c && field && batch > 0, "bad arguments to pa_apply_screen");
    return rows(c, field, batch, false, false, false, turns, scale, 0, 0, 0, (cudaStream_t)stream);
}

int pa_phase_to_turns(pa_ctx* c, const void* phi, int phi_f64, void* turns, size_t count, void* stream) {
    PA_REQUIRE(c && phi && turns, "bad arguments to pa_phase_to_turns");
    note(1);
    return check_launch(launch_phase_to_turns(phi, phi_f64, turns, c->prec == 1, count, (cudaStream_t)stream), "phase_to_turns");
}

int pa_intensity(pa_ctx* c, const void* field, void* out, int batch, void* stream) {
    PA_REQUIRE(c && field && out && batch > 0, "bad arguments to pa_intensity");
    note(1);
    return check_launch(launch_intensity(c->prec, field, out, (size_t)batch * c->n * c->n, (cudaStream_t)stream), "intensity");
}

int pa_pupil_apply(pa_ctx* c, const void* in, void* out, int batch, double radius, double sx, double sy, void* stream) {
    PA_REQUIRE(c && in && out && batch > 0, "bad arguments to pa_pupil_apply");
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called first");
    note(1);
    return check_launch(launch_pupil(c->prec, in, out, c->x, c->y, c->n, batch, (float)(radius * radius), (float)sx, (float)sy,
                                     (cudaStream_t)stream), "pupil");
}

int pa_measure(pa_ctx* c, const void* field, int batch, const float* pupils, int npupil, int per_field, double* out,
               int out_stride, void* stream) {
    PA_REQUIRE(c && field && out && batch > 0, "bad arguments to pa_measure");
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called first");
    PA_REQUIRE(npupil >= 0 && npupil <= kMaxPupils, "at most %d apertures per pa_measure call", kMaxPupils);
    PA_REQUIRE(npupil == 0 || pupils, "pupil table missing");
    PA_REQUIRE(out_stride >= kMeasureHead + npupil, "out_stride too small");
    const int nparts = c->n / 8 < 148 ? c->n / 8 : 148;
    int rc = grow((void**)&c->partials, &c->partials_bytes, (size_t)batch * nparts * (kRawMoments + kMaxPupils) * sizeof(double));
    if (rc) return rc;
    MeasureLaunch a;
    a.field = field;
    a.n = c->n;
    a.batch = batch;
    a.x = c->x;
    a.y = c->y;
    a.delta2 = c->delta * c->delta;
    a.pupils = pupils;
    a.npupil = npupil;
    a.pupils_per_field = per_field;
    a.partials = c->partials;
    a.nparts = nparts;
    a.out = out;
    a.out_stride = out_stride;
    note(2);
    return check_launch(launch_measure(c->prec, a, (cudaStream_t)stream), "measure");
}

int pa_histogram(pa_ctx* c, const double* values, size_t stride, size_t count, const double* edges, int nbins,
                 unsigned long long* counts, void* stream) {
    PA_REQUIRE(c && values && edges && counts && nbins > 0, "bad arguments to pa_histogram");
    if (count == 0) return PA_OK;
    note(1);
    return check_launch(launch_histogram(values, stride, count, edges, nbins, counts, (cudaStream_t)stream), "histogram");
}

int pa_rng_spectrum(pa_ctx* c, unsigned long long seed, unsigned long long realization0, int batch, int screen0, int nscreens,
                    int m, const float* edges, const float* psd, float* fx, float* fy, float* coef, void* stream) {
    PA_REQUIRE(c && edges && psd && fx && fy && coef && batch > 0 && nscreens > 0 && m > 0, "bad arguments to pa_rng_spectrum");
    RngLaunch a;
    a.seed = seed;
    a.realization0 = realization0;
    a.realization_stride = 1;
    a.screen0 = screen0;
    a.nscreens = nscreens;
    a.batch = batch;
    a.m = m;
    a.base = edges;
    a.psd = psd;
    a.fx = fx;
    a.fy = fy;
    a.coef = (float2*)coef;
    a.rho = nullptr;
    a.theta = nullptr;
    note(1);
    return check_launch(launch_rng_spectrum(a, (cudaStream_t)stream), "rng_spectrum");
}

int pa_fft_pass(pa_ctx* c, void* field, int batch, int kind, const void* turns, double length, double wvl, void* stream) {
    PA_REQUIRE(c && field && batch > 0 && (kind == 0 || kind == 1), "bad arguments to pa_fft_pass");
    if (kind == 0) return cols(c, field, batch, length, wvl, (cudaStream_t)stream);
    return rows(c, field, batch, true, true, false, turns, 1.0, 0, 0, 0, (cudaStream_t)stream);
}

// Token stream of one realization: SRC, then for every screen [leg] screen, then [closing leg].  A leg is
// FFT_x | columns | IFFT_x; all row-level tokens between two column passes are fused into one k_rows launch.
static int propagate_impl(pa_ctx* c, const pa_path* p, void* field, int batch, const float* fx, const float* fy, const float* coef,
                          void* stream, FusedMeasure* fm) {
    PA_REQUIRE(c && p && field && batch > 0, "bad arguments to pa_propagate");
    PA_REQUIRE(p->n_screens >= 0 && p->leg_lengths_host, "bad path description");
    PA_REQUIRE(p->n_screens == 0 || (fx && fy && coef && p->screen_scale_host), "screen coefficients missing");
    PA_REQUIRE(c->axes_set, "pa_ctx_set_axes must be called first");
    cudaStream_t st = (cudaStream_t)stream;
    const int S = p->n_screens, n = c->n;
    int rc;
    const bool tc_hoist = S > 0 && p->screen_method == PA_SCREEN_TC;
    if (S > 0) {
        // tensor-core method: operands and polynomial tables of ALL S x batch screens are prepared by three launches up
        // front (the coefficients of every path position are known), the per-position work is the contraction alone
        rc = ensure_screen_ws(c, tc_hoist ? S * batch : batch, p->m, p->m_split, p->degree, p->screen_method);
        if (rc) return rc;
        rc = grow(&c->turns, &c->turns_bytes, (size_t)batch * n * n * c->rsize());
        if (rc) return rc;
    }
    double amp, aw, ac;
    source_params(p->w0, p->wvl, p->F0, &amp, &aw, &ac);

    // state of the field between launches
    bool have_field = p->from_field != 0;   // false: the source has not been materialised yet
    bool perm = false;         // true: field is in row-spectrum form awaiting IFFT_x
    // one fused row launch: [SRC | IFFT_x]? [screen]? [FFT_x]?
    const SepTable* start_sep = nullptr;   // set when the source is carried analytically through the first leg
    const bool can_fuse = fm != nullptr && (c->n / c->e) >= 32;
    auto row_launch = [&](bool want_perm_out, const void* turns, double scale, bool final_pass = false) -> int {
        const bool src = !have_field;
        const bool fuse = final_pass && can_fuse && perm && !want_perm_out && !src;
        int r = rows(c, field, batch, perm, want_perm_out, src, turns, scale, amp, aw, ac, st, src ? start_sep : nullptr, fuse ? fm : nullptr);
        if (fuse) fm->done = true;
        have_field = true;
        perm = want_perm_out;
        return r;
    };
    if (tc_hoist) {
        rc = screens(c, fx, fy, coef, p->m, p->m_split, p->degree, p->shift_x, p->shift_y, S * batch, nullptr, nullptr, 0,
                     p->screen_method, p->coef_bound, st, 0, 0, S * batch);
        if (rc) return rc;
    }
    auto gen_screen = [&](int i) -> int {   // coefficient arrays are [S][batch][m]: one contiguous slab per path position
        const size_t o = (size_t)i * batch * p->m;
        if (tc_hoist)
            return screens(c, fx + o, fy + o, coef + 2 * o, p->m, p->m_split, p->degree, p->shift_x, p->shift_y, batch, c->turns,
                           nullptr, 0, p->screen_method, p->coef_bound, st, 1, i * batch, S * batch);
        return screens(c, fx + o, fy + o, coef + 2 * o, p->m, p->m_split, p->degree, p->shift_x, p->shift_y, batch, c->turns,
                       nullptr, 0, p->screen_method, p->coef_bound, st);
    };

    const bool analytic_first_leg = !have_field && S > 0 && c->prec == PA_C64 && c->sep_first_leg;
    if (analytic_first_leg) {
        rc = get_sep_table(c, p->leg_lengths_host[0], p->wvl, p->w0, p->F0, &start_sep);
        if (rc) return rc;
    }
    for (int i = 0; i < S; ++i) {
        const double L = p->leg_lengths_host[i];
        if (L > 0 && !(i == 0 && analytic_first_leg)) {
            if (!perm) {                       // leading FFT_x of this leg was not fused into a previous launch
                rc = row_launch(true, nullptr, 1.0);
                if (rc) return rc;
            }
            rc = cols(c, field, batch, L, p->wvl, st);
            if (rc) return rc;
        }
        rc = gen_screen(i);
        if (rc) return rc;
        const double Lnext = p->leg_lengths_host[i + 1];
        double scale = p->screen_scale_host[i];
        const bool fuse_next = Lnext > 0;
        if (i == S - 1 && !fuse_next) scale *= p->final_scale;
        rc = row_launch(fuse_next, c->turns, scale, i == S - 1 && !fuse_next);
        if (rc) return rc;
    }
    const double Llast = p->leg_lengths_host[S];
    if (Llast > 0) {
        if (!perm) {
            rc = row_launch(true, nullptr, 1.0);
            if (rc) return rc;
        }
        rc = cols(c, field, batch, Llast, p->wvl, st);
        if (rc) return rc;
        rc = row_launch(false, nullptr, p->final_scale, true);
        if (rc) return rc;
    } else if (S == 0) {
        rc = row_launch(false, nullptr, p->final_scale);
        if (rc) return rc;
    }
    return PA_OK;
}

int pa_propagate(pa_ctx* c, const pa_path* p, void* field, int batch, const float* fx, const float* fy, const float* coef,
                 void* stream) {
    return propagate_impl(c, p, field, batch, fx, fy, coef, stream, nullptr);
}

// Realizations per pass of the fused path inside pa_simulate_batch*: a batch of any size is processed in chunks of this
// many realizations (32 at 2048^2: 1.5 GiB of field + screen and 2.4 GB of tensor-core operands; measured realizations/s
// against the chunk size, device-resident: 4: 3056, 8: 3292, 16: 3385, 32: 3472 -- fewer launches and kernel tails per
// realization; more at smaller grids, fewer at the long-haul sizes).  PYATM_SIM_CHUNK overrides.
static int sim_chunk(const pa_ctx* c) {
    static const int forced = getenv("PYATM_SIM_CHUNK") ? atoi(getenv("PYATM_SIM_CHUNK")) : 0;
    if (forced > 0) return forced;
    const long long want = 32LL * 2048 * 2048 / ((long long)c->n * c->n);
    return (int)(want < 2 ? 2 : (want > 64 ? 64 : want));
}

// one chunk: coefficients already on the device, [S][batch][m] contiguous
static int simulate_chunk(pa_ctx* c, const pa_path* p, int batch, const float* fx, const float* fy, const float* coef,
                          const float* pupils_dev, int npupil, double* table_dev, int out_stride, cudaStream_t st) {
    const int n = c->n;
    // statistics only: reduce inside the final row pass and do not write the output field at all
    FusedMeasure fm{nullptr, pupils_dev, npupil, 0, false};
    FusedMeasure* fmp = nullptr;
    int rc;
    if (npupil <= kFusedPupils && !(getenv("PYATM_NO_FUSED_MEASURE") && atoi(getenv("PYATM_NO_FUSED_MEASURE")) != 0)) {
        rc = grow((void**)&c->rowsums, &c->rowsums_bytes, (size_t)batch * n * kRowSums * sizeof(double));
        if (rc) return rc;
        fm.rowsums = c->rowsums;
        fmp = &fm;
    }
    rc = propagate_impl(c, p, c->field, batch, fx, fy, coef, st, fmp);
    if (rc) return rc;
    if (fm.done) {
        PA_REQUIRE(out_stride >= kMeasureHead + npupil, "out_stride too small");
        note(1);
        return check_launch(launch_measure_rows(c->rowsums, c->y, n, batch, c->delta * c->delta, npupil, table_dev, out_stride, st),
                            "measure (fused)");
    }
    return pa_measure(c, c->field, batch, pupils_dev, npupil, 0, table_dev, out_stride, st);
}

static int simulate_common(pa_ctx* c, const pa_path* p, int batch, const float* fx_host, const float* fy_host,
                           const float* coef_host, unsigned long long seed, unsigned long long realization0,
                           const float* edges, const float* psd, const float* pupils_dev, int npupil, double* table_dev,
                           int out_stride, cudaStream_t st) {
    const int S = p->n_screens, m = p->m, n = c->n;
    const int chunk = std::min(batch, sim_chunk(c));
    int rc = grow(&c->field, &c->field_bytes, (size_t)chunk * n * n * c->csize());
    if (rc) return rc;
    const size_t rows = (size_t)(S > 0 ? S : 1);
    const size_t cnt = (size_t)chunk * rows * m;            // staging of one chunk: fx | fy | coef
    rc = grow((void**)&c->spec, &c->spec_bytes, cnt * 4 * sizeof(float));
    if (rc) return rc;
    float* fx = c->spec;
    float* fy = c->spec + cnt;
    float* coef = c->spec + 2 * cnt;
    const float* all_fx = nullptr;
    if (S > 0 && coef_host) {
        const size_t all = (size_t)batch * rows * m;
        if (batch == chunk) {
            PA_CUDA(cudaMemcpyAsync(fx, fx_host, all * sizeof(float), cudaMemcpyHostToDevice, st));
            PA_CUDA(cudaMemcpyAsync(fy, fy_host, all * sizeof(float), cudaMemcpyHostToDevice, st));
            PA_CUDA(cudaMemcpyAsync(coef, coef_host, all * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
        } else {
            // the whole batch travels host -> device once; every chunk is then gathered into the contiguous staging
            rc = grow((void**)&c->spec_all, &c->spec_all_bytes, all * 4 * sizeof(float));
            if (rc) return rc;
            all_fx = c->spec_all;
            PA_CUDA(cudaMemcpyAsync(c->spec_all, fx_host, all * sizeof(float), cudaMemcpyHostToDevice, st));
            PA_CUDA(cudaMemcpyAsync(c->spec_all + all, fy_host, all * sizeof(float), cudaMemcpyHostToDevice, st));
            PA_CUDA(cudaMemcpyAsync(c->spec_all + 2 * all, coef_host, all * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
        }
    }
    for (int c0 = 0; c0 < batch; c0 += chunk) {
        const int b = std::min(chunk, batch - c0);
        if (S > 0 && coef_host && all_fx) {
            const size_t all = (size_t)batch * rows * m, spitch = (size_t)batch * m * sizeof(float), w = (size_t)b * m * sizeof(float);
            PA_CUDA(cudaMemcpy2DAsync(fx, w, all_fx + (size_t)c0 * m, spitch, w, rows, cudaMemcpyDeviceToDevice, st));
            PA_CUDA(cudaMemcpy2DAsync(fy, w, all_fx + all + (size_t)c0 * m, spitch, w, rows, cudaMemcpyDeviceToDevice, st));
            PA_CUDA(cudaMemcpy2DAsync(coef, 2 * w, all_fx + 2 * all + (size_t)c0 * m * 2, 2 * spitch, 2 * w, rows, cudaMemcpyDeviceToDevice, st));
        } else if (S > 0 && !coef_host) {
            rc = pa_rng_spectrum(c, seed, realization0 + (unsigned long long)c0, b, 0, S, m, edges, psd, fx, fy, coef, st);
            if (rc) return rc;
        }
        rc = simulate_chunk(c, p, b, fx, fy, coef, pupils_dev, npupil, table_dev + (size_t)c0 * out_stride, out_stride, st);
        if (rc) return rc;
    }
    return PA_OK;
}

// enqueue only: host -> device copies, the whole batch, the table's device -> host copy; nothing waits
static int simulate_host_enqueue(pa_ctx* c, const pa_path* p, int batch, const float* fx_host, const float* fy_host, const float* coef_host,
                                 unsigned long long seed, unsigned long long realization0, const float* edges, const float* psd,
                                 const float* pupils_host, int npupil, double* out_host, int out_stride, cudaStream_t st) {
    PA_REQUIRE(c && p && out_host && batch > 0, "bad arguments to pa_simulate_batch");
    PA_REQUIRE(coef_host || (edges && psd) || p->n_screens == 0, "either host coefficients or ring tables are required");
    int rc = grow((void**)&c->table, &c->table_bytes, (size_t)batch * out_stride * sizeof(double));
    if (rc) return rc;
    rc = grow((void**)&c->pupils, &c->pupils_bytes, (size_t)(npupil > 0 ? npupil : 1) * 3 * sizeof(float));
    if (rc) return rc;
    if (npupil > 0) PA_CUDA(cudaMemcpyAsync(c->pupils, pupils_host, (size_t)npupil * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = simulate_common(c, p, batch, fx_host, fy_host, coef_host, seed, realization0, edges, psd, c->pupils, npupil, c->table,
                         out_stride, st);
    if (rc) return rc;
    PA_CUDA(cudaMemcpyAsync(out_host, c->table, (size_t)batch * out_stride * sizeof(double), cudaMemcpyDeviceToHost, st));
    return PA_OK;
}

int pa_simulate_batch(pa_ctx* c, const pa_path* p, int batch, const float* fx_host, const float* fy_host, const float* coef_host,
                      unsigned long long seed, unsigned long long realization0, const float* edges, const float* psd,
                      const float* pupils_host, int npupil, double* out_host, int out_stride, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = simulate_host_enqueue(c, p, batch, fx_host, fy_host, coef_host, seed, realization0, edges, psd, pupils_host, npupil, out_host,
                                   out_stride, st);
    if (rc) return rc;
    PA_CUDA(cudaStreamSynchronize(st));
    return PA_OK;
}

int pa_simulate_batch_async(pa_ctx* c, const pa_path* p, int batch, const float* fx_host, const float* fy_host, const float* coef_host,
                            unsigned long long seed, unsigned long long realization0, const float* edges, const float* psd,
                            const float* pupils_host, int npupil, double* out_host, int out_stride, void* stream) {
    return simulate_host_enqueue(c, p, batch, fx_host, fy_host, coef_host, seed, realization0, edges, psd, pupils_host, npupil, out_host,
                                 out_stride, (cudaStream_t)stream);
}

int pa_stream_synchronize(pa_ctx* c, void* stream) {
    PA_REQUIRE(c != nullptr, "bad arguments to pa_stream_synchronize: null context");
    PA_CUDA(cudaSetDevice(c->device));
    PA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PA_OK;
}

int pa_simulate_batch_device(pa_ctx* c, const pa_path* p, int batch, unsigned long long seed, unsigned long long realization0,
                             const float* edges, const float* psd, const float* pupils_dev, int npupil, double* table_dev,
                             int out_stride, void* stream) {
    PA_REQUIRE(c && p && table_dev && batch > 0, "bad arguments to pa_simulate_batch_device");
    PA_REQUIRE((edges && psd) || p->n_screens == 0, "ring tables are required");
    return simulate_common(c, p, batch, nullptr, nullptr, nullptr, seed, realization0, edges, psd, pupils_dev, npupil, table_dev,
                           out_stride, (cudaStream_t)stream);
}

}  // extern "C"
