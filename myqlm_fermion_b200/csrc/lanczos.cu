// Kernel (3): Lanczos ground-state driver on the matrix-free apply, plus the fused vector kernels
// it needs (dot, three-term recurrence update + norm, scale, axpy) and state generators.
//
// Replaces np.linalg.eigvalsh(H.get_matrix()) / scipy eigsh(H.get_matrix(sparse=True), which="SA")
// at the reference's call sites (/root/reference/tests/integration/test_integration_models.py:102-103,
// tests/integration/test_integration_adaptvqe.py:97,142).
//
// All vector kernels are HBM-streaming: grid = multiple of the SM count, 128-bit loads/stores,
// fp64 accumulation, fixed-order two-stage reductions.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace qfe {

namespace {

constexpr int VT = 256;

template <typename real_t> struct C2;
template <> struct C2<double> { using type = double2; };
template <> struct C2<float> { using type = float2; };

__device__ __forceinline__ double2 block_sum(double2 v) {
    __shared__ double2 s_red[VT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, d);
        v.y += __shfl_down_sync(0xffffffffu, v.y, d);
    }
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double2 r = make_double2(0.0, 0.0);
    if (threadIdx.x == 0)
        for (int w = 0; w < VT / 32; ++w) {
            r.x += s_red[w].x;
            r.y += s_red[w].y;
        }
    return r;
}

// <a|b> = sum conj(a) b
template <typename real_t>
__global__ void __launch_bounds__(VT) dot_kernel(const typename C2<real_t>::type* __restrict__ a, const typename C2<real_t>::type* __restrict__ b,
                                                  int64_t n, double2* __restrict__ partials) {
    double re = 0.0, im = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
        const auto u = a[i];
        const auto v = b[i];
        re += (double)u.x * (double)v.x + (double)u.y * (double)v.y;
        im += (double)u.x * (double)v.y - (double)u.y * (double)v.x;
    }
    double2 tot = block_sum(make_double2(re, im));
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// w <- w - alpha v - beta vp ; partial ||w||^2   (alpha, beta real: Hermitian Lanczos)
template <typename real_t>
__global__ void __launch_bounds__(VT) lanczos_update_kernel(typename C2<real_t>::type* __restrict__ w, const typename C2<real_t>::type* __restrict__ v,
                                                             const typename C2<real_t>::type* __restrict__ vp, double alpha, double beta, int64_t n,
                                                             double2* __restrict__ partials) {
    double nrm = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
        auto x = w[i];
        const auto a = v[i];
        double xr = (double)x.x - alpha * (double)a.x;
        double xi = (double)x.y - alpha * (double)a.y;
        if (vp) {
            const auto b = vp[i];
            xr -= beta * (double)b.x;
            xi -= beta * (double)b.y;
        }
        x.x = (real_t)xr;
        x.y = (real_t)xi;
        w[i] = x;
        nrm += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
    }
    double2 tot = block_sum(make_double2(nrm, 0.0));
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

// y <- alpha x + (accumulate ? y : 0)   (complex alpha)
template <typename real_t>
__global__ void __launch_bounds__(VT) axpy_kernel(double ar, double ai, const typename C2<real_t>::type* __restrict__ x,
                                                   typename C2<real_t>::type* __restrict__ y, int64_t n, int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
        const auto u = x[i];
        double re = ar * (double)u.x - ai * (double)u.y;
        double im = ar * (double)u.y + ai * (double)u.x;
        typename C2<real_t>::type o;
        if (accumulate) {
            o = y[i];
            re += (double)o.x;
            im += (double)o.y;
        }
        o.x = (real_t)re;
        o.y = (real_t)im;
        y[i] = o;
    }
}

template <typename real_t>
__global__ void __launch_bounds__(VT) scale_kernel(double ar, double ai, typename C2<real_t>::type* __restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
        auto u = x[i];
        const double re = ar * (double)u.x - ai * (double)u.y;
        const double im = ar * (double)u.y + ai * (double)u.x;
        u.x = (real_t)re;
        u.y = (real_t)im;
        x[i] = u;
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t v) {  // splitmix64 finaliser
    v += 0x9E3779B97F4A7C15ull;
    v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
    v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
    return v ^ (v >> 31);
}

// counter-based N(0,1) + i N(0,1): amplitude i depends only on (seed, offset + i)
template <typename real_t>
__global__ void __launch_bounds__(VT) fill_random_kernel(typename C2<real_t>::type* __restrict__ x, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
        const uint64_t h = mix64(mix64(seed) ^ (offset + (uint64_t)i));
        const uint64_t h2 = mix64(h);
        const double u1 = ((double)(h >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
        const double u2 = (double)(h2 >> 11) * (1.0 / 9007199254740992.0);
        const double r = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        typename C2<real_t>::type o;
        o.x = (real_t)(r * c);
        o.y = (real_t)(r * s);
        x[i] = o;
    }
}

// product state  (x)_q (alpha_q |0> + beta_q |1>), qubit q <-> bit n-1-q of the full index.  The factor of the low PS_LOW_BITS qubits is
// a shared-memory table built once per CTA; a CTA writes blocks of 2^PS_LOW_BITS consecutive amplitudes = (factor of the high qubits,
// one product chain per thread and block) x table: about three complex multiplications per amplitude instead of n, so the
// kernel runs at the HBM write rate.
constexpr int PS_LOW_BITS = 11;
struct ProductFactors {
    double ab[4 * 64];  // (Re alpha_q, Im alpha_q, Re beta_q, Im beta_q) per qubit: travels as the kernel parameter
};
template <typename real_t>
__global__ void __launch_bounds__(VT) product_state_kernel(typename C2<real_t>::type* __restrict__ x, int n_local, uint64_t shard, int n,
                                                            const __grid_constant__ ProductFactors f) {
    __shared__ double s_ab[4 * 64];
    __shared__ double2 s_lo[1 << PS_LOW_BITS];
    for (int i = threadIdx.x; i < 4 * n; i += VT) s_ab[i] = f.ab[i];
    __syncthreads();
    const int L = n_local < PS_LOW_BITS ? n_local : PS_LOW_BITS;
    const int n_lo = 1 << L;
    for (int j = threadIdx.x; j < n_lo; j += VT) {
        double re = 1.0, im = 0.0;
        for (int q = n - L; q < n; ++q) {
            const int bit = (j >> (n - 1 - q)) & 1;
            const double fr = s_ab[4 * q + 2 * bit], fi = s_ab[4 * q + 2 * bit + 1];
            const double t = re * fr - im * fi;
            im = re * fi + im * fr;
            re = t;
        }
        s_lo[j] = make_double2(re, im);
    }
    __syncthreads();
    const int64_t n_blocks = (int64_t(1) << n_local) >> L;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint64_t b = (shard << n_local) | ((uint64_t)blk << L);
        double re = 1.0, im = 0.0;
        for (int q = 0; q < n - L; ++q) {
            const int bit = (int)((b >> (n - 1 - q)) & 1ull);
            const double fr = s_ab[4 * q + 2 * bit], fi = s_ab[4 * q + 2 * bit + 1];
            const double t = re * fr - im * fi;
            im = re * fi + im * fr;
            re = t;
        }
        typename C2<real_t>::type* __restrict__ out = x + (blk << L);
        for (int j = threadIdx.x; j < n_lo; j += VT) {
            const double2 f = s_lo[j];
            typename C2<real_t>::type o;
            o.x = (real_t)(re * f.x - im * f.y);
            o.y = (real_t)(re * f.y + im * f.x);
            out[j] = o;
        }
    }
}

int vec_grid(qfe_ctx* ctx, int64_t n) { return grid_for(n, VT, ctx->sm_count, 16); }

template <typename real_t>
int dot_impl(qfe_ctx* ctx, const void* a, const void* b, int64_t n, double out[2]) {
    using c2 = typename C2<real_t>::type;
    const int g = vec_grid(ctx, n);
    QFE_TRY(ctx->partials.reserve(sizeof(double2) * g));
    QFE_TRY(ctx->result.reserve(sizeof(double2)));
    dot_kernel<real_t><<<g, VT, 0, ctx->stream>>>((const c2*)a, (const c2*)b, n, ctx->partials.as<double2>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_TRY(reduce_to_vals(ctx, ctx->partials.as<double2>(), 1, g, ctx->result.as<double2>()));
    QFE_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->result.p, sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    out[0] = ((double*)ctx->pinned)[0];
    out[1] = ((double*)ctx->pinned)[1];
    return QFE_OK;
}

template <typename real_t>
int update_impl(qfe_ctx* ctx, void* w, const void* v, const void* vp, double alpha, double beta, int64_t n, double* norm2) {
    using c2 = typename C2<real_t>::type;
    const int g = vec_grid(ctx, n);
    QFE_TRY(ctx->partials.reserve(sizeof(double2) * g));
    QFE_TRY(ctx->result.reserve(sizeof(double2)));
    lanczos_update_kernel<real_t><<<g, VT, 0, ctx->stream>>>((c2*)w, (const c2*)v, (const c2*)vp, alpha, beta, n, ctx->partials.as<double2>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_TRY(reduce_to_vals(ctx, ctx->partials.as<double2>(), 1, g, ctx->result.as<double2>()));
    QFE_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->result.p, sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    *norm2 = ((double*)ctx->pinned)[0];
    return QFE_OK;
}

template <typename real_t>
int axpy_impl(qfe_ctx* ctx, double ar, double ai, const void* x, void* y, int64_t n, int accumulate) {
    using c2 = typename C2<real_t>::type;
    axpy_kernel<real_t><<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(ar, ai, (const c2*)x, (c2*)y, n, accumulate);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

template <typename real_t>
int scale_impl(qfe_ctx* ctx, double ar, double ai, void* x, int64_t n) {
    using c2 = typename C2<real_t>::type;
    scale_kernel<real_t><<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(ar, ai, (c2*)x, n);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

#define DISPATCH(dtype, fn, ...) ((dtype) == QFE_C128 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

// ---- symmetric tridiagonal: lowest eigenpair by Sturm bisection + inverse iteration ---------
int sturm_count(const std::vector<double>& a, const std::vector<double>& b, int m, double x) {
    int cnt = 0;
    double d = 1.0;
    for (int i = 0; i < m; ++i) {
        const double off = i ? b[i - 1] * b[i - 1] : 0.0;
        d = (a[i] - x) - (i ? off / d : 0.0);
        if (d == 0.0) d = 1e-300;
        if (d < 0.0) ++cnt;
    }
    return cnt;
}

double lowest_eigenvalue(const std::vector<double>& a, const std::vector<double>& b, int m) {
    double lo = a[0], hi = a[0];
    for (int i = 0; i < m; ++i) {
        const double r = (i ? fabs(b[i - 1]) : 0.0) + (i + 1 < m ? fabs(b[i]) : 0.0);
        lo = std::min(lo, a[i] - r);
        hi = std::max(hi, a[i] + r);
    }
    for (int it = 0; it < 200 && hi - lo > 4e-16 * std::max(fabs(lo), fabs(hi)); ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        if (sturm_count(a, b, m, mid) >= 1)
            hi = mid;
        else
            lo = mid;
    }
    return 0.5 * (lo + hi);
}

// eigenvector of the tridiagonal (a, b) for eigenvalue theta by inverse iteration with a
// partially pivoted tridiagonal solve
void lowest_eigenvector(const std::vector<double>& a, const std::vector<double>& b, int m, double theta, std::vector<double>& y) {
    y.assign(m, 1.0 / sqrt((double)m));
    if (m == 1) {
        y[0] = 1.0;
        return;
    }
    const double scale = std::max(1.0, fabs(theta));
    const double shift = theta - 1e-13 * scale;
    std::vector<double> dl(m), d(m), du(m), du2(m);
    for (int iter = 0; iter < 4; ++iter) {
        for (int i = 0; i < m; ++i) {
            d[i] = a[i] - shift;
            du[i] = i + 1 < m ? b[i] : 0.0;
            dl[i] = i + 1 < m ? b[i] : 0.0;
            du2[i] = 0.0;
        }
        for (int i = 0; i + 1 < m; ++i) {
            if (fabs(d[i]) >= fabs(dl[i])) {
                if (d[i] == 0.0) d[i] = 1e-300;
                const double f = dl[i] / d[i];
                d[i + 1] -= f * du[i];
                y[i + 1] -= f * y[i];
            } else {
                const double f = d[i] / dl[i];
                d[i] = dl[i];
                const double tmp = d[i + 1];
                d[i + 1] = du[i] - f * tmp;
                if (i + 2 < m) {
                    du2[i] = du[i + 1];
                    du[i + 1] = -f * du2[i];
                }
                du[i] = tmp;
                const double t = y[i];
                y[i] = y[i + 1];
                y[i + 1] = t - f * y[i];
            }
        }
        if (d[m - 1] == 0.0) d[m - 1] = 1e-300;
        y[m - 1] /= d[m - 1];
        y[m - 2] = (y[m - 2] - du[m - 2] * y[m - 1]) / d[m - 2];
        for (int i = m - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / d[i];
        double nrm = 0.0;
        for (int i = 0; i < m; ++i) nrm += y[i] * y[i];
        nrm = sqrt(nrm);
        for (int i = 0; i < m; ++i) y[i] /= nrm;
    }
}

}  // namespace

int lanczos_run(qfe_ctx* ctx, const LanczosOps& ops, int64_t N, void* v, void* work, int max_iter, double tol, int dtype, int want_vector, double* eval,
                int* iters, double* resid) {
    QFE_REQUIRE(max_iter >= 1, QFE_ERR_ARG, "qfe_lanczos: max_iter must be >= 1");
    const size_t bytes = (size_t)N * (dtype == QFE_C128 ? 16 : 8);
    char* wk = (char*)work;
    void* buf[3] = {wk, wk + bytes, wk + 2 * bytes};

    std::vector<double> alpha, beta;
    double theta = 0.0, res = 0.0;
    int m_done = 0;
    std::vector<double> y;

    for (int pass = 0; pass < (want_vector ? 2 : 1); ++pass) {
        // v_cur = start / ||start||
        double nn[2];
        QFE_TRY(qfe_vec_dot(ctx, v, v, N, dtype, nn));
        QFE_TRY(ops.reduce(nn, 2));
        QFE_REQUIRE(nn[0] > 0.0, QFE_ERR_ARG, "qfe_lanczos: start vector has zero norm");
        void *cur = buf[0], *prev = buf[1], *w = buf[2];
        QFE_TRY(DISPATCH(dtype, axpy_impl, ctx, 1.0 / sqrt(nn[0]), 0.0, v, cur, N, 0));
        if (pass == 1) QFE_TRY(DISPATCH(dtype, axpy_impl, ctx, y[0], 0.0, cur, v, N, 0));
        const int m_target = pass == 0 ? max_iter : m_done;
        double beta_prev = 0.0;
        for (int j = 0; j < m_target; ++j) {
            QFE_TRY(ops.apply(cur, w));
            double a_j;
            if (pass == 0) {
                double d[2];
                QFE_TRY(qfe_vec_dot(ctx, cur, w, N, dtype, d));
                QFE_TRY(ops.reduce(d, 2));
                a_j = d[0];
                alpha.push_back(a_j);
            } else {
                a_j = alpha[j];
            }
            double nrm2 = 0.0;
            QFE_TRY(DISPATCH(dtype, update_impl, ctx, w, cur, j ? prev : nullptr, a_j, beta_prev, N, &nrm2));
            QFE_TRY(ops.reduce(&nrm2, 1));
            double b_j = sqrt(nrm2);
            if (pass == 0) {
                beta.push_back(b_j);
                const int m = j + 1;
                const bool check = (m % 5 == 0) || m == max_iter || b_j <= 1e-14 * std::max(1.0, fabs(a_j));
                if (check) {
                    theta = lowest_eigenvalue(alpha, beta, m);
                    lowest_eigenvector(alpha, beta, m, theta, y);
                    res = fabs(b_j * y[m - 1]);
                    m_done = m;
                    if (res <= tol * std::max(1.0, fabs(theta)) || b_j <= 1e-14 * std::max(1.0, fabs(a_j))) break;
                }
                m_done = m;
            } else {
                b_j = beta[j];
                if (j + 1 >= m_done) break;
            }
            if (b_j == 0.0) break;
            // v_next = w / beta_j ; rotate buffers
            QFE_TRY(DISPATCH(dtype, scale_impl, ctx, 1.0 / b_j, 0.0, w, N));
            void* t = prev;
            prev = cur;
            cur = w;
            w = t;
            beta_prev = b_j;
            if (pass == 1) QFE_TRY(DISPATCH(dtype, axpy_impl, ctx, y[j + 1], 0.0, cur, v, N, 1));
        }
        if (pass == 0) {
            theta = lowest_eigenvalue(alpha, beta, m_done);
            lowest_eigenvector(alpha, beta, m_done, theta, y);
            res = fabs(beta[m_done - 1] * y[m_done - 1]);
        }
    }
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    *eval = theta;
    if (iters) *iters = m_done;
    if (resid) *resid = res;
    return QFE_OK;
}

}  // namespace qfe

using namespace qfe;

extern "C" {

int qfe_vec_dot(qfe_ctx* ctx, const void* a, const void* b, int64_t n, int dtype, double out[2]) {
    QFE_REQUIRE(ctx && a && b && out && n >= 0, QFE_ERR_ARG, "qfe_vec_dot: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    return DISPATCH(dtype, dot_impl, ctx, a, b, n, out);
}

int qfe_vec_axpy(qfe_ctx* ctx, const double alpha[2], const void* x, void* y, int64_t n, int dtype) {
    QFE_REQUIRE(ctx && alpha && x && y && n >= 0, QFE_ERR_ARG, "qfe_vec_axpy: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    return DISPATCH(dtype, axpy_impl, ctx, alpha[0], alpha[1], x, y, n, 1);
}

int qfe_vec_scale(qfe_ctx* ctx, const double alpha[2], void* x, int64_t n, int dtype) {
    QFE_REQUIRE(ctx && alpha && x && n >= 0, QFE_ERR_ARG, "qfe_vec_scale: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    return DISPATCH(dtype, scale_impl, ctx, alpha[0], alpha[1], x, n);
}

int qfe_vec_fill_random(qfe_ctx* ctx, void* x, int64_t n, int dtype, uint64_t seed, uint64_t offset) {
    QFE_REQUIRE(ctx && x && n >= 0, QFE_ERR_ARG, "qfe_vec_fill_random: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    if (dtype == QFE_C128)
        fill_random_kernel<double><<<vec_grid(ctx, n), VT, 0, ctx->stream>>>((double2*)x, n, seed, offset);
    else
        fill_random_kernel<float><<<vec_grid(ctx, n), VT, 0, ctx->stream>>>((float2*)x, n, seed, offset);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

int qfe_vec_product_state(qfe_ctx* ctx, void* x, int n_local, int shard, int n, const double* alpha_beta, int dtype) {
    QFE_REQUIRE(ctx && x && alpha_beta && n >= 1 && n <= 64 && n_local >= 0 && n_local <= n, QFE_ERR_ARG, "qfe_vec_product_state: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    ProductFactors f;
    memset(&f, 0, sizeof(f));
    memcpy(f.ab, alpha_beta, sizeof(double) * 4 * n);
    const int64_t n_amps = int64_t(1) << n_local;
    if (dtype == QFE_C128)
        product_state_kernel<double><<<vec_grid(ctx, n_amps), VT, 0, ctx->stream>>>((double2*)x, n_local, (uint64_t)shard, n, f);
    else
        product_state_kernel<float><<<vec_grid(ctx, n_amps), VT, 0, ctx->stream>>>((float2*)x, n_local, (uint64_t)shard, n, f);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return QFE_OK;
}

int qfe_lanczos(qfe_ctx* ctx, const qfe_plan* p, void* v, void* work, int max_iter, double tol, int dtype, int want_vector, double* eval,
                int* iters, double* resid) {
    QFE_REQUIRE(ctx && p && v && work && eval, QFE_ERR_ARG, "qfe_lanczos: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    int n = 0, n_local = 0;
    QFE_TRY(qfe_plan_geometry(p, &n, &n_local));
    QFE_REQUIRE(n == n_local, QFE_ERR_STATE, "qfe_lanczos needs an unsharded plan; use qfe_lanczos_sharded");
    LanczosOps ops;
    ops.apply = [&](const void* cur, void* w) { return qfe_apply(ctx, p, cur, w, dtype); };
    ops.reduce = [](double*, int) { return (int)QFE_OK; };
    return lanczos_run(ctx, ops, int64_t(1) << n, v, work, max_iter, tol, dtype, want_vector, eval, iters, resid);
}

}  // extern "C"
