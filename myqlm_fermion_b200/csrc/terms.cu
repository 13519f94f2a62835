// Kernel (1): fermionic terms -> canonical packed Pauli terms, on the GPU.
//
// Replaces the Python object algebra of transform_to_{jw,bk,parity}_basis
// (/root/reference/qat/fermion/transforms.py:82-257) and the integral enumeration of
// ElectronicStructureHamiltonian._get_fermionic_terms (/root/reference/qat/fermion/hamiltonians.py:625-648).
//
// Pipeline (all integer work except the coefficient sums):
//   integrals --flag/scan/emit--> fermionic term records --scan(2^len)--> raw products
//   expand: one thread per (term, A/B choice): XOR masks, phase by popc        (coalesced stores)
//   stable radix sort on (slot, x, z) -> run heads -> per-run ordered sum -> drop / fold identity
// The summation order inside a key is the raw-product order (stable sort), so results are
// bit-reproducible run to run.
#include "common.cuh"

namespace qfe {

thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

namespace {

constexpr int MAX_LETTERS = 12;

// one fermionic term: up to 12 ladder operators
struct FTerm {
    double w_re, w_im;
    uint8_t mode[MAX_LETTERS];
    uint16_t dagger;  // bit i set: letter i is a creation operator
    uint16_t len;
};
static_assert(sizeof(FTerm) == 32, "FTerm layout");

// ladder strings of one encoding: a_j = 1/2 A_j + i/2 B_j ; a_j^+ = 1/2 A_j - i/2 B_j
// A_j = X^{ax} Z^{az}, B_j = i X^{bx} Z^{bz} (exactly one Y, on qubit j)
struct Ladder {
    uint64_t ax[64], az[64], bx[64], bz[64];
};

inline uint64_t qbit(int n, int q) { return uint64_t(1) << (n - 1 - q); }

// Fenwick tree of /root/reference/qat/fermion/transforms.py:19-32 as a parent array
void fenwick_parents(int lo, int hi, std::vector<int>& parent) {
    if (lo == hi) return;
    int mid = (lo + hi) / 2;
    parent[mid] = hi;
    fenwick_parents(lo, mid, parent);
    fenwick_parents(mid + 1, hi, parent);
}

int build_ladder(int n, int encoding, Ladder& L) {
    memset(&L, 0, sizeof(L));
    if (encoding == QFE_ENC_JW) {  // transforms.py:116-126
        for (int j = 0; j < n; ++j) {
            uint64_t zs = 0;
            for (int q = 0; q < j; ++q) zs |= qbit(n, q);
            L.ax[j] = qbit(n, j);
            L.az[j] = zs;
            L.bx[j] = qbit(n, j);
            L.bz[j] = zs | qbit(n, j);
        }
    } else if (encoding == QFE_ENC_PARITY) {  // transforms.py:169-179
        for (int j = 0; j < n; ++j) {
            uint64_t xs = 0;
            for (int q = j; q < n; ++q) xs |= qbit(n, q);
            L.ax[j] = xs;
            L.az[j] = j > 0 ? qbit(n, j - 1) : 0;
            L.bx[j] = xs;
            L.bz[j] = qbit(n, j);
        }
    } else if (encoding == QFE_ENC_BK) {  // transforms.py:35-79, 227-251
        std::vector<int> parent(n, -1);
        if (n > 0) fenwick_parents(0, n - 1, parent);
        for (int j = 0; j < n; ++j) {
            uint64_t u = 0, c = 0, f = 0;
            for (int a = parent[j]; a >= 0; a = parent[a]) {
                u |= qbit(n, a);
                for (int k = 0; k < j; ++k)
                    if (parent[k] == a) c |= qbit(n, k);  // children of an ancestor, index < j
            }
            for (int k = 0; k < n; ++k)
                if (parent[k] == j) f |= qbit(n, k);
            L.ax[j] = qbit(n, j) | u;
            L.az[j] = c | f;  // P = C u F
            L.bx[j] = qbit(n, j) | u;
            L.bz[j] = c | qbit(n, j);
        }
    } else {
        set_error("unknown encoding %d", encoding);
        return QFE_ERR_UNSUPPORTED;
    }
    return QFE_OK;
}

// ---- integrals -> term records ---------------------------------------------------------------
__global__ void integral_flags(const double2* __restrict__ hpq, const double2* __restrict__ hpqrs, int64_t n2, int64_t n4, double tol,
                               uint32_t* __restrict__ flags) {
    const int64_t total = n2 + n4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        double2 h = e < n2 ? hpq[e] : hpqrs[e - n2];
        flags[e] = hypot(h.x, h.y) > tol ? 1u : 0u;
    }
}

__global__ void integral_emit(const double2* __restrict__ hpq, const double2* __restrict__ hpqrs, int n, int64_t n2, int64_t n4,
                              const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, FTerm* __restrict__ terms) {
    const int64_t total = n2 + n4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (!flags[e]) continue;
        FTerm t;
        memset(&t, 0, sizeof(t));
        if (e < n2) {  // h_pq a+_p a_q
            double2 h = hpq[e];
            t.w_re = h.x;
            t.w_im = h.y;
            t.mode[0] = (uint8_t)(e / n);
            t.mode[1] = (uint8_t)(e % n);
            t.dagger = 0x1;
            t.len = 2;
        } else {  // 1/2 h_pqrs a+_p a+_q a_r a_s
            int64_t r = e - n2;
            double2 h = hpqrs[r];
            t.w_re = 0.5 * h.x;
            t.w_im = 0.5 * h.y;
            t.mode[3] = (uint8_t)(r % n);
            r /= n;
            t.mode[2] = (uint8_t)(r % n);
            r /= n;
            t.mode[1] = (uint8_t)(r % n);
            t.mode[0] = (uint8_t)(r / n);
            t.dagger = 0x3;
            t.len = 4;
        }
        terms[pos[e]] = t;
    }
}

__global__ void term_product_counts(const FTerm* __restrict__ terms, int64_t nt, uint32_t* __restrict__ counts) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += (int64_t)gridDim.x * blockDim.x)
        counts[t] = 1u << terms[t].len;
}

// ---- expand: one thread per raw product -------------------------------------------------------
__global__ void __launch_bounds__(256) expand_products(const FTerm* __restrict__ terms, int64_t nt, const uint32_t* __restrict__ raw_off,
                                                       int64_t n_raw, const Ladder* __restrict__ ladder_g, uint32_t slot,
                                                       uint64_t* __restrict__ rx, uint64_t* __restrict__ rz, double2* __restrict__ rc,
                                                       uint32_t* __restrict__ rowner, uint32_t* __restrict__ ridx) {
    __shared__ Ladder lad;
    {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(ladder_g);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&lad);
        for (int i = threadIdx.x; i < (int)(sizeof(Ladder) / 8); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_raw; i += (int64_t)gridDim.x * blockDim.x) {
        // term owning raw product i: last t with raw_off[t] <= i
        int64_t lo = 0, hi = nt - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) >> 1;
            if ((int64_t)raw_off[mid] <= i)
                lo = mid;
            else
                hi = mid - 1;
        }
        const FTerm t = terms[lo];
        const uint32_t choice = (uint32_t)(i - raw_off[lo]);
        uint64_t x = 0, z = 0;
        uint32_t k = 0;  // power of i
        for (int l = 0; l < t.len; ++l) {
            const int q = t.mode[l];
            uint64_t sx, sz;
            if ((choice >> l) & 1u) {
                sx = lad.bx[q];
                sz = lad.bz[q];
                // B carries i (its Y); the ladder operator adds -i (creation) or +i (annihilation)
                k += ((t.dagger >> l) & 1u) ? 0u : 2u;
            } else {
                sx = lad.ax[q];
                sz = lad.az[q];
            }
            k += 2u * (uint32_t)__popcll(z & sx);  // (X^x Z^z)(X^sx Z^sz) = (-1)^{|z&sx|} X^{x^sx} Z^{z^sz}
            x ^= sx;
            z ^= sz;
        }
        const double scale = ldexp(1.0, -(int)t.len);
        double re = t.w_re * scale, im = t.w_im * scale;
        switch (k & 3u) {
            case 1: { double tmp = re; re = -im; im = tmp; break; }
            case 2: re = -re; im = -im; break;
            case 3: { double tmp = re; re = im; im = -tmp; break; }
            default: break;
        }
        rx[i] = x;
        rz[i] = z;
        rc[i] = make_double2(re, im);
        rowner[i] = slot;
        ridx[i] = (uint32_t)i;
    }
}

// ---- canonicalise sorted records ---------------------------------------------------------------
__global__ void run_heads(const uint64_t* __restrict__ x, const uint64_t* __restrict__ z, const uint32_t* __restrict__ owner, int64_t n,
                          uint32_t* __restrict__ head) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        head[i] = (i == 0 || x[i] != x[i - 1] || z[i] != z[i - 1] || owner[i] != owner[i - 1]) ? 1u : 0u;
}

__global__ void run_starts(const uint32_t* __restrict__ head, const uint32_t* __restrict__ run_id, int64_t n, uint32_t* __restrict__ start) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (head[i]) start[run_id[i]] = (uint32_t)i;
}

// one thread per run: ordered sum of its coefficients; decide keep / identity
__global__ void run_reduce(const uint64_t* __restrict__ x, const uint64_t* __restrict__ z, const uint32_t* __restrict__ owner,
                           const uint32_t* __restrict__ idx, const double2* __restrict__ rc, const uint32_t* __restrict__ start, int64_t n_runs,
                           int64_t n, double drop_eps, uint64_t* __restrict__ ox, uint64_t* __restrict__ oz, uint32_t* __restrict__ oo,
                           double2* __restrict__ oc, uint32_t* __restrict__ keep, double2* __restrict__ constants) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = start[r], e = (r + 1 < n_runs) ? (int64_t)start[r + 1] : n;
        double re = 0.0, im = 0.0;
        for (int64_t j = b; j < e; ++j) {
            double2 c = rc[idx[j]];
            re += c.x;
            im += c.y;
        }
        const uint64_t xr = x[b], zr = z[b];
        const uint32_t o = owner[b];
        ox[r] = xr;
        oz[r] = zr;
        oo[r] = o;
        // Y-form (Term.coeff): c_Y = c_XZ * (-i)^{popc(x&z)}
        switch (__popcll(xr & zr) & 3) {
            case 1: { double t = re; re = im; im = -t; break; }
            case 2: re = -re; im = -im; break;
            case 3: { double t = re; re = -im; im = t; break; }
            default: break;
        }
        oc[r] = make_double2(re, im);
        if (xr == 0 && zr == 0) {
            constants[o] = make_double2(constants[o].x + re, constants[o].y + im);  // a single run per slot holds the identity
            keep[r] = 0u;
        } else {
            keep[r] = hypot(re, im) > drop_eps ? 1u : 0u;
        }
    }
}

__global__ void compact_runs(const uint64_t* __restrict__ ox, const uint64_t* __restrict__ oz, const uint32_t* __restrict__ oo,
                             const double2* __restrict__ oc, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, int64_t n_runs,
                             uint64_t* __restrict__ fx, uint64_t* __restrict__ fz, uint32_t* __restrict__ fo, double2* __restrict__ fc) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        if (!keep[r]) continue;
        const uint32_t p = pos[r];
        fx[p] = ox[r];
        fz[p] = oz[r];
        fo[p] = oo[r];
        fc[p] = oc[r];
    }
}

int count_xmasks(qfe_terms* t) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < t->T; ++i)
        if (i == 0 || t->x[i] != t->x[i - 1] || t->owner[i] != t->owner[i - 1]) ++cnt;
    t->n_xmasks = cnt;
    return QFE_OK;
}

// raw records on device (rx, rz, rc XZ-form, rowner) -> canonical host term set
int canonicalise(qfe_ctx* ctx, int n, int64_t n_raw, DevBuf& rx, DevBuf& rz, DevBuf& rc, DevBuf& rowner, DevBuf& ridx, int64_t n_slots,
                 const double* constant_in, double drop_eps, qfe_terms** out) {
    qfe_terms* t = new (std::nothrow) qfe_terms();
    QFE_REQUIRE(t, QFE_ERR_OOM, "out of host memory");
    t->n = n;
    t->n_slots = n_slots;
    t->constant.assign(2 * n_slots, 0.0);
    if (constant_in)
        for (int64_t i = 0; i < 2 * n_slots; ++i) t->constant[i] = constant_in[i];
    *out = t;
    if (n_raw == 0) return QFE_OK;

    cudaStream_t s = ctx->stream;
    const int B = 256;
    const int G = grid_for(n_raw, B, ctx->sm_count);
    DevBuf tmp, scan_scratch, head, runid, d_total;
    int owner_bits = 0;
    while ((int64_t(1) << owner_bits) < n_slots) ++owner_bits;
    int rc_ = radix_sort_terms(ctx, rx.as<uint64_t>(), rz.as<uint64_t>(), rowner.as<uint32_t>(), ridx.as<uint32_t>(), n_raw, n, n, owner_bits, tmp);
    if (rc_ != QFE_OK) return rc_;
    tmp.release();

    QFE_TRY(head.alloc(sizeof(uint32_t) * n_raw));
    QFE_TRY(runid.alloc(sizeof(uint32_t) * n_raw));
    QFE_TRY(d_total.alloc(sizeof(uint32_t) * 4));
    run_heads<<<G, B, 0, s>>>(rx.as<uint64_t>(), rz.as<uint64_t>(), rowner.as<uint32_t>(), n_raw, head.as<uint32_t>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_TRY(exclusive_scan_u32(ctx, head.as<uint32_t>(), runid.as<uint32_t>(), n_raw, scan_scratch, d_total.as<uint32_t>()));
    uint32_t n_runs32 = 0;
    QFE_CUDA(cudaMemcpyAsync(&n_runs32, d_total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    QFE_CUDA(cudaStreamSynchronize(s));
    const int64_t n_runs = n_runs32;

    DevBuf start, ox, oz, oo, oc, keep, pos, consts;
    QFE_TRY(start.alloc(sizeof(uint32_t) * (n_runs + 1)));
    QFE_TRY(ox.alloc(sizeof(uint64_t) * n_runs));
    QFE_TRY(oz.alloc(sizeof(uint64_t) * n_runs));
    QFE_TRY(oo.alloc(sizeof(uint32_t) * n_runs));
    QFE_TRY(oc.alloc(sizeof(double2) * n_runs));
    QFE_TRY(keep.alloc(sizeof(uint32_t) * n_runs));
    QFE_TRY(pos.alloc(sizeof(uint32_t) * n_runs));
    QFE_TRY(consts.alloc(sizeof(double2) * n_slots));
    QFE_CUDA(cudaMemsetAsync(consts.p, 0, sizeof(double2) * n_slots, s));
    run_starts<<<G, B, 0, s>>>(head.as<uint32_t>(), runid.as<uint32_t>(), n_raw, start.as<uint32_t>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    const int GR = grid_for(n_runs, B, ctx->sm_count);
    run_reduce<<<GR, B, 0, s>>>(rx.as<uint64_t>(), rz.as<uint64_t>(), rowner.as<uint32_t>(), ridx.as<uint32_t>(), rc.as<double2>(),
                                start.as<uint32_t>(), n_runs, n_raw, drop_eps, ox.as<uint64_t>(), oz.as<uint64_t>(), oo.as<uint32_t>(),
                                oc.as<double2>(), keep.as<uint32_t>(), consts.as<double2>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_TRY(exclusive_scan_u32(ctx, keep.as<uint32_t>(), pos.as<uint32_t>(), n_runs, scan_scratch, d_total.as<uint32_t>()));
    uint32_t n_keep32 = 0;
    QFE_CUDA(cudaMemcpyAsync(&n_keep32, d_total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    QFE_CUDA(cudaStreamSynchronize(s));
    const int64_t T = n_keep32;

    DevBuf fx, fz, fo, fc;
    QFE_TRY(fx.alloc(sizeof(uint64_t) * T));
    QFE_TRY(fz.alloc(sizeof(uint64_t) * T));
    QFE_TRY(fo.alloc(sizeof(uint32_t) * T));
    QFE_TRY(fc.alloc(sizeof(double2) * T));
    compact_runs<<<GR, B, 0, s>>>(ox.as<uint64_t>(), oz.as<uint64_t>(), oo.as<uint32_t>(), oc.as<double2>(), keep.as<uint32_t>(),
                                  pos.as<uint32_t>(), n_runs, fx.as<uint64_t>(), fz.as<uint64_t>(), fo.as<uint32_t>(), fc.as<double2>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());

    t->T = T;
    t->x.resize(T);
    t->z.resize(T);
    t->c.resize(2 * T);
    t->owner.resize(T);
    std::vector<double> cadd(2 * n_slots);
    if (T > 0) {
        QFE_CUDA(cudaMemcpyAsync(t->x.data(), fx.p, sizeof(uint64_t) * T, cudaMemcpyDeviceToHost, s));
        QFE_CUDA(cudaMemcpyAsync(t->z.data(), fz.p, sizeof(uint64_t) * T, cudaMemcpyDeviceToHost, s));
        QFE_CUDA(cudaMemcpyAsync(t->c.data(), fc.p, sizeof(double2) * T, cudaMemcpyDeviceToHost, s));
        QFE_CUDA(cudaMemcpyAsync(t->owner.data(), fo.p, sizeof(uint32_t) * T, cudaMemcpyDeviceToHost, s));
    }
    QFE_CUDA(cudaMemcpyAsync(cadd.data(), consts.p, sizeof(double2) * n_slots, cudaMemcpyDeviceToHost, s));
    QFE_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < 2 * n_slots; ++i) t->constant[i] += cadd[i];
    return count_xmasks(t);
}

int expand_and_canonicalise(qfe_ctx* ctx, int n, DevBuf& d_terms, int64_t nt, int encoding, const double constant[2], double drop_eps,
                            qfe_terms** out) {
    cudaStream_t s = ctx->stream;
    Ladder lad;
    QFE_TRY(build_ladder(n, encoding, lad));
    DevBuf d_lad, counts, raw_off, d_total, scan_scratch;
    QFE_TRY(d_lad.alloc(sizeof(Ladder)));
    QFE_CUDA(cudaMemcpyAsync(d_lad.p, &lad, sizeof(Ladder), cudaMemcpyHostToDevice, s));
    int64_t n_raw = 0;
    DevBuf rx, rz, rc, rowner, ridx;
    if (nt > 0) {
        QFE_TRY(counts.alloc(sizeof(uint32_t) * nt));
        QFE_TRY(raw_off.alloc(sizeof(uint32_t) * nt));
        QFE_TRY(d_total.alloc(16));
        term_product_counts<<<grid_for(nt, 256, ctx->sm_count), 256, 0, s>>>(d_terms.as<FTerm>(), nt, counts.as<uint32_t>());
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
        QFE_TRY(exclusive_scan_u32(ctx, counts.as<uint32_t>(), raw_off.as<uint32_t>(), nt, scan_scratch, d_total.as<uint32_t>()));
        uint32_t tot = 0;
        QFE_CUDA(cudaMemcpyAsync(&tot, d_total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        QFE_CUDA(cudaStreamSynchronize(s));
        n_raw = tot;
        QFE_TRY(rx.alloc(sizeof(uint64_t) * n_raw));
        QFE_TRY(rz.alloc(sizeof(uint64_t) * n_raw));
        QFE_TRY(rc.alloc(sizeof(double2) * n_raw));
        QFE_TRY(rowner.alloc(sizeof(uint32_t) * n_raw));
        QFE_TRY(ridx.alloc(sizeof(uint32_t) * n_raw));
        expand_products<<<grid_for(n_raw, 256, ctx->sm_count), 256, 0, s>>>(d_terms.as<FTerm>(), nt, raw_off.as<uint32_t>(), n_raw,
                                                                            d_lad.as<Ladder>(), 0u, rx.as<uint64_t>(), rz.as<uint64_t>(),
                                                                            rc.as<double2>(), rowner.as<uint32_t>(), ridx.as<uint32_t>());
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
    }
    int rc_ = canonicalise(ctx, n, n_raw, rx, rz, rc, rowner, ridx, 1, constant, drop_eps, out);
    QFE_CUDA(cudaStreamSynchronize(s));
    return rc_;
}

}  // namespace
}  // namespace qfe

using namespace qfe;

extern "C" {

const char* qfe_last_error(void) { return qfe::g_last_error.c_str(); }
int qfe_version(void) { return 100; }

int qfe_ctx_create(qfe_ctx** out, int device) {
    QFE_REQUIRE(out, QFE_ERR_ARG, "qfe_ctx_create: null out");
    int count = 0;
    QFE_CUDA(cudaGetDeviceCount(&count));
    QFE_REQUIRE(device >= 0 && device < count, QFE_ERR_ARG, "qfe_ctx_create: device %d out of range (%d visible)", device, count);
    QFE_CUDA(cudaSetDevice(device));
    qfe_ctx* c = new (std::nothrow) qfe_ctx();
    QFE_REQUIRE(c, QFE_ERR_OOM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    QFE_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    QFE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    c->pinned_bytes = 1 << 16;
    QFE_CUDA(cudaMallocHost(&c->pinned, c->pinned_bytes));
    *out = c;
    return QFE_OK;
}

int qfe_ctx_set_stream(qfe_ctx* ctx, void* cuda_stream) {
    QFE_REQUIRE(ctx, QFE_ERR_ARG, "null ctx");
    QFE_CUDA(cudaSetDevice(ctx->device));
    if (ctx->own_stream && ctx->stream) {
        QFE_CUDA(cudaStreamSynchronize(ctx->stream));
        QFE_CUDA(cudaStreamDestroy(ctx->stream));
    }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        QFE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return QFE_OK;
}

int qfe_ctx_sync(qfe_ctx* ctx) {
    QFE_REQUIRE(ctx, QFE_ERR_ARG, "null ctx");
    QFE_CUDA(cudaSetDevice(ctx->device));
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return QFE_OK;
}

int qfe_ctx_destroy(qfe_ctx* ctx) {
    if (!ctx) return QFE_OK;
    cudaSetDevice(ctx->device);
    qfe_comm_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    delete ctx;
    return QFE_OK;
}

int qfe_ctx_launch_count(const qfe_ctx* ctx, int64_t* count) {
    QFE_REQUIRE(ctx && count, QFE_ERR_ARG, "null argument");
    *count = ctx->launches;
    return QFE_OK;
}

int qfe_terms_from_integrals(qfe_ctx* ctx, int n, const double* hpq, const double* hpqrs, const double constant[2], double tol,
                             double drop_eps, int encoding, qfe_terms** out) {
    QFE_REQUIRE(ctx && hpq && out, QFE_ERR_ARG, "qfe_terms_from_integrals: null argument");
    QFE_REQUIRE(n >= 1 && n <= 64, QFE_ERR_UNSUPPORTED, "qfe_terms_from_integrals: nqbits=%d outside 1..64", n);
    QFE_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t n2 = (int64_t)n * n, n4 = hpqrs ? n2 * n2 : 0;
    DevBuf d_hpq, d_hpqrs, flags, pos, d_total, scan_scratch, d_terms;
    QFE_TRY(d_hpq.alloc(sizeof(double2) * n2));
    QFE_CUDA(cudaMemcpyAsync(d_hpq.p, hpq, sizeof(double2) * n2, cudaMemcpyHostToDevice, s));
    if (hpqrs) {
        QFE_TRY(d_hpqrs.alloc(sizeof(double2) * n4));
        QFE_CUDA(cudaMemcpyAsync(d_hpqrs.p, hpqrs, sizeof(double2) * n4, cudaMemcpyHostToDevice, s));
    }
    const int64_t total = n2 + n4;
    QFE_TRY(flags.alloc(sizeof(uint32_t) * total));
    QFE_TRY(pos.alloc(sizeof(uint32_t) * total));
    QFE_TRY(d_total.alloc(16));
    const int G = grid_for(total, 256, ctx->sm_count);
    integral_flags<<<G, 256, 0, s>>>(d_hpq.as<double2>(), d_hpqrs.as<double2>(), n2, n4, tol, flags.as<uint32_t>());
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    QFE_TRY(exclusive_scan_u32(ctx, flags.as<uint32_t>(), pos.as<uint32_t>(), total, scan_scratch, d_total.as<uint32_t>()));
    uint32_t nt32 = 0;
    QFE_CUDA(cudaMemcpyAsync(&nt32, d_total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    QFE_CUDA(cudaStreamSynchronize(s));
    const int64_t nt = nt32;
    QFE_TRY(d_terms.alloc(sizeof(FTerm) * (nt > 0 ? nt : 1)));
    if (nt > 0) {
        integral_emit<<<G, 256, 0, s>>>(d_hpq.as<double2>(), d_hpqrs.as<double2>(), n, n2, n4, flags.as<uint32_t>(), pos.as<uint32_t>(),
                                        d_terms.as<FTerm>());
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
    }
    return expand_and_canonicalise(ctx, n, d_terms, nt, encoding, constant, drop_eps, out);
}

int qfe_terms_from_fermion(qfe_ctx* ctx, int n, int64_t n_terms, const uint8_t* ops, const int32_t* qbits, const int64_t* offsets,
                           const double* coeffs, const double constant[2], double drop_eps, int encoding, qfe_terms** out) {
    QFE_REQUIRE(ctx && out && (n_terms == 0 || (ops && qbits && offsets && coeffs)), QFE_ERR_ARG, "qfe_terms_from_fermion: null argument");
    QFE_REQUIRE(n >= 1 && n <= 64, QFE_ERR_UNSUPPORTED, "qfe_terms_from_fermion: nqbits=%d outside 1..64", n);
    QFE_CUDA(cudaSetDevice(ctx->device));
    std::vector<FTerm> host(n_terms > 0 ? n_terms : 1);
    int64_t total_raw = 0;
    for (int64_t t = 0; t < n_terms; ++t) {
        const int64_t b = offsets[t], e = offsets[t + 1];
        QFE_REQUIRE(e >= b && e - b <= MAX_LETTERS, QFE_ERR_UNSUPPORTED, "term %lld has %lld ladder operators (max %d)", (long long)t,
                    (long long)(e - b), MAX_LETTERS);
        FTerm ft;
        memset(&ft, 0, sizeof(ft));
        ft.w_re = coeffs[2 * t];
        ft.w_im = coeffs[2 * t + 1];
        ft.len = (uint16_t)(e - b);
        for (int64_t l = b; l < e; ++l) {
            QFE_REQUIRE(ops[l] == 'C' || ops[l] == 'c', QFE_ERR_ARG, "FermionicTerm only accepts fermionic operators C and c (got '%c')", ops[l]);
            QFE_REQUIRE(qbits[l] >= 0 && qbits[l] < n, QFE_ERR_ARG, "mode index %d outside 0..%d", qbits[l], n - 1);
            ft.mode[l - b] = (uint8_t)qbits[l];
            if (ops[l] == 'C') ft.dagger |= (uint16_t)(1u << (l - b));
        }
        host[t] = ft;
        total_raw += int64_t(1) << ft.len;
    }
    QFE_REQUIRE(total_raw < (int64_t(1) << 32), QFE_ERR_UNSUPPORTED, "too many raw Pauli products (%lld)", (long long)total_raw);
    DevBuf d_terms;
    QFE_TRY(d_terms.alloc(sizeof(FTerm) * host.size()));
    QFE_CUDA(cudaMemcpyAsync(d_terms.p, host.data(), sizeof(FTerm) * host.size(), cudaMemcpyHostToDevice, ctx->stream));
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return expand_and_canonicalise(ctx, n, d_terms, n_terms, encoding, constant, drop_eps, out);
}

int qfe_terms_from_pauli(qfe_ctx* ctx, int n, int64_t T, const uint64_t* x, const uint64_t* z, const double* coeffs, const double constant[2],
                         double drop_eps, qfe_terms** out) {
    QFE_REQUIRE(ctx && out && (T == 0 || (x && z && coeffs)), QFE_ERR_ARG, "qfe_terms_from_pauli: null argument");
    QFE_REQUIRE(n >= 1 && n <= 64, QFE_ERR_UNSUPPORTED, "qfe_terms_from_pauli: nqbits=%d outside 1..64", n);
    QFE_REQUIRE(T < (int64_t(1) << 32), QFE_ERR_UNSUPPORTED, "too many terms");
    QFE_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t valid = n == 64 ? ~uint64_t(0) : ((uint64_t(1) << n) - 1);
    // Y-form -> XZ-form on the way in: c_XZ = c_Y * i^{popc(x&z)}
    std::vector<double> cxz(2 * (T > 0 ? T : 1));
    std::vector<uint32_t> own(T > 0 ? T : 1, 0u), idx(T > 0 ? T : 1);
    for (int64_t i = 0; i < T; ++i) {
        QFE_REQUIRE((x[i] & ~valid) == 0 && (z[i] & ~valid) == 0, QFE_ERR_ARG, "term %lld has mask bits beyond %d qubits", (long long)i, n);
        double re = coeffs[2 * i], im = coeffs[2 * i + 1];
        switch (__builtin_popcountll(x[i] & z[i]) & 3) {
            case 1: { double t = re; re = -im; im = t; break; }
            case 2: re = -re; im = -im; break;
            case 3: { double t = re; re = im; im = -t; break; }
            default: break;
        }
        cxz[2 * i] = re;
        cxz[2 * i + 1] = im;
        idx[i] = (uint32_t)i;
    }
    DevBuf rx, rz, rc, rowner, ridx;
    if (T > 0) {
        QFE_TRY(rx.alloc(sizeof(uint64_t) * T));
        QFE_TRY(rz.alloc(sizeof(uint64_t) * T));
        QFE_TRY(rc.alloc(sizeof(double2) * T));
        QFE_TRY(rowner.alloc(sizeof(uint32_t) * T));
        QFE_TRY(ridx.alloc(sizeof(uint32_t) * T));
        QFE_CUDA(cudaMemcpyAsync(rx.p, x, sizeof(uint64_t) * T, cudaMemcpyHostToDevice, s));
        QFE_CUDA(cudaMemcpyAsync(rz.p, z, sizeof(uint64_t) * T, cudaMemcpyHostToDevice, s));
        QFE_CUDA(cudaMemcpyAsync(rc.p, cxz.data(), sizeof(double2) * T, cudaMemcpyHostToDevice, s));
        QFE_CUDA(cudaMemcpyAsync(rowner.p, own.data(), sizeof(uint32_t) * T, cudaMemcpyHostToDevice, s));
        QFE_CUDA(cudaMemcpyAsync(ridx.p, idx.data(), sizeof(uint32_t) * T, cudaMemcpyHostToDevice, s));
        QFE_CUDA(cudaStreamSynchronize(s));
    }
    int rc_ = canonicalise(ctx, n, T, rx, rz, rc, rowner, ridx, 1, constant, drop_eps, out);
    QFE_CUDA(cudaStreamSynchronize(s));
    return rc_;
}

int qfe_terms_concat(qfe_ctx* ctx, const qfe_terms* const* list, int64_t K, qfe_terms** out) {
    QFE_REQUIRE(ctx && list && out && K >= 1, QFE_ERR_ARG, "qfe_terms_concat: bad argument");
    qfe_terms* t = new (std::nothrow) qfe_terms();
    QFE_REQUIRE(t, QFE_ERR_OOM, "out of host memory");
    t->n = list[0]->n;
    t->n_slots = K;
    t->constant.assign(2 * K, 0.0);
    for (int64_t k = 0; k < K; ++k) {
        const qfe_terms* s = list[k];
        if (!s || s->n != t->n || s->n_slots != 1) {
            delete t;
            set_error("qfe_terms_concat: operator %lld is null, batched, or on a different register", (long long)k);
            return QFE_ERR_ARG;
        }
        t->x.insert(t->x.end(), s->x.begin(), s->x.end());
        t->z.insert(t->z.end(), s->z.begin(), s->z.end());
        t->c.insert(t->c.end(), s->c.begin(), s->c.end());
        t->owner.insert(t->owner.end(), (size_t)s->T, (int32_t)k);
        t->constant[2 * k] = s->constant[0];
        t->constant[2 * k + 1] = s->constant[1];
        t->T += s->T;
    }
    count_xmasks(t);
    *out = t;
    return QFE_OK;
}

int qfe_terms_info(const qfe_terms* t, int* nqbits, int64_t* n_terms, int64_t* n_xmasks, int64_t* n_slots) {
    QFE_REQUIRE(t, QFE_ERR_ARG, "null terms");
    if (nqbits) *nqbits = t->n;
    if (n_terms) *n_terms = t->T;
    if (n_xmasks) *n_xmasks = t->n_xmasks;
    if (n_slots) *n_slots = t->n_slots;
    return QFE_OK;
}

int qfe_terms_export(const qfe_terms* t, uint64_t* x, uint64_t* z, double* coeffs, int32_t* owner, double* constant) {
    QFE_REQUIRE(t, QFE_ERR_ARG, "null terms");
    if (x && t->T) memcpy(x, t->x.data(), sizeof(uint64_t) * t->T);
    if (z && t->T) memcpy(z, t->z.data(), sizeof(uint64_t) * t->T);
    if (coeffs && t->T) memcpy(coeffs, t->c.data(), sizeof(double) * 2 * t->T);
    if (owner && t->T) memcpy(owner, t->owner.data(), sizeof(int32_t) * t->T);
    if (constant) memcpy(constant, t->constant.data(), sizeof(double) * 2 * t->n_slots);
    return QFE_OK;
}

int qfe_terms_destroy(qfe_terms* t) {
    delete t;
    return QFE_OK;
}

}  // extern "C"
