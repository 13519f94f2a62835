// State preparation on the device: computational-basis states and ordered products of Pauli rotations
//   psi <- exp(-i theta_K P_K) ... exp(-i theta_1 P_1) psi,
// the building block of the reference's Trotter slices (make_spin_hamiltonian_trotter_slice,
// /root/reference/qat/fermion/trotterisation.py:85-150: per term CNOT ladder + RZ(2 coeff c_t) = exp(-i coeff c_t P_t)) and of the
// UCC / ADAPT ansatz layers (/root/reference/qat/plugins/adapt_vqe.py:64-141), which the reference hands to an external simulator.
//
// One rotation alone is a read+write pass over the state (HBM-bound, 32 B per amplitude).  Consecutive rotations whose X masks fit a
// common set of <= 13 bit positions are fused: a CTA stages the 2^k amplitudes that differ only in those positions in shared
// memory (same tiling as sweep.cu, natural bit order), applies the whole batch there -- one pair update per thread and rotation,
// a barrier between rotations -- and writes the tile back once.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace qfe {

constexpr int EV_NT = 512;
constexpr int EV_MAX_TILE_BITS = 13;
constexpr int EV_MAX_ROT = 512;  // rotations per fused batch (shared-memory table)

struct RotRec {
    uint32_t xl, zl;   // tile-local X and Z masks
    uint64_t zo;       // Z bits outside the tile
    double c;          // cos(theta)
    double wr, wi;     // w = -i sin(theta) (-i)^{|x&z|}: (exp(-i theta P) psi)[b] = c psi[b] + s_b w psi[b^x], s_b = (-1)^{|z&b|}
    uint32_t kpar;     // |x&z| mod 2: s_{b^x} = s_b (-1)^kpar
    uint32_t pad;
};

struct EvolveArgs {
    void* psi;
    const RotRec* rots;
    int n_rot, k;
    unsigned long long n_tiles;
    int n_runs;
    uint8_t run_pos[24], run_len[24];
    uint64_t spread0[128], spread1[64];  // tile-local bits 0..6 and 7..12 -> shard positions
};

template <typename real_t> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float> { using type = float2; };

template <typename real_t>
__global__ void __launch_bounds__(EV_NT, 1) rotate_tile_kernel(const __grid_constant__ EvolveArgs a) {
    using cplx_t = typename Cx<real_t>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    cplx_t* tile = reinterpret_cast<cplx_t*>(smem);
    RotRec* s_rot = reinterpret_cast<RotRec*>(smem + (size_t(sizeof(cplx_t)) << a.k));
    cplx_t* __restrict__ psi = reinterpret_cast<cplx_t*>(a.psi);
    const int tid = threadIdx.x;
    const uint32_t tile_amps = 1u << a.k;
    for (int r = tid; r < a.n_rot; r += EV_NT) s_rot[r] = a.rots[r];
    for (unsigned long long tile_id = blockIdx.x; tile_id < a.n_tiles; tile_id += gridDim.x) {
        uint64_t base = 0;
        {
            unsigned long long rest = tile_id;
            for (int r = 0; r < a.n_runs; ++r) {
                const int len = a.run_len[r];
                base |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
                rest >>= len;
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < tile_amps; j += EV_NT) tile[j] = psi[base | a.spread0[j & 127] | a.spread1[j >> 7]];
        __syncthreads();
        for (int r = 0; r < a.n_rot; ++r) {
            const RotRec q = s_rot[r];
            const uint32_t outer = (uint32_t)(__popcll(q.zo & base) & 1);
            if (q.xl == 0) {  // diagonal: psi[b] *= c + s_b w
                for (uint32_t j = tid; j < tile_amps; j += EV_NT) {
                    const real_t s = ((__popc(q.zl & j) & 1) ^ outer) ? (real_t)-1 : (real_t)1;
                    const real_t fr = (real_t)q.c + s * (real_t)q.wr, fi = s * (real_t)q.wi;
                    const cplx_t v = tile[j];
                    cplx_t o;
                    o.x = fr * v.x - fi * v.y;
                    o.y = fr * v.y + fi * v.x;
                    tile[j] = o;
                }
            } else {
                // pairs (j, j ^ xl) with the top X bit of j clear: consecutive threads touch consecutive words on both sides
                const int top = 31 - __clz(q.xl);
                const uint32_t lowmask = (1u << top) - 1u;
                for (uint32_t p = tid; p < (tile_amps >> 1); p += EV_NT) {
                    const uint32_t j = ((p & ~lowmask) << 1) | (p & lowmask);
                    const uint32_t jp = j ^ q.xl;
                    const real_t s = ((__popc(q.zl & j) & 1) ^ outer) ? (real_t)-1 : (real_t)1;
                    const real_t sp = q.kpar ? -s : s;
                    const cplx_t u = tile[j], v = tile[jp];
                    const real_t c = (real_t)q.c, wr = (real_t)q.wr, wi = (real_t)q.wi;
                    cplx_t nu, nv;
                    nu.x = c * u.x + s * (wr * v.x - wi * v.y);
                    nu.y = c * u.y + s * (wr * v.y + wi * v.x);
                    nv.x = c * v.x + sp * (wr * u.x - wi * u.y);
                    nv.y = c * v.y + sp * (wr * u.y + wi * u.x);
                    tile[j] = nu;
                    tile[jp] = nv;
                }
            }
            __syncthreads();
        }
        for (uint32_t j = tid; j < tile_amps; j += EV_NT) psi[base | a.spread0[j & 127] | a.spread1[j >> 7]] = tile[j];
    }
}

// One rotation whose X mask does not fit a tile (e.g. a parity-encoded hopping term a+_p a_q, whose X string is |p - q| qubits long):
// one thread per pair (b, b ^ x) with the top X bit of b clear, straight from global memory -- a plain read + write pass.
struct WideRot {
    uint64_t x, z;
    double c, wr, wi;
    uint32_t kpar;
};

template <typename real_t>
__global__ void __launch_bounds__(256) rotate_wide_kernel(typename Cx<real_t>::type* __restrict__ psi, const unsigned long long n_pairs, const WideRot q) {
    using cplx_t = typename Cx<real_t>::type;
    const int top = 63 - __clzll((long long)q.x);
    const unsigned long long lowmask = (1ull << top) - 1ull;
    const real_t c = (real_t)q.c, wr = (real_t)q.wr, wi = (real_t)q.wi;
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long j = ((p & ~lowmask) << 1) | (p & lowmask);
        const unsigned long long jp = j ^ q.x;
        const real_t s = (__popcll(q.z & j) & 1) ? (real_t)-1 : (real_t)1;
        const real_t sp = q.kpar ? -s : s;
        const cplx_t u = psi[j], v = psi[jp];
        cplx_t nu, nv;
        nu.x = c * u.x + s * (wr * v.x - wi * v.y);
        nu.y = c * u.y + s * (wr * v.y + wi * v.x);
        nv.x = c * v.x + sp * (wr * u.x - wi * u.y);
        nv.y = c * v.y + sp * (wr * u.y + wi * u.x);
        psi[j] = nu;
        psi[jp] = nv;
    }
}

template <typename real_t>
__global__ void __launch_bounds__(256) basis_state_kernel(typename Cx<real_t>::type* __restrict__ x, int64_t n, int64_t index) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        typename Cx<real_t>::type v;
        v.x = (i == index) ? (real_t)1 : (real_t)0;
        v.y = (real_t)0;
        x[i] = v;
    }
}

}  // namespace qfe

using namespace qfe;

extern "C" {

int qfe_vec_basis_state(qfe_ctx* ctx, void* x, int64_t n, int dtype, int64_t index) {
    QFE_REQUIRE(ctx && x && n > 0, QFE_ERR_ARG, "qfe_vec_basis_state: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(index < n, QFE_ERR_ARG, "qfe_vec_basis_state: index %lld outside the %lld amplitudes", (long long)index, (long long)n);
    QFE_CUDA(cudaSetDevice(ctx->device));
    const int grid = grid_for(n, 256, ctx->sm_count, 16);
    if (dtype == QFE_C128)
        basis_state_kernel<double><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double2*>(x), n, index);
    else
        basis_state_kernel<float><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<float2*>(x), n, index);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

int qfe_pauli_rotations(qfe_ctx* ctx, int nqbits, int64_t K, const uint64_t* x, const uint64_t* z, const double* theta, void* psi, int dtype) {
    QFE_REQUIRE(ctx && psi && (K == 0 || (x && z && theta)), QFE_ERR_ARG, "qfe_pauli_rotations: null argument");
    QFE_REQUIRE(nqbits >= 1 && nqbits <= 40, QFE_ERR_ARG, "qfe_pauli_rotations: nqbits=%d outside 1..40", nqbits);
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    const uint64_t all = (uint64_t(1) << nqbits) - 1;
    for (int64_t i = 0; i < K; ++i)
        QFE_REQUIRE(!((x[i] | z[i]) & ~all), QFE_ERR_ARG, "qfe_pauli_rotations: rotation %lld acts outside the %d qubits", (long long)i, nqbits);
    const int kmax = std::min(nqbits, EV_MAX_TILE_BITS);
    const uint64_t lowmask = (uint64_t(1) << std::min(3, nqbits)) - 1;
    const size_t elt = dtype == QFE_C128 ? 16 : 8;
    // all batches are planned first: one upload of every rotation table, then the passes back to back (no allocation, no
    // synchronisation between them)
    std::vector<RotRec> recs;
    recs.reserve((size_t)K);
    std::vector<EvolveArgs> passes;
    std::vector<int64_t> pass_first;
    std::vector<WideRot> wide;        // rotations too wide for a tile, each its own pass
    std::vector<int64_t> pass_wide;   // per pass: index into `wide`, or -1 for a fused tile pass
    auto rotation_weights = [&](int64_t r, double& c, double& wr, double& wi, uint32_t& kpar) {
        const int kk = __builtin_popcountll(x[r] & z[r]) & 3;
        const double sn = std::sin(theta[r]);
        c = std::cos(theta[r]);
        // w = -i sin (-i)^k
        switch (kk) {
            case 0: wr = 0.0; wi = -sn; break;
            case 1: wr = -sn; wi = 0.0; break;
            case 2: wr = 0.0; wi = sn; break;
            default: wr = sn; wi = 0.0; break;
        }
        kpar = (uint32_t)(kk & 1);
    };
    int64_t i = 0;
    while (i < K) {
        // greedy batch: consecutive rotations whose X masks fit kmax positions together with the 3 low bits
        uint64_t L = lowmask;
        int64_t j = i;
        while (j < K && j - i < EV_MAX_ROT) {
            const uint64_t cand = L | x[j];
            if (__builtin_popcountll(cand) > kmax) break;
            L = cand;
            ++j;
        }
        if (j == i) {
            // flips more qubits than a tile holds besides the low bits: a pass of its own through global memory
            WideRot w;
            w.x = x[i];
            w.z = z[i];
            rotation_weights(i, w.c, w.wr, w.wi, w.kpar);
            EvolveArgs none;
            memset(&none, 0, sizeof(none));
            passes.push_back(none);
            pass_first.push_back((int64_t)recs.size());
            pass_wide.push_back((int64_t)wide.size());
            wide.push_back(w);
            ++i;
            continue;
        }
        // pad the tile mask to kmax positions (lowest free positions): larger tiles, fewer of them
        for (int p = 0; p < nqbits && __builtin_popcountll(L) < kmax; ++p) L |= uint64_t(1) << p;
        EvolveArgs a;
        memset(&a, 0, sizeof(a));
        a.psi = psi;
        a.k = __builtin_popcountll(L);
        a.n_tiles = 1ull << (nqbits - a.k);
        int pos[EV_MAX_TILE_BITS];
        {
            int np = 0;
            for (int p = 0; p < nqbits; ++p)
                if ((L >> p) & 1) pos[np++] = p;
        }
        auto spread = [&](uint32_t jloc) {
            uint64_t m = 0;
            for (int b = 0; b < a.k; ++b)
                if ((jloc >> b) & 1) m |= uint64_t(1) << pos[b];
            return m;
        };
        auto compress = [&](uint64_t m) {
            uint32_t o = 0;
            for (int b = 0; b < a.k; ++b)
                if ((m >> pos[b]) & 1) o |= 1u << b;
            return o;
        };
        for (uint32_t e = 0; e < 128; ++e) a.spread0[e] = spread(e & ((1u << a.k) - 1u));
        for (uint32_t e = 0; e < 64; ++e) a.spread1[e] = spread((e << 7) & ((1u << a.k) - 1u));
        {
            const uint64_t outer = all & ~L;
            int b = 0;
            while (b < nqbits) {
                if ((outer >> b) & 1) {
                    int e = b;
                    while (e < nqbits && ((outer >> e) & 1)) ++e;
                    a.run_pos[a.n_runs] = (uint8_t)b;
                    a.run_len[a.n_runs] = (uint8_t)(e - b);
                    ++a.n_runs;
                    b = e;
                } else {
                    ++b;
                }
            }
        }
        pass_first.push_back((int64_t)recs.size());
        for (int64_t r = i; r < j; ++r) {
            RotRec q;
            memset(&q, 0, sizeof(q));
            q.xl = compress(x[r]);
            q.zl = compress(z[r] & L);
            q.zo = z[r] & ~L;
            rotation_weights(r, q.c, q.wr, q.wi, q.kpar);
            recs.push_back(q);
        }
        a.n_rot = (int)(j - i);
        passes.push_back(a);
        pass_wide.push_back(-1);
        i = j;
    }
    if (passes.empty()) return QFE_OK;
    QFE_TRY(ctx->rots.reserve(sizeof(RotRec) * std::max<size_t>(recs.size(), 1)));
    if (!recs.empty()) QFE_CUDA(cudaMemcpyAsync(ctx->rots.p, recs.data(), sizeof(RotRec) * recs.size(), cudaMemcpyHostToDevice, ctx->stream));
    static int attr_device[2] = {-1, -1};  // per instantiation: the shared-memory opt-in is set once per device
    for (size_t b = 0; b < passes.size(); ++b) {
        if (pass_wide[b] >= 0) {
            const WideRot& w = wide[(size_t)pass_wide[b]];
            const unsigned long long n_pairs = 1ull << (nqbits - 1);
            const int grid = grid_for((int64_t)n_pairs, 256, ctx->sm_count, 16);
            if (dtype == QFE_C128)
                rotate_wide_kernel<double><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double2*>(psi), n_pairs, w);
            else
                rotate_wide_kernel<float><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<float2*>(psi), n_pairs, w);
            ctx->launches++;
            QFE_CUDA(cudaGetLastError());
            continue;
        }
        EvolveArgs& a = passes[b];
        a.rots = ctx->rots.as<RotRec>() + pass_first[b];
        const size_t smem = (elt << a.k) + sizeof(RotRec) * EV_MAX_ROT;
        const size_t smem_max = (elt << EV_MAX_TILE_BITS) + sizeof(RotRec) * EV_MAX_ROT;
        const int grid = (int)std::min<unsigned long long>(a.n_tiles, (unsigned long long)ctx->sm_count);
        if (dtype == QFE_C128) {
            auto kern = rotate_tile_kernel<double>;
            if (attr_device[0] != ctx->device) {
                QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
                attr_device[0] = ctx->device;
            }
            kern<<<grid, EV_NT, smem, ctx->stream>>>(a);
        } else {
            auto kern = rotate_tile_kernel<float>;
            if (attr_device[1] != ctx->device) {
                QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
                attr_device[1] = ctx->device;
            }
            kern<<<grid, EV_NT, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
    }
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));  // the call is synchronous: recs (pageable) and the table stay valid until here
    return QFE_OK;
}

}  // extern "C"
