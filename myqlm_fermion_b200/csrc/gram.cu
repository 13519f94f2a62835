// Batched overlaps G[i][j] = <bra_i|ket_j> of state vectors: the quantum Fisher information matrix of the
// natural-gradient optimizer, g_kl = 4 Re(<d_k psi|d_l psi> - <d_k psi|psi><psi|d_l psi>)
// (/root/reference/qat/fermion/naturalgradient/qfim_plugin.py:159-199, which gets every entry from its own Hadamard-test circuit), the
// energy gradient 2 Re <d_k psi|H psi>, and the S / H matrices of the quantum subspace expansion
// (/root/reference/qat/fermion/chemistry/qse.py:191-214).
//
// One pass over a panel of up to 32 bras x 32 kets: a CTA stages a chunk of 32 amplitudes of every vector in shared memory (each vector
// is read from HBM once per panel pair, 128-bit coalesced loads, the next chunk travelling in registers while the current one is
// used) and its 256 threads accumulate the 32 x 32 block in registers, 4 x 4 entries per thread over every fourth amplitude.  The
// per-CTA blocks are reduced in a fixed order (bit-reproducible).  P(P+1)/2 separate dot products would read every vector P times.
#include <algorithm>

#include "common.cuh"

namespace qfe {

namespace {

constexpr int GP = 32;          // panel size (vectors per side)
constexpr int GC = 32;          // amplitudes per staged chunk
constexpr int GROW = 2 * GP + 1;  // words per staged amplitude: 32 bras, 32 kets, one word of padding (conflict-free partner reads)

struct GramArgs {
    const void* bra[GP];
    const void* ket[GP];
    int nb, nk;
    int diag;  // the bras ARE the kets (diagonal panel of a Hermitian Gram matrix): only the 4 x 4 blocks on and above the diagonal are formed
    long long n_chunks;
    double2* partials;  // [entry (i * GP + j)][CTA]
};

template <typename real_t> struct G2;
template <> struct G2<double> { using type = double2; };
template <> struct G2<float> { using type = float2; };

template <typename real_t>
__global__ void __launch_bounds__(256) gram_panel_kernel(const __grid_constant__ GramArgs a) {
    using cplx_t = typename G2<real_t>::type;
    __shared__ double2 s[GC * GROW];  // amplitude-major: word (amp, v) at amp * GROW + v
    const int tid = threadIdx.x;
    // entry block (bi, bj) of 4 x 4 entries, amplitudes sub, sub + 4, ... of every chunk.  A diagonal panel has 36 blocks with bi <= bj
    // (row-major over the upper triangle): 144 of the 256 threads accumulate (the kernel is FP64-bound), all of them load.
    int bi = tid >> 5, bj = (tid & 31) >> 2;
    const int sub = tid & 3;
    bool active = true;
    if (a.diag) {
        int rem = tid >> 2;
        active = rem < 36;
        bi = 0;
        while (active && rem >= 8 - bi) {
            rem -= 8 - bi;
            ++bi;
        }
        bj = active ? bi + rem : 0;
        if (!active) bi = 0;
    }
    // loader role: element e = tid + 256 u is amplitude (e & 31) of vector (e >> 5): a warp reads 32 consecutive amplitudes of one vector
    const cplx_t* src[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int v = (tid + 256 * u) >> 5;
        const void* p = v < GP ? (v < a.nb ? a.bra[v] : nullptr) : (v - GP < a.nk ? a.ket[v - GP] : nullptr);
        src[u] = reinterpret_cast<const cplx_t*>(p);
    }
    double accr[4][4], acci[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 4; ++l) accr[k][l] = acci[k][l] = 0.0;

    cplx_t nxt[8];
    auto fetch = [&](long long chunk) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            cplx_t z;
            z.x = 0;
            z.y = 0;
            nxt[u] = src[u] ? src[u][chunk * GC + (tid & 31)] : z;
        }
    };
    long long chunk = blockIdx.x;
    if (chunk < a.n_chunks) fetch(chunk);
    for (; chunk < a.n_chunks; chunk += gridDim.x) {
        __syncthreads();  // the previous chunk has been consumed
#pragma unroll
        for (int u = 0; u < 8; ++u) s[(tid & 31) * GROW + ((tid + 256 * u) >> 5)] = make_double2((double)nxt[u].x, (double)nxt[u].y);
        __syncthreads();
        if (chunk + gridDim.x < a.n_chunks) fetch(chunk + gridDim.x);
        if (active)
#pragma unroll 2
        for (int amp = sub; amp < GC; amp += 4) {
            const double2* row = s + amp * GROW;
            double2 b[4], k[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                b[q] = row[4 * bi + q];
                k[q] = row[GP + 4 * bj + q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int l = 0; l < 4; ++l) {  // conj(b) k
                    accr[q][l] = fma(b[q].x, k[l].x, fma(b[q].y, k[l].y, accr[q][l]));
                    acci[q][l] = fma(b[q].x, k[l].y, fma(-b[q].y, k[l].x, acci[q][l]));
                }
        }
    }
    // the four threads of an entry block (every fourth amplitude each) add up, then one partial per entry and CTA
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            double r = accr[q][l], i = acci[q][l];
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            i += __shfl_xor_sync(0xffffffffu, i, 1);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            i += __shfl_xor_sync(0xffffffffu, i, 2);
            if (active && sub == 0) a.partials[(size_t)((4 * bi + q) * GP + 4 * bj + l) * gridDim.x + blockIdx.x] = make_double2(r, i);
        }
}

}  // namespace
}  // namespace qfe

using namespace qfe;

extern "C" {

int qfe_vec_gram(qfe_ctx* ctx, const void* const* bras, int nb, const void* const* kets, int nk, int64_t n_amps, int dtype, double* out) {
    QFE_REQUIRE(ctx && bras && out && nb >= 1, QFE_ERR_ARG, "qfe_vec_gram: bad argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(n_amps >= GC && n_amps % GC == 0, QFE_ERR_UNSUPPORTED, "qfe_vec_gram: vectors of %lld amplitudes (a multiple of %d is needed; use qfe_vec_dot)",
                (long long)n_amps, GC);
    const bool symmetric = kets == nullptr;  // G = <v_i|v_j>: Hermitian, only the panels on and above the diagonal are computed
    if (symmetric) {
        kets = bras;
        nk = nb;
    }
    QFE_REQUIRE(nk >= 1, QFE_ERR_ARG, "qfe_vec_gram: no kets");
    for (int i = 0; i < nb; ++i) QFE_REQUIRE(bras[i], QFE_ERR_ARG, "qfe_vec_gram: null vector");
    for (int j = 0; j < nk; ++j) QFE_REQUIRE(kets[j], QFE_ERR_ARG, "qfe_vec_gram: null vector");
    QFE_CUDA(cudaSetDevice(ctx->device));
    const long long n_chunks = n_amps / GC;
    const int grid = (int)std::min<long long>(n_chunks, (long long)ctx->sm_count * 2);
    QFE_TRY(ctx->partials.reserve(sizeof(double2) * (size_t)GP * GP * grid));
    QFE_TRY(ctx->result.reserve(sizeof(double2) * GP * GP));
    std::vector<double> block(2 * GP * GP);
    for (int i0 = 0; i0 < nb; i0 += GP)
        for (int j0 = symmetric ? i0 : 0; j0 < nk; j0 += GP) {
            GramArgs a;
            memset(&a, 0, sizeof(a));
            a.nb = std::min(GP, nb - i0);
            a.nk = std::min(GP, nk - j0);
            for (int i = 0; i < a.nb; ++i) a.bra[i] = bras[i0 + i];
            for (int j = 0; j < a.nk; ++j) a.ket[j] = kets[j0 + j];
            a.diag = (symmetric && j0 == i0) ? 1 : 0;  // (the host reads the upper triangle of such a panel only; the rest of the partials is unspecified)
            a.n_chunks = n_chunks;
            a.partials = ctx->partials.as<double2>();
            if (dtype == QFE_C128)
                gram_panel_kernel<double><<<grid, 256, 0, ctx->stream>>>(a);
            else
                gram_panel_kernel<float><<<grid, 256, 0, ctx->stream>>>(a);
            ctx->launches++;
            QFE_CUDA(cudaGetLastError());
            QFE_TRY(reduce_to_vals(ctx, ctx->partials.as<double2>(), GP * GP, grid, ctx->result.as<double2>()));
            QFE_CUDA(cudaMemcpyAsync(block.data(), ctx->result.p, sizeof(double2) * GP * GP, cudaMemcpyDeviceToHost, ctx->stream));
            QFE_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < a.nb; ++i)
                for (int j = 0; j < a.nk; ++j) {
                    double re = block[2 * (i * GP + j)], im = block[2 * (i * GP + j) + 1];
                    if (symmetric && j0 == i0) {  // diagonal panel: exactly Hermitian output (upper triangle kept, real diagonal)
                        if (j < i) {
                            re = block[2 * (j * GP + i)];
                            im = -block[2 * (j * GP + i) + 1];
                        } else if (j == i) {
                            im = 0.0;
                        }
                    }
                    out[2 * ((size_t)(i0 + i) * nk + (j0 + j))] = re;
                    out[2 * ((size_t)(i0 + i) * nk + (j0 + j)) + 1] = im;
                    if (symmetric && j0 > i0) {  // mirror: G[j][i] = conj(G[i][j])
                        out[2 * ((size_t)(j0 + j) * nk + (i0 + i))] = re;
                        out[2 * ((size_t)(j0 + j) * nk + (i0 + i)) + 1] = -im;
                    }
                }
        }
    return QFE_OK;
}

}  // extern "C"
