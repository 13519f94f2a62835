// Kernel (2): matrix-free Pauli-sum apply / expectation on a statevector shard.
//
// Replaces SpinHamiltonian.get_matrix (/root/reference/qat/fermion/hamiltonians.py:163-250) plus the
// matrix-vector products at the reference's call sites.  The reference materialises a 2^n x 2^n
// matrix; here every launch ("sweep") streams the shard through shared memory once:
//
//   * a CTA stages one *tile* -- 2^k amplitudes whose index differs only in the k bit positions of the sweep (k - 4 positions that
//     hold its X masks, 4 register positions they avoid) -- with coalesced 128-bit loads (128 B contiguous segments);
//   * a thread keeps the 16 amplitudes that differ on the register positions in registers; every XOR partner psi[b ^ x] of every
//     term group of the sweep is a shared-memory read at [register + immediate] (conflict-free: a warp reads a permutation of 32
//     consecutive 16 B words of one plane);
//   * signs come from popc parity: the part of z inside the tile is one __popc per (term, thread) shared by the thread's 16 rows;
//     the part outside the tile is a second mask against the tile number (real-coefficient sweeps keep their term records in the
//     kernel parameter = constant bank, next to the shared-memory pipe) or folded into the staged coefficient per tile;
//   * W_x(b) = sum_t c_t (-1)^{|z_t & b|} costs one integer multiply-add on the sign bit and one DADD per (term, thread);
//   * the expectation is reduced by warp shuffles, then per block, then by ONE fixed-order reduce kernel per class (bit-reproducible);
//   * APPLY accumulates W_x(row) * partner in registers and leaves through the tile: y is read and written in the natural order the
//     tile was staged in (128 B contiguous, 8 loads in flight per thread, its lines prefetched to L2 before the group loops);
//   * a shard with fewer tiles than SMs is worked on by up to 16 CTAs per tile, each with a slice of the groups of every range;
//   * qfe_expectation_host / qfe_expectation_class0_host upload the state on a second stream in 16 pieces and run the first sweeps
//     piece by piece behind the copy.
// Kernels: tile_sweep_kernel (generic: EXPECT / APPLY / BATCH, complex coefficients, small tiles, tiles shared by several CTAs) and, for
// real-coefficient sweeps over full 2^13-amplitude tiles, tile_expect2_kernel (<bra|H|ket>: control words in the constant bank, the
// dominant kernel of the bench), tile_expect_kernel (its predecessor: full complex transition values, QFE_EXPECT2=0) and
// tile_apply_kernel (y (+)= H|ket>); direct_kernel for shards smaller than a tile and as a cross-check.  DESIGN.md section 4 has the
// measurements behind each of them and the variants that were built and not kept.
// No tensor cores: this is an XOR-gather with +-1 weights, not a dense contraction.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "plan_host.h"

namespace qfe {

// Threads per CTA of the tile kernel: one warp per row block of 512 amplitudes.  Tiles of 2^13 amplitudes run as one CTA of 512
// threads per SM; tiles of <= 2^12 as two CTAs of 256 threads per SM, so that one CTA stages its next tile while the other computes.
constexpr int NT_MAX = 512;
constexpr int RB = PLAN_REG_BITS; // register-blocked rows per thread: R = 16 (tile-local bits 0..3)
constexpr int R = 1 << RB;
// shared-memory word of tile-local index j is (j & 15) * 513 + (j >> 4): one plane of 512 words per register row, padded to an
// odd length.  Thread t owns word t of every plane; the partner of row r for X mask (xt, xr = 0) is word t ^ xt of the same plane,
// i.e. 16 loads at [register + immediate]; a warp reads a permutation of 32 consecutive 16 B words (conflict-free), and the 8
// consecutive amplitudes a quarter warp stages land on 8 different bank groups (different planes, or consecutive words).
static_assert(PLAN_MIN_TILE_BITS == 5 + RB, "a tile must hold at least one full warp row block");
static_assert(PLAN_U_SHIFT == 5 + RB, "warp-uniform tile bits start above the lane and register bits");
static_assert((1 << (PLAN_MAX_TILE_BITS - 5 - RB)) == NT_MAX / 32, "one warp per row block of the largest tile");

// EXPECT: <bra|H|ket>; with bra tile == ket tile the own rows come from the tile and Hermitian ranges visit half the rows
// APPLY:  y (+)= H|ket>
// BATCH:  one value per group (operator pools): <bra| G_g |ket>
enum { MODE_EXPECT = 0, MODE_APPLY = 1, MODE_BATCH = 2 };

template <typename real_t> struct CplxOf;
template <> struct CplxOf<double> { using type = double2; };
template <> struct CplxOf<float> { using type = float2; };

// one staged term: coefficient (sign of the z bits outside the tile folded in per tile) and the word t_zl of plan_host.h
// (z bits on lane / row-block positions, register-row bits v, run length)
template <typename real_t, bool CPLX> struct TermRec;
template <> struct __align__(16) TermRec<double, false> { double c; uint32_t zp; uint32_t pad; };
template <> struct __align__(16) TermRec<double, true> { double c; double ci; uint32_t zp; uint32_t pad[3]; };
template <> struct __align__(8) TermRec<float, false> { float c; uint32_t zp; };
template <> struct __align__(16) TermRec<float, true> { float c; float ci; uint32_t zp; uint32_t pad; };

// one staged term of a real-coefficient sweep as it travels in the kernel parameter (constant bank): coefficient as on tile 0
// (double, or float in c_lo for complex64 states), the word t_zl and the packed outer z bits t_zoc of plan_host.h
struct TermConst { uint32_t c_lo, c_hi, zp, zoc; };
constexpr int SWEEP_CONST_TERMS = 1920;  // 30 KB of the 32 KB a kernel may take as parameters
constexpr int PROF_PHASES = 8;           // development counters: per warp, cycles spent in each phase of the tile loop

struct SweepArgs {
    const void* ket;
    const void* bra;
    void* out_vec;
    double2* partials;
    const uint32_t* g_hdr;
    const uint32_t* r_desc;
    const uint32_t* t_zl;
    const uint32_t* t_zoc;
    const double* t_cre;
    const double* t_cim;
    int n_groups, n_terms, n_ranges;
    const uint64_t* nat_off;  // PLAN_TABLE entries each, see plan_host.h
    const uint64_t* loc_off;
    const uint32_t* nat_loc;
    int k, n_runs;
    int same_tile, accumulate;
    int prefetch_dist;  // tiles ahead that go to L2 while the current one is processed (0: off)
    int split;          // CTAs that share one tile, each taking a slice of the groups of every range (shards with fewer tiles than SMs)
    int pstride;        // BATCH: partials of group g start at g * pstride
    unsigned long long slab_amps;  // APPLY with split > 1: CTA part p writes its partial result at out_vec + p * slab_amps
    unsigned long long tile_begin, n_tiles;  // the launch covers tiles [tile_begin, n_tiles)
    unsigned long long tile_stride;          // gridDim.x / split, set by the launcher
    unsigned long long ket_xor;
    unsigned long long* prof;  // development: per-warp phase cycle counters (QFE_PROF=1), else null
    int stagger;               // cycles CTA b waits before its first tile, times b: spreads the staging bursts of the CTAs over a tile period
    uint8_t run_pos[PLAN_MAX_RUNS];
    uint8_t run_len[PLAN_MAX_RUNS];
    TermConst tc[SWEEP_CONST_TERMS];  // real-coefficient sweeps only
};

static_assert(sizeof(SweepArgs) <= 32764, "kernel parameter space");

// Fixed shared-memory map (bytes), sized for the largest sweep (2^13 complex128 amplitudes, 2048 complex-coefficient terms):
// constant offsets let every access be [register + immediate] with nothing to recompute inside the group loop.
template <int NTH> struct Geo {
    static constexpr int NW = NTH / 32;
    static constexpr uint32_t PLANE = NTH + 1;              // words per register-row plane (odd: see above)
    static constexpr int CPLX_TERMS = NTH == NT_MAX ? 2048 : 1024;  // complex-coefficient terms a sweep may stage in shared memory
    static constexpr uint32_t REC_BYTES = CPLX_TERMS * 32;
    static constexpr uint32_t SM_TILE = 0;                  // 16 planes x (NTH + 1) x 16 B
    static constexpr uint32_t SM_REC = R * PLANE * 16;      // staged complex-coefficient terms
    static constexpr uint32_t SM_ACC = SM_REC + REC_BYTES / 2;  // BATCH only (half the terms staged): per-warp per-group accumulators
    static constexpr uint32_t SM_HDR = SM_REC + REC_BYTES;  // 4 KB group headers (512 x 8 B)
    static constexpr uint32_t SM_RNG = SM_HDR + 4096;       // 2 KB range descriptors (256 x 8 B)
    static constexpr uint32_t SM_NATOFF = SM_RNG + 2048;    // index tables of the sweep (plan_host.h)
    static constexpr uint32_t SM_LOCOFF = SM_NATOFF + PLAN_TABLE * 8;
    static constexpr uint32_t SM_NATLOC = SM_LOCOFF + PLAN_TABLE * 8;
    static constexpr uint32_t SM_ZOC = SM_NATLOC + PLAN_TABLE * 4;  // z bits of the staged terms outside the tile, packed against the tile number
    static constexpr uint32_t SM_RED = SM_ZOC + CPLX_TERMS * 4;     // block reduction scratch
    static constexpr uint32_t SM_PROF = SM_RED + NW * 16;           // development counters (QFE_PROF=1)
    static constexpr uint32_t SM_TEND = SM_PROF + NW * PROF_PHASES * 8;  // development: when each warp finished the tile's groups
    static constexpr uint32_t SM_TOTAL = SM_TEND + NW * 8;
};
constexpr int SWEEP_MAX_GROUPS = 512;  // ranges per sweep <= 5 skip levels x 2 polarities x 3 kinds
constexpr int BATCH_MAX_GROUPS = 128;
// staged terms per sweep by tile size: real-coefficient sweeps are bounded by the kernel parameter, complex ones by shared memory
inline int sweep_max_terms(int k, bool batch) {
    const bool big = k >= PLAN_MAX_TILE_BITS;
    if (batch) return big ? 1024 : 512;
    return big ? SWEEP_CONST_TERMS : 1024;
}
static_assert((size_t)Geo<512>::NW * BATCH_MAX_GROUPS * 16 <= Geo<512>::REC_BYTES / 2 && (size_t)Geo<256>::NW * BATCH_MAX_GROUPS * 16 <= Geo<256>::REC_BYTES / 2,
              "batch accumulators share the term area");
static_assert(Geo<512>::SM_TOTAL <= 232448 - 1024, "exceeds the 227 KB of shared memory a CTA may use");
static_assert(2 * (Geo<256>::SM_TOTAL + 1024) <= 232448, "two CTAs of 256 threads must fit one SM");

__device__ __forceinline__ double2 block_reduce_sum(double2 v, double2* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, d);
        v.y += __shfl_down_sync(0xffffffffu, v.y, d);
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double2 r = make_double2(0.0, 0.0);
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) {
            r.x += s_red[w].x;
            r.y += s_red[w].y;
        }
    }
    return r;  // valid in thread 0
}

// streaming 128-bit / 64-bit loads that do not pollute L1
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// c with its sign flipped iff `par` is odd: adding par << 31 to the high word flips the sign bit and nothing else (one IMAD / LEA)
__device__ __forceinline__ double sign_by_parity(double c, uint32_t par) {
    return __hiloint2double((int)((uint32_t)__double2hiint(c) + (par << 31)), __double2loint(c));
}
__device__ __forceinline__ float sign_by_parity(float c, uint32_t par) { return __uint_as_float(__float_as_uint(c) + (par << 31)); }

// (-1)^{bit k of v} as a floating-point factor
template <typename real_t> __device__ __forceinline__ real_t sigma(uint32_t v, int k);
template <> __device__ __forceinline__ double sigma<double>(uint32_t v, int k) { return __hiloint2double((int)(0x3ff00000u | ((v >> k) << 31)), 0); }
template <> __device__ __forceinline__ float sigma<float>(uint32_t v, int k) { return __uint_as_float(0x3f800000u | ((v >> k) << 31)); }

// one shared-memory load per staged term (real-coefficient records)
__device__ __forceinline__ void load_rec(const TermRec<double, false>* p, double& c, double& ci, uint32_t& zp) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    c = __hiloint2double((int)raw.y, (int)raw.x);
    ci = 0.0;
    zp = raw.z;
}
__device__ __forceinline__ void load_rec(const TermRec<float, false>* p, float& c, float& ci, uint32_t& zp) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    c = __uint_as_float(raw.x);
    ci = 0.0f;
    zp = raw.y;
}
__device__ __forceinline__ void load_rec(const TermRec<double, true>* p, double& c, double& ci, uint32_t& zp) {
    const double2 cc = *reinterpret_cast<const double2*>(p);
    c = cc.x;
    ci = cc.y;
    zp = p->zp;
}
__device__ __forceinline__ void load_rec(const TermRec<float, true>* p, float& c, float& ci, uint32_t& zp) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    c = __uint_as_float(raw.x);
    ci = __uint_as_float(raw.y);
    zp = raw.z;
}

template <uint32_t PL> __device__ __forceinline__ uint32_t tile_word(uint32_t j) { return (j & (uint32_t)(R - 1)) * PL + (j >> RB); }

// partner of the thread's row r: word t ^ xt of plane r (pp points at it), or of plane r ^ xr when the X mask touches register rows
template <bool XROW, uint32_t PL, typename cplx_t>
__device__ __forceinline__ cplx_t partner_row(const cplx_t* __restrict__ pp, const uint32_t xr, const int r) {
    return XROW ? pp[((uint32_t)r ^ xr) * PL] : pp[(uint32_t)r * PL];
}

// Where the staged terms of a sweep live.  Complex-coefficient sweeps keep them in shared memory (TermRec, signs flipped in place per
// tile).  Real-coefficient sweeps read them from the kernel parameter: indexed constant-bank loads run beside the shared-memory
// pipe, which the partner rows saturate (a broadcast LDS.128 costs 2 of its cycles, measured); the sign of the z bits outside the
// tile joins the parity, parity(A) ^ parity(B) = parity(A ^ B).
template <typename real_t, bool CPLX> struct Terms;
template <typename real_t> struct Terms<real_t, true> {
    const TermRec<real_t, true>* rec;
    __device__ __forceinline__ void load(const int t, const uint32_t jhi, real_t& c, real_t& ci, uint32_t& par) const {
        uint32_t zp;
        load_rec(rec + t, c, ci, zp);
        par = (uint32_t)__popc(zp & jhi);
    }
    __device__ __forceinline__ uint32_t head(const int t) const { return rec[t].zp; }
};
template <> struct Terms<double, false> {
    const SweepArgs& a;
    uint32_t tile;
    __device__ __forceinline__ void load(const int t, const uint32_t jhi, double& c, double& ci, uint32_t& par) const {
        const TermConst& q = a.tc[t];
        c = __hiloint2double((int)q.c_hi, (int)q.c_lo);
        ci = 0.0;
        par = (uint32_t)__popc((q.zp & jhi) ^ (q.zoc & tile));
    }
    __device__ __forceinline__ uint32_t head(const int t) const { return a.tc[t].zp; }
};
template <> struct Terms<float, false> {
    const SweepArgs& a;
    uint32_t tile;
    __device__ __forceinline__ void load(const int t, const uint32_t jhi, float& c, float& ci, uint32_t& par) const {
        const TermConst& q = a.tc[t];
        c = __uint_as_float(q.c_lo);
        ci = 0.0f;
        par = (uint32_t)__popc((q.zp & jhi) ^ (q.zoc & tile));
    }
    __device__ __forceinline__ uint32_t head(const int t) const { return a.tc[t].zp; }
};

// S = sum_t c_t (-1)^{|z_t & jhi|} over the staged terms [t, te).  jhi only has lane / row-block bits set, so the run fields in the
// upper half of the term word never reach the popc.  EVEN: te - t is even (KIND_FAST lists are padded): two independent chains.
template <typename real_t, bool CPLX, bool EVEN>
__device__ __forceinline__ void signed_sum(const Terms<real_t, CPLX>& terms, int t, const int te, const uint32_t jhi, real_t& Sre, real_t& Sim) {
    real_t S0 = (real_t)0, S1 = (real_t)0, I0 = (real_t)0, I1 = (real_t)0;
    if (EVEN) {
#pragma unroll 1
        for (; t < te; t += 2) {
            real_t c0, c1, d0, d1;
            uint32_t p0, p1;
            terms.load(t, jhi, c0, d0, p0);
            terms.load(t + 1, jhi, c1, d1, p1);
            S0 += sign_by_parity(c0, p0);
            S1 += sign_by_parity(c1, p1);
            if (CPLX) {
                I0 += sign_by_parity(d0, p0);
                I1 += sign_by_parity(d1, p1);
            }
        }
    } else {
#pragma unroll 1
        for (; t < te; ++t) {
            real_t c0, d0;
            uint32_t p0;
            terms.load(t, jhi, c0, d0, p0);
            S0 += sign_by_parity(c0, p0);
            if (CPLX) I0 += sign_by_parity(d0, p0);
        }
    }
    Sre = S0 + S1;
    Sim = I0 + I1;
}

// the same sum for a compile-time term count: all loads are issued before the first use (KIND_FAST groups of 2, 4 or 6 terms)
template <typename real_t, bool CPLX, int N>
__device__ __forceinline__ void signed_sum_fixed(const Terms<real_t, CPLX>& terms, const int t, const uint32_t jhi, real_t& Sre, real_t& Sim) {
    real_t c[N], d[N];
    uint32_t p[N];
#pragma unroll
    for (int i = 0; i < N; ++i) terms.load(t + i, jhi, c[i], d[i], p[i]);
    real_t S0 = (real_t)0, S1 = (real_t)0, I0 = (real_t)0, I1 = (real_t)0;
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        S0 += sign_by_parity(c[i], p[i]);
        S1 += sign_by_parity(c[i + 1], p[i + 1]);
        if (CPLX) {
            I0 += sign_by_parity(d[i], p[i]);
            I1 += sign_by_parity(d[i + 1], p[i + 1]);
        }
    }
    Sre = S0 + S1;
    Sim = I0 + I1;
}

// KIND_FAST term sum: the common even counts are straight-line code
template <typename real_t, bool CPLX>
__device__ __forceinline__ void signed_sum_fast(const Terms<real_t, CPLX>& terms, const int t, const int nt, const uint32_t jhi, real_t& Sre, real_t& Sim) {
    switch (nt) {
        case 2: signed_sum_fixed<real_t, CPLX, 2>(terms, t, jhi, Sre, Sim); break;
        case 4: signed_sum_fixed<real_t, CPLX, 4>(terms, t, jhi, Sre, Sim); break;
        case 6: signed_sum_fixed<real_t, CPLX, 6>(terms, t, jhi, Sre, Sim); break;
        default: signed_sum<real_t, CPLX, true>(terms, t, t + nt, jhi, Sre, Sim); break;
    }
}

// sum_r (-1)^{v.r} f_r over N = 2^m rows with the low m bits of v: m halving steps of fused multiply-adds
template <typename real_t, int N>
__device__ __forceinline__ real_t fold_rows(const real_t (&f)[N], const uint32_t v) {
    constexpr int TOP = (N == 16 ? 3 : 2);
    real_t g[N / 2];  // f stays intact for the next run: the first step writes the half-size copy
    {
        const real_t s = sigma<real_t>(v, TOP);
#pragma unroll
        for (int r = 0; r < N / 2; ++r) g[r] = fma(s, f[r + N / 2], f[r]);
    }
#pragma unroll
    for (int h = N / 4, bit = TOP - 1; h >= 1; h >>= 1, --bit) {
        const real_t s = sigma<real_t>(v, bit);
#pragma unroll
        for (int r = 0; r < h; ++r) g[r] = fma(s, g[r + h], g[r]);
    }
    return g[0];
}

// (sr, si) * Wc with Wc = S (real), i S (imag) or Sre + i Sim (CPLX)
template <typename real_t, bool CPLX>
__device__ __forceinline__ void times_weight(real_t& sr, real_t& si, const real_t Sre, const real_t Sim, const bool imag) {
    if (CPLX) {
        const real_t t = Sre * sr - Sim * si;
        si = Sre * si + Sim * sr;
        sr = t;
    } else if (imag) {
        const real_t t = -Sre * si;
        si = Sre * sr;
        sr = t;
    } else {
        sr *= Sre;
        si *= Sre;
    }
}

// <own| G |tile> of one group for the thread's 16 rows: sum_r conj(own_r) Wc(r) p_r with p_r the partner of row r.
// RE_ONLY (real-coefficient sweeps): only the real part of the value is wanted -- the two ranks that hold a pair of shards of a Hermitian
// class each add 2 Re <own|G|partner> (sharded.cu) -- so only Re conj(own) p (real weights) or Im conj(own) p (imaginary weights) is
// formed: half of the FP64 work of a full visit.
template <typename real_t, bool CPLX, int KIND, uint32_t PL, bool RE_ONLY = false>
__device__ __forceinline__ void group_value_full(const typename CplxOf<real_t>::type* __restrict__ pp, const uint32_t xr, const Terms<real_t, CPLX>& terms,
                                                 const int tb, const int nt, const bool imag, const uint32_t jhi,
                                                 const typename CplxOf<real_t>::type (&own)[R], real_t& out_re, real_t& out_im) {
    using cplx_t = typename CplxOf<real_t>::type;
    (void)xr;
    if (RE_ONLY && !CPLX) {
        // Re (Wc sum_r conj(own_r) p_r): Wc = S  ->  S Re(.) ;  Wc = i S  ->  -S Im(.)
        if (KIND == KIND_FAST) {
            real_t S, Si;
            signed_sum_fast<real_t, CPLX>(terms, tb, nt, jhi, S, Si);
            real_t a0 = (real_t)0, a1 = (real_t)0, a2 = (real_t)0, a3 = (real_t)0;
            if (!imag) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                    if (r & 1) {
                        a2 = fma(own[r].x, p.x, a2);
                        a3 = fma(own[r].y, p.y, a3);
                    } else {
                        a0 = fma(own[r].x, p.x, a0);
                        a1 = fma(own[r].y, p.y, a1);
                    }
                }
                out_re = S * ((a0 + a1) + (a2 + a3));
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                    if (r & 1) {
                        a2 = fma(own[r].x, p.y, a2);
                        a3 = fma(-own[r].y, p.x, a3);
                    } else {
                        a0 = fma(own[r].x, p.y, a0);
                        a1 = fma(-own[r].y, p.x, a1);
                    }
                }
                out_re = -S * ((a0 + a1) + (a2 + a3));
            }
        } else {
            real_t f[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                f[r] = imag ? -fma(own[r].x, p.y, -own[r].y * p.x) : fma(own[r].x, p.x, own[r].y * p.y);
            }
            int t = tb;
            const int te = tb + nt;
            real_t tot = (real_t)0;
#pragma unroll 1
            while (t < te) {
                const uint32_t head = terms.head(t);
                const uint32_t v = (head >> TZ_V_SHIFT) & 15u;
                const int len = (int)(head >> TZ_RUN_SHIFT);
                real_t S, Si;
                signed_sum<real_t, CPLX, false>(terms, t, t + len, jhi, S, Si);
                t += len;
                tot = fma(S, fold_rows<real_t, R>(f, v), tot);
            }
            out_re = tot;
        }
        out_im = (real_t)0;
        return;
    }
    if (KIND == KIND_FAST) {
        real_t Sre, Sim;
        signed_sum_fast<real_t, CPLX>(terms, tb, nt, jhi, Sre, Sim);
        real_t a0 = (real_t)0, a1 = (real_t)0, b0 = (real_t)0, b1 = (real_t)0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
            a0 = fma(own[r].x, p.x, a0);  // Re conj(own) p = a a' + b b'
            a1 = fma(own[r].y, p.y, a1);
            b0 = fma(own[r].x, p.y, b0);  // Im conj(own) p = a b' - b a'
            b1 = fma(-own[r].y, p.x, b1);
        }
        real_t sr = a0 + a1, si = b0 + b1;
        times_weight<real_t, CPLX>(sr, si, Sre, Sim, imag);
        out_re = sr;
        out_im = si;
    } else {
        // rows in two chunks of 8 (keeps 16 products, not 32, live); the term sums of a run are formed once per chunk
        real_t tot_re = (real_t)0, tot_im = (real_t)0;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            real_t fr[8], fi[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = 8 * ch + q;
                const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                fr[q] = fma(own[r].x, p.x, own[r].y * p.y);
                fi[q] = fma(own[r].x, p.y, -own[r].y * p.x);
            }
            int t = tb;
            const int te = tb + nt;
#pragma unroll 1
            while (t < te) {
                const uint32_t head = terms.head(t);
                const uint32_t v = (head >> TZ_V_SHIFT) & 15u;
                const int len = (int)(head >> TZ_RUN_SHIFT);
                real_t Sre, Sim;
                signed_sum<real_t, CPLX, false>(terms, t, t + len, jhi, Sre, Sim);
                t += len;
                real_t sr = fold_rows<real_t, 8>(fr, v), si = fold_rows<real_t, 8>(fi, v);
                if (ch) {
                    const real_t s3 = sigma<real_t>(v, 3);
                    sr *= s3;
                    si *= s3;
                }
                times_weight<real_t, CPLX>(sr, si, Sre, Sim, imag);
                tot_re += sr;
                tot_im += si;
            }
        }
        out_re = tot_re;
        out_im = tot_im;
    }
}

// All groups of one range (same skip bit, polarity and kind) for the thread's 16 rows.
template <typename real_t, bool CPLX, int MODE, int KIND, uint32_t PL, bool RE_ONLY = false>
__device__ __forceinline__ void process_range(const typename CplxOf<real_t>::type* __restrict__ tile, const Terms<real_t, CPLX>& terms,
                                              const uint2* __restrict__ s_hdr, double2* __restrict__ s_acc_warp, const int g0, const int g1, int tb,
                                              const uint32_t w0, const uint32_t jhi, const bool half,
                                              const typename CplxOf<real_t>::type (&own)[R], real_t (&acc_re)[R], real_t (&acc_im)[R],
                                              double& e_re, double& e_im) {
    using cplx_t = typename CplxOf<real_t>::type;
    constexpr bool PREFETCH = MODE == MODE_EXPECT;  // APPLY has no register to spare for it
    uint32_t hdr_next = PREFETCH ? s_hdr[g0].x : 0u;
#pragma unroll 1
    for (int g = g0; g < g1; ++g) {
        const uint32_t hdr = PREFETCH ? hdr_next : s_hdr[g].x;
        if (PREFETCH && g + 1 < g1) hdr_next = s_hdr[g + 1].x;  // next group's header travels while this group is processed
        const cplx_t* __restrict__ pp = tile + (w0 ^ ((hdr & HDR_XL_MASK) >> RB));  // partner thread's word of plane 0
        const uint32_t xr = hdr & (uint32_t)(R - 1);
        const int nt = (int)((hdr >> HDR_NT_SHIFT) & HDR_NT_MASK);
        const bool imag = hdr & HDR_IMAG;
        if (MODE == MODE_EXPECT && half) {
            // Hermitian group with real W: the pair (b, b^x) sums to 2 Re(conj(psi_b) W psi_p) = 2 W (a a' + b b'); only rows whose
            // skip bit equals the range polarity get here.
            if (KIND == KIND_FAST) {
                real_t S, Si;
                signed_sum_fast<real_t, CPLX>(terms, tb, nt, jhi, S, Si);
                real_t a0 = (real_t)0, a1 = (real_t)0, a2 = (real_t)0, a3 = (real_t)0;  // independent chains keep the FP64 pipe busy
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                    if (r & 1) {
                        a2 = fma(own[r].x, p.x, a2);
                        a3 = fma(own[r].y, p.y, a3);
                    } else {
                        a0 = fma(own[r].x, p.x, a0);
                        a1 = fma(own[r].y, p.y, a1);
                    }
                }
                e_re = fma((double)(S + S), (double)((a0 + a1) + (a2 + a3)), e_re);  // the pair counts twice
            } else {
                real_t f[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                    f[r] = fma(own[r].x, p.x, own[r].y * p.y);
                }
                int t = tb;
                const int te = tb + nt;
                real_t tot = (real_t)0;
#pragma unroll 1
                while (t < te) {
                    const uint32_t head = terms.head(t);
                    const uint32_t v = (head >> TZ_V_SHIFT) & 15u;
                    const int len = (int)(head >> TZ_RUN_SHIFT);
                    real_t S, Si;
                    signed_sum<real_t, CPLX, false>(terms, t, t + len, jhi, S, Si);
                    t += len;
                    tot = fma(S, fold_rows<real_t, R>(f, v), tot);
                }
                e_re += (double)(tot + tot);
            }
        } else if (MODE == MODE_EXPECT || MODE == MODE_BATCH) {
            real_t sr, si;
            group_value_full<real_t, CPLX, KIND, PL, RE_ONLY>(pp, xr, terms, tb, nt, imag, jhi, own, sr, si);
            if (MODE == MODE_EXPECT) {
                e_re += (double)sr;
                e_im += (double)si;
            } else {
                double g_re = (double)sr, g_im = (double)si;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    g_re += __shfl_down_sync(0xffffffffu, g_re, d);
                    g_im += __shfl_down_sync(0xffffffffu, g_im, d);
                }
                if ((threadIdx.x & 31) == 0) {  // only this warp writes its own row of the accumulators: deterministic
                    double2 cur = s_acc_warp[g];
                    cur.x += g_re;
                    cur.y += g_im;
                    s_acc_warp[g] = cur;
                }
            }
        } else {  // MODE_APPLY: acc_r += Wc(r) p_r
            if (KIND == KIND_FAST) {
                real_t Sre, Sim;
                signed_sum_fast<real_t, CPLX>(terms, tb, nt, jhi, Sre, Sim);
                // Wc = wr + i wi
                const real_t wr = (!CPLX && imag) ? (real_t)0 : Sre;
                const real_t wi = CPLX ? Sim : (imag ? Sre : (real_t)0);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                    if (CPLX) {
                        acc_re[r] += wr * p.x - wi * p.y;
                        acc_im[r] += wr * p.y + wi * p.x;
                    } else if (imag) {
                        acc_re[r] = fma(-wi, p.y, acc_re[r]);
                        acc_im[r] = fma(wi, p.x, acc_im[r]);
                    } else {
                        acc_re[r] = fma(wr, p.x, acc_re[r]);
                        acc_im[r] = fma(wr, p.y, acc_im[r]);
                    }
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    real_t W[8], Wi[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) W[q] = Wi[q] = (real_t)0;
                    int t = tb;
                    const int te = tb + nt;
#pragma unroll 1
                    while (t < te) {
                        const uint32_t head = terms.head(t);
                        const uint32_t v = (head >> TZ_V_SHIFT) & 15u;
                        const int len = (int)(head >> TZ_RUN_SHIFT);
                        real_t Sre, Sim;
                        signed_sum<real_t, CPLX, false>(terms, t, t + len, jhi, Sre, Sim);
                        t += len;
                        if (ch) {
                            const real_t s3 = sigma<real_t>(v, 3);
                            Sre *= s3;
                            Sim *= s3;
                        }
                        // W(q) += S (-1)^{v.q}: unfold the three low bits
                        const real_t s0 = sigma<real_t>(v, 0), s1 = sigma<real_t>(v, 1), s2 = sigma<real_t>(v, 2);
                        real_t e[8];
                        e[0] = Sre;
                        e[1] = s0 * e[0];
                        e[2] = s1 * e[0];
                        e[3] = s1 * e[1];
#pragma unroll
                        for (int q = 0; q < 4; ++q) e[4 + q] = s2 * e[q];
#pragma unroll
                        for (int q = 0; q < 8; ++q) W[q] += e[q];
                        if (CPLX) {
                            e[0] = Sim;
                            e[1] = s0 * e[0];
                            e[2] = s1 * e[0];
                            e[3] = s1 * e[1];
#pragma unroll
                            for (int q = 0; q < 4; ++q) e[4 + q] = s2 * e[q];
#pragma unroll
                            for (int q = 0; q < 8; ++q) Wi[q] += e[q];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int r = 8 * ch + q;
                        const cplx_t p = partner_row<(KIND & KIND_XROW) != 0, PL>(pp, xr, r);
                        if (CPLX) {
                            acc_re[r] += W[q] * p.x - Wi[q] * p.y;
                            acc_im[r] += W[q] * p.y + Wi[q] * p.x;
                        } else if (imag) {
                            acc_re[r] = fma(-W[q], p.y, acc_re[r]);
                            acc_im[r] = fma(W[q], p.x, acc_im[r]);
                        } else {
                            acc_re[r] = fma(W[q], p.x, acc_re[r]);
                            acc_im[r] = fma(W[q], p.y, acc_im[r]);
                        }
                    }
                }
            }
        }
        tb += nt;
    }
}

// sign of staged term t on tile T = parity(zoc_t & T) (z bits outside the tile against the tile number): moving from tile T' to T
// flips the terms with odd |zoc_t & (T ^ T')|, in place
template <typename real_t, bool CPLX>
__device__ __forceinline__ void flip_signs(TermRec<real_t, CPLX>* __restrict__ s_rec, const uint32_t* __restrict__ s_zoc, const int nT, const uint32_t diff) {
    if (diff == 0) return;
    for (int t = threadIdx.x; t < nT; t += blockDim.x) {
        if (__popc(s_zoc[t] & diff) & 1) {
            s_rec[t].c = -s_rec[t].c;
            if constexpr (CPLX) s_rec[t].ci = -s_rec[t].ci;
        }
    }
}

template <typename real_t, bool CPLX>
__device__ __forceinline__ Terms<real_t, CPLX> make_terms(const SweepArgs& a, const TermRec<real_t, CPLX>* s_rec, const uint32_t tile) {
    if constexpr (CPLX)
        return Terms<real_t, true>{s_rec};
    else
        return Terms<real_t, false>{a, tile};
}

// development counters (QFE_PROF=1): lane 0 of every warp adds the cycles since its previous stamp to phase `ph`
template <bool PROF>
__device__ __forceinline__ void prof_stamp(unsigned long long* s_prof, long long& t_prev, const int ph) {
    if (PROF) {
        const long long now = clock64();
        if ((threadIdx.x & 31) == 0) s_prof[(threadIdx.x >> 5) * PROF_PHASES + ph] += (unsigned long long)(now - t_prev);
        t_prev = now;
    }
}
enum { PH_WAIT = 0, PH_ISSUE = 1, PH_GROUPS = 2, PH_SKEW = 3, PH_LOAD = 4, PH_STAGE = 5, PH_ASYNC_TILES = 7 };

template <typename real_t, bool CPLX, int MODE, int NTH, bool PROF = false>
__global__ void __launch_bounds__(NTH, NT_MAX / NTH) tile_sweep_kernel(const __grid_constant__ SweepArgs a) {
    using G = Geo<NTH>;
    constexpr int NT = NTH, NW = G::NW;
    constexpr uint32_t PLANE = G::PLANE;
    using cplx_t = typename CplxOf<real_t>::type;
    using rec_t = TermRec<real_t, CPLX>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = a.k;
    const uint32_t tile_amps = 1u << k;
    const int nT = a.n_terms, nG = a.n_groups, nR = a.n_ranges;
    cplx_t* tile = reinterpret_cast<cplx_t*>(smem + G::SM_TILE);
    rec_t* s_rec = reinterpret_cast<rec_t*>(smem + G::SM_REC);
    uint2* s_hdr = reinterpret_cast<uint2*>(smem + G::SM_HDR);
    uint2* s_rng = reinterpret_cast<uint2*>(smem + G::SM_RNG);
    uint64_t* s_nat_off = reinterpret_cast<uint64_t*>(smem + G::SM_NATOFF);
    uint64_t* s_loc_off = reinterpret_cast<uint64_t*>(smem + G::SM_LOCOFF);
    uint32_t* s_nat_loc = reinterpret_cast<uint32_t*>(smem + G::SM_NATLOC);
    uint32_t* s_zoc = reinterpret_cast<uint32_t*>(smem + G::SM_ZOC);
    double2* s_acc = reinterpret_cast<double2*>(smem + G::SM_ACC);
    double2* s_red = reinterpret_cast<double2*>(smem + G::SM_RED);
    unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(smem + G::SM_PROF);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (PROF && tid < NW * PROF_PHASES) s_prof[tid] = 0ull;
    // development counters: every warp notes when it finished the groups of a tile; after the barrier the latest of these times splits
    // the tile period into groups (own), skew (waiting for the slowest warp) and staging (from the slowest warp's end to the next groups)
    long long* s_tend = reinterpret_cast<long long*>(smem + G::SM_TEND);
    long long t_gs = 0ll, t_end = 0ll, t_last = 0ll;
    bool have_prev = false;
    if (a.stagger > 0) {
        if (tid == 0) {
            const long long t0 = clock64(), wait = (long long)blockIdx.x * a.stagger;
            while (clock64() - t0 < wait) __nanosleep(400);
        }
        __syncthreads();
    }
    const cplx_t* __restrict__ ket = reinterpret_cast<const cplx_t*>(a.ket);
    const cplx_t* __restrict__ bra = reinterpret_cast<const cplx_t*>(a.bra);
    cplx_t* __restrict__ yout = reinterpret_cast<cplx_t*>(a.out_vec);

    // ---- once per CTA: tile-independent tables ------------------------------------------------
    for (int g = tid; g < nG; g += NT) s_hdr[g] = make_uint2(a.g_hdr[2 * g], a.g_hdr[2 * g + 1]);
    for (int g = tid; g < nR; g += NT) {
        uint2 d = make_uint2(a.r_desc[2 * g], a.r_desc[2 * g + 1]);
        if (a.split > 1) {
            // small shards: a.split CTAs stage the same tile; this one keeps its slice of the groups of every range (possibly none)
            const uint32_t split = (uint32_t)a.split, part = blockIdx.x % split;
            const uint32_t g0 = d.x & 0xffffu, len = (d.x >> 16) - g0;
            const uint32_t gs = g0 + len * part / split, ge = g0 + len * (part + 1u) / split;
            d.x = gs | (ge << 16);
            if (gs < ge) d.y = (d.y & ~0xffffu) | a.g_hdr[2 * gs + 1];  // second header word: the group's first staged term
        }
        s_rng[g] = d;
    }
    for (int i = tid; i < PLAN_TABLE; i += NT) {
        s_nat_off[i] = a.nat_off[i];
        s_loc_off[i] = a.loc_off[i];
        s_nat_loc[i] = a.nat_loc[i];
    }
    // staged terms: coefficients as on tile 0; every tile flips the signs that change against the previous tile it staged
    if constexpr (CPLX) {
        for (int t = tid; t < nT; t += NT) {
            s_rec[t].zp = a.t_zl[t];
            s_rec[t].c = (real_t)a.t_cre[t];
            s_rec[t].ci = (real_t)a.t_cim[t];
            s_zoc[t] = a.t_zoc[t];
        }
    }
    unsigned long long prev_tile = 0;
    __syncthreads();
    const uint64_t off_lo = s_nat_off[tid & 127];
    const uint32_t loc_lo = s_nat_loc[tid & 127];
    if (MODE == MODE_BATCH)
        for (int i = tid; i < NW * nG; i += NT) s_acc[i] = make_double2(0.0, 0.0);

    // one warp per row block of 512 amplitudes: thread tid owns the 16 rows jhi | r = word tid of the 16 planes
    const bool active = warp < (1 << (k - 5 - RB));
    const uint32_t jhi = (uint32_t)tid << RB;  // tile-local index of row 0 (lane and row-block bits)
    const uint32_t w0 = (uint32_t)tid;
    double e_re = 0.0, e_im = 0.0;

    // small shards: a.split CTAs stage the same tile (their range descriptors differ, see above)
    // (a.split sits in the constant bank; nothing of this is kept in registers across the group loops)
    if (MODE == MODE_APPLY) yout += (size_t)(blockIdx.x % (uint32_t)a.split) * a.slab_amps;

    for (unsigned long long tile_id = a.tile_begin + blockIdx.x / (uint32_t)a.split; tile_id < a.n_tiles; tile_id += a.tile_stride) {
        // shard position of the tile: deposit the tile number into the bits outside the tile mask
        uint64_t base = 0;
        {
            unsigned long long rest = tile_id;
            for (int r = 0; r < a.n_runs; ++r) {
                const int len = a.run_len[r];
                base |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
                rest >>= len;
            }
        }
        __syncthreads();  // previous tile fully consumed (also orders the one-time tables)
        if (PROF && have_prev && (tid & 31) == 0) {
            t_last = s_tend[0];
            for (int w = 1; w < NW; ++w) t_last = max(t_last, s_tend[w]);
            s_prof[warp * PROF_PHASES + PH_SKEW] += (unsigned long long)(t_last - t_end);
        }
        {
            const uint64_t kbase = base ^ a.ket_xor;
            // the tile is enumerated in natural order (consecutive threads = consecutive amplitudes of a 128 B segment) and
            // scattered to its tile-local words; a full tile is one batch of 16 independent 128-bit loads per thread
            if (tile_amps == R * NT) {
                // idx = tid + u * NT: the low 7 index bits are the thread's own, looked up once (off_lo, loc_lo)
                cplx_t v[R];
#pragma unroll
                for (int u = 0; u < R; ++u) v[u] = ld_stream(ket + (kbase | off_lo | s_nat_off[128 + ((tid + u * NT) >> 7)]));
                if constexpr (CPLX) flip_signs<real_t, CPLX>(s_rec, s_zoc, nT, (uint32_t)(tile_id ^ prev_tile));  // shared memory only: overlaps the loads
#pragma unroll
                for (int u = 0; u < R; ++u) tile[tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + u * NT) >> 7)])] = v[u];
            } else {
                if constexpr (CPLX) flip_signs<real_t, CPLX>(s_rec, s_zoc, nT, (uint32_t)(tile_id ^ prev_tile));
                for (uint32_t idx = tid; idx < tile_amps; idx += NT)
                    tile[tile_word<PLANE>(s_nat_loc[idx & 127] | s_nat_loc[128 + (idx >> 7)])] =
                        ld_stream(ket + (kbase | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)]));
            }
            prev_tile = tile_id;
        }
        __syncthreads();
        // next tile of this CTA -> L2 while this one is processed (two 128 B lines per thread)
        if (a.prefetch_dist > 0 && tile_id + (unsigned long long)a.prefetch_dist * a.tile_stride < a.n_tiles) {
            uint64_t nbase = 0;
            unsigned long long rest = tile_id + (unsigned long long)a.prefetch_dist * a.tile_stride;
            for (int r = 0; r < a.n_runs; ++r) {
                const int len = a.run_len[r];
                nbase |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
                rest >>= len;
            }
            nbase ^= a.ket_xor;
            for (uint32_t idx = 8u * tid; idx < tile_amps; idx += 8u * NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ket + (nbase | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)])));
        }
        // y += ...: the tile's lines of y travel to L2 while the groups are processed, for the read-modify-write that ends the tile
        if (MODE == MODE_APPLY && a.accumulate && a.prefetch_dist > 0)
            for (uint32_t idx = 8u * tid; idx < tile_amps; idx += 8u * NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(yout + (base | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)])));

        real_t acc_re[R], acc_im[R];
        if (active) {
            cplx_t own[R];
            if (MODE == MODE_APPLY) {
#pragma unroll
                for (int r = 0; r < R; ++r) acc_re[r] = acc_im[r] = (real_t)0;
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (a.same_tile) {
                        own[r] = tile[(uint32_t)r * PLANE + w0];
                    } else {
                        const uint32_t j = jhi | (uint32_t)r;
                        own[r] = bra[base | s_loc_off[j & 127] | s_loc_off[128 + (j >> 7)]];
                    }
                }
            }
            if (PROF && (tid & 31) == 0) {
                t_gs = clock64();
                if (have_prev) s_prof[warp * PROF_PHASES + PH_STAGE] += (unsigned long long)(t_gs - t_last);
            }
            double2* s_acc_warp = s_acc + (size_t)warp * nG;
            const Terms<real_t, CPLX> terms = make_terms<real_t, CPLX>(a, s_rec, (uint32_t)tile_id);
#pragma unroll 1
            for (int ri = 0; ri < nR; ++ri) {
                const uint2 d = s_rng[ri];
                const uint32_t skipbit = d.y >> RNG_SKIP_SHIFT;
                bool half = false;
                if (MODE == MODE_EXPECT) {
                    half = a.same_tile && skipbit != 0;
                    // the row blocks (warps) of the other polarity hold the partners of these pairs
                    if (half && (((jhi >> skipbit) ^ (d.y >> RNG_POL_SHIFT)) & 1u)) continue;
                }
                const int g0 = (int)(d.x & 0xffffu), g1 = (int)(d.x >> 16), t0 = (int)(d.y & 0xffffu);
                const uint32_t kind = (d.y >> RNG_KIND_SHIFT) & 3u;
                if (kind == KIND_FAST)
                    process_range<real_t, CPLX, MODE, KIND_FAST, PLANE>(tile, terms, s_hdr, s_acc_warp, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
                else if (kind == KIND_RUNS)
                    process_range<real_t, CPLX, MODE, KIND_RUNS, PLANE>(tile, terms, s_hdr, s_acc_warp, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
                else
                    process_range<real_t, CPLX, MODE, KIND_RUNS | KIND_XROW, PLANE>(tile, terms, s_hdr, s_acc_warp, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
            }
        }
        if (PROF && (tid & 31) == 0) {
            t_end = clock64();
            s_prof[warp * PROF_PHASES + PH_GROUPS] += (unsigned long long)(t_end - t_gs);
            s_prof[warp * PROF_PHASES + PH_ASYNC_TILES - 1] += 1ull;  // slot 6: tiles
            s_tend[warp] = t_end;
            have_prev = true;
        }
        if (MODE == MODE_APPLY) {
            // the result leaves through the tile, in the natural order it was staged in: 128 B contiguous per 8 threads, and the 16
            // loads of y a thread needs for y += ... are in flight together
            __syncthreads();  // every partner row has been read
            if (active) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    cplx_t o;
                    o.x = acc_re[r];
                    o.y = acc_im[r];
                    tile[(uint32_t)r * PLANE + w0] = o;
                }
            }
            __syncthreads();
            if (tile_amps == R * NT) {
#pragma unroll
                for (int h = 0; h < R; h += R / 2) {  // two batches of 8 loads in flight
                    cplx_t v[R / 2];
                    if (a.accumulate) {
#pragma unroll
                        for (int u = 0; u < R / 2; ++u) v[u] = yout[base | off_lo | s_nat_off[128 + ((tid + (h + u) * NT) >> 7)]];
                    }
#pragma unroll
                    for (int u = 0; u < R / 2; ++u) {
                        cplx_t o = tile[tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + (h + u) * NT) >> 7)])];
                        if (a.accumulate) {
                            o.x += v[u].x;
                            o.y += v[u].y;
                        }
                        yout[base | off_lo | s_nat_off[128 + ((tid + (h + u) * NT) >> 7)]] = o;
                    }
                }
            } else {
                for (uint32_t idx = tid; idx < tile_amps; idx += NT) {
                    const uint64_t gidx = base | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)];
                    cplx_t o = tile[tile_word<PLANE>(s_nat_loc[idx & 127] | s_nat_loc[128 + (idx >> 7)])];
                    if (a.accumulate) {
                        const cplx_t y = yout[gidx];
                        o.x += y.x;
                        o.y += y.y;
                    }
                    yout[gidx] = o;
                }
            }
        }
    }

    if (PROF && a.prof) {
        __syncthreads();
        if (tid < NW * PROF_PHASES) atomicAdd(a.prof + tid, s_prof[tid]);
    }
    if (MODE == MODE_EXPECT) {
        double2 tot = block_reduce_sum(make_double2(e_re, e_im), s_red);
        if (tid == 0) a.partials[blockIdx.x] = tot;
    } else if (MODE == MODE_BATCH) {
        __syncthreads();
        for (int g = tid; g < nG; g += NT) {
            double2 s = make_double2(0.0, 0.0);
            for (int w = 0; w < NW; ++w) {
                s.x += s_acc[w * nG + g].x;
                s.y += s_acc[w * nG + g].y;
            }
            a.partials[(size_t)g * a.pstride + blockIdx.x] = s;
        }
    }
}

// ---- pipelined EXPECT kernel -----------------------------------------------------------------------------------------------------
// Real-coefficient EXPECT sweeps over full tiles (2^13 amplitudes, 512 threads): the dominant launches of <psi|H|psi>.  Same tile
// layout and group loops as tile_sweep_kernel, with the serial chain between two tiles shortened (ncu source view of the generic
// kernel: the warps of a CTA spend 19 % of a 32-qubit step between the last group of one tile and the first group of the next, on a
// chain of dependent latencies rather than on bandwidth):
//   * a warp that is done with its groups requests its 16 amplitudes of the CTA's NEXT tile at once (the registers of its own rows
//     are free then; the lines were sent to L2 a tile earlier), so the load latency passes while the slower warps finish and before
//     the barrier, not after it;
//   * tile number -> shard offset is four independent table lookups (the generic kernel walks the runs of the outer mask, a dependent
//     chain of ~10 steps, twice per tile);
//   * the own rows are requested first after the barrier, the L2 prefetch of the tile after the next one is issued behind them.
// Tried and measured slower on this kernel (profiles/r02_phase_async_v1.jsonl, r02_phase3.jsonl): staging half of the next tile with
// cp.async into a third half-tile region while the groups run -- 16-byte LDGSTS scatter costs more than LDG + STS, and the latency it
// would hide is not what the tile switch waits for.
template <typename real_t> struct GeoExpect {
    using cplx_t = typename CplxOf<real_t>::type;
    static constexpr int NTH = NT_MAX, NW = NTH / 32;
    static constexpr uint32_t PLANE = NTH + 1;
    static constexpr uint32_t SM_TILE = 0;
    static constexpr uint32_t SM_HDR = ((R * PLANE * (uint32_t)sizeof(cplx_t)) + 15u) & ~15u;
    static constexpr uint32_t SM_RNG = SM_HDR + 4096;
    static constexpr uint32_t SM_NATOFF = SM_RNG + 2048;
    static constexpr uint32_t SM_LOCOFF = SM_NATOFF + PLAN_TABLE * 8;
    static constexpr uint32_t SM_BASE = SM_LOCOFF + PLAN_TABLE * 8;   // 4 x 128 entries: 7 bits of the tile number -> shard offset
    static constexpr uint32_t SM_NATLOC = SM_BASE + 4 * 128 * 8;
    static constexpr uint32_t SM_RED = SM_NATLOC + PLAN_TABLE * 4;
    static constexpr uint32_t SM_PROF = SM_RED + NW * 16;
    static constexpr uint32_t SM_TEND = SM_PROF + NW * PROF_PHASES * 8;
    static constexpr uint32_t SM_TOTAL = SM_TEND + NW * 8;
};
static_assert(GeoExpect<double>::SM_TOTAL <= 232448 - 1024, "exceeds the 227 KB of shared memory a CTA may use");

template <typename real_t, bool PROF, bool RE_ONLY>
__global__ void __launch_bounds__(NT_MAX, 1) tile_expect_kernel(const __grid_constant__ SweepArgs a) {
    using G = GeoExpect<real_t>;
    constexpr int NT = NT_MAX, NW = G::NW;
    constexpr uint32_t PLANE = G::PLANE;
    using cplx_t = typename CplxOf<real_t>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    cplx_t* const tile = reinterpret_cast<cplx_t*>(smem + G::SM_TILE);
    uint2* s_hdr = reinterpret_cast<uint2*>(smem + G::SM_HDR);
    uint2* s_rng = reinterpret_cast<uint2*>(smem + G::SM_RNG);
    uint64_t* s_nat_off = reinterpret_cast<uint64_t*>(smem + G::SM_NATOFF);
    uint64_t* s_loc_off = reinterpret_cast<uint64_t*>(smem + G::SM_LOCOFF);
    uint64_t* s_base = reinterpret_cast<uint64_t*>(smem + G::SM_BASE);
    uint32_t* s_nat_loc = reinterpret_cast<uint32_t*>(smem + G::SM_NATLOC);
    double2* s_red = reinterpret_cast<double2*>(smem + G::SM_RED);
    unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(smem + G::SM_PROF);
    long long* s_tend = reinterpret_cast<long long*>(smem + G::SM_TEND);  // development counters, as in tile_sweep_kernel
    long long t_gs = 0ll, t_end = 0ll, t_last = 0ll;
    bool have_prev = false;

    const int nG = a.n_groups, nR = a.n_ranges;
    const int tid = threadIdx.x;
    const cplx_t* __restrict__ ket = reinterpret_cast<const cplx_t*>(a.ket);
    const cplx_t* __restrict__ bra = reinterpret_cast<const cplx_t*>(a.bra);

    for (int g = tid; g < nG; g += NT) s_hdr[g] = make_uint2(a.g_hdr[2 * g], a.g_hdr[2 * g + 1]);
    for (int g = tid; g < nR; g += NT) s_rng[g] = make_uint2(a.r_desc[2 * g], a.r_desc[2 * g + 1]);
    for (int i = tid; i < PLAN_TABLE; i += NT) {
        s_nat_off[i] = a.nat_off[i];
        s_loc_off[i] = a.loc_off[i];
        s_nat_loc[i] = a.nat_loc[i];
    }
    {
        // entry (k, v): the 7 bits v of the tile number at bit 7k, deposited into the runs of the outer mask
        unsigned long long rest = (unsigned long long)(tid & 127) << (7 * (tid >> 7));
        uint64_t b = 0;
        for (int r = 0; r < a.n_runs; ++r) {
            const int len = a.run_len[r];
            b |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
            rest >>= len;
        }
        s_base[tid] = b;
    }
    if (PROF && tid < NW * PROF_PHASES) s_prof[tid] = 0ull;
    __syncthreads();

    const uint32_t jhi = (uint32_t)tid << RB, w0 = (uint32_t)tid;
    double e_re = 0.0, e_im = 0.0;
    auto base_of = [&](const unsigned long long t) {  // tile numbers stay below 2^28 (shards of at most 2^40 amplitudes, plan_create)
        return s_base[t & 127] | s_base[128 + ((t >> 7) & 127)] | s_base[256 + ((t >> 14) & 127)] | s_base[384 + ((t >> 21) & 127)];
    };
    // the thread's 16 amplitudes of a tile in natural order (idx = tid + u * NT: consecutive threads = consecutive amplitudes of a
    // 128 B segment); the low 7 index bits are the thread's own
    const uint64_t off_lo = s_nat_off[tid & 127];
    const uint32_t loc_lo = s_nat_loc[tid & 127];
    cplx_t v[R];
    uint32_t vw[R];  // their words in the tile (the same for every tile; looked up again per tile rather than held across the group loops)
    auto request = [&](const unsigned long long t) {
        const uint64_t kb = (base_of(t) ^ a.ket_xor) | off_lo;
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = ld_stream(ket + (kb | s_nat_off[128 + ((tid + u * NT) >> 7)]));
#pragma unroll
        for (int u = 0; u < R; ++u) vw[u] = tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + u * NT) >> 7)]);
    };

    unsigned long long tile_id = a.tile_begin + blockIdx.x;
    if (tile_id < a.n_tiles) request(tile_id);
    for (; tile_id < a.n_tiles; tile_id += a.tile_stride) {
        const unsigned long long next_id = tile_id + a.tile_stride;
        __syncthreads();  // every partner row of the previous tile has been read
        if (PROF && have_prev && (tid & 31) == 0) {
            t_last = s_tend[0];
            for (int w = 1; w < NW; ++w) t_last = max(t_last, s_tend[w]);
            s_prof[(tid >> 5) * PROF_PHASES + PH_SKEW] += (unsigned long long)(t_last - t_end);
        }
#pragma unroll
        for (int u = 0; u < R; ++u) tile[vw[u]] = v[u];
        __syncthreads();
        {
            cplx_t own[R];
            const uint64_t base = a.same_tile ? 0ull : base_of(tile_id);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (a.same_tile) {
                    own[r] = tile[(uint32_t)r * PLANE + w0];
                } else {
                    const uint32_t j = jhi | (uint32_t)r;
                    own[r] = bra[base | s_loc_off[j & 127] | s_loc_off[128 + (j >> 7)]];
                }
            }
            // optional (QFE_PREFETCH_DIST=1; off by default: with the request issued ahead of the barrier the L2 prefetch costs more issue
            // slots than it saves latency, 1.104 s against 1.117 s at 28 qubits): the lines of the next tile -> L2 behind the own rows
            {
                if (a.prefetch_dist > 1 && next_id < a.n_tiles) {
                    const uint64_t nb = base_of(next_id) ^ a.ket_xor;
#pragma unroll
                    for (uint32_t idx = 8u * tid; idx < (uint32_t)(R * NT); idx += 8u * NT)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ket + (nb | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)])));
                }
            }
            if (PROF && (tid & 31) == 0) {
                t_gs = clock64();
                if (have_prev) s_prof[(tid >> 5) * PROF_PHASES + PH_STAGE] += (unsigned long long)(t_gs - t_last);
            }
            real_t acc_re[R], acc_im[R];  // APPLY only: unused here
            const Terms<real_t, false> terms{a, (uint32_t)tile_id};
#pragma unroll 1
            for (int ri = 0; ri < nR; ++ri) {
                const uint2 d = s_rng[ri];
                const uint32_t skipbit = d.y >> RNG_SKIP_SHIFT;
                const bool half = a.same_tile && skipbit != 0;
                // the row blocks (warps) of the other polarity hold the partners of these pairs
                if (half && (((jhi >> skipbit) ^ (d.y >> RNG_POL_SHIFT)) & 1u)) continue;
                const int g0 = (int)(d.x & 0xffffu), g1 = (int)(d.x >> 16), t0 = (int)(d.y & 0xffffu);
                const uint32_t kind = (d.y >> RNG_KIND_SHIFT) & 3u;
                if (kind == KIND_FAST)
                    process_range<real_t, false, MODE_EXPECT, KIND_FAST, PLANE, RE_ONLY>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
                else if (kind == KIND_RUNS)
                    process_range<real_t, false, MODE_EXPECT, KIND_RUNS, PLANE, RE_ONLY>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
                else
                    process_range<real_t, false, MODE_EXPECT, KIND_RUNS | KIND_XROW, PLANE, RE_ONLY>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, half, own, acc_re, acc_im, e_re, e_im);
            }
        }
        if (PROF && (tid & 31) == 0) {
            t_end = clock64();
            s_prof[(tid >> 5) * PROF_PHASES + PH_GROUPS] += (unsigned long long)(t_end - t_gs);
            s_prof[(tid >> 5) * PROF_PHASES + PH_ASYNC_TILES - 1] += 1ull;
            s_prof[(tid >> 5) * PROF_PHASES + PH_ASYNC_TILES] += 1ull;  // marks this kernel in the counters
            s_tend[tid >> 5] = t_end;
            have_prev = true;
        }
        // the next tile's amplitudes (the CTA's last tile asks for its own again: unconditional, so the values stay in registers)
        request(next_id < a.n_tiles ? next_id : tile_id);
    }

    double2 tot = block_reduce_sum(make_double2(e_re, e_im), s_red);
    if (tid == 0) a.partials[blockIdx.x] = tot;
    if (PROF && a.prof) {
        __syncthreads();
        if (tid < NW * PROF_PHASES) atomicAdd(a.prof + tid, s_prof[tid]);
    }
}

// ---- EXPECT kernel, second generation: control words in the constant bank ----------------------------------------------------------
// Same tile layout, tile switch and arithmetic as tile_expect_kernel.  What changes is where a warp waits (ncu source view of the
// kernels above, profiles/r02_expect_28q_source_notes.txt: the shared-memory pipe is the binding resource and runs at 77 % on a large
// sweep, 63 % over the step; a LDS issued while it is busy returns after 300-1000 cycles, and the group loop took three such round
// trips in a row -- range descriptor, group header, partner rows -- besides re-reading %tid and the shared window base per group):
//   * range descriptors and group headers travel in the kernel parameter next to the term records (one record stream: header, terms,
//     header, terms ...), so every control word is a constant-bank load with a warp-uniform index, not a shared-memory round trip;
//   * the first eight partner rows of a group are requested BEFORE its term sum is formed (constant loads, popc, adds: no shared
//     memory), the other eight while the first are consumed;
//   * the term sums of KIND_RUNS groups are straight-line code for the common run lengths (a group of a weight-4 X mask is one run);
//   * %tid and the shared window base are pinned in registers (opaque to the compiler, which otherwise re-reads the special
//     registers inside the loops), partner rows are ld.shared at [register + immediate].
constexpr int E2_MAX_RANGES = 48;
constexpr int E2_MAX_REC = 2000;  // group headers + staged terms of one sweep

struct Expect2Args {
    const void* ket;
    const void* bra;
    double2* partials;
    const uint64_t* nat_off;  // PLAN_TABLE entries each, see plan_host.h
    const uint64_t* loc_off;
    const uint32_t* nat_loc;
    unsigned long long tile_begin, n_tiles;  // the launch covers tiles [tile_begin, n_tiles)
    unsigned long long tile_stride;          // gridDim.x
    unsigned long long ket_xor;
    unsigned long long* prof;  // development: per-warp phase cycle counters (QFE_PROF=1), else null
    int n_ranges, n_runs, same_tile, n_rec;
    uint8_t run_pos[PLAN_MAX_RUNS];
    uint8_t run_len[PLAN_MAX_RUNS];
    // range ri: .x = record index of its first group header | number of groups << 16; .y = kind / polarity / skip bit as in r_desc
    uint2 rng[E2_MAX_RANGES];
    // group: one header record (.x = header word a of plan_host.h) followed by its staged terms (.x/.y = coefficient bits: double, or
    // float in .x for complex64 states; .z = t_zl, .w = t_zoc)
    uint4 rec[E2_MAX_REC];
};
static_assert(sizeof(Expect2Args) <= 32764, "kernel parameter space");

// shared-memory accesses to the tile by 32-bit shared address.  Every access to the tile in the kernels that use these is one of these
// volatile statements: they keep their order among themselves and against the barriers; no "memory" clobber, so that the table lookups
// (plain loads from regions that are never written after the prologue) can move across them.
template <typename real_t> struct Smem;
template <> struct Smem<double> {
    static __device__ __forceinline__ double2 ld(const uint32_t addr) {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void st(const uint32_t addr, const double2 v) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y));
    }
};
template <> struct Smem<float> {
    static __device__ __forceinline__ float2 ld(const uint32_t addr) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void st(const uint32_t addr, const float2 v) {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y));
    }
};

template <typename real_t> __device__ __forceinline__ real_t rec_coeff(const uint4& q);
template <> __device__ __forceinline__ double rec_coeff<double>(const uint4& q) { return __hiloint2double((int)q.y, (int)q.x); }
template <> __device__ __forceinline__ float rec_coeff<float>(const uint4& q) { return __uint_as_float(q.x); }

// c_t (-1)^{|z_t & b|} of staged term record i for the thread's rows (jhi) on tile `tile`
template <typename real_t>
__device__ __forceinline__ real_t e2_term(const Expect2Args& a, const int i, const uint32_t jhi, const uint32_t tile) {
    const uint4 q = a.rec[i];
    return sign_by_parity(rec_coeff<real_t>(q), (uint32_t)__popc((q.z & jhi) ^ (q.w & tile)));
}
template <typename real_t, int N>
__device__ __forceinline__ real_t e2_sum_fixed(const Expect2Args& a, const int t, const uint32_t jhi, const uint32_t tile) {
    real_t c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) c[i] = e2_term<real_t>(a, t + i, jhi, tile);  // all loads before the first use
    real_t S0 = c[0], S1 = N > 1 ? c[1] : (real_t)0;
#pragma unroll
    for (int i = 2; i < N; ++i) {
        if (i & 1)
            S1 += c[i];
        else
            S0 += c[i];
    }
    return S0 + S1;
}
// signed sum of the n staged terms from record t on
template <typename real_t>
__device__ __forceinline__ real_t e2_sum(const Expect2Args& a, int t, const int n, const uint32_t jhi, const uint32_t tile) {
    // compare-and-branch in the order of frequency (weight-4 X masks carry 4, then 6 terms); a jump table costs an LDC + BRX round trip
    if (n == 4) return e2_sum_fixed<real_t, 4>(a, t, jhi, tile);
    if (n == 6) return e2_sum_fixed<real_t, 6>(a, t, jhi, tile);
    if (n == 2) return e2_sum_fixed<real_t, 2>(a, t, jhi, tile);
    if (n == 5) return e2_sum_fixed<real_t, 5>(a, t, jhi, tile);
    if (n == 3) return e2_sum_fixed<real_t, 3>(a, t, jhi, tile);
    if (n == 1) return e2_sum_fixed<real_t, 1>(a, t, jhi, tile);
    real_t S0 = (real_t)0, S1 = (real_t)0;
    const int te = t + n;
#pragma unroll 1
    for (; t + 4 <= te; t += 4) {
        const real_t c0 = e2_term<real_t>(a, t, jhi, tile), c1 = e2_term<real_t>(a, t + 1, jhi, tile);
        const real_t c2 = e2_term<real_t>(a, t + 2, jhi, tile), c3 = e2_term<real_t>(a, t + 3, jhi, tile);
        S0 += c0 + c2;
        S1 += c1 + c3;
    }
#pragma unroll 1
    for (; t < te; ++t) S0 += e2_term<real_t>(a, t, jhi, tile);
    return S0 + S1;
}

// All groups of one range for the thread's 16 rows.  rp: record index of the first group header; sbase: shared address of the tile.
// AHEAD: partner rows requested before the term sum is formed; every row that is consumed requests the row AHEAD further on.
template <typename real_t, int KIND, bool HALF, bool RE_ONLY, int AHEAD = R / 2>
__device__ __forceinline__ void e2_range(const Expect2Args& a, int rp, const int ng, const uint32_t sbase, const uint32_t tid, const uint32_t jhi,
                                         const uint32_t tile_no, const typename CplxOf<real_t>::type (&own)[R], double& e_re, double& e_im) {
    using cplx_t = typename CplxOf<real_t>::type;
    constexpr uint32_t ROWB = GeoExpect<real_t>::PLANE * (uint32_t)sizeof(cplx_t);  // bytes between the planes of two register rows
    constexpr bool XROW = (KIND & KIND_XROW) != 0;
    uint32_t hdr_next = a.rec[rp].x;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        const uint32_t hdr = hdr_next;
        const int nt = (int)((hdr >> HDR_NT_SHIFT) & HDR_NT_MASK);
        const int t0 = rp + 1;
        rp = t0 + nt;
        hdr_next = a.rec[rp].x;  // the record after the last group of a sweep exists (zero): no bound to check
        const bool imag = hdr & HDR_IMAG;
        const uint32_t xr = hdr & (uint32_t)(R - 1);
        // the partner thread's word of plane 0: thread index ^ thread bits of the X mask
        const uint32_t paddr = sbase + (tid ^ ((hdr & HDR_XL_MASK) >> RB)) * (uint32_t)sizeof(cplx_t);
        auto row_addr = [&](const int r) { return XROW ? paddr + (((uint32_t)r ^ xr) * ROWB) : paddr + (uint32_t)r * ROWB; };
        cplx_t p[R];
#pragma unroll
        for (int r = 0; r < AHEAD; ++r) p[r] = Smem<real_t>::ld(row_addr(r));  // travel while the term sum is formed
        if (KIND == KIND_FAST) {
            const real_t S = e2_sum<real_t>(a, t0, nt, jhi, tile_no);
            if (HALF || RE_ONLY) {
                // HALF: the pair (b, b^x) sums to 2 S (a a' + b b').  RE_ONLY: Re(Wc conj(own) p) = S Re(.) or, Wc = i S, -S Im(.)
                const bool im = !HALF && imag;
                real_t a0 = (real_t)0, a1 = (real_t)0, a2 = (real_t)0, a3 = (real_t)0;
                if (!im) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r + AHEAD < R) p[r + AHEAD] = Smem<real_t>::ld(row_addr(r + AHEAD));
                        if (r & 1) {
                            a2 = fma(own[r].x, p[r].x, a2);
                            a3 = fma(own[r].y, p[r].y, a3);
                        } else {
                            a0 = fma(own[r].x, p[r].x, a0);
                            a1 = fma(own[r].y, p[r].y, a1);
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r + AHEAD < R) p[r + AHEAD] = Smem<real_t>::ld(row_addr(r + AHEAD));
                        if (r & 1) {
                            a2 = fma(own[r].x, p[r].y, a2);
                            a3 = fma(-own[r].y, p[r].x, a3);
                        } else {
                            a0 = fma(own[r].x, p[r].y, a0);
                            a1 = fma(-own[r].y, p[r].x, a1);
                        }
                    }
                }
                const real_t s = (a0 + a1) + (a2 + a3);
                if (HALF)
                    e_re = fma((double)(S + S), (double)s, e_re);
                else
                    e_re = fma((double)(im ? -S : S), (double)s, e_re);
            } else {
                real_t a0 = (real_t)0, a1 = (real_t)0, b0 = (real_t)0, b1 = (real_t)0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r + AHEAD < R) p[r + AHEAD] = Smem<real_t>::ld(row_addr(r + AHEAD));
                    a0 = fma(own[r].x, p[r].x, a0);  // Re conj(own) p = a a' + b b'
                    a1 = fma(own[r].y, p[r].y, a1);
                    b0 = fma(own[r].x, p[r].y, b0);  // Im conj(own) p = a b' - b a'
                    b1 = fma(-own[r].y, p[r].x, b1);
                }
                real_t sr = a0 + a1, si = b0 + b1;
                times_weight<real_t, false>(sr, si, S, (real_t)0, imag);
                e_re += (double)sr;
                e_im += (double)si;
            }
        } else if (HALF || RE_ONLY) {
            // one real number per row, folded with the sign pattern of every run
            const bool im = !HALF && imag;
            int t = t0;
            const int te = t0 + nt;
            uint32_t head = a.rec[t].z;
            int len = (int)(head >> TZ_RUN_SHIFT);
            real_t S = e2_sum<real_t>(a, t, len, jhi, tile_no);  // first run: formed while the rows travel
            real_t f[R];
            if (!im) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r + AHEAD < R) p[r + AHEAD] = Smem<real_t>::ld(row_addr(r + AHEAD));
                    f[r] = fma(own[r].x, p[r].x, own[r].y * p[r].y);
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r + AHEAD < R) p[r + AHEAD] = Smem<real_t>::ld(row_addr(r + AHEAD));
                    f[r] = fma(own[r].y, p[r].x, -own[r].x * p[r].y);  // -Im conj(own) p
                }
            }
            real_t tot = S * fold_rows<real_t, R>(f, (head >> TZ_V_SHIFT) & 15u);
            t += len;
#pragma unroll 1
            while (t < te) {
                head = a.rec[t].z;
                len = (int)(head >> TZ_RUN_SHIFT);
                S = e2_sum<real_t>(a, t, len, jhi, tile_no);
                t += len;
                tot = fma(S, fold_rows<real_t, R>(f, (head >> TZ_V_SHIFT) & 15u), tot);
            }
            e_re += HALF ? (double)(tot + tot) : (double)tot;
        } else {
            // full complex value: rows in two chunks of 8 (keeps 16 products, not 32, live); the term sums of a run are formed per chunk
            real_t tot_re = (real_t)0, tot_im = (real_t)0;
#pragma unroll
            for (int r = AHEAD; r < R / 2; ++r) p[r] = Smem<real_t>::ld(row_addr(r));  // the first chunk is rows 0..7
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                real_t fr[8], fi[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = 8 * ch + q;
                    fr[q] = fma(own[r].x, p[q].x, own[r].y * p[q].y);
                    fi[q] = fma(own[r].x, p[q].y, -own[r].y * p[q].x);
                }
                if (ch == 0) {  // the rows of the second chunk travel during the run loop of the first
#pragma unroll
                    for (int q = 0; q < 8; ++q) p[q] = Smem<real_t>::ld(row_addr(8 + q));
                }
                int t = t0;
                const int te = t0 + nt;
#pragma unroll 1
                while (t < te) {
                    const uint32_t head = a.rec[t].z;
                    const uint32_t v = (head >> TZ_V_SHIFT) & 15u;
                    const int len = (int)(head >> TZ_RUN_SHIFT);
                    const real_t S = e2_sum<real_t>(a, t, len, jhi, tile_no);
                    t += len;
                    real_t sr = fold_rows<real_t, 8>(fr, v), si = fold_rows<real_t, 8>(fi, v);
                    if (ch) {
                        const real_t s3 = sigma<real_t>(v, 3);
                        sr *= s3;
                        si *= s3;
                    }
                    times_weight<real_t, false>(sr, si, S, (real_t)0, imag);
                    tot_re += sr;
                    tot_im += si;
                }
            }
            e_re += (double)tot_re;
            e_im += (double)tot_im;
        }
    }
}

template <typename real_t, bool PROF, bool RE_ONLY>
__global__ void __launch_bounds__(NT_MAX, 1) tile_expect2_kernel(const __grid_constant__ Expect2Args a) {
    using G = GeoExpect<real_t>;
    constexpr int NT = NT_MAX, NW = G::NW;
    constexpr uint32_t PLANE = G::PLANE;
    using cplx_t = typename CplxOf<real_t>::type;
    constexpr uint32_t ROWB = PLANE * (uint32_t)sizeof(cplx_t);
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* s_nat_off = reinterpret_cast<uint64_t*>(smem + G::SM_NATOFF);
    uint64_t* s_loc_off = reinterpret_cast<uint64_t*>(smem + G::SM_LOCOFF);
    uint64_t* s_base = reinterpret_cast<uint64_t*>(smem + G::SM_BASE);
    uint32_t* s_nat_loc = reinterpret_cast<uint32_t*>(smem + G::SM_NATLOC);
    double2* s_red = reinterpret_cast<double2*>(smem + G::SM_RED);
    unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(smem + G::SM_PROF);
    long long* s_tend = reinterpret_cast<long long*>(smem + G::SM_TEND);  // development counters, as in tile_sweep_kernel
    long long t_gs = 0ll, t_end = 0ll, t_last = 0ll;
    bool have_prev = false;

    // %tid and the shared window base, opaque to the compiler: it keeps them in registers instead of re-reading the special registers
    uint32_t tid, sbase = (uint32_t)__cvta_generic_to_shared(smem + G::SM_TILE);
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    asm volatile("" : "+r"(sbase));
    const int nR = a.n_ranges;
    const cplx_t* __restrict__ ket = reinterpret_cast<const cplx_t*>(a.ket);
    const cplx_t* __restrict__ bra = reinterpret_cast<const cplx_t*>(a.bra);

    for (int i = (int)tid; i < PLAN_TABLE; i += NT) {
        s_nat_off[i] = a.nat_off[i];
        s_loc_off[i] = a.loc_off[i];
        s_nat_loc[i] = a.nat_loc[i];
    }
    {
        // entry (k, v): the 7 bits v of the tile number at bit 7k, deposited into the runs of the outer mask
        unsigned long long rest = (unsigned long long)(tid & 127u) << (7 * (tid >> 7));
        uint64_t b = 0;
        for (int r = 0; r < a.n_runs; ++r) {
            const int len = a.run_len[r];
            b |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
            rest >>= len;
        }
        s_base[tid] = b;
    }
    if (PROF && tid < NW * PROF_PHASES) s_prof[tid] = 0ull;
    __syncthreads();

    const uint32_t jhi = tid << RB;
    const uint32_t paddr0 = sbase + tid * (uint32_t)sizeof(cplx_t);  // the thread's word of plane 0
    double e_re = 0.0, e_im = 0.0;
    auto base_of = [&](const unsigned long long t) {  // tile numbers stay below 2^28 (shards of at most 2^40 amplitudes, plan_create)
        return s_base[t & 127] | s_base[128 + ((t >> 7) & 127)] | s_base[256 + ((t >> 14) & 127)] | s_base[384 + ((t >> 21) & 127)];
    };
    // the thread's 16 amplitudes of a tile in natural order (idx = tid + u * NT: consecutive threads = consecutive amplitudes of a
    // 128 B segment); the low 7 index bits are the thread's own
    const uint64_t off_lo = s_nat_off[tid & 127u];
    const uint32_t loc_lo = s_nat_loc[tid & 127u];
    cplx_t v[R];
    uint32_t vw[R];  // shared addresses of their words in the tile (the same for every tile; looked up per tile rather than held across the group loops)
    auto request = [&](const unsigned long long t) {
        const uint64_t kb = (base_of(t) ^ a.ket_xor) | off_lo;
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = ld_stream(ket + (kb | s_nat_off[128 + ((tid + u * NT) >> 7)]));
#pragma unroll
        for (int u = 0; u < R; ++u) vw[u] = sbase + tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + u * NT) >> 7)]) * (uint32_t)sizeof(cplx_t);
    };

    unsigned long long tile_id = a.tile_begin + blockIdx.x;
    if (tile_id < a.n_tiles) request(tile_id);
    for (; tile_id < a.n_tiles; tile_id += a.tile_stride) {
        const unsigned long long next_id = tile_id + a.tile_stride;
        __syncthreads();  // every partner row of the previous tile has been read
        if (PROF && have_prev && (tid & 31u) == 0) {
            t_last = s_tend[0];
            for (int w = 1; w < NW; ++w) t_last = max(t_last, s_tend[w]);
            s_prof[(tid >> 5) * PROF_PHASES + PH_SKEW] += (unsigned long long)(t_last - t_end);
        }
#pragma unroll
        for (int u = 0; u < R; ++u) Smem<real_t>::st(vw[u], v[u]);
        if (PROF && have_prev && (tid & 31u) == 0) {  // [0]: last warp done -> this warp's amplitudes of the next tile have arrived and are stored
            t_gs = clock64();
            s_prof[(tid >> 5) * PROF_PHASES + PH_WAIT] += (unsigned long long)(t_gs - t_last);
        }
        __syncthreads();
        if (PROF && have_prev && (tid & 31u) == 0) {  // [1]: -> every warp has stored (barrier)
            const long long t_b2 = clock64();
            s_prof[(tid >> 5) * PROF_PHASES + PH_ISSUE] += (unsigned long long)(t_b2 - t_gs);
        }
        {
            cplx_t own[R];
            if (a.same_tile) {
#pragma unroll
                for (int r = 0; r < R; ++r) own[r] = Smem<real_t>::ld(paddr0 + (uint32_t)r * ROWB);
            } else {
                const uint64_t base = base_of(tile_id);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t j = jhi | (uint32_t)r;
                    own[r] = bra[base | s_loc_off[j & 127u] | s_loc_off[128 + (j >> 7)]];
                }
            }
            if (PROF && (tid & 31u) == 0) {
                t_gs = clock64();
                if (have_prev) s_prof[(tid >> 5) * PROF_PHASES + PH_STAGE] += (unsigned long long)(t_gs - t_last);
            }
            const uint32_t tile_no = (uint32_t)tile_id;
#pragma unroll 1
            for (int ri = 0; ri < nR; ++ri) {
                const uint2 d = a.rng[ri];
                const uint32_t skipbit = d.y >> RNG_SKIP_SHIFT;
                const bool half = a.same_tile && skipbit != 0;
                // the row blocks (warps) of the other polarity hold the partners of these pairs
                if (half && (((jhi >> skipbit) ^ (d.y >> RNG_POL_SHIFT)) & 1u)) continue;
                const int rp = (int)(d.x & 0xffffu), ng = (int)(d.x >> 16);
                const uint32_t kind = (d.y >> RNG_KIND_SHIFT) & 3u;
                if (half) {
                    if (kind == KIND_FAST)
                        e2_range<real_t, KIND_FAST, true, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                    else if (kind == KIND_RUNS)
                        e2_range<real_t, KIND_RUNS, true, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                    else
                        e2_range<real_t, KIND_RUNS | KIND_XROW, true, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                } else {
                    if (kind == KIND_FAST)
                        e2_range<real_t, KIND_FAST, false, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                    else if (kind == KIND_RUNS)
                        e2_range<real_t, KIND_RUNS, false, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                    else
                        e2_range<real_t, KIND_RUNS | KIND_XROW, false, RE_ONLY>(a, rp, ng, sbase, tid, jhi, tile_no, own, e_re, e_im);
                }
            }
        }
        if (PROF && (tid & 31u) == 0) {
            t_end = clock64();
            s_prof[(tid >> 5) * PROF_PHASES + PH_GROUPS] += (unsigned long long)(t_end - t_gs);
            s_prof[(tid >> 5) * PROF_PHASES + PH_ASYNC_TILES - 1] += 1ull;
            s_prof[(tid >> 5) * PROF_PHASES + PH_ASYNC_TILES] += 1ull;  // marks the pipelined kernels in the counters
            s_tend[tid >> 5] = t_end;
            have_prev = true;
        }
        // the next tile's amplitudes (the CTA's last tile asks for its own again: unconditional, so the values stay in registers)
        request(next_id < a.n_tiles ? next_id : tile_id);
    }

    double2 tot = block_reduce_sum(make_double2(e_re, e_im), s_red);
    if (tid == 0) a.partials[blockIdx.x] = tot;
    if (PROF && a.prof) {
        __syncthreads();
        if (tid < NW * PROF_PHASES) atomicAdd(a.prof + tid, s_prof[tid]);
    }
}

// ---- pipelined APPLY kernel --------------------------------------------------------------------------------------------------------
// y (+)= H|ket> for real-coefficient sweeps over full tiles: the group loops and the write-back through the tile of tile_sweep_kernel
// <APPLY>, with the tile switch of tile_expect_kernel -- the next ket tile is requested as soon as the accumulators have left the
// registers (before the read-modify-write of y, which then overlaps those loads), tile offsets come from tables.
template <typename real_t>
__global__ void __launch_bounds__(NT_MAX, 1) tile_apply_kernel(const __grid_constant__ SweepArgs a) {
    using G = GeoExpect<real_t>;
    constexpr int NT = NT_MAX;
    constexpr uint32_t PLANE = G::PLANE;
    using cplx_t = typename CplxOf<real_t>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    cplx_t* const tile = reinterpret_cast<cplx_t*>(smem + G::SM_TILE);
    uint2* s_hdr = reinterpret_cast<uint2*>(smem + G::SM_HDR);
    uint2* s_rng = reinterpret_cast<uint2*>(smem + G::SM_RNG);
    uint64_t* s_nat_off = reinterpret_cast<uint64_t*>(smem + G::SM_NATOFF);
    uint64_t* s_base = reinterpret_cast<uint64_t*>(smem + G::SM_BASE);
    uint32_t* s_nat_loc = reinterpret_cast<uint32_t*>(smem + G::SM_NATLOC);

    const int nG = a.n_groups, nR = a.n_ranges;
    const int tid = threadIdx.x;
    const cplx_t* __restrict__ ket = reinterpret_cast<const cplx_t*>(a.ket);
    cplx_t* __restrict__ yout = reinterpret_cast<cplx_t*>(a.out_vec);

    for (int g = tid; g < nG; g += NT) s_hdr[g] = make_uint2(a.g_hdr[2 * g], a.g_hdr[2 * g + 1]);
    for (int g = tid; g < nR; g += NT) s_rng[g] = make_uint2(a.r_desc[2 * g], a.r_desc[2 * g + 1]);
    for (int i = tid; i < PLAN_TABLE; i += NT) {
        s_nat_off[i] = a.nat_off[i];
        s_nat_loc[i] = a.nat_loc[i];
    }
    {
        unsigned long long rest = (unsigned long long)(tid & 127) << (7 * (tid >> 7));
        uint64_t b = 0;
        for (int r = 0; r < a.n_runs; ++r) {
            const int len = a.run_len[r];
            b |= (rest & ((1ull << len) - 1ull)) << a.run_pos[r];
            rest >>= len;
        }
        s_base[tid] = b;
    }
    __syncthreads();

    const uint32_t jhi = (uint32_t)tid << RB, w0 = (uint32_t)tid;
    auto base_of = [&](const unsigned long long t) {
        return s_base[t & 127] | s_base[128 + ((t >> 7) & 127)] | s_base[256 + ((t >> 14) & 127)] | s_base[384 + ((t >> 21) & 127)];
    };
    const uint64_t off_lo = s_nat_off[tid & 127];
    const uint32_t loc_lo = s_nat_loc[tid & 127];
    cplx_t v[R];
    auto request = [&](const unsigned long long t) {
        const uint64_t kb = (base_of(t) ^ a.ket_xor) | off_lo;
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = ld_stream(ket + (kb | s_nat_off[128 + ((tid + u * NT) >> 7)]));
    };

    unsigned long long tile_id = a.tile_begin + blockIdx.x;
    if (tile_id < a.n_tiles) request(tile_id);
    for (; tile_id < a.n_tiles; tile_id += a.tile_stride) {
        const unsigned long long next_id = tile_id + a.tile_stride;
        const uint64_t ybase = base_of(tile_id) | off_lo;
        __syncthreads();  // the write-back of the previous tile has read the tile
#pragma unroll
        for (int u = 0; u < R; ++u) tile[tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + u * NT) >> 7)])] = v[u];
        __syncthreads();
        // y += ...: the tile's lines of y travel to L2 while the groups are processed, for the read-modify-write that ends the tile
        if (a.accumulate && a.prefetch_dist > 0)
#pragma unroll
            for (uint32_t idx = 8u * tid; idx < (uint32_t)(R * NT); idx += 8u * NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(yout + (base_of(tile_id) | s_nat_off[idx & 127] | s_nat_off[128 + (idx >> 7)])));
        {
            real_t acc_re[R], acc_im[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc_re[r] = acc_im[r] = (real_t)0;
            cplx_t own[R];  // EXPECT / BATCH only: unused here
            double e_re = 0.0, e_im = 0.0;
            const Terms<real_t, false> terms{a, (uint32_t)tile_id};
#pragma unroll 1
            for (int ri = 0; ri < nR; ++ri) {
                const uint2 d = s_rng[ri];
                const int g0 = (int)(d.x & 0xffffu), g1 = (int)(d.x >> 16), t0 = (int)(d.y & 0xffffu);
                const uint32_t kind = (d.y >> RNG_KIND_SHIFT) & 3u;
                if (kind == KIND_FAST)
                    process_range<real_t, false, MODE_APPLY, KIND_FAST, PLANE>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, false, own, acc_re, acc_im, e_re, e_im);
                else if (kind == KIND_RUNS)
                    process_range<real_t, false, MODE_APPLY, KIND_RUNS, PLANE>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, false, own, acc_re, acc_im, e_re, e_im);
                else
                    process_range<real_t, false, MODE_APPLY, KIND_RUNS | KIND_XROW, PLANE>(tile, terms, s_hdr, nullptr, g0, g1, t0, w0, jhi, false, own, acc_re, acc_im, e_re, e_im);
            }
            __syncthreads();  // every partner row has been read: the result leaves through the tile
#pragma unroll
            for (int r = 0; r < R; ++r) {
                cplx_t o;
                o.x = acc_re[r];
                o.y = acc_im[r];
                tile[(uint32_t)r * PLANE + w0] = o;
            }
        }
        // the accumulators are in the tile: their registers take the next ket tile, which travels during the write-back
        request(next_id < a.n_tiles ? next_id : tile_id);
        __syncthreads();
        // natural order: y is read and written in 128 B contiguous segments, 8 independent loads in flight per thread
#pragma unroll
        for (int h = 0; h < R; h += R / 2) {
            cplx_t yv[R / 2];
            if (a.accumulate) {
#pragma unroll
                for (int u = 0; u < R / 2; ++u) yv[u] = yout[ybase | s_nat_off[128 + ((tid + (h + u) * NT) >> 7)]];
            }
#pragma unroll
            for (int u = 0; u < R / 2; ++u) {
                cplx_t o = tile[tile_word<PLANE>(loc_lo | s_nat_loc[128 + ((tid + (h + u) * NT) >> 7)])];
                if (a.accumulate) {
                    o.x += yv[u].x;
                    o.y += yv[u].y;
                }
                yout[ybase | s_nat_off[128 + ((tid + (h + u) * NT) >> 7)]] = o;
            }
        }
    }
}

// ---- direct kernel: one thread per row, partners straight from global memory ------------------
// Used for shards smaller than a tile (n_local < 8) and as an independent cross-check of the tile
// kernel in the tests (QFE_FORCE_DIRECT=1).
struct DirectArgs {
    const void* ket;
    const void* bra;
    void* out_vec;
    double2* partials;
    const uint64_t* x;
    const uint64_t* z;
    const double* cre;
    const double* cim;
    const int64_t* range_begin;  // BATCH: per-group term ranges (device), indexed from range0
    int64_t tb, te;              // EXPECT / APPLY: the class's term range
    int range0;
    int n_local;
    int accumulate;
};

template <typename real_t, int MODE>
__global__ void __launch_bounds__(256) direct_kernel(const DirectArgs a) {
    using cplx_t = typename CplxOf<real_t>::type;
    __shared__ double2 s_red[8];
    const cplx_t* __restrict__ ket = reinterpret_cast<const cplx_t*>(a.ket);
    const cplx_t* __restrict__ bra = reinterpret_cast<const cplx_t*>(a.bra);
    cplx_t* __restrict__ yout = reinterpret_cast<cplx_t*>(a.out_vec);
    int64_t tb = a.tb, te = a.te;
    if (MODE == MODE_BATCH) {
        tb = a.range_begin[a.range0 + blockIdx.y];
        te = a.range_begin[a.range0 + blockIdx.y + 1];
    }
    const uint64_t n_amps = 1ull << a.n_local;
    double e_re = 0.0, e_im = 0.0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_amps; b += (uint64_t)gridDim.x * blockDim.x) {
        double acc_re = 0.0, acc_im = 0.0;
        for (int64_t t = tb; t < te; ++t) {
            const cplx_t p = ket[b ^ a.x[t]];
            double cr = a.cre[t], ci = a.cim[t];
            if (__popcll(a.z[t] & b) & 1) {
                cr = -cr;
                ci = -ci;
            }
            acc_re += cr * (double)p.x - ci * (double)p.y;
            acc_im += cr * (double)p.y + ci * (double)p.x;
        }
        if (MODE == MODE_APPLY) {
            cplx_t y;
            if (a.accumulate) {
                y = yout[b];
                y.x += (real_t)acc_re;
                y.y += (real_t)acc_im;
            } else {
                y.x = (real_t)acc_re;
                y.y = (real_t)acc_im;
            }
            yout[b] = y;
        } else {
            const cplx_t v = bra[b];
            e_re += (double)v.x * acc_re + (double)v.y * acc_im;
            e_im += (double)v.x * acc_im - (double)v.y * acc_re;
        }
    }
    if (MODE != MODE_APPLY) {
        double2 tot = block_reduce_sum(make_double2(e_re, e_im), s_red);
        if (threadIdx.x == 0) a.partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// vals[r] = sum_b partials[r * n_blocks + b], one warp per range, fixed order
__global__ void __launch_bounds__(128) reduce_partials(const double2* __restrict__ partials, int n_ranges, int n_blocks, double2* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_ranges) return;
    double sx = 0.0, sy = 0.0;
    for (int b = lane; b < n_blocks; b += 32) {
        const double2 p = partials[(size_t)r * n_blocks + b];
        sx += p.x;
        sy += p.y;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sx += __shfl_down_sync(0xffffffffu, sx, d);
        sy += __shfl_down_sync(0xffffffffu, sy, d);
    }
    if (lane == 0) vals[r] = make_double2(sx, sy);
}

// y (+)= sum_p slab[p], fixed order: ends an APPLY class whose tiles were shared by several CTAs
template <typename cplx_t>
__global__ void __launch_bounds__(256) combine_slabs(const cplx_t* __restrict__ slabs, int n_parts, int64_t n_amps, cplx_t* __restrict__ y, int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += (int64_t)gridDim.x * blockDim.x) {
        cplx_t v;
        if (accumulate) {
            v = y[i];
        } else {
            v.x = 0;
            v.y = 0;
        }
        for (int p = 0; p < n_parts; ++p) {
            const cplx_t w = slabs[(size_t)p * n_amps + i];
            v.x += w.x;
            v.y += w.y;
        }
        y[i] = v;
    }
}

int reduce_to_vals(qfe_ctx* ctx, const double2* partials, int n_ranges, int n_blocks, double2* vals) {
    reduce_partials<<<(n_ranges + 3) / 4, 128, 0, ctx->stream>>>(partials, n_ranges, n_blocks, vals);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

}  // namespace qfe

using namespace qfe;

struct qfe_plan {
    PlanHost h;
    PlanLimits lim;
    bool batch = false;
    bool force_direct = false;
    bool hermitian = false;
    int device = 0;
    DevBuf g_hdr, r_desc, t_zl, t_zoc, t_cre, t_cim, d_x, d_z, d_cre, d_cim, dg_begin, nat_off, loc_off, nat_loc;
};

namespace qfe {

template <typename T>
static int upload(DevBuf& buf, const std::vector<T>& v, cudaStream_t s) {
    QFE_TRY(buf.alloc(sizeof(T) * (v.empty() ? 1 : v.size())));
    if (!v.empty()) QFE_CUDA(cudaMemcpyAsync(buf.p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
    return QFE_OK;
}

static int find_class(const qfe_plan* p, int x_hi) {
    for (size_t i = 0; i < p->h.classes.size(); ++i)
        if (p->h.classes[i].x_hi == x_hi) return (int)i;
    return -1;
}

// CTA size and grid of a sweep: 2^13-amplitude tiles = one CTA of 512 threads per SM, smaller tiles = two CTAs of 256 threads
static int sweep_threads(int k) { return k >= PLAN_MAX_TILE_BITS ? NT_MAX : NT_MAX / 2; }
static int64_t sweep_capacity(const qfe_ctx* ctx, const SweepHost& sw) { return (int64_t)ctx->sm_count * (NT_MAX / sweep_threads(sw.k)); }
// A shard with fewer tiles than the GPU has CTA slots: up to SWEEP_MAX_SPLIT CTAs stage the same tile and share the groups of each of
// its ranges (a tile takes ~20 us to work through alone, whatever the register size).  QFE_SPLIT=1 turns this off (development knob).
constexpr int SWEEP_MAX_SPLIT = 16;
static int sweep_split(const qfe_ctx* ctx, const SweepHost& sw, bool any_group_count = false) {
    static const int cap_env = [] {
        const char* e = getenv("QFE_SPLIT");
        return e ? std::max(1, atoi(e)) : SWEEP_MAX_SPLIT;
    }();
    const int64_t cap = sweep_capacity(ctx, sw);
    if (sw.n_tiles * 2 > cap || (!any_group_count && sw.group_end - sw.group_begin < 2)) return 1;
    return (int)std::min<int64_t>(cap / sw.n_tiles, std::min(cap_env, SWEEP_MAX_SPLIT));
}
static int sweep_grid(const qfe_ctx* ctx, const SweepHost& sw, int split) {
    return split > 1 ? (int)(sw.n_tiles * split) : (int)std::min<int64_t>(sw.n_tiles, sweep_capacity(ctx, sw));
}

template <typename real_t, bool CPLX, int MODE, int NTH>
static int launch_tile_n(qfe_ctx* ctx, SweepArgs& args, int grid) {
    auto kern = tile_sweep_kernel<real_t, CPLX, MODE, NTH>;
    constexpr size_t smem = Geo<NTH>::SM_TOTAL;
    static int attr_device = -1;  // once per instantiation (and again if another device's context comes by)
    if (attr_device != ctx->device) {
        QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = ctx->device;
    }
    args.tile_stride = (unsigned long long)(grid / std::max(args.split, 1));
    kern<<<grid, NTH, smem, ctx->stream>>>(args);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

// QFE_PIPELINED=0 keeps real-coefficient EXPECT sweeps on the generic kernel; QFE_PROF=1 runs the instrumented instances
static bool env_flag(const char* name, bool dflt) {
    const char* e = getenv(name);
    return e ? (e[0] != '0') : dflt;
}
static bool prof_enabled() {
    static const bool on = env_flag("QFE_PROF", false);
    return on;
}

template <typename real_t, bool PROF, bool RE_ONLY>
static int launch_expect_pipelined_p(qfe_ctx* ctx, SweepArgs& args, int grid) {
    auto kern = tile_expect_kernel<real_t, PROF, RE_ONLY>;
    constexpr size_t smem = GeoExpect<real_t>::SM_TOTAL;
    static int attr_device = -1;
    if (attr_device != ctx->device) {
        QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = ctx->device;
    }
    args.tile_stride = (unsigned long long)grid;
    kern<<<grid, NT_MAX, smem, ctx->stream>>>(args);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

static int prof_buffer(qfe_ctx* ctx, SweepArgs& args) {
    if (!ctx->prof.p) {
        QFE_TRY(ctx->prof.alloc(sizeof(unsigned long long) * (NT_MAX / 32) * PROF_PHASES));
        QFE_CUDA(cudaMemsetAsync(ctx->prof.p, 0, ctx->prof.bytes, ctx->stream));
    }
    args.prof = ctx->prof.as<unsigned long long>();
    return QFE_OK;
}

// real-coefficient EXPECT sweep over full tiles, one CTA per tile slot
static bool pipelined_eligible(const SweepArgs& args, bool cplx) {
    static const bool on = env_flag("QFE_PIPELINED", true);
    return on && !cplx && args.k == PLAN_MAX_TILE_BITS && args.split == 1;
}
static int launch_expect_pipelined(qfe_ctx* ctx, SweepArgs& args, int grid, int dtype, bool real_only) {
    if (prof_enabled()) {
        QFE_TRY(prof_buffer(ctx, args));
        return dtype == QFE_C128 ? launch_expect_pipelined_p<double, true, false>(ctx, args, grid) : launch_expect_pipelined_p<float, true, false>(ctx, args, grid);
    }
    if (real_only && !args.same_tile)  // (on the state's own tile the Hermitian half visit already forms the real part only)
        return dtype == QFE_C128 ? launch_expect_pipelined_p<double, false, true>(ctx, args, grid) : launch_expect_pipelined_p<float, false, true>(ctx, args, grid);
    return dtype == QFE_C128 ? launch_expect_pipelined_p<double, false, false>(ctx, args, grid) : launch_expect_pipelined_p<float, false, false>(ctx, args, grid);
}

// second-generation EXPECT kernel (control words in the constant bank); QFE_EXPECT2=0 keeps the first generation
static bool expect2_eligible(const SweepArgs& args) {
    static const bool on = env_flag("QFE_EXPECT2", true);
    return on && args.n_ranges <= E2_MAX_RANGES && args.n_terms + args.n_groups + 1 <= E2_MAX_REC;
}
static void fill_expect2_args(const qfe_plan* p, const SweepHost& s, const SweepArgs& a, int dtype, Expect2Args& b) {
    b.ket = a.ket;
    b.bra = a.bra;
    b.partials = a.partials;
    b.nat_off = a.nat_off;
    b.loc_off = a.loc_off;
    b.nat_loc = a.nat_loc;
    b.tile_begin = a.tile_begin;
    b.n_tiles = a.n_tiles;
    b.ket_xor = a.ket_xor;
    b.prof = a.prof;
    b.n_ranges = a.n_ranges;
    b.n_runs = a.n_runs;
    b.same_tile = a.same_tile;
    memcpy(b.run_pos, a.run_pos, sizeof(b.run_pos));
    memcpy(b.run_len, a.run_len, sizeof(b.run_len));
    // the record stream (plan_host.cpp: build_sweep_records, interpreted on the CPU by tests/csrc/plan_emul.cpp) goes into the parameter
    static thread_local SweepRecords recs;
    build_sweep_records(p->h, s, dtype != QFE_C128, recs);
    memcpy(b.rng, recs.rng.data(), recs.rng.size() * sizeof(uint32_t));
    memcpy(b.rec, recs.rec.data(), recs.rec.size() * sizeof(uint32_t));
    b.n_rec = (int)(recs.rec.size() / 4) - 1;
}
template <typename real_t, bool PROF, bool RE_ONLY>
static int launch_expect2_p(qfe_ctx* ctx, Expect2Args& args, int grid) {
    auto kern = tile_expect2_kernel<real_t, PROF, RE_ONLY>;
    constexpr size_t smem = GeoExpect<real_t>::SM_TOTAL;
    static int attr_device = -1;
    if (attr_device != ctx->device) {
        QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = ctx->device;
    }
    args.tile_stride = (unsigned long long)grid;
    kern<<<grid, NT_MAX, smem, ctx->stream>>>(args);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}
static int launch_expect2(qfe_ctx* ctx, const qfe_plan* p, const SweepHost& sw, SweepArgs& a, int grid, int dtype, bool real_only) {
    if (prof_enabled()) QFE_TRY(prof_buffer(ctx, a));
    static thread_local Expect2Args b;  // 32 KB: filled per launch, copied into the launch by the runtime
    fill_expect2_args(p, sw, a, dtype, b);
    if (prof_enabled()) return dtype == QFE_C128 ? launch_expect2_p<double, true, false>(ctx, b, grid) : launch_expect2_p<float, true, false>(ctx, b, grid);
    if (real_only && !a.same_tile)  // (on the state's own tile the Hermitian half visit already forms the real part only)
        return dtype == QFE_C128 ? launch_expect2_p<double, false, true>(ctx, b, grid) : launch_expect2_p<float, false, true>(ctx, b, grid);
    return dtype == QFE_C128 ? launch_expect2_p<double, false, false>(ctx, b, grid) : launch_expect2_p<float, false, false>(ctx, b, grid);
}

template <typename real_t>
static int launch_apply_pipelined_t(qfe_ctx* ctx, SweepArgs& args, int grid) {
    auto kern = tile_apply_kernel<real_t>;
    constexpr size_t smem = GeoExpect<real_t>::SM_TOTAL;
    static int attr_device = -1;
    if (attr_device != ctx->device) {
        QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_device = ctx->device;
    }
    args.tile_stride = (unsigned long long)grid;
    kern<<<grid, NT_MAX, smem, ctx->stream>>>(args);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}
static int launch_apply_pipelined(qfe_ctx* ctx, SweepArgs& args, int grid, int dtype) {
    return dtype == QFE_C128 ? launch_apply_pipelined_t<double>(ctx, args, grid) : launch_apply_pipelined_t<float>(ctx, args, grid);
}
template <typename real_t, bool CPLX, int MODE>
static int launch_tile(qfe_ctx* ctx, SweepArgs& args, int grid) {
    if constexpr (MODE == MODE_EXPECT && !CPLX && std::is_same<real_t, double>::value) {
        if (prof_enabled() && sweep_threads(args.k) == NT_MAX) {  // instrumented instance of the synchronously staged kernel
            QFE_TRY(prof_buffer(ctx, args));
            auto kern = tile_sweep_kernel<double, false, MODE_EXPECT, NT_MAX, true>;
            constexpr size_t smem = Geo<NT_MAX>::SM_TOTAL;
            static int attr_device = -1;
            if (attr_device != ctx->device) {
                QFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_device = ctx->device;
            }
            args.tile_stride = (unsigned long long)(grid / std::max(args.split, 1));
            kern<<<grid, NT_MAX, smem, ctx->stream>>>(args);
            ctx->launches++;
            QFE_CUDA(cudaGetLastError());
            return QFE_OK;
        }
    }
    return sweep_threads(args.k) == NT_MAX ? launch_tile_n<real_t, CPLX, MODE, NT_MAX>(ctx, args, grid) : launch_tile_n<real_t, CPLX, MODE, NT_MAX / 2>(ctx, args, grid);
}

template <typename real_t, int MODE>
static int launch_tile_c(qfe_ctx* ctx, SweepArgs& args, int grid, bool cplx) {
    return cplx ? launch_tile<real_t, true, MODE>(ctx, args, grid) : launch_tile<real_t, false, MODE>(ctx, args, grid);
}

template <int MODE>
static int launch_tile_d(qfe_ctx* ctx, SweepArgs& args, int grid, bool cplx, int dtype) {
    return dtype == QFE_C128 ? launch_tile_c<double, MODE>(ctx, args, grid, cplx) : launch_tile_c<float, MODE>(ctx, args, grid, cplx);
}

template <int MODE>
static int launch_direct(qfe_ctx* ctx, const DirectArgs& args, dim3 grid, int dtype) {
    if (dtype == QFE_C128)
        direct_kernel<double, MODE><<<grid, 256, 0, ctx->stream>>>(args);
    else
        direct_kernel<float, MODE><<<grid, 256, 0, ctx->stream>>>(args);
    ctx->launches++;
    QFE_CUDA(cudaGetLastError());
    return QFE_OK;
}

static void fill_sweep_args(const qfe_plan* p, const SweepHost& s, int dtype, SweepArgs& a) {
    a.g_hdr = p->g_hdr.as<uint32_t>() + 2 * (size_t)s.group_begin;
    a.r_desc = p->r_desc.as<uint32_t>() + 2 * (size_t)s.range_begin;
    a.n_ranges = s.range_end - s.range_begin;
    a.t_zl = p->t_zl.as<uint32_t>() + s.term_begin;
    a.t_zoc = p->t_zoc.as<uint32_t>() + s.term_begin;
    a.t_cre = p->t_cre.as<double>() + s.term_begin;
    a.t_cim = p->t_cim.as<double>() + s.term_begin;
    a.n_groups = s.group_end - s.group_begin;
    a.n_terms = (int)(s.term_end - s.term_begin);
    a.k = s.k;
    static const int prefetch_dist = [] {
        const char* pd = getenv("QFE_PREFETCH_DIST");  // development knob (2: also in the pipelined EXPECT kernel)
        return pd ? atoi(pd) : 1;
    }();
    a.prefetch_dist = prefetch_dist;
    static const int stagger_span = [] {
        const char* e = getenv("QFE_STAGGER");  // development knob: cycles over which the first tiles of the CTAs are spread
        return e ? atoi(e) : 0;
    }();
    a.stagger = (stagger_span > 0 && s.n_tiles >= 148 * 16) ? stagger_span / 148 : 0;
    a.split = 1;
    a.n_runs = s.n_runs;
    a.tile_begin = 0;
    a.n_tiles = (unsigned long long)s.n_tiles;
    a.ket_xor = s.ket_xor;
    memcpy(a.run_pos, s.run_pos, sizeof(a.run_pos));
    memcpy(a.run_len, s.run_len, sizeof(a.run_len));
    const size_t si = (size_t)(&s - p->h.sweeps.data());
    if (!s.complex_w) {  // real-coefficient sweep: the staged terms travel in the kernel parameter
        const PlanHost& h = p->h;
        for (int t = 0; t < a.n_terms; ++t) {
            const int64_t i = s.term_begin + t;
            TermConst& q = a.tc[t];
            if (dtype == QFE_C128) {
                uint64_t bits;
                memcpy(&bits, &h.t_cre[i], 8);
                q.c_lo = (uint32_t)bits;
                q.c_hi = (uint32_t)(bits >> 32);
            } else {
                const float f = (float)h.t_cre[i];
                memcpy(&q.c_lo, &f, 4);
                q.c_hi = 0;
            }
            q.zp = h.t_zl[i];
            q.zoc = h.t_zoc[i];
        }
    }
    a.nat_off = p->nat_off.as<uint64_t>() + si * PLAN_TABLE;
    a.loc_off = p->loc_off.as<uint64_t>() + si * PLAN_TABLE;
    a.nat_loc = p->nat_loc.as<uint32_t>() + si * PLAN_TABLE;
}

// Runs every sweep of one class in EXPECT or BATCH mode and leaves per-range values in host_vals
// (EXPECT: one value per sweep; BATCH: one per group, sweep order).
// part / n_parts: only the sweeps si with (si - first sweep of the class) % n_parts == part run (the direct kernel, which has no
// sweeps, runs in part 0 only): the parts of a class add up to the whole class.
// Upload pipeline of qfe_expectation_host: the state arrives in `chunks` equal pieces (its top qubits select the piece); a sweep
// whose tile leaves those qubits alone can run on a piece as soon as it has landed, so the first sweeps hide the host->device copy.
struct UploadPipe {
    int chunks = 1;
    int max_early = 0;                   // sweeps that run piece by piece
    std::vector<cudaEvent_t> ready;      // ready[c]: piece c is in device memory
    int (*issue_copy)(void*, int) = nullptr;  // enqueues the copy of piece c on the copy stream and records ready[c]
    void* cookie = nullptr;
};

static int run_class_reduce(qfe_ctx* ctx, const qfe_plan* p, int ci, const void* bra, const void* ket, int dtype, int part, int n_parts,
                            std::vector<double>& host_vals, std::vector<int32_t>& val_slot, const UploadPipe* pipe = nullptr, bool real_only = false) {
    const ClassHost& cls = p->h.classes[ci];
    const bool batch = p->batch;
    const bool tiles = p->h.use_tiles && !p->force_direct;
    cudaStream_t s = ctx->stream;
    host_vals.clear();
    val_slot.clear();
    int64_t n_vals = 0;
    auto mine = [&](int si) { return (si - cls.sweep_begin) % n_parts == part; };
    // sweeps that run piece by piece while the state is uploaded (EXPECT on the state itself only)
    std::vector<char> early(cls.sweep_end - cls.sweep_begin, 0);
    int n_early = 0;
    if (pipe && tiles && !batch && bra == ket && pipe->chunks > 1) {
        int lc = 0;
        while ((1 << lc) < pipe->chunks) ++lc;
        const uint64_t top = ((uint64_t(1) << lc) - 1) << (p->h.n_local - lc);
        for (int si = cls.sweep_begin; si < cls.sweep_end && n_early < pipe->max_early; ++si) {
            const SweepHost& sw = p->h.sweeps[si];
            if (mine(si) && sw.ket_xor == 0 && (sw.tile_mask & top) == 0 && sw.n_tiles >= (int64_t)pipe->chunks * ctx->sm_count) {
                early[si - cls.sweep_begin] = 1;
                ++n_early;
            }
        }
    }
    if (tiles) {
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si)
            if (mine(si)) n_vals += batch ? (p->h.sweeps[si].group_end - p->h.sweeps[si].group_begin) : (early[si - cls.sweep_begin] ? pipe->chunks : 1);
    } else {
        n_vals = part != 0 ? 0 : (batch ? (cls.dgroup_end - cls.dgroup_begin) : 1);
    }
    if (n_vals == 0) return QFE_OK;
    QFE_TRY(ctx->result.reserve(sizeof(double2) * n_vals));
    double2* d_vals = ctx->result.as<double2>();
    int64_t vpos = 0;
    if (tiles) {
        // every value (sweep, or group of a sweep in BATCH mode) owns one row of `row` per-CTA partials; rows of sweeps with a smaller
        // grid keep their zeros, and ONE fixed-order reduction ends the class
        size_t row = 1;
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si)
            if (mine(si)) row = std::max(row, (size_t)sweep_grid(ctx, p->h.sweeps[si], sweep_split(ctx, p->h.sweeps[si])));
        QFE_REQUIRE(n_vals <= INT32_MAX && row <= INT32_MAX, QFE_ERR_ARG, "too many values in one class");
        QFE_TRY(ctx->partials.reserve(sizeof(double2) * row * (size_t)n_vals));
        QFE_CUDA(cudaMemsetAsync(ctx->partials.p, 0, sizeof(double2) * row * (size_t)n_vals, s));
        auto launch = [&](int si, int64_t tile_begin, int64_t tile_end) -> int {
            const SweepHost& sw = p->h.sweeps[si];
            SweepArgs a;
            memset(&a, 0, sizeof(a));
            fill_sweep_args(p, sw, dtype, a);
            a.ket = ket;
            a.bra = bra;
            a.same_tile = (bra == ket && sw.ket_xor == 0) ? 1 : 0;
            a.split = sweep_split(ctx, sw);
            a.pstride = (int)row;
            a.tile_begin = (unsigned long long)tile_begin;
            a.n_tiles = (unsigned long long)tile_end;
            const int grid = tile_end - tile_begin == sw.n_tiles ? sweep_grid(ctx, sw, a.split)
                                                                 : (int)std::min<int64_t>(tile_end - tile_begin, sweep_capacity(ctx, sw));
            const int n_ranges = batch ? a.n_groups : 1;
            a.partials = ctx->partials.as<double2>() + (size_t)vpos * row;
            if (batch)
                QFE_TRY(launch_tile_d<MODE_BATCH>(ctx, a, grid, sw.complex_w, dtype));
            else if (pipelined_eligible(a, sw.complex_w) && expect2_eligible(a) && (a.same_tile || real_only))  // (full complex visits: first generation)
                QFE_TRY(launch_expect2(ctx, p, sw, a, grid, dtype, real_only));
            else if (pipelined_eligible(a, sw.complex_w))
                QFE_TRY(launch_expect_pipelined(ctx, a, grid, dtype, real_only));
            else
                QFE_TRY(launch_tile_d<MODE_EXPECT>(ctx, a, grid, sw.complex_w, dtype));
            for (int g = 0; g < n_ranges; ++g) val_slot.push_back(batch ? p->h.g_slot[sw.group_begin + g] : 0);
            vpos += n_ranges;
            return QFE_OK;
        };
        if (pipe && pipe->chunks > 1) {
            for (int c = 0; c < pipe->chunks; ++c) {
                QFE_TRY(pipe->issue_copy(pipe->cookie, c));
                QFE_CUDA(cudaStreamWaitEvent(s, pipe->ready[c], 0));
                for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) {
                    if (!early[si - cls.sweep_begin]) continue;
                    const int64_t per = p->h.sweeps[si].n_tiles / pipe->chunks;  // the top qubits of the tile number select the piece
                    QFE_TRY(launch(si, per * c, per * (c + 1)));
                }
            }
        }
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si)
            if (mine(si) && !early[si - cls.sweep_begin]) QFE_TRY(launch(si, 0, p->h.sweeps[si].n_tiles));
        QFE_TRY(reduce_to_vals(ctx, ctx->partials.as<double2>(), (int)n_vals, (int)row, d_vals));
    } else {
        DirectArgs a;
        memset(&a, 0, sizeof(a));
        a.ket = ket;
        a.bra = bra;
        a.x = p->d_x.as<uint64_t>();
        a.z = p->d_z.as<uint64_t>();
        a.cre = p->d_cre.as<double>();
        a.cim = p->d_cim.as<double>();
        a.range_begin = p->dg_begin.as<int64_t>();
        a.range0 = cls.dgroup_begin;
        a.tb = cls.dterm_begin;
        a.te = cls.dterm_end;
        a.n_local = p->h.n_local;
        const int64_t n_amps = int64_t(1) << p->h.n_local;
        const int gx = (int)std::min<int64_t>((n_amps + 255) / 256, (int64_t)ctx->sm_count * 4);
        const int n_ranges = (int)n_vals;
        QFE_TRY(ctx->partials.reserve(sizeof(double2) * (size_t)n_ranges * gx));
        a.partials = ctx->partials.as<double2>();
        // gridDim.y is limited to 65535 ranges per launch
        for (int r0 = 0; r0 < n_ranges; r0 += 32768) {
            const int nr = std::min(32768, n_ranges - r0);
            DirectArgs b = a;
            b.range0 = cls.dgroup_begin + r0;
            b.partials = a.partials + (size_t)r0 * gx;
            if (batch)
                QFE_TRY(launch_direct<MODE_BATCH>(ctx, b, dim3(gx, nr), dtype));
            else
                QFE_TRY(launch_direct<MODE_EXPECT>(ctx, b, dim3(gx, 1), dtype));
        }
        reduce_partials<<<(n_ranges + 3) / 4, 128, 0, s>>>(a.partials, n_ranges, gx, d_vals);
        ctx->launches++;
        QFE_CUDA(cudaGetLastError());
        for (int g = 0; g < n_ranges; ++g) val_slot.push_back(batch ? p->h.dg_slot[cls.dgroup_begin + g] : 0);
        vpos = n_ranges;
    }
    host_vals.resize(2 * n_vals);
    QFE_CUDA(cudaMemcpyAsync(host_vals.data(), d_vals, sizeof(double2) * n_vals, cudaMemcpyDeviceToHost, s));
    QFE_CUDA(cudaStreamSynchronize(s));
    return QFE_OK;
}

}  // namespace qfe

extern "C" {

int qfe_plan_create(qfe_ctx* ctx, const qfe_terms* t, int n_local, int shard, int tile_bits, qfe_plan** out) {
    QFE_REQUIRE(ctx && t && out, QFE_ERR_ARG, "qfe_plan_create: null argument");
    QFE_REQUIRE(n_local >= 0 && n_local <= t->n, QFE_ERR_ARG, "qfe_plan_create: n_local=%d outside 0..%d", n_local, t->n);
    QFE_REQUIRE(n_local <= 40, QFE_ERR_UNSUPPORTED, "qfe_plan_create: shards above 2^40 amplitudes are not supported");
    QFE_CUDA(cudaSetDevice(ctx->device));
    qfe_plan* p = new (std::nothrow) qfe_plan();
    QFE_REQUIRE(p, QFE_ERR_OOM, "out of host memory");
    p->device = ctx->device;
    p->batch = t->n_slots > 1;
    const char* fd = getenv("QFE_FORCE_DIRECT");
    p->force_direct = (tile_bits < 0) || (fd && fd[0] == '1');
    PlanLimits lim;
    const char* tb_env = getenv("QFE_TILE_BITS");  // development: default tile size without touching the callers
    lim.tile_bits = tile_bits > 0 ? tile_bits : (tb_env && atoi(tb_env) > 0 ? atoi(tb_env) : 13);
    if (lim.tile_bits > 13) lim.tile_bits = 13;  // 2^13 complex128 amplitudes = 128 KB of the 227 KB shared memory
    lim.max_terms = sweep_max_terms(std::min(n_local, lim.tile_bits), p->batch);
    lim.max_groups = p->batch ? BATCH_MAX_GROUPS : SWEEP_MAX_GROUPS;
    if (!p->batch) {  // every real-coefficient sweep over full tiles fits the record stream of tile_expect2_kernel (headers + terms + one pad)
        lim.group_overhead = 1;
        lim.max_slots = E2_MAX_REC - 1;
    }
    p->lim = lim;
    p->hermitian = true;
    for (int64_t i = 0; i < t->T && p->hermitian; ++i) p->hermitian = t->c[2 * i + 1] == 0.0;
    for (int64_t i = 0; i < t->n_slots && p->hermitian; ++i) p->hermitian = t->constant[2 * i + 1] == 0.0;
    TermView v;
    v.n = t->n;
    v.T = t->T;
    v.x = t->x.data();
    v.z = t->z.data();
    v.c = t->c.data();
    v.owner = t->owner.data();
    v.n_slots = t->n_slots;
    v.constant = t->constant.data();
    const char* err = nullptr;
    if (build_plan(v, n_local, shard, lim, p->h, &err) != 0) {
        set_error("%s", err ? err : "build_plan failed");
        delete p;
        return QFE_ERR_ARG;
    }
    cudaStream_t s = ctx->stream;
    int rc = QFE_OK;
    if ((rc = upload(p->g_hdr, p->h.g_hdr, s)) || (rc = upload(p->r_desc, p->h.r_desc, s)) || (rc = upload(p->t_zl, p->h.t_zl, s)) || (rc = upload(p->t_zoc, p->h.t_zoc, s)) ||
        (rc = upload(p->t_cre, p->h.t_cre, s)) || (rc = upload(p->t_cim, p->h.t_cim, s)) || (rc = upload(p->d_x, p->h.d_x, s)) ||
        (rc = upload(p->d_z, p->h.d_z, s)) || (rc = upload(p->d_cre, p->h.d_cre, s)) || (rc = upload(p->d_cim, p->h.d_cim, s)) ||
        (rc = upload(p->dg_begin, p->h.dg_begin, s)) || (rc = upload(p->nat_off, p->h.nat_off, s)) || (rc = upload(p->loc_off, p->h.loc_off, s)) ||
        (rc = upload(p->nat_loc, p->h.nat_loc, s))) {
        delete p;
        return rc;
    }
    QFE_CUDA(cudaStreamSynchronize(s));
    *out = p;
    return QFE_OK;
}

int qfe_plan_info(const qfe_plan* p, int64_t* n_sweeps, int64_t* n_groups, int64_t* n_terms, int32_t* n_classes, int32_t* tile_bits) {
    QFE_REQUIRE(p, QFE_ERR_ARG, "null plan");
    const bool tiles = p->h.use_tiles && !p->force_direct;
    if (n_sweeps) *n_sweeps = tiles ? (int64_t)p->h.sweeps.size() : (int64_t)p->h.classes.size();
    if (n_groups) *n_groups = (int64_t)p->h.dg_slot.size();
    if (n_terms) *n_terms = (int64_t)p->h.d_x.size();
    if (n_classes) *n_classes = (int32_t)p->h.classes.size();
    if (tile_bits) *tile_bits = tiles ? p->h.k : 0;
    return QFE_OK;
}

int qfe_plan_hermitian(const qfe_plan* p, int* hermitian) {
    QFE_REQUIRE(p && hermitian, QFE_ERR_ARG, "null argument");
    *hermitian = p->hermitian ? 1 : 0;
    return QFE_OK;
}

int qfe_plan_work(const qfe_plan* p, int64_t* half_groups, int64_t* full_groups, int64_t* half_terms, int64_t* full_terms) {
    QFE_REQUIRE(p, QFE_ERR_ARG, "null plan");
    int64_t hg = 0, fg = 0, ht = 0, ft = 0;
    const PlanHost& h = p->h;
    for (const SweepHost& s : h.sweeps)
        for (int r = s.range_begin; r < s.range_end; ++r) {
            const uint32_t d0 = h.r_desc[2 * r], d1 = h.r_desc[2 * r + 1];
            const bool half = (d1 >> RNG_SKIP_SHIFT) != 0;
            for (uint32_t g = (d0 & 0xffffu); g < (d0 >> 16); ++g) {
                const int64_t nt = (h.g_hdr[2 * (s.group_begin + g)] >> HDR_NT_SHIFT) & HDR_NT_MASK;
                (half ? hg : fg) += 1;
                (half ? ht : ft) += nt;
            }
        }
    if (half_groups) *half_groups = hg;
    if (full_groups) *full_groups = fg;
    if (half_terms) *half_terms = ht;
    if (full_terms) *full_terms = ft;
    return QFE_OK;
}

int qfe_plan_geometry(const qfe_plan* p, int* nqbits, int* n_local) {
    QFE_REQUIRE(p, QFE_ERR_ARG, "null plan");
    if (nqbits) *nqbits = p->h.n;
    if (n_local) *n_local = p->h.n_local;
    return QFE_OK;
}

int qfe_plan_slots(const qfe_plan* p, int64_t* n_slots) {
    QFE_REQUIRE(p && n_slots, QFE_ERR_ARG, "null argument");
    *n_slots = p->h.n_slots;
    return QFE_OK;
}

int qfe_plan_classes(const qfe_plan* p, int32_t* x_hi, int32_t capacity) {
    QFE_REQUIRE(p && x_hi, QFE_ERR_ARG, "null argument");
    QFE_REQUIRE(capacity >= (int32_t)p->h.classes.size(), QFE_ERR_ARG, "qfe_plan_classes: capacity %d < %zu classes", capacity, p->h.classes.size());
    for (size_t i = 0; i < p->h.classes.size(); ++i) x_hi[i] = p->h.classes[i].x_hi;
    return QFE_OK;
}

int qfe_dev_phase_cycles(qfe_ctx* ctx, uint64_t* cycles, int reset) {
    QFE_REQUIRE(ctx && cycles, QFE_ERR_ARG, "qfe_dev_phase_cycles: null argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)(NT_MAX / 32) * PROF_PHASES;
    for (size_t i = 0; i < n; ++i) cycles[i] = 0;
    if (!ctx->prof.p) return QFE_OK;
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    QFE_CUDA(cudaMemcpy(cycles, ctx->prof.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) QFE_CUDA(cudaMemset(ctx->prof.p, 0, n * sizeof(uint64_t)));
    return QFE_OK;
}

int qfe_plan_destroy(qfe_plan* p) {
    if (p) cudaSetDevice(p->device);
    delete p;
    return QFE_OK;
}

int qfe_expectation_class(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, double* out) {
    return qfe_expectation_class_part(ctx, p, x_hi, bra, ket, dtype, 0, 1, out);
}

int qfe_expectation_class_part(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, int part, int n_parts, double* out) {
    // QFE_FORCE_REAL_ONLY=1 (development / tests): run the real-part-only kernels the sharded driver uses for its pair ranks, so that they can
    // be checked on one GPU; the imaginary part of the result is then not meaningful
    static const bool force_real = env_flag("QFE_FORCE_REAL_ONLY", false);
    return expectation_class_part_impl(ctx, p, x_hi, bra, ket, dtype, part, n_parts, force_real, out);
}

}  // extern "C"

namespace qfe {
// real_only: the caller uses the real part of the result only (imaginary parts come back as computed or as zero)
int expectation_class_part_impl(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, int part, int n_parts, bool real_only,
                                double* out) {
    QFE_REQUIRE(ctx && p && bra && ket && out, QFE_ERR_ARG, "qfe_expectation_class: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, QFE_ERR_ARG, "qfe_expectation_class_part: part %d of %d", part, n_parts);
    QFE_CUDA(cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < 2 * p->h.n_slots; ++i) out[i] = 0.0;
    const int ci = find_class(p, x_hi);
    if (ci < 0) return QFE_OK;  // no term flips these high qubits
    std::vector<double> vals;
    std::vector<int32_t> slots;
    QFE_TRY(run_class_reduce(ctx, p, ci, bra, ket, dtype, part, n_parts, vals, slots, nullptr, real_only));
    for (size_t i = 0; i < slots.size(); ++i) {
        out[2 * slots[i]] += vals[2 * i];
        out[2 * slots[i] + 1] += vals[2 * i + 1];
    }
    return QFE_OK;
}
}  // namespace qfe

extern "C" {

int qfe_expectation(qfe_ctx* ctx, const qfe_plan* p, const void* psi, int dtype, double out[2]) {
    QFE_REQUIRE(ctx && p && psi && out, QFE_ERR_ARG, "qfe_expectation: null argument");
    QFE_REQUIRE(p->h.n_local == p->h.n, QFE_ERR_STATE, "qfe_expectation needs an unsharded plan (n_local == nqbits); use qfe_expectation_class");
    QFE_REQUIRE(p->h.n_slots == 1, QFE_ERR_STATE, "qfe_expectation needs a single-operator term set");
    return qfe_expectation_class(ctx, p, 0, psi, psi, dtype, out);
}

namespace {
struct HostCopy {
    qfe_ctx* ctx;
    const char* src;
    char* dst;
    size_t piece;
    std::vector<cudaEvent_t>* ready;
};
int issue_piece(void* cookie, int c) {
    HostCopy* h = static_cast<HostCopy*>(cookie);
    // modest copies: the source may be pageable, and a piece must not wait behind the whole state
    const size_t sub = size_t(1) << 28;
    for (size_t off = 0; off < h->piece; off += sub) {
        const size_t nb = std::min(sub, h->piece - off);
        QFE_CUDA(cudaMemcpyAsync(h->dst + c * h->piece + off, h->src + c * h->piece + off, nb, cudaMemcpyHostToDevice, h->ctx->copy_stream));
    }
    QFE_CUDA(cudaEventRecord((*h->ready)[c], h->ctx->copy_stream));
    return QFE_OK;
}
}  // namespace

int qfe_expectation_class0_host(qfe_ctx* ctx, const qfe_plan* p, const void* psi_host, void* psi_dev, int dtype, double out[2]) {
    QFE_REQUIRE(ctx && p && psi_host && psi_dev && out, QFE_ERR_ARG, "qfe_expectation_class0_host: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(p->h.n_slots == 1, QFE_ERR_STATE, "qfe_expectation_class0_host needs a single-operator term set");
    QFE_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t(1) << p->h.n_local) * (dtype == QFE_C128 ? 16 : 8);
    // The shard is uploaded on a second stream in 16 pieces (selected by its top 4 qubits); the first sweeps whose tiles leave those
    // qubits alone follow the upload piece by piece, so that most of the copy hides behind them.  QFE_HOST_EARLY = number of such
    // sweeps (0: plain upload, then all sweeps).
    static const int early_env = [] {
        const char* e = getenv("QFE_HOST_EARLY");
        return e ? atoi(e) : 20;
    }();
    const int ci = find_class(p, 0);
    const bool tiles = p->h.use_tiles && !p->force_direct;
    const int chunks = 16;
    out[0] = out[1] = 0.0;
    if (early_env <= 0 || ci < 0 || !tiles || bytes < (size_t(1) << 30)) {
        const size_t sub = size_t(1) << 28;  // modest copies: the source may be pageable
        for (size_t off = 0; off < bytes; off += sub) {
            const size_t nb = std::min(sub, bytes - off);
            QFE_CUDA(cudaMemcpyAsync((char*)psi_dev + off, (const char*)psi_host + off, nb, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (ci < 0) {
            QFE_CUDA(cudaStreamSynchronize(ctx->stream));
            return QFE_OK;
        }
        return qfe_expectation_class(ctx, p, 0, psi_dev, psi_dev, dtype, out);
    }
    if (!ctx->copy_stream) QFE_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while ((int)ctx->copy_events.size() < chunks + 1) {
        cudaEvent_t e;
        QFE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->copy_events.push_back(e);
    }
    // the destination may still be read by work queued on the compute stream
    QFE_CUDA(cudaEventRecord(ctx->copy_events[chunks], ctx->stream));
    QFE_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[chunks], 0));
    std::vector<cudaEvent_t> ready(ctx->copy_events.begin(), ctx->copy_events.begin() + chunks);
    HostCopy hc{ctx, (const char*)psi_host, (char*)psi_dev, bytes / chunks, &ready};
    UploadPipe pipe;
    pipe.chunks = chunks;
    pipe.max_early = early_env;
    pipe.ready = ready;
    pipe.issue_copy = issue_piece;
    pipe.cookie = &hc;
    std::vector<double> vals;
    std::vector<int32_t> slots;
    QFE_TRY(run_class_reduce(ctx, p, ci, psi_dev, psi_dev, dtype, 0, 1, vals, slots, &pipe));
    for (size_t i = 0; i < slots.size(); ++i) {
        out[0] += vals[2 * i];
        out[1] += vals[2 * i + 1];
    }
    return QFE_OK;
}

int qfe_expectation_host(qfe_ctx* ctx, const qfe_plan* p, const void* psi_host, int dtype, double out[2]) {
    QFE_REQUIRE(ctx && p && psi_host && out, QFE_ERR_ARG, "qfe_expectation_host: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(p->h.n_local == p->h.n, QFE_ERR_STATE, "qfe_expectation_host needs an unsharded plan");
    QFE_CUDA(cudaSetDevice(ctx->device));
    QFE_TRY(ctx->hostvec.reserve((size_t(1) << p->h.n) * (dtype == QFE_C128 ? 16 : 8)));
    return qfe_expectation_class0_host(ctx, p, psi_host, ctx->hostvec.p, dtype, out);
}

int qfe_apply_class(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* ket, void* y, int dtype, int accumulate) {
    QFE_REQUIRE(ctx && p && ket && y, QFE_ERR_ARG, "qfe_apply_class: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_REQUIRE(p->h.n_slots == 1, QFE_ERR_STATE, "qfe_apply_class needs a single-operator term set");
    QFE_REQUIRE(ket != y, QFE_ERR_ARG, "qfe_apply_class: in-place apply is not supported");
    QFE_CUDA(cudaSetDevice(ctx->device));
    const int ci = find_class(p, x_hi);
    const size_t bytes = (size_t(1) << p->h.n_local) * (dtype == QFE_C128 ? 16 : 8);
    if (ci < 0) {
        if (!accumulate) QFE_CUDA(cudaMemsetAsync(y, 0, bytes, ctx->stream));
        return QFE_OK;
    }
    const ClassHost& cls = p->h.classes[ci];
    const bool tiles = p->h.use_tiles && !p->force_direct;
    if (tiles) {
        if (cls.sweep_begin == cls.sweep_end && !accumulate) QFE_CUDA(cudaMemsetAsync(y, 0, bytes, ctx->stream));
        // small shards: the CTAs that share a tile leave their partial results in slabs, summed into y in a fixed order at the end
        int split = cls.sweep_begin < cls.sweep_end ? SWEEP_MAX_SPLIT : 1;
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) split = std::min(split, sweep_split(ctx, p->h.sweeps[si], true));  // one split for the class
        if (split > 1) QFE_TRY(ctx->slabs.reserve(bytes * (size_t)split));
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) {
            const SweepHost& sw = p->h.sweeps[si];
            SweepArgs a;
            memset(&a, 0, sizeof(a));
            fill_sweep_args(p, sw, dtype, a);
            a.ket = ket;
            a.bra = ket;
            a.same_tile = 0;
            a.split = split;
            if (split > 1) {
                a.out_vec = ctx->slabs.p;
                a.slab_amps = 1ull << p->h.n_local;
                a.accumulate = si != cls.sweep_begin;
            } else {
                a.out_vec = y;
                a.accumulate = (si == cls.sweep_begin) ? (accumulate ? 1 : 0) : 1;
            }
            if (pipelined_eligible(a, sw.complex_w))
                QFE_TRY(launch_apply_pipelined(ctx, a, sweep_grid(ctx, sw, split), dtype));
            else
                QFE_TRY(launch_tile_d<MODE_APPLY>(ctx, a, sweep_grid(ctx, sw, split), sw.complex_w, dtype));
        }
        if (split > 1) {
            const int64_t n_amps = int64_t(1) << p->h.n_local;
            const int gx = (int)std::min<int64_t>((n_amps + 255) / 256, (int64_t)ctx->sm_count * 8);
            if (dtype == QFE_C128)
                combine_slabs<double2><<<gx, 256, 0, ctx->stream>>>(ctx->slabs.as<double2>(), split, n_amps, (double2*)y, accumulate ? 1 : 0);
            else
                combine_slabs<float2><<<gx, 256, 0, ctx->stream>>>(ctx->slabs.as<float2>(), split, n_amps, (float2*)y, accumulate ? 1 : 0);
            ctx->launches++;
            QFE_CUDA(cudaGetLastError());
        }
    } else {
        DirectArgs a;
        memset(&a, 0, sizeof(a));
        a.ket = ket;
        a.bra = ket;
        a.out_vec = y;
        a.x = p->d_x.as<uint64_t>();
        a.z = p->d_z.as<uint64_t>();
        a.cre = p->d_cre.as<double>();
        a.cim = p->d_cim.as<double>();
        a.tb = cls.dterm_begin;
        a.te = cls.dterm_end;
        a.n_local = p->h.n_local;
        a.accumulate = accumulate ? 1 : 0;
        const int64_t n_amps = int64_t(1) << p->h.n_local;
        const int gx = (int)std::min<int64_t>((n_amps + 255) / 256, (int64_t)ctx->sm_count * 4);
        QFE_TRY(launch_direct<MODE_APPLY>(ctx, a, dim3(gx, 1), dtype));
    }
    return QFE_OK;
}

int qfe_apply(qfe_ctx* ctx, const qfe_plan* p, const void* psi, void* y, int dtype) {
    QFE_REQUIRE(ctx && p && psi && y, QFE_ERR_ARG, "qfe_apply: null argument");
    QFE_REQUIRE(p->h.n_local == p->h.n, QFE_ERR_STATE, "qfe_apply needs an unsharded plan (n_local == nqbits); use qfe_apply_class");
    return qfe_apply_class(ctx, p, 0, psi, y, dtype, 0);
}

int qfe_commutator_gradients(qfe_ctx* ctx, const qfe_plan* h_plan, const qfe_plan* pool_plan, const void* psi, void* scratch, int dtype,
                             double* grads) {
    QFE_REQUIRE(ctx && h_plan && pool_plan && psi && scratch && grads, QFE_ERR_ARG, "qfe_commutator_gradients: null argument");
    QFE_REQUIRE(h_plan->h.n_local == h_plan->h.n && pool_plan->h.n_local == pool_plan->h.n && h_plan->h.n == pool_plan->h.n, QFE_ERR_STATE,
                "qfe_commutator_gradients needs unsharded plans on the same register");
    // phi = H psi, then g_k = 2 Im <phi| tau_k |psi>
    QFE_TRY(qfe_apply(ctx, h_plan, psi, scratch, dtype));
    std::vector<double> vals(2 * pool_plan->h.n_slots);
    QFE_TRY(qfe_expectation_class(ctx, pool_plan, 0, scratch, psi, dtype, vals.data()));
    for (int64_t k = 0; k < pool_plan->h.n_slots; ++k) grads[k] = 2.0 * vals[2 * k + 1];
    return QFE_OK;
}

}  // extern "C"
