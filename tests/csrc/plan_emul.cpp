// TEST INFRASTRUCTURE.  CPU interpreter of the sweep planner's output (myqlm_fermion_b200/csrc/plan_host.*):
// walks every sweep / tile / group / term of one x_hi class exactly as the tile kernel indexes them,
// in plain scalar complex128.  Lets the planner (coverage of the X masks by tile masks, tile geometry,
// mask compression, coefficient folding) be checked on a machine without a GPU.  Not shipped, not
// linked into libqfe.so.
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../myqlm_fermion_b200/csrc/plan_host.h"

using cplx = std::complex<double>;
using namespace qfe;

extern "C" int plan_emul_run(int n, int64_t T, const uint64_t* x, const uint64_t* z, const double* c, const int32_t* owner, int64_t n_slots,
                             const double* constant, int n_local, int shard, int tile_bits, int max_terms, int max_groups, int x_hi,
                             const double* bra_, const double* ket_, double* out_slots, double* y_, int64_t* stats /* sweeps, groups, terms, classes */,
                             int half_visit) {
    TermView v;
    v.n = n;
    v.T = T;
    v.x = x;
    v.z = z;
    v.c = c;
    v.owner = owner;
    v.n_slots = n_slots;
    v.constant = constant;
    PlanLimits lim;
    lim.tile_bits = tile_bits;
    lim.max_terms = max_terms;
    lim.max_groups = max_groups;
    PlanHost p;
    const char* err = nullptr;
    if (build_plan(v, n_local, shard, lim, p, &err) != 0) return 1;
    if (!p.use_tiles) return 2;
    const cplx* bra = reinterpret_cast<const cplx*>(bra_);
    const cplx* ket = reinterpret_cast<const cplx*>(ket_);
    cplx* y = reinterpret_cast<cplx*>(y_);
    cplx* out = reinterpret_cast<cplx*>(out_slots);
    stats[0] = stats[1] = stats[2] = 0;
    stats[3] = (int64_t)p.classes.size();
    // every group of every class must be covered by exactly one sweep
    {
        size_t ng_tiled = p.g_hdr.size() / 2, ng_direct = p.dg_slot.size();
        if (ng_tiled != ng_direct || p.t_zl.size() != p.d_x.size()) return 3;
    }
    for (const ClassHost& cls : p.classes) {
        if (cls.x_hi != x_hi) continue;
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) {
            const SweepHost& s = p.sweeps[si];
            stats[0]++;
            stats[1] += s.group_end - s.group_begin;
            stats[2] += s.term_end - s.term_begin;
            const uint32_t tile_amps = 1u << s.k;
            std::vector<cplx> tile(tile_amps);
            std::vector<char> seen;
            if (si == cls.sweep_begin) seen.assign(size_t(1) << n_local, 0);
            for (int64_t t = 0; t < s.n_tiles; ++t) {
                const uint64_t base = tile_base(s, (uint64_t)t);
                if (base & s.tile_mask) return 4;
                for (uint32_t j = 0; j < tile_amps; ++j) {
                    const uint64_t pos = tile_spread(s, j);
                    if (pos & ~s.tile_mask) return 5;
                    if (compress_to_tile(s, pos) != j) return 6;
                    tile[j] = ket[(base ^ s.ket_xor) | pos];
                    if (!seen.empty()) {
                        if (seen[base | pos]) return 7;  // tiles must partition the shard
                        seen[base | pos] = 1;
                    }
                }
                // <psi|G|psi> with bra == ket on the same tile may visit half the rows of a Hermitian group (HDR_HALF)
                const bool same = (bra == ket) && s.ket_xor == 0 && half_visit;
                for (uint32_t j = 0; j < tile_amps; ++j) {
                    const uint64_t row = base | tile_spread(s, j);
                    int64_t tb = s.term_begin;
                    cplx acc_row(0.0, 0.0);
                    for (int g = s.group_begin; g < s.group_end; ++g) {
                        const uint32_t ha = p.g_hdr[2 * g], hb = p.g_hdr[2 * g + 1];
                        const uint32_t xl = ha & HDR_XL_MASK;
                        const int nt = (int)((ha >> HDR_NT_SHIFT) & HDR_NT_MASK);
                        const bool imag = ha & HDR_IMAG;
                        const uint32_t skip = hb & HDR_SKIP_MASK;
                        if (((ha & HDR_HALF) != 0) != (skip != 0)) return 10;
                        if (skip && (skip != (1u << (31 - __builtin_clz(xl))) || skip < (1u << PLAN_U_SHIFT))) return 11;  // highest X bit, warp-uniform
                        cplx w(0.0, 0.0);
                        uint32_t zreg0 = 0;
                        bool single = true;
                        for (int64_t tt = tb; tt < tb + nt; ++tt) {
                            const uint32_t zl = p.t_zl[tt] & 0xffffu, pat = (p.t_zl[tt] >> 16) & 0xffu;
                            if (zl & 7u) return 12;  // register bits live in the pattern only
                            const uint32_t jr = j & 7u;
                            const int par = (__builtin_popcount(zl & j) + (int)((pat >> jr) & 1u) + __builtin_popcountll(p.t_zo[tt] & base)) & 1;
                            if (tt == tb)
                                zreg0 = pat;
                            else if (pat != zreg0)
                                single = false;
                            cplx cc = imag ? cplx(0.0, p.t_cre[tt]) : cplx(p.t_cre[tt], p.t_cim[tt]);
                            w += par ? -cc : cc;
                        }
                        if (single != ((ha & HDR_SINGLE) != 0)) return 13;
                        if (single && ((hb >> HDR_PAT_SHIFT) & 0xffu) != zreg0) return 14;
                        tb += nt;
                        const cplx contrib = w * tile[j ^ xl];
                        acc_row += contrib;
                        if (same && skip) {
                            if (!(j & skip)) out[p.g_slot[g]] += 2.0 * std::real(std::conj(bra[row]) * contrib);
                        } else {
                            out[p.g_slot[g]] += std::conj(bra[row]) * contrib;
                        }
                    }
                    if (tb != s.term_end) return 8;
                    y[row] += acc_row;
                }
            }
            if (!seen.empty())
                for (char f : seen)
                    if (!f) return 9;
        }
    }
    return 0;
}

// plan statistics only (no state vectors): sweeps, groups, terms, classes + per-sweep group counts
extern "C" int plan_emul_stats(int n, int64_t T, const uint64_t* x, const uint64_t* z, const double* c, const int32_t* owner, int64_t n_slots,
                               const double* constant, int n_local, int shard, int tile_bits, int max_terms, int max_groups, int64_t* stats,
                               int32_t* groups_per_sweep, int32_t capacity) {
    TermView v;
    v.n = n;
    v.T = T;
    v.x = x;
    v.z = z;
    v.c = c;
    v.owner = owner;
    v.n_slots = n_slots;
    v.constant = constant;
    PlanLimits lim;
    lim.tile_bits = tile_bits;
    lim.max_terms = max_terms;
    lim.max_groups = max_groups;
    PlanHost p;
    const char* err = nullptr;
    if (build_plan(v, n_local, shard, lim, p, &err) != 0) return 1;
    stats[0] = (int64_t)p.sweeps.size();
    stats[1] = (int64_t)p.g_hdr.size() / 2;
    stats[2] = (int64_t)p.t_zl.size();
    stats[3] = (int64_t)p.classes.size();
    for (size_t i = 0; i < p.sweeps.size() && (int32_t)i < capacity; ++i) groups_per_sweep[i] = p.sweeps[i].group_end - p.sweeps[i].group_begin;
    return 0;
}
