// TEST INFRASTRUCTURE.  CPU interpreter of the sweep planner's output (myqlm_fermion_b200/csrc/plan_host.*):
// walks every sweep / tile / group / term of one x_hi class exactly as the tile kernel indexes them,
// in plain scalar complex128.  Lets the planner (coverage of the X masks by tile masks, tile geometry,
// mask compression, coefficient folding) be checked on a machine without a GPU.  Not shipped, not
// linked into libqfe.so.
#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../myqlm_fermion_b200/csrc/plan_host.h"

using cplx = std::complex<double>;
using namespace qfe;

static int emul_split = 1;
// number of slices every range is walked in (the kernel's split of small shards); 1 = off
extern "C" void plan_emul_set_split(int split) { emul_split = split < 1 ? 1 : split; }

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
        if (ng_tiled != ng_direct || p.t_zl.size() < p.d_x.size() || p.t_zl.size() > p.d_x.size() + ng_direct) return 3;  // + even-length padding of KIND_FAST groups
    }
    for (const ClassHost& cls : p.classes) {
        if (cls.x_hi != x_hi) continue;
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) {
            const SweepHost& s = p.sweeps[si];
            stats[0]++;
            stats[1] += s.group_end - s.group_begin;
            stats[2] += s.term_end - s.term_begin;
            const uint32_t tile_amps = 1u << s.k;
            // staging tables: natural index (coalesced order) -> shard offset / tile-local index, tile-local index -> shard offset
            if (p.nat_off.size() != p.sweeps.size() * PLAN_TABLE || p.nat_loc.size() != p.nat_off.size() || p.loc_off.size() != p.nat_off.size()) return 22;
            {
                const uint64_t* no = p.nat_off.data() + (size_t)si * PLAN_TABLE;
                const uint32_t* nl = p.nat_loc.data() + (size_t)si * PLAN_TABLE;
                const uint64_t* lo = p.loc_off.data() + (size_t)si * PLAN_TABLE;
                if ((s.tile_mask & 7u) != 7u) return 23;  // 128 B segments: the three low bits are in every tile
                std::vector<char> seen_loc(tile_amps, 0);
                for (uint32_t nat = 0; nat < tile_amps; ++nat) {
                    const uint64_t off = no[nat & 127] | no[128 + (nat >> 7)];
                    const uint32_t loc = nl[nat & 127] | nl[128 + (nat >> 7)];
                    if ((off & 7u) != (nat & 7u) || (off & ~s.tile_mask)) return 24;
                    if (loc >= tile_amps || seen_loc[loc]) return 25;
                    seen_loc[loc] = 1;
                    if ((lo[loc & 127] | lo[128 + (loc >> 7)]) != off || tile_spread(s, loc) != off) return 26;
                }
            }
            std::vector<cplx> tile(tile_amps);
            std::vector<char> seen;
            if (si == cls.sweep_begin) seen.assign(size_t(1) << n_local, 0);
            for (int64_t t = 0; t < s.n_tiles; ++t) {
                const uint64_t base = tile_base(s, (uint64_t)t);
                if (base & s.tile_mask) return 4;
                for (int64_t tt = s.term_begin; tt < s.term_end; ++tt)  // packed outer z bits against the tile number
                    if ((__builtin_popcountll(p.t_zo[tt] & base) ^ __builtin_popcount(p.t_zoc[tt] & (uint32_t)t)) & 1) return 27;
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
                // <psi|G|psi> with bra == ket on the same tile may visit half the rows of a Hermitian range (skip bit)
                const bool same = (bra == ket) && s.ket_xor == 0 && half_visit;
                for (uint32_t j = 0; j < tile_amps; ++j) {
                    const uint64_t row = base | tile_spread(s, j);
                    cplx acc_row(0.0, 0.0);
                    int g_expect = 0;
                    int64_t t_expect = 0;
                    for (int ri = s.range_begin; ri < s.range_end; ++ri) {
                        const uint32_t d0 = p.r_desc[2 * ri], d1 = p.r_desc[2 * ri + 1];
                        const int g0 = (int)(d0 & 0xffffu), g1 = (int)(d0 >> 16);
                        const int kind = (int)((d1 >> RNG_KIND_SHIFT) & 3u);
                        const uint32_t pol = (d1 >> RNG_POL_SHIFT) & 1u;
                        const uint32_t skipbit = d1 >> RNG_SKIP_SHIFT;
                        if (g0 != g_expect || (int64_t)(d1 & 0xffffu) != t_expect || g1 <= g0) return 10;  // ranges tile the sweep's groups in order
                        if (skipbit && (skipbit < (uint32_t)PLAN_U_SHIFT || skipbit >= (uint32_t)s.k || s.ket_xor != 0)) return 11;
                        if (kind != KIND_FAST && kind != KIND_RUNS && kind != (KIND_RUNS | KIND_XROW)) return 12;
                        // the groups of a range in `emul_split` slices, each entered through the second header word of its first
                        // group, exactly as the CTAs that share a tile of a small shard do (emul_split = 1: the whole range)
                        int64_t tb_end = s.term_begin + t_expect;
                        for (int part = 0; part < emul_split; ++part) {
                        const int gs = g0 + (int)((uint32_t)(g1 - g0) * (uint32_t)part / (uint32_t)emul_split);
                        const int ge = g0 + (int)((uint32_t)(g1 - g0) * (uint32_t)(part + 1) / (uint32_t)emul_split);
                        if (gs == ge) continue;
                        int64_t tb = emul_split > 1 ? s.term_begin + (int64_t)p.g_hdr[2 * (s.group_begin + gs) + 1] : tb_end;
                        if (tb != tb_end) return 29;  // slices follow each other in the staged terms
                        for (int gl = gs; gl < ge; ++gl) {
                            const int g = s.group_begin + gl;
                            const uint32_t ha = p.g_hdr[2 * g];
                            const uint32_t xl = ha & HDR_XL_MASK;
                            const int nt = (int)((ha >> HDR_NT_SHIFT) & HDR_NT_MASK);
                            const bool imag = ha & HDR_IMAG;
                            if ((int64_t)p.g_hdr[2 * g + 1] != tb - s.term_begin) return 28;  // second word: first staged term of the group
                            if (skipbit && ((31 - __builtin_clz(xl)) != (int)skipbit || imag)) return 13;  // the skip bit is the top X bit
                            if (((kind & KIND_XROW) != 0) != ((xl & 15u) != 0)) return 21;  // X bits on register rows are flagged
                            cplx w(0.0, 0.0);
                            // walk the runs exactly as the kernel does: the first term of a run carries its length
                            int64_t tt = tb;
                            uint32_t last_v = 0;
                            bool first_run = true;
                            while (tt < tb + nt) {
                                const uint32_t head = p.t_zl[tt];
                                const uint32_t v = (head >> TZ_V_SHIFT) & 15u, len = head >> TZ_RUN_SHIFT;
                                if (len == 0 || tt + len > tb + nt) return 14;
                                if (!first_run && v <= last_v) return 17;  // runs ascend in v
                                if (kind == KIND_FAST && (v != 0 || len != (uint32_t)nt || (nt & 1))) return 15;
                                for (uint32_t q = 0; q < len; ++q) {
                                    const uint32_t w32 = p.t_zl[tt + q];
                                    if (((w32 >> TZ_V_SHIFT) & 15u) != v || (q > 0 && (w32 >> TZ_RUN_SHIFT) != 0)) return 16;
                                    if (w32 & 0xffffu & ~TZ_HI_MASK) return 18;  // register bits live in the v field only
                                    const int par = (__builtin_popcount(w32 & TZ_HI_MASK & j) + __builtin_popcount(v & j) + __builtin_popcountll(p.t_zo[tt + q] & base)) & 1;
                                    cplx cc = imag ? cplx(0.0, p.t_cre[tt + q]) : cplx(p.t_cre[tt + q], p.t_cim[tt + q]);
                                    w += par ? -cc : cc;
                                }
                                last_v = v;
                                first_run = false;
                                tt += len;
                            }
                            tb += nt;
                            const cplx contrib = w * tile[j ^ xl];
                            acc_row += contrib;
                            if (same && skipbit) {
                                if (((j >> skipbit) & 1u) == pol) out[p.g_slot[g]] += 2.0 * std::real(std::conj(bra[row]) * contrib);
                            } else {
                                out[p.g_slot[g]] += std::conj(bra[row]) * contrib;
                            }
                        }
                        tb_end = tb;
                        }
                        g_expect = g1;
                        t_expect = tb_end - s.term_begin;
                    }
                    if (g_expect != s.group_end - s.group_begin || s.term_begin + t_expect != s.term_end) return 8;
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

// per-group dump of a plan (development aid for tools/plan_stats.py): one CSV line per group
extern "C" int plan_emul_dump(int n, int64_t T, const uint64_t* x, const uint64_t* z, const double* c, const int32_t* owner, int64_t n_slots,
                              const double* constant, int n_local, int shard, int tile_bits, int max_terms, int max_groups, const char* path) {
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
    FILE* f = fopen(path, "w");
    if (!f) return 2;
    fprintf(f, "sweep,range,kind,skipbit,pol,group,nt,xword,imag,tile_mask,x_hi\n");
    for (size_t si = 0; si < p.sweeps.size(); ++si) {
        const SweepHost& s = p.sweeps[si];
        for (int r = s.range_begin; r < s.range_end; ++r) {
            const uint32_t d0 = p.r_desc[2 * r], d1 = p.r_desc[2 * r + 1];
            for (uint32_t g = (d0 & 0xffffu); g < (d0 >> 16); ++g) {
                const uint32_t h = p.g_hdr[2 * (s.group_begin + g)];
                fprintf(f, "%zu,%d,%u,%u,%u,%u,%u,%u,%u,%llu,%d\n", si, r - s.range_begin, (d1 >> RNG_KIND_SHIFT) & 3u, d1 >> RNG_SKIP_SHIFT, (d1 >> RNG_POL_SHIFT) & 1u, g,
                        (h >> HDR_NT_SHIFT) & HDR_NT_MASK, h & HDR_XL_MASK, (h & HDR_IMAG) ? 1u : 0u, (unsigned long long)s.tile_mask, (int)p.classes[(size_t)s.cls].x_hi);
            }
        }
    }
    fclose(f);
    return 0;
}

// <bra|H_class|ket> of one x_hi class from the RECORD STREAMS of its sweeps (build_sweep_records), interpreted the way
// tile_expect2_kernel walks them: range word -> (first record, group count, kind, polarity, skip bit); header record -> X mask, term
// count, imaginary flag; term records -> coefficient bits, t_zl (lane / row-block z bits, register-row bits v, run length) and the outer z
// bits packed against the tile number.  half_visit: Hermitian ranges count 2 Re over the rows whose skip bit equals the polarity.
// single_precision: the coefficients travel as floats.  Returns 0, or a positive code naming the violated property; -1 if a sweep of the
// class has complex coefficients or the register is split over operators (those do not take this path).
extern "C" int plan_emul_run_records(int n, int64_t T, const uint64_t* x, const uint64_t* z, const double* c, int n_local, int shard, int tile_bits,
                                     int max_terms, int max_groups, int max_slots, int x_hi, const double* bra_, const double* ket_, int half_visit,
                                     int single_precision, double* out2, int64_t* stats /* sweeps, max records of a sweep, max ranges of a sweep */) {
    std::vector<int32_t> owner((size_t)T, 0);
    const double zero[2] = {0.0, 0.0};
    TermView v;
    v.n = n;
    v.T = T;
    v.x = x;
    v.z = z;
    v.c = c;
    v.owner = owner.data();
    v.n_slots = 1;
    v.constant = zero;
    PlanLimits lim;
    lim.tile_bits = tile_bits;
    lim.max_terms = max_terms;
    lim.max_groups = max_groups;
    lim.group_overhead = max_slots > 0 ? 1 : 0;
    lim.max_slots = max_slots;
    PlanHost p;
    const char* err = nullptr;
    if (build_plan(v, n_local, shard, lim, p, &err) != 0) return 1;
    if (!p.use_tiles) return 2;
    const cplx* bra = reinterpret_cast<const cplx*>(bra_);
    const cplx* ket = reinterpret_cast<const cplx*>(ket_);
    cplx out(0.0, 0.0);
    stats[0] = stats[1] = stats[2] = 0;
    SweepRecords recs;
    for (const ClassHost& cls : p.classes) {
        if (cls.x_hi != x_hi) continue;
        for (int si = cls.sweep_begin; si < cls.sweep_end; ++si) {
            const SweepHost& s = p.sweeps[si];
            if (s.complex_w) return -1;
            build_sweep_records(p, s, single_precision != 0, recs);
            const int n_ranges = s.range_end - s.range_begin;
            const int64_t n_rec = (int64_t)recs.rec.size() / 4;
            stats[0]++;
            stats[1] = std::max<int64_t>(stats[1], n_rec);
            stats[2] = std::max<int64_t>(stats[2], n_ranges);
            if ((int64_t)recs.rng.size() != 2 * (int64_t)n_ranges) return 3;
            if (n_rec != (s.group_end - s.group_begin) + (s.term_end - s.term_begin) + 1) return 4;  // headers + terms + the closing record
            if (max_slots > 0 && n_rec - 1 > max_slots && s.group_end - s.group_begin > 1) return 5;  // the planner's bound
            for (int q = 0; q < 4; ++q)
                if (recs.rec[4 * (n_rec - 1) + q] != 0u) return 6;
            const uint32_t tile_amps = 1u << s.k;
            std::vector<cplx> tile(tile_amps);
            for (int64_t t = 0; t < s.n_tiles; ++t) {
                const uint64_t base = tile_base(s, (uint64_t)t);
                for (uint32_t j = 0; j < tile_amps; ++j) tile[j] = ket[(base ^ s.ket_xor) | tile_spread(s, j)];
                const bool same = (bra == ket) && s.ket_xor == 0;
                for (uint32_t j = 0; j < tile_amps; ++j) {
                    const cplx own = bra[base | tile_spread(s, j)];
                    const uint32_t jhi = j & ~15u, jr = j & 15u;
                    int64_t next_rec = 0;
                    for (int ri = 0; ri < n_ranges; ++ri) {
                        const uint32_t d0 = recs.rng[2 * ri], d1 = recs.rng[2 * ri + 1];
                        int64_t rp = d0 & 0xffffu;
                        const int ng = (int)(d0 >> 16);
                        if (rp != next_rec || (d1 & 0xffffu)) return 7;  // the ranges follow each other in the stream
                        const uint32_t skipbit = d1 >> RNG_SKIP_SHIFT, pol = (d1 >> RNG_POL_SHIFT) & 1u, kind = (d1 >> RNG_KIND_SHIFT) & 3u;
                        const bool half = same && half_visit && skipbit != 0;
                        for (int g = 0; g < ng; ++g) {
                            const uint32_t hdr = recs.rec[4 * rp];
                            if (recs.rec[4 * rp + 1] | recs.rec[4 * rp + 2] | recs.rec[4 * rp + 3]) return 8;
                            const uint32_t xl = hdr & HDR_XL_MASK;
                            const int nt = (int)((hdr >> HDR_NT_SHIFT) & HDR_NT_MASK);
                            const bool imag = hdr & HDR_IMAG;
                            if (rp + 1 + nt >= n_rec) return 9;
                            if (((kind & KIND_XROW) != 0) != ((xl & 15u) != 0)) return 10;
                            double w = 0.0;
                            int64_t tt = rp + 1;
                            while (tt < rp + 1 + nt) {  // run by run, as the kernel does for KIND_RUNS; KIND_FAST is one run with v = 0
                                const uint32_t head = recs.rec[4 * tt + 2];
                                const uint32_t vrow = (head >> TZ_V_SHIFT) & 15u, len = head >> TZ_RUN_SHIFT;
                                if (len == 0 || tt + len > rp + 1 + nt) return 11;
                                if (kind == KIND_FAST && (vrow != 0 || (nt & 1))) return 12;
                                double srun = 0.0;
                                for (uint32_t q = 0; q < len; ++q) {
                                    const uint32_t* r4 = &recs.rec[4 * (tt + q)];
                                    double coeff;
                                    if (single_precision) {
                                        float f;
                                        memcpy(&f, &r4[0], 4);
                                        coeff = f;
                                    } else {
                                        const uint64_t bits = (uint64_t)r4[0] | ((uint64_t)r4[1] << 32);
                                        memcpy(&coeff, &bits, 8);
                                    }
                                    // the kernel's parity: popc((t_zl & jhi) ^ (t_zoc & tile number)); the run fields of t_zl sit above bit 15
                                    const int par = __builtin_popcount((r4[2] & jhi) ^ (r4[3] & (uint32_t)t)) & 1;
                                    srun += par ? -coeff : coeff;
                                }
                                w += (__builtin_popcount(vrow & jr) & 1) ? -srun : srun;
                                tt += len;
                            }
                            const cplx wc = imag ? cplx(0.0, w) : cplx(w, 0.0);
                            const cplx contrib = std::conj(own) * wc * tile[j ^ xl];
                            if (half) {
                                if (imag || (31 - __builtin_clz(xl)) != (int)skipbit) return 13;
                                if (((j >> skipbit) & 1u) == pol) out += 2.0 * std::real(contrib);
                            } else {
                                out += contrib;
                            }
                            rp += 1 + nt;
                        }
                        next_rec = rp;
                    }
                    if (next_rec != n_rec - 1) return 14;
                }
            }
        }
    }
    out2[0] = out.real();
    out2[1] = out.imag();
    return 0;
}
