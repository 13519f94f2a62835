// Sweep planner (host).  See plan_host.h for the vocabulary.
#include "plan_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace qfe {

namespace {

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

struct Group {
    int32_t slot;
    uint64_t x_lo;
    int64_t tb;  // first term (index into the canonical term arrays)
    int32_t nt;
};

// gather-form coefficient of term i: (H psi)[b] = sum_t cg_t (-1)^{|z_t & b|} psi[b ^ x_t]
// cg = c_Y * (-i)^{|x&z|} * (-1)^{|z_hi & shard|}
inline void gather_coeff(const TermView& t, int64_t i, int n_local, int shard, double& re, double& im) {
    re = t.c[2 * i];
    im = t.c[2 * i + 1];
    switch (popc64(t.x[i] & t.z[i]) & 3) {
        case 1: { double a = re; re = im; im = -a; break; }
        case 2: re = -re; im = -im; break;
        case 3: { double a = re; re = -im; im = a; break; }
        default: break;
    }
    const uint64_t z_hi = n_local >= 64 ? 0 : (t.z[i] >> n_local);
    if (popc64(z_hi & (uint64_t)shard) & 1) {
        re = -re;
        im = -im;
    }
}

// outer-mask runs and tile count of a sweep whose tile mask and pos[] are set
void finish_sweep_geometry(SweepHost& s, int n_local) {
    const uint64_t all = n_local >= 64 ? ~uint64_t(0) : ((uint64_t(1) << n_local) - 1);
    const uint64_t outer = all & ~s.tile_mask;
    s.n_runs = 0;
    int b = 0;
    while (b < n_local) {
        if ((outer >> b) & 1) {
            int e = b;
            while (e < n_local && ((outer >> e) & 1)) ++e;
            s.run_pos[s.n_runs] = (uint8_t)b;
            s.run_len[s.n_runs] = (uint8_t)(e - b);
            ++s.n_runs;
            b = e;
        } else {
            ++b;
        }
    }
    s.n_tiles = int64_t(1) << (n_local - s.k);
}

}  // namespace

uint32_t compress_to_tile(const SweepHost& s, uint64_t mask) {
    uint32_t out = 0;
    for (int i = 0; i < s.k; ++i)
        if ((mask >> s.pos[i]) & 1) out |= 1u << i;
    return out;
}

uint64_t tile_base(const SweepHost& s, uint64_t tile) {
    uint64_t base = 0;
    for (int r = 0; r < s.n_runs; ++r) {
        const uint64_t m = (uint64_t(1) << s.run_len[r]) - 1;
        base |= (tile & m) << s.run_pos[r];
        tile >>= s.run_len[r];
    }
    return base;
}

uint64_t tile_spread(const SweepHost& s, uint32_t j) {
    uint64_t out = 0;
    for (int i = 0; i < s.k; ++i)
        if ((j >> i) & 1) out |= uint64_t(1) << s.pos[i];
    return out;
}

// natural index: bit i <-> i-th smallest position of the tile mask
uint64_t tile_natural_spread(const SweepHost& s, uint32_t nat) {
    uint64_t out = 0, rest = s.tile_mask;
    for (int i = 0; rest; ++i, rest &= rest - 1)
        if ((nat >> i) & 1) out |= uint64_t(1) << __builtin_ctzll(rest);
    return out;
}

namespace {

// bits of `mask` on the outer positions (outside the tile mask), packed in the order tile_base deposits the tile number
uint32_t compress_outer(const SweepHost& s, uint64_t mask) {
    uint32_t out = 0;
    int at = 0;
    for (int r = 0; r < s.n_runs; ++r) {
        out |= (uint32_t)((mask >> s.run_pos[r]) & ((uint64_t(1) << s.run_len[r]) - 1)) << at;
        at += s.run_len[r];
    }
    return out;
}

void append_tables(const SweepHost& s, PlanHost& out) {
    for (uint32_t e = 0; e < (uint32_t)PLAN_TABLE; ++e) {
        const uint32_t idx = e < 128 ? e : ((e - 128) << 7);  // entries 0..127: index bits 0..6; 128..191: bits 7..12
        const uint32_t in_tile = idx & ((1u << s.k) - 1u);
        const uint64_t off = tile_natural_spread(s, in_tile);
        out.nat_off.push_back(off);
        out.nat_loc.push_back(compress_to_tile(s, off));
        out.loc_off.push_back(tile_spread(s, in_tile));
    }
}

}  // namespace

int build_plan(const TermView& t, int n_local, int shard, const PlanLimits& lim, PlanHost& out, const char** err) {
    static const char* e_arg = "build_plan: bad geometry";
    if (n_local < 0 || n_local > t.n || t.n > 64 || t.n - n_local > 30 || shard < 0 || (t.n - n_local < 31 && shard >= (1 << (t.n - n_local)))) {
        if (err) *err = e_arg;
        return 1;
    }
    if (lim.tile_bits > PLAN_MAX_TILE_BITS || lim.max_terms > (int)HDR_NT_MASK || lim.max_terms < 1 || lim.max_groups < 1) {
        if (err) *err = "build_plan: limits out of range";
        return 1;
    }
    out = PlanHost();
    out.n = t.n;
    out.n_local = n_local;
    out.shard = shard;
    out.n_slots = t.n_slots;
    const uint64_t lomask = n_local >= 64 ? ~uint64_t(0) : ((uint64_t(1) << n_local) - 1);

    // terms: canonical set + one identity term per slot with a non-zero constant (it is just the
    // (x=0, z=0) member of the diagonal group; contributes constant * <bra|ket>)
    struct TRec {
        int32_t slot;
        uint64_t x, z;
        double re, im;
        bool herm;  // Term.coeff is real: the term is Hermitian on its own
    };
    std::vector<TRec> recs;
    recs.reserve(t.T + t.n_slots);
    {
        int64_t i = 0;
        for (int64_t s = 0; s < t.n_slots; ++s) {
            if (t.constant && (t.constant[2 * s] != 0.0 || t.constant[2 * s + 1] != 0.0)) recs.push_back({(int32_t)s, 0, 0, t.constant[2 * s], t.constant[2 * s + 1], t.constant[2 * s + 1] == 0.0});
            while (i < t.T && (t.owner ? t.owner[i] : 0) == s) {
                TRec r;
                r.slot = (int32_t)s;
                r.x = t.x[i];
                r.z = t.z[i];
                gather_coeff(t, i, n_local, shard, r.re, r.im);
                r.herm = t.c[2 * i + 1] == 0.0;
                recs.push_back(r);
                ++i;
            }
        }
        if (i != t.T) {
            if (err) *err = "build_plan: terms are not sorted by slot";
            return 1;
        }
    }
    const int64_t NT = (int64_t)recs.size();

    // groups per class
    std::map<int32_t, std::vector<Group>> by_class;
    for (int64_t i = 0; i < NT;) {
        int64_t j = i;
        while (j < NT && recs[j].slot == recs[i].slot && recs[j].x == recs[i].x) ++j;
        const int32_t x_hi = n_local >= 64 ? 0 : (int32_t)(recs[i].x >> n_local);
        // split oversized groups: (W1 + W2) psi[b^x] = W1 psi[b^x] + W2 psi[b^x]
        for (int64_t b = i; b < j; b += lim.max_terms) {
            Group g;
            g.slot = recs[i].slot;
            g.x_lo = recs[i].x & lomask;
            g.tb = b;
            g.nt = (int32_t)std::min<int64_t>(lim.max_terms, j - b);
            by_class[x_hi].push_back(g);
        }
        i = j;
    }

    const int k = std::min(n_local, lim.tile_bits);
    out.k = k;
    out.use_tiles = k >= PLAN_MIN_TILE_BITS;
    const int low = std::min(PLAN_LOW_BITS, k);
    const uint64_t lowmask = (uint64_t(1) << low) - 1;
    const int xcap = k - PLAN_REG_BITS;  // X positions of a tile (lane and row-block bits)

    for (auto& kv : by_class) {
        ClassHost cls;
        cls.x_hi = kv.first;
        std::vector<Group>& groups = kv.second;
        const int ci = (int)out.classes.size();

        // ---- direct layout ----
        cls.dterm_begin = (int64_t)out.d_x.size();
        cls.dgroup_begin = (int32_t)out.dg_slot.size();
        for (const Group& g : groups) {
            out.dg_slot.push_back(g.slot);
            out.dg_begin.push_back((int64_t)out.d_x.size());
            for (int64_t i = g.tb; i < g.tb + g.nt; ++i) {
                out.d_x.push_back(recs[i].x & lomask);
                out.d_z.push_back(recs[i].z & lomask);
                out.d_cre.push_back(recs[i].re);
                out.d_cim.push_back(recs[i].im);
            }
        }
        cls.dterm_end = (int64_t)out.d_x.size();
        cls.dgroup_end = (int32_t)out.dg_slot.size();

        // ---- tiled layout: greedy cover of the groups' x_lo by the X positions of tile masks ----
        cls.sweep_begin = (int32_t)out.sweeps.size();
        if (out.use_tiles) {
            const size_t ng = groups.size();
            std::vector<uint64_t> need(ng);
            std::vector<char> done(ng, 0);
            size_t left = ng;
            for (size_t g = 0; g < ng; ++g) need[g] = groups[g].x_lo;
            const double fbase = 8.0;
            double fpow[PLAN_MAX_TILE_BITS + 2];
            for (int m = 0; m <= PLAN_MAX_TILE_BITS + 1; ++m) fpow[m] = std::pow(fbase, -m);
            // wide groups (more X bits than a tile has X positions) are reached through a non-zero ket_xor and / or
            // register rows; narrow ones are covered first with ket_xor = 0 and X masks that avoid the register positions
            while (left > 0) {
                // seed: virtual all-zero mask while narrow groups remain, else the first wide group
                uint64_t seed = 0;
                bool have_narrow = false;
                for (size_t g = 0; g < ng && !have_narrow; ++g)
                    if (!done[g] && popc64(need[g]) <= xcap) have_narrow = true;
                if (!have_narrow) {
                    for (size_t g = 0; g < ng; ++g)
                        if (!done[g]) {
                            seed = need[g];
                            break;
                        }
                }
                uint64_t Xs = 0;  // X positions
                for (int step = 0; step < xcap; ++step) {
                    const int cap = xcap - step - 1;  // slots left after this pick
                    double gain[64];
                    for (int p = 0; p < 64; ++p) gain[p] = 0.0;
                    for (size_t g = 0; g < ng; ++g) {
                        if (done[g]) continue;
                        uint64_t miss = (need[g] ^ seed) & ~Xs;
                        const int m = popc64(miss);
                        if (m == 0 || m - 1 > cap) continue;
                        const double d = fpow[m - 1] - (m <= cap ? fpow[m] : 0.0);
                        while (miss) {
                            gain[__builtin_ctzll(miss)] += d;
                            miss &= miss - 1;
                        }
                    }
                    // a tile holds the low bits and 4 register positions besides the X positions: keep room for them
                    int best = -1;
                    for (int p = 0; p < n_local; ++p)
                        if (!((Xs >> p) & 1) && (best < 0 || gain[p] > gain[best])) best = p;
                    if (gain[best] == 0.0) {
                        // nothing left to gain: pad with the lowest unused positions above what the register rows will take
                        int spare = 0;
                        best = -1;
                        for (int p = 0; p < n_local && best < 0; ++p)
                            if (!((Xs >> p) & 1) && ++spare > PLAN_REG_BITS) best = p;
                        if (best < 0) break;
                    }
                    Xs |= uint64_t(1) << best;
                }
                // register positions: the low bits the X positions left out, then the free positions on which the fewest of the
                // groups this sweep will take carry Z bits (no Z bit on the register rows = one weight for the 16 rows)
                uint64_t Rm = lowmask & ~Xs;
                {
                    double zcost[64];
                    for (int p = 0; p < 64; ++p) zcost[p] = 0.0;
                    for (size_t g = 0; g < ng; ++g) {
                        if (done[g] || ((need[g] ^ seed) & ~Xs)) continue;
                        const Group& gr = groups[g];
                        uint64_t zany = 0, zdiff = 0;
                        for (int64_t i = gr.tb; i < gr.tb + gr.nt; ++i) {
                            zany |= recs[i].z;
                            zdiff |= recs[i].z ^ recs[gr.tb].z;
                        }
                        for (uint64_t m = zany & lomask & ~Xs; m; m &= m - 1) zcost[__builtin_ctzll(m)] += 1.0;
                        for (uint64_t m = zdiff & lomask & ~Xs; m; m &= m - 1) zcost[__builtin_ctzll(m)] += 3.0;  // several runs
                    }
                    while (popc64(Rm) < PLAN_REG_BITS) {
                        int best = -1;
                        for (int p = 0; p < n_local; ++p)
                            if (!(((Xs | Rm) >> p) & 1) && (best < 0 || zcost[p] < zcost[best])) best = p;
                        if (best < 0) break;
                        Rm |= uint64_t(1) << best;
                    }
                }
                for (int p = 0; p < n_local && popc64(Xs | Rm) < k; ++p)  // early exit of the pick loop: fill the tile
                    if (!(((Xs | Rm) >> p) & 1)) Xs |= uint64_t(1) << p;
                const uint64_t L = Xs | Rm;
                const uint64_t kxor = seed & ~L;
                SweepHost s;
                s.cls = ci;
                s.k = k;
                s.tile_mask = L;
                s.ket_xor = kxor;
                // take coverable groups in order until the staging capacity is reached.  A narrow group that touches a register
                // position waits for a sweep whose X positions hold all its bits (first pass); wide ones have no such sweep.
                std::vector<size_t> take;
                int64_t nterms = 0;
                // A sweep that stays small streams the whole shard for little work: it then also takes the pending groups whose X mask
                // touches its register positions (slower per group, but one pass over the shard less later on).
                const char* fill_env = getenv("QFE_PLAN_FILL");  // development knob: 0 disables the fill
                const int fill_below = fill_env ? atoi(fill_env) : lim.fill_below;
                for (int pass = 0; pass < 2 && (take.empty() || (pass == 1 && (int)take.size() < fill_below)); ++pass)
                    for (size_t g = 0; g < ng; ++g) {
                        if (done[g] || ((need[g] & ~L) != kxor)) continue;
                        const bool touches_rows = (need[g] & Rm) && popc64(need[g]) <= xcap;
                        if (pass == 0 ? touches_rows : (!touches_rows && !take.empty())) continue;  // pass 1 adds what pass 0 deferred
                        const int nt_padded = (groups[g].nt + 1) & ~1;  // staged term lists are padded to an even length
                        if ((int)take.size() >= lim.max_groups || nterms + nt_padded > lim.max_terms) break;
                        if (lim.max_slots > 0 && !take.empty() && nterms + nt_padded + lim.group_overhead * ((int64_t)take.size() + 1) > lim.max_slots) break;
                        take.push_back(g);
                        nterms += nt_padded;
                    }
                // classify coefficients
                bool cplx = false;
                std::vector<char> imag(take.size(), 0);
                for (size_t a = 0; a < take.size(); ++a) {
                    const Group& g = groups[take[a]];
                    bool all_re = true, all_im = true;
                    for (int64_t i = g.tb; i < g.tb + g.nt; ++i) {
                        if (recs[i].im != 0.0) all_re = false;
                        if (recs[i].re != 0.0) all_im = false;
                    }
                    if (all_re)
                        imag[a] = 0;
                    else if (all_im)
                        imag[a] = 1;
                    else
                        cplx = true;
                }
                s.complex_w = cplx;
                // tile-local bit order: register rows, then lanes (the low bits among the X positions first: consecutive
                // amplitudes of a staged segment then differ in the lowest lane bits), then row blocks.  The row-block
                // positions are picked to hit as many Hermitian groups as possible: those are visited on half the row blocks.
                {
                    std::vector<int> rows, lanes, ublk, cand;
                    for (int p = 0; p < n_local; ++p) {
                        if ((Rm >> p) & 1) rows.push_back(p);
                        if (((Xs & lowmask) >> p) & 1) lanes.push_back(p);
                        if (((Xs & ~lowmask) >> p) & 1) cand.push_back(p);
                    }
                    const int nu = k - PLAN_U_SHIFT;
                    std::vector<char> hit(take.size(), 0);
                    for (size_t a = 0; a < take.size(); ++a) {
                        const Group& g = groups[take[a]];
                        bool eligible = !cplx && !imag[a] && kxor == 0;
                        for (int64_t i = g.tb; i < g.tb + g.nt && eligible; ++i)
                            if (!recs[i].herm) eligible = false;
                        if (!eligible) hit[a] = 1;  // never half-visited: not worth a row-block position
                    }
                    for (int iu = 0; iu < nu; ++iu) {
                        size_t best = 0;
                        double best_w = -1.0;
                        for (size_t c = 0; c < cand.size(); ++c) {
                            double w = 0.0;
                            for (size_t a = 0; a < take.size(); ++a)
                                if (!hit[a] && ((groups[take[a]].x_lo >> cand[c]) & 1)) w += 48.0 + 6.0 * groups[take[a]].nt;
                            if (w > best_w) {
                                best_w = w;
                                best = c;
                            }
                        }
                        for (size_t a = 0; a < take.size(); ++a)
                            if ((groups[take[a]].x_lo >> cand[best]) & 1) hit[a] = 1;
                        ublk.push_back(cand[best]);
                        cand.erase(cand.begin() + best);
                    }
                    std::sort(ublk.begin(), ublk.end());
                    for (int p : cand) lanes.push_back(p);
                    int i = 0;
                    for (int p : rows) s.pos[i++] = (uint8_t)p;
                    for (int p : lanes) s.pos[i++] = (uint8_t)p;
                    for (int p : ublk) s.pos[i++] = (uint8_t)p;
                }
                finish_sweep_geometry(s, n_local);
                s.group_begin = (int32_t)(out.g_hdr.size() / 2);
                s.term_begin = (int64_t)out.t_zl.size();
                s.range_begin = (int32_t)(out.r_desc.size() / 2);
                const uint32_t regmask = (1u << PLAN_REG_BITS) - 1u;
                // classify every group, then order the sweep's groups by (skip bit, polarity, kind) so that the kernel
                // walks homogeneous ranges: whole ranges are skipped / dispatched, not single groups
                struct GInfo {
                    size_t gi;
                    uint32_t xl;
                    bool imag;
                    int kind, skipbit, pol;
                    std::vector<int64_t> order;  // term emit order
                };
                auto zreg_of = [&](int64_t i) { return compress_to_tile(s, (recs[i].z & lomask) & L) & regmask; };
                std::vector<GInfo> info(take.size());
                double pol_load[PLAN_MAX_TILE_BITS + 1][2] = {{0.0}};
                for (size_t a = 0; a < take.size(); ++a) {
                    const Group& g = groups[take[a]];
                    GInfo& gi = info[a];
                    gi.gi = take[a];
                    gi.xl = compress_to_tile(s, g.x_lo);
                    gi.imag = !cplx && imag[a];
                    bool fast = true, hermitian = true;
                    for (int64_t i = g.tb; i < g.tb + g.nt; ++i) {
                        if (zreg_of(i) != 0) fast = false;
                        if (!recs[i].herm) hermitian = false;
                    }
                    gi.skipbit = 0;
                    gi.pol = 0;
                    if (hermitian && !cplx && !gi.imag && kxor == 0 && (gi.xl >> PLAN_U_SHIFT) != 0) gi.skipbit = 31 - __builtin_clz(gi.xl);
                    for (int64_t i = g.tb; i < g.tb + g.nt; ++i) gi.order.push_back(i);
                    const bool xrow = (gi.xl & regmask) != 0;
                    gi.kind = (fast && !xrow) ? KIND_FAST : (KIND_RUNS | (xrow ? KIND_XROW : 0));
                    if (!fast) std::stable_sort(gi.order.begin(), gi.order.end(), [&](int64_t p, int64_t q) { return zreg_of(p) < zreg_of(q); });
                }
                // Polarity of the half-visited groups: the rows whose skip bit equals `pol` do the work of a group, the others skip it.  Every
                // warp waits for the slowest one at the end of a tile, so the two halves of every skip-bit level must carry equal work:
                // longest-processing-time order (heaviest group first, each to the lighter half) on a cost model of one group visit
                // (QFE_PLAN_COST=base,per_term,per_run; QFE_PLAN_BALANCE=0: assignment in group order, the round-1 behaviour).
                {
                    static const int balance = [] {
                        const char* e = getenv("QFE_PLAN_BALANCE");
                        return e ? atoi(e) : 1;
                    }();
                    static double cost_w[3] = {48.0, 6.0, 0.0};
                    static const bool cost_init = [] {
                        const char* e = getenv("QFE_PLAN_COST");
                        if (e) sscanf(e, "%lf,%lf,%lf", &cost_w[0], &cost_w[1], &cost_w[2]);
                        return true;
                    }();
                    (void)cost_init;
                    std::vector<std::pair<double, size_t>> half;  // (cost, index into info)
                    for (size_t a = 0; a < info.size(); ++a) {
                        if (info[a].skipbit == 0) continue;
                        const Group& g = groups[info[a].gi];
                        int runs = 0;
                        if (info[a].kind != KIND_FAST) {
                            uint32_t seen = 0;
                            for (int64_t i = g.tb; i < g.tb + g.nt; ++i) seen |= 1u << zreg_of(i);
                            runs = __builtin_popcount(seen);
                        }
                        half.push_back({cost_w[0] + cost_w[1] * g.nt + cost_w[2] * runs, a});
                    }
                    if (balance) std::stable_sort(half.begin(), half.end(), [](const std::pair<double, size_t>& p, const std::pair<double, size_t>& q) { return p.first > q.first; });
                    for (const auto& h : half) {
                        GInfo& gi = info[h.second];
                        gi.pol = pol_load[gi.skipbit][1] < pol_load[gi.skipbit][0] ? 1 : 0;
                        pol_load[gi.skipbit][gi.pol] += h.first;
                    }
                }
                std::vector<size_t> ord(take.size());
                for (size_t a = 0; a < ord.size(); ++a) ord[a] = a;
                auto range_key = [&](const GInfo& A) { return (A.skipbit << 3) | (A.pol << 2) | A.kind; };
                std::stable_sort(ord.begin(), ord.end(), [&](size_t p, size_t q) { return range_key(info[p]) < range_key(info[q]); });
                for (size_t oi = 0; oi < ord.size(); ++oi) {
                    const GInfo& gi = info[ord[oi]];
                    const Group& g = groups[gi.gi];
                    const uint32_t g_local = (uint32_t)(out.g_hdr.size() / 2 - s.group_begin);
                    const uint32_t t_local = (uint32_t)(out.t_zl.size() - s.term_begin);
                    const bool new_range = oi == 0 || range_key(info[ord[oi - 1]]) != range_key(gi);
                    if (new_range) {
                        out.r_desc.push_back(g_local | ((g_local + 1) << 16));
                        out.r_desc.push_back(t_local | ((uint32_t)gi.kind << RNG_KIND_SHIFT) | ((uint32_t)gi.pol << RNG_POL_SHIFT) |
                                             ((uint32_t)gi.skipbit << RNG_SKIP_SHIFT));
                    } else {
                        uint32_t& d0 = out.r_desc[out.r_desc.size() - 2];
                        d0 = (d0 & 0xffffu) | ((g_local + 1) << 16);
                    }
                    size_t run_head = 0;
                    for (size_t oi2 = 0; oi2 < gi.order.size(); ++oi2) {
                        const int64_t i = gi.order[oi2];
                        const uint32_t zl = compress_to_tile(s, (recs[i].z & lomask) & L);
                        const uint32_t v = zl & regmask;
                        if (oi2 == 0 || v != ((out.t_zl[run_head] >> TZ_V_SHIFT) & regmask)) run_head = out.t_zl.size();
                        out.t_zl.push_back((zl & TZ_HI_MASK) | (v << TZ_V_SHIFT));
                        out.t_zl[run_head] += 1u << TZ_RUN_SHIFT;
                        out.t_zo.push_back((recs[i].z & lomask) & ~L);
                        out.t_zoc.push_back(compress_outer(s, (recs[i].z & lomask) & ~L));
                        out.t_cre.push_back(gi.imag ? recs[i].im : recs[i].re);
                        out.t_cim.push_back(gi.imag ? 0.0 : recs[i].im);
                    }
                    int nt_emit = g.nt;
                    if (gi.kind == KIND_FAST && (g.nt & 1)) {  // zero-coefficient copy of the last term: keeps the list even
                        out.t_zl.push_back(out.t_zl.back() & ~(0xfffu << TZ_RUN_SHIFT));
                        out.t_zl[run_head] += 1u << TZ_RUN_SHIFT;
                        out.t_zo.push_back(out.t_zo.back());
                        out.t_zoc.push_back(out.t_zoc.back());
                        out.t_cre.push_back(0.0);
                        out.t_cim.push_back(0.0);
                        ++nt_emit;
                    }
                    out.g_hdr.push_back(gi.xl | ((uint32_t)nt_emit << HDR_NT_SHIFT) | (gi.imag ? HDR_IMAG : 0u));
                    out.g_hdr.push_back(t_local);  // first staged term of the group: lets a CTA start in the middle of a range
                    out.g_slot.push_back(g.slot);
                    done[gi.gi] = 1;
                    --left;
                }
                append_tables(s, out);
                s.range_end = (int32_t)(out.r_desc.size() / 2);
                s.group_end = (int32_t)(out.g_hdr.size() / 2);
                s.term_end = (int64_t)out.t_zl.size();
                out.sweeps.push_back(s);
                if (take.empty()) {  // cannot happen: every group fits an empty sweep
                    if (err) *err = "build_plan: internal error (empty sweep)";
                    return 1;
                }
            }
        }
        cls.sweep_end = (int32_t)out.sweeps.size();
        out.classes.push_back(cls);
    }
    out.dg_begin.push_back((int64_t)out.d_x.size());
    return 0;
}

void build_sweep_records(const PlanHost& p, const SweepHost& s, bool single_precision, SweepRecords& out) {
    out.rng.clear();
    out.rec.clear();
    for (int32_t r = s.range_begin; r < s.range_end; ++r) {
        const uint32_t d0 = p.r_desc[2 * (size_t)r], d1 = p.r_desc[2 * (size_t)r + 1];
        const uint32_t g0 = d0 & 0xffffu, g1 = d0 >> 16;
        out.rng.push_back((uint32_t)(out.rec.size() / 4) | ((g1 - g0) << 16));
        out.rng.push_back(d1 & ~0xffffu);
        for (uint32_t g = g0; g < g1; ++g) {
            const uint32_t hdr = p.g_hdr[2 * (size_t)(s.group_begin + g)], t0 = p.g_hdr[2 * (size_t)(s.group_begin + g) + 1];
            const uint32_t nt = (hdr >> HDR_NT_SHIFT) & HDR_NT_MASK;
            out.rec.insert(out.rec.end(), {hdr, 0u, 0u, 0u});
            for (uint32_t t = t0; t < t0 + nt; ++t) {
                const size_t i = (size_t)s.term_begin + t;
                uint32_t lo, hi = 0;
                if (single_precision) {
                    const float f = (float)p.t_cre[i];
                    memcpy(&lo, &f, 4);
                } else {
                    uint64_t bits;
                    memcpy(&bits, &p.t_cre[i], 8);
                    lo = (uint32_t)bits;
                    hi = (uint32_t)(bits >> 32);
                }
                out.rec.insert(out.rec.end(), {lo, hi, p.t_zl[i], p.t_zoc[i]});
            }
        }
    }
    out.rec.insert(out.rec.end(), {0u, 0u, 0u, 0u});
}

}  // namespace qfe
