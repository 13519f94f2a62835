// Host-side sweep planner of kernel (2).  Pure C++ (no CUDA) so that the planner's output can be
// interpreted on the CPU by the test-only emulator in tests/csrc/.
//
// A *group* is the set of terms of one operator slot that share an X mask.  Groups whose X mask
// has high part x_hi (the bits above n_local) form *class* x_hi: they read the ket shard of rank
// shard^x_hi.  Inside a class the groups are covered by *sweeps*: a sweep fixes a set L of k
// shard-local bit positions (the *tile mask*): 4 *register positions* that (almost) no X mask of the sweep touches, and k-4
// *X positions* (5 lane bits, up to 4 row-block bits).  Every group of the sweep has x_lo = ket_xor + (bits inside L) with one
// ket_xor outside L, so a CTA that stages the 2^k amplitudes {base | spread(j)} of one *tile* in shared memory finds every
// XOR partner in that tile.  One sweep = one kernel launch = one pass over the shard in HBM, evaluating all its groups.
// Tile-local index j: bits 0..3 = register rows (16 amplitudes a thread keeps in registers), bits 4..8 = lane, bits 9.. =
// row block (warp).  The 3 lowest shard bits are always in the tile (128 B contiguous segments in HBM).
#pragma once

#include <cstdint>
#include <vector>

namespace qfe {

struct TermView {
    int n = 0;
    int64_t T = 0;
    const uint64_t* x = nullptr;
    const uint64_t* z = nullptr;
    const double* c = nullptr;  // Y-form (Term.coeff), interleaved complex
    const int32_t* owner = nullptr;
    int64_t n_slots = 1;
    const double* constant = nullptr;  // 2*n_slots
};

constexpr int PLAN_MAX_TILE_BITS = 13;  // 2^13 complex128 amplitudes = 128 KB of shared memory
constexpr int PLAN_LOW_BITS = 3;        // 8 amplitudes = 128 B (c128) contiguous per gather segment: always inside the tile
constexpr int PLAN_REG_BITS = 4;        // tile-local bits [0, 4) index the 16 rows a thread keeps in registers
constexpr int PLAN_U_SHIFT = PLAN_REG_BITS + 5;  // tile-local bits >= 9 are warp-uniform (row-block index = warp index)
constexpr int PLAN_MIN_TILE_BITS = PLAN_U_SHIFT; // one full warp row block; below this the direct kernel is used
constexpr int PLAN_MAX_RUNS = 24;
constexpr int PLAN_TABLE = 192;         // per-sweep lookup tables: 128 entries for index bits 0..6, 64 for bits 7..12

// group header word a
constexpr uint32_t HDR_XL_MASK = 0x1fffu;   // bits 0..12: tile-local X mask (bits 0..3 = register rows, 4..12 = thread)
constexpr int HDR_NT_SHIFT = 13;            // bits 13..24: number of staged terms (<= 4095)
constexpr uint32_t HDR_NT_MASK = 0xfffu;
constexpr uint32_t HDR_IMAG = 1u << 25;     // gather coefficients purely imaginary (stored in t_cre)
// group header word b: unused (0)

// staged term word t_zl: bits 4..12 = z bits on the lane / row-block positions of the tile; bits 16..19 = z bits on the register
// rows (v); bits 20..31 = length of the run of equal v this term starts (0 for the other terms of the run).
constexpr uint32_t TZ_HI_MASK = ((1u << PLAN_MAX_TILE_BITS) - 1u) & ~((1u << PLAN_REG_BITS) - 1u);
constexpr int TZ_V_SHIFT = 16, TZ_RUN_SHIFT = 20;

// A sweep's groups are ordered into *ranges* of equal (skip bit, polarity, kind):
//   skip bit: tile-local index (>= PLAN_U_SHIFT) of the top X bit of a Hermitian group with real gather coefficients, 0 if
//             none.  <psi|G|psi> on one tile equals 2 Re over the rows whose skip bit equals the range's polarity, so the row
//             blocks (warps) of the other polarity skip the range.  The planner balances the two polarities of every skip bit
//             so that all row blocks carry the same work;
//   kind:     how W(row) = sum_t c_t (-1)^{|z_t & row|} is formed for the thread's 16 register rows, and where the partner rows are.
enum { KIND_FAST = 0,  // no z bit on the register rows: W is the same for the 16 rows; term list padded to an even length
       KIND_RUNS = 1,  // terms sorted by their register-row z bits v: one signed sum S per run, <.> folded with (-1)^{v.r}
       KIND_XROW = 2   // flag: the X mask touches register rows (partner row = row ^ xr); combined with KIND_RUNS
};
// range descriptor: word 0 = first group | (one past the last group) << 16 (sweep-local);
//                   word 1 = first term (sweep-local) | kind << 16 | polarity << 18 | skip bit << 24
constexpr int RNG_KIND_SHIFT = 16, RNG_POL_SHIFT = 18, RNG_SKIP_SHIFT = 24;

struct PlanLimits {
    int tile_bits = 13;
    int max_terms = 2048;  // per sweep (shared-memory staging capacity)
    int max_groups = 512;
    int group_overhead = 0;  // staging slots a group costs besides its terms (1 when the group headers travel with the terms, sweep.cu)
    int max_slots = 0;       // terms + group_overhead * groups of a sweep stay within this (0: no such bound)
    int fill_below = 32;   // a sweep with fewer groups also takes pending groups whose X mask touches its register positions
                           // (measured optimum at 28 qubits: 1.19 s against 1.22 s without, 1.21 s with 48)
};

struct SweepHost {
    int cls = 0;            // index into PlanHost::classes
    int k = 0;              // tile bits
    uint64_t tile_mask = 0; // L
    uint64_t ket_xor = 0;   // shard-local bits outside L shared by every X mask of the sweep: ket tile = base ^ ket_xor
    uint8_t pos[PLAN_MAX_TILE_BITS] = {0};  // tile-local bit i <-> shard position pos[i]
    int n_runs = 0;         // runs of the outer mask (shard bits outside L), LSB first
    uint8_t run_pos[PLAN_MAX_RUNS] = {0};
    uint8_t run_len[PLAN_MAX_RUNS] = {0};
    int64_t n_tiles = 0;
    int32_t group_begin = 0, group_end = 0;
    int64_t term_begin = 0, term_end = 0;
    int32_t range_begin = 0, range_end = 0;
    bool complex_w = false;  // some group mixes real and imaginary gather coefficients
};

struct ClassHost {
    int32_t x_hi = 0;
    int32_t sweep_begin = 0, sweep_end = 0;
    int64_t dterm_begin = 0, dterm_end = 0;    // direct-kernel term range
    int32_t dgroup_begin = 0, dgroup_end = 0;  // direct-kernel group range
};

struct PlanHost {
    int n = 0, n_local = 0, shard = 0, k = 0;
    int64_t n_slots = 1;
    bool use_tiles = false;
    std::vector<ClassHost> classes;  // ascending x_hi
    std::vector<SweepHost> sweeps;
    // tiled layout (sweep order); two header words per group: the HDR_* fields, then the sweep-local index of its first staged term
    std::vector<uint32_t> g_hdr;
    std::vector<int32_t> g_slot;
    std::vector<uint32_t> r_desc;  // two words per range
    // per sweep, PLAN_TABLE entries each: *natural* index (bit i <-> i-th smallest position of L; what the coalesced staging loop
    // enumerates) -> shard offset and -> tile-local index; tile-local index -> shard offset
    std::vector<uint64_t> nat_off, loc_off;
    std::vector<uint32_t> nat_loc;
    std::vector<uint32_t> t_zl;    // staged term word, see the TZ_* constants
    std::vector<uint64_t> t_zo;    // shard-local z bits outside the tile
    std::vector<uint32_t> t_zoc;   // the same bits packed against the tile number: sign of the term on tile t = parity(t_zoc & t)
    std::vector<double> t_cre, t_cim;  // gather-form coefficients, rank sign folded; imag groups keep Im in t_cre
    // direct layout (class order, original term order inside a class)
    std::vector<uint64_t> d_x, d_z;  // shard-local masks
    std::vector<double> d_cre, d_cim;
    std::vector<int32_t> dg_slot;      // per direct group
    std::vector<int64_t> dg_begin;     // per direct group: first term (size n_dgroups+1)
};

// Record stream of ONE real-coefficient sweep, as tile_expect2_kernel reads it from its kernel parameter (sweep.cu, Expect2Args):
//   rng: two words per range -- word 0 = record index of the first group header | number of groups << 16; word 1 = kind / polarity /
//        skip bit as in r_desc (low 16 bits zero);
//   rec: four words per record -- a group is one header record (word 0 = header word a) followed by its staged terms (words 0/1 =
//        coefficient bits: a double, or a float in word 0 for complex64 states; word 2 = t_zl; word 3 = t_zoc); one all-zero record ends
//        the stream (the kernel prefetches the header after the last group).
struct SweepRecords {
    std::vector<uint32_t> rng;
    std::vector<uint32_t> rec;
};
void build_sweep_records(const PlanHost& p, const SweepHost& s, bool single_precision, SweepRecords& out);

// returns 0 on success, nonzero with *err set otherwise
int build_plan(const TermView& t, int n_local, int shard, const PlanLimits& lim, PlanHost& out, const char** err);

// tile-local compression of a mask that lies inside the sweep's tile mask
uint32_t compress_to_tile(const SweepHost& s, uint64_t mask);
// shard position of tile-local index j for tile number `tile`
uint64_t tile_base(const SweepHost& s, uint64_t tile);
uint64_t tile_spread(const SweepHost& s, uint32_t j);
uint64_t tile_natural_spread(const SweepHost& s, uint32_t nat);

}  // namespace qfe
