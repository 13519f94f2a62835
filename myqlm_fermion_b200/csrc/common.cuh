// Shared internals of libqfe.so (not part of the ABI; the ABI is include/qfe.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/qfe.h"

namespace qfe {

void set_error(const char* fmt, ...);

#define QFE_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            qfe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return err__ == cudaErrorMemoryAllocation ? QFE_ERR_OOM : QFE_ERR_CUDA;            \
        }                                                                                      \
    } while (0)

#define QFE_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            qfe::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

#define QFE_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != QFE_OK) return rc__; \
    } while (0)

// RAII device buffer (raw cudaMalloc; freed with the owner)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    int alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        QFE_CUDA(cudaMalloc(&p, n));
        bytes = n;
        return QFE_OK;
    }
    // grow-only
    int reserve(size_t n) {
        if (n <= bytes) return QFE_OK;
        return alloc(n);
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace qfe

struct qfe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int64_t launches = 0;
    qfe::DevBuf partials;   // per-block partial sums of the sweep kernels
    qfe::DevBuf result;     // final reduced scalars
    qfe::DevBuf hostvec;    // device landing buffer of qfe_expectation_host
    cudaStream_t copy_stream = nullptr;      // upload stream of qfe_expectation_host
    std::vector<cudaEvent_t> copy_events;
    qfe::DevBuf rots;       // rotation tables of qfe_pauli_rotations (all batches of one call)
    qfe::DevBuf prof;       // development: phase cycle counters of the sweep kernels (QFE_PROF=1)
    qfe::DevBuf slabs;      // per-part partial results of APPLY sweeps whose tiles are shared by several CTAs (small shards)
    void* pinned = nullptr; // small pinned host buffer for scalar read-back
    size_t pinned_bytes = 0;
    // multi-GPU (sharded.cu): the NCCL communicator of this rank, a stream for the shard exchanges and the events that order them
    void* comm = nullptr;   // ncclComm_t
    int comm_rank = 0, comm_nranks = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t comm_ev[3] = {nullptr, nullptr, nullptr};  // [b]: partner buffer b has arrived; [2]: compute stream position at the start of a walk
    cudaEvent_t comm_done[2] = {nullptr, nullptr};         // [b]: the sweeps that read partner buffer b are enqueued
    qfe::DevBuf comm_scratch;                              // scalars on their way through an all-reduce
    int64_t exchanges = 0, exchanged_bytes = 0;
    std::vector<std::pair<int, double>> class_ms;          // last qfe_expectation_sharded*: (x_hi, host milliseconds incl. the wait for the exchange)
};

struct qfe_terms {
    int n = 0;               // qubits
    int64_t T = 0;           // terms (identity excluded)
    int64_t n_slots = 1;     // operators in the batch
    // canonical host copy, ascending (slot, x, z); coefficients in Y-form (Term.coeff)
    std::vector<uint64_t> x, z;
    std::vector<double> c;         // 2*T
    std::vector<int32_t> owner;    // T
    std::vector<double> constant;  // 2*n_slots
    int64_t n_xmasks = 0;          // distinct (slot, x)
};

namespace qfe {

inline int grid_for(int64_t n, int block, int sm_count, int per_sm = 8) {
    int64_t g = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- device scan / sort primitives (scan_sort.cu) ------------------------------------------
// exclusive prefix sum of n uint32 (in place allowed: out may equal in); total optional (device ptr)
int exclusive_scan_u32(qfe_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, DevBuf& scratch, uint32_t* d_total);
// stable LSD radix sort of records (x, z, idx) on the low `n_bits_z` bits of z then the low `n_bits_x`
// bits of x (x most significant), plus `n_bits_owner` bits of the 32-bit owner key (most significant of all).
int radix_sort_terms(qfe_ctx* ctx, uint64_t* x, uint64_t* z, uint32_t* owner, uint32_t* idx, int64_t n, int n_bits_x,
                     int n_bits_z, int n_bits_owner, DevBuf& tmp);

// vals[r] = sum over blocks of partials[r * n_blocks + b] in a fixed order (sweep.cu)
int reduce_to_vals(qfe_ctx* ctx, const double2* partials, int n_ranges, int n_blocks, double2* vals);

// qfe_expectation_class_part with the option to form only the real part of the value (sweep.cu; used by the sharded driver, whose pair
// ranks add 2 Re <own|G|partner>)
int expectation_class_part_impl(qfe_ctx* ctx, const qfe_plan* p, int x_hi, const void* bra, const void* ket, int dtype, int part, int n_parts, bool real_only,
                                double* out);

// Lanczos recurrence on an abstract operator (lanczos.cu): `apply(cur, w)` enqueues w = H cur on ctx->stream for this rank's part of the
// vectors (n_amps amplitudes each), `reduce(vals, n)` sums host scalars over the ranks (identity on one GPU).
struct LanczosOps {
    std::function<int(const void*, void*)> apply;
    std::function<int(double*, int)> reduce;
};
int lanczos_run(qfe_ctx* ctx, const LanczosOps& ops, int64_t n_amps, void* v, void* work, int max_iter, double tol, int dtype, int want_vector,
                double* eval, int* iters, double* resid);

}  // namespace qfe
