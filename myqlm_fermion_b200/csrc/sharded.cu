// Multi-GPU drivers behind the C ABI: the statevector is sharded by its high qubits over the ranks of one NCCL communicator that the
// qfe_ctx owns (one process per GPU).  Rank r holds the 2^n_local amplitudes whose top index bits equal r; a term whose
// X mask has high part x_hi maps shard r onto shard r ^ x_hi, so the terms of a plan fall into classes, class 0 is local and every
// other class needs ONE pairwise exchange of the shard with rank r ^ x_hi: ncclSend + ncclRecv in one group, over NVLink / NVSwitch.
// With room for two partner buffers the exchange of class k+1 runs on a second stream while the sweeps of class k run.
//
// The reference (/root/reference, pure Python) has no distributed path; these entry points replace the same call sites as the
// single-GPU ones (qat/plugins/adapt_vqe.py:181-185, 206; qat/fermion/vqe.py:53-59; tests/integration/test_integration_models.py:102-103)
// for registers that do not fit one GPU.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already mapped into the process -- torch's -- or the system one), so
// libqfe.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>  // types only; every call goes through the table below
#include <nvtx3/nvToolsExt.h>  // header-only: ranges per class / exchange show up in Nsight when a tool is attached

#include <chrono>
#include <cmath>

#include "common.cuh"

namespace qfe {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return QFE_OK;
    void* h = nullptr;
    const char* env = getenv("QFE_NCCL_LIB");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses (e.g. torch.distributed's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    QFE_REQUIRE(h, QFE_ERR_UNSUPPORTED, "NCCL not found (libnccl.so.2; set QFE_NCCL_LIB): %s", dlerror());
    NcclApi a;
    a.handle = h;
#define QFE_NCCL_SYM(field, name)                                                        \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                       \
    QFE_REQUIRE(a.field, QFE_ERR_UNSUPPORTED, "NCCL symbol %s missing", name)
    QFE_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    QFE_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    QFE_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    QFE_NCCL_SYM(Send, "ncclSend");
    QFE_NCCL_SYM(Recv, "ncclRecv");
    QFE_NCCL_SYM(AllReduce, "ncclAllReduce");
    QFE_NCCL_SYM(GroupStart, "ncclGroupStart");
    QFE_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    QFE_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    QFE_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef QFE_NCCL_SYM
    g_nccl = a;
    return QFE_OK;
}

#define QFE_NCCL(call)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != ncclSuccess) {                                                                              \
            qfe::set_error("%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));    \
            return QFE_ERR_COMM;                                                                               \
        }                                                                                                      \
    } while (0)

inline ncclComm_t comm_of(const qfe_ctx* ctx) { return static_cast<ncclComm_t>(ctx->comm); }

size_t shard_bytes(const qfe_plan* p, int dtype) {
    int n = 0, n_local = 0;
    qfe_plan_geometry(p, &n, &n_local);
    return (size_t(1) << n_local) * (dtype == QFE_C128 ? 16 : 8);
}

// Pairwise exchange of this rank's shard with rank ^ x_hi into `dst`, on `stream`.
int exchange(qfe_ctx* ctx, const void* shard, void* dst, size_t bytes, int x_hi, cudaStream_t stream) {
    const int peer = ctx->comm_rank ^ x_hi;
    QFE_REQUIRE(peer >= 0 && peer < ctx->comm_nranks, QFE_ERR_STATE, "class %d has no partner among %d ranks", x_hi, ctx->comm_nranks);
    char label[48];
    snprintf(label, sizeof(label), "qfe exchange x_hi=%d", x_hi);
    nvtxRangePushA(label);
    struct Pop { ~Pop() { nvtxRangePop(); } } pop;
    QFE_NCCL(g_nccl.GroupStart());
    QFE_NCCL(g_nccl.Send(shard, bytes / 8, ncclFloat64, peer, comm_of(ctx), stream));
    QFE_NCCL(g_nccl.Recv(dst, bytes / 8, ncclFloat64, peer, comm_of(ctx), stream));
    QFE_NCCL(g_nccl.GroupEnd());
    ctx->exchanges++;
    ctx->exchanged_bytes += (int64_t)bytes;
    return QFE_OK;
}

// in-place sum over the ranks of n host doubles (through a small device buffer; one NCCL all-reduce)
int allreduce_host(qfe_ctx* ctx, double* vals, int n) {
    if (ctx->comm_nranks <= 1) return QFE_OK;
    QFE_TRY(ctx->comm_scratch.reserve(sizeof(double) * (size_t)std::max(n, 2)));
    QFE_CUDA(cudaMemcpyAsync(ctx->comm_scratch.p, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    QFE_NCCL(g_nccl.AllReduce(ctx->comm_scratch.p, ctx->comm_scratch.p, (size_t)n, ncclFloat64, ncclSum, comm_of(ctx), ctx->stream));
    QFE_CUDA(cudaMemcpyAsync(vals, ctx->comm_scratch.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    QFE_CUDA(cudaStreamSynchronize(ctx->stream));
    return QFE_OK;
}

// The classes of a plan and the exchange pipeline over them: next() hands out, class by class, the buffer that holds the ket shard of
// rank ^ x_hi (the own shard for class 0).  With two partner buffers the exchange of the following non-local class is already in
// flight on the communication stream while the caller computes on the current one.
struct ClassWalk {
    qfe_ctx* ctx;
    const void* shard;
    char* partner;
    size_t bytes;
    int n_buf;
    std::vector<int32_t> classes;
    std::vector<int> nonlocal;  // indices into classes
    size_t issued = 0;          // non-local classes whose exchange has been enqueued
    size_t taken = 0;           // non-local classes handed to the caller

    int init(qfe_ctx* c, const qfe_plan* p, const void* s, void* partner_buf, size_t partner_bytes, int dtype) {
        ctx = c;
        shard = s;
        partner = static_cast<char*>(partner_buf);
        bytes = shard_bytes(p, dtype);
        int32_t nc = 0;
        QFE_TRY(qfe_plan_info(p, nullptr, nullptr, nullptr, &nc, nullptr));
        classes.assign((size_t)std::max(nc, 1), 0);
        QFE_TRY(qfe_plan_classes(p, classes.data(), (int32_t)classes.size()));
        classes.resize((size_t)nc);
        for (size_t i = 0; i < classes.size(); ++i)
            if (classes[i] != 0) nonlocal.push_back((int)i);
        n_buf = (int)std::min<size_t>(2, bytes ? partner_bytes / bytes : 0);
        QFE_REQUIRE(nonlocal.empty() || (partner && n_buf >= 1), QFE_ERR_ARG, "partner buffer too small: %zu bytes, a shard has %zu", partner_bytes, bytes);
        QFE_REQUIRE(nonlocal.empty() || ctx->comm, QFE_ERR_STATE, "the plan has non-local classes but the context has no communicator (qfe_comm_init)");
        if (!nonlocal.empty()) {
            // the communication stream starts behind whatever produced the shard on the compute stream
            QFE_CUDA(cudaEventRecord(ctx->comm_ev[2], ctx->stream));
            QFE_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[2], 0));
        }
        return QFE_OK;
    }
    int issue_one() {
        const int b = (int)(issued % (size_t)n_buf);
        // the buffer was read by the class that used it last: its sweeps are on the compute stream
        if (issued >= (size_t)n_buf) QFE_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_done[b], 0));
        QFE_TRY(exchange(ctx, shard, partner + (size_t)b * bytes, bytes, classes[(size_t)nonlocal[issued]], ctx->comm_stream));
        QFE_CUDA(cudaEventRecord(ctx->comm_ev[b], ctx->comm_stream));
        ++issued;
        return QFE_OK;
    }
    // buffer holding the ket shard for class index ci; the compute stream is made to wait for its arrival
    int ket_for(size_t ci, const void** ket) {
        if (classes[ci] == 0) {
            *ket = shard;
            // an exchange may run behind the local class
            if (n_buf >= 2 && issued == taken && issued < nonlocal.size()) QFE_TRY(issue_one());
            return QFE_OK;
        }
        if (issued == taken) QFE_TRY(issue_one());
        const int b = (int)(taken % (size_t)n_buf);
        QFE_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[b], 0));
        *ket = partner + (size_t)b * bytes;
        ++taken;
        if (n_buf >= 2 && issued == taken && issued < nonlocal.size()) QFE_TRY(issue_one());  // the next one travels during this class
        return QFE_OK;
    }
    // call when the work on class index ci has been enqueued on the compute stream
    int done_with(size_t ci) {
        if (classes[ci] == 0) return QFE_OK;
        const int b = (int)((taken - 1) % (size_t)n_buf);
        QFE_CUDA(cudaEventRecord(ctx->comm_done[b], ctx->stream));
        return QFE_OK;
    }
};

}  // namespace

int comm_allreduce_host(qfe_ctx* ctx, double* vals, int n) { return allreduce_host(ctx, vals, n); }

}  // namespace qfe

using namespace qfe;

extern "C" {

int qfe_comm_unique_id(uint8_t id[128]) {
    QFE_REQUIRE(id, QFE_ERR_ARG, "qfe_comm_unique_id: null argument");
    QFE_TRY(load_nccl());
    ncclUniqueId u;
    QFE_NCCL(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return QFE_OK;
}

int qfe_comm_init(qfe_ctx* ctx, const uint8_t id[128], int rank, int nranks) {
    QFE_REQUIRE(ctx && id, QFE_ERR_ARG, "qfe_comm_init: null argument");
    QFE_REQUIRE(nranks >= 1 && (nranks & (nranks - 1)) == 0, QFE_ERR_ARG, "qfe_comm_init: %d ranks is not a power of two", nranks);
    QFE_REQUIRE(rank >= 0 && rank < nranks, QFE_ERR_ARG, "qfe_comm_init: rank %d of %d", rank, nranks);
    QFE_REQUIRE(!ctx->comm, QFE_ERR_STATE, "qfe_comm_init: the context already has a communicator");
    QFE_TRY(load_nccl());
    QFE_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm = nullptr;
    QFE_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_nranks = nranks;
    QFE_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; ++i) QFE_CUDA(cudaEventCreateWithFlags(&ctx->comm_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) QFE_CUDA(cudaEventCreateWithFlags(&ctx->comm_done[i], cudaEventDisableTiming));
    return QFE_OK;
}

int qfe_comm_info(const qfe_ctx* ctx, int* rank, int* nranks, int* nccl_version, int64_t* exchanges, int64_t* exchanged_bytes) {
    QFE_REQUIRE(ctx, QFE_ERR_ARG, "qfe_comm_info: null ctx");
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (nranks) *nranks = ctx->comm ? ctx->comm_nranks : 1;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    if (exchanges) *exchanges = ctx->exchanges;
    if (exchanged_bytes) *exchanged_bytes = ctx->exchanged_bytes;
    return QFE_OK;
}

int qfe_comm_class_times(const qfe_ctx* ctx, int32_t* x_hi, double* ms, int32_t capacity, int32_t* n) {
    QFE_REQUIRE(ctx && n, QFE_ERR_ARG, "qfe_comm_class_times: null argument");
    *n = (int32_t)ctx->class_ms.size();
    for (int32_t i = 0; i < *n && i < capacity; ++i) {
        if (x_hi) x_hi[i] = ctx->class_ms[(size_t)i].first;
        if (ms) ms[i] = ctx->class_ms[(size_t)i].second;
    }
    return QFE_OK;
}

int qfe_comm_destroy(qfe_ctx* ctx) {
    if (!ctx || !ctx->comm) return QFE_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->comm_stream);
    g_nccl.CommDestroy(comm_of(ctx));
    ctx->comm = nullptr;
    cudaStreamDestroy(ctx->comm_stream);
    ctx->comm_stream = nullptr;
    for (int i = 0; i < 3; ++i) cudaEventDestroy(ctx->comm_ev[i]);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->comm_done[i]);
    ctx->comm_nranks = 1;
    ctx->comm_rank = 0;
    return QFE_OK;
}

int qfe_allreduce_sum(qfe_ctx* ctx, double* values, int n) {
    QFE_REQUIRE(ctx && values && n >= 0, QFE_ERR_ARG, "qfe_allreduce_sum: bad argument");
    QFE_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return QFE_OK;
    return allreduce_host(ctx, values, n);
}

int qfe_exchange_shard(qfe_ctx* ctx, const void* shard, void* partner, int64_t n_amps, int dtype, int x_hi) {
    QFE_REQUIRE(ctx && shard && partner && n_amps > 0, QFE_ERR_ARG, "qfe_exchange_shard: bad argument");
    QFE_REQUIRE(ctx->comm, QFE_ERR_STATE, "qfe_exchange_shard: no communicator (qfe_comm_init)");
    QFE_REQUIRE(x_hi > 0 && x_hi < ctx->comm_nranks, QFE_ERR_ARG, "qfe_exchange_shard: x_hi %d outside 1..%d", x_hi, ctx->comm_nranks - 1);
    QFE_CUDA(cudaSetDevice(ctx->device));
    QFE_TRY(exchange(ctx, shard, partner, (size_t)n_amps * (dtype == QFE_C128 ? 16 : 8), x_hi, ctx->stream));
    return QFE_OK;
}

// <bra|H|ket> over the whole register, identical on every rank.
static int expectation_sharded_impl(qfe_ctx* ctx, const qfe_plan* p, const void* bra, const void* ket, void* partner, size_t partner_bytes, int dtype,
                                    double* out, const double* class0 /* already evaluated share of class 0, or null */) {
    int64_t n_slots64 = 0;
    {
        int n = 0, nl = 0;
        QFE_TRY(qfe_plan_geometry(p, &n, &nl));
        QFE_REQUIRE(ctx->comm_nranks == (1 << (n - nl)), QFE_ERR_STATE, "the plan splits the register into %d shards, the communicator has %d ranks",
                    1 << (n - nl), ctx->comm_nranks);
    }
    QFE_TRY(qfe_plan_slots(p, &n_slots64));
    const size_t n_slots = (size_t)n_slots64;
    int hermitian = 0;
    QFE_TRY(qfe_plan_hermitian(p, &hermitian));
    // <psi|H|psi> with every class Hermitian: the shard pairs (r, r^x_hi) and (r^x_hi, r) of a class give conjugate numbers.  The two
    // ranks of a pair split the sweeps of the class between them (each with its own shard as the bra) and count their half twice.
    const bool split = (bra == ket) && n_slots == 1 && hermitian;
    std::vector<double> acc(2 * n_slots, 0.0), part(2 * n_slots, 0.0);
    if (class0) {
        acc[0] += class0[0];
        acc[1] += class0[1];
    }
    ClassWalk walk;
    QFE_TRY(walk.init(ctx, p, ket, partner, partner_bytes, dtype));
    ctx->class_ms.clear();
    for (size_t ci = 0; ci < walk.classes.size(); ++ci) {
        const int x_hi = walk.classes[ci];
        if (x_hi == 0 && class0) continue;
        char label[48];
        snprintf(label, sizeof(label), "qfe expectation class x_hi=%d", x_hi);
        nvtxRangePushA(label);
        struct Pop { ~Pop() { nvtxRangePop(); } } pop;
        const auto t0 = std::chrono::steady_clock::now();
        struct Clock {
            qfe_ctx* c; int x; std::chrono::steady_clock::time_point t;
            ~Clock() { c->class_ms.push_back({x, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count()}); }
        } clock{ctx, x_hi, t0};
        const void* k = nullptr;
        QFE_TRY(walk.ket_for(ci, &k));
        if (x_hi != 0 && split) {
            const int pivot = __builtin_ctz((unsigned)x_hi);
            QFE_TRY(expectation_class_part_impl(ctx, p, x_hi, bra, k, dtype, (ctx->comm_rank >> pivot) & 1, 2, /*real_only=*/true, part.data()));
            acc[0] += 2.0 * part[0];
        } else {
            QFE_TRY(qfe_expectation_class(ctx, p, x_hi, bra, k, dtype, part.data()));
            for (size_t i = 0; i < 2 * n_slots; ++i) acc[i] += part[i];
        }
        QFE_TRY(walk.done_with(ci));
    }
    QFE_TRY(allreduce_host(ctx, acc.data(), (int)acc.size()));
    for (size_t i = 0; i < 2 * n_slots; ++i) out[i] = acc[i];
    return QFE_OK;
}

int qfe_expectation_sharded(qfe_ctx* ctx, const qfe_plan* p, const void* bra_shard, const void* ket_shard, void* partner, int64_t partner_bytes, int dtype,
                            double* out) {
    QFE_REQUIRE(ctx && p && ket_shard && out, QFE_ERR_ARG, "qfe_expectation_sharded: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    return expectation_sharded_impl(ctx, p, bra_shard ? bra_shard : ket_shard, ket_shard, partner, (size_t)std::max<int64_t>(partner_bytes, 0), dtype, out, nullptr);
}

int qfe_expectation_sharded_host(qfe_ctx* ctx, const qfe_plan* p, const void* shard_host, void* shard_dev, void* partner, int64_t partner_bytes, int dtype,
                                 double out[2]) {
    QFE_REQUIRE(ctx && p && shard_host && shard_dev && out, QFE_ERR_ARG, "qfe_expectation_sharded_host: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    double local[2] = {0.0, 0.0};
    // the upload hides behind the sweeps of the local class; the exchanges start from the device copy afterwards
    QFE_TRY(qfe_expectation_class0_host(ctx, p, shard_host, shard_dev, dtype, local));
    return expectation_sharded_impl(ctx, p, shard_dev, shard_dev, partner, (size_t)std::max<int64_t>(partner_bytes, 0), dtype, out, local);
}

int qfe_apply_sharded(qfe_ctx* ctx, const qfe_plan* p, const void* ket_shard, void* y_shard, void* partner, int64_t partner_bytes, int dtype) {
    QFE_REQUIRE(ctx && p && ket_shard && y_shard, QFE_ERR_ARG, "qfe_apply_sharded: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    ClassWalk walk;
    QFE_TRY(walk.init(ctx, p, ket_shard, partner, (size_t)std::max<int64_t>(partner_bytes, 0), dtype));
    bool first = true;
    for (size_t ci = 0; ci < walk.classes.size(); ++ci) {
        char label[48];
        snprintf(label, sizeof(label), "qfe apply class x_hi=%d", walk.classes[ci]);
        nvtxRangePushA(label);
        struct Pop { ~Pop() { nvtxRangePop(); } } pop;
        const void* k = nullptr;
        QFE_TRY(walk.ket_for(ci, &k));
        QFE_TRY(qfe_apply_class(ctx, p, walk.classes[ci], k, y_shard, dtype, first ? 0 : 1));
        QFE_TRY(walk.done_with(ci));
        first = false;
    }
    if (first) QFE_CUDA(cudaMemsetAsync(y_shard, 0, walk.bytes, ctx->stream));
    return QFE_OK;
}

int qfe_commutator_gradients_sharded(qfe_ctx* ctx, const qfe_plan* h_plan, const qfe_plan* pool_plan, const void* psi_shard, void* scratch, void* partner,
                                     int64_t partner_bytes, int dtype, double* grads) {
    QFE_REQUIRE(ctx && h_plan && pool_plan && psi_shard && scratch && grads, QFE_ERR_ARG, "qfe_commutator_gradients_sharded: null argument");
    // phi = H psi (sharded like psi), then g_k = 2 Im <phi| tau_k |psi>: every pool class reads the psi shard of rank ^ x_hi
    QFE_TRY(qfe_apply_sharded(ctx, h_plan, psi_shard, scratch, partner, partner_bytes, dtype));
    int64_t K = 0;
    QFE_TRY(qfe_plan_slots(pool_plan, &K));
    std::vector<double> vals(2 * (size_t)K);
    QFE_TRY(qfe_expectation_sharded(ctx, pool_plan, scratch, psi_shard, partner, partner_bytes, dtype, vals.data()));
    for (int64_t k = 0; k < K; ++k) grads[k] = 2.0 * vals[2 * k + 1];
    return QFE_OK;
}

int qfe_vec_dot_sharded(qfe_ctx* ctx, const void* a, const void* b, int64_t n_amps, int dtype, double out[2]) {
    QFE_TRY(qfe_vec_dot(ctx, a, b, n_amps, dtype, out));
    return allreduce_host(ctx, out, 2);
}

int qfe_lanczos_sharded(qfe_ctx* ctx, const qfe_plan* p, void* v, void* work, void* partner, int64_t partner_bytes, int max_iter, double tol, int dtype,
                        int want_vector, double* eval, int* iters, double* resid) {
    QFE_REQUIRE(ctx && p && v && work && eval, QFE_ERR_ARG, "qfe_lanczos_sharded: null argument");
    QFE_REQUIRE(dtype == QFE_C128 || dtype == QFE_C64, QFE_ERR_ARG, "unknown dtype %d", dtype);
    QFE_CUDA(cudaSetDevice(ctx->device));
    int n = 0, n_local = 0;
    QFE_TRY(qfe_plan_geometry(p, &n, &n_local));
    LanczosOps ops;
    ops.apply = [&](const void* cur, void* w) { return qfe_apply_sharded(ctx, p, cur, w, partner, partner_bytes, dtype); };
    ops.reduce = [&](double* vals, int cnt) { return allreduce_host(ctx, vals, cnt); };
    return lanczos_run(ctx, ops, int64_t(1) << n_local, v, work, max_iter, tol, dtype, want_vector, eval, iters, resid);
}

}  // extern "C"
