// nccl_glue.cu -- the one exchange step of the path behind the C ABI: NCCL collectives launched on the
// library's compute stream, no host synchronisation in between.  NCCL is loaded at run time (dlopen of libnccl.so.2 --
// inside a PyTorch process that is the copy torch already mapped), so the library itself links against nothing but
// cudart and single-GPU users never need NCCL installed.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace rs {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

const NcclApi *nccl_api() {
    std::call_once(g_nccl_once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        g_nccl.ok = false;
        if (!h) return;
#define RS_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return;
        RS_SYM(GetUniqueId, "ncclGetUniqueId") RS_SYM(CommInitRank, "ncclCommInitRank") RS_SYM(CommDestroy, "ncclCommDestroy")
        RS_SYM(AllGather, "ncclAllGather") RS_SYM(AllReduce, "ncclAllReduce") RS_SYM(GetErrorString, "ncclGetErrorString")
#undef RS_SYM
        g_nccl.ok = true;
    });
    return g_nccl.ok ? &g_nccl : nullptr;
}

int nccl_fail(const char *what, int r) {
    const NcclApi *N = nccl_api();
    return set_error(RS_ERR_CUDA, "%s: NCCL error %d (%s)", what, r, N ? N->GetErrorString((ncclResult_t)r) : "libnccl not loaded");
}

// in-place all-gather of equal slices of a device buffer: rank r owns [r * count, (r+1) * count) words
int nccl_allgather_inplace_u64(void *comm, uint64_t *d_buf, size_t count, int rank) {
    const NcclApi *N = nccl_api();
    if (!N) return set_error(RS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    const ncclResult_t r = N->AllGather(d_buf + (size_t)rank * count, d_buf, count, ncclUint64, (ncclComm_t)comm, current_stream());
    if (r != ncclSuccess) return nccl_fail("ncclAllGather", (int)r);
    return RS_OK;
}
int nccl_allreduce_sum_u64(void *comm, uint64_t *d_buf, size_t count) {
    const NcclApi *N = nccl_api();
    if (!N) return set_error(RS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    const ncclResult_t r = N->AllReduce(d_buf, d_buf, count, ncclUint64, ncclSum, (ncclComm_t)comm, current_stream());
    if (r != ncclSuccess) return nccl_fail("ncclAllReduce", (int)r);
    return RS_OK;
}

}  // namespace rs

using namespace rs;

extern "C" {
#pragma GCC visibility push(default)

int rs_nccl_unique_id(uint8_t *id128) {
    RS_REQUIRE(id128, RS_ERR_ARG, "rs_nccl_unique_id: null argument");
    const NcclApi *N = nccl_api();
    if (!N) return set_error(RS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    const ncclResult_t r = N->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", (int)r);
    memcpy(id128, &id, 128);
    return RS_OK;
}

int rs_nccl_comm_create(const uint8_t *id128, int rank, int world, void **comm_out) {
    RS_REQUIRE(id128 && comm_out && world >= 1 && rank >= 0 && rank < world, RS_ERR_ARG, "rs_nccl_comm_create: bad argument");
    const NcclApi *N = nccl_api();
    if (!N) return set_error(RS_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    const ncclResult_t r = N->CommInitRank(&c, world, id, rank);
    if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", (int)r);
    *comm_out = (void *)c;
    return RS_OK;
}

void rs_nccl_comm_destroy(void *comm) {
    const NcclApi *N = nccl_api();
    if (N && comm) N->CommDestroy((ncclComm_t)comm);
}

// every rank contributes n words (host) and receives world * n words (host), rank-major: the per-shard commitments /
// evaluations that go into the host Fiat-Shamir transcript (tens of KiB)
int rs_allgather_u64(void *comm, int rank, int world, const uint64_t *host_in, size_t n, uint64_t *host_out) {
    RS_REQUIRE(comm && host_in && host_out && world >= 1 && rank >= 0 && rank < world, RS_ERR_ARG, "rs_allgather_u64: bad argument");
    if (n == 0) return RS_OK;
    uint64_t *d = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d, (size_t)world * n * 8, current_stream()));
    cudaError_t e = cudaMemcpyAsync(d + (size_t)rank * n, host_in, n * 8, cudaMemcpyHostToDevice, current_stream());
    int rc = (e == cudaSuccess) ? nccl_allgather_inplace_u64(comm, d, n, rank) : set_error(RS_ERR_CUDA, "rs_allgather_u64: %s", cudaGetErrorString(e));
    if (rc == RS_OK) {
        e = cudaMemcpyAsync(host_out, d, (size_t)world * n * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_allgather_u64: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(d, current_stream());
    return rc;
}

#pragma GCC visibility pop
}
