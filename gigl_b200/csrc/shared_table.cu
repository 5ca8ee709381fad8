// shared_table.cu - the node-feature table sharded over the GPUs of one NVSwitch box, seen by every GPU as ONE
// flat array.
//
// features are sharded by contiguous node-id range (51 GB at N = 1e8, F = 128 -> 6.4 GB per
// GPU; MAG240M's 751 GB does not fit one GPU at all) and the gather kernel pulls remote neighbour rows over NVLink.
// Instead of an id -> (owner, offset) lookup in every kernel, the shards are stitched together with the CUDA virtual
// memory API: every process reserves one virtual range of n_shards * shard_bytes, creates its own shard as a
// shareable physical allocation (cuMemCreate, POSIX file descriptor handle), imports the peers' shards
// (cuMemImportFromShareableHandle) and maps shard k at offset k * shard_bytes.  Row v of the table is then at
// base + v * F * 4 on every GPU; a load of a remote row is an ordinary global load that the MMU routes over
// NVLink / NVSwitch.  All gather kernels (batch_collate.cu, sage_aggregate.cu) work on it unchanged.
//
// The file descriptors travel between the per-GPU processes over a Unix socket (host plumbing in
// gigl_b200/sharding.py); the reference has no equivalent (its features travel inside every sample proto).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace {

struct DriverApi {
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Export)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*Import)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*Reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
};

template <typename Fn>
bool load_sym(const char* name, Fn*& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
    fn = reinterpret_cast<Fn*>(p);
    return true;
}

DriverApi& api() {
    static DriverApi a;
    if (!a.ok) {
        a.ok = load_sym("cuMemGetAllocationGranularity", a.GetGranularity) && load_sym("cuMemCreate", a.Create) &&
               load_sym("cuMemExportToShareableHandle", a.Export) && load_sym("cuMemImportFromShareableHandle", a.Import) &&
               load_sym("cuMemAddressReserve", a.Reserve) && load_sym("cuMemMap", a.Map) && load_sym("cuMemSetAccess", a.SetAccess) &&
               load_sym("cuMemUnmap", a.Unmap) && load_sym("cuMemRelease", a.Release) && load_sym("cuMemAddressFree", a.AddressFree);
    }
    return a;
}

CUmemAllocationProp shard_prop(int device) {
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

int drv_fail(gigl_ctx* ctx, CUresult r, const char* what) {
    return gigl_fail(ctx, r == CUDA_ERROR_OUT_OF_MEMORY ? GIGL_E_NOMEM : GIGL_E_CUDA, std::string(what) + " failed (CUresult " + std::to_string((int)r) + ")");
}

}  // namespace

struct gigl_shared_table {
    gigl_ctx* ctx = nullptr;
    int32_t n_shards = 0, my_shard = 0, F = 0;
    int64_t rows_per_shard = 0;
    size_t shard_bytes = 0;
    CUdeviceptr base = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<char> mapped;
    int export_fd = -1;
};

extern "C" {

int gigl_shared_table_row_granule(gigl_ctx* ctx, int32_t F, int64_t* rows) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, F >= 1 && rows, "bad arguments");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    GIGL_CUDA(ctx, cudaFree(0));
    DriverApi& d = api();
    if (!d.ok) return gigl_fail(ctx, GIGL_E_CUDA, "the CUDA virtual memory management API is not available from this driver");
    CUmemAllocationProp prop = shard_prop(ctx->device);
    size_t gran = 0;
    CUresult r = d.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS) return drv_fail(ctx, r, "cuMemGetAllocationGranularity");
    // smallest row count whose bytes are a multiple of the granularity: gran / gcd(gran, 4 F)
    size_t a = gran, b = (size_t)F * 4;
    while (b) {
        const size_t t = a % b;
        a = b;
        b = t;
    }
    *rows = (int64_t)(gran / a);
    return GIGL_OK;
}

int gigl_shared_table_create(gigl_ctx* ctx, int32_t n_shards, int32_t my_shard, int64_t rows_per_shard, int32_t F,
                             gigl_shared_table** out, int32_t* export_fd) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out && export_fd && n_shards >= 1 && my_shard >= 0 && my_shard < n_shards && rows_per_shard >= 1 && F >= 1, "bad arguments");
    int64_t granule = 0;
    int rc = gigl_shared_table_row_granule(ctx, F, &granule);
    if (rc != GIGL_OK) return rc;
    GIGL_CHECK(ctx, rows_per_shard % granule == 0, "rows_per_shard must be a multiple of gigl_shared_table_row_granule()");
    DriverApi& d = api();
    gigl_shared_table* t = new (std::nothrow) gigl_shared_table();
    if (!t) return gigl_fail(ctx, GIGL_E_NOMEM, "out of host memory");
    t->ctx = ctx;
    t->n_shards = n_shards;
    t->my_shard = my_shard;
    t->F = F;
    t->rows_per_shard = rows_per_shard;
    t->shard_bytes = (size_t)rows_per_shard * F * 4;
    t->handles.assign(n_shards, 0);
    t->mapped.assign(n_shards, 0);
    CUresult r = d.Reserve(&t->base, t->shard_bytes * n_shards, 0, 0, 0);
    if (r != CUDA_SUCCESS) {
        delete t;
        return drv_fail(ctx, r, "cuMemAddressReserve");
    }
    CUmemAllocationProp prop = shard_prop(ctx->device);
    r = d.Create(&t->handles[my_shard], t->shard_bytes, &prop, 0);
    if (r != CUDA_SUCCESS) {
        d.AddressFree(t->base, t->shard_bytes * n_shards);
        delete t;
        return drv_fail(ctx, r, "cuMemCreate");
    }
    int fd = -1;
    r = d.Export(&fd, t->handles[my_shard], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.Release(t->handles[my_shard]);
        d.AddressFree(t->base, t->shard_bytes * n_shards);
        delete t;
        return drv_fail(ctx, r, "cuMemExportToShareableHandle");
    }
    t->export_fd = fd;
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = ctx->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    const CUdeviceptr at = t->base + (CUdeviceptr)my_shard * t->shard_bytes;
    r = d.Map(at, t->shard_bytes, 0, t->handles[my_shard], 0);
    if (r == CUDA_SUCCESS) r = d.SetAccess(at, t->shard_bytes, &acc, 1);
    if (r != CUDA_SUCCESS) {
        gigl_shared_table_destroy(t);
        return drv_fail(ctx, r, "cuMemMap / cuMemSetAccess (own shard)");
    }
    t->mapped[my_shard] = 1;
    *export_fd = fd;
    *out = t;
    return GIGL_OK;
}

int gigl_shared_table_attach(gigl_shared_table* t, int32_t shard, int32_t fd) {
    if (!t) return gigl_fail(nullptr, GIGL_E_INVALID, "null table");
    gigl_ctx* ctx = t->ctx;
    GIGL_CHECK(ctx, shard >= 0 && shard < t->n_shards && shard != t->my_shard && !t->mapped[shard] && fd >= 0, "bad shard / fd");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    DriverApi& d = api();
    CUresult r = d.Import(&t->handles[shard], (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return drv_fail(ctx, r, "cuMemImportFromShareableHandle");
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = ctx->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    const CUdeviceptr at = t->base + (CUdeviceptr)shard * t->shard_bytes;
    r = d.Map(at, t->shard_bytes, 0, t->handles[shard], 0);
    if (r != CUDA_SUCCESS) return drv_fail(ctx, r, "cuMemMap (peer shard)");
    r = d.SetAccess(at, t->shard_bytes, &acc, 1);
    if (r != CUDA_SUCCESS) {
        d.Unmap(at, t->shard_bytes);
        return drv_fail(ctx, r, "cuMemSetAccess (peer shard; is P2P / NVLink access available between the two GPUs?)");
    }
    t->mapped[shard] = 1;
    return GIGL_OK;
}

int gigl_shared_table_ptrs(const gigl_shared_table* t, float** base_dev, float** my_shard_dev, int64_t* total_rows) {
    if (!t) return GIGL_E_INVALID;
    if (base_dev) *base_dev = (float*)t->base;
    if (my_shard_dev) *my_shard_dev = (float*)(t->base + (CUdeviceptr)t->my_shard * t->shard_bytes);
    if (total_rows) *total_rows = t->rows_per_shard * t->n_shards;
    return GIGL_OK;
}

void gigl_shared_table_destroy(gigl_shared_table* t) {
    if (!t) return;
    DriverApi& d = api();
    cudaSetDevice(t->ctx->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < t->n_shards; ++k) {
        if (t->mapped[k]) d.Unmap(t->base + (CUdeviceptr)k * t->shard_bytes, t->shard_bytes);
        if (t->handles[k]) d.Release(t->handles[k]);
    }
    if (t->base) d.AddressFree(t->base, t->shard_bytes * t->n_shards);
    if (t->export_fd >= 0) close(t->export_fd);
    delete t;
}

}  // extern "C"
