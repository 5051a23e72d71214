// Host side of the TMA path: tensor-map encoding through the driver entry point (no -lcuda link
// dependency, so the library loads on machines without a driver and fails only when used).
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "tma.cuh"

namespace b200ca {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

// Descriptors live in GLOBAL memory (one 128-byte slot each, carved from per-device slabs): on this
// driver (580.x) a `__grid_constant__` CUtensorMap kernel parameter makes UTMALDG fault with "illegal
// instruction", a descriptor read through a global pointer works (tools/tma_probe.cu variants 4 / 12).
typedef std::tuple<int, const void *, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t> MapKey;
static std::map<MapKey, const CUtensorMap *> g_maps;
static std::mutex g_maps_mu;
struct Slab {
    CUtensorMap *base;
    int used;
    int dev;
};
static std::vector<Slab> g_slabs;
constexpr int SLAB_SLOTS = 512;

void clear_tensor_map_cache() {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    g_maps.clear();
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &s : g_slabs) {
        cudaSetDevice(s.dev);
        cudaDeviceSynchronize();
        cudaFree(s.base);
    }
    cudaSetDevice(cur);
    g_slabs.clear();
    (void)cudaGetLastError();
}

// caller holds g_maps_mu
static int alloc_slot(int dev, CUtensorMap **out) {
    for (auto &s : g_slabs)
        if (s.dev == dev && s.used < SLAB_SLOTS) {
            *out = s.base + s.used++;
            return 0;
        }
    Slab s;
    s.dev = dev;
    s.used = 0;
    s.base = nullptr;
    B200CA_CUDA(cudaMalloc(&s.base, sizeof(CUtensorMap) * SLAB_SLOTS));
    g_slabs.push_back(s);
    *out = g_slabs.back().base + g_slabs.back().used++;
    return 0;
}

int get_tensor_map_3d(const CUtensorMap **out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                      uint32_t b0, uint32_t b1) {
    int dev = 0;
    B200CA_CUDA(cudaGetDevice(&dev));
    MapKey key(dev, base, d0, d1, d2, s1, s2, b0, b1);
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return B200CA_E_DRIVER;
    }
    if (((uintptr_t)base & 15) || (s1 * 4) % 16 || (s2 * 4) % 16 || (b0 * 4) % 16 || b0 > 256 || b1 > 256) {
        set_error("tensor map: misaligned base/strides/box (base %p s1 %llu s2 %llu box %u x %u)", base, (unsigned long long)s1,
                  (unsigned long long)s2, b0, b1);
        return B200CA_E_BADARG;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1 * sizeof(float), s2 * sizeof(float)};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return B200CA_E_DRIVER;
    }
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        CUtensorMap *slot = nullptr;
        int rc = alloc_slot(dev, &slot);
        if (rc) return rc;
        // blocking copy (cold path): the descriptor is visible to every later launch on any stream
        B200CA_CUDA(cudaMemcpy(slot, &m, sizeof(m), cudaMemcpyHostToDevice));
        g_maps[key] = slot;
        *out = slot;
    }
    return 0;
}

}  // namespace b200ca

extern "C" int b200ca_shutdown(void) {
    b200ca::clear_tensor_map_cache();
    b200ca::scratch_free_all();
    return 0;
}
