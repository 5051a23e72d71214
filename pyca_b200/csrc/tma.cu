// Host side of the TMA path: tensor-map encoding through the driver entry point (no -lcuda link
// dependency, so the library loads on machines without a driver and fails only when used).
#include <map>
#include <mutex>
#include <tuple>

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

typedef std::tuple<int, const void *, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t> MapKey;
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

void clear_tensor_map_cache() {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    g_maps.clear();
}

int get_tensor_map_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
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
        if (g_maps.size() > 4096) g_maps.clear();
        g_maps[key] = m;
    }
    *out = m;
    return 0;
}

}  // namespace b200ca

extern "C" int b200ca_shutdown(void) {
    b200ca::clear_tensor_map_cache();
    return 0;
}
