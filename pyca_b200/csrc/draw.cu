// Fused draw(): state -> RGB on the device, for sm_100a.
//
// The reference's per-frame path after step() is draw() -- a handful of torch elementwise ops producing the float
// (3,H,W) `_worldmap` (ca2d.py:117-124, lenia.py:481-500, macelenia.py:365-385, nca.py:74-87) -- followed by
// Automaton.worldmap (automaton.py:209-216): 255 * permute(2,1,0) -> D2H of a FLOAT (W,H,3) tensor -> uint8 cast on the
// host.  Here one kernel per model reads the state in its native layout (packed bits, framed planes, NCHW), applies the
// model's colour rule, and writes (a) the float `_worldmap` (still the public attribute, and the echo state of CA2D) and
// (b) the final (W,H,3) uint8 frame, transposed through shared memory so both sides are coalesced.  The host then copies
// 3 bytes per cell instead of 12.
//
//   frame[x][y][c] = (uint8) trunc(255.f * worldmap[c][y][x])        -- the same fp32 product and truncation as
//                                                                        (255 * t).numpy().astype(np.uint8)
#include "common.cuh"

using namespace b200ca;

namespace {

constexpr int TILE = 32;

struct GolSrc {
    const uint32_t *bits;
    float *wm;  // (3,H,W) in: previous frame's worldmap (echo); out: the new one
    int stride;
    float d0, d1, d2;  // decay_speed * highlight_color
    __device__ __forceinline__ void rgb(int y, int x, int H, int W, float &r, float &g, float &b) const {
        const float w = (float)((bits[(size_t)y * stride + (x >> 5)] >> (x & 31)) & 1u);
        const size_t o = (size_t)y * W + x, P = (size_t)H * W;
        // echo = clamp(worldmap - decay, 0, 1);  worldmap' = clamp(world + echo, 0, 1)   (ca2d.py:121-124)
        r = fminf(fmaxf(w + fminf(fmaxf(wm[o] - d0, 0.f), 1.f), 0.f), 1.f);
        g = fminf(fmaxf(w + fminf(fmaxf(wm[o + P] - d1, 0.f), 1.f), 0.f), 1.f);
        b = fminf(fmaxf(w + fminf(fmaxf(wm[o + 2 * P] - d2, 0.f), 1.f), 0.f), 1.f);
    }
};

struct LeniaSrc {
    const float *st;    // framed planes of the shown world: plane c at st + c * plane
    const float *food;  // dense (H,W) added to all three colours (macelenia.py:379-380) or nullptr
    size_t plane;
    int pitch, C;
    __device__ __forceinline__ void rgb(int y, int x, int H, int W, float &r, float &g, float &b) const {
        const size_t o = (size_t)(y + B200CA_FRAME) * pitch + x + B200CA_FRAME;
        // C == 1: grey; C == 2: (c0, c1, 0); C >= 3: the first three channels   (lenia.py:489-495)
        r = st[o];
        g = C == 1 ? r : st[o + plane];
        b = C == 1 ? r : (C == 2 ? 0.f : st[o + 2 * plane]);
        if (food) {
            const float f = food[(size_t)y * W + x];
            r += f, g += f, b += f;
        }
        r = fminf(fmaxf(r, 0.f), 1.f), g = fminf(fmaxf(g, 0.f), 1.f), b = fminf(fmaxf(b, 0.f), 1.f);
    }
};

struct NcaSrc {
    const float *st;  // dense (n,H,W) planes of the shown world; channels 0..3 = RGBA (nca.py:267-271)
    __device__ __forceinline__ void rgb(int y, int x, int H, int W, float &r, float &g, float &b) const {
        const size_t o = (size_t)y * W + x, P = (size_t)H * W;
        const float a = fminf(fmaxf(st[o + 3 * P], 0.f), 1.f);
        // pic = clamp(argb, 0, 1); rgb = pic[:3] * pic[3:]  (blend on black, nca.py:79-82)
        r = fminf(fmaxf(st[o], 0.f), 1.f) * a;
        g = fminf(fmaxf(st[o + P], 0.f), 1.f) * a;
        b = fminf(fmaxf(st[o + 2 * P], 0.f), 1.f) * a;
    }
};

struct WmSrc {
    const float *wm;  // an existing float (3,H,W) worldmap (any draw() the reference or a subclass produced)
    __device__ __forceinline__ void rgb(int y, int x, int H, int W, float &r, float &g, float &b) const {
        const size_t o = (size_t)y * W + x, P = (size_t)H * W;
        r = wm[o], g = wm[o + P], b = wm[o + 2 * P];
    }
};

// grid (ceil(W/32), ceil(H/32)), block (32, 8).  Reads run along x, the uint8 frame is written along (y, c).
template <typename Src>
__global__ void __launch_bounds__(256) draw_kernel(const Src src, float *__restrict__ wm, uint8_t *__restrict__ frame, int H, int W) {
    __shared__ uint8_t tile[TILE][TILE * 3 + 4];  // [x][3*y + c], rows padded to 100 bytes
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    const int x = x0 + threadIdx.x;
    const size_t P = (size_t)H * W;
#pragma unroll
    for (int k = 0; k < TILE / 8; ++k) {
        const int yl = threadIdx.y + 8 * k, y = y0 + yl;
        if (x < W && y < H) {
            float r, g, b;
            src.rgb(y, x, H, W, r, g, b);
            if (wm) {
                const size_t o = (size_t)y * W + x;
                wm[o] = r, wm[o + P] = g, wm[o + 2 * P] = b;
            }
            tile[threadIdx.x][3 * yl] = (uint8_t)__float2uint_rz(255.f * r);
            tile[threadIdx.x][3 * yl + 1] = (uint8_t)__float2uint_rz(255.f * g);
            tile[threadIdx.x][3 * yl + 2] = (uint8_t)__float2uint_rz(255.f * b);
        }
    }
    if (!frame) return;
    __syncthreads();
    // column x of the tile = up to 96 consecutive bytes of frame row x
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int ny = min(TILE, H - y0) * 3;
    for (int e = tid; e < TILE * TILE * 3; e += 256) {
        const int xl = e / (TILE * 3), bi = e % (TILE * 3);
        if (x0 + xl < W && bi < ny) frame[((size_t)(x0 + xl) * H + y0) * 3 + bi] = tile[xl][bi];
    }
}

template <typename Src>
int launch_draw(const Src &src, float *wm, uint8_t *frame, int H, int W, void *stream) {
    B200CA_REQUIRE(H >= 1 && W >= 1 && (wm || frame), "draw: need H,W >= 1 and at least one output");
    dim3 grid(ceil_div(W, TILE), ceil_div(H, TILE));
    draw_kernel<Src><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, wm, frame, H, W);
    B200CA_LAUNCH_CHECK("draw_kernel");
    return 0;
}

}  // namespace

extern "C" int b200ca_draw_gol(const uint32_t *bits, float *worldmap, uint8_t *frame, int H, int W, int stride_words, float decay_r,
                               float decay_g, float decay_b, void *stream) {
    B200CA_REQUIRE(bits && worldmap, "draw_gol: bits and worldmap (the echo state) are required");
    GolSrc s{bits, worldmap, stride_words, decay_r, decay_g, decay_b};
    return launch_draw(s, worldmap, frame, H, W, stream);
}

extern "C" int b200ca_draw_lenia(const float *state, const float *food, float *worldmap, uint8_t *frame, int batch_index, int C,
                                 int H, int W, void *stream) {
    B200CA_REQUIRE(state && C >= 1 && batch_index >= 0, "draw_lenia: bad arguments");
    const int pitch = b200ca_frame_pitch(W);
    const size_t plane = (size_t)(H + 2 * B200CA_FRAME) * pitch;
    LeniaSrc s{state + (size_t)batch_index * C * plane, food, plane, pitch, C};
    return launch_draw(s, worldmap, frame, H, W, stream);
}

extern "C" int b200ca_draw_nca(const float *state, float *worldmap, uint8_t *frame, int batch_index, int n, int H, int W, void *stream) {
    B200CA_REQUIRE(state && n >= 4 && batch_index >= 0, "draw_nca: need a state with >= 4 channels (RGBA)");
    NcaSrc s{state + (size_t)batch_index * n * H * W};
    return launch_draw(s, worldmap, frame, H, W, stream);
}

extern "C" int b200ca_frame_pack_rgb(const float *worldmap, uint8_t *frame, int H, int W, void *stream) {
    B200CA_REQUIRE(worldmap && frame, "frame_pack_rgb: null pointer");
    WmSrc s{worldmap};
    return launch_draw(s, nullptr, frame, H, W, stream);
}
