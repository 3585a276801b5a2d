// Shared declarations for the sm_100a kernels of the polar-contour hot path.
#pragma once
#include <utility>
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ycr_b200.h"

#define YCR_C YCR_CONTOUR_POINTS
#define YCR_EMPTY 0xFFFFFFFFu
#define YCR_FLOOR 1e-6f          // utils/tal.py:1189-1191,1275-1277,1455; nn/modules/head.py:480
#define YCR_GATE_DEG 3.0         // utils/tal.py:1185,1270
#define YCR_NUM_SMS 148
#define YCR_SMEM_MAX 232448      // 227 KB: opt-in shared memory per block on sm_100
// shared memory of k_resolve_image: [G][levels] rectangles, [A] pick words, 4 arrays of G*topk, 5 arrays of G
static inline int64_t ycr_resolve_smem_bytes(int64_t A, int64_t G, int64_t topk) {
    const int64_t pos_cap = (G * topk > 0) ? G * topk : 1;
    return G * YCR_MAX_LEVELS * 16 + A * 4 + pos_cap * 16 + G * 20 + 64;
}

void ycr_set_error(const char* fmt, ...);

// Optional per-kernel CUDA-event timing (ycr_profile_begin/read in the C ABI); a no-op unless enabled.
enum { YCR_T_SETUP = 0, YCR_T_CAND = 1, YCR_T_TOPK = 2, YCR_T_RESOLVE = 3, YCR_T_POS = 4, YCR_T_STREAM = 5,
       YCR_T_FINAL = 6, YCR_T_DECODE = 7, YCR_T_NMS_FILTER = 8, YCR_T_NMS_SORT = 9, YCR_T_NMS_SUPPRESS = 10,
       YCR_T_COUNT = 16 };
void ycr_prof_mark(int tag, int end, cudaStream_t st);

// Programmatic dependent launch (sm_90+): the kernels of the training path are launched with
// programmaticStreamSerializationAllowed, execute pdl_enter() as their first statement - wait until the
// preceding grid in the stream has completed and its writes are visible, then allow the next grid to be
// launched - so the launch and block scheduling of kernel N+1 overlap the tail of kernel N instead of following
// it.  YCR_PDL=0 in the environment turns the attribute off (plain stream order; pdl_enter is then a no-op).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
inline bool ycr_pdl_enabled() {
    static const int on = [] { const char* e = getenv("YCR_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
    return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t ycr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ycr_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
struct YcrProfScope {
    int tag; cudaStream_t st;
    YcrProfScope(int t, cudaStream_t s) : tag(t), st(s) { ycr_prof_mark(tag, 0, st); }
    ~YcrProfScope() { ycr_prof_mark(tag, 1, st); }
};

#define YCR_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ycr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return YCR_E_CUDA;                                                            \
        }                                                                                 \
    } while (0)

#define YCR_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ycr_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return YCR_E_CUDA;                                                            \
        }                                                                                 \
    } while (0)

// Device-side copy of the grid with prefix offsets.
struct GridDev {
    int n_levels;
    int h[YCR_MAX_LEVELS], w[YCR_MAX_LEVELS];
    int off[YCR_MAX_LEVELS + 1];  // anchor offset of each level; off[n_levels] = A
    float stride[YCR_MAX_LEVELS];
};

static inline GridDev make_grid_dev(const ycr_grid_t* g) {
    GridDev d{};
    d.n_levels = g->n_levels;
    int acc = 0;
    for (int l = 0; l < YCR_MAX_LEVELS; ++l) {
        d.off[l] = acc;
        if (l < g->n_levels) {
            d.h[l] = g->h[l];
            d.w[l] = g->w[l];
            d.stride[l] = g->stride[l];
            acc += g->h[l] * g->w[l];
        }
    }
    d.off[YCR_MAX_LEVELS] = acc;
    for (int l = g->n_levels; l <= YCR_MAX_LEVELS; ++l) d.off[l] = acc;
    return d;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Simple bump allocator over the caller's workspace.
struct WsAlloc {
    char* base;
    size_t off;
    size_t cap;
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
