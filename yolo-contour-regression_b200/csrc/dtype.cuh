// Head outputs arrive in fp32, or - under autocast, the reference's default (engine/trainer.py:332,
// engine/validator.py:103-104) - in fp16 / bf16.  The kernels read them in place and compute in fp32, as the
// reference does after its own up-casts (utils/loss.py:861; BCE-with-logits and pow are fp32 ops under autocast);
// gradients are written back in the input type.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

enum { YCR_F32 = 0, YCR_F16 = 1, YCR_BF16 = 2 };

static inline int ycr_dtype_size(int dt) { return dt == YCR_F32 ? 4 : 2; }

#ifdef __CUDACC__
// element i of an array of the run-time type dt
__device__ __forceinline__ float ycr_ld(const void* p, int64_t i, int dt) {
    if (dt == YCR_F32) return reinterpret_cast<const float*>(p)[i];
    if (dt == YCR_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// The same load without the conversion: the raw bits of element i (an fp32 word, or a 16-bit value zero-extended).
// A kernel that requests values long before it uses them keeps the raw words in registers - converting at the load
// would make every load wait for its own data - and turns them into floats with ycr_from_raw at the point of use.
__device__ __forceinline__ uint32_t ycr_ld_raw(const void* p, int64_t i, int dt) {
    if (dt == YCR_F32) return reinterpret_cast<const uint32_t*>(p)[i];
    return reinterpret_cast<const unsigned short*>(p)[i];
}

__device__ __forceinline__ float ycr_from_raw(uint32_t raw, int dt) {
    if (dt == YCR_F32) return __uint_as_float(raw);
    if (dt == YCR_F16) return __half2float(__ushort_as_half((unsigned short)raw));
    return __uint_as_float(raw << 16);   // a bf16 is the upper half of the fp32 with the same value
}

__device__ __forceinline__ void ycr_st(void* p, int64_t i, float v, int dt) {
    if (dt == YCR_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (dt == YCR_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// the value a tensor of type dt would hold (the reference's sigmoid / target_scores.to(dtype) stay in the input type)
__device__ __forceinline__ float ycr_round_to(float v, int dt) {
    if (dt == YCR_F16) return __half2float(__float2half_rn(v));
    if (dt == YCR_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

template <typename T> struct YcrType;
template <> struct YcrType<float> {
    static constexpr int code = YCR_F32;
    static __device__ __forceinline__ float4 ld4cs(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void st4cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
};
template <> struct YcrType<__half> {
    static constexpr int code = YCR_F16;
    static __device__ __forceinline__ float4 ld4cs(const __half* p) {
        const uint2 r = __ldcs(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ void st4cs(__half* p, float4 v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned*>(&a);
        r.y = *reinterpret_cast<const unsigned*>(&b);
        __stcs(reinterpret_cast<uint2*>(p), r);
    }
};
template <> struct YcrType<__nv_bfloat16> {
    static constexpr int code = YCR_BF16;
    static __device__ __forceinline__ float4 ld4cs(const __nv_bfloat16* p) {
        const uint2 r = __ldcs(reinterpret_cast<const uint2*>(p));
        // a bf16 is the upper half of the fp32 with the same value
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xffff0000u));
    }
    static __device__ __forceinline__ void st4cs(__nv_bfloat16* p, float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned*>(&a);
        r.y = *reinterpret_cast<const unsigned*>(&b);
        __stcs(reinterpret_cast<uint2*>(p), r);
    }
};
#endif
