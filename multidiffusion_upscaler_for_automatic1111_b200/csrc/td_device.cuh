// Device-side helpers: element traits with explicit IEEE roundings, 128-bit
// read-only loads, unaligned-vector extraction by funnel shift.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace td {

// ---- element traits ------------------------------------------------------
// to_f32 is exact; from_f32 is round-to-nearest-even -- the same rounding
// torch applies when an fp32 intermediate is stored to a half/bf16 tensor.
template <typename T> struct Elem;
template <> struct Elem<__half> {
    static constexpr int kDtype = 0;
    __device__ static __forceinline__ float to_f32(__half v) { return __half2float(v); }
    __device__ static __forceinline__ __half from_f32(float f) { return __float2half_rn(f); }
    __device__ static __forceinline__ float bits_to_f32(uint16_t b) { return __half2float(__ushort_as_half(b)); }
    __device__ static __forceinline__ uint16_t f32_to_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kDtype = 1;
    __device__ static __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f32(float f) { return __float2bfloat16_rn(f); }
    __device__ static __forceinline__ float bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
    __device__ static __forceinline__ uint16_t f32_to_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
};
template <> struct Elem<float> {
    static constexpr int kDtype = 2;
    __device__ static __forceinline__ float to_f32(float v) { return v; }
    __device__ static __forceinline__ float from_f32(float f) { return f; }
};

// round an fp32 value through T (the "store to x_buffer, load it back" of the reference)
template <typename T> __device__ __forceinline__ float round_through(float f) { return Elem<T>::to_f32(Elem<T>::from_f32(f)); }
template <> __device__ __forceinline__ float round_through<float>(float f) { return f; }

// ---- 128-bit global access -------------------------------------------------
__device__ __forceinline__ uint4 ldg128(const void* p) {  // read-only path, keeps L1 (neighbour threads re-use the line)
    return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void stg128_stream(void* p, const uint4& v) {  // write-once data: do not pollute L1
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- unaligned 16-byte window out of two aligned 16-byte chunks -----------
// Given chunk A = elements [8k, 8k+8) and chunk B = [8k+8, 8k+16) of a row of
// 16-bit elements and a shift s in [0,8), returns elements [8k+s, 8k+s+8).
// `s` is uniform across the threads that handle the same tile column, so the
// switch does not diverge.
__device__ __forceinline__ uint4 window16(const uint4& A, const uint4& B, int s) {
    const uint32_t sh = (uint32_t)(s & 1) * 16u;
    uint4 o;
    switch (s >> 1) {
        case 0:
            o.x = __funnelshift_r(A.x, A.y, sh); o.y = __funnelshift_r(A.y, A.z, sh);
            o.z = __funnelshift_r(A.z, A.w, sh); o.w = __funnelshift_r(A.w, B.x, sh);
            break;
        case 1:
            o.x = __funnelshift_r(A.y, A.z, sh); o.y = __funnelshift_r(A.z, A.w, sh);
            o.z = __funnelshift_r(A.w, B.x, sh); o.w = __funnelshift_r(B.x, B.y, sh);
            break;
        case 2:
            o.x = __funnelshift_r(A.z, A.w, sh); o.y = __funnelshift_r(A.w, B.x, sh);
            o.z = __funnelshift_r(B.x, B.y, sh); o.w = __funnelshift_r(B.y, B.z, sh);
            break;
        default:
            o.x = __funnelshift_r(A.w, B.x, sh); o.y = __funnelshift_r(B.x, B.y, sh);
            o.z = __funnelshift_r(B.y, B.z, sh); o.w = __funnelshift_r(B.z, B.w, sh);
            break;
    }
    return o;
}
// Same for 32-bit elements: chunk = 4 elements, shift s in [0,4).
__device__ __forceinline__ uint4 window32(const uint4& A, const uint4& B, int s) {
    switch (s) {
        case 0: return A;
        case 1: return make_uint4(A.y, A.z, A.w, B.x);
        case 2: return make_uint4(A.z, A.w, B.x, B.y);
        default: return make_uint4(A.w, B.x, B.y, B.z);
    }
}

template <typename T> struct Vec {  // one 16-byte vector of T
    static constexpr int kElems = 16 / (int)sizeof(T);
    static constexpr int kLog2 = (sizeof(T) == 2) ? 3 : 2;
    __device__ static __forceinline__ uint4 window(const uint4& A, const uint4& B, int s) {
        if constexpr (sizeof(T) == 2) return window16(A, B, s);
        else return window32(A, B, s);
    }
    // element j of a packed vector as fp32
    __device__ static __forceinline__ float get(const uint4& v, int j) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if constexpr (sizeof(T) == 2) {
            const uint32_t word = w[j >> 1];
            return Elem<T>::bits_to_f32((uint16_t)((j & 1) ? (word >> 16) : (word & 0xffffu)));
        } else {
            return __uint_as_float(w[j]);
        }
    }
};

// number of entries <= val in a non-decreasing int16 array (upper bound)
__device__ __forceinline__ int upper_bound16(const short* a, int n, int val) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)a[mid] <= val) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace td
