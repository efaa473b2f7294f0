// Per-sampler-step hot path of tiled diffusion on sm_100a:
//   scatter  (crop-with-overlap of the latent into the tile batch)      multidiffusion.py:155
//   blend    (ordered overlap accumulate + normalise, gather form)      multidiffusion.py:166-167,208
//            (gaussian-weighted, pre-normalised accumulate)             mixtureofdiffusers.py:122-126
//
// Design (B200-first):
//   * ONE launch per step for all T tiles (the reference: T/TB cats + T..3T slice-adds).
//   * Gather form: each thread owns one 16-byte vector of the output canvas and
//     visits the tiles covering it in ascending tile index, rounding through the
//     canvas dtype after every add -> bit-identical to the reference's in-place
//     `x_buffer[slicer] += tile` sequence, with no atomics and no RMW traffic.
//   * Tile x-origins are arbitrary (46, 92, 231, ...), so tile rows and canvas rows
//     are mutually misaligned.  Every access is still a 128-bit aligned load: a
//     thread reads the two aligned 16-byte chunks that straddle its window and
//     extracts it with funnel shifts; the shift is uniform per tile column.
//   * Grid geometry (<= 256 row + 256 col origins) and the UNet's per-batch output
//     pointers travel in the kernel parameter block: no device tables, no H2D.
//   * All memory-bound: no tensor cores here on purpose.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "td_b200.h"
#include "td_device.cuh"
#include "td_internal.h"

namespace {

using namespace td;

struct GeomParams {
    int H, W, th, tw, rows, cols, N, C;
    float inv_dx, inv_dy;  // 1/stride estimates for the origin search (any value is safe)
    short ys[TD_MAX_GRID_DIM];
    short xs[TD_MAX_GRID_DIM];
};

struct BlendParams {
    GeomParams g;
    int tile_bs, num_batches;
    long long tile_stride;  // N*C*th*tw elements
    const void* batch_ptrs[TD_MAX_BATCH_PTRS];
};

enum { MODE_MD = 0, MODE_MOD = 1 };

// largest i with a[i] <= val (or -1); `a` non-decreasing.  The estimate only
// affects speed, never the result.
__device__ __forceinline__ int last_le(const short* a, int n, int val, float inv_d) {
    int i = min(n - 1, max(0, (int)((float)val * inv_d)));
    while (i + 1 < n && (int)a[i + 1] <= val) ++i;
    while (i >= 0 && (int)a[i] > val) --i;
    return i;
}

__device__ __forceinline__ void load_origins(const GeomParams& g, short* s_ys, short* s_xs) {
    for (int i = threadIdx.x; i < g.rows; i += blockDim.x) s_ys[i] = g.ys[i];
    for (int i = threadIdx.x; i < g.cols; i += blockDim.x) s_xs[i] = g.xs[i];
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Scatter, vector path: one thread = one aligned 16-byte store into the tile
// batch, fed by two aligned 16-byte loads of the (misaligned) canvas row.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
scatter_vec_kernel(const __grid_constant__ GeomParams g, const T* __restrict__ x, T* __restrict__ tiles,
                   int tile_begin, long long total_vecs) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_vecs) return;
    const int twv = g.tw >> L2V;
    int rem_i;
    long long rem = idx;
    const int uv = (int)(rem % twv); rem /= twv;
    const int v = (int)(rem % g.th); rem /= g.th;
    const int plane = (int)(rem % (g.N * g.C)); rem /= (g.N * g.C);  // n*C + c
    const int tl = (int)rem;                                        // tile index local to this launch
    (void)rem_i;
    const int t = tile_begin + tl;
    const int r = t / g.cols, c = t - r * g.cols;
    const int y = (int)g.ys[r] + v;
    const int u0 = (int)g.xs[c] + uv * VEC;  // first canvas column of this vector
    const T* row = x + ((long long)plane * g.H + y) * g.W;
    const int k = u0 >> L2V, s = u0 & (VEC - 1);
    const uint4 A = ldg128(row + (size_t)k * VEC);
    uint4 B = make_uint4(0, 0, 0, 0);
    if (s != 0) B = ldg128(row + (size_t)(k + 1) * VEC);  // in-bounds: it holds element u0+VEC-1 < W, W % VEC == 0
    const uint4 o = Vec<T>::window(A, B, s);
    stg128(tiles + idx * VEC, o);  // idx enumerates the tile batch in memory order
}

// Scatter, generic path: any width / alignment, one element per thread.
template <typename T>
__global__ void __launch_bounds__(256)
scatter_generic_kernel(const __grid_constant__ GeomParams g, const T* __restrict__ x, T* __restrict__ tiles,
                       int tile_begin, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    long long rem = idx;
    const int u = (int)(rem % g.tw); rem /= g.tw;
    const int v = (int)(rem % g.th); rem /= g.th;
    const int plane = (int)(rem % (g.N * g.C)); rem /= (g.N * g.C);
    const int t = tile_begin + (int)rem;
    const int r = t / g.cols, c = t - r * g.cols;
    tiles[idx] = x[((long long)plane * g.H + (int)g.ys[r] + v) * g.W + (int)g.xs[c] + u];
}

// ---------------------------------------------------------------------------
// Blend, vector path.
// ---------------------------------------------------------------------------
template <typename T, int MODE, bool WRITE_BUF>
__global__ void __launch_bounds__(256)
blend_grid_vec_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ weights,
                      const float* __restrict__ tile_weights, const float* __restrict__ rescale,
                      float* __restrict__ out_f32, T* __restrict__ out_buf, long long total_vecs) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    __shared__ short s_ys[TD_MAX_GRID_DIM];
    __shared__ short s_xs[TD_MAX_GRID_DIM];
    const GeomParams& g = p.g;
    load_origins(g, s_ys, s_xs);

    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_vecs) return;
    const int wv = g.W >> L2V;
    long long rem = idx;
    const int xv = (int)(rem % wv); rem /= wv;
    const int y = (int)(rem % g.H); rem /= g.H;
    const int plane = (int)rem;  // n*C + c
    const int x0 = xv * VEC;

    // covering tile rows for y, covering tile cols for any pixel of [x0, x0+VEC)
    const int r_lo = last_le(s_ys, g.rows, y - g.th, g.inv_dy) + 1;
    const int r_hi = last_le(s_ys, g.rows, y, g.inv_dy);
    const int c_lo = last_le(s_xs, g.cols, x0 - g.tw, g.inv_dx) + 1;
    const int c_hi = last_le(s_xs, g.cols, x0 + VEC - 1, g.inv_dx);

    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.0f;

    float rs[VEC];
    if constexpr (MODE == MODE_MOD) {
        const float4* rp = reinterpret_cast<const float4*>(rescale + (long long)y * g.W + x0);
#pragma unroll
        for (int q = 0; q < VEC / 4; ++q) {
            const float4 f = __ldg(rp + q);
            rs[4 * q + 0] = f.x; rs[4 * q + 1] = f.y; rs[4 * q + 2] = f.z; rs[4 * q + 3] = f.w;
        }
    }

    const int twv = g.tw >> L2V;
    for (int r = r_lo; r <= r_hi; ++r) {
        const int v = y - (int)s_ys[r];
        const long long row_off = ((long long)plane * g.th + v) * g.tw;
        for (int c = c_lo; c <= c_hi; ++c) {
            const int t = r * g.cols + c;
            const int b = t / p.tile_bs;
            const T* trow = reinterpret_cast<const T*>(p.batch_ptrs[b]) + (long long)(t - b * p.tile_bs) * p.tile_stride + row_off;
            const int u0 = x0 - (int)s_xs[c];  // tile-local column of element 0 (may be <0 or >tw-VEC at tile edges)
            const int k = u0 >> L2V, s = u0 & (VEC - 1);
            uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
            if (k >= 0 && k < twv) A = ldg128(trow + (long long)k * VEC);
            if (s != 0 && k + 1 >= 0 && k + 1 < twv) B = ldg128(trow + (long long)(k + 1) * VEC);
            const uint4 e = Vec<T>::window(A, B, s);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool valid = (unsigned)(u0 + j) < (unsigned)g.tw;
                float val = Vec<T>::get(e, j);
                if constexpr (MODE == MODE_MOD) {
                    // w = tile_weights * rescale_factor[slicer]; x_tile_out * w   (two fp32 roundings, no FMA)
                    const float tw_ = valid ? __ldg(tile_weights + (long long)v * g.tw + (u0 + j)) : 0.0f;
                    val = __fmul_rn(val, __fmul_rn(tw_, rs[j]));
                }
                const float sum = round_through<T>(__fadd_rn(acc[j], val));
                acc[j] = valid ? sum : acc[j];
            }
        }
    }

    const long long o = ((long long)plane * g.H + y) * g.W + x0;
    if constexpr (MODE == MODE_MD) {
        // x_out = where(weights > 1, x_buffer / weights, x_buffer)  -- fp32, IEEE divide
        const float4* wp = reinterpret_cast<const float4*>(weights + (long long)y * g.W + x0);
        float4* op = reinterpret_cast<float4*>(out_f32 + o);
#pragma unroll
        for (int q = 0; q < VEC / 4; ++q) {
            const float4 w = __ldg(wp + q);
            float4 f;
            f.x = w.x > 1.0f ? __fdiv_rn(acc[4 * q + 0], w.x) : acc[4 * q + 0];
            f.y = w.y > 1.0f ? __fdiv_rn(acc[4 * q + 1], w.y) : acc[4 * q + 1];
            f.z = w.z > 1.0f ? __fdiv_rn(acc[4 * q + 2], w.z) : acc[4 * q + 2];
            f.w = w.w > 1.0f ? __fdiv_rn(acc[4 * q + 3], w.w) : acc[4 * q + 3];
            op[q] = f;
        }
    }
    if constexpr (WRITE_BUF || MODE == MODE_MOD) {
        uint4 pk;
        if constexpr (sizeof(T) == 2) {
            uint32_t w32[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                w32[q] = (uint32_t)Elem<T>::f32_to_bits(acc[2 * q]) | ((uint32_t)Elem<T>::f32_to_bits(acc[2 * q + 1]) << 16);
            pk = make_uint4(w32[0], w32[1], w32[2], w32[3]);
        } else {
            pk = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
        }
        stg128(out_buf + o, pk);
    }
}

// Blend, generic path: one element per thread, any width / alignment, tile
// dtype and canvas dtype independent.
template <typename TIn, typename TAcc, int MODE>
__global__ void __launch_bounds__(256)
blend_grid_generic_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ weights,
                          const float* __restrict__ tile_weights, const float* __restrict__ rescale,
                          float* __restrict__ out_f32, TAcc* __restrict__ out_buf, long long total) {
    __shared__ short s_ys[TD_MAX_GRID_DIM];
    __shared__ short s_xs[TD_MAX_GRID_DIM];
    const GeomParams& g = p.g;
    load_origins(g, s_ys, s_xs);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    long long rem = idx;
    const int x = (int)(rem % g.W); rem /= g.W;
    const int y = (int)(rem % g.H); rem /= g.H;
    const int plane = (int)rem;
    const int r_lo = last_le(s_ys, g.rows, y - g.th, g.inv_dy) + 1;
    const int r_hi = last_le(s_ys, g.rows, y, g.inv_dy);
    const int c_lo = last_le(s_xs, g.cols, x - g.tw, g.inv_dx) + 1;
    const int c_hi = last_le(s_xs, g.cols, x, g.inv_dx);
    float acc = 0.0f;
    float rsc = 0.0f;
    if constexpr (MODE == MODE_MOD) rsc = rescale[(long long)y * g.W + x];
    for (int r = r_lo; r <= r_hi; ++r) {
        const int v = y - (int)s_ys[r];
        for (int c = c_lo; c <= c_hi; ++c) {
            const int t = r * g.cols + c;
            const int b = t / p.tile_bs;
            const int u = x - (int)s_xs[c];
            const TIn* tp = reinterpret_cast<const TIn*>(p.batch_ptrs[b]) + (long long)(t - b * p.tile_bs) * p.tile_stride +
                            ((long long)plane * g.th + v) * g.tw + u;
            float val = Elem<TIn>::to_f32(*tp);
            if constexpr (MODE == MODE_MOD) val = __fmul_rn(val, __fmul_rn(tile_weights[(long long)v * g.tw + u], rsc));
            acc = round_through<TAcc>(__fadd_rn(acc, val));
        }
    }
    if constexpr (MODE == MODE_MD) {
        const float w = weights[(long long)y * g.W + x];
        out_f32[idx] = w > 1.0f ? __fdiv_rn(acc, w) : acc;
    }
    if (out_buf != nullptr) out_buf[idx] = Elem<TAcc>::from_f32(acc);
}

// ---------------------------------------------------------------------------
// Host-side dispatch
// ---------------------------------------------------------------------------
int fill_geom(const td_grid* g, int N, int C, GeomParams* o) {
    if (g == nullptr) { td_set_error("null grid"); return TD_ERR_INVALID_ARG; }
    if (N <= 0 || C <= 0) { td_set_error("N and C must be positive (N=%d C=%d)", N, C); return TD_ERR_INVALID_ARG; }
    if (g->rows <= 0 || g->cols <= 0 || g->rows > TD_MAX_GRID_DIM || g->cols > TD_MAX_GRID_DIM ||
        g->num_tiles != g->rows * g->cols) {
        td_set_error("grid not initialised (rows=%d cols=%d)", g->rows, g->cols);
        return TD_ERR_INVALID_ARG;
    }
    if (g->H >= 32768 || g->W >= 32768) { td_set_error("canvas %dx%d too large", g->H, g->W); return TD_ERR_UNSUPPORTED; }
    o->H = g->H; o->W = g->W; o->th = g->tile_h; o->tw = g->tile_w;
    o->rows = g->rows; o->cols = g->cols; o->N = N; o->C = C;
    for (int i = 0; i < g->rows; ++i) o->ys[i] = (short)g->ys[i];
    for (int i = 0; i < g->cols; ++i) o->xs[i] = (short)g->xs[i];
    o->inv_dx = g->cols > 1 ? (float)(g->cols - 1) / (float)std::max(1, g->W - g->tile_w) : 0.0f;
    o->inv_dy = g->rows > 1 ? (float)(g->rows - 1) / (float)std::max(1, g->H - g->tile_h) : 0.0f;
    return TD_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        td_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return TD_ERR_CUDA;
    }
    return TD_OK;
}

inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

template <typename T>
int launch_scatter(const GeomParams& gp, const void* x, void* tiles, int tile_begin, int n_tiles, bool vec, cudaStream_t st) {
    const long long elems = (long long)n_tiles * gp.N * gp.C * gp.th * gp.tw;
    if (elems == 0) return TD_OK;
    if (vec) {
        const long long total = elems / Vec<T>::kElems;
        scatter_vec_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>(gp, (const T*)x, (T*)tiles, tile_begin, total);
    } else {
        scatter_generic_kernel<T><<<blocks_for(elems, 256), 256, 0, st>>>(gp, (const T*)x, (T*)tiles, tile_begin, elems);
    }
    return check_launch("td_scatter_tiles");
}

template <typename T, int MODE>
int launch_blend_vec(const BlendParams& bp, const float* weights, const float* tile_weights, const float* rescale,
                     float* out_f32, void* out_buf, cudaStream_t st) {
    const GeomParams& g = bp.g;
    const long long total = (long long)g.N * g.C * g.H * (g.W / Vec<T>::kElems);
    const unsigned nb = blocks_for(total, 256);
    if (MODE == MODE_MD && out_buf != nullptr)
        blend_grid_vec_kernel<T, MODE, true><<<nb, 256, 0, st>>>(bp, weights, tile_weights, rescale, out_f32, (T*)out_buf, total);
    else
        blend_grid_vec_kernel<T, MODE, false><<<nb, 256, 0, st>>>(bp, weights, tile_weights, rescale, out_f32, (T*)out_buf, total);
    return check_launch("td_blend (vec)");
}

template <typename TIn, typename TAcc, int MODE>
int launch_blend_generic(const BlendParams& bp, const float* weights, const float* tile_weights, const float* rescale,
                         float* out_f32, void* out_buf, cudaStream_t st) {
    const GeomParams& g = bp.g;
    const long long total = (long long)g.N * g.C * g.H * g.W;
    blend_grid_generic_kernel<TIn, TAcc, MODE><<<blocks_for(total, 256), 256, 0, st>>>(bp, weights, tile_weights, rescale,
                                                                                       out_f32, (TAcc*)out_buf, total);
    return check_launch("td_blend (generic)");
}

template <int MODE>
int dispatch_generic(int tile_dtype, int acc_dtype, const BlendParams& bp, const float* w, const float* tw, const float* rs,
                     float* out_f32, void* out_buf, cudaStream_t st) {
#define TD_CASE(TI, TA, TIn, TAcc) \
    if (tile_dtype == TI && acc_dtype == TA) return launch_blend_generic<TIn, TAcc, MODE>(bp, w, tw, rs, out_f32, out_buf, st);
    TD_CASE(TD_F16, TD_F16, __half, __half)
    TD_CASE(TD_F16, TD_BF16, __half, __nv_bfloat16)
    TD_CASE(TD_F16, TD_F32, __half, float)
    TD_CASE(TD_BF16, TD_F16, __nv_bfloat16, __half)
    TD_CASE(TD_BF16, TD_BF16, __nv_bfloat16, __nv_bfloat16)
    TD_CASE(TD_BF16, TD_F32, __nv_bfloat16, float)
    TD_CASE(TD_F32, TD_F16, float, __half)
    TD_CASE(TD_F32, TD_BF16, float, __nv_bfloat16)
    TD_CASE(TD_F32, TD_F32, float, float)
#undef TD_CASE
    td_set_error("unknown dtype pair (%d, %d)", tile_dtype, acc_dtype);
    return TD_ERR_INVALID_ARG;
}

int fill_blend(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N, int C, int tile_dtype,
               int acc_dtype, BlendParams* bp) {
    int st = fill_geom(g, N, C, &bp->g);
    if (st != TD_OK) return st;
    if (td_dtype_size(tile_dtype) == 0 || td_dtype_size(acc_dtype) == 0) { td_set_error("unknown dtype"); return TD_ERR_INVALID_ARG; }
    if (batch_ptrs == nullptr || num_batches <= 0 || tile_bs <= 0) { td_set_error("bad batch pointer table"); return TD_ERR_INVALID_ARG; }
    if (num_batches > TD_MAX_BATCH_PTRS) {
        td_set_error("%d batch tensors exceed TD_MAX_BATCH_PTRS=%d (concatenate them first)", num_batches, TD_MAX_BATCH_PTRS);
        return TD_ERR_UNSUPPORTED;
    }
    if ((long long)num_batches * tile_bs < g->num_tiles || (long long)(num_batches - 1) * tile_bs >= g->num_tiles) {
        td_set_error("batch table (%d x %d) does not cover %d tiles exactly", num_batches, tile_bs, g->num_tiles);
        return TD_ERR_INVALID_ARG;
    }
    for (int b = 0; b < num_batches; ++b) {
        if (batch_ptrs[b] == nullptr) { td_set_error("batch_ptrs[%d] is null", b); return TD_ERR_INVALID_ARG; }
        bp->batch_ptrs[b] = batch_ptrs[b];
    }
    bp->tile_bs = tile_bs;
    bp->num_batches = num_batches;
    bp->tile_stride = (long long)N * C * g->tile_h * g->tile_w;
    return TD_OK;
}

bool blend_vec_ok(const BlendParams& bp, int tile_dtype, int acc_dtype, std::initializer_list<const void*> ptrs) {
    if (tile_dtype != acc_dtype) return false;
    const int vec = 16 / td_dtype_size(tile_dtype);
    if (bp.g.W % vec != 0 || bp.g.tw % vec != 0) return false;
    for (int b = 0; b < bp.num_batches; ++b)
        if (!aligned16(bp.batch_ptrs[b])) return false;
    for (const void* p : ptrs)
        if (p != nullptr && !aligned16(p)) return false;
    return true;
}

}  // namespace

extern "C" int td_scatter_tiles(const td_grid* g, const void* x, void* tiles, int N, int C, int dtype, int tile_begin,
                                int tile_end, uint32_t flags, void* stream) {
    GeomParams gp;
    int st = fill_geom(g, N, C, &gp);
    if (st != TD_OK) return st;
    if (x == nullptr || tiles == nullptr) { td_set_error("td_scatter_tiles: null tensor"); return TD_ERR_INVALID_ARG; }
    if (tile_begin < 0 || tile_end < tile_begin || tile_end > g->num_tiles) {
        td_set_error("td_scatter_tiles: tile range [%d,%d) outside [0,%d)", tile_begin, tile_end, g->num_tiles);
        return TD_ERR_INVALID_ARG;
    }
    const int es = td_dtype_size(dtype);
    if (es == 0) { td_set_error("td_scatter_tiles: unknown dtype %d", dtype); return TD_ERR_INVALID_ARG; }
    const int vec = 16 / es;
    const bool vec_ok = !(flags & TD_FLAG_FORCE_GENERIC) && gp.W % vec == 0 && gp.tw % vec == 0 && aligned16(x) && aligned16(tiles);
    cudaStream_t s = (cudaStream_t)stream;
    // fp16 and bf16 are both moved as opaque 16-bit words
    if (es == 2) return launch_scatter<__half>(gp, x, tiles, tile_begin, tile_end - tile_begin, vec_ok, s);
    return launch_scatter<float>(gp, x, tiles, tile_begin, tile_end - tile_begin, vec_ok, s);
}

extern "C" int td_blend_multidiffusion(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N,
                                       int C, int tile_dtype, int acc_dtype, const float* weights, float* x_out,
                                       void* x_buffer, uint32_t flags, void* stream) {
    BlendParams bp;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (weights == nullptr || x_out == nullptr) { td_set_error("td_blend_multidiffusion: null weights / x_out"); return TD_ERR_INVALID_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (!(flags & TD_FLAG_FORCE_GENERIC) && blend_vec_ok(bp, tile_dtype, acc_dtype, {weights, x_out, x_buffer})) {
        switch (tile_dtype) {
            case TD_F16: return launch_blend_vec<__half, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
            case TD_BF16: return launch_blend_vec<__nv_bfloat16, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
            default: return launch_blend_vec<float, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
        }
    }
    return dispatch_generic<MODE_MD>(tile_dtype, acc_dtype, bp, weights, nullptr, nullptr, x_out, x_buffer, s);
}

extern "C" int td_blend_mixture(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N, int C,
                                int tile_dtype, int acc_dtype, const float* tile_weights, const float* rescale,
                                void* x_buffer, uint32_t flags, void* stream) {
    BlendParams bp;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (tile_weights == nullptr || rescale == nullptr || x_buffer == nullptr) {
        td_set_error("td_blend_mixture: null tile_weights / rescale / x_buffer");
        return TD_ERR_INVALID_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!(flags & TD_FLAG_FORCE_GENERIC) && blend_vec_ok(bp, tile_dtype, acc_dtype, {rescale, x_buffer})) {
        switch (tile_dtype) {
            case TD_F16: return launch_blend_vec<__half, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
            case TD_BF16: return launch_blend_vec<__nv_bfloat16, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
            default: return launch_blend_vec<float, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
        }
    }
    return dispatch_generic<MODE_MOD>(tile_dtype, acc_dtype, bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
}
