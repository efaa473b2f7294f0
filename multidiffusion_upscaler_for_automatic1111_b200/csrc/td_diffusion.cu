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
#include <atomic>
#include <mutex>

#include "td_b200.h"
#include "td_device.cuh"
#include "td_internal.h"
#include "td_tma.cuh"

namespace {

using namespace td;

struct GeomParams {
    int H, W, th, tw, rows, cols, N, C;
    float inv_dx, inv_dy;  // 1/stride estimates for the origin search (any value is safe)
    unsigned nc_magic, twv_magic;  // fast division by N*C and by tw/VEC (operands < 2^16)
    int dbg_no_tiles;              // TD_FLAG_DBG_NO_TILES: skip every tile visit (measures the launch + epilogue floor)
    short ys[TD_MAX_GRID_DIM];
    short xs[TD_MAX_GRID_DIM];
};

struct BlendParams {
    GeomParams g;
    int tile_bs, num_batches;
    unsigned bs_magic;  // fast division by tile_bs
    // cp.async kernel: tiles touching patch row / patch col i (first index, count), computed on the host
    unsigned char prow_lo[TD_MAX_GRID_DIM], prow_n[TD_MAX_GRID_DIM];
    unsigned char pcol_lo[TD_MAX_GRID_DIM], pcol_n[TD_MAX_GRID_DIM];
    const uint32_t* wait_flags;  // tile shard: spin until wait_flags[i] >= *wait_value for i < wait_world
    int wait_world;
    const uint32_t* wait_value;  // this rank's device-side step counter (bumped by td_peer_signal)
    int py0, py_count;      // cp.async kernels: first patch row and patch-row count of this launch (row-range blend; 0, 0 = all)
    int own_r_lo, own_r_hi; // row-strip shard: tile rows [lo, hi) are this rank's own outputs -- a CTA whose visits stay inside them
                            // skips the peer wait (lo == hi: every CTA waits)
    int sched;              // 0: blockIdx.y = patch row, .z = plane group; 1 / 2: blockIdx.z = patch row ascending / descending
                            // (slowest-varying), so that the CTAs that must wait for a neighbour's halo are scheduled last
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

// floor(n / d) for n, d < 2^16 with magic = ceil(2^32 / d); d == 1 is passed as magic 0.
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned magic) { return magic ? __umulhi(n, magic) : n; }

// ---------------------------------------------------------------------------
// Scatter, vector path.  One block = `rows_per_block` rows of one (tile, n, c)
// plane; tile origin and the funnel shift are block-uniform.  Each thread moves
// 16-byte vectors: two aligned 128-bit loads of the (misaligned) canvas row,
// one aligned 128-bit store into the tile batch; 4 vectors in flight per thread.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
scatter_vec_kernel(const __grid_constant__ GeomParams g, const T* __restrict__ x, T* __restrict__ tiles,
                   int tile_begin, int rows_per_block) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int UNROLL = 4;
    const unsigned tp = blockIdx.x;                       // (tile local, n, c) plane of the tile batch
    const unsigned tl = fastdiv(tp, g.nc_magic);
    const int plane = (int)(tp - tl * (unsigned)(g.N * g.C));
    const int t = tile_begin + (int)tl;
    const int r = t / g.cols, c = t - r * g.cols;
    const int xs = (int)g.xs[c], ys = (int)g.ys[r];
    const int twv = g.tw >> L2V;
    const int v0 = blockIdx.y * rows_per_block;
    const int nvec = min(rows_per_block, g.th - v0) * twv;
    const int s = xs & (VEC - 1), k0 = xs >> L2V;         // block-uniform shift
    const T* src = x + ((long long)plane * g.H + ys + v0) * g.W;
    T* dst = tiles + ((long long)tp * g.th + v0) * g.tw;
    for (int base = threadIdx.x; base < nvec; base += 128 * UNROLL) {
        uint4 A[UNROLL], B[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const int i = base + q * 128;
            A[q] = B[q] = make_uint4(0, 0, 0, 0);
            if (i < nvec) {
                const unsigned vi = fastdiv((unsigned)i, g.twv_magic);
                const int uv = i - (int)vi * twv;
                const T* row = src + (long long)vi * g.W + (long long)(k0 + uv) * VEC;
                A[q] = ldg128(row);
                if (s != 0) B[q] = ldg128(row + VEC);   // in-bounds: holds element xs+uv*VEC+VEC-1 < W, W % VEC == 0
            }
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const int i = base + q * 128;
            if (i < nvec) stg128(dst + (long long)i * VEC, Vec<T>::window(A[q], B[q], s));
        }
    }
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
//
// Work decomposition: a warp owns an 8-vector x 8-row patch of one canvas plane
// (lane = 8 vectors x 4 rows, 2 rows per thread); a block is 4 warps stacked in y.
// The covering tile rows / cols are computed once per warp (uniform loops); a lane
// takes part in a tile visit through predicated loads only.  Out-of-tile chunks are
// never loaded and read as -0.0, the exact additive identity (x + -0.0 == x for
// every x, signed zeros included), and because tw % VEC == 0 chunk validity IS
// element validity -- so the MultiDiffusion path needs no per-element masks and
// accumulates with packed HADD2: round_half(a + b) of the exact sum, which equals
// torch's half(float(a) + float(b)) (fp32 has >= 2p+2 bits: double rounding is innocuous).
// ---------------------------------------------------------------------------
template <typename T> struct PackedAdd;  // acc (+)= e on one 32-bit word of packed T
template <> struct PackedAdd<__half> {
    __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
        __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
};
template <> struct PackedAdd<__nv_bfloat16> {
    __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
        __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
};
template <> struct PackedAdd<float> {
    __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
        return __float_as_uint(__fadd_rn(__uint_as_float(a), __uint_as_float(b)));
    }
};

constexpr int kBlendWarps = 2;       // warps per block, stacked in y
constexpr int kWarpVecs = 8;         // vectors per warp row (lane = 8 vectors x 4 rows)
constexpr int kGroup = 2;            // tile visits whose loads are in flight together

template <int MODE> struct BlendShape { static constexpr int kRPT = (MODE == 0) ? 4 : 2; };  // rows per thread

template <typename T, int MODE, bool WRITE_BUF>
__global__ void __launch_bounds__(kBlendWarps * 32)
blend_grid_vec_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ weights,
                      const float* __restrict__ tile_weights, const float* __restrict__ rescale,
                      float* __restrict__ out_f32, T* __restrict__ out_buf) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int RPT = BlendShape<MODE>::kRPT;
    constexpr int WROWS = 4 * RPT;                 // rows per warp patch
    constexpr uint32_t NEG0 = sizeof(T) == 2 ? 0x80008000u : 0x80000000u;
    __shared__ short s_ys[TD_MAX_GRID_DIM];
    __shared__ short s_xs[TD_MAX_GRID_DIM];
    const GeomParams& g = p.g;
    if (p.wait_flags != nullptr) {   // tile shard: peers' tile outputs must be complete before they are read
        if ((int)threadIdx.x < p.wait_world) {
            const uint32_t want = *p.wait_value;
            uint32_t v;
            unsigned long long spins = 0;
            do {   // bounded: a peer that never signals (crashed / interrupted rank) becomes a launch error, not a hung GPU
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_flags + threadIdx.x) : "memory");
                if (++spins > (1ull << 31)) __trap();
            } while ((int32_t)(v - want) < 0);
        }
    }
    load_origins(g, s_ys, s_xs);   // (contains the __syncthreads that publishes the wait)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & 7, ly = lane >> 3;
    const int plane = blockIdx.z;
    const int wv = g.W >> L2V;
    const int xv0 = blockIdx.x * kWarpVecs;                           // warp's first vector
    const int yw = (blockIdx.y * kBlendWarps + warp) * WROWS;         // warp's first row
    if (yw >= g.H) return;                                            // warp-uniform
    const int x_lo = xv0 * VEC, x_hi = min(x_lo + kWarpVecs * VEC, g.W) - 1;
    const int y_hi = min(yw + WROWS, g.H) - 1;

    // tiles touching this warp's patch (uniform across the warp)
    const int r_lo = last_le(s_ys, g.rows, yw - g.th, g.inv_dy) + 1;
    const int r_hi = last_le(s_ys, g.rows, y_hi, g.inv_dy);
    const int c_lo = last_le(s_xs, g.cols, x_lo - g.tw, g.inv_dx) + 1;
    const int c_hi = last_le(s_xs, g.cols, x_hi, g.inv_dx);
    const int nc = c_hi - c_lo + 1;
    const int nv = g.dbg_no_tiles ? 0 : (r_hi - r_lo + 1) * nc;       // visits, in ascending tile index

    const int xv = xv0 + lx;
    const bool x_ok = xv < wv;
    const int x0 = xv * VEC;
    int y[RPT];
    bool y_ok[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) { y[q] = yw + ly + 4 * q; y_ok[q] = x_ok && y[q] < g.H; }

    constexpr bool kPacked = (MODE == MODE_MD);   // accumulate directly in T (packed words)
    uint4 pacc[RPT];
    float facc[RPT][kPacked ? 1 : VEC];
    float rs[RPT][kPacked ? 1 : VEC];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        pacc[q] = make_uint4(0, 0, 0, 0);
        if constexpr (!kPacked) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) { facc[q][j] = 0.0f; rs[q][j] = 0.0f; }
            if (y_ok[q]) {
                const float4* rp = reinterpret_cast<const float4*>(rescale + (long long)y[q] * g.W + x0);
#pragma unroll
                for (int h = 0; h < VEC / 4; ++h) {
                    const float4 f = __ldg(rp + h);
                    rs[q][4 * h + 0] = f.x; rs[q][4 * h + 1] = f.y; rs[q][4 * h + 2] = f.z; rs[q][4 * h + 3] = f.w;
                }
            }
        }
    }

    const int twv = g.tw >> L2V;
    const int plane_off = plane * g.th * g.tw;     // < 2^31: checked on the host
    int r_cur = r_lo, c_cur = c_lo;
    for (int i0 = 0; i0 < nv; i0 += kGroup) {
        uint4 A[kGroup][RPT], B[kGroup][RPT];
        int sft[kGroup], u0s[kGroup], vrow[kGroup][RPT];
        bool live[kGroup][RPT];
        // ---- phase 1: issue every load of the group -----------------------------------------
#pragma unroll
        for (int gi = 0; gi < kGroup; ++gi) {
            const bool in = i0 + gi < nv;
            const int r = in ? r_cur : r_lo, c = in ? c_cur : c_lo;
            if (in) { if (++c_cur > c_hi) { c_cur = c_lo; ++r_cur; } }
            const unsigned t = (unsigned)(r * g.cols + c);
            const unsigned b = fastdiv(t, p.bs_magic);
            const T* tbase = reinterpret_cast<const T*>(p.batch_ptrs[b]) + (long long)(t - b * (unsigned)p.tile_bs) * p.tile_stride;
            const int ysr = (int)s_ys[r];
            const int u0 = x0 - (int)s_xs[c];      // tile-local column of element 0; u0 & (VEC-1) is warp-uniform
            const int k = u0 >> L2V, s = u0 & (VEC - 1);
            sft[gi] = s; u0s[gi] = u0;
            const bool okA = in && (unsigned)k < (unsigned)twv;
            const bool okB = in && s != 0 && (unsigned)(k + 1) < (unsigned)twv;
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int v = y[q] - ysr;
                const bool rok = y_ok[q] && (unsigned)v < (unsigned)g.th;
                vrow[gi][q] = v; live[gi][q] = rok && in;
                A[gi][q] = make_uint4(NEG0, NEG0, NEG0, NEG0);
                B[gi][q] = A[gi][q];
                const T* trow = tbase + (plane_off + v * g.tw + k * VEC);
                if (rok && okA) A[gi][q] = ldg128(trow);
                if (rok && okB) B[gi][q] = ldg128(trow + VEC);
            }
        }
        // ---- phase 2: consume in tile order ---------------------------------------------------
#pragma unroll
        for (int gi = 0; gi < kGroup; ++gi) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const uint4 e = Vec<T>::window(A[gi][q], B[gi][q], sft[gi]);
                if constexpr (kPacked) {
                    pacc[q].x = PackedAdd<T>::add(pacc[q].x, e.x);
                    pacc[q].y = PackedAdd<T>::add(pacc[q].y, e.y);
                    pacc[q].z = PackedAdd<T>::add(pacc[q].z, e.z);
                    pacc[q].w = PackedAdd<T>::add(pacc[q].w, e.w);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const int u = u0s[gi] + j;
                        const bool valid = live[gi][q] && (unsigned)u < (unsigned)g.tw;
                        // w = tile_weights * rescale_factor[slicer]; x_tile_out * w   (two fp32 roundings, no FMA)
                        const float tw_ = valid ? __ldg(tile_weights + vrow[gi][q] * g.tw + u) : 0.0f;
                        const float val = __fmul_rn(Vec<T>::get(e, j), __fmul_rn(tw_, rs[q][j]));
                        const float sum = round_through<T>(__fadd_rn(facc[q][j], val));
                        facc[q][j] = valid ? sum : facc[q][j];
                    }
                }
            }
        }
    }

#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        if (!y_ok[q]) continue;
        const long long o = ((long long)plane * g.H + y[q]) * g.W + x0;
        float acc[VEC];
        if constexpr (kPacked) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = Vec<T>::get(pacc[q], j);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = facc[q][j];
        }
        if constexpr (MODE == MODE_MD) {
            // x_out = where(weights > 1, x_buffer / weights, x_buffer)  -- fp32, IEEE divide
            const float4* wp = reinterpret_cast<const float4*>(weights + (long long)y[q] * g.W + x0);
            float4* op = reinterpret_cast<float4*>(out_f32 + o);
#pragma unroll
            for (int h = 0; h < VEC / 4; ++h) {
                const float4 w = __ldg(wp + h);
                float4 f;
                f.x = w.x > 1.0f ? __fdiv_rn(acc[4 * h + 0], w.x) : acc[4 * h + 0];
                f.y = w.y > 1.0f ? __fdiv_rn(acc[4 * h + 1], w.y) : acc[4 * h + 1];
                f.z = w.z > 1.0f ? __fdiv_rn(acc[4 * h + 2], w.z) : acc[4 * h + 2];
                f.w = w.w > 1.0f ? __fdiv_rn(acc[4 * h + 3], w.w) : acc[4 * h + 3];
                op[h] = f;
            }
        }
        if constexpr (WRITE_BUF || MODE == MODE_MOD) {
            uint4 pk;
            if constexpr (kPacked) {
                pk = pacc[q];
            } else if constexpr (sizeof(T) == 2) {
                uint32_t w32[4];
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    w32[h] = (uint32_t)Elem<T>::f32_to_bits(acc[2 * h]) | ((uint32_t)Elem<T>::f32_to_bits(acc[2 * h + 1]) << 16);
                pk = make_uint4(w32[0], w32[1], w32[2], w32[3]);
            } else {
                pk = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
            }
            stg128(out_buf + o, pk);
        }
    }
}

// ---------------------------------------------------------------------------
// Blend + normalise, MultiDiffusion, cp.async path (the default when it applies).
//
// A CTA owns a 8-row x 8-vector (TD_AS_ROWS) patch of one canvas plane.  For every tile touching the
// patch (ascending tile index) the threads copy the aligned superset of the intersection
// (9 chunks x 8 rows) global -> shared with 16-byte cp.async; chunks outside the tile
// use src-size 0, i.e. the copy itself zero-fills them (exact: the accumulator starts at
// +0 and can never become -0).  Every copy of the CTA is in flight before the first wait:
// memory-level parallelism without staging registers and without the per-box cost of TMA.
// The consume loop is specialised on the per-tile shift s (uniform over the CTA): two
// LDS.128, at most four funnel shifts with an immediate, four packed adds.
// With integer weights (MultiDiffusion's always are) and their correctly rounded
// reciprocals supplied, the IEEE divide is q = a*rcp; r = fma(-q, w, a); q' = fma(r, rcp, q)
// -- correctly rounded for every 16-bit numerator and w <= 4096 (checked exhaustively by
// td_debug_check_fast_div / tests).
// ---------------------------------------------------------------------------
#ifndef TD_AS_X
#define TD_AS_X 8                                // tuning knob (-DTD_AS_X=): vectors per patch row
#endif
constexpr int kAsX = TD_AS_X;                    // vectors per patch row
#ifndef TD_AS_ROWS
#define TD_AS_ROWS 8                             // tuning knob (-DTD_AS_ROWS=): 8 / 16 / 32 measured 8.89 / 9.18 / 10.7 us (cfg2), 8 kept
#endif
constexpr int kAsY = TD_AS_ROWS;                 // patch rows
constexpr int kAsVecThreads = kAsX * kAsY;       // 64 threads own one output vector each
constexpr int kAsChunks = kAsX + 1;              // staged chunks per row (aligned superset)
constexpr int kAsSlots = kAsY * kAsChunks;       // 72 copy slots per tile visit: one per thread
constexpr int kAsThreads = (kAsSlots + 31) / 32 * 32;   // 96 = 3 warps: 72 copy threads
constexpr int kAsStage = kAsSlots * 16;          // 1152 bytes per tile visit

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// 16-byte copy that writes zeros instead (no global access) when `neg_if_skip` is negative
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, long long gmem_src, int neg_if_skip) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.lt.s32 p, %2, 0;\n\t"
        "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t}"
        ::"r"(smem_dst), "l"(gmem_src), "r"(neg_if_skip)
        : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <typename T, int S>
__device__ __forceinline__ void consume_shifted(uint4& acc, const unsigned char* p) {
    const uint4 A = lds128(p);
    uint4 e;
    if constexpr (S == 0) {
        e = A;
    } else {
        const uint4 B = lds128(p + 16);
        e = Vec<T>::window(A, B, S);    // S is a compile-time constant: register selection / immediate shifts only
    }
    acc.x = PackedAdd<T>::add(acc.x, e.x);
    acc.y = PackedAdd<T>::add(acc.y, e.y);
    acc.z = PackedAdd<T>::add(acc.z, e.z);
    acc.w = PackedAdd<T>::add(acc.w, e.w);
}

__device__ __forceinline__ float div_exact_small_int(float a, float w, float rcp) {
    const float q = __fmul_rn(a, rcp);
    const float r = __fmaf_rn(-q, w, a);
    const float q2 = __fmaf_rn(r, rcp, q);
    // w > 0: the quotient carries a's sign (this also keeps -0 / w = -0, which the fma chain turns into +0)
    return __uint_as_float((__float_as_uint(q2) & 0x7fffffffu) | (__float_as_uint(a) & 0x80000000u));
}

// Programmatic dependent launch (sm_90+): a kernel launched with the stream-serialisation attribute may become
// resident while its predecessor in the stream drains.  pdl_wait() blocks until the predecessor has completed and
// its memory is visible -- it precedes every global read or write below, so stream order is kept; what overlaps
// is the launch itself, the parameter fetch and the index arithmetic.  Both are no-ops for an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct __align__(16) VisitEntry {   // per tile visit of a CTA, computed once by one thread
    long long origin;               // byte address of tile element (v0, k0*VEC) of this plane (may lie before the tile)
    int v0, k0;                     // tile row / tile chunk of the patch origin (may be negative)
};
constexpr int kAsMaxVisits = 64;

template <typename T, int S, int PPC>
__device__ __forceinline__ void consume_planes(uint4 (&acc)[PPC], const unsigned char* p) {
#pragma unroll
    for (int q = 0; q < PPC; ++q) consume_shifted<T, S>(acc[q], p + q * kAsStage);
}

// PPC = planes per CTA: the (n, c) planes of a patch share the visit table, the zero-fill predicate, the weights
// and the shift, so a CTA that blends PPC planes amortises its prologue and runs PPC independent add chains.
template <typename T, bool WRITE_BUF, bool FASTDIV, int PPC>
__global__ void __launch_bounds__(kAsThreads)
blend_md_async_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ weights, const float* __restrict__ rcp_weights,
                      float* __restrict__ out_f32, T* __restrict__ out_buf) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int BX = kAsX * VEC;
    extern __shared__ __align__(16) unsigned char td_smem[];
    __shared__ VisitEntry s_visit[kAsMaxVisits];
    __shared__ int s_shift[kAsMaxVisits];
    const GeomParams& g = p.g;
    const int tid = threadIdx.x;
    const int zy = p.sched == 0 ? (int)blockIdx.y : (p.sched == 2 ? p.py_count - 1 - (int)blockIdx.z : (int)blockIdx.z);
    const int plane = (p.sched == 0 ? (int)blockIdx.z : (int)blockIdx.y) * PPC;   // first plane of this CTA (PPC divides N*C)
    const int by = zy + p.py0;                // patch row (a row-range launch starts at py0)
    const int x_lo = blockIdx.x * BX, y_lo = by * kAsY;
    pdl_launch_dependents();

    // tiles touching this patch: host-computed per patch row / col (uniform constant-bank reads)
    const int r_lo = p.prow_lo[by], c_lo = p.pcol_lo[blockIdx.x];
    const int nc = p.pcol_n[blockIdx.x];
    const int nv = g.dbg_no_tiles ? 0 : (int)p.prow_n[by] * nc;   // <= the host's stage count

    // this thread's output vector; its weights are fetched now so that their latency hides behind the tile copies
    const int tx = tid % kAsX, ty = tid / kAsX;
    const int x0 = x_lo + tx * VEC, y = y_lo + ty;
    const bool inside = tid < kAsVecThreads && x0 < g.W && y < g.H;
    const long long wo = (long long)y * g.W + x0;
    float4 wv[VEC / 4], rv[VEC / 4];
    pdl_wait();
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
        wv[h] = inside ? __ldg(reinterpret_cast<const float4*>(weights + wo) + h) : make_float4(1.f, 1.f, 1.f, 1.f);
        if constexpr (FASTDIV) rv[h] = inside ? __ldg(reinterpret_cast<const float4*>(rcp_weights + wo) + h) : make_float4(1.f, 1.f, 1.f, 1.f);
    }

    const bool reads_halo = r_lo < p.own_r_lo || r_lo + (int)p.prow_n[by] > p.own_r_hi;   // always true when own_r_lo == own_r_hi
    if (p.wait_flags != nullptr && reads_halo && tid >= kAsThreads - 32 && tid - (kAsThreads - 32) < p.wait_world) {
        // tile shard: the peers' tile outputs must be complete before any copy reads them (last warp spins,
        // the barrier below publishes it to the CTA)
        const uint32_t want = *p.wait_value;
        uint32_t v;
        unsigned long long spins = 0;
        do {   // bounded: a peer that never signals becomes a launch error, not a hung GPU
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_flags + (tid - (kAsThreads - 32))) : "memory");
            if (++spins > (1ull << 31)) __trap();
        } while ((int32_t)(v - want) < 0);
    }

    // ---- per-visit constants: thread i prepares visit i -------------------------------------------------
    const long long plane_bytes = (long long)g.th * g.tw * (long long)sizeof(T);
    if (tid < nv) {
        const int ri = tid / nc, ci = tid - ri * nc;
        const int r = r_lo + ri, c = c_lo + ci;
        const unsigned t = (unsigned)(r * g.cols + c);
        const unsigned b = fastdiv(t, p.bs_magic);
        const int u0 = x_lo - (int)g.xs[c];
        const int v0 = y_lo - (int)g.ys[r], k0 = u0 >> L2V;
        const long long elem = (long long)(t - b * (unsigned)p.tile_bs) * p.tile_stride + (long long)v0 * g.tw + (long long)k0 * VEC;
        VisitEntry e;
        e.origin = (long long)reinterpret_cast<uintptr_t>(p.batch_ptrs[b]) + elem * (long long)sizeof(T) + (long long)plane * plane_bytes;
        e.v0 = v0;
        e.k0 = k0;
        s_visit[tid] = e;
        s_shift[tid] = u0 & (VEC - 1);
    }
    __syncthreads();

    // ---- issue every copy of this CTA: src = origin(visit) + offset(thread) + plane, zero-fill outside the tile -----
    if (tid < kAsSlots) {
        const int row = tid / kAsChunks, j = tid - row * kAsChunks;    // this thread's copy slot (same in every stage)
        const int th1 = g.th - 1 - row, tw1 = (g.tw >> L2V) - 1 - j;   // v0 <= th1 and k0 <= tw1 <=> inside the tile (upper bounds)
        const long long off = ((long long)row * g.tw + (long long)j * VEC) * (long long)sizeof(T);
        uint32_t dst = smem_u32(td_smem) + (uint32_t)tid * 16u;
#pragma unroll 4
        for (int i = 0; i < nv; ++i, dst += PPC * kAsStage) {
            const VisitEntry e = s_visit[i];
            // negative (= outside the tile) iff v < 0, v >= th, k < 0 or k >= twv, with v = v0 + row, k = k0 + j
            const int skip = (e.v0 + row) | (th1 - e.v0) | (e.k0 + j) | (tw1 - e.k0);
#pragma unroll
            for (int q = 0; q < PPC; ++q) cp_async16_zfill(dst + q * kAsStage, e.origin + off + q * plane_bytes, skip);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (tid >= kAsVecThreads) return;

    // ---- consume in tile order ------------------------------------------------------------------------
    const unsigned char* mine = td_smem + (ty * kAsChunks + tx) * 16;
    uint4 acc[PPC];
#pragma unroll
    for (int q = 0; q < PPC; ++q) acc[q] = make_uint4(0, 0, 0, 0);
    {
        for (int i = 0; i < nv; ++i, mine += PPC * kAsStage) {
            const int s = s_shift[i];   // uniform over the CTA
            if constexpr (VEC == 8) {
                switch (s) {
                    case 0: consume_planes<T, 0, PPC>(acc, mine); break;
                    case 1: consume_planes<T, 1, PPC>(acc, mine); break;
                    case 2: consume_planes<T, 2, PPC>(acc, mine); break;
                    case 3: consume_planes<T, 3, PPC>(acc, mine); break;
                    case 4: consume_planes<T, 4, PPC>(acc, mine); break;
                    case 5: consume_planes<T, 5, PPC>(acc, mine); break;
                    case 6: consume_planes<T, 6, PPC>(acc, mine); break;
                    default: consume_planes<T, 7, PPC>(acc, mine); break;
                }
            } else {
                switch (s) {
                    case 0: consume_planes<T, 0, PPC>(acc, mine); break;
                    case 1: consume_planes<T, 1, PPC>(acc, mine); break;
                    case 2: consume_planes<T, 2, PPC>(acc, mine); break;
                    default: consume_planes<T, 3, PPC>(acc, mine); break;
                }
            }
        }
    }

    // ---- normalise + store --------------------------------------------------------------------------
    if (!inside) return;
#pragma unroll
    for (int q = 0; q < PPC; ++q) {
        const long long o = ((long long)(plane + q) * g.H + y) * g.W + x0;
        float4* op = reinterpret_cast<float4*>(out_f32 + o);
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {
            const float4 w = wv[h];
            const float a0 = Vec<T>::get(acc[q], 4 * h + 0), a1 = Vec<T>::get(acc[q], 4 * h + 1);
            const float a2 = Vec<T>::get(acc[q], 4 * h + 2), a3 = Vec<T>::get(acc[q], 4 * h + 3);
            float4 f;   // x_out = where(weights > 1, x_buffer / weights, x_buffer)  -- fp32, correctly rounded divide
            if constexpr (FASTDIV) {
                const float4 rc = rv[h];
                f.x = w.x > 1.0f ? div_exact_small_int(a0, w.x, rc.x) : a0;
                f.y = w.y > 1.0f ? div_exact_small_int(a1, w.y, rc.y) : a1;
                f.z = w.z > 1.0f ? div_exact_small_int(a2, w.z, rc.z) : a2;
                f.w = w.w > 1.0f ? div_exact_small_int(a3, w.w, rc.w) : a3;
            } else {
                f.x = w.x > 1.0f ? __fdiv_rn(a0, w.x) : a0;
                f.y = w.y > 1.0f ? __fdiv_rn(a1, w.y) : a1;
                f.z = w.z > 1.0f ? __fdiv_rn(a2, w.z) : a2;
                f.w = w.w > 1.0f ? __fdiv_rn(a3, w.w) : a3;
            }
            op[h] = f;
        }
        if constexpr (WRITE_BUF) stg128(out_buf + o, acc[q]);
    }
}

// ---------------------------------------------------------------------------
// Blend, Mixture of Diffusers, cp.async path: same staging as blend_md_async_kernel; the consume
// step reproduces mixtureofdiffusers.py:125-126 per element --
//     w = tile_weights[v, u] * rescale[y, x]      (fp32 product, own rounding)
//     acc = round_T(float(acc) + float(tile) * w) (separate multiply and add, no FMA)
// The gaussian tile weights are read with three aligned 128-bit loads per vector (the 4-float shift
// is uniform over the CTA).  Elements outside the tile are skipped explicitly (see mod_accumulate).
// ---------------------------------------------------------------------------
// [jlo, jhi): the elements of this vector that lie inside the tile.  Elements outside must not be touched at all:
// with gaussian weights the accumulator CAN be -0.0 (a tiny negative sum rounds to -0 in fp16), and -0 + (+0) = +0.
template <typename T, int O>
__device__ __forceinline__ void mod_accumulate(float (&acc)[Vec<T>::kElems], const uint4& e, const float4& w0, const float4& w1,
                                               const float4& w2, const float (&rs)[Vec<T>::kElems], int jlo, int jhi) {
    constexpr int VEC = Vec<T>::kElems;
    const float wl[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
    if (jlo == 0 && jhi == VEC) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float w = __fmul_rn(wl[O + j], rs[j]);
            const float val = __fmul_rn(Vec<T>::get(e, j), w);
            acc[j] = round_through<T>(__fadd_rn(acc[j], val));
        }
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float w = __fmul_rn(wl[O + j], rs[j]);
            const float val = __fmul_rn(Vec<T>::get(e, j), w);
            const float sum = round_through<T>(__fadd_rn(acc[j], val));
            acc[j] = (j >= jlo && j < jhi) ? sum : acc[j];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kAsThreads)
blend_mod_async_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ tile_weights, const float* __restrict__ rescale,
                       T* __restrict__ out_buf) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int BX = kAsX * VEC;
    extern __shared__ __align__(16) unsigned char td_smem[];
    __shared__ VisitEntry s_visit[kAsMaxVisits];
    __shared__ int s_shift[kAsMaxVisits];
    const GeomParams& g = p.g;
    const int tid = threadIdx.x;
    const int plane = blockIdx.z;
    const int x_lo = blockIdx.x * BX, y_lo = blockIdx.y * kAsY;
    const int r_lo = p.prow_lo[blockIdx.y], c_lo = p.pcol_lo[blockIdx.x];
    const int nc = p.pcol_n[blockIdx.x];
    const int nv = (int)p.prow_n[blockIdx.y] * nc;
    pdl_launch_dependents();

    const int tx = tid % kAsX, ty = tid / kAsX;
    const int x0 = x_lo + tx * VEC, y = y_lo + ty;
    const bool inside = tid < kAsVecThreads && x0 < g.W && y < g.H;
    float rs[VEC];
    pdl_wait();
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
        const float4 f = inside ? __ldg(reinterpret_cast<const float4*>(rescale + (long long)y * g.W + x0) + h) : make_float4(0.f, 0.f, 0.f, 0.f);
        rs[4 * h + 0] = f.x; rs[4 * h + 1] = f.y; rs[4 * h + 2] = f.z; rs[4 * h + 3] = f.w;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)   // 1/0 = inf marks a pixel no tile covers: it receives only zero fill, keep 0 * w finite
        if (isinf(rs[j])) rs[j] = 0.0f;

    if (tid < nv) {
        const int ri = tid / nc, ci = tid - ri * nc;
        const int r = r_lo + ri, c = c_lo + ci;
        const unsigned t = (unsigned)(r * g.cols + c);
        const unsigned b = fastdiv(t, p.bs_magic);
        const int u0 = x_lo - (int)g.xs[c];
        const int v0 = y_lo - (int)g.ys[r], k0 = u0 >> L2V;
        const long long elem = (long long)(t - b * (unsigned)p.tile_bs) * p.tile_stride + (long long)plane * g.th * g.tw +
                               (long long)v0 * g.tw + (long long)k0 * VEC;
        VisitEntry e;
        e.origin = (long long)reinterpret_cast<uintptr_t>(p.batch_ptrs[b]) + elem * (long long)sizeof(T);
        e.v0 = v0;
        e.k0 = k0;
        s_visit[tid] = e;
        s_shift[tid] = u0 & (VEC - 1);
    }
    __syncthreads();
    if (tid < kAsSlots) {
        const int row = tid / kAsChunks, j = tid - row * kAsChunks;
        const int th1 = g.th - 1 - row, tw1 = (g.tw >> L2V) - 1 - j;
        const long long off = ((long long)row * g.tw + (long long)j * VEC) * (long long)sizeof(T);
        uint32_t dst = smem_u32(td_smem) + (uint32_t)tid * 16u;
#pragma unroll 4
        for (int i = 0; i < nv; ++i, dst += kAsStage) {
            const VisitEntry e = s_visit[i];
            const int skip = (e.v0 + row) | (th1 - e.v0) | (e.k0 + j) | (tw1 - e.k0);
            cp_async16_zfill(dst, e.origin + off, skip);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (!inside) return;

    const unsigned char* mine = td_smem + (ty * kAsChunks + tx) * 16;
    const int tw4 = g.tw >> 2;   // tile-weight row in float4 chunks (tw % 8 == 0 on this path)
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.0f;
    for (int i = 0; i < nv; ++i, mine += kAsStage) {
        const VisitEntry e = s_visit[i];
        const int s = s_shift[i];
        const int v = e.v0 + ty;
        if ((unsigned)v >= (unsigned)g.th) continue;        // this row of the stage is all zero fill: nothing to add
        const uint4 A = lds128(mine);
        uint4 ev = A;
        if (s != 0) ev = Vec<T>::window(A, lds128(mine + 16), s);
        const int u = e.k0 * VEC + s + tx * VEC;            // tile column of element 0 (may be outside the tile)
        const int q0 = u >> 2, o = u & 3;                   // o == s & 3: uniform over the CTA
        const float4* wrow = reinterpret_cast<const float4*>(tile_weights + (long long)v * g.tw);
        const float4 w0 = __ldg(wrow + min(max(q0, 0), tw4 - 1));
        const float4 w1 = __ldg(wrow + min(max(q0 + 1, 0), tw4 - 1));
        const float4 w2 = __ldg(wrow + min(max(q0 + 2, 0), tw4 - 1));
        const int jlo = max(0, -u), jhi = min(VEC, g.tw - u);
        if (jhi <= jlo) continue;                           // vector entirely outside this tile
        switch (o) {
            case 0: mod_accumulate<T, 0>(acc, ev, w0, w1, w2, rs, jlo, jhi); break;
            case 1: mod_accumulate<T, 1>(acc, ev, w0, w1, w2, rs, jlo, jhi); break;
            case 2: mod_accumulate<T, 2>(acc, ev, w0, w1, w2, rs, jlo, jhi); break;
            default: mod_accumulate<T, 3>(acc, ev, w0, w1, w2, rs, jlo, jhi); break;
        }
    }

    const long long o_ = ((long long)plane * g.H + y) * g.W + x0;
    uint4 pk;
    if constexpr (sizeof(T) == 2) {
        uint32_t w32[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            w32[h] = (uint32_t)Elem<T>::f32_to_bits(acc[2 * h]) | ((uint32_t)Elem<T>::f32_to_bits(acc[2 * h + 1]) << 16);
        pk = make_uint4(w32[0], w32[1], w32[2], w32[3]);
    } else {
        pk = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
    }
    stg128(out_buf + o_, pk);
}

// ---------------------------------------------------------------------------
// Persistent, software-pipelined form of the cp.async blend (TD_FLAG_PIPELINE, opt-in): a CTA walks
// patches pi = blockIdx.x, += gridDim.x with two stage sets -- the copies of patch n+1
// are in flight while patch n is consumed, normalised and stored, so DRAM stays busy
// instead of all CTAs of a wave issuing, waiting and computing in lock step.  Measured on B200
// (cfg2): 13.9 us vs 10.0 us for one-patch-per-CTA -- 15 resident warps per SM and three
// barriers per patch lose more than the overlap wins -- so it is not the default.
// ---------------------------------------------------------------------------
template <typename T, bool WRITE_BUF, bool FASTDIV>
__global__ void __launch_bounds__(kAsThreads)
blend_md_pipe_kernel(const __grid_constant__ BlendParams p, const float* __restrict__ weights, const float* __restrict__ rcp_weights,
                     float* __restrict__ out_f32, T* __restrict__ out_buf, int nv_cap, int px_count, int py_count, int total_patches) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int BX = kAsX * VEC;
    extern __shared__ __align__(16) unsigned char td_smem[];
    __shared__ VisitEntry s_visit[2][kAsMaxVisits];
    __shared__ int s_shift[2][kAsMaxVisits];
    const GeomParams& g = p.g;
    const int tid = threadIdx.x;
    const int set_bytes = nv_cap * kAsStage;

    // per-thread constants: copy slot (same in every stage) and output vector
    const int row = tid / kAsChunks, j = tid - row * kAsChunks;
    const int th1 = g.th - 1 - row, tw1 = (g.tw >> L2V) - 1 - j;
    const long long off = ((long long)row * g.tw + (long long)j * VEC) * (long long)sizeof(T);
    const uint32_t smem0 = smem_u32(td_smem);
    const int tx = tid % kAsX, ty = tid / kAsX;
    const bool vec_thread = tid < kAsVecThreads;

    struct Patch { int x_lo, y_lo, plane, nv; };
    auto decode = [&](int pi) {
        Patch q;
        const int px = pi % px_count, t = pi / px_count;
        const int py = t % py_count;
        q.plane = t / py_count;
        q.x_lo = px * BX; q.y_lo = py * kAsY;
        q.nv = g.dbg_no_tiles ? 0 : (int)p.prow_n[py] * (int)p.pcol_n[px];
        return q;
    };
    auto build_table = [&](const Patch& q, int set) {   // thread i prepares visit i
        if (tid < q.nv) {
            const int px = q.x_lo / BX, py = q.y_lo / kAsY;
            const int nc = p.pcol_n[px];
            const int ri = tid / nc, ci = tid - ri * nc;
            const int r = (int)p.prow_lo[py] + ri, c = (int)p.pcol_lo[px] + ci;
            const unsigned t = (unsigned)(r * g.cols + c);
            const unsigned b = fastdiv(t, p.bs_magic);
            const int u0 = q.x_lo - (int)g.xs[c];
            const int v0 = q.y_lo - (int)g.ys[r], k0 = u0 >> L2V;
            const long long elem = (long long)(t - b * (unsigned)p.tile_bs) * p.tile_stride + (long long)q.plane * g.th * g.tw +
                                   (long long)v0 * g.tw + (long long)k0 * VEC;
            VisitEntry e;
            e.origin = (long long)reinterpret_cast<uintptr_t>(p.batch_ptrs[b]) + elem * (long long)sizeof(T);
            e.v0 = v0;
            e.k0 = k0;
            s_visit[set][tid] = e;
            s_shift[set][tid] = u0 & (VEC - 1);
        }
    };
    auto issue = [&](const Patch& q, int set) {
        if (tid < kAsSlots) {
            uint32_t dst = smem0 + (uint32_t)(set * set_bytes) + (uint32_t)tid * 16u;
#pragma unroll 4
            for (int i = 0; i < q.nv; ++i, dst += kAsStage) {
                const VisitEntry e = s_visit[set][i];
                const int skip = (e.v0 + row) | (th1 - e.v0) | (e.k0 + j) | (tw1 - e.k0);
                cp_async16_zfill(dst, e.origin + off, skip);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float4 wv[VEC / 4], rv[VEC / 4], wn[VEC / 4], rn[VEC / 4];
    auto prefetch_weights = [&](const Patch& q) {
        const int x0 = q.x_lo + tx * VEC, y = q.y_lo + ty;
        const bool inside = vec_thread && x0 < g.W && y < g.H;
        const long long wo = (long long)y * g.W + x0;
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {
            wn[h] = inside ? __ldg(reinterpret_cast<const float4*>(weights + wo) + h) : make_float4(1.f, 1.f, 1.f, 1.f);
            if constexpr (FASTDIV) rn[h] = inside ? __ldg(reinterpret_cast<const float4*>(rcp_weights + wo) + h) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
    };

    int pi = blockIdx.x;
    if (pi >= total_patches) return;
    Patch nxt = decode(pi);
    build_table(nxt, 0);
    __syncthreads();
    issue(nxt, 0);
    prefetch_weights(nxt);
    int b = 0;
    while (true) {
        const Patch cur = nxt;
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) { wv[h] = wn[h]; if constexpr (FASTDIV) rv[h] = rn[h]; }
        const int next_pi = pi + gridDim.x;
        const bool has_next = next_pi < total_patches;
        __syncthreads();                       // everyone is done with set / table 1-b (consumed one iteration ago)
        if (has_next) { nxt = decode(next_pi); build_table(nxt, 1 - b); }
        __syncthreads();                       // table 1-b visible
        if (has_next) { issue(nxt, 1 - b); prefetch_weights(nxt); }
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();                       // set b has landed for every thread

        if (vec_thread) {
            const unsigned char* mine = td_smem + b * set_bytes + (ty * kAsChunks + tx) * 16;
            uint4 acc = make_uint4(0, 0, 0, 0);
            for (int i = 0; i < cur.nv; ++i, mine += kAsStage) {
                const int s = s_shift[b][i];   // uniform over the CTA
                if constexpr (VEC == 8) {
                    switch (s) {
                        case 0: consume_shifted<T, 0>(acc, mine); break;
                        case 1: consume_shifted<T, 1>(acc, mine); break;
                        case 2: consume_shifted<T, 2>(acc, mine); break;
                        case 3: consume_shifted<T, 3>(acc, mine); break;
                        case 4: consume_shifted<T, 4>(acc, mine); break;
                        case 5: consume_shifted<T, 5>(acc, mine); break;
                        case 6: consume_shifted<T, 6>(acc, mine); break;
                        default: consume_shifted<T, 7>(acc, mine); break;
                    }
                } else {
                    switch (s) {
                        case 0: consume_shifted<T, 0>(acc, mine); break;
                        case 1: consume_shifted<T, 1>(acc, mine); break;
                        case 2: consume_shifted<T, 2>(acc, mine); break;
                        default: consume_shifted<T, 3>(acc, mine); break;
                    }
                }
            }
            const int x0 = cur.x_lo + tx * VEC, y = cur.y_lo + ty;
            if (x0 < g.W && y < g.H) {
                const long long o = ((long long)cur.plane * g.H + y) * g.W + x0;
                float4* op = reinterpret_cast<float4*>(out_f32 + o);
#pragma unroll
                for (int h = 0; h < VEC / 4; ++h) {
                    const float4 w = wv[h];
                    const float a0 = Vec<T>::get(acc, 4 * h + 0), a1 = Vec<T>::get(acc, 4 * h + 1);
                    const float a2 = Vec<T>::get(acc, 4 * h + 2), a3 = Vec<T>::get(acc, 4 * h + 3);
                    float4 f;   // x_out = where(weights > 1, x_buffer / weights, x_buffer)  -- fp32, correctly rounded divide
                    if constexpr (FASTDIV) {
                        const float4 rc = rv[h];
                        f.x = w.x > 1.0f ? div_exact_small_int(a0, w.x, rc.x) : a0;
                        f.y = w.y > 1.0f ? div_exact_small_int(a1, w.y, rc.y) : a1;
                        f.z = w.z > 1.0f ? div_exact_small_int(a2, w.z, rc.z) : a2;
                        f.w = w.w > 1.0f ? div_exact_small_int(a3, w.w, rc.w) : a3;
                    } else {
                        f.x = w.x > 1.0f ? __fdiv_rn(a0, w.x) : a0;
                        f.y = w.y > 1.0f ? __fdiv_rn(a1, w.y) : a1;
                        f.z = w.z > 1.0f ? __fdiv_rn(a2, w.z) : a2;
                        f.w = w.w > 1.0f ? __fdiv_rn(a3, w.w) : a3;
                    }
                    op[h] = f;
                }
                if constexpr (WRITE_BUF) stg128(out_buf + o, acc);
            }
        }
        if (!has_next) break;
        pi = next_pi;
        b ^= 1;
    }
}

// exhaustive check of div_exact_small_int against the IEEE divide: every 16-bit pattern of T x w in [1, max_w]
template <typename T>
__global__ void check_fast_div_kernel(int max_w, unsigned long long* mismatches) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // 16-bit pattern
    if (idx >= 65536u) return;
    const float a = Elem<T>::bits_to_f32((uint16_t)idx);
    if (isnan(a) || isinf(a)) return;
    unsigned long long bad = 0;
    for (int wi = 1; wi <= max_w; ++wi) {
        const float w = (float)wi;
        const float rcp = __fdiv_rn(1.0f, w);
        const float q = div_exact_small_int(a, w, rcp);
        const float ref = __fdiv_rn(a, w);
        if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------------------
// Blend + normalise, MultiDiffusion, TMA path (TD_FLAG_TMA; kept as the measured alternative).
//
// A CTA owns an 8-row x 16-vector patch of one canvas plane.  For every tile that
// touches the patch (ascending tile index) one thread issues ONE cp.async.bulk.tensor
// box load: the copy engine clips the box against the tile and zero-fills the rest.
// All boxes of a CTA are in flight at once (one mbarrier each, no staging registers);
// the threads then walk the stages in tile order.  TMA wants a 16-byte aligned inner
// start coordinate, so the box is the aligned superset of the patch (BX + VEC columns)
// and the per-tile shift s = (x_lo - xs) mod VEC -- uniform over the CTA -- is applied
// on the shared-memory read: two aligned LDS.128 + funnel shift, then a packed add.
// Zero fill is exact: the accumulator starts at +0 and can never become -0.
// ---------------------------------------------------------------------------
constexpr int kTmaBXV = 16;   // vectors per patch row
constexpr int kTmaBY = 8;     // patch rows
constexpr int kTmaThreads = kTmaBXV * kTmaBY;
constexpr int kTmaMaxVisits = 96;

template <int NB>
struct TmaBlendParams {
    GeomParams g;
    int tile_bs;
    unsigned bs_magic;
    int nv_cap;
    CUtensorMap maps[NB];
};

template <typename T, bool WRITE_BUF, int NB>
__global__ void __launch_bounds__(kTmaThreads)
blend_md_tma_kernel(const __grid_constant__ TmaBlendParams<NB> p, const float* __restrict__ weights,
                    float* __restrict__ out_f32, T* __restrict__ out_buf) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    constexpr int BX = kTmaBXV * VEC;
    constexpr int PITCH = (BX + VEC) * (int)sizeof(T);   // bytes per staged row (aligned superset)
    constexpr int STAGE = PITCH * kTmaBY;                // multiple of 128 B
    extern __shared__ __align__(128) unsigned char td_smem[];
    __shared__ short s_ys[TD_MAX_GRID_DIM];
    __shared__ short s_xs[TD_MAX_GRID_DIM];
    __shared__ signed char s_shift[kTmaMaxVisits];
    const GeomParams& g = p.g;
    uint64_t* bars = reinterpret_cast<uint64_t*>(td_smem + (size_t)p.nv_cap * STAGE);
    load_origins(g, s_ys, s_xs);

    const int plane = blockIdx.z;
    const int x_lo = blockIdx.x * BX, y_lo = blockIdx.y * kTmaBY;
    const int x_hi = min(x_lo + BX, g.W) - 1, y_hi = min(y_lo + kTmaBY, g.H) - 1;
    const int r_lo = last_le(s_ys, g.rows, y_lo - g.th, g.inv_dy) + 1;
    const int r_hi = last_le(s_ys, g.rows, y_hi, g.inv_dy);
    const int c_lo = last_le(s_xs, g.cols, x_lo - g.tw, g.inv_dx) + 1;
    const int c_hi = last_le(s_xs, g.cols, x_hi, g.inv_dx);
    const int nc = c_hi - c_lo + 1;
    const int nv = g.dbg_no_tiles ? 0 : (r_hi - r_lo + 1) * nc;   // <= nv_cap by construction on the host

    // visit i is owned by thread i: init its barrier, publish its shift, launch its box
    for (int i = threadIdx.x; i < nv; i += kTmaThreads) {
        const int ri = i / nc, ci = i - ri * nc;
        const int r = r_lo + ri, c = c_lo + ci;
        const unsigned t = (unsigned)(r * g.cols + c);
        const unsigned b = fastdiv(t, p.bs_magic);
        const int pz = (int)(t - b * (unsigned)p.tile_bs) * (g.N * g.C) + plane;   // plane inside batch tensor b
        const int u0 = x_lo - (int)s_xs[c];          // tile-local column of the patch's first pixel
        const int k = u0 >> L2V;                     // floor: aligned chunk holding it
        s_shift[i] = (signed char)(u0 & (VEC - 1));
        mbar_init(&bars[i], 1);
        fence_proxy_async();
        mbar_arrive_expect_tx(&bars[i], STAGE);
        tma_load_3d(td_smem + (size_t)i * STAGE, &p.maps[b], k * VEC, y_lo - (int)s_ys[r], pz, &bars[i]);
    }
    __syncthreads();

    const int tx = threadIdx.x % kTmaBXV, ty = threadIdx.x / kTmaBXV;
    const unsigned char* mine = td_smem + (size_t)ty * PITCH + (size_t)tx * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < nv; ++i) {
        const int s = (int)s_shift[i];
        mbar_wait(&bars[i], 0);
        const uint4 A = lds128(mine + (size_t)i * STAGE);
        uint4 e = A;
        if (s != 0) {
            const uint4 B = lds128(mine + (size_t)i * STAGE + 16);
            e = Vec<T>::window(A, B, s);
        }
        acc.x = PackedAdd<T>::add(acc.x, e.x);
        acc.y = PackedAdd<T>::add(acc.y, e.y);
        acc.z = PackedAdd<T>::add(acc.z, e.z);
        acc.w = PackedAdd<T>::add(acc.w, e.w);
    }

    const int x0 = x_lo + tx * VEC, y = y_lo + ty;
    if (x0 >= g.W || y >= g.H) return;
    const long long o = ((long long)plane * g.H + y) * g.W + x0;
    const float4* wp = reinterpret_cast<const float4*>(weights + (long long)y * g.W + x0);
    float4* op = reinterpret_cast<float4*>(out_f32 + o);
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
        const float4 w = __ldg(wp + h);
        const float a0 = Vec<T>::get(acc, 4 * h + 0), a1 = Vec<T>::get(acc, 4 * h + 1);
        const float a2 = Vec<T>::get(acc, 4 * h + 2), a3 = Vec<T>::get(acc, 4 * h + 3);
        float4 f;   // x_out = where(weights > 1, x_buffer / weights, x_buffer)  -- fp32, IEEE divide
        f.x = w.x > 1.0f ? __fdiv_rn(a0, w.x) : a0;
        f.y = w.y > 1.0f ? __fdiv_rn(a1, w.y) : a1;
        f.z = w.z > 1.0f ? __fdiv_rn(a2, w.z) : a2;
        f.w = w.w > 1.0f ? __fdiv_rn(a3, w.w) : a3;
        op[h] = f;
    }
    if constexpr (WRITE_BUF) stg128(out_buf + o, acc);
}

// ---------------------------------------------------------------------------
// Scatter, TMA path: a CTA moves one whole (tile, n, c) plane.  One thread issues up to
// kScBoxes box loads (aligned superset of `rb` tile rows each, one mbarrier per box) from
// the canvas, all in flight at once; the threads then walk the boxes in order: re-align
// (two LDS.128 + funnel shift specialised on the block-uniform shift) and write the tile
// batch with aligned 128-bit stores, so the first stores overlap the later loads.
// ---------------------------------------------------------------------------
constexpr int kScBoxes = 8;
constexpr int kScThreads = 256;

struct TmaScatterParams {
    GeomParams g;
    CUtensorMap src;   // canvas  [N*C][H][W], box [1][rb][tw + VEC]
};

template <typename T, int S>
__device__ __forceinline__ void scatter_rows(const unsigned char* box, T* dst, int nvec, int twv, unsigned twv_magic, int pitch) {
    for (int i = threadIdx.x; i < nvec; i += kScThreads) {
        const unsigned vi = fastdiv((unsigned)i, twv_magic);
        const int uv = i - (int)vi * twv;
        const unsigned char* src = box + (size_t)vi * pitch + (size_t)uv * 16;
        const uint4 A = lds128(src);
        uint4 e = A;
        if constexpr (S != 0) e = Vec<T>::window(A, lds128(src + 16), S);
        stg128(dst + (long long)i * Vec<T>::kElems, e);
    }
}

template <typename T>
__global__ void __launch_bounds__(kScThreads)
scatter_tma_kernel(const __grid_constant__ TmaScatterParams p, T* __restrict__ tiles, int tile_begin, int rb, int nboxes, int box_bytes) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int L2V = Vec<T>::kLog2;
    extern __shared__ __align__(128) unsigned char td_smem[];
    __shared__ uint64_t bars[kScBoxes];
    const GeomParams& g = p.g;
    const unsigned tp = blockIdx.x;
    const unsigned tl = fastdiv(tp, g.nc_magic);
    const int plane = (int)(tp - tl * (unsigned)(g.N * g.C));
    const int t = tile_begin + (int)tl;
    const int r = t / g.cols, c = t - r * g.cols;
    const int xs = (int)g.xs[c], ys = (int)g.ys[r];
    const int pitch = (g.tw + VEC) * (int)sizeof(T);
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int b = 0; b < nboxes; ++b) mbar_init(&bars[b], 1);
        fence_proxy_async();
    }
    pdl_wait();   // x is read (TMA) and tiles are written only after the previous kernel in the stream has completed
    if (threadIdx.x == 0) {
        for (int b = 0; b < nboxes; ++b) {
            mbar_arrive_expect_tx(&bars[b], (uint32_t)(pitch * rb));
            tma_load_3d(td_smem + (size_t)b * box_bytes, &p.src, (xs >> L2V) * VEC, ys + b * rb, plane, &bars[b]);
        }
    }
    __syncthreads();
    const int s = xs & (VEC - 1);
    const int twv = g.tw >> L2V;
    T* dst = tiles + (long long)tp * g.th * g.tw;
    for (int b = 0; b < nboxes; ++b) {
        const int nvec = min(rb, g.th - b * rb) * twv;
        const unsigned char* box = td_smem + (size_t)b * box_bytes;
        T* d = dst + (long long)b * rb * g.tw;
        mbar_wait(&bars[b], 0);
        if constexpr (VEC == 8) {
            switch (s) {
                case 0: scatter_rows<T, 0>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 1: scatter_rows<T, 1>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 2: scatter_rows<T, 2>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 3: scatter_rows<T, 3>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 4: scatter_rows<T, 4>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 5: scatter_rows<T, 5>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 6: scatter_rows<T, 6>(box, d, nvec, twv, g.twv_magic, pitch); break;
                default: scatter_rows<T, 7>(box, d, nvec, twv, g.twv_magic, pitch); break;
            }
        } else {
            switch (s) {
                case 0: scatter_rows<T, 0>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 1: scatter_rows<T, 1>(box, d, nvec, twv, g.twv_magic, pitch); break;
                case 2: scatter_rows<T, 2>(box, d, nvec, twv, g.twv_magic, pitch); break;
                default: scatter_rows<T, 3>(box, d, nvec, twv, g.twv_magic, pitch); break;
            }
        }
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
// magic = ceil(2^32 / d) for fastdiv(); 0 encodes d == 1
unsigned magic_u16(unsigned d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + d - 1) / d); }

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
    const unsigned nc = (unsigned)(N * C);
    if (nc >= 65536u) { td_set_error("N*C = %u too large", nc); return TD_ERR_UNSUPPORTED; }
    o->nc_magic = magic_u16(nc);
    o->twv_magic = 0;  // set by the vector launchers (depends on the element size)
    o->dbg_no_tiles = 0;
    o->inv_dx = g->cols > 1 ? (float)(g->cols - 1) / (float)std::max(1, g->W - g->tile_w) : 0.0f;
    o->inv_dy = g->rows > 1 ? (float)(g->rows - 1) / (float)std::max(1, g->H - g->tile_h) : 0.0f;
    return TD_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Launch with the programmatic-stream-serialisation attribute (see pdl_wait): inside a stream or a captured graph
// the kernel may start while its predecessor drains.  tl_pdl is set per C-ABI call from TD_FLAG_NO_PDL.
thread_local bool tl_pdl = true;
thread_local cudaError_t tl_launch_err = cudaSuccess;

template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = tl_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) tl_launch_err = e;   // reported by check_launch
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = tl_launch_err;
    tl_launch_err = cudaSuccess;
    if (e != cudaSuccess) {
        td_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return TD_ERR_CUDA;
    }
    return TD_OK;
}

inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

template <typename T>
int launch_scatter(GeomParams gp, const void* x, void* tiles, int tile_begin, int n_tiles, bool vec, cudaStream_t st) {
    const long long elems = (long long)n_tiles * gp.N * gp.C * gp.th * gp.tw;
    if (elems == 0) return TD_OK;
    if (vec) {
        const int twv = gp.tw / Vec<T>::kElems;
        gp.twv_magic = magic_u16((unsigned)twv);
        const int rows_per_block = std::max(1, std::min(gp.th, (128 * 4 + twv - 1) / twv));  // ~4 vectors per thread
        if ((long long)rows_per_block * twv >= 65536) { td_set_error("tile row too wide for the vector scatter"); return TD_ERR_UNSUPPORTED; }
        dim3 grid((unsigned)(n_tiles * gp.N * gp.C), (unsigned)((gp.th + rows_per_block - 1) / rows_per_block));
        scatter_vec_kernel<T><<<grid, 128, 0, st>>>(gp, (const T*)x, (T*)tiles, tile_begin, rows_per_block);
    } else {
        scatter_generic_kernel<T><<<blocks_for(elems, 256), 256, 0, st>>>(gp, (const T*)x, (T*)tiles, tile_begin, elems);
    }
    return check_launch("td_scatter_tiles");
}

template <typename T, int MODE>
int launch_blend_vec(const BlendParams& bp, const float* weights, const float* tile_weights, const float* rescale,
                     float* out_f32, void* out_buf, cudaStream_t st) {
    const GeomParams& g = bp.g;
    const int wv = g.W / Vec<T>::kElems;
    const int rows_per_block = kBlendWarps * 4 * BlendShape<MODE>::kRPT;
    dim3 grid((unsigned)((wv + kWarpVecs - 1) / kWarpVecs), (unsigned)((g.H + rows_per_block - 1) / rows_per_block),
              (unsigned)(g.N * g.C));
    if (grid.y > 65535u || grid.z > 65535u) { td_set_error("canvas too large for the vector blend grid"); return TD_ERR_UNSUPPORTED; }
    if (MODE == MODE_MD && out_buf != nullptr)
        blend_grid_vec_kernel<T, MODE, true><<<grid, kBlendWarps * 32, 0, st>>>(bp, weights, tile_weights, rescale, out_f32, (T*)out_buf);
    else
        blend_grid_vec_kernel<T, MODE, false><<<grid, kBlendWarps * 32, 0, st>>>(bp, weights, tile_weights, rescale, out_f32, (T*)out_buf);
    return check_launch("td_blend (vec)");
}

// max number of tile rows (cols) touching any BY-row (BX-px) patch: bounds the TMA stage count
int max_union(const int32_t* org, int n, int extent, int size, int patch) {
    int best = 0;
    for (int lo = 0; lo < size; lo += patch) {
        const int hi = std::min(lo + patch, size) - 1;
        int cnt = 0;
        for (int i = 0; i < n; ++i)
            if (org[i] <= hi && org[i] + extent > lo) ++cnt;
        best = std::max(best, cnt);
    }
    return best;
}

// dynamic shared memory a kernel may request without opting in (48 KB minus its static allocation, with margin)
constexpr int kNoOptInSmem = 40 * 1024;

// The opt-in above 48 KB of dynamic shared memory is a per-device function attribute: remember the size configured
// for each device ordinal (a host may drive several GPUs from one process), set it under a lock.
constexpr int kMaxDevices = 64;
struct SmemOptIn {
    std::atomic<int> bytes[kMaxDevices];
    SmemOptIn() { for (auto& b : bytes) b.store(kNoOptInSmem, std::memory_order_relaxed); }
};

template <typename KernelT>
int ensure_dyn_smem(KernelT kernel, int bytes, SmemOptIn* configured) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) { td_set_error("cudaGetDevice failed"); return TD_ERR_CUDA; }
    std::atomic<int>* slot = dev < kMaxDevices ? &configured->bytes[dev] : nullptr;
    if (slot != nullptr && bytes <= slot->load(std::memory_order_acquire)) return TD_OK;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (slot != nullptr && bytes <= slot->load(std::memory_order_relaxed)) return TD_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { td_set_error("cudaFuncSetAttribute(%d B smem): %s", bytes, cudaGetErrorString(e)); return TD_ERR_CUDA; }
    if (slot != nullptr) slot->store(bytes, std::memory_order_release);
    return TD_OK;
}

template <typename T, bool WRITE_BUF, int NB>
int launch_blend_tma_nb(const td_grid* g, const BlendParams& bp, int tile_dtype, const float* weights, float* out_f32, void* out_buf,
                        int nv_cap, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kTmaBXV * VEC;
    constexpr int STAGE = (BX + VEC) * (int)sizeof(T) * kTmaBY;
    static TmaBlendParams<NB> tp;   // filled per launch under a lock (the params are copied at launch)
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    tp.g = bp.g; tp.tile_bs = bp.tile_bs; tp.bs_magic = bp.bs_magic; tp.nv_cap = nv_cap;
    const int NC = bp.g.N * bp.g.C;
    for (int b = 0; b < bp.num_batches; ++b) {
        const int nt = std::min(bp.tile_bs, g->num_tiles - b * bp.tile_bs);
        int rc = td_encode_tensor_map_3d(&tp.maps[b], bp.batch_ptrs[b], tile_dtype, (uint64_t)nt * NC, (uint64_t)bp.g.th,
                                         (uint64_t)bp.g.tw, kTmaBY, BX + VEC);
        if (rc != TD_OK) return rc;
    }
    const int smem = nv_cap * (STAGE + 8);
    static SmemOptIn configured;
    int rc = ensure_dyn_smem(blend_md_tma_kernel<T, WRITE_BUF, NB>, smem, &configured);
    if (rc != TD_OK) return rc;
    dim3 grid((unsigned)((bp.g.W + BX - 1) / BX), (unsigned)((bp.g.H + kTmaBY - 1) / kTmaBY), (unsigned)NC);
    blend_md_tma_kernel<T, WRITE_BUF, NB><<<grid, kTmaThreads, smem, st>>>(tp, weights, out_f32, (T*)out_buf);
    return check_launch("td_blend_multidiffusion (tma)");
}

// returns TD_OK if launched, 1 if the TMA path does not apply (caller falls back), <0 on error
template <typename T>
int try_launch_blend_tma(const td_grid* g, const BlendParams& bp, int tile_dtype, const float* weights, float* out_f32,
                         void* out_buf, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kTmaBXV * VEC;
    constexpr int STAGE = (BX + VEC) * (int)sizeof(T) * kTmaBY;
    if (bp.g.N * bp.g.C > 65535 || (bp.g.H + kTmaBY - 1) / kTmaBY > 65535) return 1;
    const int nv_cap = max_union(g->ys, g->rows, g->tile_h, g->H, kTmaBY) * max_union(g->xs, g->cols, g->tile_w, g->W, BX);
    if (nv_cap <= 0 || nv_cap > kTmaMaxVisits || nv_cap * (STAGE + 8) > 200 * 1024) return 1;
    const bool wb = out_buf != nullptr;
    if (bp.num_batches <= 32)
        return wb ? launch_blend_tma_nb<T, true, 32>(g, bp, tile_dtype, weights, out_f32, out_buf, nv_cap, st)
                  : launch_blend_tma_nb<T, false, 32>(g, bp, tile_dtype, weights, out_f32, out_buf, nv_cap, st);
    return wb ? launch_blend_tma_nb<T, true, TD_MAX_BATCH_PTRS>(g, bp, tile_dtype, weights, out_f32, out_buf, nv_cap, st)
              : launch_blend_tma_nb<T, false, TD_MAX_BATCH_PTRS>(g, bp, tile_dtype, weights, out_f32, out_buf, nv_cap, st);
}

template <typename T>
int try_launch_scatter_tma(GeomParams gp, const void* x, void* tiles, int dtype, int tile_begin, int n_tiles, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    const int es = (int)sizeof(T);
    if (gp.tw + VEC > 256 || n_tiles <= 0) return 1;
    const long long planes_out = (long long)n_tiles * gp.N * gp.C;
    if (planes_out > 0x7fffffffLL) return 1;
    const int pitch = (gp.tw + VEC) * es;
    // whole plane per CTA in <= kScBoxes boxes of rb rows (<= 256 rows per box, ~8 KB each when possible)
    int rb = std::max(1, std::min(256, 8192 / pitch));
    rb = std::max(rb, (gp.th + kScBoxes - 1) / kScBoxes);
    if (rb > 256) return 1;
    const int nboxes = (gp.th + rb - 1) / rb;
    const int box_bytes = (rb * pitch + 127) / 128 * 128;
    const int smem = nboxes * box_bytes;
    if (smem > 200 * 1024) return 1;
    const int twv = gp.tw / VEC;
    if ((long long)rb * twv >= 65536) return 1;
    gp.twv_magic = magic_u16((unsigned)twv);
    static TmaScatterParams tp;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    tp.g = gp;
    int rc = td_encode_tensor_map_3d(&tp.src, x, dtype, (uint64_t)gp.N * gp.C, (uint64_t)gp.H, (uint64_t)gp.W, (uint32_t)rb,
                                     (uint32_t)(gp.tw + VEC));
    if (rc != TD_OK) return rc;
    static SmemOptIn configured;
    rc = ensure_dyn_smem(scatter_tma_kernel<T>, smem, &configured);
    if (rc != TD_OK) return rc;
    launch_pdl(scatter_tma_kernel<T>, dim3((unsigned)planes_out), dim3(kScThreads), (size_t)smem, st, tp, (T*)tiles, tile_begin, rb, nboxes, box_bytes);
    return check_launch("td_scatter_tiles (tma)");
}

// tiles touching each patch along one axis: first index and count (both fit a byte: <= 256 tiles per axis)
void fill_patch_table(const int32_t* org, int n, int extent, int size, int patch, unsigned char* lo_out, unsigned char* n_out) {
    int idx = 0;
    for (int lo = 0; lo < size; lo += patch, ++idx) {
        const int hi = std::min(lo + patch, size) - 1;
        int first = -1, cnt = 0;
        for (int i = 0; i < n; ++i)
            if (org[i] <= hi && org[i] + extent > lo) { if (first < 0) first = i; ++cnt; }
        lo_out[idx] = (unsigned char)std::max(first, 0);
        n_out[idx] = (unsigned char)std::min(cnt, 255);
    }
}

int sm_count() {
    static int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        return v;
    }();
    return n;
}

template <typename T, bool WRITE_BUF, bool FASTDIV, int PPC>
int launch_blend_async_ppc(const BlendParams& bp, const float* weights, const float* rcp_weights, float* out_f32, void* out_buf,
                           int nv_cap, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kAsX * VEC;
    const int px = (bp.g.W + BX - 1) / BX, py = (bp.g.H + kAsY - 1) / kAsY, planes = bp.g.N * bp.g.C;
    const int smem = nv_cap * PPC * kAsStage;
    static SmemOptIn configured;
    int rc = ensure_dyn_smem(blend_md_async_kernel<T, WRITE_BUF, FASTDIV, PPC>, smem, &configured);
    if (rc != TD_OK) return rc;
    const unsigned gy = (unsigned)(bp.py_count > 0 ? bp.py_count : py), gz = (unsigned)(planes / PPC);
    dim3 grid((unsigned)px, bp.sched ? gz : gy, bp.sched ? gy : gz);
    launch_pdl(blend_md_async_kernel<T, WRITE_BUF, FASTDIV, PPC>, grid, dim3(kAsThreads), (size_t)smem, st, bp, weights, rcp_weights,
               out_f32, (T*)out_buf);
    return TD_OK;
}

// planes per CTA (tl_ppc_max is set per C-ABI call: TD_FLAG_ONE_PLANE forces 1)
#ifndef TD_AS_PPC
#define TD_AS_PPC 2   // tuning knob (-DTD_AS_PPC=): 1 / 2 / 4
#endif
thread_local int tl_ppc_max = TD_AS_PPC;

template <typename T, bool WRITE_BUF, bool FASTDIV>
int launch_blend_async_impl(const BlendParams& bp, const float* weights, const float* rcp_weights, float* out_f32, void* out_buf,
                            int nv_cap, bool pipelined, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kAsX * VEC;
    const int px = (bp.g.W + BX - 1) / BX, py = (bp.g.H + kAsY - 1) / kAsY, planes = bp.g.N * bp.g.C;
    if (pipelined) {
        const int smem = 2 * nv_cap * kAsStage;
        static SmemOptIn configured;
        int rc = ensure_dyn_smem(blend_md_pipe_kernel<T, WRITE_BUF, FASTDIV>, smem, &configured);
        if (rc != TD_OK) return rc;
        const long long total = (long long)px * py * planes;
        const int per_sm = std::max(1, std::min(8, (220 * 1024) / (smem + 3 * 1024)));   // CTAs that fit one SM's shared memory
        const int grid = (int)std::min<long long>(total, (long long)sm_count() * per_sm);
        blend_md_pipe_kernel<T, WRITE_BUF, FASTDIV><<<grid, kAsThreads, smem, st>>>(bp, weights, rcp_weights, out_f32, (T*)out_buf, nv_cap,
                                                                                   px, py, (int)total);
    } else {
        int rc;
        if (tl_ppc_max >= 4 && planes % 4 == 0 && 4 * nv_cap * kAsStage <= 64 * 1024)
            rc = launch_blend_async_ppc<T, WRITE_BUF, FASTDIV, 4>(bp, weights, rcp_weights, out_f32, out_buf, nv_cap, st);
        else if (tl_ppc_max >= 2 && planes % 2 == 0 && 2 * nv_cap * kAsStage <= 64 * 1024)
            rc = launch_blend_async_ppc<T, WRITE_BUF, FASTDIV, 2>(bp, weights, rcp_weights, out_f32, out_buf, nv_cap, st);
        else
            rc = launch_blend_async_ppc<T, WRITE_BUF, FASTDIV, 1>(bp, weights, rcp_weights, out_f32, out_buf, nv_cap, st);
        if (rc != TD_OK) return rc;
    }
    return check_launch("td_blend_multidiffusion (cp.async)");
}

template <typename T, bool WRITE_BUF>
int launch_blend_async(const td_grid* g, const BlendParams& bp_in, const float* weights, const float* rcp_weights, float* out_f32,
                       void* out_buf, int nv_cap, bool pipelined, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kAsX * VEC;
    BlendParams bp = bp_in;
    fill_patch_table(g->ys, g->rows, g->tile_h, g->H, kAsY, bp.prow_lo, bp.prow_n);
    fill_patch_table(g->xs, g->cols, g->tile_w, g->W, BX, bp.pcol_lo, bp.pcol_n);
    if (bp.wait_flags != nullptr && bp.own_r_hi > bp.own_r_lo && bp.py_count > 0 && !pipelined) {
        // row-strip shard: only the CTAs that read a neighbour's band wait; they sit at the end of the row range the halo
        // enters from and are scheduled last (patch row = slowest grid dimension, walked away from that end)
        int first_wait = -1, last_wait = -1;
        for (int by = bp.py0; by < bp.py0 + bp.py_count; ++by) {
            const int lo = bp.prow_lo[by], n = bp.prow_n[by];
            if (n > 0 && (lo < bp.own_r_lo || lo + n > bp.own_r_hi)) { if (first_wait < 0) first_wait = by; last_wait = by; }
        }
        const int mid2 = 2 * bp.py0 + bp.py_count - 1;      // 2 x centre of the range
        bp.sched = (first_wait >= 0 && first_wait + last_wait <= mid2) ? 2 : 1;
    }
    if (rcp_weights != nullptr && sizeof(T) == 2)
        return launch_blend_async_impl<T, WRITE_BUF, true>(bp, weights, rcp_weights, out_f32, out_buf, nv_cap, pipelined, st);
    return launch_blend_async_impl<T, WRITE_BUF, false>(bp, weights, nullptr, out_f32, out_buf, nv_cap, pipelined, st);
}

// Mixture of Diffusers on the cp.async kernel; 0 launched, 1 not applicable, <0 error
template <typename T>
int try_launch_blend_mod_async(const td_grid* g, const BlendParams& bp_in, const float* tile_weights, const float* rescale, void* out_buf,
                               cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kAsX * VEC;
    if (bp_in.g.N * bp_in.g.C > 65535 || (bp_in.g.H + kAsY - 1) / kAsY > 65535 || bp_in.tile_stride >= (1ll << 31)) return 1;
    if ((bp_in.g.H + kAsY - 1) / kAsY > TD_MAX_GRID_DIM || (bp_in.g.W + BX - 1) / BX > TD_MAX_GRID_DIM) return 1;
    if ((reinterpret_cast<uintptr_t>(tile_weights) & 15u) != 0 || (bp_in.g.tw & 7) != 0) return 1;
    const int nv_cap = max_union(g->ys, g->rows, g->tile_h, g->H, kAsY) * max_union(g->xs, g->cols, g->tile_w, g->W, BX);
    if (nv_cap <= 0 || nv_cap > kAsMaxVisits || nv_cap > kAsThreads || nv_cap * kAsStage > 200 * 1024) return 1;
    BlendParams bp = bp_in;
    fill_patch_table(g->ys, g->rows, g->tile_h, g->H, kAsY, bp.prow_lo, bp.prow_n);
    fill_patch_table(g->xs, g->cols, g->tile_w, g->W, BX, bp.pcol_lo, bp.pcol_n);
    const int smem = nv_cap * kAsStage;
    static SmemOptIn configured;
    int rc = ensure_dyn_smem(blend_mod_async_kernel<T>, smem, &configured);
    if (rc != TD_OK) return rc;
    dim3 grid((unsigned)((bp.g.W + BX - 1) / BX), (unsigned)((bp.g.H + kAsY - 1) / kAsY), (unsigned)(bp.g.N * bp.g.C));
    launch_pdl(blend_mod_async_kernel<T>, grid, dim3(kAsThreads), (size_t)smem, st, bp, tile_weights, rescale, (T*)out_buf);
    return check_launch("td_blend_mixture (cp.async)");
}

// returns TD_OK if launched, 1 if the path does not apply (caller falls back), <0 on error
template <typename T>
int try_launch_blend_async(const td_grid* g, const BlendParams& bp, const float* weights, const float* rcp_weights, float* out_f32,
                           void* out_buf, bool pipelined, cudaStream_t st) {
    constexpr int VEC = Vec<T>::kElems;
    constexpr int BX = kAsX * VEC;
    if (bp.g.N * bp.g.C > 65535 || (bp.g.H + kAsY - 1) / kAsY > 65535) return 1;
    if (bp.tile_stride >= (1ll << 31)) return 1;
    const int nv_cap = max_union(g->ys, g->rows, g->tile_h, g->H, kAsY) * max_union(g->xs, g->cols, g->tile_w, g->W, BX);
    if (nv_cap <= 0 || nv_cap > kAsMaxVisits || nv_cap > kAsThreads || nv_cap * kAsStage > 200 * 1024) return 1;
    if ((bp.g.H + kAsY - 1) / kAsY > TD_MAX_GRID_DIM || (bp.g.W + BX - 1) / BX > TD_MAX_GRID_DIM) return 1;   // patch tables
    if (pipelined && 2 * nv_cap * kAsStage > 200 * 1024) pipelined = false;
    return out_buf != nullptr ? launch_blend_async<T, true>(g, bp, weights, rcp_weights, out_f32, out_buf, nv_cap, pipelined, st)
                              : launch_blend_async<T, false>(g, bp, weights, rcp_weights, out_f32, out_buf, nv_cap, pipelined, st);
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
    bp->wait_flags = nullptr; bp->wait_world = 0; bp->wait_value = nullptr;
    bp->py0 = 0; bp->py_count = 0;
    bp->own_r_lo = 0; bp->own_r_hi = 0; bp->sched = 0;
    bp->bs_magic = magic_u16((unsigned)tile_bs);
    bp->num_batches = num_batches;
    bp->tile_stride = (long long)N * C * g->tile_h * g->tile_w;
    return TD_OK;
}

bool blend_vec_ok(const BlendParams& bp, int tile_dtype, int acc_dtype, std::initializer_list<const void*> ptrs) {
    if (tile_dtype != acc_dtype) return false;
    if (bp.tile_stride >= (1ll << 31)) return false;  // the vector kernel uses 32-bit offsets inside a tile
    const int vec = 16 / td_dtype_size(tile_dtype);
    if (bp.g.W % vec != 0 || bp.g.tw % vec != 0) return false;
    for (int b = 0; b < bp.num_batches; ++b)
        if (!aligned16(bp.batch_ptrs[b])) return false;
    for (const void* p : ptrs)
        if (p != nullptr && !aligned16(p)) return false;
    return true;
}

}  // namespace

__global__ void empty_kernel() {}

extern "C" int td_debug_launch_empty(int blocks, int threads, void* stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024) { td_set_error("td_debug_launch_empty: bad launch shape"); return TD_ERR_INVALID_ARG; }
    empty_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>();
    return check_launch("td_debug_launch_empty");
}

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
    tl_pdl = !(flags & TD_FLAG_NO_PDL);
    const bool vec_ok = !(flags & TD_FLAG_FORCE_GENERIC) && gp.W % vec == 0 && gp.tw % vec == 0 && aligned16(x) && aligned16(tiles);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec_ok && (flags & TD_FLAG_ROWS)) {   // opt-in: row-block form (td_rows.cu); falls through when not applicable
        const int rc = td_rows_try_launch_scatter(g, x, tiles, N, C, dtype, tile_begin, tile_end, (flags & TD_FLAG_NO_PDL) ? 0 : 1, stream);
        if (rc <= 0) return rc;
    }
    if (vec_ok && !(flags & TD_FLAG_NO_TMA)) {
        const int rc = es == 2 ? try_launch_scatter_tma<__half>(gp, x, tiles, dtype, tile_begin, tile_end - tile_begin, s)
                               : try_launch_scatter_tma<float>(gp, x, tiles, dtype, tile_begin, tile_end - tile_begin, s);
        if (rc <= 0) return rc;   // launched (0) or hard error (<0); 1 = not applicable -> vector kernel
    }
    // fp16 and bf16 are both moved as opaque 16-bit words
    if (es == 2) return launch_scatter<__half>(gp, x, tiles, tile_begin, tile_end - tile_begin, vec_ok, s);
    return launch_scatter<float>(gp, x, tiles, tile_begin, tile_end - tile_begin, vec_ok, s);
}

extern "C" int td_blend_multidiffusion(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N,
                                       int C, int tile_dtype, int acc_dtype, const float* weights, const float* rcp_weights,
                                       float* x_out, void* x_buffer, uint32_t flags, void* stream) {
    BlendParams bp;
    tl_pdl = !(flags & TD_FLAG_NO_PDL);
    tl_ppc_max = (flags & TD_FLAG_ONE_PLANE) ? 1 : TD_AS_PPC;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (weights == nullptr || x_out == nullptr) { td_set_error("td_blend_multidiffusion: null weights / x_out"); return TD_ERR_INVALID_ARG; }
    bp.g.dbg_no_tiles = (flags & TD_FLAG_DBG_NO_TILES) ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!(flags & TD_FLAG_FORCE_GENERIC) && blend_vec_ok(bp, tile_dtype, acc_dtype, {weights, rcp_weights, x_out, x_buffer})) {
        int rc = 1;
        if (flags & TD_FLAG_ROWS) {
            // opt-in: row-block form (td_rows.cu); falls through to the default kernels when not applicable
            rc = td_rows_try_launch_md(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, weights, rcp_weights, x_out, x_buffer,
                                       (flags & TD_FLAG_NO_PDL) ? 0 : 1, (flags & TD_FLAG_DBG_NO_TILES) ? 1 : 0, stream);
            if (rc <= 0) return rc;
        }
        if (flags & TD_FLAG_STRIP) {   // opt-in: strip form (td_strip.cu); falls through to the default when not applicable
            rc = td_strip_try_launch(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, weights, rcp_weights, x_out, x_buffer,
                                     (flags & TD_FLAG_NO_PDL) ? 0 : 1, (flags & TD_FLAG_ONE_PLANE) ? 1 : 2, stream);
            if (rc <= 0) return rc;
        }
        if (flags & TD_FLAG_TMA) {
            switch (tile_dtype) {
                case TD_F16: rc = try_launch_blend_tma<__half>(g, bp, tile_dtype, weights, x_out, x_buffer, s); break;
                case TD_BF16: rc = try_launch_blend_tma<__nv_bfloat16>(g, bp, tile_dtype, weights, x_out, x_buffer, s); break;
                default: rc = try_launch_blend_tma<float>(g, bp, tile_dtype, weights, x_out, x_buffer, s); break;
            }
        } else if (!(flags & TD_FLAG_NO_TMA)) {
            const bool pipe = (flags & TD_FLAG_PIPELINE) != 0;   // measured slower on B200; opt-in
            switch (tile_dtype) {
                case TD_F16: rc = try_launch_blend_async<__half>(g, bp, weights, rcp_weights, x_out, x_buffer, pipe, s); break;
                case TD_BF16: rc = try_launch_blend_async<__nv_bfloat16>(g, bp, weights, rcp_weights, x_out, x_buffer, pipe, s); break;
                default: rc = try_launch_blend_async<float>(g, bp, weights, nullptr, x_out, x_buffer, pipe, s); break;
            }
        }
        if (rc <= 0) return rc;   // launched or hard error; 1 = not applicable -> register-staged kernel
        switch (tile_dtype) {
            case TD_F16: return launch_blend_vec<__half, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
            case TD_BF16: return launch_blend_vec<__nv_bfloat16, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
            default: return launch_blend_vec<float, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
        }
    }
    return dispatch_generic<MODE_MD>(tile_dtype, acc_dtype, bp, weights, nullptr, nullptr, x_out, x_buffer, s);
}

extern "C" int td_blend_multidiffusion_rows(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N,
                                            int C, int tile_dtype, int acc_dtype, const float* weights, const float* rcp_weights,
                                            float* x_out, void* x_buffer, int row_begin, int row_end, const uint32_t* wait_flags,
                                            int wait_count, const uint32_t* wait_value, int own_band_begin, int own_band_end,
                                            uint32_t flags, void* stream) {
    BlendParams bp;
    tl_pdl = !(flags & TD_FLAG_NO_PDL);
    tl_ppc_max = (flags & TD_FLAG_ONE_PLANE) ? 1 : TD_AS_PPC;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (weights == nullptr || x_out == nullptr) { td_set_error("td_blend_multidiffusion_rows: null weights / x_out"); return TD_ERR_INVALID_ARG; }
    if (row_begin < 0 || row_end > g->H || row_begin > row_end || row_begin % 8 != 0 || (row_end % 8 != 0 && row_end != g->H)) {
        td_set_error("td_blend_multidiffusion_rows: rows [%d,%d) must lie in [0,%d) on multiples of 8", row_begin, row_end, g->H);
        return TD_ERR_INVALID_ARG;
    }
    if (wait_count < 0 || wait_count > TD_MAX_PEERS || (wait_count > 0 && (wait_flags == nullptr || wait_value == nullptr))) {
        td_set_error("td_blend_multidiffusion_rows: bad wait table");
        return TD_ERR_INVALID_ARG;
    }
    if (row_begin == row_end) return TD_OK;
    if (!blend_vec_ok(bp, tile_dtype, acc_dtype, {weights, rcp_weights, x_out, x_buffer})) {
        td_set_error("td_blend_multidiffusion_rows: needs the vector path (same tile / canvas dtype, W and tile_w multiples of the vector, 16-byte aligned buffers)");
        return TD_ERR_UNSUPPORTED;
    }
    bp.py0 = row_begin / 8;
    bp.py_count = (row_end - row_begin + 7) / 8;
    if (wait_count > 0) { bp.wait_flags = wait_flags; bp.wait_world = wait_count; bp.wait_value = wait_value; }
    if (own_band_begin < 0 || own_band_end > g->rows || own_band_begin > own_band_end) {
        td_set_error("td_blend_multidiffusion_rows: own bands [%d,%d) outside the %d tile rows", own_band_begin, own_band_end, g->rows);
        return TD_ERR_INVALID_ARG;
    }
    if (wait_count > 0 && own_band_end > own_band_begin) { bp.own_r_lo = own_band_begin; bp.own_r_hi = own_band_end; }   // -> launch_blend_async
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    switch (tile_dtype) {
        case TD_F16: rc = try_launch_blend_async<__half>(g, bp, weights, rcp_weights, x_out, x_buffer, false, s); break;
        case TD_BF16: rc = try_launch_blend_async<__nv_bfloat16>(g, bp, weights, rcp_weights, x_out, x_buffer, false, s); break;
        default: rc = try_launch_blend_async<float>(g, bp, weights, nullptr, x_out, x_buffer, false, s); break;
    }
    if (rc == 1) { td_set_error("td_blend_multidiffusion_rows: geometry outside the cp.async kernel's limits"); return TD_ERR_UNSUPPORTED; }
    return rc;
}

extern "C" int td_debug_check_fast_div(int dtype, int max_w, unsigned long long* mismatches_dev, void* stream) {
    if (mismatches_dev == nullptr || max_w < 1 || (dtype != TD_F16 && dtype != TD_BF16)) {
        td_set_error("td_debug_check_fast_div: bad arguments");
        return TD_ERR_INVALID_ARG;
    }
    if (dtype == TD_F16) check_fast_div_kernel<__half><<<256, 256, 0, (cudaStream_t)stream>>>(max_w, mismatches_dev);
    else check_fast_div_kernel<__nv_bfloat16><<<256, 256, 0, (cudaStream_t)stream>>>(max_w, mismatches_dev);
    return check_launch("td_debug_check_fast_div");
}

extern "C" int td_blend_multidiffusion_peer(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N,
                                            int C, int tile_dtype, int acc_dtype, const float* weights, float* x_out,
                                            void* x_buffer, const uint32_t* wait_flags, int world, const uint32_t* wait_value,
                                            uint32_t flags, void* stream) {
    BlendParams bp;
    tl_pdl = !(flags & TD_FLAG_NO_PDL);
    tl_ppc_max = (flags & TD_FLAG_ONE_PLANE) ? 1 : TD_AS_PPC;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (weights == nullptr || x_out == nullptr) { td_set_error("td_blend_multidiffusion_peer: null weights / x_out"); return TD_ERR_INVALID_ARG; }
    if (wait_flags == nullptr || wait_value == nullptr || world <= 0 || world > TD_MAX_PEERS) { td_set_error("td_blend_multidiffusion_peer: bad wait table"); return TD_ERR_INVALID_ARG; }
    if (!blend_vec_ok(bp, tile_dtype, acc_dtype, {weights, x_out, x_buffer})) {
        td_set_error("td_blend_multidiffusion_peer: needs the vector path (same tile / canvas dtype, W and tile_w multiples of the vector, 16-byte aligned buffers)");
        return TD_ERR_UNSUPPORTED;
    }
    bp.wait_flags = wait_flags; bp.wait_world = world; bp.wait_value = wait_value;
    cudaStream_t s = (cudaStream_t)stream;
    if (flags & TD_FLAG_PEER_ASYNC) {   // cp.async-staged kernel over peer pointers: measured slower over NVLink than direct loads (opt-in)
        int rc;
        switch (tile_dtype) {
            case TD_F16: rc = try_launch_blend_async<__half>(g, bp, weights, nullptr, x_out, x_buffer, false, s); break;
            case TD_BF16: rc = try_launch_blend_async<__nv_bfloat16>(g, bp, weights, nullptr, x_out, x_buffer, false, s); break;
            default: rc = try_launch_blend_async<float>(g, bp, weights, nullptr, x_out, x_buffer, false, s); break;
        }
        if (rc <= 0) return rc;
    }
    switch (tile_dtype) {
        case TD_F16: return launch_blend_vec<__half, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
        case TD_BF16: return launch_blend_vec<__nv_bfloat16, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
        default: return launch_blend_vec<float, MODE_MD>(bp, weights, nullptr, nullptr, x_out, x_buffer, s);
    }
}

extern "C" int td_blend_mixture(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs, int N, int C,
                                int tile_dtype, int acc_dtype, const float* tile_weights, const float* rescale,
                                void* x_buffer, uint32_t flags, void* stream) {
    BlendParams bp;
    tl_pdl = !(flags & TD_FLAG_NO_PDL);
    tl_ppc_max = (flags & TD_FLAG_ONE_PLANE) ? 1 : TD_AS_PPC;
    int st = fill_blend(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, acc_dtype, &bp);
    if (st != TD_OK) return st;
    if (tile_weights == nullptr || rescale == nullptr || x_buffer == nullptr) {
        td_set_error("td_blend_mixture: null tile_weights / rescale / x_buffer");
        return TD_ERR_INVALID_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!(flags & TD_FLAG_FORCE_GENERIC) && blend_vec_ok(bp, tile_dtype, acc_dtype, {rescale, x_buffer})) {
        if (flags & TD_FLAG_ROWS) {   // opt-in: row-block form (td_rows.cu)
            const int rc = td_rows_try_launch_mod(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, tile_weights, rescale, x_buffer,
                                                  (flags & TD_FLAG_NO_PDL) ? 0 : 1, stream);
            if (rc <= 0) return rc;
        }
        if (flags & TD_FLAG_STRIP) {   // opt-in: strip form (td_strip.cu); falls through to the default when not applicable
            const int rc = td_strip_try_launch_mod(g, batch_ptrs, num_batches, tile_bs, N, C, tile_dtype, tile_weights, rescale, x_buffer,
                                                   (flags & TD_FLAG_NO_PDL) ? 0 : 1, stream);
            if (rc <= 0) return rc;
        }
        if (!(flags & TD_FLAG_NO_TMA)) {
            int rc;
            switch (tile_dtype) {
                case TD_F16: rc = try_launch_blend_mod_async<__half>(g, bp, tile_weights, rescale, x_buffer, s); break;
                case TD_BF16: rc = try_launch_blend_mod_async<__nv_bfloat16>(g, bp, tile_weights, rescale, x_buffer, s); break;
                default: rc = try_launch_blend_mod_async<float>(g, bp, tile_weights, rescale, x_buffer, s); break;
            }
            if (rc <= 0) return rc;   // launched or hard error; 1 = not applicable -> register-staged kernel
        }
        switch (tile_dtype) {
            case TD_F16: return launch_blend_vec<__half, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
            case TD_BF16: return launch_blend_vec<__nv_bfloat16, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
            default: return launch_blend_vec<float, MODE_MOD>(bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
        }
    }
    return dispatch_generic<MODE_MOD>(tile_dtype, acc_dtype, bp, nullptr, tile_weights, rescale, nullptr, x_buffer, s);
}
