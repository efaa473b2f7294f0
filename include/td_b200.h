/*
 * td_b200.h -- C-ABI of the B200-native tiled-diffusion / tiled-VAE hot path.
 *
 * The reference (pkuliyi2015/multidiffusion-upscaler-for-automatic1111 @ 22798f6)
 * has NO FFI boundary: its hot path is eager PyTorch inside Python classes
 * (SURVEY.md section 8(b)).  This header is the boundary our Python host classes
 * (same names / arguments as the reference's tile-method + VAEHook surface) bind
 * with ctypes; INTEGRATION.md shows the binding a reference maintainer would add.
 * Each entry point cites the reference code it replaces (file:line relative to
 * the reference root).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, a cudaStream_t passed as void*;
 *   - DEVICE pointers are `tensor.data_ptr()`; the caller (PyTorch) owns all memory;
 *   - no allocation, no host<->device sync, no exceptions: every function returns
 *     TD_OK (0) or a negative td_status, and td_last_error() gives the message;
 *   - all work is enqueued on the given stream (reference: "current stream").
 */
#ifndef TD_B200_H
#define TD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD_ABI_VERSION 1

typedef enum td_status {
    TD_OK = 0,
    TD_ERR_INVALID_ARG = -1, /* null pointer, negative size, out-of-range index */
    TD_ERR_UNSUPPORTED = -2, /* legal in the reference but not on this path (e.g. >256 grid rows) */
    TD_ERR_CUDA = -3,        /* launch / driver error; see td_last_error() */
    TD_ERR_CAPACITY = -4     /* caller-provided output array too small */
} td_status;

typedef enum td_dtype { TD_F16 = 0, TD_BF16 = 1, TD_F32 = 2 } td_dtype;

/* blend flags */
#define TD_FLAG_FORCE_GENERIC 1u /* use the scalar any-alignment kernel (test / fallback path) */
#define TD_FLAG_NO_TMA 2u        /* skip the TMA kernels, use the register-staged vector kernels  */
#define TD_FLAG_DBG_NO_TILES 0x100u /* measurement aid: skip all tile visits (launch + epilogue floor) */

#define TD_MAX_GRID_DIM 256   /* max tile rows / cols of a grid plan            */
#define TD_MAX_BATCH_PTRS 128 /* max UNet output batch tensors per blend launch */

const char* td_last_error(void);
int td_abi_version(void);
/* measurement aid: an empty kernel launch (launch-latency floor of the bench harness) */
int td_debug_launch_empty(int blocks, int threads, void* stream);

/* ------------------------------------------------------------------------- *
 *  Host bookkeeping (integer-exact; Python float64 semantics reproduced with
 *  C double).  No GPU needed.
 * ------------------------------------------------------------------------- */

/* split_bboxes -- tile_utils/utils.py:160-177.
 * Writes up to `cap` tiles as (x, y, w, h) int32 quadruples in row-major order
 * (row outer, col inner).  Returns the tile count T (>= 1) or a td_status.
 * out_cols / out_rows may be NULL. */
int td_split_bboxes(int w, int h, int tile_w, int tile_h, int overlap,
                    int32_t* out_xywh, int cap, int* out_cols, int* out_rows);

/* splitable -- tile_utils/utils.py:151-158 (w, h in IMAGE pixels). Returns 0/1. */
int td_splitable(int w, int h, int tile_w, int tile_h, int overlap);

/* gaussian_weights -- tile_utils/utils.py:180-194.  out: fp32 [tile_h*tile_w]. */
int td_gaussian_weights(int tile_w, int tile_h, float* out);

/* Grid plan = what init_grid_bbox leaves on the delegate
 * (tile_methods/abstractdiffusion.py:172-186): clamped tile size and overlap,
 * separable tile origins, tile count and re-balanced tile batch size. */
typedef struct td_grid {
    int32_t H, W;           /* latent canvas, abstractdiffusion.py:25-26           */
    int32_t tile_h, tile_w; /* min(tile, canvas), :176-177                         */
    int32_t overlap;        /* clamped with the UNclamped tile args, :178          */
    int32_t rows, cols;     /* utils.py:161-162                                    */
    int32_t num_tiles;      /* rows*cols                                           */
    int32_t num_batches;    /* ceil(T / tile_bs), :184                             */
    int32_t tile_bs;        /* ceil(T / num_batches), :185                         */
    int32_t ys[TD_MAX_GRID_DIM]; /* tile-row origins, utils.py:169                 */
    int32_t xs[TD_MAX_GRID_DIM]; /* tile-col origins, utils.py:171                 */
} td_grid;

/* init_grid_bbox -- abstractdiffusion.py:172-186.  w, h are LATENT sizes. */
int td_grid_init(td_grid* g, int w, int h, int tile_w, int tile_h, int overlap, int tile_bs);

/* Weight canvas of split_bboxes -- utils.py:167,175: fp32 [H*W] zeros, then
 * `+= init_weight` per tile in list order.  tile_weights NULL => 1.0
 * (abstractdiffusion.py:188-190), else fp32 [tile_h*tile_w]
 * (mixtureofdiffusers.py:38-43).  HOST arrays. */
int td_grid_weights(const td_grid* g, const float* tile_weights, float* out_weights);

/* rescale_factor = 1 / weights -- mixtureofdiffusers.py:32 (fp32 IEEE divide). */
int td_rescale_factor(const float* weights, float* out, int64_t n);

/* ------------------------------------------------------------------------- *
 *  Device kernels: per-sampler-step hot path.
 * ------------------------------------------------------------------------- */

/* Scatter -- multidiffusion.py:155, mixtureofdiffusers.py:88,104:
 *   tiles[(t - tile_begin)*N + n, c, v, u] = x[n, c, ys[t/cols] + v, xs[t%cols] + u]
 * for t in [tile_begin, tile_end): ONE launch for the whole step (the reference
 * issues one torch.cat per tile batch).  x: [N,C,H,W] contiguous; tiles:
 * [(tile_end-tile_begin)*N, C, tile_h, tile_w] contiguous; same dtype. */
int td_scatter_tiles(const td_grid* g, const void* x, void* tiles, int N, int C, int dtype,
                     int tile_begin, int tile_end, uint32_t flags, void* stream);

/* Blend + normalise, MultiDiffusion -- multidiffusion.py:166-167 + :208 fused,
 * gather form, one launch per step:
 *   acc = 0 (acc_dtype);  for covering tiles t ascending: acc = round_acc(float(acc) + float(tile_t))
 *   x_out = weights > 1 ? float(acc) / weights : float(acc)          (fp32, IEEE divide)
 * which is bit-identical to the reference's sequential `x_buffer[slicer] +=`.
 * batch_ptrs: HOST array of `num_batches` DEVICE pointers, batch b holding tiles
 * [b*tile_bs, min((b+1)*tile_bs, T)) as [.*N, C, tile_h, tile_w] contiguous
 * (the UNet's output tensors; peer-GPU pointers are allowed).
 * weights: fp32 [H*W] device.  x_out: fp32 [N,C,H,W].  x_buffer (acc_dtype,
 * [N,C,H,W]) is optional (NULL = do not materialise abstractdiffusion.py:24). */
int td_blend_multidiffusion(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs,
                            int N, int C, int tile_dtype, int acc_dtype, const float* weights,
                            float* x_out, void* x_buffer, uint32_t flags, void* stream);

/* Blend, Mixture of Diffusers -- mixtureofdiffusers.py:122-126, returns x_buffer (:169):
 *   w   = tile_weights[v,u] * rescale[y,x]                       (fp32 product)
 *   acc = round_acc(float(acc) + float(tile_t) * w)              (separate mul / add roundings)
 * tile_weights: fp32 [tile_h*tile_w]; rescale: fp32 [H*W]; x_buffer: acc_dtype [N,C,H,W]. */
int td_blend_mixture(const td_grid* g, const void* const* batch_ptrs, int num_batches, int tile_bs,
                     int N, int C, int tile_dtype, int acc_dtype, const float* tile_weights,
                     const float* rescale, void* x_buffer, uint32_t flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TD_B200_H */
