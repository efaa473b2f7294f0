// Host-side tile bookkeeping of the C-ABI (no GPU needed).
//
// The reference does this arithmetic in Python: `/` is float64 division,
// math.ceil works on that float64, int() truncates toward zero.  C `double`
// reproduces every intermediate bit-for-bit, so the integer results are exact.
//
//   td_split_bboxes     <- tile_utils/utils.py:160-177
//   td_splitable        <- tile_utils/utils.py:151-158
//   td_gaussian_weights <- tile_utils/utils.py:180-194
//   td_grid_init        <- tile_methods/abstractdiffusion.py:172-186
//   td_grid_weights     <- tile_utils/utils.py:167,175 (+ abstractdiffusion.py:182)
//   td_rescale_factor   <- tile_methods/mixtureofdiffusers.py:32
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "td_b200.h"
#include "td_internal.h"

static thread_local char g_err[512] = "";

void td_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* td_last_error(void) { return g_err; }
extern "C" int td_abi_version(void) { return TD_ABI_VERSION; }

namespace {

// cols = math.ceil((w - overlap) / (tile_w - overlap))          utils.py:161
inline int ceil_div_py(int num, int den) { return (int)std::ceil((double)num / (double)den); }

struct Split {
    int cols, rows;
    double dx, dy;
};

inline int split_counts(int w, int h, int tile_w, int tile_h, int overlap, Split* s) {
    if (w <= 0 || h <= 0 || tile_w <= 0 || tile_h <= 0 || overlap < 0) {
        td_set_error("split_bboxes: non-positive size (w=%d h=%d tile=%dx%d overlap=%d)", w, h, tile_w, tile_h, overlap);
        return TD_ERR_INVALID_ARG;
    }
    if (tile_w == overlap || tile_h == overlap) {  // Python: ZeroDivisionError at utils.py:161-162
        td_set_error("split_bboxes: overlap %d equals the tile size (%dx%d): division by zero", overlap, tile_w, tile_h);
        return TD_ERR_INVALID_ARG;
    }
    if (tile_w > w || tile_h > h) {  // not reachable through init_grid_bbox (tile is clamped first)
        td_set_error("split_bboxes: tile %dx%d larger than canvas %dx%d", tile_w, tile_h, w, h);
        return TD_ERR_INVALID_ARG;
    }
    // NB: overlap > tile is legal when the tile was clamped to the canvas
    // (abstractdiffusion.py:176-178): (w-ov)/(tw-ov) is then negative/negative.
    s->cols = ceil_div_py(w - overlap, tile_w - overlap);
    s->rows = ceil_div_py(h - overlap, tile_h - overlap);
    if (s->cols < 0) s->cols = 0;  // Python: range(negative) is empty -> no tiles
    if (s->rows < 0) s->rows = 0;
    s->dx = s->cols > 1 ? (double)(w - tile_w) / (double)(s->cols - 1) : 0.0;  // utils.py:163
    s->dy = s->rows > 1 ? (double)(h - tile_h) / (double)(s->rows - 1) : 0.0;  // utils.py:164
    return TD_OK;
}

// x = min(int(col * dx), w - tile_w)                              utils.py:171
inline int origin(int idx, double d, int limit) { return std::min((int)((double)idx * d), limit); }

}  // namespace

extern "C" int td_split_bboxes(int w, int h, int tile_w, int tile_h, int overlap, int32_t* out_xywh, int cap,
                               int* out_cols, int* out_rows) {
    Split s;
    int st = split_counts(w, h, tile_w, tile_h, overlap, &s);
    if (st != TD_OK) return st;
    if (out_cols) *out_cols = s.cols;
    if (out_rows) *out_rows = s.rows;
    const long long T = (long long)s.cols * s.rows;
    if (out_xywh == nullptr) return (int)T;  // count query
    if (T > cap) {
        td_set_error("split_bboxes: %lld tiles exceed capacity %d", T, cap);
        return TD_ERR_CAPACITY;
    }
    int32_t* o = out_xywh;
    for (int row = 0; row < s.rows; ++row) {
        const int y = origin(row, s.dy, h - tile_h);
        for (int col = 0; col < s.cols; ++col) {
            const int x = origin(col, s.dx, w - tile_w);
            *o++ = x; *o++ = y; *o++ = tile_w; *o++ = tile_h;
        }
    }
    return (int)T;
}

extern "C" int td_splitable(int w, int h, int tile_w, int tile_h, int overlap) {
    w /= 8; h /= 8;  // opt_f, utils.py:152
    const int min_tile = std::min(tile_w, tile_h);
    if (overlap >= min_tile) overlap = min_tile - 4;
    if (tile_w - overlap <= 0 || tile_h - overlap <= 0) {
        td_set_error("splitable: tile %dx%d too small for overlap %d", tile_w, tile_h, overlap);
        return TD_ERR_INVALID_ARG;
    }
    const int cols = ceil_div_py(w - overlap, tile_w - overlap);
    const int rows = ceil_div_py(h - overlap, tile_h - overlap);
    return (cols > 1 || rows > 1) ? 1 : 0;
}

extern "C" int td_gaussian_weights(int tile_w, int tile_h, float* out) {
    if (tile_w <= 0 || tile_h <= 0 || out == nullptr) {
        td_set_error("gaussian_weights: bad args");
        return TD_ERR_INVALID_ARG;
    }
    // f(x, mid) = exp(-(x-mid)*(x-mid) / (tile_w*tile_w) / (2*var)) / sqrt(2*pi*var), var = 0.01
    // evaluated left to right in float64 exactly as utils.py:189 writes it.
    const double var = 0.01;
    const double norm = std::sqrt(2.0 * M_PI * var);
    const double tw2 = (double)(tile_w * tile_w);
    auto f = [&](double x, double mid) { return std::exp(-(x - mid) * (x - mid) / tw2 / (2.0 * var)) / norm; };
    const double xmid = (double)(tile_w - 1) / 2.0;  // utils.py:190
    const double ymid = (double)tile_h / 2.0;        // utils.py:191 (asymmetric on purpose)
    for (int y = 0; y < tile_h; ++y) {
        const double yp = f((double)y, ymid);
        for (int x = 0; x < tile_w; ++x) out[(size_t)y * tile_w + x] = (float)(yp * f((double)x, xmid));  // np.outer -> fp32
    }
    return TD_OK;
}

extern "C" int td_feather_mask(int w, int h, double ratio, float* out) {
    if (w <= 0 || h <= 0 || out == nullptr) {
        td_set_error("feather_mask: bad args");
        return TD_ERR_INVALID_ARG;
    }
    for (size_t i = 0; i < (size_t)w * h; ++i) out[i] = 1.0f;
    // feather_radius = int(min(w//2, h//2) * ratio): float64 product, truncation (utils.py:200)
    const int radius = (int)((double)std::min(w / 2, h / 2) * ratio);
    for (int i = 0; i < h / 2; ++i)
        for (int j = 0; j < w / 2; ++j) {
            const int dist = std::min(i, j);
            if (dist >= radius) continue;
            // (dist / radius) ** 2 in float64, stored into a float32 array (utils.py:208)
            const float wt = (float)std::pow((double)dist / (double)radius, 2.0);
            out[(size_t)i * w + j] = wt;
            out[(size_t)i * w + (w - j - 1)] = wt;
            out[(size_t)(h - i - 1) * w + j] = wt;
            out[(size_t)(h - i - 1) * w + (w - j - 1)] = wt;
        }
    return TD_OK;
}

extern "C" int td_custom_bbox_rect(double x, double y, double w, double h, int canvas_w, int canvas_h, int32_t* out_xywh) {
    if (out_xywh == nullptr || canvas_w <= 0 || canvas_h <= 0) {
        td_set_error("custom_bbox_rect: bad args");
        return TD_ERR_INVALID_ARG;
    }
    if (x > 1.0 || y > 1.0 || w <= 0.0 || h <= 0.0) return 0;   // skipped by the reference (abstractdiffusion.py:207)
    // int(x * W), math.ceil(w * W) on float64 products; then clamps (abstractdiffusion.py:208-215)
    int xi = (int)(x * (double)canvas_w);
    int yi = (int)(y * (double)canvas_h);
    int wi = (int)std::ceil(w * (double)canvas_w);
    int hi = (int)std::ceil(h * (double)canvas_h);
    xi = std::max(0, xi);
    yi = std::max(0, yi);
    wi = std::min(canvas_w - xi, wi);
    hi = std::min(canvas_h - yi, hi);
    out_xywh[0] = xi; out_xywh[1] = yi; out_xywh[2] = wi; out_xywh[3] = hi;
    return 1;
}

extern "C" int td_grid_init(td_grid* g, int w, int h, int tile_w, int tile_h, int overlap, int tile_bs) {
    if (g == nullptr || tile_bs <= 0) {
        td_set_error("grid_init: bad args");
        return TD_ERR_INVALID_ARG;
    }
    if (tile_w <= 0 || tile_h <= 0 || w <= 0 || h <= 0) {
        td_set_error("grid_init: non-positive size");
        return TD_ERR_INVALID_ARG;
    }
    std::memset(g, 0, sizeof(*g));
    g->W = w; g->H = h;
    g->tile_w = std::min(tile_w, w);                                     // abstractdiffusion.py:176
    g->tile_h = std::min(tile_h, h);                                     // :177
    g->overlap = std::max(0, std::min(overlap, std::min(tile_w, tile_h) - 4));  // :178 (unclamped tile args)
    Split s;
    int st = split_counts(w, h, g->tile_w, g->tile_h, g->overlap, &s);
    if (st != TD_OK) return st;
    if (s.rows > TD_MAX_GRID_DIM || s.cols > TD_MAX_GRID_DIM) {
        td_set_error("grid_init: %dx%d tile grid exceeds TD_MAX_GRID_DIM=%d", s.rows, s.cols, TD_MAX_GRID_DIM);
        return TD_ERR_UNSUPPORTED;
    }
    g->rows = s.rows; g->cols = s.cols;
    for (int r = 0; r < s.rows; ++r) g->ys[r] = origin(r, s.dy, h - g->tile_h);
    for (int c = 0; c < s.cols; ++c) g->xs[c] = origin(c, s.dx, w - g->tile_w);
    g->num_tiles = s.rows * s.cols;                                      // :183
    if (g->num_tiles == 0) {  // Python: ZeroDivisionError at :185
        td_set_error("grid_init: the split produced no tiles (w=%d h=%d tile=%dx%d overlap=%d)", w, h, g->tile_w, g->tile_h, g->overlap);
        return TD_ERR_INVALID_ARG;
    }
    g->num_batches = ceil_div_py(g->num_tiles, tile_bs);                 // :184
    g->tile_bs = ceil_div_py(g->num_tiles, g->num_batches);              // :185
    return g->num_tiles;
}

extern "C" int td_grid_weights(const td_grid* g, const float* tile_weights, float* out) {
    if (g == nullptr || out == nullptr) {
        td_set_error("grid_weights: null");
        return TD_ERR_INVALID_ARG;
    }
    const int W = g->W, H = g->H, tw = g->tile_w, th = g->tile_h;
    std::fill(out, out + (size_t)W * H, 0.0f);
    for (int r = 0; r < g->rows; ++r)
        for (int c = 0; c < g->cols; ++c) {  // list order: row outer, col inner
            const int x0 = g->xs[c], y0 = g->ys[r];
            for (int v = 0; v < th; ++v) {
                float* o = out + (size_t)(y0 + v) * W + x0;
                if (tile_weights) {
                    const float* tw_row = tile_weights + (size_t)v * tw;
                    for (int u = 0; u < tw; ++u) o[u] = o[u] + tw_row[u];
                } else {
                    for (int u = 0; u < tw; ++u) o[u] = o[u] + 1.0f;
                }
            }
        }
    return TD_OK;
}

extern "C" int td_rescale_factor(const float* weights, float* out, int64_t n) {
    if (weights == nullptr || out == nullptr || n < 0) {
        td_set_error("rescale_factor: bad args");
        return TD_ERR_INVALID_ARG;
    }
    for (int64_t i = 0; i < n; ++i) out[i] = 1.0f / weights[i];  // IEEE; inf where uncovered, like torch
    return TD_OK;
}
