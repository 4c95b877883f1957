/*
 * swfr.h - C ABI of libswfr_b200.so, the B200-native (sm_100a) replacement for the shape -> pixels path
 * of open-flash/swf-renderer.
 *
 * The boundary mirrors the reference Rust crate's public API (the crate keeps its traits; a Rust shim binds
 * these entry points).  Every entry point cites the reference interface it replaces,
 * paths relative to the reference repository:
 *
 *   swfr_create / swfr_destroy        HeadlessGfxRenderer::new / Drop         rs/src/headless_renderer.rs:60-64, 871-904
 *                                     createRenderer / destroyRenderer        rs/src/wasm.rs:60-99
 *   swfr_register_shape               ClientAssetStore::register_shape        rs/src/asset.rs:9-12
 *                                     HeadlessGfxRenderer::define_shape       rs/src/headless_renderer.rs:229-231
 *                                     decodeSwfShape (compile, cached)        ts/src/lib/shape/decode-swf-shape.ts:22-39
 *   swfr_register_morph_shape         ClientAssetStore::register_morph_shape  rs/src/asset.rs:9-12
 *                                     decodeSwfMorphShape                     ts/src/lib/shape/decode-swf-morph-shape.ts:21-41
 *   swfr_register_bitmap[_xswfbmp]    Renderer.addBitmap(tag)                 ts/src/lib/renderer.ts:4-8
 *                                     NodeCanvasBitmapService.addBitmap       ts/src/lib/renderers/node-canvas-bitmap-service.ts:14-37
 *                                     decodeXSwfBmpSync                       ts/src/lib/decode-x-swf-bmp.ts:9-41
 *   swfr_render                       SwfRenderer::render(stage)              rs/src/swf_renderer.rs:3-5
 *                                     Renderer::set_stage + get_image         rs/src/renderer.rs:81-87, headless_renderer.rs:233-244
 *                                     CanvasRenderer.render(stage)            ts/src/lib/renderers/canvas-renderer.ts:61-78
 *   swfr_render_batch                 the same, N stages per launch (frame batches / morph-ratio sweeps)
 *   swfr_read_image                   HeadlessGfxRenderer::download_image     rs/src/headless_renderer.rs:725-868
 *                                     Image{meta{width,height,stride},data}   rs/src/renderer.rs:89-103
 *                                     canvas.toBuffer("image/png") un-premultiply  ts/src/test/node-canvas-renderer.spec.ts:134-147
 *   swfr_debug_*                      parity taps (no reference counterpart)
 *
 * Conventions: every function returns a swfr_status (0 = ok, negative = error; the reference's
 * &'static str / panic / throw sites map to these codes).  Inputs are borrowed for the duration of the call
 * only (as register_shape(&tag) borrows); outputs are caller-allocated.  A handle owns one CUDA device and one
 * stream and is NOT thread-safe (the reference takes &mut self); different handles may be driven from different
 * threads or processes (that is how frames are sharded over 8 GPUs).  There is no CPU fallback: without a CUDA
 * device swfr_create fails with SWFR_ERR_CUDA.
 */
#ifndef SWFR_H
#define SWFR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWFR_ABI_VERSION 3

typedef enum swfr_status {
  SWFR_OK = 0,
  SWFR_ERR_INVALID_HANDLE = -1,
  SWFR_ERR_INVALID_ID = -2,        /* unknown shape / morph shape / bitmap id (reference: panic, BitmapNotFound) */
  SWFR_ERR_INVALID_FILL_ID = -3,   /* "Invalid fill ID" (decode-swf-shape.ts:418-420) */
  SWFR_ERR_UNSUPPORTED_STYLE = -4, /* NotImplementedFillStyle / NotImplementedLineStyle / "Unknown fill type" */
  SWFR_ERR_OOM = -5,
  SWFR_ERR_CUDA = -6,
  SWFR_ERR_INVALID_ARGUMENT = -7,
  SWFR_ERR_MALFORMED = -8          /* e.g. morph move_to without morph_move_to, bad x-swf-bmp stream */
} swfr_status;

typedef struct swfr_renderer swfr_renderer; /* opaque */
typedef struct swfr_batch swfr_batch;       /* opaque: a set of stages resident in HBM */

/* ---- swf-tree 0.8.0 POD mirrors (a Rust shim converts #[repr(C)] field by field) ------------------- */

typedef struct swfr_rgba8 { uint8_t r, g, b, a; } swfr_rgba8; /* swf_tree::StraightSRgba8 */

/* swf_tree::Matrix: Sfixed16P16 epsilons + twips translation */
typedef struct swfr_swf_matrix {
  int32_t scale_x, scale_y, rotate_skew0, rotate_skew1;
  int32_t translate_x, translate_y;
} swfr_swf_matrix;

typedef enum swfr_fill_type {
  SWFR_FILL_SOLID = 0,
  SWFR_FILL_LINEAR_GRADIENT = 1,
  SWFR_FILL_RADIAL_GRADIENT = 2,
  SWFR_FILL_FOCAL_GRADIENT = 3,
  SWFR_FILL_BITMAP = 4
} swfr_fill_type;

typedef enum swfr_spread { SWFR_SPREAD_PAD = 0, SWFR_SPREAD_REFLECT = 1, SWFR_SPREAD_REPEAT = 2 } swfr_spread;
typedef enum swfr_color_space { SWFR_COLOR_SRGB = 0, SWFR_COLOR_LINEAR_RGB = 1 } swfr_color_space;

typedef struct swfr_color_stop {
  uint8_t ratio;
  swfr_rgba8 color;
  swfr_rgba8 morph_color; /* morph shapes only */
} swfr_color_stop;

typedef struct swfr_gradient {
  uint8_t spread;      /* swfr_spread */
  uint8_t color_space; /* swfr_color_space */
  uint16_t n_colors;
  const swfr_color_stop *colors;
} swfr_gradient;

typedef struct swfr_fill_style {
  uint32_t type; /* swfr_fill_type */
  swfr_rgba8 color;
  swfr_rgba8 morph_color; /* morph shapes only */
  swfr_swf_matrix matrix;
  swfr_gradient gradient;
  int16_t focal_point; /* Sfixed8P8 epsilons */
  uint16_t bitmap_id;
  uint8_t repeating;
  uint8_t smoothed; /* carried, ignored like the reference (canvas-renderer.ts:295-316) */
} swfr_fill_style;

typedef struct swfr_line_style {
  uint16_t width;       /* twips */
  uint16_t morph_width; /* morph shapes only */
  swfr_fill_style fill;
  /* caps / joins / scaling flags of LineStyle2 are not consulted by the reference renderer */
} swfr_line_style;

typedef struct swfr_styles {
  uint32_t n_fill;
  const swfr_fill_style *fill;
  uint32_t n_line;
  const swfr_line_style *line;
} swfr_styles;

typedef enum swfr_record_type { SWFR_RECORD_EDGE = 0, SWFR_RECORD_STYLE_CHANGE = 1 } swfr_record_type;

/* swf_tree::ShapeRecord / MorphShapeRecord as a tagged POD */
typedef struct swfr_shape_record {
  uint32_t type; /* swfr_record_type */
  /* edge */
  int32_t delta_x, delta_y;
  int32_t control_delta_x, control_delta_y;
  int32_t morph_delta_x, morph_delta_y;
  int32_t morph_control_delta_x, morph_control_delta_y;
  uint8_t has_control_delta, has_morph_control_delta;
  /* style change */
  uint8_t has_move_to, has_morph_move_to, has_left_fill, has_right_fill, has_line_style, has_new_styles;
  int32_t move_to_x, move_to_y;
  int32_t morph_move_to_x, morph_move_to_y;
  uint32_t left_fill, right_fill, line_style;
  const swfr_styles *new_styles;
} swfr_shape_record;

/* swf_tree::tags::DefineShape / DefineMorphShape (fields the renderer reads) */
typedef struct swfr_define_shape {
  uint16_t id;
  int32_t bounds[4];       /* x_min, x_max, y_min, y_max (twips) */
  int32_t morph_bounds[4]; /* morph shapes only */
  swfr_styles initial_styles;
  uint32_t n_records;
  const swfr_shape_record *records;
} swfr_define_shape;

/* ---- stage (rs/src/stage.rs:4-59) ------------------------------------------------------------------- */

typedef enum swfr_primitive_kind { SWFR_PRIM_SHAPE = 0, SWFR_PRIM_MORPH_SHAPE = 1 } swfr_primitive_kind;

/* swf_tree::ColorTransformWithAlpha: per channel c' = clamp(((c * mult) >> 8) + add, 0, 255) on the straight (not
 * premultiplied) 8-bit colour; mult is Sfixed8P8 (epsilons: 256 = 1.0), add is a signed 16-bit integer.  The reference
 * renderer has no colour-transform input at all (its display list carries a matrix and a ratio); this is the SWF
 * semantics, defined here by the oracle ("parity unpinned"):
 *   solid fills      the fill colour is transformed before it is premultiplied;
 *   gradients, bitmaps   the evaluated premultiplied pixel is un-premultiplied ((c * 255 + a / 2) / a), transformed and
 *                    premultiplied again ((c' * a' + 127) / 255) - integer arithmetic throughout. */
typedef struct swfr_color_transform {
  int16_t red_mult, green_mult, blue_mult, alpha_mult;
  int16_t red_add, green_add, blue_add, alpha_add;
} swfr_color_transform;

/* DisplayPrimitive::{Shape(StoredShape), MorphShape(StoredMorphShape)} */
typedef struct swfr_display_primitive {
  uint32_t kind;   /* swfr_primitive_kind */
  uint32_t id;     /* ShapeId / MorphShapeId returned by swfr_register_* */
  float matrix[6]; /* Matrix2D: [scale_x, scale_y, rotate_skew0, rotate_skew1, translate_x, translate_y];
                      x' = m0*x + m3*y + m4, y' = m2*x + m1*y + m5 (twips) */
  uint16_t ratio;  /* MorphRatio (rs/src/stage.rs:28-34): 0 = start, 65535 = end; the lerp factor is ratio / 65535 */
  uint16_t flags;  /* SWFR_PRIM_RATIO_F32: ratio_f replaces ratio */
  float ratio_f;   /* the TypeScript renderer's MorphShape.ratio, a number in 0..1 (ts/src/lib/display/morph-shape.ts:5-10,
                      canvas-renderer.ts:190-205); lets a caller ask for exactly 0.5 */
  swfr_color_transform color_transform; /* read when flags has SWFR_PRIM_COLOR_TRANSFORM */
} swfr_display_primitive;
#define SWFR_PRIM_RATIO_F32 1u
#define SWFR_PRIM_COLOR_TRANSFORM 2u

typedef struct swfr_stage {
  swfr_rgba8 background_color; /* carried; the reference TS renderer ignores it (canvas-renderer.ts:70-72) */
  uint32_t n_primitives;
  const swfr_display_primitive *display_root;
} swfr_stage;

/* ---- display tree (ts/src/lib/display/) ---------------------------------------------------------- */

/* DisplayObjectType (ts/src/lib/display/display-object-type.ts:1-5), same order */
typedef enum swfr_display_object_type {
  SWFR_DISPLAY_CONTAINER = 0,
  SWFR_DISPLAY_MORPH_SHAPE = 1,
  SWFR_DISPLAY_SHAPE = 2
} swfr_display_object_type;

/* DisplayObject = DisplayObjectContainer | MorphShape | Shape.  The reference embeds the definition tag and
 * compiles it on first use (canvas-renderer.ts:96-112); here definitions are registered first and referenced by id. */
typedef struct swfr_display_object {
  uint32_t type;           /* swfr_display_object_type */
  uint32_t id;             /* ShapeId / MorphShapeId (shapes and morph shapes) */
  uint8_t has_matrix;      /* `matrix?: Matrix` */
  swfr_swf_matrix matrix;  /* swf-tree Matrix: Sfixed16P16 epsilons + twips */
  float ratio;             /* MorphShape.ratio, a number in 0..1 */
  uint8_t has_color_transform; /* applies to everything below a container (concatenated down the tree: the child's
                                  transform first, mult = (Pm * Cm) >> 8, add = ((Pm * Ca) >> 8) + Pa, clamped to int16) */
  swfr_color_transform color_transform;
  uint32_t n_children;     /* containers */
  const struct swfr_display_object *children;
} swfr_display_object;

/* Stage (ts/src/lib/display/stage.ts:7-18): size in pixels, root children */
typedef struct swfr_display_stage {
  uint8_t has_background_color;
  swfr_rgba8 background_color;
  uint32_t width, height;
  uint32_t n_children;
  const swfr_display_object *children;
} swfr_display_stage;

/* renderStage / drawDisplayObject / drawContainer (canvas-renderer.ts:69-94, 131-145): depth-first walk in paint
 * order; a container's matrix applies to everything below it (save / applyMatrix / restore).  Matrices are composed
 * in FP64 like Cairo's CTM and rounded to the Rust API's f32 Matrix2D at the leaves; morph primitives carry the
 * float ratio (SWFR_PRIM_RATIO_F32).  Host only.  Writes up to `cap` primitives, *n = the number needed. */
int swfr_flatten_display_stage(const swfr_display_stage *stage, swfr_display_primitive *out, uint32_t cap, uint32_t *n);

/* ---- lifecycle -------------------------------------------------------------------------------------- */

/* Creates a renderer with a width x height RGBA8 viewport on CUDA device `device`, with its own stream. */
int swfr_create(int device, uint32_t width, uint32_t height, swfr_renderer **out);
/* Same, with an existing CUDA stream (a cudaStream_t passed as void*) as the renderer's stream: every render is
 * ordered INTO it - the stream waits for the render's last kernel, so work the caller enqueues on it after swfr_render
 * sees finished frames, and events recorded on it time the renders.  The kernels themselves run on the renderer's own
 * pass streams (two passes of a batch overlap, and a render starts while the tail of its predecessor still runs). */
int swfr_create_on_stream(int device, uint32_t width, uint32_t height, void *cuda_stream, swfr_renderer **out);
void swfr_destroy(swfr_renderer *r);
const char *swfr_last_error(const swfr_renderer *r);
const char *swfr_status_string(int status);
uint32_t swfr_abi_version(void);

/* Options: SWFR_OPT_RETAIN_COMPILED (default 1) keeps the compiled paths of every definition on the host for
 * the swfr_debug_compiled / swfr_debug_segments taps; SWFR_OPT_FRAMES_PER_PASS (default 32) bounds how many frames
 * share one set of launches and one working set; SWFR_OPT_PROFILE = 1 records CUDA events at stage boundaries;
 * SWFR_OPT_HOST_THREADS = threads that flatten stages into draw items (0 = default: min(8, hardware threads)). */
typedef enum swfr_option {
  SWFR_OPT_RETAIN_COMPILED = 1,
  SWFR_OPT_FRAMES_PER_PASS = 2,
  SWFR_OPT_PROFILE = 3,
  SWFR_OPT_HOST_THREADS = 4,
  SWFR_OPT_OCCLUSION_CHUNKS = 7,   /* depth chunks for occlusion culling: the items of a frame are binned in this many
                                      ranges from the top (last painted) down, and geometry under an opaque full-tile
                                      cover found by an upper chunk is skipped.  0 (default): automatic (4 for frames
                                      of >= 1024 items, else 1); 1: no culling - every edge and record is produced, which
                                      the swfr_debug_edges / swfr_debug_tile_counts taps need; up to 8.  Pixels are
                                      identical for every setting */
  SWFR_OPT_DEBUG_TINY_ARENA = 6,   /* tests only: working arrays start at a few hundred entries, so every render has to
                                      grow them and re-run (swfr_stats.retries > 0) */
  SWFR_OPT_CLEAR_TO_BACKGROUND = 5 /* 0 (default): frames start transparent and Stage.background_color is ignored, like
                                      the TypeScript renderer and the headless Rust renderer (canvas-renderer.ts:70-72,
                                      headless_renderer.rs:611-615); 1: frames start from the opaque background colour,
                                      like the windowed Rust renderer (gfx_renderer.rs:292-301) */
} swfr_option;
int swfr_set_option(swfr_renderer *r, uint32_t key, uint64_t value);

/* ---- asset store ------------------------------------------------------------------------------------ */

int swfr_register_shape(swfr_renderer *r, const swfr_define_shape *tag, uint32_t *out_shape_id);
int swfr_register_morph_shape(swfr_renderer *r, const swfr_define_shape *tag, uint32_t *out_morph_shape_id);
/* straight (non-premultiplied) RGBA8 rows, `stride` bytes apart */
int swfr_register_bitmap(swfr_renderer *r, uint16_t bitmap_id, uint32_t width, uint32_t height, const uint8_t *rgba,
                         size_t stride);
/* DefineBitmap with media type image/x-swf-bmp (format 3: colormapped 8-bit + zlib) */
int swfr_register_bitmap_xswfbmp(swfr_renderer *r, uint16_t bitmap_id, const uint8_t *data, size_t len);
/* The decoder alone (host only): writes straight RGBA8 when `rgba` is non-NULL and cap is large enough. */
int swfr_decode_xswfbmp(const uint8_t *data, size_t len, uint8_t *rgba, uint64_t cap, uint32_t *width, uint32_t *height);

/* ---- rendering -------------------------------------------------------------------------------------- */

/* Renders one stage into frame 0.  Asynchronous on the renderer's stream. */
int swfr_render(swfr_renderer *r, const swfr_stage *stage);
/* Renders n stages into frames 0..n-1 with one set of launches per SWFR_OPT_FRAMES_PER_PASS frames.  Asynchronous:
 * the call flattens the stages on the host and uploads them while up to two earlier renders are still in flight,
 * waits for all but the newest of those and enqueues this one behind it - so a caller that streams batches (render k,
 * read k async, render k+1, ...) keeps the host flattening, both PCIe directions and the GPU busy at the same time.
 * A render is settled lazily (device counters: working memory that overflowed -> grown and the render re-run,
 * BitmapNotFound, statistics): such results surface at the next swfr_sync / swfr_read_image / swfr_get_stats. */
int swfr_render_batch(swfr_renderer *r, const swfr_stage *stages, uint32_t n);

/* Stages kept resident in HBM, for repeated rendering without host traffic. */
int swfr_batch_create(swfr_renderer *r, const swfr_stage *stages, uint32_t n, swfr_batch **out);
int swfr_batch_render(swfr_renderer *r, swfr_batch *batch);
void swfr_batch_destroy(swfr_renderer *r, swfr_batch *batch);

int swfr_sync(swfr_renderer *r);

/* Copies frame `frame` of the last render to host memory: RGBA8, `stride` >= 4*width bytes per row.
 * premultiplied != 0 returns the canvas' internal premultiplied pixels; 0 returns straight alpha with the
 * PNG-export rounding of the reference test (c = (c*255 + a/2) / a).  Synchronises the stream. */
int swfr_read_image(swfr_renderer *r, uint32_t frame, uint8_t *dst, size_t stride, int premultiplied);
/* Copies frames [first, first+count) premultiplied and tightly packed (width*height*4 bytes each) into pinned
 * or pageable host memory without synchronising.  While the render is in flight each pass' frames are copied on a
 * separate copy stream as soon as that pass is done, overlapping the passes that follow and - if the caller goes on
 * to the next swfr_render_batch - the start of the next render (which waits before overwriting a frame still being
 * copied).  `dst` is valid after swfr_sync(). */
int swfr_read_frames_async(swfr_renderer *r, uint32_t first, uint32_t count, uint8_t *dst);
/* Device pointer of frame 0 of the last render (premultiplied RGBA8, frames width*height*4 bytes apart). */
int swfr_device_frames(swfr_renderer *r, void **out_ptr, uint32_t *out_n_frames);

/* The outline the device stroker generates for one draw of a morph shape's strokes at `ratio` (lerped path and width,
 * round caps and joins: canvas-renderer.ts:252-266), computed by the same code on the host - no renderer, no CUDA: 8
 * doubles per segment (curve flag, line path index, x0, y0, cx, cy, x1, y1 in twips). */
int swfr_debug_morph_stroke(const swfr_define_shape *tag, double ratio, double *out, uint64_t cap, uint64_t *n_segs,
                            uint32_t *n_paths);

/* ---- optional gather of finished frames onto one GPU (off the hot path) --------------------
 * Frames shard over renderers (one per GPU, frame f -> renderer f mod N) with no collective; a caller that wants all
 * of them on one GPU afterwards moves them GPU to GPU over NVLink / NVSwitch with plain asynchronous peer copies:
 *   1. every source renderer:   swfr_sync(src); swfr_export_frames(src, &exp[k]);   (then any barrier / message)
 *   2. the gathering renderer:  swfr_gather_frames(dst, exp, N, &ptr, &n);  swfr_sync(dst);
 * exp[k] may come from another process (one process per GPU: it carries a CUDA IPC memory handle; the bytes of the
 * struct are what the processes exchange) or from the same process (one thread per GPU: the pointer is used directly
 * and peer access is enabled).  Frame slot s of exp[k] lands in slot s * N + k of the gather buffer, i.e. in global
 * frame order; one strided copy (cudaMemcpy2DAsync, row = one frame) per source, on the gathering renderer's copy
 * stream.  The exported store must stay untouched (no render on the source) until the gatherer's swfr_sync. */
typedef struct swfr_frames_export {
  uint8_t ipc_handle[64];  /* cudaIpcMemHandle_t of the frame store */
  uint64_t pid;            /* exporting process */
  uint64_t device_ptr;     /* frame 0 in the exporting process' address space */
  uint64_t offset;         /* of frame 0 inside the exported allocation */
  uint64_t frame_bytes;    /* width * height * 4 */
  int32_t device;          /* CUDA ordinal of the exporting renderer */
  uint32_t n_frames;       /* frames of its last render */
  uint32_t width, height;
} swfr_frames_export;
int swfr_export_frames(swfr_renderer *r, swfr_frames_export *out);
/* `*out_ptr` = the gather buffer on this renderer's device (premultiplied RGBA8, frames in global order), `*out_n` =
 * the frames it holds; valid after swfr_sync().  `*out_ms` (optional) is filled by the NEXT swfr_sync with the device
 * time of the copies (CUDA events on the copy stream). */
int swfr_gather_frames(swfr_renderer *r, const swfr_frames_export *sources, uint32_t n_sources, void **out_ptr, uint32_t *out_n);
int swfr_gather_last_ms(swfr_renderer *r, float *out_ms);

/* Renderer.render(stage) of the TypeScript API (ts/src/lib/renderer.ts:4-8, canvas-renderer.ts:61-78): flattens the
 * display tree and renders it into frame 0 (n trees into frames 0..n-1).  stage.width / height must equal the
 * renderer's viewport. */
int swfr_render_display_stage(swfr_renderer *r, const swfr_display_stage *stage);
int swfr_render_display_stages(swfr_renderer *r, const swfr_display_stage *stages, uint32_t n);

/* ---- image files (host only) -------------------------------------------------------------------------- */

/* write_pam (rs/src/pam.rs:3-34) / imageDataToPam (ts/src/lib/image-data-to-pam.ts:8-30): "P7" header + tight RGBA
 * rows.  Pass out = NULL to query the size; *n = bytes needed / written. */
int swfr_write_pam(const uint8_t *rgba, uint32_t width, uint32_t height, size_t stride, uint8_t *out, uint64_t cap, uint64_t *n);
/* canvas.toBuffer("image/png") of the reference test (node-canvas-renderer.spec.ts:134-147): 8-bit RGBA PNG of
 * straight-alpha pixels (swfr_read_image with premultiplied = 0). */
int swfr_write_png(const uint8_t *rgba, uint32_t width, uint32_t height, size_t stride, uint8_t *out, uint64_t cap, uint64_t *n);

/* ---- statistics of the last render (after swfr_sync) ------------------------------------------------ */

typedef struct swfr_stats {
  uint64_t n_primitives, n_path_instances, n_segments, n_edges, n_slots, n_records, n_tiles;
  uint64_t algorithmic_bytes; /* B_seg + B_draw + 2*8*E_tile + 4*W*H per frame, summed */
  uint64_t fine_slots;        /* (path, tile) slots the coverage kernel composited (after occlusion culling) */
  uint64_t fine_records;      /* binned records it read; n_records - fine_records were binned but hidden */
  uint32_t kernel_launches;   /* kernels launched by the last swfr_*render* call */
  uint32_t retries;           /* re-runs caused by working-memory growth */
} swfr_stats;
int swfr_get_stats(swfr_renderer *r, swfr_stats *out);
/* With SWFR_OPT_PROFILE = 1 every pass records CUDA events (on the renderer's stream) at its stage boundaries;
 * this returns the per-stage device time of the last render summed over its passes, and the number of passes
 * (= launches of each stage's kernels).  Stage names via swfr_stage_name(i). */
int swfr_get_stage_times(swfr_renderer *r, float *ms, uint32_t cap, uint32_t *n_stages, uint32_t *n_passes);
const char *swfr_stage_name(uint32_t i);

/* ---- parity taps ------------------------------------------------------------------------------------ */

/* Compiled path commands of a registered definition, in the reference's CommandType encoding
 * (LineTo=0, CurveTo=1, MoveTo=2; ts/src/lib/shape/path.ts:4-8).  Per command 1 + 8 doubles:
 * type, then x/endX, y/endY, controlX, controlY for the start state and the same four for the end state.
 * path_info per path: [n_commands, has_fill, has_line].  Pass NULL buffers to query sizes. */
int swfr_debug_compiled(swfr_renderer *r, uint32_t kind, uint32_t id, double *commands, uint64_t commands_cap,
                        uint64_t *n_commands, int32_t *path_info, uint64_t path_cap, uint64_t *n_paths);
/* The compiler alone (host only, no CUDA device needed): same outputs as swfr_debug_compiled and
 * swfr_debug_segments for a tag that is not registered anywhere. */
int swfr_compile_debug(const swfr_define_shape *tag, int morph, double *commands, uint64_t commands_cap,
                       uint64_t *n_commands, int32_t *path_info, uint64_t path_cap, uint64_t *n_paths, double *segs,
                       uint64_t segs_cap, uint64_t *n_segs);
/* Device segment store of a definition after implicit close + stroke expansion: per segment
 * [is_curve, path, x0,y0,cx,cy,x1,y1 (start), same six (end)] as doubles. */
int swfr_debug_segments(swfr_renderer *r, uint32_t kind, uint32_t id, double *segs, uint64_t cap, uint64_t *n);
/* Flattened 24.8 edges (x0,y0,x1,y1) and their path-instance index for one frame of the last render. */
int swfr_debug_edges(swfr_renderer *r, uint32_t frame, int32_t *edges, int32_t *edge_path, uint64_t cap, uint64_t *n);
/* Binned records per 16x16 tile (row-major ceil(h/16) x ceil(w/16)) for one frame of the last render. */
int swfr_debug_tile_counts(swfr_renderer *r, uint32_t frame, uint32_t *counts, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SWFR_H */
