// Device data layout and kernel launchers of libswfr_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "host_types.h"
#include "stroke_core.h"

namespace swfr {

constexpr int kTile = 16;          // px
constexpr int kTileFx = 4096;      // 16 px in 1/256 px
constexpr int kMaxLenFx = 16384;   // flattened edges span at most 64 px per axis
constexpr int kNumSM = 148;        // B200
constexpr int kGroupTiles = 8;     // tile columns per candidate list (128 px)
constexpr int kStageBlock = 256;   // staging entries per block (blocks are private to one warp of the binning pass)
constexpr int kMaxChunks = 8;      // depth chunks for occlusion culling
constexpr int kBackdropSmall = 1024;  // grids up to this many slots get their backdrop prefix from one warp

// One draw item after host flattening of the stage: 48 bytes.
struct DrawItem {
  float m[6];           // Matrix2D order
  uint32_t seg_first;   // first segment of the definition in its store
  uint32_t paint_first; // first DefPaint of the definition
  uint32_t path_off;    // first path instance of this item
  uint32_t frame;       // frame index within the batch
  uint16_t ratio;       // MorphRatio (rs/src/stage.rs:28-34): ratio / 65535
  uint16_t kind;        // ITEM_STATIC / ITEM_MORPH / ITEM_DYNAMIC, | ITEM_RATIO_F32 when ratio_f32 replaces ratio
  float ratio_f32;      // the TypeScript renderer's ratio (a number in 0..1, morph-shape.ts:5-10)
};
enum : uint16_t {
  ITEM_STATIC = 0,   // segments in the static store, paints in the definition paint store
  ITEM_MORPH = 1,    // start/end segments in the morph store
  ITEM_DYNAMIC = 2,  // segments and paints in the batch's own stores (morph-shape strokes expanded for this draw)
  ITEM_KIND_MASK = 0xff,
  ITEM_RATIO_F32 = 0x100,
  ITEM_CX = 0x200  // the draw carries a colour transform (RenderArgs.item_cx)
};

// Per path instance, read by binning and by the fine kernel: 16 bytes.
struct PathRec {
  uint32_t xy0;    // bx0 | by0 << 16   (tile bbox origin)
  uint32_t wh;     // bw  | bh  << 16   (0 when empty)
  uint32_t info;   // paint type | flags << 8 | frame << 16   (flag bit0: opaque source, bit1: stroke outline -> sampled coverage)
  uint32_t color;  // premultiplied RGBA8 of a solid paint
};

// Per path instance paint parameters for non-solid paints: 80 bytes (five 16-byte words: k_fine stages the
// instance it composites in shared memory with one load per lane).
struct PaintInst {
  float inv[6];  // device px -> fill space
  float focal, omf;  // gradients: focal point, 1 - focal^2; bitmaps: 1 / rx, 1 / ry
  float rx, ry;      // bitmaps: footprint in texels; gradients: rx = 1 / omf
  unsigned long long ptr;  // ramp pointer (gradients) or texture object (bitmaps)
  int32_t bw, bh;          // bitmap size
  uint32_t spread, repeating;
  float inv_bw, inv_bh;    // bitmaps: 1 / bw, 1 / bh (float division)
  uint32_t cx_on, pad;     // cx_on != 0: the draw's colour transform applies to every evaluated pixel (k_fine<true> only)
  int16_t cx[8];           // swfr_color_transform: red, green, blue, alpha mult (8.8), then the four adds
};
static_assert(sizeof(PaintInst) == 96, "PaintInst is staged as six uint4");

// One draw of a morph shape's strokes: the device stroker (k_stroke) writes its outline into the batch's dynamic
// segment store, seg_cap entries from seg_first on (unused ones are marked empty), and the bounds of every visible line
// path into the batch's dynamic paints.
struct StrokeJob {
  uint32_t line_first, line_count;  // the morph shape's lines (stroke::LineDev)
  uint32_t seg_first, seg_cap;      // into the dynamic segment store
  uint32_t paint_first, path_count; // into the dynamic paints: the lines visible at this ratio, in order
  uint32_t item, morph_id;          // host bookkeeping: the draw item of the outline, the morph shape
  double ratio;
};
constexpr uint32_t kNullSegment = 0xffffffffu;  // path_flags of an unused entry of the dynamic segment store

struct BitmapDev {
  unsigned long long tex;
  int32_t w, h;
  uint32_t opaque, valid;
};

// Counters written by the device, read by the host after a sync.
constexpr int kFineSliceFrames = 16;  // k_fine is launched per slice of this many frames, so that finished frames can
constexpr int kMaxFineSlices = 16;    // leave for the host while the rest of the pass is still being composited

struct Totals {
  uint32_t n_edges, n_slots, n_records;
  uint32_t overflow;  // bit0 edges, bit1 slots, bit2 records, bit3 candidate lists, bit4 row lists, bit5 arena dirty (see k_init),
                      // bit6 a stroke outline outgrew the segments reserved for it (k_stroke)
  uint32_t error;     // bit0: unknown bitmap id
  uint32_t work[kMaxFineSlices];  // fine-kernel tile queues, one per slice of frames
  uint32_t n_list;    // candidate-list entries
  uint32_t n_big_chunk[8];    // per depth chunk: visible path instances with large tile grids
  uint32_t n_small_chunk[8];  // ... with small tile grids
  uint32_t n_alive_items[8];  // ... draw items with a visible path
  uint32_t n_rowent;  // row-list entries
  uint32_t n_stage_blocks;  // staging blocks handed out by the binning pass
  uint32_t overflow_stage;  // the staging buffer was too small (found while binning, after the scans)
  uint32_t fine_hits;       // (path, tile) slots composited by k_fine
  uint32_t fine_records;    // records read by k_fine (the rest of n_records was binned but hidden)
};

struct Caps {
  uint32_t edges, slots, records, list, rows, stage;  // stage: entries, a multiple of the staging block size
};

// Everything a render launch needs (device pointers unless noted).
struct RenderArgs {
  int width, height, tiles_x, tiles_y;
  uint32_t n_items, n_seginst, n_paths, n_frames;  // host-known
  const DrawItem *items;
  const uint32_t *item_seg_off;   // n_items + 1
  const uint32_t *item_path_off;  // n_items + 1
  const uint32_t *frame_path_off; // n_frames + 1
  const uint32_t *frame_bg;       // n_frames: premultiplied RGBA8 every pixel of the frame starts from
  const SegStatic *segs_static;
  const SegMorph *segs_morph;
  const DefPaint *def_paints;
  SegStatic *segs_dynamic;         // per batch (ITEM_DYNAMIC): written by k_stroke
  DefPaint *paints_dynamic;        // per batch (ITEM_DYNAMIC): bounds written by k_stroke
  const StrokeJob *jobs;           // stroke jobs of the pass
  uint32_t n_jobs;                 // host-known
  const stroke::LineDev *mlines;   // morph lines and their commands (asset store)
  const stroke::CmdDev *mcmds;
  const uint32_t *ramps;  // kRampSize premultiplied RGBA8 entries per gradient
  const BitmapDev *bitmaps;
  uint32_t *seg_edge_off;   // n_seginst + 1 (piece counts, then exclusive scan)
  uint32_t *seg_item;       // n_seginst: draw item of each segment instance
  PathRec *path_rec;        // n_paths
  PaintInst *paint_inst;    // n_paths
  uint32_t *path_slot_off;  // n_paths + 1
  uint32_t *path_rec_base;  // n_paths: first record of the path instance
  int4 *edges;              // caps.edges
  uint32_t *edge_pid;       // caps.edges: path instance of each edge
  uint32_t *slot_count;     // caps.slots (+1): records per slot while binning; all zero between renders (k_scatter counts back down)
  int32_t *slot_backdrop;   // caps.slots: winding deltas posted while binning; all zero between chunks (k_cover zeroes what it reads)
  int32_t *slot_wind;       // caps.slots: winding number at the left edge of the slot's tile (k_cover -> k_fine)
  uint32_t *slot_off;       // caps.slots + 1: end of the slot's record range, relative to path_rec_base
  unsigned long long *records;  // caps.records
  uint4 *stage;             // caps.stage: (record lo, record hi, slot, path instance) in binning order
  uint32_t *stage_used;     // caps.stage / 256: entries used in each staging block
  // occlusion culling by depth chunks: the items of every frame are split into n_chunks ranges
  // in paint order; chunks are binned from the top one down, and what a chunk finds completely covered by an opaque
  // path hides the geometry of the chunks below it
  uint32_t has_cx;              // host-known: a draw of the pass carries a colour transform (k_fine<true> as well)
  const int16_t *item_cx;       // [n_items][8] colour transforms (swfr_color_transform), or null when has_cx == 0
  uint32_t has_sampled;         // host-known: the pass draws stroke outlines (k_fine<true>: coverage by sub-scanlines for them)
  uint32_t n_chunks;            // host-known
  const uint32_t *chunk_items;  // (n_chunks + 1) * n_frames: first item of chunk c in frame f at [c * n_frames + f]
  uint32_t *tile_cover;         // n_frames * tiles: 1 + the highest path instance that covers the tile opaquely, 0 = none
  uint32_t *path_alive;         // n_paths: 0 when every tile of the path's bbox is covered from above
  uint32_t *cover_bits;         // n_frames * tiles_y * cover_words: one bit per tile, set when tile_cover != 0
  uint32_t cover_words;         // host-known: 32-bit words per tile row = ceil(tiles_x / 32)
  uint32_t *chunk_edge;         // kMaxChunks + 1: edge cursor at the start of each depth chunk (unordered edges)
  uint32_t *frames;         // n_frames * width * height
  uint32_t *scan_tmp;       // >= 4096 words
  // candidate lists: for every (frame, tile row, group of kGroupTiles tile columns) the path instances whose
  // tile bbox overlaps it, in paint order
  uint32_t groups_x, n_lists;  // host-known: n_lists = n_frames * tiles_y * groups_x
  uint32_t *list_off;          // n_lists + 1 (counts, then exclusive scan)
  uint32_t *list_items;        // caps.list
  uint32_t *row_count;         // n_frames * tiles_y: path instances whose bbox covers the tile row
  uint32_t *row_off;           // n_frames * tiles_y + 1
  uint2 *row_items;            // caps.rows: (path instance, bx0 | bw << 16) per row, in paint order
  uint32_t *big_chunk;         // n_paths: ... of the depth chunk being processed
  uint32_t *small_chunk;       // n_paths: visible path instances with small tile grids of the depth chunk being processed
  uint32_t *path_item;         // n_paths: draw item of each path instance
  uint32_t *item_alive;        // n_items: 1 when any path instance of the draw item is visible
  uint32_t *alive_items;       // n_items: those draw items, listed per depth chunk
  uint32_t *alive_paths;       // n_paths: the visible path instances of frame f, in paint order, at [frame_path_off[f] ..)
  uint32_t *alive_count;       // n_frames: how many
  uint32_t *arena_dirty;       // 1 word per arena: set when a render was aborted (its self-cleaning arrays are dirty)
  Totals *totals;
  Caps caps;
};

// Pipeline stages as seen by the profiler hooks (swfr_get_stage_times).
static_assert(kMaxChunks == 8, "Totals holds 8 per-chunk counters");
constexpr int kNumStages = 6;
const char *stage_name(int i);


// Enqueues every kernel of one render on `stream`; returns the number of kernels launched.
// `ev` (optional) points at kNumStages + 1 events recorded at the stage boundaries.
// `slice_done` (optional): one event per slice of frames (fine_slices(a.n_frames) of them), recorded when that slice's
// frames are final.
int launch_render(const RenderArgs &a, cudaStream_t stream, cudaEvent_t *ev = nullptr, cudaEvent_t *slice_done = nullptr);
// Frames per slice / number of slices of a pass of n_frames frames.
inline uint32_t fine_slice_frames(uint32_t n_frames) {
  uint32_t f = kFineSliceFrames;
  while ((n_frames + f - 1) / f > (uint32_t)kMaxFineSlices) f *= 2;
  return f;
}
inline uint32_t fine_slices(uint32_t n_frames) {
  uint32_t f = fine_slice_frames(n_frames);
  return n_frames ? (n_frames + f - 1) / f : 1;
}

void launch_unpremultiply(const uint32_t *src, uint32_t *dst, uint64_t n_px, cudaStream_t stream);
void launch_premultiply(const uint8_t *src, size_t stride, uint32_t w, uint32_t h, uint32_t *dst, uint32_t *translucent,
                        cudaStream_t stream);
void launch_xswfbmp_expand(const uint8_t *inflated, uint32_t colors, uint32_t w, uint32_t h, uint32_t padded, uint32_t *dst,
                           cudaStream_t stream);
void launch_tile_counts(const RenderArgs &a, uint32_t frame, uint32_t *counts, cudaStream_t stream);

}  // namespace swfr
