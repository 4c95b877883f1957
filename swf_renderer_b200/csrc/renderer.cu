// Host side of libswfr_b200: asset store, stage flattening, working-memory management and the C ABI
// (include/swfr.h).  Mirrors the reference Rust crate's surface:
//   rs/src/asset.rs:9-20            ClientAssetStore / ServerAssetStore
//   rs/src/swf_renderer.rs:3-5      SwfRenderer::render(stage)
//   rs/src/stage.rs:4-59            Stage, DisplayPrimitive, Matrix2D, MorphRatio
//   rs/src/renderer.rs:81-103       Renderer::set_stage, Image{meta,data}
//   rs/src/headless_renderer.rs:60-64, 229-244, 725-868   new / define_shape / get_image / download_image
// There is no CPU fallback: every entry point that renders needs a CUDA device.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <new>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached

#include "kernels.h"
#include "stroke_core.h"

using namespace swfr;

namespace {
// NVTX range over a scope: host flattening, uploads, the enqueue of every pass and the settling of a
// render show up as named ranges on the timeline of nsys / ncu, with the kernels of the pass nested under them.
struct Range {
  explicit Range(const char *name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};
}  // namespace

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t reserve(size_t bytes, bool keep = false, cudaStream_t st = 0) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = std::max(bytes, cap + cap / 2);
    ncap = (ncap + 255) & ~(size_t)255;
    void *np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) return e;
    if (p) {
      if (keep) cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
      cudaStreamSynchronize(st);
      cudaFree(p);
    }
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t ncap = (bytes + bytes / 4 + 4095) & ~(size_t)4095;
    cudaError_t e = cudaHostAlloc(&p, ncap, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = ncap;
    return e;
  }
};

// One set of launches: a contiguous range of frames whose working set shares the arena.
struct Pass {
  uint32_t f0 = 0, n_frames = 0, n_items = 0, n_seginst = 0, n_paths = 0;
  // offsets into the batch-wide host/device arrays
  size_t items_at = 0, seg_off_at = 0, path_off_at = 0, frame_off_at = 0, chunk_at = 0;
  uint32_t n_chunks = 1;  // depth chunks for occlusion culling
  uint32_t job_first = 0, n_jobs = 0;  // stroke jobs of the pass' frames
  bool has_sampled = false;  // a frame of the pass draws a stroke outline: k_fine with the sub-scanline coverage routine
  bool has_cx = false;       // a draw of the pass carries a colour transform
};

struct BitmapRes {
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex = 0;
};

}  // namespace

// Host arrays of a batch live in pinned memory (written once by the stage flattener, read by the H2D copies).
template <class T>
struct PinnedArr {
  PinnedBuf buf;
  size_t n = 0;
  cudaError_t resize(size_t count) {
    n = count;
    return buf.reserve(std::max<size_t>(count * sizeof(T), 64));
  }
  T *data() const { return reinterpret_cast<T *>(buf.p); }
  T &operator[](size_t i) const { return data()[i]; }
  size_t size() const { return n; }
  size_t bytes() const { return n * sizeof(T); }
};

struct swfr_batch {
  uint32_t n_frames = 0;
  std::vector<Pass> passes;
  PinnedArr<DrawItem> items;
  PinnedArr<swfr_color_transform> item_cx;  // parallel to items; empty unless a draw of the batch carries a colour transform
  PinnedArr<uint32_t> seg_off, path_off, frame_off;  // concatenated per pass ([n+1] each)
  PinnedArr<uint32_t> chunk_items;                   // per pass: (n_chunks + 1) x n_frames first items of the depth chunks
  PinnedArr<uint32_t> frame_bg;                      // per frame: premultiplied RGBA8 the frame starts from
  // Strokes of morph shapes: every such draw is a job for the device stroker (k_stroke), which writes the outline's
  // segments into the batch's dynamic segment store (device only: dyn_seg_count entries) and the bounds into its paints
  PinnedArr<DefPaint> dyn_paints;
  PinnedArr<StrokeJob> jobs;  // in frame order: the jobs of a pass are contiguous
  size_t dyn_seg_count = 0;
  DevBuf d_item_cx;
  DevBuf d_items, d_seg_off, d_path_off, d_frame_off, d_dyn_segs, d_dyn_paints, d_frame_bg, d_chunk_items, d_jobs;
  bool resident = false;
  uint64_t n_prims = 0, n_seginst = 0, n_paths = 0;
  cudaEvent_t uploaded = nullptr;  // recorded on the upload stream after the H2D copies
  ~swfr_batch() {
    if (uploaded) cudaEventDestroy(uploaded);
  }
};

struct swfr_renderer {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint32_t width = 0, height = 0, tiles_x = 0, tiles_y = 0;
  std::string last_error;
  bool retain_compiled = true;
  uint32_t frames_per_pass = 32;  // measured best on the 10 k shapes stream (16: 3.31, 32: 2.88, 64: 3.06 ms per 64-frame step)

  // ---- asset store (host mirrors + device copies) ----
  std::vector<SegStatic> h_static;
  std::vector<SegMorph> h_morph;
  std::vector<DefPaint> h_paints;
  std::vector<uint32_t> h_ramps;
  std::vector<DefEntry> shape_defs, morph_defs;
  // Morph shapes with visible strokes: their line paths live on the device (lines + commands, both morph states); the
  // outline of every draw is generated there (k_stroke).  seg_cap = segments reserved per draw: the largest outline
  // found at five ratios at registration plus a margin; an outline that outgrows it is found by the device and the
  // batch is laid out again with exact counts (relayout_strokes).
  struct MorphStroke {
    uint32_t line_first = 0, line_count = 0;  // into h_mlines; line_count == 0: no visible stroke
    uint32_t seg_cap = 0;
  };
  std::vector<MorphStroke> morph_strokes;  // by morph shape id
  std::vector<stroke::LineDev> h_mlines;
  std::vector<stroke::CmdDev> h_mcmds;
  DevBuf d_mlines, d_mcmds;
  size_t up_mlines = 0, up_mcmds = 0;
  std::vector<std::unique_ptr<CompiledDef>> shape_dbg, morph_dbg;
  DevBuf d_static, d_morph, d_paints, d_ramps, d_bitmaps;
  size_t up_static = 0, up_morph = 0, up_paints = 0, up_ramps = 0;  // elements already uploaded
  std::vector<BitmapRes> bitmaps;                                   // by id, lazily sized 65536
  std::vector<BitmapDev> h_bitmaps;

  // ---- working memory ----
  // Arenas: consecutive passes of a batch alternate between them and between the pass streams, so that the many
  // short, latency-bound kernels at the front of one pass run under the long coverage kernel of its neighbour.
  struct Arena {
    DevBuf seg_edge_off, seg_item, path_rec, paint_inst, path_slot_off, path_rec_base, edges, edge_pid, slot_count, slot_backdrop, slot_wind, slot_off, records, scan_tmp, list_off, list_items, row_count, row_off, row_items, stage, stage_used, tile_cover, path_alive, cover_bits, big_chunk, small_chunk, path_item, item_alive, alive_items, alive_paths, alive_count, chunk_edge, dirty;
  };
  static constexpr int kArenas = 4;
  Arena arena[kArenas];
  int n_arenas = 2;                         // arenas (and pass streams) in use: pass i runs in arena / on pass stream i % n_arenas
  // Passes never run on `stream` itself: pass i of every render goes to pass_stream[i % n_arenas] (so the passes of
  // consecutive renders that share an arena are ordered by their stream, and nothing else orders them: render k + 1
  // starts while the last passes of render k are still running), and `stream` only waits for the render's last
  // kernels (join), so that whatever the caller enqueues on it afterwards sees finished frames.
  cudaStream_t pass_stream[kArenas] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t join_ev[kArenas] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t render_done = nullptr;  // recorded on `stream` after the join of the newest render
  // pass layout of the newest render: a render with a different layout writes frame ranges on other streams than its
  // predecessor did, and then waits for the predecessor as a whole
  uint32_t layout_frames = 0, layout_fpp = 0;
  DevBuf frames, scratch, scratch2;
  Caps caps{0, 0, 0, 0, 0, 0};
  // Renders in flight.  A render is *settled* (its device counters read back: overflow of the working memory ->
  // grow and re-run, errors, statistics) lazily: when a later render needs the host to catch up, at swfr_sync, or
  // before anything reads its results synchronously - never between two renders that are enqueued back to back.
  static constexpr int kSlots = 3;  // counter blocks / scratch batches: at most two renders are in flight
  struct CopyReq {
    uint32_t first, count;
    uint8_t *dst;
  };
  struct InFlight {
    swfr_batch *b = nullptr;
    int slot = 0;           // totals[slot] / pin_totals[slot]
    uint32_t launches = 0;
    std::vector<CopyReq> copy_reqs;  // swfr_read_frames_async requests (re-issued if the render has to be re-run)
  };
  std::deque<InFlight> inflight;  // oldest first
  DevBuf totals[kSlots];
  PinnedBuf pin_totals[kSlots];
  cudaEvent_t slot_done[kSlots] = {nullptr, nullptr, nullptr};  // recorded on `stream` after the counters' D2H copy
  int next_slot = 0;
  swfr_batch scratch_batch[kSlots];  // swfr_render / swfr_render_batch rotate, so that the stages of render k + 1 are
  int scratch_ix = 0;                // flattened and uploaded while renders k - 1 and k are still on the GPU
  uint32_t host_threads = 0;    // stage flattening threads (0 = min(8, hardware))
  bool clear_to_background = false;
  uint32_t occlusion_chunks = 0;  // 0 = automatic, 1 = no culling, n = n depth chunks
  bool tiny_arena = false;  // debug: start every working array at a few hundred entries so that growth + re-run is exercised
  cudaStream_t up_stream = nullptr;

  // ---- newest render ----
  swfr_batch *last = nullptr;
  uint32_t frames_rendered = 0;
  swfr_stats stats{};  // of the newest settled render
  std::vector<Totals> last_totals;
  size_t arena_pass = 0;  // index of the pass whose working set is in the arena
  bool profile = false;   // record CUDA events at the stage boundaries of every pass
  std::vector<cudaEvent_t> prof_events;
  size_t prof_passes = 0;
  float stage_ms[kNumStages] = {0};
  uint32_t stage_launches = 0;
  // readback overlapped with rendering: one event per slice of a pass, copies on their own stream
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> pass_done;  // kMaxFineSlices events per pass: slice k of pass i at [i * kMaxFineSlices + k]
  bool copy_pending = false;
  // frame ranges whose device->host copy may still be in flight when the next render starts: that render's pass
  // which overwrites the range waits for the event (recorded on the copy stream)
  struct CopyFence {
    uint32_t first, count;
    cudaEvent_t done;
  };
  std::vector<CopyFence> copy_fences;
  std::vector<cudaEvent_t> fence_pool;
  // optional gather of other renderers' frames (swfr_gather_frames)
  DevBuf gathered;
  cudaEvent_t gather_ev[2] = {nullptr, nullptr};
  bool gather_timed = false;
  float gather_ms = 0.f;
  struct Import {
    uint8_t handle[64];
    void *base;
  };
  std::vector<Import> imports;  // IPC allocations opened so far (kept open: opening is expensive)
};

namespace {

int fail(swfr_renderer *r, int code, const std::string &msg) {
  if (r) r->last_error = msg;
  return code;
}

// No C++ exception may cross the C ABI: host allocations and the shape compiler run inside this guard.
template <class F>
int guarded(swfr_renderer *r, F &&f) {
  try {
    return f();
  } catch (const std::bad_alloc &) {
    return fail(r, SWFR_ERR_OOM, "host allocation failed");
  } catch (const std::exception &e) {
    return fail(r, SWFR_ERR_INVALID_ARGUMENT, e.what());
  } catch (...) {
    return fail(r, SWFR_ERR_INVALID_ARGUMENT, "unexpected exception");
  }
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t _e = (call);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return fail(r, _e == cudaErrorMemoryAllocation ? SWFR_ERR_OOM : SWFR_ERR_CUDA,                    \
                  std::string(#call) + ": " + cudaGetErrorString(_e));                                  \
  } while (0)

int flush_store(swfr_renderer *r) {
  cudaStream_t st = r->stream;
  auto up = [&](DevBuf &d, const void *h, size_t elem, size_t n, size_t &done) -> cudaError_t {
    if (n == done) return cudaSuccess;
    cudaError_t e = d.reserve(std::max<size_t>(n * elem, 256), true, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync((char *)d.p + done * elem, (const char *)h + done * elem, (n - done) * elem,
                        cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host vectors may be reallocated by later registrations
    done = n;
    return e;
  };
  CK(up(r->d_static, r->h_static.data(), sizeof(SegStatic), r->h_static.size(), r->up_static));
  CK(up(r->d_morph, r->h_morph.data(), sizeof(SegMorph), r->h_morph.size(), r->up_morph));
  CK(up(r->d_paints, r->h_paints.data(), sizeof(DefPaint), r->h_paints.size(), r->up_paints));
  CK(up(r->d_ramps, r->h_ramps.data(), sizeof(uint32_t), r->h_ramps.size(), r->up_ramps));
  CK(up(r->d_mlines, r->h_mlines.data(), sizeof(stroke::LineDev), r->h_mlines.size(), r->up_mlines));
  CK(up(r->d_mcmds, r->h_mcmds.data(), sizeof(stroke::CmdDev), r->h_mcmds.size(), r->up_mcmds));
  CK(r->d_mlines.reserve(256));
  CK(r->d_mcmds.reserve(256));
  CK(r->d_static.reserve(256));
  CK(r->d_morph.reserve(256));
  CK(r->d_paints.reserve(256));
  CK(r->d_ramps.reserve(256));
  if (!r->d_bitmaps.p) {
    CK(r->d_bitmaps.reserve(65536 * sizeof(BitmapDev)));
    CK(cudaMemsetAsync(r->d_bitmaps.p, 0, 65536 * sizeof(BitmapDev), st));
  }
  return SWFR_OK;
}

// Morph lines as the device stroker reads them: lines + commands, both morph states, appended to the two stores.
static void morph_lines_to_device(const std::vector<MorphLine> &src, std::vector<stroke::LineDev> &lines, std::vector<stroke::CmdDev> &cmds) {
  for (const MorphLine &ml : src) {
    stroke::LineDev ln{};
    ln.cmd_first = (uint32_t)cmds.size();
    ln.cmd_count = (uint32_t)ml.commands.size();
    ln.w0 = ml.w0;
    ln.w1 = ml.w1;
    memcpy(ln.color0, ml.color0, 4);
    memcpy(ln.color1, ml.color1, 4);
    for (const Command &c : ml.commands) {
      stroke::CmdDev cd{};
      cd.type = c.type;
      for (int k = 0; k < 4; k++) cd.s[k] = c.s[k], cd.e[k] = c.e[k];
      cmds.push_back(cd);
    }
    lines.push_back(ln);
  }
}

// Visible line paths and outline segments of one draw of a morph shape's strokes at ratio r (host run of the
// generator the device runs per draw: registration-time estimate and the exact counts of a relayout).
static void count_morph_strokes(const swfr_renderer *r, const swfr_renderer::MorphStroke &ms, double ratio, uint32_t *n_paths,
                                uint32_t *n_segs) {
  double width_state = 1.0;
  uint32_t paths = 0, segs = 0;
  for (uint32_t l = 0; l < ms.line_count; l++) {
    const stroke::LineDev &ln = r->h_mlines[ms.line_first + l];
    const double w = stroke::lerp(ln.w0, ln.w1, ratio);
    if (w > 0) width_state = w;
    const double al = stroke::lerp(ln.color0[3] / 255.0, ln.color1[3] / 255.0, ratio);
    if (al <= 0) continue;  // composites nothing
    if (n_segs) {
      stroke::Sink sink{nullptr, 0, 0, paths, {1.f, 1.f, 0.f, 0.f}, false, 0.f, 0.f, 0.f, 0.f};
      stroke::stroke_line(r->h_mcmds.data() + ln.cmd_first, ln.cmd_count, ratio, width_state, sink);
      segs += sink.n;
    }
    paths++;
  }
  if (n_paths) *n_paths = paths;
  if (n_segs) *n_segs = segs;
}

// Flattens stages into draw items (reference: CanvasRenderer.renderStage / drawDisplayObject,
// ts/src/lib/renderers/canvas-renderer.ts:69-94) and splits them into passes.  Two sweeps over the stages, both
// parallel over frames: (1) validate ids, count segment / path instances per frame and expand the strokes of morph
// shapes for their ratio, (2) after a prefix over the frames of each pass, write the draw items and their offsets
// straight into the batch's pinned arrays.
int build_batch(swfr_renderer *r, const swfr_stage *stages, uint32_t n, swfr_batch &b) {
  Range range("swfr: flatten stages");
  b.n_frames = n;
  b.passes.clear();
  b.n_prims = b.n_seginst = b.n_paths = 0;
  b.resident = false;
  const uint32_t fpp = std::max<uint32_t>(1, r->frames_per_pass);
  uint64_t total_prims = 0;
  for (uint32_t f = 0; f < n; f++) {
    if (stages[f].n_primitives && !stages[f].display_root)
      return fail(r, SWFR_ERR_INVALID_ARGUMENT, "stage.display_root is NULL");
    total_prims += stages[f].n_primitives;
  }
  if (total_prims > 0x7ffffff0ull) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "too many display primitives");
  uint32_t nt = r->host_threads ? r->host_threads : std::min<uint32_t>(8, std::max(1u, std::thread::hardware_concurrency()));
  nt = std::max<uint32_t>(1, std::min<uint32_t>(nt, n));
  if (total_prims < 20000) nt = 1;

  struct DynItem {  // stroke outlines of one morph primitive (a job of the device stroker)
    uint32_t prim, seg_at, seg_count, paint_at, paint_count, morph_id;
    double ratio;
  };
  struct FrameSum {
    uint64_t seg = 0, path = 0, items = 0;
    bool sampled = false;  // the frame draws a stroke outline
    bool cx = false;       // a draw of the frame carries a colour transform
    int err = SWFR_OK;
    uint32_t bad_id = 0;
    std::vector<DynItem> dyn;
    uint64_t dyn_segs = 0;  // segments reserved for the frame's stroke outlines
    std::vector<DefPaint> dyn_paints;
  };
  std::vector<FrameSum> sums(n);
  auto lookup = [&](const swfr_display_primitive &pr, int &err) -> const DefEntry * {
    if (pr.kind == SWFR_PRIM_SHAPE) {
      if (pr.id >= r->shape_defs.size()) {
        err = SWFR_ERR_INVALID_ID;
        return nullptr;
      }
      return &r->shape_defs[pr.id];
    }
    if (pr.kind == SWFR_PRIM_MORPH_SHAPE) {
      if (pr.id >= r->morph_defs.size()) {
        err = SWFR_ERR_INVALID_ID;
        return nullptr;
      }
      return &r->morph_defs[pr.id];
    }
    err = SWFR_ERR_INVALID_ARGUMENT;
    return nullptr;
  };
  // an identity transform is no transform (un-premultiplying and premultiplying a gradient pixel again is lossy)
  auto prim_cx = [](const swfr_display_primitive &pr) -> bool {
    if (!(pr.flags & SWFR_PRIM_COLOR_TRANSFORM)) return false;
    const swfr_color_transform &t = pr.color_transform;
    return !(t.red_mult == 256 && t.green_mult == 256 && t.blue_mult == 256 && t.alpha_mult == 256 && t.red_add == 0 &&
             t.green_add == 0 && t.blue_add == 0 && t.alpha_add == 0);
  };
  auto prim_ratio = [](const swfr_display_primitive &pr) -> double {
    return (pr.flags & SWFR_PRIM_RATIO_F32) ? (double)pr.ratio_f : (double)pr.ratio / 65535.0;
  };
  auto run = [&](auto &&fn) {
    // frames t, t + nt, ... go to worker t; if the process cannot start a thread, the caller does that share itself
    // (no exception may cross the C ABI)
    std::vector<std::thread> th;
    std::vector<uint32_t> mine{0u};
    for (uint32_t t = 1; t < nt; t++) {
      try {
        th.emplace_back(fn, t);
      } catch (...) {
        mine.push_back(t);
      }
    }
    for (uint32_t t : mine) fn(t);
    for (std::thread &x : th) x.join();
  };
  run([&](uint32_t t) {
    for (uint32_t f = t; f < n; f += nt) try {  // an exception in a worker thread must not reach std::terminate
      FrameSum &s = sums[f];
      const swfr_stage &st = stages[f];
      for (uint32_t i = 0; i < st.n_primitives; i++) {
        const swfr_display_primitive &pr = st.display_root[i];
        int err = SWFR_OK;
        const DefEntry *de = lookup(pr, err);
        if (!de) {
          s.err = err;
          s.bad_id = pr.id;
          break;
        }
        s.seg += de->seg_count;
        s.path += de->path_count;
        s.items += 1;
        s.sampled |= de->has_sampled != 0;
        s.cx |= prim_cx(pr);
        if (pr.kind == SWFR_PRIM_MORPH_SHAPE && r->morph_strokes[pr.id].line_count) {
          // the lines that are visible at this ratio (lerped alpha > 0) become the paths of one more draw item; their
          // geometry is left to the device (no per-draw host geometry), `seg_cap` segments are reserved for it
          const swfr_renderer::MorphStroke &ms = r->morph_strokes[pr.id];
          DynItem d;
          d.prim = i;
          d.morph_id = pr.id;
          d.ratio = prim_ratio(pr);
          d.seg_at = (uint32_t)s.dyn_segs;
          d.paint_at = (uint32_t)s.dyn_paints.size();
          for (uint32_t l = 0; l < ms.line_count; l++) {
            const stroke::LineDev &ln = r->h_mlines[ms.line_first + l];
            if (stroke::lerp(ln.color0[3] / 255.0, ln.color1[3] / 255.0, d.ratio) <= 0) continue;  // composites nothing
            DefPaint p{};
            p.type = PAINT_SOLID;
            p.lut = -1;
            memcpy(p.color0, ln.color0, 4);
            memcpy(p.color1, ln.color1, 4);
            p.flags |= PF_COLOR_MORPH | PF_SAMPLED;
            p.bounds[0] = p.bounds[1] = 1.0f;  // empty until the device stroker has written the outline's bounds
            s.dyn_paints.push_back(p);
          }
          // (SWFR_OPT_DEBUG_TINY_ARENA: far too little room, so that the relayout after an outgrown outline is exercised)
          d.seg_count = r->tiny_arena ? std::min<uint32_t>(ms.seg_cap, 8u) : ms.seg_cap;
          d.paint_count = (uint32_t)s.dyn_paints.size() - d.paint_at;
          if (d.paint_count) {
            s.sampled = true;
            s.dyn_segs += d.seg_count;
            s.dyn.push_back(d);
            s.seg += d.seg_count;
            s.path += d.paint_count;
            s.items += 1;
          }
        }
      }
    } catch (...) {
      sums[f].err = SWFR_ERR_OOM;
    }
  });
  for (uint32_t f = 0; f < n; f++) {
    if (sums[f].err == SWFR_ERR_OOM) return fail(r, SWFR_ERR_OOM, "host allocation failed while flattening the stages");
    if (sums[f].err == SWFR_ERR_INVALID_ID) return fail(r, SWFR_ERR_INVALID_ID, "unknown shape id " + std::to_string(sums[f].bad_id));
    if (sums[f].err != SWFR_OK) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "unknown display primitive kind");
  }

  // pass layout + per-frame bases
  struct FrameBase {
    size_t item_at, seg_off_at, path_off_at, frame_off_at, dyn_seg_at, dyn_paint_at, job_at;
    uint32_t seg0, path0;
  };
  std::vector<FrameBase> base(n);
  size_t items_at = 0, seg_off_at = 0, path_off_at = 0, frame_off_at = 0, dyn_seg_at = 0, dyn_paint_at = 0, job_at = 0;
  for (uint32_t f0 = 0; f0 < n; f0 += fpp) {
    Pass p;
    p.f0 = f0;
    p.n_frames = std::min(fpp, n - f0);
    p.items_at = items_at;
    p.seg_off_at = seg_off_at;
    p.path_off_at = path_off_at;
    p.frame_off_at = frame_off_at;
    uint64_t seg_run = 0, path_run = 0;
    size_t it_run = 0;
    p.job_first = (uint32_t)job_at;
    for (uint32_t f = f0; f < f0 + p.n_frames; f++) {
      base[f] = FrameBase{items_at + it_run, seg_off_at + it_run, path_off_at + it_run, frame_off_at + (f - f0),
                          dyn_seg_at,        dyn_paint_at,        job_at,               (uint32_t)seg_run,    (uint32_t)path_run};
      seg_run += sums[f].seg;
      path_run += sums[f].path;
      it_run += sums[f].items;
      p.has_sampled |= sums[f].sampled;
      p.has_cx |= sums[f].cx;
      dyn_seg_at += sums[f].dyn_segs;
      dyn_paint_at += sums[f].dyn_paints.size();
      job_at += sums[f].dyn.size();
    }
    p.n_jobs = (uint32_t)job_at - p.job_first;
    if (seg_run > 0xfffffff0ull || path_run > 0xfffffff0ull || it_run > 0xfffffff0ull)
      return fail(r, SWFR_ERR_INVALID_ARGUMENT, "a pass exceeds 2^32 segment instances; lower SWFR_OPT_FRAMES_PER_PASS");
    p.n_items = (uint32_t)it_run;
    p.n_seginst = (uint32_t)seg_run;
    p.n_paths = (uint32_t)path_run;
    items_at += it_run;
    seg_off_at += it_run + 1;
    path_off_at += it_run + 1;
    frame_off_at += p.n_frames + 1;
    b.n_seginst += seg_run;
    b.n_paths += path_run;
    b.passes.push_back(p);
  }
  if (dyn_seg_at > 0xfffffff0ull) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "too many stroke segments in one batch");
  // Depth chunks (occlusion culling): the items of every frame are split in paint order into n_chunks ranges whose
  // sizes double from the top (last painted) chunk downwards - most tiles are covered by the first few opaque
  // shapes from the top, and everything under a cover is neither flattened nor binned.
  size_t chunk_at = 0;
  for (Pass &p : b.passes) {
    uint64_t max_items = 0;
    for (uint32_t f = p.f0; f < p.f0 + p.n_frames; f++) max_items = std::max<uint64_t>(max_items, sums[f].items);
    p.n_chunks = r->occlusion_chunks ? r->occlusion_chunks : (max_items >= 1024 ? 4u : 1u);
    p.chunk_at = chunk_at;
    chunk_at += (size_t)(p.n_chunks + 1) * p.n_frames;
  }
  CK(b.chunk_items.resize(chunk_at));
  for (const Pass &p : b.passes) {
    uint32_t it_run = 0;
    const uint64_t denom = (1ull << p.n_chunks) - 1;
    for (uint32_t f = p.f0; f < p.f0 + p.n_frames; f++) {
      const uint64_t ni = sums[f].items;
      for (uint32_t c = 0; c <= p.n_chunks; c++) {
        // chunk c starts after the bottom (2^n - 2^(n-c)) / (2^n - 1) of the items
        const uint64_t above = ((1ull << (p.n_chunks - c)) - 1) * ni / denom;
        b.chunk_items[p.chunk_at + (size_t)c * p.n_frames + (f - p.f0)] = it_run + (uint32_t)(ni - above);
      }
      it_run += (uint32_t)ni;
    }
  }
  b.n_prims = total_prims;
  CK(b.items.resize(items_at));
  bool any_cx = false;
  for (const Pass &p : b.passes) any_cx |= p.has_cx;
  CK(b.item_cx.resize(any_cx ? items_at : 0));
  CK(b.seg_off.resize(seg_off_at));
  CK(b.path_off.resize(path_off_at));
  CK(b.frame_off.resize(frame_off_at));
  b.dyn_seg_count = dyn_seg_at;
  CK(b.jobs.resize(job_at));
  CK(b.dyn_paints.resize(dyn_paint_at));
  CK(b.frame_bg.resize(n));
  for (uint32_t f = 0; f < n; f++) {
    const swfr_rgba8 &c = stages[f].background_color;
    // gfx_renderer.rs:292-301: clear colour (r, g, b, 1.0); otherwise transparent black (canvas-renderer.ts:70-72)
    b.frame_bg[f] = r->clear_to_background ? ((uint32_t)c.r | ((uint32_t)c.g << 8) | ((uint32_t)c.b << 16) | 0xff000000u) : 0u;
  }
  for (const Pass &p : b.passes) {  // closing entries of each pass
    b.seg_off[p.seg_off_at + p.n_items] = p.n_seginst;
    b.path_off[p.path_off_at + p.n_items] = p.n_paths;
    b.frame_off[p.frame_off_at + p.n_frames] = p.n_paths;
  }
  run([&](uint32_t t) {
    for (uint32_t f = t; f < n; f += nt) {
      const swfr_stage &st = stages[f];
      const FrameBase &fb = base[f];
      const FrameSum &fs = sums[f];
      DrawItem *items = b.items.data() + fb.item_at;
      swfr_color_transform *cxs = b.item_cx.size() ? b.item_cx.data() + fb.item_at : nullptr;
      const swfr_color_transform cx_identity = {256, 256, 256, 256, 0, 0, 0, 0};
      uint32_t *so = b.seg_off.data() + fb.seg_off_at, *po = b.path_off.data() + fb.path_off_at;
      uint32_t seg_run = fb.seg0, path_run = fb.path0;
      b.frame_off[fb.frame_off_at] = path_run;
      const uint32_t local_frame = f % fpp;
      if (!fs.dyn_paints.empty())
        memcpy(b.dyn_paints.data() + fb.dyn_paint_at, fs.dyn_paints.data(), fs.dyn_paints.size() * sizeof(DefPaint));
      size_t k = 0, next_dyn = 0;
      for (uint32_t i = 0; i < st.n_primitives; i++) {
        const swfr_display_primitive &pr = st.display_root[i];
        int err = SWFR_OK;
        const DefEntry *de = lookup(pr, err);
        DrawItem it;
        memcpy(it.m, pr.matrix, sizeof it.m);
        it.seg_first = de->seg_first;
        it.paint_first = de->paint_first;
        it.path_off = path_run;
        it.frame = local_frame;
        it.ratio = pr.ratio;
        const bool has_cx = prim_cx(pr);
        const uint16_t item_flags = (uint16_t)(((pr.flags & SWFR_PRIM_RATIO_F32) ? ITEM_RATIO_F32 : 0) | (has_cx ? ITEM_CX : 0));
        it.kind = (uint16_t)((de->is_morph ? ITEM_MORPH : ITEM_STATIC) | item_flags);
        it.ratio_f32 = pr.ratio_f;
        items[k] = it;
        if (cxs) cxs[k] = has_cx ? pr.color_transform : cx_identity;
        so[k] = seg_run;
        po[k] = path_run;
        k++;
        seg_run += de->seg_count;
        path_run += de->path_count;
        if (next_dyn < fs.dyn.size() && fs.dyn[next_dyn].prim == i) {  // the primitive's strokes, painted after its fills
          const DynItem &d = fs.dyn[next_dyn];
          it.seg_first = (uint32_t)(fb.dyn_seg_at + d.seg_at);
          it.paint_first = (uint32_t)(fb.dyn_paint_at + d.paint_at);
          it.path_off = path_run;
          it.kind = (uint16_t)(ITEM_DYNAMIC | item_flags);
          items[k] = it;
          if (cxs) cxs[k] = has_cx ? pr.color_transform : cx_identity;
          const swfr_renderer::MorphStroke &ms = r->morph_strokes[d.morph_id];
          StrokeJob job{};
          job.line_first = ms.line_first;
          job.line_count = ms.line_count;
          job.seg_first = it.seg_first;
          job.seg_cap = d.seg_count;
          job.paint_first = it.paint_first;
          job.path_count = d.paint_count;
          job.item = (uint32_t)(fb.item_at + k);  // index into the batch's items
          job.morph_id = d.morph_id;
          job.ratio = d.ratio;
          b.jobs[fb.job_at + next_dyn] = job;
          next_dyn++;
          so[k] = seg_run;
          po[k] = path_run;
          k++;
          seg_run += d.seg_count;
          path_run += d.paint_count;
        }
      }
    }
  });
  return SWFR_OK;
}

// Enqueues the H2D copies of a batch on the upload stream (its host arrays are pinned and its device arrays are its
// own, so this may run while the previous render is still on the GPU) and records b.uploaded.
int upload_batch(swfr_renderer *r, swfr_batch &b) {
  Range range("swfr: upload batch");
  if (!r->up_stream) CK(cudaStreamCreateWithFlags(&r->up_stream, cudaStreamNonBlocking));
  if (!b.uploaded) CK(cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming));
  cudaStream_t st = r->up_stream;
  CK(b.d_items.reserve(std::max<size_t>(b.items.bytes(), 256)));
  CK(b.d_seg_off.reserve(std::max<size_t>(b.seg_off.bytes(), 256)));
  CK(b.d_path_off.reserve(std::max<size_t>(b.path_off.bytes(), 256)));
  CK(b.d_frame_off.reserve(std::max<size_t>(b.frame_off.bytes(), 256)));
  if (b.items.bytes()) CK(cudaMemcpyAsync(b.d_items.p, b.items.data(), b.items.bytes(), cudaMemcpyHostToDevice, st));
  if (b.item_cx.bytes()) {
    CK(b.d_item_cx.reserve(b.item_cx.bytes()));
    CK(cudaMemcpyAsync(b.d_item_cx.p, b.item_cx.data(), b.item_cx.bytes(), cudaMemcpyHostToDevice, st));
  }
  if (b.seg_off.bytes()) CK(cudaMemcpyAsync(b.d_seg_off.p, b.seg_off.data(), b.seg_off.bytes(), cudaMemcpyHostToDevice, st));
  if (b.path_off.bytes()) CK(cudaMemcpyAsync(b.d_path_off.p, b.path_off.data(), b.path_off.bytes(), cudaMemcpyHostToDevice, st));
  if (b.frame_off.bytes())
    CK(cudaMemcpyAsync(b.d_frame_off.p, b.frame_off.data(), b.frame_off.bytes(), cudaMemcpyHostToDevice, st));
  CK(b.d_chunk_items.reserve(std::max<size_t>(b.chunk_items.bytes(), 256)));
  if (b.chunk_items.bytes())
    CK(cudaMemcpyAsync(b.d_chunk_items.p, b.chunk_items.data(), b.chunk_items.bytes(), cudaMemcpyHostToDevice, st));
  CK(b.d_frame_bg.reserve(std::max<size_t>(b.frame_bg.bytes(), 256)));
  if (b.frame_bg.bytes()) CK(cudaMemcpyAsync(b.d_frame_bg.p, b.frame_bg.data(), b.frame_bg.bytes(), cudaMemcpyHostToDevice, st));
  CK(b.d_dyn_segs.reserve(std::max<size_t>(b.dyn_seg_count * sizeof(SegStatic), 256)));  // written by the device stroker
  CK(b.d_dyn_paints.reserve(std::max<size_t>(b.dyn_paints.bytes(), 256)));
  CK(b.d_jobs.reserve(std::max<size_t>(b.jobs.bytes(), 256)));
  if (b.jobs.bytes()) CK(cudaMemcpyAsync(b.d_jobs.p, b.jobs.data(), b.jobs.bytes(), cudaMemcpyHostToDevice, st));
  if (b.dyn_paints.bytes())
    CK(cudaMemcpyAsync(b.d_dyn_paints.p, b.dyn_paints.data(), b.dyn_paints.bytes(), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(b.uploaded, st));
  b.resident = true;
  return SWFR_OK;
}

int finish(swfr_renderer *r);

// Sizes the working memory for batch b.  Runs twice: a dry run that only finds out whether any array has to grow -
// if so, the renders in flight are settled first (they use the arrays that would be freed) - then the real one.
int ensure_arena(swfr_renderer *r, const swfr_batch &b) {
  uint32_t max_seg = 0, max_paths = 0, max_items = 0, max_frames = 1;
  for (const Pass &p : b.passes) {
    max_seg = std::max(max_seg, p.n_seginst);
    max_paths = std::max(max_paths, p.n_paths);
    max_items = std::max(max_items, p.n_items);
    max_frames = std::max(max_frames, p.n_frames);
  }
  const int n_arenas = (int)std::min<size_t>(b.passes.size(), (size_t)r->n_arenas);  // a single pass needs one working set
  Caps want = r->caps;
  if (r->tiny_arena) {  // SWFR_OPT_DEBUG_TINY_ARENA: no heuristics, every array has to grow through settle()
    want.edges = std::max<uint32_t>(want.edges, 512);
    want.slots = std::max<uint32_t>(want.slots, 512);
    want.records = std::max<uint32_t>(want.records, 512);
    want.list = std::max<uint32_t>(want.list, 512);
    want.rows = std::max<uint32_t>(want.rows, 512);
    want.stage = std::max<uint32_t>(want.stage, 1024);
  } else {
    want.edges = std::max<uint32_t>(want.edges, std::max<uint32_t>(1u << 16, max_seg * 4));
    want.slots = std::max<uint32_t>(want.slots, std::max<uint32_t>(1u << 16, max_paths * 32));
    want.records = std::max<uint32_t>(want.records, std::max<uint32_t>(1u << 17, want.edges * 2));
    want.list = std::max<uint32_t>(want.list, std::max<uint32_t>(1u << 16, max_paths * 12));
    want.rows = std::max<uint32_t>(want.rows, std::max<uint32_t>(1u << 16, max_paths * 8));
    // staging: the records, plus one partly filled block per binning warp (at most kNumSM * 16 * 8 warps, one per 32 edges)
    uint64_t warps = std::min<uint64_t>((uint64_t)kNumSM * 16 * 8, (uint64_t)want.edges / 32 + 1);
    uint64_t st = (uint64_t)want.records + want.records / 4 + 2 * warps * kStageBlock + 65535u;  // a partly filled block + one taken in advance per warp
    want.stage = std::max<uint32_t>(want.stage, (uint32_t)std::min<uint64_t>(st & ~255ull, 0xffffff00ull));
  }
  const uint32_t groups_x = (r->tiles_x + kGroupTiles - 1) / kGroupTiles;
  size_t max_lists = (size_t)std::max<uint32_t>(1, r->frames_per_pass) * r->tiles_y * groups_x;
  for (const Pass &p : b.passes) max_lists = std::max(max_lists, (size_t)p.n_frames * r->tiles_y * groups_x);
  const size_t max_rows = max_lists / groups_x;
  const size_t tiles = (size_t)r->tiles_x * r->tiles_y;
  for (int dry = 1; dry >= 0; dry--) {
    bool grow = false;
    cudaError_t err = cudaSuccess;
    auto need = [&](DevBuf &d, size_t bytes, bool zeroed = false) {
      if (bytes <= d.cap || err != cudaSuccess) return;
      if (dry) {
        grow = true;
        return;
      }
      err = d.reserve(bytes);
      // self-cleaning arrays (slot record counters, winding deltas) start from zero
      if (zeroed && err == cudaSuccess) err = cudaMemset(d.p, 0, d.cap);
    };
    for (int k = 0; k < n_arenas; k++) {
      swfr_renderer::Arena &A = r->arena[k];
      need(A.seg_edge_off, ((size_t)max_seg + 1) * 4 + 256);
      need(A.seg_item, (size_t)max_seg * 4 + 256);
      need(A.path_rec, (size_t)max_paths * sizeof(PathRec) + 256);
      need(A.paint_inst, (size_t)max_paths * sizeof(PaintInst) + 256);
      need(A.path_slot_off, ((size_t)max_paths + 1) * 4 + 256);
      need(A.path_rec_base, ((size_t)max_paths + 1) * 4 + 256);
      need(A.big_chunk, (size_t)max_paths * 4 + 256);
      need(A.small_chunk, (size_t)max_paths * 4 + 256);
      need(A.path_item, (size_t)max_paths * 4 + 256);
      need(A.item_alive, (size_t)max_items * 4 + 256);
      need(A.alive_items, (size_t)max_items * 4 + 256);
      need(A.path_alive, (size_t)max_paths * 4 + 256);
      need(A.alive_paths, (size_t)max_paths * 4 + 256);
      need(A.alive_count, (size_t)max_frames * 4 + 256);
      need(A.tile_cover, (size_t)max_frames * tiles * 4 + 256);
      need(A.cover_bits, (size_t)max_frames * (r->tiles_x + 1) * (r->tiles_y + 1) * 4 + 256);
      need(A.scan_tmp, 8192 * 4);
      need(A.chunk_edge, 64 * 4);
      need(A.dirty, 256, true);
      need(A.list_off, (max_lists + 1) * 4 + 256);
      need(A.list_items, (size_t)want.list * 4);
      need(A.row_count, (max_rows + 1) * 4 + 256);
      need(A.row_off, (max_rows + 1) * 4 + 256);
      need(A.row_items, (size_t)want.rows * 8);
      need(A.edges, (size_t)want.edges * 16);
      need(A.edge_pid, (size_t)want.edges * 4);
      need(A.slot_count, ((size_t)want.slots + 1) * 4, true);
      need(A.slot_backdrop, ((size_t)want.slots + 1) * 4, true);
      need(A.slot_wind, ((size_t)want.slots + 1) * 4);
      need(A.slot_off, ((size_t)want.slots + 1) * 4);
      need(A.records, (size_t)want.records * 8);
      need(A.stage, (size_t)want.stage * 16);
      need(A.stage_used, (size_t)(want.stage / 256 + 1) * 4);
    }
    for (int k = 0; k < swfr_renderer::kSlots; k++) need(r->totals[k], std::max<size_t>(b.passes.size(), 1) * sizeof(Totals));
    need(r->frames, std::max<size_t>((size_t)b.n_frames * r->width * r->height * 4, 256));
    for (int k = 0; k < swfr_renderer::kSlots; k++)  // pinned mirrors of the counters: a D2H copy may be in flight into them
      if (std::max<size_t>(b.passes.size(), 1) * sizeof(Totals) > r->pin_totals[k].cap) {
        if (dry)
          grow = true;
        else if (err == cudaSuccess)
          err = r->pin_totals[k].reserve(std::max<size_t>(b.passes.size(), 16) * sizeof(Totals));
      }
    if (err != cudaSuccess)
      return fail(r, err == cudaErrorMemoryAllocation ? SWFR_ERR_OOM : SWFR_ERR_CUDA,
                  std::string("working memory: ") + cudaGetErrorString(err));
    if (dry && !grow) break;
    if (dry) {
      int rc = finish(r);
      if (rc != SWFR_OK) return rc;
    }
  }
  r->caps = want;
  return SWFR_OK;
}

RenderArgs make_args(swfr_renderer *r, const swfr_batch &b, const Pass &p, size_t pass_index, int slot) {
  RenderArgs a{};
  const swfr_renderer::Arena &A = r->arena[pass_index % (size_t)r->n_arenas];
  a.width = (int)r->width;
  a.height = (int)r->height;
  a.tiles_x = (int)r->tiles_x;
  a.tiles_y = (int)r->tiles_y;
  a.n_items = p.n_items;
  a.n_seginst = p.n_seginst;
  a.n_paths = p.n_paths;
  a.n_frames = p.n_frames;
  a.items = b.d_items.as<DrawItem>() + p.items_at;
  a.item_seg_off = b.d_seg_off.as<uint32_t>() + p.seg_off_at;
  a.item_path_off = b.d_path_off.as<uint32_t>() + p.path_off_at;
  a.frame_path_off = b.d_frame_off.as<uint32_t>() + p.frame_off_at;
  a.segs_static = r->d_static.as<SegStatic>();
  a.segs_morph = r->d_morph.as<SegMorph>();
  a.def_paints = r->d_paints.as<DefPaint>();
  a.frame_bg = b.d_frame_bg.as<uint32_t>() + p.f0;
  a.n_chunks = p.n_chunks;
  a.has_sampled = p.has_sampled ? 1u : 0u;
  a.has_cx = p.has_cx ? 1u : 0u;
  a.item_cx = p.has_cx ? b.d_item_cx.as<int16_t>() + p.items_at * 8 : nullptr;
  a.chunk_items = b.d_chunk_items.as<uint32_t>() + p.chunk_at;
  a.tile_cover = A.tile_cover.as<uint32_t>();
  a.path_alive = A.path_alive.as<uint32_t>();
  a.cover_bits = A.cover_bits.as<uint32_t>();
  a.cover_words = (r->tiles_x + 31) / 32;
  a.chunk_edge = A.chunk_edge.as<uint32_t>();
  a.arena_dirty = A.dirty.as<uint32_t>();
  a.segs_dynamic = b.d_dyn_segs.as<SegStatic>();
  a.paints_dynamic = b.d_dyn_paints.as<DefPaint>();
  a.jobs = b.d_jobs.as<StrokeJob>() + p.job_first;
  a.n_jobs = p.n_jobs;
  a.mlines = r->d_mlines.as<stroke::LineDev>();
  a.mcmds = r->d_mcmds.as<stroke::CmdDev>();
  a.ramps = r->d_ramps.as<uint32_t>();
  a.bitmaps = r->d_bitmaps.as<BitmapDev>();
  a.seg_edge_off = A.seg_edge_off.as<uint32_t>();
  a.seg_item = A.seg_item.as<uint32_t>();
  a.path_rec = A.path_rec.as<PathRec>();
  a.paint_inst = A.paint_inst.as<PaintInst>();
  a.path_slot_off = A.path_slot_off.as<uint32_t>();
  a.path_rec_base = A.path_rec_base.as<uint32_t>();
  a.edges = A.edges.as<int4>();
  a.edge_pid = A.edge_pid.as<uint32_t>();
  a.slot_count = A.slot_count.as<uint32_t>();
  a.slot_backdrop = A.slot_backdrop.as<int32_t>();
  a.slot_wind = A.slot_wind.as<int32_t>();
  a.slot_off = A.slot_off.as<uint32_t>();
  a.records = A.records.as<unsigned long long>();
  a.stage = A.stage.as<uint4>();
  a.stage_used = A.stage_used.as<uint32_t>();
  a.frames = r->frames.as<uint32_t>() + (size_t)p.f0 * r->width * r->height;
  a.scan_tmp = A.scan_tmp.as<uint32_t>();
  a.groups_x = (r->tiles_x + kGroupTiles - 1) / kGroupTiles;
  a.n_lists = p.n_frames * r->tiles_y * a.groups_x;
  a.list_off = A.list_off.as<uint32_t>();
  a.row_count = A.row_count.as<uint32_t>();
  a.row_off = A.row_off.as<uint32_t>();
  a.row_items = A.row_items.as<uint2>();
  a.list_items = A.list_items.as<uint32_t>();
  a.big_chunk = A.big_chunk.as<uint32_t>();
  a.small_chunk = A.small_chunk.as<uint32_t>();
  a.path_item = A.path_item.as<uint32_t>();
  a.item_alive = A.item_alive.as<uint32_t>();
  a.alive_items = A.alive_items.as<uint32_t>();
  a.alive_paths = A.alive_paths.as<uint32_t>();
  a.alive_count = A.alive_count.as<uint32_t>();
  a.totals = r->totals[slot].as<Totals>() + pass_index;
  a.caps = r->caps;
  return a;
}

int settle(swfr_renderer *r, size_t keep);

int ensure_streams(swfr_renderer *r, int n) {
  for (int k = 0; k < n; k++)
    if (!r->pass_stream[k]) {
      CK(cudaStreamCreateWithFlags(&r->pass_stream[k], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&r->join_ev[k], cudaEventDisableTiming));
    }
  if (!r->render_done) CK(cudaEventCreateWithFlags(&r->render_done, cudaEventDisableTiming));
  for (int k = 0; k < swfr_renderer::kSlots; k++)
    if (!r->slot_done[k]) CK(cudaEventCreateWithFlags(&r->slot_done[k], cudaEventDisableTiming | cudaEventBlockingSync));
  return SWFR_OK;
}

// Enqueues every pass of batch b.  `settled_first`: the caller has synchronised everything (re-run after an overflow):
// all passes go to `stream` one after the other.
int enqueue_passes(swfr_renderer *r, swfr_batch &b, int slot, bool serial, uint32_t *launches_out) {
  uint32_t launches = 0;
  const bool overlap = !serial && b.passes.size() > 1 && !r->profile && r->n_arenas > 1;
  const int n_streams = overlap ? (int)std::min<size_t>(b.passes.size(), (size_t)r->n_arenas) : 1;
  if (!serial) {
    // a render laid out differently from its predecessor writes frame ranges from other streams than that one did
    const uint32_t fpp = std::max<uint32_t>(1, r->frames_per_pass);
    const bool same = r->layout_frames == b.n_frames && r->layout_fpp == (overlap ? fpp : 0u);
    for (int k = 0; k < n_streams; k++) {
      if (b.uploaded) CK(cudaStreamWaitEvent(r->pass_stream[k], b.uploaded, 0));
      if (!same) CK(cudaStreamWaitEvent(r->pass_stream[k], r->render_done, 0));
    }
    r->layout_frames = b.n_frames;
    r->layout_fpp = overlap ? fpp : 0u;
  }
  for (size_t i = 0; i < b.passes.size(); i++) {
    const int sk = overlap ? (int)(i % (size_t)r->n_arenas) : 0;
    cudaStream_t st = serial ? r->stream : r->pass_stream[sk];
    if (!serial)  // a device->host copy of an earlier render may still be reading the frames this pass overwrites
      for (const swfr_renderer::CopyFence &cf : r->copy_fences)
        if (cf.first < b.passes[i].f0 + b.passes[i].n_frames && b.passes[i].f0 < cf.first + cf.count)
          CK(cudaStreamWaitEvent(st, cf.done, 0));
    Range range("swfr: enqueue pass");
    launches += (uint32_t)launch_render(make_args(r, b, b.passes[i], i, slot), st,
                                        (r->profile && !serial) ? r->prof_events.data() + i * (kNumStages + 1) : nullptr,
                                        r->pass_done.data() + i * kMaxFineSlices);
  }
  if (!serial) {
    for (int k = 0; k < n_streams; k++) {
      CK(cudaEventRecord(r->join_ev[k], r->pass_stream[k]));
      CK(cudaStreamWaitEvent(r->stream, r->join_ev[k], 0));
    }
    CK(cudaEventRecord(r->render_done, r->stream));
    for (const swfr_renderer::CopyFence &cf : r->copy_fences) r->fence_pool.push_back(cf.done);
    r->copy_fences.clear();
  }
  CK(cudaMemcpyAsync(r->pin_totals[slot].p, r->totals[slot].p, b.passes.size() * sizeof(Totals), cudaMemcpyDeviceToHost, r->stream));
  CK(cudaEventRecord(r->slot_done[slot], r->stream));
  CK(cudaGetLastError());
  if (launches_out) *launches_out = launches;
  return SWFR_OK;
}

int launch_batch(swfr_renderer *r, swfr_batch &b) {
  // The newest render in flight keeps the GPU busy while this one is enqueued behind it; older ones are settled now
  // (their counters are usually back already).  Profiling runs read per-render events: nothing stays in flight.
  int rc = settle(r, r->profile ? 0 : 1);
  if (rc != SWFR_OK) return rc;
  rc = flush_store(r);
  if (rc != SWFR_OK) return rc;
  rc = ensure_arena(r, b);
  if (rc != SWFR_OK) return rc;
  rc = ensure_streams(r, r->n_arenas);
  if (rc != SWFR_OK) return rc;
  r->prof_passes = 0;
  if (r->profile) {
    size_t need = b.passes.size() * (kNumStages + 1);
    while (r->prof_events.size() < need) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      r->prof_events.push_back(e);
    }
    r->prof_passes = b.passes.size();
  }
  while (r->pass_done.size() < b.passes.size() * kMaxFineSlices) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    r->pass_done.push_back(e);
  }
  swfr_renderer::InFlight in;
  in.b = &b;
  in.slot = r->next_slot;
  r->next_slot = (r->next_slot + 1) % swfr_renderer::kSlots;
  rc = enqueue_passes(r, b, in.slot, false, &in.launches);
  if (rc != SWFR_OK) return rc;
  r->inflight.push_back(std::move(in));
  r->last = &b;
  r->arena_pass = b.passes.empty() ? 0 : b.passes.size() - 1;
  r->frames_rendered = b.n_frames;
  return SWFR_OK;
}

// Statistics and device-reported errors of a render whose counters are in `tot`.
int account(swfr_renderer *r, const swfr_batch &b, const std::vector<Totals> &tot, uint32_t launches, uint32_t retries) {
  memset(&r->stats, 0, sizeof r->stats);
  r->stats.kernel_launches = launches;
  r->stats.retries = retries;
  uint32_t err = 0;
  r->stats.n_primitives = b.n_prims;
  r->stats.n_segments = b.n_seginst;
  r->stats.n_path_instances = b.n_paths;
  r->stats.n_tiles = (uint64_t)r->tiles_x * r->tiles_y * b.n_frames;
  for (const Totals &t : tot) {
    r->stats.n_edges += t.n_edges;
    r->stats.n_slots += t.n_slots;
    r->stats.n_records += t.n_records;
    r->stats.fine_slots += t.fine_hits;
    r->stats.fine_records += t.fine_records;
    err |= t.error;
  }
  // segments read once, draw items read once, binned records written + read once (8 B each),
  // framebuffer written once (frames start from a clear, so there is no load).
  r->stats.algorithmic_bytes = b.n_seginst * sizeof(SegStatic) + b.n_prims * sizeof(DrawItem) +
                               2ull * 8ull * r->stats.n_records + 4ull * r->width * r->height * b.n_frames;
  if (err & 1u) return fail(r, SWFR_ERR_INVALID_ID, "BitmapNotFound: a bitmap fill references an unregistered bitmap id");
  return SWFR_OK;
}

bool any_overflow(const Totals *t, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (t[i].overflow | t[i].overflow_stage) return true;
  return false;
}

// A stroke outline outgrew the segments reserved for it (overflow bit 6, found by k_stroke): the exact size of every
// outline of the batch is counted on the host (the same generator, count only), the dynamic segment store and the
// per-item segment offsets of every pass are laid out again, and the estimate of the morph shapes is raised so that
// the next batch reserves enough.  Everything is synchronised when this runs (recover).
int relayout_strokes(swfr_renderer *r, swfr_batch &b) {
  Range range("swfr: relayout stroke outlines");
  std::vector<uint32_t> exact(b.jobs.size());
  for (size_t j = 0; j < b.jobs.size(); j++) {
    swfr_renderer::MorphStroke &ms = r->morph_strokes[b.jobs[j].morph_id];
    count_morph_strokes(r, ms, b.jobs[j].ratio, nullptr, &exact[j]);
    ms.seg_cap = std::max(ms.seg_cap, exact[j] + exact[j] / 4 + 16);
  }
  size_t ji = 0, dyn_at = 0;
  b.n_seginst = 0;
  for (Pass &p : b.passes) {
    uint32_t *so = b.seg_off.data() + p.seg_off_at;
    DrawItem *items = b.items.data() + p.items_at;
    uint64_t seg_run = 0;
    uint32_t prev_old = so[0];
    for (uint32_t k = 0; k < p.n_items; k++) {
      const uint32_t next_old = so[k + 1];
      uint32_t cnt = next_old - prev_old;
      prev_old = next_old;
      if ((items[k].kind & ITEM_KIND_MASK) == ITEM_DYNAMIC) {
        if (ji >= b.jobs.size() || b.jobs[ji].item != p.items_at + k) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "stroke jobs out of step");
        cnt = exact[ji];
        items[k].seg_first = (uint32_t)dyn_at;
        b.jobs[ji].seg_first = (uint32_t)dyn_at;
        b.jobs[ji].seg_cap = cnt;
        dyn_at += cnt;
        ji++;
      }
      so[k] = (uint32_t)seg_run;
      seg_run += cnt;
    }
    if (seg_run > 0xfffffff0ull) return fail(r, SWFR_ERR_OOM, "a pass exceeds 2^32 segment instances");
    so[p.n_items] = (uint32_t)seg_run;
    p.n_seginst = (uint32_t)seg_run;
    b.n_seginst += seg_run;
  }
  b.dyn_seg_count = dyn_at;
  int rc = upload_batch(r, b);
  if (rc != SWFR_OK) return rc;
  CK(cudaStreamSynchronize(r->up_stream));
  return SWFR_OK;
}

// Working memory overflowed in the oldest render in flight (the later ones ran with the same arrays and have
// overwritten its frames): everything is synchronised, then every render in flight is run again, in order and alone,
// growing the arrays until it fits, and its read-back requests are served again.
int recover(swfr_renderer *r) {
  Range range("swfr: grow working memory and re-run");
  CK(cudaStreamSynchronize(r->stream));
  if (r->copy_stream) CK(cudaStreamSynchronize(r->copy_stream));
  const size_t fb = (size_t)r->width * r->height * 4;
  int result = SWFR_OK;
  while (!r->inflight.empty()) {
    swfr_renderer::InFlight in = std::move(r->inflight.front());
    r->inflight.pop_front();
    swfr_batch &b = *in.b;
    const size_t np = b.passes.size();
    uint32_t retries = 0;
    const Totals *pt = reinterpret_cast<const Totals *>(r->pin_totals[in.slot].p);
    for (int guard = 0;; guard++) {
      if (any_overflow(pt, np)) {
        if (guard >= 12) {
          r->inflight.clear();
          return fail(r, SWFR_ERR_OOM, "working memory kept overflowing");
        }
        Caps want = r->caps;
        bool strokes_outgrown = false;
        auto grow = [](uint32_t need) { return (uint32_t)std::min<uint64_t>((uint64_t)need + need / 4 + 1024, 0xfffffff0ull); };
        for (size_t i = 0; i < np; i++) {
          const Totals &t = pt[i];
          if (t.overflow & 1u) want.edges = std::max(want.edges, grow(t.n_edges));
          if (t.overflow & 2u) want.slots = std::max(want.slots, grow(t.n_slots));
          if (t.overflow & 4u) want.records = std::max(want.records, grow(t.n_records));
          if (t.overflow & 8u) want.list = std::max(want.list, grow(t.n_list));
          if (t.overflow & 16u) want.rows = std::max(want.rows, grow(t.n_rowent));
          if (t.overflow & 1u) want.records = std::max(want.records, want.edges * 2);
          if (t.overflow_stage) {
            uint64_t need = ((uint64_t)t.n_stage_blocks + t.n_stage_blocks / 8 + 64) * 256;
            want.stage = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want.stage, need), 0xffffff00ull);
          }
          strokes_outgrown |= (t.overflow & 64u) != 0;
        }
        if (strokes_outgrown) {
          int rc = relayout_strokes(r, b);
          if (rc != SWFR_OK) {
            r->inflight.clear();
            return rc;
          }
          uint32_t max_seg = 0;
          for (const Pass &p : b.passes) max_seg = std::max(max_seg, p.n_seginst);
          for (int k = 0; k < (int)std::min<size_t>(np, (size_t)r->n_arenas); k++) {
            CK(r->arena[k].seg_edge_off.reserve(((size_t)max_seg + 1) * 4 + 256));
            CK(r->arena[k].seg_item.reserve((size_t)max_seg * 4 + 256));
          }
          want.edges = std::max<uint32_t>(want.edges, std::max<uint32_t>(1u << 16, max_seg * 4));
        }
        if (!r->tiny_arena)
          want.stage = std::max<uint32_t>(want.stage, (uint32_t)std::min<uint64_t>(((uint64_t)want.records + want.records / 4 + 65535u) & ~255ull, 0xffffff00ull));
        for (int k = 0; k < (int)std::min<size_t>(np, (size_t)r->n_arenas); k++) {
          swfr_renderer::Arena &A = r->arena[k];
          CK(A.list_items.reserve((size_t)want.list * 4));
          CK(A.row_items.reserve((size_t)want.rows * 8));
          CK(A.stage.reserve((size_t)want.stage * 16));
          CK(A.stage_used.reserve((size_t)(want.stage / 256 + 1) * 4));
          CK(A.edges.reserve((size_t)want.edges * 16));
          CK(A.edge_pid.reserve((size_t)want.edges * 4));
          CK(A.slot_count.reserve(((size_t)want.slots + 1) * 4));
          CK(A.slot_backdrop.reserve(((size_t)want.slots + 1) * 4));
          CK(A.slot_wind.reserve(((size_t)want.slots + 1) * 4));
          CK(A.slot_off.reserve(((size_t)want.slots + 1) * 4));
          CK(A.records.reserve((size_t)want.records * 8));
        }
        r->caps = want;
        retries++;
      } else if (guard > 0) {
        break;  // this run fitted
      }
      // an aborted render leaves the self-cleaning arrays dirty
      for (int k = 0; k < r->n_arenas; k++) {
        swfr_renderer::Arena &A = r->arena[k];
        if (A.slot_count.p) CK(cudaMemsetAsync(A.slot_count.p, 0, A.slot_count.cap, r->stream));
        if (A.slot_backdrop.p) CK(cudaMemsetAsync(A.slot_backdrop.p, 0, A.slot_backdrop.cap, r->stream));
        if (A.dirty.p) CK(cudaMemsetAsync(A.dirty.p, 0, A.dirty.cap, r->stream));
      }
      uint32_t launches = 0;
      int rc = enqueue_passes(r, b, in.slot, true, &launches);
      if (rc != SWFR_OK) {
        r->inflight.clear();
        return rc;
      }
      in.launches += launches;
      CK(cudaStreamSynchronize(r->stream));
    }
    for (const swfr_renderer::CopyReq &q : in.copy_reqs)
      CK(cudaMemcpyAsync(q.dst, (const char *)r->frames.p + (size_t)q.first * fb, (size_t)q.count * fb, cudaMemcpyDeviceToHost, r->stream));
    CK(cudaStreamSynchronize(r->stream));
    r->last_totals.assign(pt, pt + np);
    r->arena_pass = np ? np - 1 : 0;
    int rc = account(r, b, r->last_totals, in.launches, retries);
    if (rc != SWFR_OK) result = rc;
  }
  return result;
}

// Settles the renders in flight, oldest first, until at most `keep` remain.
int settle(swfr_renderer *r, size_t keep) {
  if (r->inflight.size() <= keep) return SWFR_OK;
  Range range("swfr: settle renders");
  int result = SWFR_OK;
  while (r->inflight.size() > keep) {
    swfr_renderer::InFlight &in = r->inflight.front();
    CK(cudaEventSynchronize(r->slot_done[in.slot]));
    const size_t np = in.b->passes.size();
    const Totals *pt = reinterpret_cast<const Totals *>(r->pin_totals[in.slot].p);
    if (any_overflow(pt, np)) {
      int rc = recover(r);
      return rc != SWFR_OK ? rc : result;
    }
    r->last_totals.assign(pt, pt + np);
    int rc = account(r, *in.b, r->last_totals, in.launches, 0);
    if (rc != SWFR_OK) result = rc;
    if (r->prof_passes && r->inflight.size() == 1) {
      for (int k = 0; k < kNumStages; k++) r->stage_ms[k] = 0.f;
      for (size_t i = 0; i < r->prof_passes; i++) {
        cudaEvent_t *ev = r->prof_events.data() + i * (kNumStages + 1);
        for (int k = 0; k < kNumStages; k++) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) r->stage_ms[k] += ms;
        }
      }
      r->stage_launches = (uint32_t)r->prof_passes;
      r->prof_passes = 0;
    }
    r->inflight.pop_front();
  }
  return result;
}

// Waits for every render in flight, grows working memory and re-runs what overflowed it.
int finish(swfr_renderer *r) { return settle(r, 0); }

int register_def(swfr_renderer *r, const swfr_define_shape *tag, bool morph, uint32_t *out_id) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!tag || !out_id) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL argument");
  auto def = std::make_unique<CompiledDef>();
  std::string err;
  int rc = compile_definition(tag, morph, *def, err);
  if (rc != SWFR_OK) return fail(r, rc, err);
  DefEntry de{};
  de.is_morph = morph ? 1 : 0;
  de.paint_first = (uint32_t)r->h_paints.size();
  de.path_count = (uint32_t)def->paints.size();
  de.seg_count = (uint32_t)def->segs.size();
  for (const DefPaint &p : def->paints)
    if (p.flags & PF_SAMPLED) de.has_sampled = 1;
  size_t ramp_base = r->h_ramps.size() / kRampSize;
  for (auto &l : def->luts) r->h_ramps.insert(r->h_ramps.end(), l.begin(), l.end());
  for (DefPaint p : def->paints) {
    if (p.lut >= 0) p.lut += (int32_t)ramp_base;
    r->h_paints.push_back(p);
  }
  if (morph) {
    de.seg_first = (uint32_t)r->h_morph.size();
    r->h_morph.insert(r->h_morph.end(), def->segs.begin(), def->segs.end());
    *out_id = (uint32_t)r->morph_defs.size();
    r->morph_defs.push_back(de);
    swfr_renderer::MorphStroke ms;
    if (def->has_visible_morph_stroke) {
      ms.line_first = (uint32_t)r->h_mlines.size();
      ms.line_count = (uint32_t)def->morph_lines.size();
      morph_lines_to_device(def->morph_lines, r->h_mlines, r->h_mcmds);
      // room reserved per draw: the largest outline at five ratios plus a quarter
      uint32_t most = 0;
      for (int q = 0; q <= 4; q++) {
        uint32_t n_segs = 0;
        count_morph_strokes(r, ms, q / 4.0, nullptr, &n_segs);
        most = std::max(most, n_segs);
      }
      ms.seg_cap = most + most / 4 + 16;
    }
    r->morph_strokes.push_back(ms);
    r->morph_dbg.push_back(r->retain_compiled ? std::move(def) : nullptr);
  } else {
    de.seg_first = (uint32_t)r->h_static.size();
    for (const SegMorph &s : def->segs) {
      SegStatic g;
      memcpy(g.p, s.s, sizeof g.p);
      g.path_flags = s.path_flags;
      r->h_static.push_back(g);
    }
    *out_id = (uint32_t)r->shape_defs.size();
    r->shape_defs.push_back(de);
    r->shape_dbg.push_back(r->retain_compiled ? std::move(def) : nullptr);
  }
  return SWFR_OK;
}

}  // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================

extern "C" {

uint32_t swfr_abi_version(void) { return SWFR_ABI_VERSION; }

const char *swfr_status_string(int s) {
  switch (s) {
    case SWFR_OK: return "ok";
    case SWFR_ERR_INVALID_HANDLE: return "invalid handle";
    case SWFR_ERR_INVALID_ID: return "invalid id";
    case SWFR_ERR_INVALID_FILL_ID: return "invalid fill id";
    case SWFR_ERR_UNSUPPORTED_STYLE: return "unsupported style";
    case SWFR_ERR_OOM: return "out of memory";
    case SWFR_ERR_CUDA: return "CUDA error";
    case SWFR_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SWFR_ERR_MALFORMED: return "malformed input";
    default: return "unknown status";
  }
}

const char *swfr_last_error(const swfr_renderer *r) { return r ? r->last_error.c_str() : "invalid handle"; }

int swfr_create_on_stream(int device, uint32_t width, uint32_t height, void *cuda_stream, swfr_renderer **out) {
  if (!out) return SWFR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (width == 0 || height == 0 || width > 16384 || height > 16384) return SWFR_ERR_INVALID_ARGUMENT;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return SWFR_ERR_CUDA;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return SWFR_ERR_CUDA;
  swfr_renderer *r = new swfr_renderer();
  r->device = device;
  r->width = width;
  r->height = height;
  r->tiles_x = (width + kTile - 1) / kTile;
  r->tiles_y = (height + kTile - 1) / kTile;
  if (cuda_stream) {
    r->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete r;
      return SWFR_ERR_CUDA;
    }
    r->own_stream = true;
  }
  if (const char *e = getenv("SWFR_FRAMES_PER_PASS")) r->frames_per_pass = (uint32_t)std::min(std::max(1, atoi(e)), 4096);
  if (const char *e = getenv("SWFR_ARENAS")) r->n_arenas = std::min(std::max(1, atoi(e)), (int)swfr_renderer::kArenas);
  *out = r;
  return SWFR_OK;
}

int swfr_create(int device, uint32_t width, uint32_t height, swfr_renderer **out) {
  return swfr_create_on_stream(device, width, height, nullptr, out);
}

void swfr_destroy(swfr_renderer *r) {
  if (!r) return;
  cudaSetDevice(r->device);
  cudaStreamSynchronize(r->stream);
  for (BitmapRes &b : r->bitmaps) {
    if (b.tex) cudaDestroyTextureObject(b.tex);
    if (b.arr) cudaFreeArray(b.arr);
  }
  if (r->copy_stream) {
    cudaStreamSynchronize(r->copy_stream);
    cudaStreamDestroy(r->copy_stream);
  }
  for (int k = 0; k < swfr_renderer::kArenas; k++)
    if (r->pass_stream[k]) {
      cudaStreamSynchronize(r->pass_stream[k]);
      cudaStreamDestroy(r->pass_stream[k]);
      cudaEventDestroy(r->join_ev[k]);
    }
  if (r->render_done) cudaEventDestroy(r->render_done);
  for (const swfr_renderer::Import &im : r->imports) cudaIpcCloseMemHandle(im.base);
  for (int k = 0; k < 2; k++)
    if (r->gather_ev[k]) cudaEventDestroy(r->gather_ev[k]);
  for (int k = 0; k < swfr_renderer::kSlots; k++)
    if (r->slot_done[k]) cudaEventDestroy(r->slot_done[k]);
  if (r->up_stream) {
    cudaStreamSynchronize(r->up_stream);
    cudaStreamDestroy(r->up_stream);
  }
  for (const swfr_renderer::CopyFence &cf : r->copy_fences) cudaEventDestroy(cf.done);
  for (cudaEvent_t e : r->fence_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : r->pass_done) cudaEventDestroy(e);
  for (cudaEvent_t e : r->prof_events) cudaEventDestroy(e);
  if (r->own_stream) cudaStreamDestroy(r->stream);
  delete r;
}

int swfr_set_option(swfr_renderer *r, uint32_t key, uint64_t value) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  switch (key) {
    case 1: r->retain_compiled = value != 0; return SWFR_OK;
    case 2: r->frames_per_pass = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, value), 4096); return SWFR_OK;  // grid.y = frames
    case 3: r->profile = value != 0; return SWFR_OK;
    case 4: r->host_threads = (uint32_t)std::min<uint64_t>(value, 256); return SWFR_OK;
    case 5: r->clear_to_background = value != 0; return SWFR_OK;
    case 7:
      if (value > (uint64_t)kMaxChunks) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "at most 8 occlusion chunks");
      r->occlusion_chunks = (uint32_t)value;
      return SWFR_OK;
    case 6:
      r->tiny_arena = value != 0;
      if (r->tiny_arena) r->caps = Caps{0, 0, 0, 0, 0, 0};
      return SWFR_OK;
    default: return fail(r, SWFR_ERR_INVALID_ARGUMENT, "unknown option");
  }
}

int swfr_register_shape(swfr_renderer *r, const swfr_define_shape *tag, uint32_t *out_id) {
  return guarded(r, [&] { return register_def(r, tag, false, out_id); });
}
int swfr_register_morph_shape(swfr_renderer *r, const swfr_define_shape *tag, uint32_t *out_id) {
  return guarded(r, [&] { return register_def(r, tag, true, out_id); });
}

namespace {

// Installs a tight premultiplied RGBA8 image that already sits in device memory (r->scratch) as bitmap `id`:
// 2D array + point-sampling texture object + table entry.
int install_bitmap(swfr_renderer *r, uint16_t id, uint32_t w, uint32_t h, bool opaque) {
  BitmapRes &res = r->bitmaps[id];
  if (res.tex) cudaDestroyTextureObject(res.tex);
  if (res.arr) cudaFreeArray(res.arr);
  res = BitmapRes{};
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
  CK(cudaMallocArray(&res.arr, &cd, w, h));
  CK(cudaMemcpy2DToArrayAsync(res.arr, 0, 0, r->scratch.p, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, r->stream));
  cudaResourceDesc rd{};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = res.arr;
  cudaTextureDesc td{};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;  // filtering is done explicitly (box footprint), texels fetched exactly
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  CK(cudaCreateTextureObject(&res.tex, &rd, &td, nullptr));
  BitmapDev bd{};
  bd.tex = (unsigned long long)res.tex;
  bd.w = (int32_t)w;
  bd.h = (int32_t)h;
  bd.opaque = opaque ? 1 : 0;
  bd.valid = 1;
  r->h_bitmaps[id] = bd;
  CK(cudaMemcpyAsync(r->d_bitmaps.as<BitmapDev>() + id, &r->h_bitmaps[id], sizeof(BitmapDev), cudaMemcpyHostToDevice,
                     r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return SWFR_OK;
}

int begin_bitmap(swfr_renderer *r) {
  cudaSetDevice(r->device);
  int rc = finish(r);
  if (rc != SWFR_OK) return rc;
  rc = flush_store(r);
  if (rc != SWFR_OK) return rc;
  if (r->bitmaps.empty()) {
    r->bitmaps.resize(65536);
    r->h_bitmaps.resize(65536);
  }
  return SWFR_OK;
}

}  // namespace

// Straight RGBA8 from the host: uploaded as it is, premultiplied on the device (what a Canvas stores).
int swfr_register_bitmap(swfr_renderer *r, uint16_t id, uint32_t w, uint32_t h, const uint8_t *rgba, size_t stride) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!rgba || w == 0 || h == 0 || stride < (size_t)w * 4) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "bad bitmap");
  int rc = begin_bitmap(r);
  if (rc != SWFR_OK) return rc;
  const size_t src_bytes = (size_t)(h - 1) * stride + (size_t)w * 4, dst_bytes = (size_t)w * h * 4;
  CK(r->scratch.reserve(dst_bytes));
  CK(r->scratch2.reserve(src_bytes + 16));
  uint32_t *flag = reinterpret_cast<uint32_t *>((char *)r->scratch2.p + ((src_bytes + 3) & ~(size_t)3));
  CK(cudaMemcpyAsync(r->scratch2.p, rgba, src_bytes, cudaMemcpyHostToDevice, r->stream));
  CK(cudaMemsetAsync(flag, 0, 4, r->stream));
  launch_premultiply(r->scratch2.as<uint8_t>(), stride, w, h, r->scratch.as<uint32_t>(), flag, r->stream);
  uint32_t translucent = 0;
  CK(cudaMemcpyAsync(&translucent, flag, 4, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return install_bitmap(r, id, w, h, translucent == 0);
}

// DefineBitmap image/x-swf-bmp: zlib inflate on the host (a serial bit stream), colour-table expansion on the device
// (one byte per pixel crosses PCIe instead of four).
int swfr_register_bitmap_xswfbmp(swfr_renderer *r, uint16_t id, const uint8_t *data, size_t len) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!data) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL data");
  return guarded(r, [&]() -> int {
  std::vector<uint8_t> inflated;
  uint32_t w = 0, h = 0, colors = 0, padded = 0;
  std::string err;
  int rc = inflate_xswfbmp(data, len, inflated, &w, &h, &colors, &padded, err);
  if (rc != SWFR_OK) return fail(r, rc, err);
  if (w == 0 || h == 0) return fail(r, SWFR_ERR_MALFORMED, "x-swf-bmp: empty image");
  rc = begin_bitmap(r);
  if (rc != SWFR_OK) return rc;
  CK(r->scratch.reserve((size_t)w * h * 4));
  CK(r->scratch2.reserve(inflated.size()));
  CK(cudaMemcpyAsync(r->scratch2.p, inflated.data(), inflated.size(), cudaMemcpyHostToDevice, r->stream));
  launch_xswfbmp_expand(r->scratch2.as<uint8_t>(), colors, w, h, padded, r->scratch.as<uint32_t>(), r->stream);
  CK(cudaStreamSynchronize(r->stream));  // `inflated` is pageable host memory
  return install_bitmap(r, id, w, h, true);
  });
}

int swfr_decode_xswfbmp(const uint8_t *data, size_t len, uint8_t *rgba, uint64_t cap, uint32_t *w, uint32_t *h) {
  if (!data) return SWFR_ERR_INVALID_ARGUMENT;
  return guarded(nullptr, [&]() -> int {
  std::vector<uint8_t> out;
  std::string err;
  uint32_t ww = 0, hh = 0;
  int rc = decode_xswfbmp(data, len, out, &ww, &hh, err);
  if (rc != SWFR_OK) return rc;
  if (w) *w = ww;
  if (h) *h = hh;
  if (rgba && cap < out.size()) return SWFR_ERR_INVALID_ARGUMENT;  // the caller's buffer is too small for w x h x 4 bytes
  if (rgba && !out.empty()) memcpy(rgba, out.data(), out.size());
  return SWFR_OK;
  });
}

int swfr_render_batch(swfr_renderer *r, const swfr_stage *stages, uint32_t n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!stages || n == 0) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no stages");
  cudaSetDevice(r->device);
  // Earlier renders (up to two) are still on the GPU: flatten and upload this one's stages meanwhile, into the
  // scratch batch none of them uses, and only then let the host catch up with all but the newest of them.
  r->scratch_ix = (r->scratch_ix + 1) % swfr_renderer::kSlots;
  swfr_batch &b = r->scratch_batch[r->scratch_ix];
  return guarded(r, [&]() -> int {
    for (;;) {  // only when renders of caller-owned batches were interleaved: settle up to the last user of this scratch
      size_t at = r->inflight.size();
      for (size_t i = 0; i < r->inflight.size(); i++)
        if (r->inflight[i].b == &b) at = i;
      if (at == r->inflight.size()) break;
      int rc0 = settle(r, r->inflight.size() - 1 - at);
      if (rc0 != SWFR_OK) return rc0;
    }
    if (r->last == &b) {  // the taps and frames_rendered describe a batch that is being rebuilt
      r->last = nullptr;
      r->frames_rendered = 0;
    }
    int rc = build_batch(r, stages, n, b);
    if (rc != SWFR_OK) return rc;
    rc = upload_batch(r, b);
    if (rc != SWFR_OK) return rc;
    return launch_batch(r, b);
  });
}

int swfr_render(swfr_renderer *r, const swfr_stage *stage) { return swfr_render_batch(r, stage, 1); }

int swfr_render_display_stages(swfr_renderer *r, const swfr_display_stage *stages, uint32_t n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!stages || n == 0) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no stages");
  return guarded(r, [&]() -> int {
  std::vector<std::vector<swfr_display_primitive>> prims(n);
  std::vector<swfr_stage> flat(n);
  for (uint32_t f = 0; f < n; f++) {
    const swfr_display_stage &ds = stages[f];
    if (ds.width != r->width || ds.height != r->height)
      return fail(r, SWFR_ERR_INVALID_ARGUMENT, "stage size differs from the renderer's viewport");
    uint32_t count = 0;
    int rc = swfr_flatten_display_stage(&ds, nullptr, 0, &count);
    if (rc != SWFR_OK) return fail(r, rc, "UnexpectedDisplayObjectType");
    prims[f].resize(std::max<uint32_t>(count, 1));
    rc = swfr_flatten_display_stage(&ds, prims[f].data(), count, &count);
    if (rc != SWFR_OK) return fail(r, rc, "UnexpectedDisplayObjectType");
    flat[f].background_color = ds.has_background_color ? ds.background_color : swfr_rgba8{0, 0, 0, 0};
    flat[f].n_primitives = count;
    flat[f].display_root = prims[f].data();
  }
  return swfr_render_batch(r, flat.data(), n);
  });
}

int swfr_render_display_stage(swfr_renderer *r, const swfr_display_stage *stage) {
  return swfr_render_display_stages(r, stage, 1);
}

int swfr_batch_create(swfr_renderer *r, const swfr_stage *stages, uint32_t n, swfr_batch **out) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!stages || n == 0 || !out) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no stages");
  cudaSetDevice(r->device);
  return guarded(r, [&]() -> int {
    auto b = std::make_unique<swfr_batch>();
    int rc = build_batch(r, stages, n, *b);
    if (rc != SWFR_OK) return rc;
    rc = upload_batch(r, *b);
    if (rc != SWFR_OK) return rc;
    CK(cudaStreamSynchronize(r->up_stream));
    *out = b.release();
    return SWFR_OK;
  });
}

int swfr_batch_render(swfr_renderer *r, swfr_batch *b) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!b) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL batch");
  cudaSetDevice(r->device);
  return launch_batch(r, *b);
}

void swfr_batch_destroy(swfr_renderer *r, swfr_batch *b) {
  if (!b) return;
  if (r) {
    cudaSetDevice(r->device);
    finish(r);
    if (r->last == b) r->last = nullptr;
  }
  delete b;
}

int swfr_sync(swfr_renderer *r) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  cudaSetDevice(r->device);
  int rc = finish(r);
  if (rc != SWFR_OK) return rc;
  CK(cudaStreamSynchronize(r->stream));
  if (r->copy_pending) {
    CK(cudaStreamSynchronize(r->copy_stream));
    r->copy_pending = false;
  }
  if (r->gather_timed) {
    r->gather_timed = false;
    if (cudaEventElapsedTime(&r->gather_ms, r->gather_ev[0], r->gather_ev[1]) != cudaSuccess) r->gather_ms = 0.f;
  }
  return SWFR_OK;
}

int swfr_get_stage_times(swfr_renderer *r, float *ms, uint32_t cap, uint32_t *n_stages, uint32_t *n_passes) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  int rc = swfr_sync(r);
  if (rc != SWFR_OK) return rc;
  for (uint32_t k = 0; k < (uint32_t)kNumStages && ms && k < cap; k++) ms[k] = r->stage_ms[k];
  if (n_stages) *n_stages = kNumStages;
  if (n_passes) *n_passes = r->stage_launches;
  return SWFR_OK;
}

const char *swfr_stage_name(uint32_t i) { return stage_name((int)i); }

int swfr_get_stats(swfr_renderer *r, swfr_stats *out) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!out) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL out");
  int rc = swfr_sync(r);
  *out = r->stats;
  return rc;
}

int swfr_read_image(swfr_renderer *r, uint32_t frame, uint8_t *dst, size_t stride, int premultiplied) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!dst || stride < (size_t)r->width * 4) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "bad destination");
  if (frame >= r->frames_rendered) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "Failed to render: no such frame");
  cudaSetDevice(r->device);
  int rc = finish(r);
  if (rc != SWFR_OK) return rc;
  size_t npx = (size_t)r->width * r->height;
  const uint32_t *src = r->frames.as<uint32_t>() + (size_t)frame * npx;
  if (!premultiplied) {
    CK(r->scratch.reserve(npx * 4));
    launch_unpremultiply(src, r->scratch.as<uint32_t>(), npx, r->stream);
    src = r->scratch.as<uint32_t>();
  }
  CK(cudaMemcpy2DAsync(dst, stride, src, (size_t)r->width * 4, (size_t)r->width * 4, r->height, cudaMemcpyDeviceToHost,
                       r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return SWFR_OK;
}

int swfr_read_frames_async(swfr_renderer *r, uint32_t first, uint32_t count, uint8_t *dst) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!dst || count == 0 || first >= r->frames_rendered || count > r->frames_rendered - first)
    return fail(r, SWFR_ERR_INVALID_ARGUMENT, "bad frame range");
  cudaSetDevice(r->device);
  size_t fb = (size_t)r->width * r->height * 4;
  if (r->inflight.empty() || !r->last) {
    CK(cudaMemcpyAsync(dst, (const char *)r->frames.p + (size_t)first * fb, (size_t)count * fb, cudaMemcpyDeviceToHost,
                       r->stream));
    return SWFR_OK;
  }
  // the render is still in flight: copy each pass' frames as soon as that pass is done, on the copy stream, so the
  // transfer overlaps the passes that follow
  if (!r->copy_stream) CK(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
  const swfr_batch &b = *r->last;
  for (size_t i = 0; i < b.passes.size(); i++) {
    const Pass &p = b.passes[i];
    const uint32_t fs = fine_slice_frames(p.n_frames), ns = fine_slices(p.n_frames);
    for (uint32_t k = 0; k < ns; k++) {  // every slice of the pass' frames leaves as soon as it is final
      const uint32_t s0 = p.f0 + k * fs, s1 = std::min(p.f0 + p.n_frames, s0 + fs);
      uint32_t lo = std::max(first, s0), hi = std::min(first + count, s1);
      if (lo >= hi) continue;
      CK(cudaStreamWaitEvent(r->copy_stream, r->pass_done[i * kMaxFineSlices + k], 0));
      CK(cudaMemcpyAsync(dst + (size_t)(lo - first) * fb, (const char *)r->frames.p + (size_t)lo * fb, (size_t)(hi - lo) * fb,
                         cudaMemcpyDeviceToHost, r->copy_stream));
      cudaEvent_t done;
      if (!r->fence_pool.empty()) {
        done = r->fence_pool.back();
        r->fence_pool.pop_back();
      } else {
        CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      }
      CK(cudaEventRecord(done, r->copy_stream));
      r->copy_fences.push_back(swfr_renderer::CopyFence{lo, hi - lo, done});
    }
  }
  r->copy_pending = true;
  r->inflight.back().copy_reqs.push_back(swfr_renderer::CopyReq{first, count, dst});
  return SWFR_OK;
}

int swfr_device_frames(swfr_renderer *r, void **out_ptr, uint32_t *out_n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (out_ptr) *out_ptr = r->frames.p;
  if (out_n) *out_n = r->frames_rendered;
  return SWFR_OK;
}

// ---- optional peer gather of finished frames -----------------------------------------------

int swfr_export_frames(swfr_renderer *r, swfr_frames_export *out) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!out) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL out");
  cudaSetDevice(r->device);
  int rc = swfr_sync(r);
  if (rc != SWFR_OK) return rc;
  if (!r->frames.p || r->frames_rendered == 0) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no rendered frames to export");
  memset(out, 0, sizeof *out);
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, r->frames.p));
  static_assert(sizeof(h) == sizeof(out->ipc_handle), "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out->ipc_handle, &h, sizeof h);
  out->pid = (uint64_t)getpid();
  out->device_ptr = (uint64_t)(uintptr_t)r->frames.p;
  out->offset = 0;  // cudaMalloc'd on its own: the handle opens at frame 0
  out->frame_bytes = (uint64_t)r->width * r->height * 4;
  out->device = r->device;
  out->n_frames = r->frames_rendered;
  out->width = r->width;
  out->height = r->height;
  return SWFR_OK;
}

int swfr_gather_frames(swfr_renderer *r, const swfr_frames_export *src, uint32_t n_src, void **out_ptr, uint32_t *out_n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!src || n_src == 0 || !out_ptr || !out_n) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL argument");
  cudaSetDevice(r->device);
  return guarded(r, [&]() -> int {
    const uint64_t fb = (uint64_t)r->width * r->height * 4;
    uint64_t total = 0;
    uint32_t slots = 0;
    for (uint32_t k = 0; k < n_src; k++) {
      if (src[k].width != r->width || src[k].height != r->height || src[k].frame_bytes != fb)
        return fail(r, SWFR_ERR_INVALID_ARGUMENT, "a source renders another viewport size");
      // frame f lives on source f mod N in slot f div N: source k holds ceil((total - k) / N) frames
      slots = std::max(slots, src[k].n_frames);
      total += src[k].n_frames;
    }
    for (uint32_t k = 0; k < n_src; k++) {
      const uint64_t want = (total + n_src - 1 - k) / n_src;
      if (src[k].n_frames != want)
        return fail(r, SWFR_ERR_INVALID_ARGUMENT, "source " + std::to_string(k) + " holds " + std::to_string(src[k].n_frames) +
                                                      " frames, round-robin sharding of " + std::to_string(total) + " expects " +
                                                      std::to_string(want));
    }
    if (total > 0xffffffffull) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "too many frames");
    CK(r->gathered.reserve(std::max<size_t>((size_t)slots * n_src * fb, 256)));
    if (!r->copy_stream) CK(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++)
      if (!r->gather_ev[k]) CK(cudaEventCreate(&r->gather_ev[k]));
    std::vector<const void *> ptrs(n_src);
    for (uint32_t k = 0; k < n_src; k++) {
      if (src[k].pid == (uint64_t)getpid()) {  // same process (one thread per GPU): the pointer itself
        ptrs[k] = (const void *)(uintptr_t)src[k].device_ptr;
        if (src[k].device != r->device) {
          int can = 0;
          CK(cudaDeviceCanAccessPeer(&can, r->device, src[k].device));
          if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(src[k].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
          }  // without peer access the copy is staged through the host by the driver
        }
      } else {  // another process: open its allocation (once)
        void *base = nullptr;
        for (const swfr_renderer::Import &im : r->imports)
          if (memcmp(im.handle, src[k].ipc_handle, 64) == 0) base = im.base;
        if (!base) {
          cudaIpcMemHandle_t h;
          memcpy(&h, src[k].ipc_handle, sizeof h);
          CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
          swfr_renderer::Import im;
          memcpy(im.handle, src[k].ipc_handle, 64);
          im.base = base;
          r->imports.push_back(im);
        }
        ptrs[k] = (const char *)base + src[k].offset;
      }
    }
    // the copies: one strided copy per source (row = one frame, destination pitch = N frames), after this renderer's
    // own work (its frame store may be one of the sources)
    if (!r->render_done) CK(cudaEventCreateWithFlags(&r->render_done, cudaEventDisableTiming));
    CK(cudaEventRecord(r->render_done, r->stream));
    CK(cudaStreamWaitEvent(r->copy_stream, r->render_done, 0));
    CK(cudaEventRecord(r->gather_ev[0], r->copy_stream));
    for (uint32_t k = 0; k < n_src; k++) {
      if (src[k].n_frames == 0) continue;
      CK(cudaMemcpy2DAsync((char *)r->gathered.p + (size_t)k * fb, (size_t)n_src * fb, ptrs[k], (size_t)fb, (size_t)fb, src[k].n_frames,
                           cudaMemcpyDefault, r->copy_stream));
    }
    CK(cudaEventRecord(r->gather_ev[1], r->copy_stream));
    r->copy_pending = true;
    r->gather_timed = true;
    *out_ptr = r->gathered.p;
    *out_n = (uint32_t)total;
    return SWFR_OK;
  });
}

int swfr_gather_last_ms(swfr_renderer *r, float *out_ms) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  if (!out_ms) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "NULL out");
  int rc = swfr_sync(r);
  if (rc != SWFR_OK) return rc;
  *out_ms = r->gather_ms;
  return SWFR_OK;
}

// ---- parity taps ------------------------------------------------------------------------------------

int swfr_debug_compiled(swfr_renderer *r, uint32_t kind, uint32_t id, double *commands, uint64_t commands_cap,
                        uint64_t *n_commands, int32_t *path_info, uint64_t path_cap, uint64_t *n_paths) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  auto &v = kind == SWFR_PRIM_MORPH_SHAPE ? r->morph_dbg : r->shape_dbg;
  if (id >= v.size() || !v[id]) return fail(r, SWFR_ERR_INVALID_ID, "definition unknown or not retained");
  const CompiledDef &d = *v[id];
  uint64_t nc = 0;
  for (size_t p = 0; p < d.paths.size(); p++) {
    const CompiledPath &cp = d.paths[p];
    if (path_info && p < path_cap) {
      path_info[3 * p] = (int32_t)cp.commands.size();
      path_info[3 * p + 1] = cp.has_fill;
      path_info[3 * p + 2] = cp.has_line;
    }
    for (const Command &c : cp.commands) {
      if (commands && nc < commands_cap) {
        double *o = commands + 9 * nc;
        o[0] = c.type;
        for (int k = 0; k < 4; k++) o[1 + k] = c.s[k], o[5 + k] = c.e[k];
      }
      nc++;
    }
  }
  if (n_commands) *n_commands = nc;
  if (n_paths) *n_paths = d.paths.size();
  return SWFR_OK;
}

// Host-only: compiles a tag without a renderer (no CUDA needed) and returns commands, path info and segments.
int swfr_compile_debug(const swfr_define_shape *tag, int morph, double *commands, uint64_t commands_cap,
                       uint64_t *n_commands, int32_t *path_info, uint64_t path_cap, uint64_t *n_paths, double *segs,
                       uint64_t segs_cap, uint64_t *n_segs) {
  if (!tag) return SWFR_ERR_INVALID_ARGUMENT;
  CompiledDef d;
  std::string err;
  int rc = compile_definition(tag, morph != 0, d, err);
  if (rc != SWFR_OK) return rc;
  uint64_t nc = 0;
  for (size_t p = 0; p < d.paths.size(); p++) {
    const CompiledPath &cp = d.paths[p];
    if (path_info && p < path_cap) {
      path_info[3 * p] = (int32_t)cp.commands.size();
      path_info[3 * p + 1] = cp.has_fill;
      path_info[3 * p + 2] = cp.has_line;
    }
    for (const Command &c : cp.commands) {
      if (commands && nc < commands_cap) {
        double *o = commands + 9 * nc;
        o[0] = c.type;
        for (int k = 0; k < 4; k++) o[1 + k] = c.s[k], o[5 + k] = c.e[k];
      }
      nc++;
    }
  }
  if (n_commands) *n_commands = nc;
  if (n_paths) *n_paths = d.paths.size();
  for (size_t i = 0; i < d.segs.size() && segs && i < segs_cap; i++) {
    double *o = segs + 14 * i;
    o[0] = d.segs[i].path_flags >> 31;
    o[1] = d.segs[i].path_flags & 0x7fffffffu;
    for (int k = 0; k < 6; k++) o[2 + k] = d.segs[i].s[k], o[8 + k] = d.segs[i].e[k];
  }
  if (n_segs) *n_segs = d.segs.size();
  return SWFR_OK;
}

int swfr_debug_segments(swfr_renderer *r, uint32_t kind, uint32_t id, double *segs, uint64_t cap, uint64_t *n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  auto &v = kind == SWFR_PRIM_MORPH_SHAPE ? r->morph_dbg : r->shape_dbg;
  if (id >= v.size() || !v[id]) return fail(r, SWFR_ERR_INVALID_ID, "definition unknown or not retained");
  const CompiledDef &d = *v[id];
  for (size_t i = 0; i < d.segs.size() && segs && i < cap; i++) {
    double *o = segs + 14 * i;
    o[0] = d.segs[i].path_flags >> 31;
    o[1] = d.segs[i].path_flags & 0x7fffffffu;
    for (int k = 0; k < 6; k++) o[2 + k] = d.segs[i].s[k], o[8 + k] = d.segs[i].e[k];
  }
  if (n) *n = d.segs.size();
  return SWFR_OK;
}

// Host run of the device stroker's generator (stroke_core.h) for one morph-shape tag at one ratio: what k_stroke writes
// for a draw.  Host only (no renderer, no CUDA).  out: 8 doubles per segment (curve, path, x0, y0, cx, cy, x1, y1).
int swfr_debug_morph_stroke(const swfr_define_shape *tag, double ratio, double *out, uint64_t cap, uint64_t *n_segs,
                            uint32_t *n_paths) {
  if (!tag) return SWFR_ERR_INVALID_ARGUMENT;
  return guarded(nullptr, [&]() -> int {
    CompiledDef def;
    std::string err;
    int rc = compile_definition(tag, true, def, err);
    if (rc != SWFR_OK) return rc;
    std::vector<stroke::LineDev> lines;
    std::vector<stroke::CmdDev> cmds;
    morph_lines_to_device(def.morph_lines, lines, cmds);
    std::vector<SegStatic> segs;
    double width_state = 1.0;
    uint32_t path = 0;
    for (const stroke::LineDev &ln : lines) {
      const double w = stroke::lerp(ln.w0, ln.w1, ratio);
      if (w > 0) width_state = w;
      if (stroke::lerp(ln.color0[3] / 255.0, ln.color1[3] / 255.0, ratio) <= 0) continue;
      stroke::Sink count{nullptr, 0, 0, path, {1.f, 1.f, 0.f, 0.f}, false, 0.f, 0.f, 0.f, 0.f};
      stroke::stroke_line(cmds.data() + ln.cmd_first, ln.cmd_count, ratio, width_state, count);
      const size_t at = segs.size();
      segs.resize(at + count.n);
      stroke::Sink sink{segs.data() + at, count.n, 0, path, {1.f, 1.f, 0.f, 0.f}, false, 0.f, 0.f, 0.f, 0.f};
      stroke::stroke_line(cmds.data() + ln.cmd_first, ln.cmd_count, ratio, width_state, sink);
      path++;
    }
    for (size_t i = 0; i < segs.size() && out && i < cap; i++) {
      double *o = out + 8 * i;
      o[0] = segs[i].path_flags >> 31;
      o[1] = segs[i].path_flags & 0x7fffffffu;
      for (int k = 0; k < 6; k++) o[2 + k] = segs[i].p[k];
    }
    if (n_segs) *n_segs = segs.size();
    if (n_paths) *n_paths = path;
    return SWFR_OK;
  });
}

static int debug_pass(swfr_renderer *r, uint32_t frame, const Pass **pass, size_t *index) {
  if (!r->last || frame >= r->frames_rendered) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no such frame");
  int rc = finish(r);
  if (rc != SWFR_OK) return rc;
  if (!r->last) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no render to inspect");
  const swfr_batch &b = *r->last;
  // the arena holds the working set of one pass only (the last one launched)
  if (r->arena_pass >= b.passes.size()) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "no render to inspect");
  const Pass &p = b.passes[r->arena_pass];
  if (frame < p.f0 || frame >= p.f0 + p.n_frames)
    return fail(r, SWFR_ERR_INVALID_ARGUMENT, "debug taps need the frame to be in the last pass (render fewer frames)");
  if (p.n_chunks > 1)
    return fail(r, SWFR_ERR_INVALID_ARGUMENT,
                "debug taps need the complete edge / record lists: set SWFR_OPT_OCCLUSION_CHUNKS to 1 (hidden geometry is "
                "skipped otherwise)");
  *pass = &p;
  *index = r->arena_pass;
  return SWFR_OK;
}

int swfr_debug_edges(swfr_renderer *r, uint32_t frame, int32_t *edges, int32_t *edge_path, uint64_t cap, uint64_t *n) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  cudaSetDevice(r->device);
  const Pass *p;
  size_t pi;
  int rc = debug_pass(r, frame, &p, &pi);
  if (rc != SWFR_OK) return rc;
  const swfr_renderer::Arena &A = r->arena[pi % (size_t)r->n_arenas];
  const swfr_batch &b = *r->last;
  // frame -> path range -> item range -> segment-instance range -> edge range
  uint32_t lf = frame - p->f0;
  uint32_t path_lo = b.frame_off[p->frame_off_at + lf], path_hi = b.frame_off[p->frame_off_at + lf + 1];
  const uint32_t *po = &b.path_off[p->path_off_at], *so = &b.seg_off[p->seg_off_at];
  uint32_t i_lo = 0, i_hi = p->n_items;
  // items of a frame are contiguous; find them by their frame tag
  while (i_lo < p->n_items && b.items[p->items_at + i_lo].frame < lf) i_lo++;
  i_hi = i_lo;
  while (i_hi < p->n_items && b.items[p->items_at + i_hi].frame == lf) i_hi++;
  (void)po;
  (void)path_hi;
  uint32_t s_lo = so[i_lo], s_hi = so[i_hi];
  uint32_t e_lo = 0, e_hi = 0;
  CK(cudaMemcpy(&e_lo, A.seg_edge_off.as<uint32_t>() + s_lo, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&e_hi, A.seg_edge_off.as<uint32_t>() + s_hi, 4, cudaMemcpyDeviceToHost));
  uint64_t cnt = e_hi - e_lo;
  if (n) *n = cnt;
  uint64_t take = std::min<uint64_t>(cnt, cap);
  if (edges && take) CK(cudaMemcpy(edges, A.edges.as<int4>() + e_lo, take * 16, cudaMemcpyDeviceToHost));
  if (edge_path && take) {
    CK(cudaMemcpy(edge_path, A.edge_pid.as<uint32_t>() + e_lo, take * 4, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < take; i++) edge_path[i] -= (int32_t)path_lo;  // index within the frame
  }
  return SWFR_OK;
}

int swfr_debug_tile_counts(swfr_renderer *r, uint32_t frame, uint32_t *counts, uint64_t cap) {
  if (!r) return SWFR_ERR_INVALID_HANDLE;
  cudaSetDevice(r->device);
  const Pass *p;
  size_t pi;
  int rc = debug_pass(r, frame, &p, &pi);
  if (rc != SWFR_OK) return rc;
  size_t nt = (size_t)r->tiles_x * r->tiles_y;
  if (!counts || cap < nt) return fail(r, SWFR_ERR_INVALID_ARGUMENT, "counts buffer too small");
  CK(r->scratch.reserve(std::max<size_t>(nt * 4, (size_t)r->width * r->height * 4)));
  CK(cudaMemsetAsync(r->scratch.p, 0, nt * 4, r->stream));
  RenderArgs a = make_args(r, *r->last, *p, pi, 0);
  launch_tile_counts(a, frame - p->f0, r->scratch.as<uint32_t>(), r->stream);
  CK(cudaMemcpyAsync(counts, r->scratch.p, nt * 4, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return SWFR_OK;
}

}  // extern "C"
