"""ctypes binding of include/swfr.h (libswfr_b200.so).

The library is the product: it must be present (built in-tree by ``__graft_entry__.build()`` or
``make -C swf_renderer_b200/csrc``) and there is no fallback of any kind when it is missing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswfr_b200.so")

OK = 0
ERR_INVALID_HANDLE, ERR_INVALID_ID, ERR_INVALID_FILL_ID, ERR_UNSUPPORTED_STYLE = -1, -2, -3, -4
ERR_OOM, ERR_CUDA, ERR_INVALID_ARGUMENT, ERR_MALFORMED = -5, -6, -7, -8

FILL_SOLID, FILL_LINEAR_GRADIENT, FILL_RADIAL_GRADIENT, FILL_FOCAL_GRADIENT, FILL_BITMAP = 0, 1, 2, 3, 4
SPREAD_PAD, SPREAD_REFLECT, SPREAD_REPEAT = 0, 1, 2
COLOR_SRGB, COLOR_LINEAR_RGB = 0, 1
RECORD_EDGE, RECORD_STYLE_CHANGE = 0, 1
PRIM_SHAPE, PRIM_MORPH_SHAPE = 0, 1
PRIM_RATIO_F32 = 1
PRIM_COLOR_TRANSFORM = 2
OPT_RETAIN_COMPILED, OPT_FRAMES_PER_PASS, OPT_PROFILE, OPT_HOST_THREADS, OPT_CLEAR_TO_BACKGROUND = 1, 2, 3, 4, 5
OPT_DEBUG_TINY_ARENA = 6
OPT_OCCLUSION_CHUNKS = 7


class Rgba8(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


class SwfMatrix(C.Structure):
    _fields_ = [
        ("scale_x", C.c_int32),
        ("scale_y", C.c_int32),
        ("rotate_skew0", C.c_int32),
        ("rotate_skew1", C.c_int32),
        ("translate_x", C.c_int32),
        ("translate_y", C.c_int32),
    ]


class ColorStop(C.Structure):
    _fields_ = [("ratio", C.c_uint8), ("color", Rgba8), ("morph_color", Rgba8)]


class Gradient(C.Structure):
    _fields_ = [
        ("spread", C.c_uint8),
        ("color_space", C.c_uint8),
        ("n_colors", C.c_uint16),
        ("colors", C.POINTER(ColorStop)),
    ]


class FillStyle(C.Structure):
    _fields_ = [
        ("type", C.c_uint32),
        ("color", Rgba8),
        ("morph_color", Rgba8),
        ("matrix", SwfMatrix),
        ("gradient", Gradient),
        ("focal_point", C.c_int16),
        ("bitmap_id", C.c_uint16),
        ("repeating", C.c_uint8),
        ("smoothed", C.c_uint8),
    ]


class LineStyle(C.Structure):
    _fields_ = [("width", C.c_uint16), ("morph_width", C.c_uint16), ("fill", FillStyle)]


class Styles(C.Structure):
    _fields_ = [
        ("n_fill", C.c_uint32),
        ("fill", C.POINTER(FillStyle)),
        ("n_line", C.c_uint32),
        ("line", C.POINTER(LineStyle)),
    ]


class ShapeRecord(C.Structure):
    _fields_ = [
        ("type", C.c_uint32),
        ("delta_x", C.c_int32),
        ("delta_y", C.c_int32),
        ("control_delta_x", C.c_int32),
        ("control_delta_y", C.c_int32),
        ("morph_delta_x", C.c_int32),
        ("morph_delta_y", C.c_int32),
        ("morph_control_delta_x", C.c_int32),
        ("morph_control_delta_y", C.c_int32),
        ("has_control_delta", C.c_uint8),
        ("has_morph_control_delta", C.c_uint8),
        ("has_move_to", C.c_uint8),
        ("has_morph_move_to", C.c_uint8),
        ("has_left_fill", C.c_uint8),
        ("has_right_fill", C.c_uint8),
        ("has_line_style", C.c_uint8),
        ("has_new_styles", C.c_uint8),
        ("move_to_x", C.c_int32),
        ("move_to_y", C.c_int32),
        ("morph_move_to_x", C.c_int32),
        ("morph_move_to_y", C.c_int32),
        ("left_fill", C.c_uint32),
        ("right_fill", C.c_uint32),
        ("line_style", C.c_uint32),
        ("new_styles", C.POINTER(Styles)),
    ]


class DefineShape(C.Structure):
    _fields_ = [
        ("id", C.c_uint16),
        ("bounds", C.c_int32 * 4),
        ("morph_bounds", C.c_int32 * 4),
        ("initial_styles", Styles),
        ("n_records", C.c_uint32),
        ("records", C.POINTER(ShapeRecord)),
    ]


class ColorTransform(C.Structure):  # swfr_color_transform: mult in Sfixed8P8 epsilons (256 = 1.0), add in integers
    _fields_ = [
        ("red_mult", C.c_int16),
        ("green_mult", C.c_int16),
        ("blue_mult", C.c_int16),
        ("alpha_mult", C.c_int16),
        ("red_add", C.c_int16),
        ("green_add", C.c_int16),
        ("blue_add", C.c_int16),
        ("alpha_add", C.c_int16),
    ]


class DisplayPrimitive(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32),
        ("id", C.c_uint32),
        ("matrix", C.c_float * 6),
        ("ratio", C.c_uint16),
        ("flags", C.c_uint16),
        ("ratio_f", C.c_float),
        ("color_transform", ColorTransform),
    ]


class Stage(C.Structure):
    _fields_ = [
        ("background_color", Rgba8),
        ("n_primitives", C.c_uint32),
        ("display_root", C.POINTER(DisplayPrimitive)),
    ]


class DisplayObject(C.Structure):
    pass


DisplayObject._fields_ = [
    ("type", C.c_uint32),
    ("id", C.c_uint32),
    ("has_matrix", C.c_uint8),
    ("matrix", SwfMatrix),
    ("ratio", C.c_float),
    ("has_color_transform", C.c_uint8),
    ("color_transform", ColorTransform),
    ("n_children", C.c_uint32),
    ("children", C.POINTER(DisplayObject)),
]


class DisplayStage(C.Structure):
    _fields_ = [
        ("has_background_color", C.c_uint8),
        ("background_color", Rgba8),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("n_children", C.c_uint32),
        ("children", C.POINTER(DisplayObject)),
    ]


DISPLAY_CONTAINER, DISPLAY_MORPH_SHAPE, DISPLAY_SHAPE = 0, 1, 2


class Stats(C.Structure):
    _fields_ = [
        ("n_primitives", C.c_uint64),
        ("n_path_instances", C.c_uint64),
        ("n_segments", C.c_uint64),
        ("n_edges", C.c_uint64),
        ("n_slots", C.c_uint64),
        ("n_records", C.c_uint64),
        ("n_tiles", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
        ("fine_slots", C.c_uint64),
        ("fine_records", C.c_uint64),
        ("kernel_launches", C.c_uint32),
        ("retries", C.c_uint32),
    ]


# every symbol include/swfr.h declares, with its prototype
PROTOTYPES = {
    "swfr_abi_version": (C.c_uint32, []),
    "swfr_status_string": (C.c_char_p, [C.c_int]),
    "swfr_last_error": (C.c_char_p, [C.c_void_p]),
    "swfr_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "swfr_create_on_stream": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "swfr_destroy": (None, [C.c_void_p]),
    "swfr_set_option": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64]),
    "swfr_register_shape": (C.c_int, [C.c_void_p, C.POINTER(DefineShape), C.POINTER(C.c_uint32)]),
    "swfr_register_morph_shape": (C.c_int, [C.c_void_p, C.POINTER(DefineShape), C.POINTER(C.c_uint32)]),
    "swfr_register_bitmap": (C.c_int, [C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "swfr_register_bitmap_xswfbmp": (C.c_int, [C.c_void_p, C.c_uint16, C.c_void_p, C.c_size_t]),
    "swfr_decode_xswfbmp": (
        C.c_int,
        [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    ),
    "swfr_render": (C.c_int, [C.c_void_p, C.POINTER(Stage)]),
    "swfr_render_batch": (C.c_int, [C.c_void_p, C.POINTER(Stage), C.c_uint32]),
    "swfr_flatten_display_stage": (
        C.c_int,
        [C.POINTER(DisplayStage), C.POINTER(DisplayPrimitive), C.c_uint32, C.POINTER(C.c_uint32)],
    ),
    "swfr_render_display_stage": (C.c_int, [C.c_void_p, C.POINTER(DisplayStage)]),
    "swfr_render_display_stages": (C.c_int, [C.c_void_p, C.POINTER(DisplayStage), C.c_uint32]),
    "swfr_write_pam": (
        C.c_int,
        [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)],
    ),
    "swfr_write_png": (
        C.c_int,
        [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)],
    ),
    "swfr_batch_create": (C.c_int, [C.c_void_p, C.POINTER(Stage), C.c_uint32, C.POINTER(C.c_void_p)]),
    "swfr_batch_render": (C.c_int, [C.c_void_p, C.c_void_p]),
    "swfr_batch_destroy": (None, [C.c_void_p, C.c_void_p]),
    "swfr_sync": (C.c_int, [C.c_void_p]),
    "swfr_read_image": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int]),
    "swfr_read_frames_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "swfr_device_frames": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "swfr_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "swfr_get_stage_times": (
        C.c_int,
        [C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
    ),
    "swfr_stage_name": (C.c_char_p, [C.c_uint32]),
    "swfr_debug_compiled": (
        C.c_int,
        [
            C.c_void_p,
            C.c_uint32,
            C.c_uint32,
            C.c_void_p,
            C.c_uint64,
            C.POINTER(C.c_uint64),
            C.c_void_p,
            C.c_uint64,
            C.POINTER(C.c_uint64),
        ],
    ),
    "swfr_debug_segments": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "swfr_debug_edges": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "swfr_debug_tile_counts": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
}



class FramesExport(C.Structure):
    """swfr_frames_export: what a renderer hands to the renderer that gathers its frames (plain bytes; other processes
    receive them through any control-plane message)."""

    _fields_ = [
        ("ipc_handle", C.c_uint8 * 64),
        ("pid", C.c_uint64),
        ("device_ptr", C.c_uint64),
        ("offset", C.c_uint64),
        ("frame_bytes", C.c_uint64),
        ("device", C.c_int32),
        ("n_frames", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
    ]


PROTOTYPES["swfr_export_frames"] = (C.c_int, [C.c_void_p, C.POINTER(FramesExport)])
PROTOTYPES["swfr_gather_frames"] = (
    C.c_int,
    [C.c_void_p, C.POINTER(FramesExport), C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)],
)
PROTOTYPES["swfr_gather_last_ms"] = (C.c_int, [C.c_void_p, C.POINTER(C.c_float)])

PROTOTYPES["swfr_debug_morph_stroke"] = (
    C.c_int,
    [C.POINTER(DefineShape), C.c_double, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)],
)

_LIB = None


def load():
    """Load libswfr_b200.so and attach prototypes.  Raises if the library was not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libswfr_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C swf_renderer_b200/csrc`. There is no CPU fallback." % LIB_PATH
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


PROTOTYPES["swfr_compile_debug"] = (
    C.c_int,
    [
        C.POINTER(DefineShape),
        C.c_int,
        C.c_void_p,
        C.c_uint64,
        C.POINTER(C.c_uint64),
        C.c_void_p,
        C.c_uint64,
        C.POINTER(C.c_uint64),
        C.c_void_p,
        C.c_uint64,
        C.POINTER(C.c_uint64),
    ],
)
