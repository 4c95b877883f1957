"""swf_renderer_b200 - B200-native (sm_100a) replacement for the shape -> pixels path of open-flash/swf-renderer.

The package is a thin host-side mirror of the reference renderer interface (``renderer.py``) over the C ABI in
``include/swfr.h`` (``capi.py``); all rasterization runs in the hand-written CUDA kernels of ``csrc/``.
"""
from .renderer import (  # noqa: F401
    HeadlessRenderer,
    Image,
    ImageMetadata,
    Matrix2D,
    Stage,
    StoredMorphShape,
    StoredShape,
    SwfrError,
    compile_tag,
    decode_x_swf_bmp,
)
from . import display  # noqa: F401,E402  (mirror of the TypeScript display tree / Renderer interface)
