"""Oracle restatement of ts/src/lib/decode-x-swf-bmp.ts:9-41 (TEST INFRASTRUCTURE).

``image/x-swf-bmp`` format 3: u8 format id, u16le width, u16le height, u8 (colour count - 1),
then a zlib stream holding an RGB palette followed by palette indices, rows padded to 4 bytes.
Out-of-range indices decode to opaque black (decode-x-swf-bmp.ts:35-36).  Alpha is always 255.
Pinned byte-exact against tests/bitmap/homestuck-beta-3.pam (ts/src/test/decode-bitmap.spec.ts:31-36).
"""
import zlib

import numpy as np


def decode_x_swf_bmp(data: bytes) -> np.ndarray:
    """Return straight RGBA8 as an (h, w, 4) uint8 array."""
    if data[0] != 3:
        raise ValueError("UnsupportedXSwfBmpFormatId: %d" % data[0])
    width = data[1] | (data[2] << 8)
    height = data[3] | (data[4] << 8)
    padded = width + ((4 - (width % 4)) % 4)
    color_count = data[5] + 1
    src = zlib.decompress(bytes(data[6:]))
    table = 3 * color_count
    pal = np.zeros((256, 4), dtype=np.uint8)
    pal[:, 3] = 255  # out-of-range index => 0x000000ff
    pal[:color_count, :3] = np.frombuffer(src[:table], dtype=np.uint8).reshape(color_count, 3)
    idx = np.frombuffer(src[table : table + padded * height], dtype=np.uint8).reshape(height, padded)[:, :width]
    return pal[idx]


def define_bitmap_rgba(tag: dict) -> np.ndarray:
    """ts/src/lib/renderers/node-canvas-bitmap-service.ts:14-37 on a ``define-bitmap`` AST dict."""
    if tag["media_type"] != "image/x-swf-bmp":
        raise ValueError("NotImplemented: Support for %s images" % tag["media_type"])
    return decode_x_swf_bmp(bytes.fromhex(tag["data"]))


def to_pam(rgba: np.ndarray) -> bytes:
    """ts/src/lib/image-data-to-pam.ts / rs/src/pam.rs:3-34 (P7 RGB_ALPHA)."""
    h, w = rgba.shape[:2]
    head = "P7\nWIDTH %d\nHEIGHT %d\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" % (w, h)
    return head.encode() + rgba.tobytes()
