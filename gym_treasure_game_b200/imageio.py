"""Image files without third-party packages (``_TreasureGameDrawer.draw_to_file``,
``_treasure_game_drawer.py:233-236``, uses ``pygame.image.save``): PNG through zlib, 24-bit BMP."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def _png(rgb: np.ndarray) -> bytes:
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))          # filter type 0 per scanline

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def _bmp(rgb: np.ndarray) -> bytes:
    h, w, _ = rgb.shape
    pad = (-3 * w) % 4
    rows = b"".join(rgb[y, :, ::-1].tobytes() + b"\x00" * pad for y in range(h - 1, -1, -1))   # BGR, bottom-up
    header = struct.pack("<2sIHHI", b"BM", 54 + len(rows), 0, 0, 54)
    info = struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(rows), 2835, 2835, 0, 0)
    return header + info + rows


def save_rgb(fname: str, rgb: np.ndarray) -> None:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    if rgb.ndim != 3 or rgb.shape[2] != 3:
        raise ValueError("expected an (H, W, 3) uint8 image")
    data = _bmp(rgb) if fname.lower().endswith(".bmp") else _png(rgb)
    with open(fname, "wb") as f:
        f.write(data)
