"""CPU: host-side logic of the product -- level text front end, render assets, sharding --
against the oracle's independent restatements."""
import numpy as np
import pytest

import py_oracle as po
import ref_harness as rh
import render_oracle as ro
from gym_treasure_game_b200 import Level, sprites
from gym_treasure_game_b200.vector_env import shard_range


def test_default_level_matches_oracle_level():
    a, b = Level.default(), po.default_level()
    assert list(a.tiles) == b.tiles
    assert [tuple(o) for o in a.objects] == [tuple(o) for o in b.objects]
    assert [tuple(t) for t in a.triggers] == [tuple(t) for t in b.triggers]
    assert a.obs_dim == 9 and a.frame_size == (624, 672)
    assert a.state_descriptors() == ["playerx", "playery", "handle1.angle", "handle2.angle", "key.x", "key.y",
                                     "bolt.locked", "goldcoin.x", "goldcoin.y"]


@pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")
def test_reference_three_file_format_and_descriptors():
    import os
    ref = rh.load_reference()
    d = ref.default_dir
    files = [os.path.join(d, f) for f in ("domain.txt", "domain-objects.txt", "domain-interactions.txt")]
    lv = Level.from_reference_files(*files)
    assert lv == Level.default()
    assert lv.state_descriptors() == ref.Impl(*files).get_state_descriptors()      # impl:380-400


def test_mirror_and_round_trip():
    lv = Level.default()
    m = lv.mirrored()
    assert m.mirrored() == lv
    assert Level.from_strings(*m.to_strings()) == m
    mo = po.mirrored_level(po.default_level())
    assert list(m.tiles) == mo.tiles and [tuple(o) for o in m.objects] == [tuple(o) for o in mo.objects]


def test_background_equals_oracle_renderer():
    for lv, lvt in ((Level.default(), po.default_level()), (Level.default().mirrored(), po.mirrored_level(po.default_level()))):
        assert np.array_equal(sprites.compose_background(lv), ro.background(lvt.tiles))


def test_dynamic_atlas_equals_oracle_scaling():
    atlas = sprites.dynamic_atlas()
    names = ["door_closed", "door_open", "key", "gold", "bolt_open", "bolt_locked", "handle_base", "hero"]
    for i, n in enumerate(names):
        assert np.array_equal(atlas[i], ro._scaled(n)), n
    assert np.array_equal(atlas[8], ro._scaled("hero")[:, ::-1])           # pre-flipped hero (drawer.py:160)
    assert atlas.shape == (9, 48, 48, 4) and atlas.dtype == np.uint8


def test_blend_matches_oracle_blit():
    rng = np.random.default_rng(0)
    dst = rng.integers(0, 256, (48, 48, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (48, 48, 4), dtype=np.uint8)
    src[:8, :, 3] = 0
    src[8:16, :, 3] = 255
    a = dst.copy()
    sprites.blend_over(a, src)
    b = dst.copy()
    ro._blit(b, src, 0, 0)
    assert np.array_equal(a, b)


def test_shard_range_partitions():
    for total, world in ((1 << 20, 8), (1000, 3), (7, 8), (4096, 1)):
        ranges = [shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_image_writers_roundtrip(tmp_path):
    """draw_to_file's writers (PNG through zlib, 24-bit BMP) decode back to the same pixels (no GPU, no PIL)."""
    import struct
    import zlib
    import numpy as np
    from gym_treasure_game_b200.imageio import save_rgb
    img = np.random.default_rng(3).integers(0, 256, (13, 7, 3), dtype=np.uint8)     # odd width: BMP row padding
    png, bmp = str(tmp_path / "a.png"), str(tmp_path / "a.bmp")
    save_rgb(png, img); save_rgb(bmp, img)
    data = open(png, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        chunks[tag] = body
        pos += 12 + n
    assert struct.unpack(">IIBBBBB", chunks[b"IHDR"]) == (7, 13, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), dtype=np.uint8).reshape(13, 1 + 21)
    assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(13, 7, 3), img)
    raw = open(bmp, "rb").read()
    off, w, h, bpp = struct.unpack("<I", raw[10:14])[0], *struct.unpack("<ii", raw[18:26]), struct.unpack("<H", raw[28:30])[0]
    assert (w, h, bpp) == (7, 13, 24)
    stride = (3 * w + 3) // 4 * 4
    body = np.frombuffer(raw[off:], dtype=np.uint8).reshape(h, stride)[:, :3 * w].reshape(h, w, 3)
    assert np.array_equal(body[::-1, :, ::-1], img)
    with pytest.raises(ValueError):
        save_rgb(png, img[:, :, 0])
