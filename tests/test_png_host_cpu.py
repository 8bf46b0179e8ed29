"""CPU: the library's host-side PNG decoder (csrc/png_host.cu; `ops.decode_png_rgb8` / `ops.decode_png_gray16`) against PIL, which is what
the reference's loaders call (src/data_utils.py:134-165 `Image.open(path).convert('RGB')`, :167-234 `np.array(Image.open(path))` of 16-bit
depth maps).  Bit-exact for every colour type PIL writes, every zlib level and all five scanline filters; corrupt files fail with a
message instead of returning garbage."""
import io
import struct
import zlib

import numpy as np
import pytest

PIL = pytest.importorskip('PIL.Image')
from tta_depth_completion_b200 import ops


def png_bytes(img, **kw):
    buf = io.BytesIO()
    img.save(buf, format='PNG', **kw)
    return buf.getvalue()


def smooth_rgb(h, w, seed):
    """image-like content (gradients + texture): PIL's adaptive filter heuristic then uses Sub / Up / Average / Paeth rows, not only None"""
    r = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) % 256)], -1).astype(np.int32)
    return np.clip(base + r.randint(-12, 13, size=(h, w, 3)), 0, 255).astype(np.uint8)


def filter_types(data):
    """the scanline filter bytes actually present in a PNG file (parsed independently of the library)"""
    off, idat, hdr = 8, b'', None
    while off < len(data):
        n, t = struct.unpack('>I4s', data[off:off + 8])
        if t == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', data[off + 8:off + 8 + 13])
        if t == b'IDAT':
            idat += data[off + 8:off + 8 + n]
        off += 12 + n
    w, h, depth, color = hdr[:4]
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color]
    stride = w * ch * depth // 8
    raw = zlib.decompress(idat)
    return {raw[(stride + 1) * y] for y in range(h)}


@pytest.mark.parametrize('h,w', [(1, 1), (7, 13), (64, 96), (352, 1216)])
@pytest.mark.parametrize('level', [0, 1, 6, 9])
def test_rgb8_matches_pil(h, w, level):
    arr = smooth_rgb(h, w, h * 31 + w + level)
    data = png_bytes(PIL.fromarray(arr, 'RGB'), compress_level=level)
    want = np.asarray(PIL.open(io.BytesIO(data)).convert('RGB'))
    got = ops.decode_png_rgb8(data)
    assert got.dtype == np.uint8 and got.shape == (h, w, 3)
    assert np.array_equal(got, want) and np.array_equal(got, arr)
    assert ops.png_info(data) == (w, h, 3, 8)


def test_all_five_filters_are_exercised():
    seen = set()
    for seed in range(6):
        arr = smooth_rgb(96, 160, seed)
        if seed % 2:
            arr = np.random.RandomState(seed).randint(0, 256, size=arr.shape).astype(np.uint8)      # noise: 'None' rows
        data = png_bytes(PIL.fromarray(arr, 'RGB'), compress_level=6)
        seen |= filter_types(data)
        assert np.array_equal(ops.decode_png_rgb8(data), arr)
    # hand-built files: one per filter type, so the test does not depend on PIL's heuristic
    arr = smooth_rgb(24, 40, 99)
    for ft in range(5):
        data = encode_with_filter(arr, ft)
        assert filter_types(data) == {ft}
        assert np.array_equal(np.asarray(PIL.open(io.BytesIO(data)).convert('RGB')), arr)       # the hand-built file is a valid PNG
        assert np.array_equal(ops.decode_png_rgb8(data), arr), ft
        seen.add(ft)
    assert seen == {0, 1, 2, 3, 4}


def encode_with_filter(arr, ft):
    h, w, c = arr.shape
    bpp = c
    rows = b''
    prev = np.zeros(w * c, np.int32)
    for y in range(h):
        cur = arr[y].reshape(-1).astype(np.int32)
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - ((left + prev) >> 1)
        else:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        rows += bytes([ft]) + (f & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack('>I', len(d)) + t + d + struct.pack('>I', zlib.crc32(t + d) & 0xffffffff)
    ihdr = struct.pack('>IIBBBBB', w, h, 8, 2, 0, 0, 0)
    comp = zlib.compress(rows, 6)
    half = len(comp) // 2                                           # two IDAT chunks: the stream continues across chunk boundaries
    return b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', ihdr) + chunk(b'IDAT', comp[:half]) + chunk(b'IDAT', comp[half:]) + chunk(b'IEND', b'')


@pytest.mark.parametrize('mode', ['RGBA', 'L', 'LA', 'P'])
def test_convert_rgb_of_other_colour_types(mode):
    arr = smooth_rgb(40, 56, 5)
    img = PIL.fromarray(arr, 'RGB')
    if mode == 'RGBA':
        a = np.random.RandomState(1).randint(0, 256, size=arr.shape[:2]).astype(np.uint8)
        img = PIL.fromarray(np.dstack([arr, a]), 'RGBA')
    elif mode == 'L':
        img = img.convert('L')
    elif mode == 'LA':
        img = img.convert('LA')
    else:
        img = img.convert('P', palette=PIL.Palette.ADAPTIVE, colors=200) if hasattr(PIL, 'Palette') else img.quantize(200)
    data = png_bytes(img)
    want = np.asarray(PIL.open(io.BytesIO(data)).convert('RGB'))
    assert np.array_equal(ops.decode_png_rgb8(data), want)


@pytest.mark.parametrize('h,w', [(3, 5), (48, 80), (352, 1216)])
def test_depth16_matches_pil(h, w):
    r = np.random.RandomState(h + w)
    depth = (r.rand(h, w) * 80 * 256).astype(np.uint16)
    depth[r.rand(h, w) > 0.05] = 0                                  # sparse, as a projected LiDAR scan
    data = png_bytes(PIL.fromarray(depth, 'I;16'))
    want = np.array(PIL.open(io.BytesIO(data)))
    got = ops.decode_png_gray16(data)
    assert got.dtype == np.uint16 and np.array_equal(got, want) and np.array_equal(got, depth)
    assert ops.png_info(data) == (w, h, 1, 16)
    # the reference's load_depth (src/data_utils.py:204-234): z = array / 256, z[z <= 0] = 0
    z_ref = np.array(PIL.open(io.BytesIO(data)), dtype=np.float32) / 256.0
    z_ref[z_ref <= 0] = 0.0
    assert np.array_equal(got.astype(np.float32) / 256.0, z_ref)


def test_corrupt_files_fail_loudly():
    arr = smooth_rgb(16, 16, 3)
    data = bytearray(png_bytes(PIL.fromarray(arr, 'RGB')))
    with pytest.raises(RuntimeError, match='signature'):
        ops.decode_png_rgb8(b'not a png' * 10)
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x40                                       # flips a bit inside IDAT: the chunk CRC catches it
    with pytest.raises(RuntimeError, match='CRC'):
        ops.decode_png_rgb8(bytes(bad))
    with pytest.raises(RuntimeError, match='IEND|past the end'):
        ops.decode_png_rgb8(bytes(data[:len(data) - 20]))
    with pytest.raises(RuntimeError, match='grey'):
        ops.decode_png_gray16(bytes(data))                           # an RGB file is not a depth map
    img = PIL.fromarray(arr, 'RGB')
    buf = io.BytesIO()
    try:
        img.save(buf, format='PNG', interlace=1)                    # not every Pillow build writes Adam7
    except Exception:
        return
    if buf.getvalue()[8 + 8 + 12] == 1:
        with pytest.raises(RuntimeError, match='interlace'):
            ops.decode_png_rgb8(buf.getvalue())
