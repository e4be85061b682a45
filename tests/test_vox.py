"""MagicaVoxel reader (libvrt_host.so vrt_vox_*) against the reference's loader (vox/loader.zig).

The first three tests are the reference's ENTIRE test suite for this repo (loader.zig:265-281), ported one to one; the
rest exercise parseBuffer on synthetic files and — when the reference tree is present, i.e. in the build container —
on its two assets and its default-palette table."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi

REF = "/root/reference"


def validate(buf: bytes) -> int:
    b = (C.c_uint8 * len(buf)).from_buffer_copy(buf)
    return ffi.host_lib().vrt_vox_validate_header(b, len(buf))


def test_validate_header_valid_header_accepted():  # loader.zig:265-269
    assert validate(b"VOX " + bytes([150, 0, 0, 0]) + b"MAIN") == 0


def test_validate_header_invalid_id_detected():  # loader.zig:271-275
    assert validate(b"!VOX" + bytes([150, 0, 0, 0]) + b"MAIN") == -10  # ParseError.InvalidId


def test_validate_header_invalid_version_detected():  # loader.zig:277-281
    assert validate(b"VOX " + bytes([169, 0, 0, 0]) + b"MAIN") == -14  # ParseError.UnexpectedVersion


def test_validate_header_missing_main():
    assert validate(b"VOX " + bytes([150, 0, 0, 0]) + b"NIAM") == -15  # ParseError.InvalidFileContent
    assert validate(b"VOX ") == -15


def chunk(tag: bytes, content: bytes, children: bytes = b"") -> bytes:
    return tag + struct.pack("<ii", len(content), len(children)) + content + children


def make_vox(models, palette=None, pack=True, extra=b""):
    body = b""
    if pack:
        body += chunk(b"PACK", struct.pack("<i", len(models)))
    for size, voxels in models:
        body += chunk(b"SIZE", struct.pack("<iii", *size))
        body += chunk(b"XYZI", struct.pack("<i", len(voxels)) + b"".join(bytes(v) for v in voxels))
    body += extra
    if palette is not None:
        body += chunk(b"RGBA", bytes(palette))
    return b"VOX " + struct.pack("<i", 150) + chunk(b"MAIN", b"", body)


def test_parse_two_models_with_palette_and_unknown_chunk():
    pal = bytes((i * 7) & 0xFF for i in range(1024))
    m0 = ((4, 5, 6), [(1, 2, 3, 9), (0, 0, 0, 1)])
    m1 = ((2, 2, 2), [(1, 1, 1, 255)])
    vox = zv.Vox(make_vox([m0, m1], palette=pal, extra=chunk(b"nTRN", b"\x00" * 8)))
    assert vox.num_models == 2
    assert vox.size(0) == (4, 5, 6) and vox.size(1) == (2, 2, 2)
    assert vox.xyzi(0).tolist() == [[1, 2, 3, 9], [0, 0, 0, 1]] and vox.xyzi(1).tolist() == [[1, 1, 1, 255]]
    p = vox.palette
    assert p[0].tolist() == [0, 0, 0, 1]                       # loader.zig:167-172
    assert p[1:255].tobytes() == pal[: 254 * 4]                # file colour i -> palette[i + 1], 254 entries (loader.zig:173-182)
    # without PACK: one model (loader.zig:73-76)
    vox = zv.Vox(make_vox([m0], pack=False))
    assert vox.num_models == 1 and vox.size(0) == (4, 5, 6)


def test_parse_errors():
    good = make_vox([((1, 1, 1), [(0, 0, 0, 1)])])
    for cut in (10, 25, len(good) - 3):
        with pytest.raises(ffi.VrtError) as e:
            zv.Vox(good[:cut])
        assert e.value.code == -15
    bad_size = good.replace(b"SIZE", b"SIZF")
    with pytest.raises(ffi.VrtError) as e:
        zv.Vox(bad_size)
    assert e.value.code == -11
    with pytest.raises(ffi.VrtError) as e:
        zv.Vox(good.replace(b"XYZI", b"XYZJ"))
    assert e.value.code == -12
    with pytest.raises(ffi.VrtError) as e:
        zv.Vox(b"XOV " + good[4:])
    assert e.value.code == -10
    zv.Vox(b"XOV " + good[4:], strict=False)  # non-strict parsing skips validateHeader (loader.zig:42-44)
    with pytest.raises(ffi.VrtError) as e:
        zv.Vox(path="/nonexistent/file.vox")
    assert e.value.code == -16


def test_default_palette_and_materials():
    vox = zv.Vox(make_vox([((1, 1, 1), [(0, 0, 0, 1)])]))
    p = vox.palette
    assert p[0].tolist() == [0, 0, 0, 0] and p[1].tolist() == [255, 255, 255, 255]
    assert p[2].tolist() == [255, 255, 0xCC, 255] and p[7].tolist() == [255, 0xCC, 255, 255] and p[37].tolist() == [0xCC, 255, 255, 255]
    assert p[215].tolist() == [0, 0, 0x33, 255] and p[216].tolist() == [0xEE, 0, 0, 255] and p[255].tolist() == [0x11, 0x11, 0x11, 255]
    mats = zv.terrain_materials()
    n = vox.materials(mats, 8)  # main.zig:96-108: after the 8 terrain materials
    assert n == 248
    assert mats[8]["type"] == 2 and mats[8]["type_data"] == np.float32(1.52)  # palette[0] has alpha 0 < 0.8 -> dielectric
    assert mats[9]["type"] == 0 and mats[9]["albedo_r"] == 1.0 and mats[10]["albedo_b"] == np.float32(0xCC / 255.0)
    assert mats[7]["type"] == 1  # terrain materials untouched


def test_insert_into_grid_swaps_y_and_z():
    vox = zv.Vox(make_vox([((4, 4, 4), [(1, 2, 3, 5)])]))
    g = ffi.Grid((4, 4, 4))
    assert vox.insert_into(g, offset=(4, 0, 8), material_base=8) == 0  # insert(x+4, z+0, y+8, 5+8) (main.zig:110-118)
    ref = ffi.Grid((4, 4, 4))
    assert ref.insert(1 + 4, 3 + 0, 2 + 8, 13) == 0
    assert np.array_equal(g.occupancy, ref.occupancy) and np.array_equal(g.material_indices, ref.material_indices)
    assert np.array_equal(g.brick_indices, ref.brick_indices)
    assert vox.insert_into(g, offset=(16, 0, 0)) == -1  # out of the grid: the reference asserts


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_assets_and_default_palette_table():
    doom = zv.Vox(path=os.path.join(REF, "assets/models/doom.vox"))
    assert doom.num_models == 1 and doom.size(0) == (126, 126, 126) and len(doom.xyzi(0)) == 3894
    monu = zv.Vox(path=os.path.join(REF, "assets/models/monu10.vox"))
    assert monu.size(0) == (72, 72, 126) and len(monu.xyzi(0)) == 150764
    assert (monu.xyzi(0)[:, 0] < 72).all() and (monu.xyzi(0)[:, 2] < 126).all()
    # main.zig:77-118: doom.vox into the reference's default 128x64x128-brick grid at offset (200, 50, 150)
    g = ffi.Grid((128, 64, 128), min_point=(-32.0, -16.0, -32.0), scale=0.5)
    assert doom.insert_into(g, offset=(200, 50, 150), material_base=8) == 0
    occ = np.unpackbits(g.occupancy[: g.active_bricks * 8], bitorder="little").sum()
    assert occ == len(np.unique(doom.xyzi(0)[:, :3], axis=0))
    # the procedural default palette equals the reference's 256-constant table (loader.zig:246-263)
    text = open(os.path.join(REF, "src/modules/voxel_rt/vox/loader.zig")).read()
    table = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{8}", text[text.index("const default_rgba"):])][:256]
    assert len(table) == 256
    ours = zv.Vox(make_vox([((1, 1, 1), [])])).palette.view("<u4").reshape(-1)
    assert ours.tolist() == table
