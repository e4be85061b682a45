"""Regenerates tests/golden/*.npz from the CPU oracle.  Run from the repo root: python tests/golden/make_golden.py

The vectors are what the REFERENCE'S OWN SHADER TEXT computes: every case is rendered twice — by the hand-written oracle and
by oracle/_ref/libref_shader.so (brick_raytracer.comp + rand.comp / image.frag compiled by g++, oracle/ref_shim/) — and the
script refuses to write a file unless the two agree bit for bit (RGBA8 and primary hit records).  The oracle supplies the
extra AOV fields (cell / voxel indices, step counters) the shader does not output.  /root/reference does not travel to the
GPU box, these files do.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import zig_vulkan_b200 as zv  # noqa: E402
from zig_vulkan_b200 import scenes  # noqa: E402
from oracle import orc  # noqa: E402

POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))

# name -> (n_voxels, brick_dim, width, height, sun, sun_radius, spp, max_bounce, pose)
CASES = {
    "c1_64_256x256": (64, 4, 256, 256, False, 0.0, 1, 0, POSE0),
    "c3_128_320x180_sun": (128, 4, 320, 180, True, 0.0, 1, 0, POSE0),
    "look_64_160x90_spp2_bounce2": (64, 4, 160, 90, True, 5.0, 2, 2, POSE0),
    "bd8_64_160x90_sun": (64, 8, 160, 90, True, 0.0, 1, 0, POSE0),
    "bd16_128_160x90_sun": (128, 16, 160, 90, True, 0.0, 1, 0, POSE0),
    "inside_64_160x90_sun": (64, 4, 160, 90, True, 0.0, 1, 0, dict(origin=(3.0, 4.0, 5.0), euler_deg=(-10.0, 140.0, 0.0))),
}


def render_case(case):
    n, bd, w, h, sun_on, radius, spp, bounce, pose = case
    grid = scenes.build_grid(n, brick_dim=bd)
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    cam = scenes.camera(w, h, spp=spp, max_bounce=bounce, **pose)
    sun = scenes.sun(sun_on, radius)
    img, aov, cnt = sc.render(cam, sun, aov=True)
    return grid, cam, sun, img, aov, cnt


def check_against_reference_text(case, img, aov):
    """The same case through the reference's shader text (oracle/_ref): must be identical."""
    from oracle import ref

    n, bd, w, h, sun_on, radius, spp, bounce, pose = case
    grid = scenes.build_grid(n, brick_dim=bd)
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    rimg, hits = ref.render(sc, scenes.camera(w, h, spp=spp, max_bounce=bounce, **pose), scenes.sun(sun_on, radius), hits=True)
    hit = (aov["flags"] & 1) != 0
    ok = np.array_equal(rimg, img) and np.array_equal(hit, hits["hit"] != 0) and np.array_equal(aov["material"][hit], hits["index"][hit])
    for f in ("t", "point", "normal"):
        ok = ok and np.array_equal(aov[f].view(np.uint32)[hit], hits[f].view(np.uint32)[hit])
    if not ok:
        raise SystemExit("make_golden: the oracle and the reference's shader text disagree — nothing written")


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, case in CASES.items():
        _, _, _, img, aov, cnt = render_case(case)
        check_against_reference_text(case, img, aov)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), rgba=img, aov=aov, counters=np.array([cnt[k] for k in orc.COUNTER_NAMES], dtype=np.uint64))
        print(name, cnt)
    make_denoise_golden(out_dir)


def make_denoise_golden(out_dir):
    """The present pass (image.frag) over the traced c1 frame: same size RGBA, and a 1.5x BGRA target."""
    from oracle import ref

    _, _, _, img, _, _ = render_case(CASES["c1_64_256x256"])
    if not (np.array_equal(ref.present(img), orc.denoise(img)) and
            np.array_equal(ref.present(img, out_width=384, out_height=216, flags=1), orc.denoise(img, out_width=384, out_height=216, flags=1))):
        raise SystemExit("make_golden: the denoise oracle and image.frag disagree — nothing written")
    np.savez_compressed(os.path.join(out_dir, "denoise_c1_64_256x256.npz"), traced=img, denoised=orc.denoise(img),
                        denoised_384x216_bgra=orc.denoise(img, out_width=384, out_height=216, flags=1))


if __name__ == "__main__":
    main()
