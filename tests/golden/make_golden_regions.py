"""Regenerate tests/golden/cornell_sample_regions.json and tests/golden/thenextweek_sample_regions.json from the
reference's published renders.

Reads /root/reference/sample/therestofyourlife.png (900x900, 8-bit, the book-3 Cornell box at
src/main.rs defaults: width 900, 1000 spp, depth 100) -- the ONLY result-pinning artefact the
reference ships (it has no tests).  Only region statistics are committed, not the image.
sample/thenextweek.png (900x900) is the book-2 final scene (src/scene.rs:732-874) rendered by the reference; its
scene is unseeded, so only regions on the seed-independent objects are kept (no ground boxes; not the blue
subsurface sphere either, whose published pixels hold clamped fireflies), as LINEAR means
((v + 0.5) / 256)^2 -- the inverse of Vec3::to_color at the centre of the 8-bit bin -- of unsaturated pixels.
Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_golden_regions.py
"""
import hashlib
import json
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/sample/therestofyourlife.png"
# name: (x0, x1, y0, y1) in image coordinates (y down), SURVEY.md Appendix C
REGIONS = {
    "whole": (0, 900, 0, 900),
    "light_centre": (400, 500, 125, 145),
    "green_wall": (60, 140, 400, 500),
    "red_wall": (760, 840, 400, 500),
    "back_wall": (400, 500, 250, 350),
    "floor_front": (300, 400, 820, 860),
    "ceiling": (200, 300, 60, 100),
    "tall_box_front": (300, 440, 450, 700),
    "glass_sphere_centre": (540, 600, 660, 720),
    "caustic": (545, 605, 798, 810),
    "border": (0, 15, 0, 15),
}


def main():
    raw = open(SRC, "rb").read()
    im = np.asarray(Image.open(SRC).convert("RGB")).astype(np.float64)
    assert im.shape == (900, 900, 3)
    nz = np.argwhere(im.sum(axis=2) > 0)
    out = {
        "source": "sample/therestofyourlife.png",
        "sha256": hashlib.sha256(raw).hexdigest(),
        "size": [900, 900],
        "scene": "cornell_box (src/scene.rs:630-730), 1000 spp, depth 100, to_color (src/vec3.rs:54-61)",
        "nonblack_bbox_rows": [int(nz[:, 0].min()), int(nz[:, 0].max())],
        "nonblack_bbox_cols": [int(nz[:, 1].min()), int(nz[:, 1].max())],
        "regions": {},
    }
    for name, (x0, x1, y0, y1) in REGIONS.items():
        out["regions"][name] = {"box_xyxy": [x0, x1, y0, y1], "mean_rgb8": [round(float(v), 3) for v in im[y0:y1, x0:x1].mean(axis=(0, 1))]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_sample_regions.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


SRC2 = "/root/reference/sample/thenextweek.png"
REGIONS2 = {
    "fog_upper_right": (650, 850, 150, 250),   # the rho = 1e-4 medium that fills the room, in front of the dark wall
    "fog_mid_left": (260, 330, 250, 330),
    "moving_sphere": (100, 180, 220, 300),     # MovingSphere, Lambertian (0.7, 0.3, 0.1), motion-blurred
    "fuzzy_metal": (350, 450, 400, 500),       # Metal (0.8, 0.8, 0.9) fuzz 10
    "earth_land": (100, 170, 450, 520),        # ImageTexture earthmap.png: Asia
    "earth_ocean": (50, 110, 560, 620),        # ... Indian Ocean
    "grey_sphere": (730, 800, 600, 680),       # NoiseTexture(0.1)
    "white_cluster": (520, 640, 320, 420),     # 1000 random white spheres, rotated 15 degrees and translated
}


def main2():
    raw = open(SRC2, "rb").read()
    im = np.asarray(Image.open(SRC2).convert("RGB")).astype(np.float64)
    assert im.shape == (900, 900, 3)
    out = {
        "source": "sample/thenextweek.png",
        "sha256": hashlib.sha256(raw).hexdigest(),
        "size": [900, 900],
        "scene": "final_scene (src/scene.rs:732-874), unseeded; linear = ((v + 0.5) / 256)^2",
        "regions": {},
    }
    for name, (x0, x1, y0, y1) in REGIONS2.items():
        px = im[y0:y1, x0:x1]
        assert (px >= 255).any(axis=2).mean() < 1e-3, name  # (3 clamped pixels of 12 000 in the cluster)
        lin = ((px + 0.5) / 256.0) ** 2
        out["regions"][name] = {"box_xyxy": [x0, x1, y0, y1], "mean_linear": [round(float(v), 5) for v in lin.mean(axis=(0, 1))],
                                "mean_rgb8": [round(float(v), 3) for v in px.mean(axis=(0, 1))]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "thenextweek_sample_regions.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


SRC3 = "/root/reference/sample/inoneweekend.png"
# The book-1 final scene (1024x576), rendered before the reference grew lights: sky-lit, legacy Material::scatter.
# HEAD's scene.rs no longer has this variant; the host front end authors it as `book1_cover`.  Regions on the
# sky and the three big spheres, which do not depend on the unseeded small spheres (beyond their reflections).
REGIONS3 = {
    "sky_top": (100, 900, 0, 30),            # blue is clamped at 255 here: only red and green are compared
    "metal_top": (600, 740, 70, 150),        # Metal (0.7, 0.6, 0.5) fuzz 0 reflecting the sky
    "metal_upper_mid": (560, 780, 170, 215),
    "brown_sphere": (345, 395, 90, 150),     # Lambertian (0.4, 0.2, 0.1)
    "glass_lower": (420, 480, 185, 225),     # Dielectric 1.5: the sky seen through the lower half
    "ground_far_left": (20, 300, 138, 150),  # Lambertian (0.5, 0.5, 0.5) near the horizon
}


def main3():
    raw = open(SRC3, "rb").read()
    im = np.asarray(Image.open(SRC3).convert("RGB")).astype(np.float64)
    assert im.shape == (576, 1024, 3)
    out = {
        "source": "sample/inoneweekend.png",
        "sha256": hashlib.sha256(raw).hexdigest(),
        "size": [1024, 576],
        "scene": "book-1 final scene (book1_cover), sky-lit, legacy integrator; linear = ((v + 0.5) / 256)^2",
        "regions": {},
    }
    for name, (x0, x1, y0, y1) in REGIONS3.items():
        px = im[y0:y1, x0:x1]
        lin = ((px + 0.5) / 256.0) ** 2
        out["regions"][name] = {"box_xyxy": [x0, x1, y0, y1], "mean_linear": [round(float(v), 5) for v in lin.mean(axis=(0, 1))],
                                "clamped_fraction": [round(float(v), 4) for v in (px >= 255).mean(axis=(0, 1))]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inoneweekend_sample_regions.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
    main2()
    main3()
