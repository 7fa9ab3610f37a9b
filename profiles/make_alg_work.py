"""Algorithmic work per ray segment for each bench scene.

The event counts N_* come from the INSTRUMENTED ORACLE, i.e. from the reference's own traversal
order (left-then-right, no reordering), so that a smarter GPU traversal cannot inflate its own
roofline fraction.  Unit constants are op counts of the reference formulas:
  flops/ray = 27 N_node + 24 N_sph_rej + 60 N_sph_acc + 36 N_msph + 12 N_rect_rej + 30 N_rect_acc
            + 12 N_translate + 30 N_rotate + 25 N_medium + 165 N_diffuse + 60 N_dielectric + 52 N_metal
  bytes/ray = 32 N_node + 16 N_sph + 36 N_msph + 24 N_rect + 24 N_box + 12 N_translate + 8 N_rotate
            + 12 N_medium + 3 N_texel + 1344/7 N_perlin_noise_calls
(rect tests made inside Rect::pdf_value are part of the 165-flop diffuse bounce, not of N_rect_*).
Writes profiles/alg_work_per_ray.json.  CPU only:  python profiles/make_alg_work.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vecchio_b200 as vb  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SCENES = [("cornell_box", 0, 200, 32, 100), ("cornell_smoke", 0, 200, 32, 100), ("random_spheres_demo", 0, 400, 16, 50),
          ("final_scene", 0, 200, 16, 100), ("stress_spheres", 1000, 320, 4, 50)]


def main():
    out = {}
    for name, param, W, spp, depth in SCENES:
        s = vb.Scene(name, seed=1, param=param)
        cam = s.next_camera()
        o = po.OracleScene(s)
        _, _, st = o.render(cam, vb.render_params(W, s.height_for(W), spp, depth, seed=1))
        c = st.as_dict()
        rays = c["rays"]
        flops = (27 * c["n_node"] + 24 * c["n_sph_rej"] + 60 * c["n_sph_acc"] + 36 * c["n_msph"] + 12 * c["n_rect_rej"] +
                 30 * c["n_rect_acc"] + 12 * c["n_translate"] + 30 * c["n_rotate"] + 25 * c["n_medium"] +
                 165 * c["n_diffuse"] + 60 * c["n_dielectric"] + 52 * c["n_metal"])
        nbytes = (32 * c["n_node"] + 16 * (c["n_sph_rej"] + c["n_sph_acc"]) + 36 * c["n_msph"] + 24 * (c["n_rect_rej"] + c["n_rect_acc"]) +
                  24 * c["n_box"] + 12 * c["n_translate"] + 8 * c["n_rotate"] + 12 * c["n_medium"] + 3 * c["n_texel"] + 192 * c["n_perlin"])
        per_ray = {k: round(v / rays, 4) for k, v in c.items() if k.startswith("n_")}
        out[name] = {"flops_per_ray": round(flops / rays, 2), "bytes_per_ray": round(nbytes / rays, 2),
                     "rays_per_path": round(rays / c["paths"], 4), "events_per_ray": per_ray,
                     "sample": f"{W}x{s.height_for(W)} at {spp} spp, depth {depth}, scene seed 1" + (f", param {param}" if param else "")}
        print(name, out[name]["flops_per_ray"], out[name]["bytes_per_ray"], out[name]["rays_per_path"])
    json.dump(out, open(os.path.join(ROOT, "profiles", "alg_work_per_ray.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
