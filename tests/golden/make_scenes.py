"""Writes the scene INPUTS the BASELINE.json configs name (C1 = the reference's Cornell box, C2 = its smallpt scene)
as JSON in the reference's scene format (src/scene.rs:185-850), plus the cube mesh they instance.

The GPU box has no /root/reference, so the test inputs are generated from the values below (scene data, not code):
    python tests/golden/make_scenes.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "scenes")

FILM = {"width": 800, "height": 600, "samples": 4, "frames": 1, "start_frame": 0, "end_frame": 0, "scene_time": 0,
        "filter": {"type": "mitchell_netravali", "width": 2.0, "height": 2.0, "b": 1.0 / 3.0, "c": 1.0 / 3.0}}
CAMERA = {"fov": 30, "transform": [{"type": "translate", "translation": [0, 12, -60]}]}
INTEGRATOR = {"type": "pathtracer", "min_depth": 4, "max_depth": 8}


def wall(name, material, steps):
    return {"name": name, "type": "receiver", "material": material, "geometry": {"type": "plane"}, "transform": steps}


def walls(scale_back, scale_side, scale_floor, mats):
    sc = lambda s: {"type": "scale", "scaling": s}
    tr = lambda t: {"type": "translate", "translation": t}
    ry = lambda r: {"type": "rotate_y", "rotation": r}
    rx = lambda r: {"type": "rotate_x", "rotation": r}
    return {"type": "group", "name": "walls", "transform": [tr([0, 12, 0])], "objects": [
        wall("back_wall", mats[0], [sc(scale_back), tr([0, 0, 20])]),
        wall("left_wall", mats[1], [sc(scale_side), ry(90.0), tr([-15.0, 0, 0])]),
        wall("right_wall", mats[2], [sc(scale_side), ry(-90.0), tr([15.0, 0, 0])]),
        wall("top_wall", mats[0], [sc(scale_floor), rx(90.0), tr([0.0, 12, 0])]),
        wall("bottom_wall", mats[0], [sc(scale_floor), rx(90), tr([0.0, -12, 0])])]}


def cornell():
    cube = lambda name, s, r, t: {"name": name, "type": "receiver", "material": "white_plastic",
                                  "geometry": {"type": "mesh", "file": "models/unit_cube.obj", "model": "Cube"},
                                  "transform": [{"type": "scale", "scaling": s}, {"type": "rotate_y", "rotation": r}, {"type": "translate", "translation": t}]}
    return {"film": FILM, "camera": CAMERA, "integrator": INTEGRATOR, "materials": [
        {"type": "matte", "name": "white_wall", "diffuse": [0.740063, 0.742313, 0.733934], "roughness": 1.0},
        {"type": "matte", "name": "red_wall", "diffuse": [0.366046, 0.0371827, 0.0416385], "roughness": 1.0},
        {"type": "matte", "name": "green_wall", "diffuse": [0.162928, 0.408903, 0.0833759], "roughness": 1.0},
        {"type": "plastic", "name": "white_plastic", "diffuse": [0.8, 0.8, 0.8], "gloss": [0.6, 0.6, 0.6], "roughness": 0.5}],
        "objects": [walls([15, 12, 1], [20, 12, 1], [15, 20, 1], ["white_wall", "red_wall", "green_wall"]),
                    {"name": "light", "type": "emitter", "material": "white_wall", "emitter": "area", "emission": [1, 0.772549, 0.560784, 40],
                     "geometry": {"type": "rectangle", "width": 6, "height": 6},
                     "transform": [{"type": "rotate_x", "rotation": 90}, {"type": "translate", "translation": [0, 23.8, 0]}]},
                    cube("tall_cube", [4, 10, 4], -20, [-6, 5, 6]), cube("short_block", [4, 5, 4], 15, [4, 2.5, -3.0])]}


def smallpt():
    film = dict(FILM, samples=16)
    sphere = lambda name, mat, t: {"name": name, "type": "receiver", "material": mat, "geometry": {"type": "sphere", "radius": 1.0},
                                   "transform": [{"type": "scale", "scaling": 5.0}, {"type": "translate", "translation": t}]}
    return {"film": film, "camera": CAMERA, "integrator": INTEGRATOR, "materials": [
        {"type": "matte", "name": "white_wall", "diffuse": [1.0, 1.0, 1.0], "roughness": 1.0},
        {"type": "matte", "name": "red_wall", "diffuse": [1.0, 0.2, 0.2], "roughness": 1.0},
        {"type": "matte", "name": "blue_wall", "diffuse": [0.2, 0.2, 1.0], "roughness": 1.0},
        {"type": "metal", "name": "metal", "refractive_index": [0.155265, 0.116723, 0.138381], "absorption_coefficient": [4.82835, 3.12225, 2.14696], "roughness": 0.2},
        {"type": "plastic", "name": "plastic", "gloss": [0.8, 0.8, 0.8], "diffuse": [0.8, 0.2, 0.2], "roughness": 0.02},
        {"type": "glass", "name": "glass", "reflect": [1.0, 1.0, 1.0], "transmit": [1.0, 1.0, 1.0], "eta": 1.52}],
        "objects": [walls(32.0, 32.0, 32.0, ["white_wall", "red_wall", "blue_wall"]),
                    sphere("metal_sphere", "metal", [-6.0, 5.0, 8.0]), sphere("glass_sphere", "glass", [6.0, 5.0, -2.0]),
                    {"name": "light", "type": "emitter", "material": "white_wall", "emitter": "area", "emission": [0.780131, 0.780409, 0.775833, 60],
                     "geometry": {"type": "sphere", "radius": 1.0}, "transform": [{"type": "translate", "translation": [0.0, 22, 0]}]}]}


# the unit cube the Cornell scene instances: 8 corners (two carry the modelling tool's 1e-6 jitter), one uv chart and one
# normal per face, six quads
CUBE_V = [(1, -1, -1), (1, -1, 1), (-1, -1, 1), (-1, -1, -1), (1, 1, -0.999999), (0.999999, 1, 1.000001), (-1, 1, 1), (-1, 1, -1)]
CUBE_VT = [(0.0, 0.334353), (0.332314, 0.333333), (0.333333, 0.665647), (0.001019, 0.666667), (1.0, 0.001019), (0.998981, 0.333333),
           (0.666667, 0.332314), (0.667686, 0.0), (1.0, 0.665647), (0.667686, 0.666667), (0.666667, 0.334353), (0.334353, 0.666667),
           (0.333333, 0.334353), (0.665647, 0.333333), (0.666667, 0.665647), (0.333333, 0.332314), (0.00102, 0.333333), (0.0, 0.00102),
           (0.332314, 0.0), (0.333333, 0.001019), (0.665647, 0.0), (0.334353, 0.333333)]
CUBE_VN = [(0, -1, 0), (0, 1, 0), (1, 0, 0), (-0.0, -0.0, 1), (-1, -0.0, -0.0), (0, 0, -1)]
CUBE_F = [[(1, 1, 1), (2, 2, 1), (3, 3, 1), (4, 4, 1)], [(5, 5, 2), (8, 6, 2), (7, 7, 2), (6, 8, 2)], [(1, 6, 3), (5, 9, 3), (6, 10, 3), (2, 11, 3)],
          [(2, 12, 4), (6, 13, 4), (7, 14, 4), (3, 15, 4)], [(3, 16, 5), (7, 17, 5), (8, 18, 5), (4, 19, 5)], [(5, 20, 6), (1, 21, 6), (4, 7, 6), (8, 22, 6)]]


def cube_obj():
    lines = ["o Cube"]
    lines += ["v %.6f %.6f %.6f" % v for v in CUBE_V]
    lines += ["vt %.6f %.6f" % v for v in CUBE_VT]
    lines += ["vn %.6f %.6f %.6f" % v for v in CUBE_VN]
    lines += ["f " + " ".join("%d/%d/%d" % c for c in f) for f in CUBE_F]
    return "\n".join(lines) + "\n"


def tr15_like():
    """C5-shaped input (BASELINE configs[4], the reference's tr15 scene): every feature tr15.json uses (plus an animated field
    of view, scene.rs:286-304) — B-spline keyframed camera, keyframed groups with default degree 3 and repeated knots, keyframed objects inside keyframed groups, area
    lights on disks with keyframed emission whose 4th component ramps from 0, OBJ meshes, a MERL material — at a size the
    oracle renders in seconds. The models and the MERL table are stand-ins (the originals are not redistributable)."""
    tr = lambda t: {"type": "translate", "translation": t}
    sc = lambda s: {"type": "scale", "scaling": s}
    rx = lambda r: {"type": "rotate_x", "rotation": r}
    ry = lambda r: {"type": "rotate_y", "rotation": r}
    rz = lambda r: {"type": "rotate_z", "rotation": r}
    cp = lambda *steps: {"transform": list(steps)}
    film = dict(FILM, width=1920, height=1080, samples=2048, frames=50, start_frame=0, end_frame=49, scene_time=25)
    camera = {"fov": [40, 40, 34, 44], "fov_knots": [0, 0, 0, 12, 25, 25, 25], "fov_spline_degree": 2, "keyframes": {"control_points": [
        cp(rx(15), tr([0, 12, -46])), cp(rx(15), tr([0, 12, -46])), cp(rx(9), tr([0, 14, -50])), cp(rx(9), tr([0, 14, -50])),
        cp(rx(8), ry(25), tr([-19, 14, -40])), cp(rx(8), ry(35), tr([-13, 14, -42])), cp(rx(8), ry(45), tr([-3, 16, -44]))],
        "knots": [0, 0, 0, 0, 6, 12, 18, 25, 25, 25, 25]}}
    cube = lambda name, mat, key: dict({"name": name, "type": "receiver", "material": mat,
                                        "geometry": {"type": "mesh", "file": "models/unit_cube.obj", "model": "Cube"}}, **key)
    sphere = lambda name, mat, key: dict({"name": name, "type": "receiver", "material": mat, "geometry": {"type": "sphere", "radius": 1.0}}, **key)
    rise = {"keyframes": {"control_points": [cp(tr([0, 12, 0])), cp(tr([0, 12, 0])), cp(tr([0, 10, 0])), cp(tr([0, 10, 0]))],
                          "knots": [2.5, 2.5, 2.5, 2.5, 6.0, 6.0, 6.0, 6.0]}}
    return {"film": film, "camera": camera, "integrator": {"type": "pathtracer", "min_depth": 5, "max_depth": 10}, "materials": [
        {"type": "matte", "name": "white_wall", "diffuse": [0.740063, 0.742313, 0.733934], "roughness": 1.0},
        {"type": "matte", "name": "red_wall", "diffuse": [0.366046, 0.0371827, 0.0416385], "roughness": 1.0},
        {"type": "matte", "name": "green_wall", "diffuse": [0.162928, 0.408903, 0.0833759], "roughness": 1.0},
        {"type": "plastic", "name": "plastic", "diffuse": [0.2, 0.5, 0.8], "gloss": [0.7, 0.7, 0.7], "roughness": 0.1},
        {"type": "metal", "name": "metal", "refractive_index": [0.155265, 0.116723, 0.138381], "absorption_coefficient": [4.82835, 3.12225, 2.14696], "roughness": 0.3},
        {"type": "glass", "name": "glass", "reflect": [1.0, 1.0, 1.0], "transmit": [1.0, 1.0, 1.0], "eta": 1.52},
        {"type": "merl", "name": "measured", "file": "merl/synthetic.binary"}],
        "objects": [
            dict({k: v for k, v in walls([15, 12, 1], [20, 12, 1], [15, 20, 1], ["white_wall", "red_wall", "green_wall"]).items() if k != "transform"}, **rise),
            {"type": "group", "name": "toys", "keyframes": {"control_points": [cp(tr([-4, 0, 0])), cp(ry(40), tr([0, 1, 2])), cp(ry(80), tr([4, 0, 0])),
                                                                               cp(ry(120), tr([2, 2, -3])), cp(ry(200), tr([-2, 0, -2]))],
                                                            "knots": [0, 0, 0, 0, 12, 25, 25, 25, 25]},
             "objects": [
                 cube("spinning_block", "measured", {"keyframes": {"control_points": [cp(sc([2, 3, 2]), tr([-5, 3, 4])), cp(sc([2, 3, 2]), ry(120), tr([-5, 5, 4])),
                                                                                          cp(sc([2, 3, 2]), ry(240), tr([-5, 3, 4]))], "degree": 2, "knots": [0, 0, 0, 25, 25, 25]}}),
                 sphere("metal_ball", "metal", {"transform": [sc(2.5), tr([4, 2.5, 2])]}),
                 sphere("glass_ball", "glass", {"keyframes": {"control_points": [cp(sc(2.0), tr([0, 8, -4])), cp(sc(3.0), tr([1, 10, -4]))], "degree": 1,
                                                              "knots": [0, 0, 25, 25]}})]},
            cube("static_block", "plastic", {"transform": [sc([3, 1, 3]), ry(15), tr([7, 1, 8])]}),
            {"name": "large_light1", "type": "emitter", "material": "white_wall", "emitter": "area",
             "emission": [{"time": 0, "color": [1, 0.772549, 0.560784, 0]}, {"time": 3.5, "color": [1, 0.772549, 0.560784, 0]},
                          {"time": 5, "color": [1, 0.772549, 0.560784, 110]}, {"time": 7, "color": [0.88, 0.772549, 0.560784, 150]}],
             "geometry": {"type": "disk", "radius": 5, "inner_radius": 0.0}, "transform": [ry(30), rz(-40), tr([-8, 17, -8])]},
            {"name": "panel", "type": "emitter", "material": "white_wall", "emitter": "area", "emission": [1, 1, 1, 12],
             "geometry": {"type": "rectangle", "width": 6, "height": 6}, "transform": [rx(90), tr([0, 21.5, 0])]},
            {"name": "spark", "type": "emitter", "emitter": "point", "emission": [{"time": 0, "color": [1, 0.9, 0.8, 40]}, {"time": 25, "color": [0.6, 0.7, 1.0, 160]}],
             "keyframes": {"control_points": [cp(tr([-10, 18, -12])), cp(tr([10, 20, -10]))], "degree": 1, "knots": [0, 0, 25, 25]}}]}


def write_synthetic_merl(path):
    """A MERL-format file (int32 dims 90,90,180 then r,g,b planes of float64, material/merl.rs:51-84) holding an analytic lobe:
    real measured BRDF files are not redistributable. 35 MB, so it is generated where needed instead of committed."""
    import numpy as np
    os.makedirs(os.path.dirname(path), exist_ok=True)
    th = (np.arange(90) / 90.0) ** 2 * (np.pi / 2)
    td = np.arange(90) / 90.0 * (np.pi / 2)
    base = (np.exp(-(th[:, None] ** 2) / 0.05) * 4.0 + 0.2) * (0.3 + 0.04 + 0.96 * (1 - np.cos(td[None, :])) ** 5)
    t = np.repeat(base[:, :, None], 180, axis=2).reshape(-1)
    with open(path, "wb") as f:
        f.write(np.array([90, 90, 180], np.int32).tobytes())
        for scale, k in ((1500.0, 0.9), (1500.0, 0.7), (1500.0 / 1.66, 0.5)):
            f.write((t * k * scale).astype(np.float64).tobytes())


if __name__ == "__main__":
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)
    json.dump(cornell(), open(os.path.join(OUT, "c1_cornell_box.json"), "w"), separators=(",", ":"))
    json.dump(smallpt(), open(os.path.join(OUT, "c2_smallpt.json"), "w"), separators=(",", ":"))
    json.dump(tr15_like(), open(os.path.join(OUT, "c5_tr15_like.json"), "w"), separators=(",", ":"))
    open(os.path.join(OUT, "models", "unit_cube.obj"), "w").write(cube_obj())
    print("wrote", OUT)
