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


if __name__ == "__main__":
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)
    json.dump(cornell(), open(os.path.join(OUT, "c1_cornell_box.json"), "w"), separators=(",", ":"))
    json.dump(smallpt(), open(os.path.join(OUT, "c2_smallpt.json"), "w"), separators=(",", ":"))
    open(os.path.join(OUT, "models", "unit_cube.obj"), "w").write(cube_obj())
    print("wrote", OUT)
