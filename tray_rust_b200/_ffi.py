"""ctypes mirror of include/trb.h (the C ABI) and loaders for the shared libraries.

The product library is ``tray_rust_b200/lib/libtrb.so`` (CUDA kernels + host code,
built by ``__graft_entry__.build()``). There is NO CPU fallback: if the library is
missing, ``load_trb()`` raises.

The parity oracle has its own bindings under ``oracle/pyoracle.py`` (test infrastructure);
nothing in this package imports, loads or links it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)

TRB_ABI_VERSION = 4
TRB_OK, TRB_INVALID_ARG, TRB_CUDA, TRB_OOM, TRB_UNSUPPORTED, TRB_IO, TRB_NO_DEVICE, TRB_NCCL = range(8)
INST_RECEIVER, INST_EMITTER_AREA, INST_EMITTER_POINT = 0, 1, 2
SHAPE_NONE, SHAPE_SPHERE, SHAPE_DISK, SHAPE_RECT, SHAPE_MESH = 0, 1, 2, 3, 4
MAT_MATTE, MAT_PLASTIC, MAT_METAL, MAT_SPECULAR_METAL, MAT_GLASS, MAT_ROUGH_GLASS, MAT_MERL = range(7)
FILTER_MITCHELL_NETRAVALI, FILTER_GAUSSIAN = 0, 1
INTEGRATOR_PATH, INTEGRATOR_WHITTED, INTEGRATOR_NORMALS_DEBUG = 0, 1, 2
RENDER_STATS, RENDER_NO_UPDATE, RENDER_REFERENCE_SHADOW, RENDER_MEGAKERNEL, RENDER_TIME_TRACE = 1, 2, 4, 8, 16
MISS = 0xFFFFFFFF
BVH_LEAF = 0x80000000
MERL_TABLE_FLOATS = 90 * 90 * 180 * 3

u32, f32 = C.c_uint32, C.c_float


class Keyframe(C.Structure):
    _fields_ = [("translation", f32 * 3), ("rotation", f32 * 4), ("scaling", f32 * 3)]


class Spline(C.Structure):
    _fields_ = [("degree", u32), ("n_ctrl", u32), ("ctrl_first", u32), ("n_knots", u32), ("knot_first", u32)]


class ColorKey(C.Structure):
    _fields_ = [("rgba", f32 * 4), ("time", f32)]


class Instance(C.Structure):
    _fields_ = [("kind", u32), ("shape", u32), ("p0", f32), ("p1", f32), ("mesh", u32), ("material", u32),
                ("spline_first", u32), ("n_splines", u32), ("emission_first", u32), ("n_emission", u32)]


class Mesh(C.Structure):
    _fields_ = [("n_verts", u32), ("n_tris", u32), ("positions", C.POINTER(f32)), ("normals", C.POINTER(f32)),
                ("texcoords", C.POINTER(f32)), ("indices", C.POINTER(u32))]


class Image(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("rgba8", C.POINTER(C.c_uint8)), ("time", f32), ("pad", u32)]


class Texture(C.Structure):
    _fields_ = [("first_image", u32), ("n_images", u32)]


class Material(C.Structure):
    _fields_ = [("type", u32), ("c0", f32 * 3), ("c1", f32 * 3), ("roughness", f32), ("eta", f32), ("merl", u32), ("tex", u32 * 4)]


class Camera(C.Structure):
    _fields_ = [("spline_first", u32), ("n_splines", u32), ("fov", f32), ("shutter_size", f32), ("active_at", u32),
                ("fov_degree", u32), ("n_fov_ctrl", u32), ("fov_ctrl_first", u32), ("n_fov_knots", u32),
                ("fov_knot_first", u32)]


class Film(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("samples", u32), ("frames", u32), ("start_frame", u32),
                ("end_frame", u32), ("scene_time", f32), ("filter_type", u32), ("filter_w", f32), ("filter_h", f32),
                ("filter_b", f32), ("filter_c", f32)]


class Integrator(C.Structure):
    _fields_ = [("type", u32), ("min_depth", u32), ("max_depth", u32)]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", u32), ("film", Film), ("integrator", Integrator),
                ("n_cameras", u32), ("cameras", C.POINTER(Camera)),
                ("n_instances", u32), ("instances", C.POINTER(Instance)),
                ("n_splines", u32), ("splines", C.POINTER(Spline)),
                ("n_keyframes", u32), ("keyframes", C.POINTER(Keyframe)),
                ("n_knots", u32), ("knots", C.POINTER(f32)),
                ("n_color_keys", u32), ("color_keys", C.POINTER(ColorKey)),
                ("n_meshes", u32), ("meshes", C.POINTER(Mesh)),
                ("n_materials", u32), ("materials", C.POINTER(Material)),
                ("n_merl", u32), ("merl_tables", C.POINTER(C.POINTER(f32))),
                ("n_fov_floats", u32), ("fov_floats", C.POINTER(f32)),
                ("n_textures", u32), ("textures", C.POINTER(Texture)),
                ("n_images", u32), ("images", C.POINTER(Image))]


class RenderCfg(C.Structure):
    _fields_ = [("spp", u32), ("sample_first", u32), ("sample_count", u32), ("block_start", u32),
                ("block_count", u32), ("current_frame", u32), ("seed", u32), ("flags", u32),
                ("shard_index", u32), ("shard_count", u32), ("shard_chunk", u32)]


class Stats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("rays_mis", C.c_uint64), ("rays_continuation", C.c_uint64), ("node_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("inst_tests", C.c_uint64), ("kernel_ms", f32), ("update_ms", f32)]

    def rays_total(self):
        return self.rays_primary + self.rays_shadow + self.rays_mis + self.rays_continuation

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Ray(C.Structure):
    _fields_ = [("o", f32 * 3), ("d", f32 * 3), ("min_t", f32), ("max_t", f32)]


class Hit(C.Structure):
    _fields_ = [("t", f32), ("inst", u32), ("prim", u32), ("pad", u32)]


class Sample(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("r", f32), ("g", f32), ("b", f32)]


class BvhNode(C.Structure):
    _fields_ = [("bmin", f32 * 3), ("bmax", f32 * 3), ("a", u32), ("b", u32)]


# numpy dtypes with identical layout
import numpy as np  # noqa: E402

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("min_t", "<f4"), ("max_t", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("inst", "<u4"), ("prim", "<u4"), ("pad", "<u4")])
SAMPLE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("g", "<f4"), ("b", "<f4")])
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("a", "<u4"), ("b", "<u4")])

TRB_SYMBOLS = [
    "trb_scene_create", "trb_scene_load_json", "trb_scene_destroy", "trb_scene_info", "trb_scene_update_frame",
    "trb_render", "trb_render_device", "trb_intersect", "trb_intersect_device", "trb_camera_rays",
    "trb_render_samples", "trb_film_to_srgb8", "trb_block_list", "trb_scene_get_bvh", "trb_scene_get_transform",
    "trb_scene_get_filter_table", "trb_last_error", "trb_abi_version", "trb_desc_load_json", "trb_desc_free",
    "trb_host_build_bvh", "trb_host_keyframe_transform", "trb_host_animated_transform", "trb_host_animated_color", "trb_host_quad_check", "trb_selftest_box", "trb_launch_count", "trb_scene_trace_time", "trb_scene_check_error", "trb_scene_set_option", "trb_write_png",
    "trb_nccl_unique_id", "trb_comm_create", "trb_comm_destroy", "trb_comm_info", "trb_comm_reduce_film", "trb_render_sharded",
    "trb_group_create", "trb_group_load_json", "trb_group_render", "trb_group_scene", "trb_group_destroy",
]

_trb = None


def trb_path():
    return os.path.join(_HERE, "lib", "libtrb.so")


def load_trb():
    """Load the product library. Raises if it has not been built — never falls back."""
    global _trb
    if _trb is not None:
        return _trb
    p = trb_path()
    if not os.path.exists(p):
        raise RuntimeError("libtrb.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "tray_rust_b200 has no CPU fallback." % p)
    lib = C.CDLL(p)
    vp, sz = C.c_void_p, C.c_size_t
    lib.trb_last_error.restype = C.c_char_p
    lib.trb_abi_version.restype = u32
    lib.trb_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(vp)]
    lib.trb_scene_load_json.argtypes = [C.c_char_p, u32, u32, u32, C.c_int, C.POINTER(vp)]
    lib.trb_scene_destroy.argtypes = [vp]
    lib.trb_scene_destroy.restype = None
    lib.trb_scene_info.argtypes = [vp] + [C.POINTER(u32)] * 6
    lib.trb_scene_update_frame.argtypes = [vp, u32, f32, f32]
    lib.trb_render.argtypes = [vp, C.POINTER(RenderCfg), vp, C.POINTER(Stats)]
    lib.trb_render_device.argtypes = [vp, C.POINTER(RenderCfg), vp, vp, vp]
    lib.trb_intersect.argtypes = [vp, sz, vp, vp, C.POINTER(Stats)]
    lib.trb_intersect_device.argtypes = [vp, sz, vp, vp, vp, vp]
    lib.trb_camera_rays.argtypes = [vp, C.POINTER(RenderCfg), sz, vp, vp]
    lib.trb_render_samples.argtypes = [vp, C.POINTER(RenderCfg), sz, vp, C.POINTER(Stats)]
    lib.trb_film_to_srgb8.argtypes = [vp, vp, vp]
    lib.trb_block_list.argtypes = [vp, u32, u32, C.POINTER(u32), vp, u32]
    lib.trb_scene_get_bvh.argtypes = [vp, C.c_int, C.POINTER(u32), vp, C.POINTER(u32), vp]
    lib.trb_scene_get_transform.argtypes = [vp, u32, vp, vp]
    lib.trb_scene_get_filter_table.argtypes = [vp, vp]
    lib.trb_desc_load_json.argtypes = [C.c_char_p, u32, u32, u32, C.POINTER(C.POINTER(SceneDesc))]
    lib.trb_desc_free.argtypes = [C.POINTER(SceneDesc)]
    lib.trb_desc_free.restype = None
    lib.trb_host_build_bvh.argtypes = [vp, u32, u32, C.POINTER(u32), vp, vp]
    lib.trb_host_keyframe_transform.argtypes = [C.POINTER(Keyframe), vp, vp]
    lib.trb_host_animated_transform.argtypes = [C.POINTER(SceneDesc), u32, u32, f32, vp, vp]
    lib.trb_host_animated_color.argtypes = [C.POINTER(SceneDesc), u32, u32, f32, vp]
    lib.trb_host_quad_check.argtypes = [vp, u32, vp, u32, C.POINTER(u32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.trb_selftest_box.argtypes = [u32, u32, C.POINTER(C.c_uint64)]
    lib.trb_launch_count.restype = C.c_uint64
    lib.trb_scene_trace_time.argtypes = [vp, C.POINTER(f32), C.POINTER(u32)]
    lib.trb_write_png.argtypes = [C.c_char_p, vp, u32, u32]
    lib.trb_scene_check_error.argtypes = [vp]
    lib.trb_scene_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    lib.trb_nccl_unique_id.argtypes = [vp]
    lib.trb_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.trb_comm_destroy.argtypes = [vp]
    lib.trb_comm_destroy.restype = None
    lib.trb_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.trb_comm_reduce_film.argtypes = [vp, vp, sz, C.c_int, vp]
    lib.trb_render_sharded.argtypes = [vp, vp, C.POINTER(RenderCfg), C.c_int, vp, C.POINTER(Stats)]
    lib.trb_group_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.trb_group_load_json.argtypes = [C.c_char_p, u32, u32, u32, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.trb_group_render.argtypes = [vp, C.POINTER(RenderCfg), vp, C.POINTER(Stats)]
    lib.trb_group_scene.argtypes = [vp, C.c_int]
    lib.trb_group_scene.restype = vp
    lib.trb_group_destroy.argtypes = [vp]
    lib.trb_group_destroy.restype = None
    _trb = lib
    return lib


def ptr(a):
    """void* of a contiguous numpy array"""
    return a.ctypes.data_as(C.c_void_p)
