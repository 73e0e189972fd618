"""ctypes mirror of include/akari_b200.h and include/akari_b200_host.h (POD structs only)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

AKR_OK = 0
AKR_ERR_INVALID_ARGUMENT = 1
AKR_ERR_CUDA = 2
AKR_ERR_UNSUPPORTED = 3
AKR_ERR_OUT_OF_MEMORY = 4
AKR_ERR_STATE = 5

AKR_SVM_MAX_ARGS = 25
AKR_SAMPLER_INDEPENDENT = 0
AKR_SAMPLER_PMJ02BN = 1
AKR_FILTER_BOX = 0
AKR_FILTER_GAUSSIAN = 1


class AkrSvmNode(C.Structure):
    _fields_ = [("op", C.c_uint32), ("n_args", C.c_uint32), ("a", C.c_uint32 * AKR_SVM_MAX_ARGS)]


class AkrShaderKind(C.Structure):
    _fields_ = [("nodes", C.POINTER(AkrSvmNode)), ("n_nodes", C.c_uint32)]


class AkrShaderRef(C.Structure):
    _fields_ = [("shader_kind", C.c_uint32), ("data_offset", C.c_uint32)]


class AkrMesh(C.Structure):
    _fields_ = [
        ("vertices", C.POINTER(C.c_float)),
        ("indices", C.POINTER(C.c_uint32)),
        ("normals", C.POINTER(C.c_float)),
        ("uvs", C.POINTER(C.c_float)),
        ("tangents", C.POINTER(C.c_float)),
        ("material_slots", C.POINTER(C.c_uint32)),
        ("n_vertices", C.c_uint32),
        ("n_triangles", C.c_uint32),
        ("n_material_slots", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class AkrInstance(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 16),
        ("geom_id", C.c_uint32),
        ("flags", C.c_uint32),
        ("materials", C.POINTER(AkrShaderRef)),
        ("n_materials", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class AkrPerspectiveCamera(C.Structure):
    _fields_ = [
        ("c2w", C.c_float * 16),
        ("fov", C.c_float),
        ("lens_radius", C.c_float),
        ("focal_length", C.c_float),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class AkrSceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("n_meshes", C.c_uint32),
        ("n_instances", C.c_uint32),
        ("n_shader_kinds", C.c_uint32),
        ("meshes", C.POINTER(AkrMesh)),
        ("instances", C.POINTER(AkrInstance)),
        ("shader_kinds", C.POINTER(AkrShaderKind)),
        ("shader_data", C.POINTER(C.c_uint8)),
        ("shader_data_size", C.c_size_t),
        ("camera", AkrPerspectiveCamera),
        ("images", C.c_void_p),
        ("n_images", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class AkrPtConfig(C.Structure):
    _fields_ = [
        ("spp", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("spp_per_pass", C.c_uint32),
        ("rr_depth", C.c_uint32),
        ("use_nee", C.c_uint32),
        ("indirect_only", C.c_uint32),
        ("force_diffuse", C.c_uint32),
        ("pixel_offset", C.c_int32 * 2),
        ("debug_depth", C.c_int32),
    ]


class AkrSamplerConfig(C.Structure):
    _fields_ = [("type", C.c_uint32), ("_pad", C.c_uint32), ("seed", C.c_uint64)]


class AkrFilterConfig(C.Structure):
    _fields_ = [("type", C.c_uint32), ("radius", C.c_float)]


class AkrTile(C.Structure):
    _fields_ = [("y0", C.c_uint32), ("y1", C.c_uint32), ("block_rows", C.c_uint32), ("n_shards", C.c_uint32), ("shard", C.c_uint32),
                ("_pad", C.c_uint32)]


class AkrStats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64),
        ("segments", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("gpu_ms", C.c_double),
        ("gpu_ms_kernel", C.c_double * 8),
        ("launches_kernel", C.c_uint64 * 8),
        ("shaded_hits", C.c_uint64),
    ]


class AkrEngineOptions(C.Structure):
    _fields_ = [
        ("wave_size", C.c_uint32),
        ("_unused0", C.c_uint32),
        ("profile_stages", C.c_uint32),
        ("trace_mode", C.c_uint32),
        ("fused", C.c_uint32),
        ("smem_node_kb", C.c_uint32),
        ("aov_mask", C.c_uint32),
        ("_reserved", C.c_uint32 * 1),
    ]


class AkrAovConfig(C.Structure):
    _fields_ = [("spp", C.c_uint32), ("aov", C.c_uint32), ("remap", C.c_uint32), ("_pad", C.c_uint32)]


AOV_NAMES = ["ns", "ng", "tangent", "bitangent", "albedo", "roughness"]  # AKR_AOV_* order (aov.rs:9-22)


class AkrRenderTask(C.Structure):
    _fields_ = [
        ("pt", AkrPtConfig),
        ("sampler", AkrSamplerConfig),
        ("filter", AkrFilterConfig),
        ("out", C.c_char * 512),
        ("method", C.c_uint32),
        ("aov", AkrAovConfig),
    ]


HOST_LIB = os.path.join(HERE, "libakari_b200_host.so")
# AKR_B200_CUDA_LIB points at an alternative build of the same library (kernel-tuning experiments only)
CUDA_LIB = os.environ.get("AKR_B200_CUDA_LIB") or os.path.join(HERE, "libakari_b200.so")

# every symbol the two headers declare (tests check the built libraries export exactly these)
HOST_SYMBOLS = [
    "akr_host_load_scene", "akr_host_free_scene", "akr_host_scene_desc", "akr_host_scene_set_resolution",
    "akr_host_parse_method_file", "akr_host_parse_method_string", "akr_host_default_task",
    "akr_host_write_image", "akr_host_last_error",
]
CUDA_SYMBOLS = [
    "akr_b200_create", "akr_b200_destroy", "akr_b200_last_error", "akr_b200_set_stream",
    "akr_b200_upload_sampler_tables", "akr_b200_upload_albedo_table", "akr_b200_upload_scene",
    "akr_b200_begin", "akr_b200_render_pass", "akr_b200_render_pt", "akr_b200_render_aov", "akr_b200_synchronize",
    "akr_b200_download_film", "akr_b200_resolve_film", "akr_b200_resolve_film_device",
    "akr_b200_get_stats", "akr_b200_reset_stats", "akr_b200_set_engine_options",
    "akr_b200_debug_first_hits", "akr_b200_tile_rows",
]


def load_host_lib():
    if not os.path.exists(HOST_LIB):
        raise RuntimeError(f"{HOST_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(HOST_LIB)
    lib.akr_host_load_scene.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.akr_host_load_scene.restype = C.c_int
    lib.akr_host_free_scene.argtypes = [C.c_void_p]
    lib.akr_host_free_scene.restype = None
    lib.akr_host_scene_desc.argtypes = [C.c_void_p]
    lib.akr_host_scene_desc.restype = C.POINTER(AkrSceneDesc)
    lib.akr_host_scene_set_resolution.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    lib.akr_host_scene_set_resolution.restype = C.c_int
    lib.akr_host_parse_method_file.argtypes = [C.c_char_p, C.POINTER(AkrRenderTask)]
    lib.akr_host_parse_method_file.restype = C.c_int
    lib.akr_host_parse_method_string.argtypes = [C.c_char_p, C.POINTER(AkrRenderTask)]
    lib.akr_host_parse_method_string.restype = C.c_int
    lib.akr_host_default_task.argtypes = [C.POINTER(AkrRenderTask)]
    lib.akr_host_default_task.restype = None
    lib.akr_host_write_image.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint32]
    lib.akr_host_write_image.restype = C.c_int
    lib.akr_host_last_error.argtypes = []
    lib.akr_host_last_error.restype = C.c_char_p
    return lib


def load_cuda_lib():
    """Load the CUDA C-ABI library.  There is no CPU fallback: a missing library is an error."""
    if not os.path.exists(CUDA_LIB):
        raise RuntimeError(f"{CUDA_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(CUDA_LIB)
    vp = C.c_void_p
    lib.akr_b200_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.akr_b200_destroy.argtypes = [vp]
    lib.akr_b200_destroy.restype = None
    lib.akr_b200_last_error.argtypes = [vp]
    lib.akr_b200_last_error.restype = C.c_char_p
    lib.akr_b200_set_stream.argtypes = [vp, vp]
    lib.akr_b200_upload_sampler_tables.argtypes = [vp, vp, vp]
    lib.akr_b200_upload_albedo_table.argtypes = [vp, vp]
    lib.akr_b200_upload_scene.argtypes = [vp, C.POINTER(AkrSceneDesc)]
    lib.akr_b200_begin.argtypes = [vp, C.POINTER(AkrPtConfig), C.POINTER(AkrSamplerConfig),
                                   C.POINTER(AkrFilterConfig), C.POINTER(AkrTile)]
    lib.akr_b200_render_pass.argtypes = [vp, C.c_uint32, C.c_int]
    lib.akr_b200_render_pt.argtypes = [vp, C.POINTER(AkrPtConfig), C.POINTER(AkrSamplerConfig),
                                       C.POINTER(AkrFilterConfig), C.POINTER(AkrTile)]
    lib.akr_b200_render_aov.argtypes = [vp, C.POINTER(AkrAovConfig), C.POINTER(AkrSamplerConfig), C.POINTER(AkrFilterConfig), C.POINTER(AkrTile)]
    lib.akr_b200_synchronize.argtypes = [vp]
    lib.akr_b200_download_film.argtypes = [vp, vp, C.c_size_t]
    lib.akr_b200_resolve_film.argtypes = [vp, vp, C.c_size_t, C.c_int]
    lib.akr_b200_resolve_film_device.argtypes = [vp, vp, C.c_size_t, C.c_int]
    lib.akr_b200_get_stats.argtypes = [vp, C.POINTER(AkrStats)]
    lib.akr_b200_reset_stats.argtypes = [vp]
    lib.akr_b200_set_engine_options.argtypes = [vp, C.POINTER(AkrEngineOptions)]
    lib.akr_b200_debug_first_hits.argtypes = [vp, vp, vp, C.c_size_t]
    lib.akr_b200_tile_rows.argtypes = [C.POINTER(AkrTile)]
    lib.akr_b200_tile_rows.restype = C.c_uint32
    for name in CUDA_SYMBOLS:
        if name not in ("akr_b200_destroy", "akr_b200_last_error", "akr_b200_tile_rows"):
            getattr(lib, name).restype = C.c_int
    return lib
