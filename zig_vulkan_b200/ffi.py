"""ctypes bindings of include/vrt.h (libvrt.so) and include/vrt_host.h (libvrt_host.so).

Struct layouts mirror the reference's ``extern struct``s byte for byte (Camera.zig:183-193, Sun.zig:13-18,
brick/State.zig:60-79, gpu_types.zig:16-32); tests/test_abi.py checks sizes and offsets.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

VRT_ABI_VERSION = 1
VRT_FLAG_AOV = 1
VRT_FLAG_BASELINE = 2
VRT_FLAG_INTERLEAVE = 4
VRT_EXCHANGE_ALLGATHER = 0
VRT_EXCHANGE_PEER_STORE = 1
VRT_EXCHANGE_PEER_FLAGS = 2
VRT_EXCHANGE_HOST = 3
VRT_EXCHANGE_PEER_PUSH = 4
VRT_EXCHANGE_PEER_TILES = 5
VRT_SCHED_STATIC, VRT_SCHED_LPT, VRT_SCHED_DEAL, VRT_SCHED_SHARED = 0, 1, 2, 3
VRT_NCCL_ID_BYTES = 128
VRT_IPC_HANDLE_BYTES = 64

VRT_OK, VRT_E_INVALID, VRT_E_OOM, VRT_E_RANGE, VRT_E_CUDA, VRT_E_NCCL, VRT_E_STATE = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "VRT_OK", -1: "VRT_E_INVALID", -2: "VRT_E_OOM", -3: "VRT_E_RANGE", -4: "VRT_E_CUDA", -5: "VRT_E_NCCL", -6: "VRT_E_STATE"}


class VrtError(RuntimeError):
    """A vrt_* call returned a negative status (the Zig binding maps these to an error set)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class GridState(C.Structure):  # State.zig:60-79
    _fields_ = [
        ("voxel_dim_x", C.c_uint32), ("voxel_dim_y", C.c_uint32), ("voxel_dim_z", C.c_uint32),
        ("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
        ("padding1", C.c_uint32), ("padding2", C.c_uint32),
        ("min_point_base_t", C.c_float * 4),
        ("max_point_scale", C.c_float * 4),
    ]


class CameraDevice(C.Structure):  # Camera.zig:183-193
    _fields_ = [
        ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("_pad0", C.c_uint32 * 2),
        ("horizontal", C.c_float * 3), ("_pad1", C.c_float),
        ("vertical", C.c_float * 3), ("_pad2", C.c_float),
        ("lower_left_corner", C.c_float * 3), ("_pad3", C.c_float),
        ("origin", C.c_float * 3), ("_pad4", C.c_float),
        ("samples_per_pixel", C.c_int32), ("max_bounce", C.c_int32), ("_pad5", C.c_uint32 * 2),
    ]


class SunDevice(C.Structure):  # Sun.zig:13-18
    _fields_ = [("position", C.c_float * 3), ("enabled", C.c_uint32), ("color", C.c_float * 3), ("radius", C.c_float)]


class DenoiseParams(C.Structure):  # GraphicsPipeline.zig:27-32 (PushConstant); defaults :34-39
    _fields_ = [("samples", C.c_int32), ("distribution_bias", C.c_float), ("pixel_multiplier", C.c_float), ("inverse_hue_tolerance", C.c_float)]

    @classmethod
    def default(cls):
        return cls(20, 0.6, 1.5, 20.0)


VRT_DENOISE_BGRA = 1


class BenchmarkReport(C.Structure):  # Benchmark.Report + what Report.print logs (Benchmark.zig:81-139)
    _fields_ = [("min_frame_ms", C.c_float), ("max_frame_ms", C.c_float), ("avg_frame_ms", C.c_float), ("frames", C.c_uint32),
                ("voxel_dim", C.c_uint32 * 3), ("sun_enabled", C.c_uint32), ("image_width", C.c_uint32), ("image_height", C.c_uint32),
                ("max_bounce", C.c_int32), ("samples_per_pixel", C.c_int32)]


class Material(C.Structure):  # gpu_types.zig:16-32
    _fields_ = [("type", C.c_uint32), ("albedo_r", C.c_float), ("albedo_g", C.c_float), ("albedo_b", C.c_float), ("type_data", C.c_float)]


class Aov(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32), ("grid_index", C.c_uint32), ("voxel_index", C.c_uint32), ("material", C.c_uint32),
        ("t", C.c_float), ("point", C.c_float * 3), ("normal", C.c_float * 3),
        ("shadow_grid_index", C.c_uint32), ("shadow_voxel_index", C.c_uint32),
        ("grid_steps", C.c_uint32), ("voxel_steps", C.c_uint32), ("status_fetches", C.c_uint32),
    ]


AOV_DTYPE = np.dtype(
    [
        ("flags", "<u4"), ("grid_index", "<u4"), ("voxel_index", "<u4"), ("material", "<u4"),
        ("t", "<f4"), ("point", "<f4", 3), ("normal", "<f4", 3),
        ("shadow_grid_index", "<u4"), ("shadow_voxel_index", "<u4"),
        ("grid_steps", "<u4"), ("voxel_steps", "<u4"), ("status_fetches", "<u4"),
    ]
)
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("_pad0", "<f4"), ("direction", "<f4", 3), ("_pad1", "<f4")])
RAY_HIT_DTYPE = np.dtype([("hit", "<u4"), ("grid_index", "<u4"), ("voxel_index", "<u4"), ("material", "<u4"), ("t", "<f4"), ("normal", "<f4", 3)])
MATERIAL_DTYPE = np.dtype([("type", "<u4"), ("albedo_r", "<f4"), ("albedo_g", "<f4"), ("albedo_b", "<f4"), ("type_data", "<f4")])


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "primary_hits", "shadow_rays", "grid_steps", "voxel_steps", "status_fetches", "bricks_entered", "hits")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
        ("brick_dim", C.c_uint32), ("material_capacity", C.c_uint32), ("n_bricks", C.c_uint64), ("n_brick_alloc", C.c_uint64),
        ("device", C.c_int32), ("flags", C.c_uint32), ("row_begin", C.c_uint32), ("row_end", C.c_uint32),
        ("part_rank", C.c_uint32), ("part_world", C.c_uint32),
    ]


class HcamConfig(C.Structure):  # Camera.Config
    _fields_ = [
        ("viewport_height", C.c_float), ("origin", C.c_float * 3), ("samples_per_pixel", C.c_int32), ("max_bounce", C.c_int32),
        ("turn_rate", C.c_float), ("normal_speed", C.c_float), ("sprint_speed", C.c_float), ("user_input_disabled", C.c_uint32),
    ]


class HsunConfig(C.Structure):  # Sun.Config
    _fields_ = [
        ("animate", C.c_uint32), ("animate_speed", C.c_float), ("enabled", C.c_uint32), ("color", C.c_float * 3),
        ("radius", C.c_float), ("sun_distance", C.c_float),
    ]


class RendererConfig(C.Structure):  # VoxelRT.Config
    _fields_ = [
        ("internal_resolution_width", C.c_uint32), ("internal_resolution_height", C.c_uint32), ("material_buffer", C.c_uint32),
        ("camera", HcamConfig), ("sun", HsunConfig), ("device", C.c_int32), ("flags", C.c_uint32),
        ("row_begin", C.c_uint32), ("row_end", C.c_uint32),
    ]


# every symbol include/vrt.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
VRT_SYMBOLS = {
    "vrt_init": (C.c_int, [C.POINTER(_P), C.POINTER(Config)]),
    "vrt_deinit": (None, [_P]),
    "vrt_last_error": (C.c_char_p, [_P]),
    "vrt_upload_grid_state": (C.c_int, [_P, C.POINTER(GridState)]),
    "vrt_upload_materials": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_upload_brick_statuses": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_upload_brick_indices": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_upload_brick_occupancy": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_upload_brick_start_indices": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_upload_material_indices": (C.c_int, [_P, _SZ, _P, _SZ]),
    "vrt_trace": (C.c_int, [_P, C.POINTER(CameraDevice), C.POINTER(SunDevice)]),
    "vrt_sync": (C.c_int, [_P]),
    "vrt_read_framebuffer": (C.c_int, [_P, _P, _SZ]),
    "vrt_trace_to_host": (C.c_int, [_P, C.POINTER(CameraDevice), C.POINTER(SunDevice), _P, _SZ]),
    "vrt_trace_to_host_async": (C.c_int, [_P, C.POINTER(CameraDevice), C.POINTER(SunDevice), _P, _SZ]),
    "vrt_trace_rays": (C.c_int, [_P, _P, _P, _SZ]),
    "vrt_trace_rays_host": (C.c_int, [_P, _P, _P, _SZ]),
    "vrt_read_aov": (C.c_int, [_P, _P, _SZ]),
    "vrt_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "vrt_last_trace_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "vrt_last_trace_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "vrt_last_trace_launches": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "vrt_debug_force_accel_rebuild": (C.c_int, [_P]),
    "vrt_debug_tile_stats": (C.c_int, [_P, _P, _SZ]),
    "vrt_set_schedule": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "vrt_sched_get_costs": (C.c_int, [_P, _P, _SZ]),
    "vrt_sched_set_costs": (C.c_int, [_P, _P, _SZ]),
    "vrt_set_stream": (C.c_int, [_P, _P]),
    "vrt_insert_voxels": (C.c_int, [_P, _P, _SZ, C.POINTER(C.c_uint32)]),
    "vrt_download_buffer": (C.c_int, [_P, C.c_uint32, _SZ, _P, _SZ]),
    "vrt_denoise": (C.c_int, [_P, C.POINTER(DenoiseParams), C.c_uint32, C.c_uint32, C.c_uint32]),
    "vrt_read_denoised": (C.c_int, [_P, _P, _SZ]),
    "vrt_denoised_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "vrt_last_denoise_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "vrt_attach_framebuffer": (C.c_int, [_P, _P, _SZ]),
    "vrt_framebuffer_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "vrt_comm_get_unique_id": (C.c_int, [_P]),
    "vrt_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vrt_comm_get_ipc_handle": (C.c_int, [_P, _P]),
    "vrt_comm_open_peers": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vrt_comm_set_exchange": (C.c_int, [_P, C.c_uint32]),
}

VRT_HOST_SYMBOLS = {
    "vrt_grid_create": (_P, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_float * 3), C.c_float, C.c_float]),
    "vrt_grid_destroy": (None, [_P]),
    "vrt_grid_insert": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]),
    "vrt_grid_insert_many": (C.c_int, [_P, _P, _SZ]),
    "vrt_grid_active_bricks": (C.c_uint32, [_P]),
    "vrt_grid_brick_dim": (C.c_uint32, [_P]),
    "vrt_grid_brick_alloc": (C.c_uint64, [_P]),
    "vrt_grid_get_state": (None, [_P, C.POINTER(GridState)]),
    "vrt_grid_statuses": (_P, [_P, C.POINTER(C.c_uint64)]),
    "vrt_grid_brick_indices": (_P, [_P, C.POINTER(C.c_uint64)]),
    "vrt_grid_occupancy": (_P, [_P, C.POINTER(C.c_uint64)]),
    "vrt_grid_start_indices": (_P, [_P, C.POINTER(C.c_uint64)]),
    "vrt_grid_material_indices": (_P, [_P, C.POINTER(C.c_uint64)]),
    "vrt_grid_delta_peek": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vrt_grid_delta_reset": (None, [_P, C.c_int]),
    "vrt_hcam_default_config": (None, [C.POINTER(HcamConfig)]),
    "vrt_hcam_create": (_P, [C.c_float, C.c_uint32, C.c_uint32, C.POINTER(HcamConfig)]),
    "vrt_hcam_destroy": (None, [_P]),
    "vrt_hcam_device": (None, [_P, C.POINTER(CameraDevice)]),
    "vrt_hcam_set_origin": (None, [_P, C.POINTER(C.c_float * 3)]),
    "vrt_hcam_translate": (None, [_P, C.c_float, C.POINTER(C.c_float * 3)]),
    "vrt_hcam_turn_pitch": (None, [_P, C.c_float]),
    "vrt_hcam_turn_yaw": (None, [_P, C.c_float]),
    "vrt_hcam_reset": (None, [_P]),
    "vrt_hcam_activate_sprint": (None, [_P]),
    "vrt_hcam_disable_sprint": (None, [_P]),
    "vrt_hcam_disable_input": (None, [_P]),
    "vrt_hcam_enable_input": (None, [_P]),
    "vrt_hcam_set_orientation": (None, [_P, C.POINTER(C.c_float * 4), C.POINTER(C.c_float * 4)]),
    "vrt_hcam_set_euler_deg": (None, [_P, C.c_float, C.c_float, C.c_float]),
    "vrt_hsun_default_config": (None, [C.POINTER(HsunConfig)]),
    "vrt_hsun_create": (_P, [C.POINTER(HsunConfig)]),
    "vrt_hsun_destroy": (None, [_P]),
    "vrt_hsun_device": (None, [_P, C.POINTER(SunDevice)]),
    "vrt_hsun_update": (None, [_P, C.c_float]),
    "vrt_renderer_default_config": (None, [C.POINTER(RendererConfig)]),
    "vrt_renderer_create": (C.c_int, [C.POINTER(_P), _P, C.POINTER(RendererConfig)]),
    "vrt_renderer_destroy": (None, [_P]),
    "vrt_renderer_last_error": (C.c_char_p, [_P]),
    "vrt_renderer_camera": (_P, [_P]),
    "vrt_renderer_sun": (_P, [_P]),
    "vrt_renderer_ctx": (_P, [_P]),
    "vrt_renderer_push_materials": (C.c_int, [_P, _P, _SZ]),
    "vrt_renderer_update_grid_delta": (C.c_int, [_P]),
    "vrt_renderer_update_sun": (None, [_P, C.c_float]),
    "vrt_renderer_draw": (C.c_int, [_P]),
    "vrt_renderer_draw_to_host": (C.c_int, [_P, _P, _SZ]),
    "vrt_renderer_present_to_host": (C.c_int, [_P, C.POINTER(DenoiseParams), C.c_uint32, C.c_uint32, C.c_uint32, _P, _SZ]),
    "vrt_scene_terrain_materials": (C.c_uint32, [_P, C.c_uint32]),
    "vrt_scene_synthetic": (C.c_int, [C.c_uint32, C.c_uint32, _P, _P]),
    "vrt_scene_synthetic_box": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P]),
    "vrt_scene_synthetic_fill": (C.c_int, [_P, C.c_uint32]),
    "vrt_vox_validate_header": (C.c_int, [_P, _SZ]),
    "vrt_vox_parse": (C.c_int, [C.POINTER(_P), _P, _SZ, C.c_int]),
    "vrt_vox_load": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int]),
    "vrt_vox_destroy": (None, [_P]),
    "vrt_vox_num_models": (C.c_int32, [_P]),
    "vrt_vox_model_size": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32 * 3)]),
    "vrt_vox_model_xyzi": (_P, [_P, C.c_int32, C.POINTER(C.c_uint64)]),
    "vrt_vox_palette": (_P, [_P]),
    "vrt_vox_materials": (C.c_uint32, [_P, _P, C.c_uint32, C.c_uint32]),
    "vrt_vox_insert_into_grid": (C.c_int, [_P, C.c_int32, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "vrt_bench_path_pose": (None, [C.c_float, C.c_float, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 4)]),
    "vrt_benchmark_create": (_P, [_P, _P, C.c_int, C.c_float, C.c_float]),
    "vrt_benchmark_destroy": (None, [_P]),
    "vrt_benchmark_update": (C.c_int, [_P, C.c_float]),
    "vrt_benchmark_get_report": (None, [_P, C.POINTER(BenchmarkReport)]),
    "vrt_renderer_run_benchmark": (C.c_int, [_P, C.c_float, C.c_float, C.POINTER(BenchmarkReport)]),
}


def _load(name: str, symbols: dict) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make` (or __graft_entry__.build()). There is no fallback path.")
    dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for sym, (res, args) in symbols.items():
        fn = getattr(dll, sym)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return dll


_lib = None
_host = None


def lib() -> C.CDLL:
    """libvrt.so (the C ABI of include/vrt.h)."""
    global _lib
    if _lib is None:
        _lib = _load("libvrt.so", VRT_SYMBOLS)
    return _lib


def host_lib() -> C.CDLL:
    """libvrt_host.so (include/vrt_host.h); depends on libvrt.so."""
    global _host
    if _host is None:
        lib()
        _host = _load("libvrt_host.so", VRT_HOST_SYMBOLS)
    return _host


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def terrain_materials(capacity: int = 256) -> np.ndarray:
    """The 8 terrain materials (terrain/terrain.zig:130-196) padded with zeros to `capacity` entries."""
    out = np.zeros(capacity, dtype=MATERIAL_DTYPE)
    host_lib().vrt_scene_terrain_materials(_ptr(out), capacity)
    return out


def bench_path_pose(t: float, extent_scale: float = 1.0):
    o = (C.c_float * 3)()
    q = (C.c_float * 4)()
    host_lib().vrt_bench_path_pose(t, extent_scale, C.byref(o), C.byref(q))
    return list(o), list(q)


class Benchmark:
    """Benchmark (voxel_rt/Benchmark.zig): moves a HostCamera along the reference's fly-through by accumulated frame time."""

    def __init__(self, camera: "HostCamera", grid: "Grid | None" = None, sun_enabled: bool = True, duration_s: float = 60.0, extent_scale: float = 1.0):
        self._h = host_lib()
        self.camera = camera
        self.handle = self._h.vrt_benchmark_create(camera.handle, grid.handle if grid is not None else None, int(sun_enabled), duration_s, extent_scale)
        if not self.handle:
            raise ValueError("vrt_benchmark_create rejected its arguments")

    def update(self, dt: float) -> bool:
        return bool(self._h.vrt_benchmark_update(self.handle, dt))

    @property
    def report(self) -> BenchmarkReport:
        r = BenchmarkReport()
        self._h.vrt_benchmark_get_report(self.handle, C.byref(r))
        return r

    def close(self):
        if self.handle:
            self._h.vrt_benchmark_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


VOX_ERRORS = {-10: "InvalidId", -11: "ExpectedSizeHeader", -12: "ExpectedXyziHeader", -13: "ExpectedRgbaHeader", -14: "UnexpectedVersion",
              -15: "InvalidFileContent", -16: "IoError"}


class Vox:
    """A parsed MagicaVoxel file (vox/loader.zig, vox/types.zig)."""

    def __init__(self, data: bytes | None = None, path: str | None = None, strict: bool = True):
        h = host_lib()
        self._h = h
        handle = C.c_void_p()
        if path is not None:
            rc = h.vrt_vox_load(C.byref(handle), path.encode(), 1 if strict else 0)
        else:
            buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
            rc = h.vrt_vox_parse(C.byref(handle), buf, len(data), 1 if strict else 0)
        if rc != 0:
            raise VrtError(rc, "vox: " + VOX_ERRORS.get(rc, str(rc)))
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                self._h.vrt_vox_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def num_models(self) -> int:
        return self._h.vrt_vox_num_models(self.handle)

    def size(self, model=0):
        s = (C.c_int32 * 3)()
        self._h.vrt_vox_model_size(self.handle, model, C.byref(s))
        return tuple(s)

    def xyzi(self, model=0) -> np.ndarray:
        n = C.c_uint64()
        p = self._h.vrt_vox_model_xyzi(self.handle, model, C.byref(n))
        if not n.value:
            return np.zeros((0, 4), dtype=np.uint8)
        return np.frombuffer((C.c_uint8 * (n.value * 4)).from_address(p), dtype=np.uint8).reshape(-1, 4).copy()

    @property
    def palette(self) -> np.ndarray:
        return np.frombuffer((C.c_uint8 * 1024).from_address(self._h.vrt_vox_palette(self.handle)), dtype=np.uint8).reshape(256, 4).copy()

    def materials(self, materials: np.ndarray, material_base: int) -> int:
        """main.zig:96-108: palette -> materials[material_base:], in place."""
        assert materials.dtype == MATERIAL_DTYPE and materials.flags["C_CONTIGUOUS"]
        return self._h.vrt_vox_materials(self.handle, _ptr(materials), materials.shape[0], material_base)

    def insert_into(self, grid: "Grid", offset=(0, 0, 0), material_base=0, model=0) -> int:
        """main.zig:110-118."""
        return self._h.vrt_vox_insert_into_grid(self.handle, model, grid.handle, offset[0], offset[1], offset[2], material_base)


class Grid:
    """BrickGrid (brick/Grid.zig)."""

    def __init__(self, dim, brick_dim=4, brick_alloc=0, min_point=(0.0, 0.0, 0.0), scale=1.0, base_t=0.01):
        h = host_lib()
        self._h = h
        self.handle = h.vrt_grid_create(dim[0], dim[1], dim[2], brick_dim, brick_alloc, C.byref(_f3(min_point)), scale, base_t)
        if not self.handle:
            raise VrtError(-1, f"vrt_grid_create({dim}, brick_dim={brick_dim}) failed")

    def close(self):
        if self.handle:
            self._h.vrt_grid_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert(self, x, y, z, material) -> int:
        return self._h.vrt_grid_insert(self.handle, x, y, z, material)

    def insert_many(self, xyzm: np.ndarray) -> int:
        a = np.ascontiguousarray(xyzm, dtype=np.uint32).reshape(-1, 4)
        return self._h.vrt_grid_insert_many(self.handle, _ptr(a), a.shape[0])

    def fill_synthetic(self, seed=420) -> int:
        return self._h.vrt_scene_synthetic_fill(self.handle, seed)

    @property
    def state(self) -> GridState:
        s = GridState()
        self._h.vrt_grid_get_state(self.handle, C.byref(s))
        return s

    @property
    def brick_dim(self) -> int:
        return self._h.vrt_grid_brick_dim(self.handle)

    @property
    def brick_alloc(self) -> int:
        return self._h.vrt_grid_brick_alloc(self.handle)

    @property
    def active_bricks(self) -> int:
        return self._h.vrt_grid_active_bricks(self.handle)

    def _array(self, fn, dtype) -> np.ndarray:
        n = C.c_uint64()
        p = fn(self.handle, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_uint8 * (n.value * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype)  # view into the grid's memory; valid while the grid lives

    @property
    def statuses(self):
        return self._array(self._h.vrt_grid_statuses, np.uint32)

    @property
    def brick_indices(self):
        return self._array(self._h.vrt_grid_brick_indices, np.uint32)

    @property
    def occupancy(self):
        return self._array(self._h.vrt_grid_occupancy, np.uint8)

    @property
    def start_indices(self):
        return self._array(self._h.vrt_grid_start_indices, np.uint32)

    @property
    def material_indices(self):
        return self._array(self._h.vrt_grid_material_indices, np.uint8)

    def delta(self, which: int):
        a, b = C.c_uint64(), C.c_uint64()
        rc = self._h.vrt_grid_delta_peek(self.handle, which, C.byref(a), C.byref(b))
        return rc, a.value, b.value

    def delta_reset(self, which: int):
        self._h.vrt_grid_delta_reset(self.handle, which)


class HostCamera:
    """Camera (voxel_rt/Camera.zig)."""

    def __init__(self, fov_deg, width, height, origin=(0.0, 0.0, 0.0), samples_per_pixel=2, max_bounce=2, handle=None):
        self._h = host_lib()
        self._owned = handle is None
        if handle is None:
            cfg = HcamConfig()
            self._h.vrt_hcam_default_config(C.byref(cfg))
            cfg.origin[:] = [float(v) for v in origin]
            cfg.samples_per_pixel = samples_per_pixel
            cfg.max_bounce = max_bounce
            handle = self._h.vrt_hcam_create(fov_deg, width, height, C.byref(cfg))
            if not handle:
                raise VrtError(-1, "vrt_hcam_create failed")
        self.handle = handle

    def __del__(self):
        try:
            if self._owned and self.handle:
                self._h.vrt_hcam_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def device(self) -> CameraDevice:
        d = CameraDevice()
        self._h.vrt_hcam_device(self.handle, C.byref(d))
        return d

    def set_origin(self, origin):
        self._h.vrt_hcam_set_origin(self.handle, C.byref(_f3(origin)))

    def translate(self, dt, by):
        self._h.vrt_hcam_translate(self.handle, dt, C.byref(_f3(by)))

    def turn_pitch(self, angle):
        self._h.vrt_hcam_turn_pitch(self.handle, angle)

    def turn_yaw(self, angle):
        self._h.vrt_hcam_turn_yaw(self.handle, angle)

    def reset(self):
        self._h.vrt_hcam_reset(self.handle)

    def set_orientation(self, yaw_wxyz, pitch_wxyz=None):
        y = (C.c_float * 4)(*[float(v) for v in yaw_wxyz])
        p = (C.c_float * 4)(*[float(v) for v in pitch_wxyz]) if pitch_wxyz is not None else None
        self._h.vrt_hcam_set_orientation(self.handle, C.byref(y), C.byref(p) if p is not None else None)

    def set_euler_deg(self, x, y, z):
        self._h.vrt_hcam_set_euler_deg(self.handle, x, y, z)


class HostSun:
    """Sun (voxel_rt/Sun.zig)."""

    def __init__(self, enabled=True, radius=5.0, animate=True, color=(1.0, 1.1, 1.0), sun_distance=1000.0, animate_speed=0.1, handle=None):
        self._h = host_lib()
        self._owned = handle is None
        if handle is None:
            cfg = HsunConfig()
            self._h.vrt_hsun_default_config(C.byref(cfg))
            cfg.enabled = 1 if enabled else 0
            cfg.radius = radius
            cfg.animate = 1 if animate else 0
            cfg.color[:] = [float(v) for v in color]
            cfg.sun_distance = sun_distance
            cfg.animate_speed = animate_speed
            handle = self._h.vrt_hsun_create(C.byref(cfg))
        self.handle = handle

    def __del__(self):
        try:
            if self._owned and self.handle:
                self._h.vrt_hsun_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def device(self) -> SunDevice:
        d = SunDevice()
        self._h.vrt_hsun_device(self.handle, C.byref(d))
        return d

    def update(self, dt):
        self._h.vrt_hsun_update(self.handle, dt)


class Context:
    """One vrt_ctx: ComputePipeline + its buffers + the target image (include/vrt.h)."""

    def __init__(self, width, height, n_bricks, brick_dim=4, n_brick_alloc=0, material_capacity=256, device=0, flags=0, rows=(0, 0), part=None, handle=None):
        self._l = lib()
        self.width, self.height = width, height
        self.n_bricks, self.brick_dim, self.n_brick_alloc = n_bricks, brick_dim, n_brick_alloc or n_bricks
        self._owned = handle is None
        if handle is None:
            if part is not None:  # (rank, world): interleaved 4-row strips
                flags |= VRT_FLAG_INTERLEAVE
            cfg = Config(C.sizeof(Config), VRT_ABI_VERSION, width, height, brick_dim, material_capacity, n_bricks, n_brick_alloc, device, flags, rows[0], rows[1],
                         part[0] if part else 0, part[1] if part else 0)
            h = C.c_void_p()
            rc = self._l.vrt_init(C.byref(h), C.byref(cfg))
            if rc != 0:
                raise VrtError(rc, self._l.vrt_last_error(None).decode())
            handle = h
        self.handle = handle

    def close(self):
        if self._owned and self.handle:
            self._l.vrt_deinit(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VrtError(rc, self._l.vrt_last_error(self.handle).decode())

    def upload_grid_state(self, state: GridState):
        self._check(self._l.vrt_upload_grid_state(self.handle, C.byref(state)))
        self._grid_dims = (state.dim_x, state.dim_y, state.dim_z)

    def upload_grid_delta(self, grid: "Grid") -> int:
        """VoxelRT.updateGridDelta (VoxelRT.zig:107-172): ship the five dirty ranges of the host grid.  Returns the bytes sent."""
        up = [self.upload_brick_statuses, self.upload_brick_indices, self.upload_brick_occupancy, self.upload_brick_start_indices, self.upload_material_indices]
        arr = [grid.statuses, grid.brick_indices, grid.occupancy, grid.start_indices, grid.material_indices]
        sent = 0
        for which in range(5):
            active, lo, hi = grid.delta(which)
            if active:
                up[which](lo, arr[which][lo:hi])
                sent += int(arr[which][lo:hi].nbytes)
                grid.delta_reset(which)
        return sent

    def _upload(self, fn, offset, data: np.ndarray, dtype):
        a = np.ascontiguousarray(data, dtype=dtype)
        self._check(fn(self.handle, offset, _ptr(a), a.shape[0]))

    def upload_materials(self, offset, data):
        self._upload(self._l.vrt_upload_materials, offset, data, MATERIAL_DTYPE)

    def upload_brick_statuses(self, offset, data):
        self._upload(self._l.vrt_upload_brick_statuses, offset, data, np.uint32)

    def upload_brick_indices(self, offset, data):
        self._upload(self._l.vrt_upload_brick_indices, offset, data, np.uint32)

    def upload_brick_occupancy(self, offset, data):
        self._upload(self._l.vrt_upload_brick_occupancy, offset, data, np.uint8)

    def upload_brick_start_indices(self, offset, data):
        self._upload(self._l.vrt_upload_brick_start_indices, offset, data, np.uint32)

    def upload_material_indices(self, offset, data):
        self._upload(self._l.vrt_upload_material_indices, offset, data, np.uint8)

    def upload_grid(self, grid: Grid, materials: np.ndarray):
        """transferGridState + one full transfer of every array (what the first updateGridDelta amounts to)."""
        self.upload_grid_state(grid.state)
        self.upload_materials(0, materials)
        self.upload_brick_statuses(0, grid.statuses)
        self.upload_brick_indices(0, grid.brick_indices)
        self.upload_brick_occupancy(0, grid.occupancy)
        self.upload_brick_start_indices(0, grid.start_indices)
        self.upload_material_indices(0, grid.material_indices)

    def trace(self, camera: CameraDevice, sun: SunDevice):
        self._check(self._l.vrt_trace(self.handle, C.byref(camera), C.byref(sun)))

    def sync(self):
        self._check(self._l.vrt_sync(self.handle))

    def read_framebuffer(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(self._l.vrt_read_framebuffer(self.handle, _ptr(out), out.nbytes))
        return out

    def insert_voxels(self, xyzm: np.ndarray, active_bricks: int) -> int:
        """BrickGrid.insert for a batch of {x, y, z, material} uint32 quadruples on the device; returns the new active brick count."""
        xyzm = np.ascontiguousarray(xyzm, dtype=np.uint32).reshape(-1, 4)
        active = C.c_uint32(active_bricks)
        self._check(self._l.vrt_insert_voxels(self.handle, _ptr(xyzm), len(xyzm), C.byref(active)))
        return active.value

    def download_buffer(self, which: int) -> np.ndarray:
        """One of the five grid buffers (binding number 3..7) as the host would have uploaded it."""
        n, dt = {3: (self.n_bricks + 31) // 32, 4: self.n_bricks}.get(which), np.uint32
        bits = self.brick_dim ** 3
        alloc = self.n_brick_alloc
        if which == 5:
            n, dt = alloc * bits // 8, np.uint8
        elif which == 6:
            n = alloc
        elif which == 7:
            n, dt = alloc * bits, np.uint8
        out = np.empty(n, dtype=dt)
        self._check(self._l.vrt_download_buffer(self.handle, which, 0, _ptr(out), n))
        return out

    def denoise(self, params: "DenoiseParams | None" = None, out_width: int | None = None, out_height: int | None = None, flags: int = 0) -> np.ndarray:
        """image.frag over the framebuffer last traced into; returns the (out_height, out_width, 4) uint8 result."""
        params = params or DenoiseParams.default()
        ow, oh = out_width or self.width, out_height or self.height
        self._check(self._l.vrt_denoise(self.handle, C.byref(params), ow, oh, flags))
        out = np.empty((oh, ow, 4), dtype=np.uint8)
        self._check(self._l.vrt_read_denoised(self.handle, _ptr(out), out.nbytes))
        return out

    def last_denoise_ms(self) -> float:
        ms = C.c_float()
        self._check(self._l.vrt_last_denoise_ms(self.handle, C.byref(ms)))
        return ms.value

    def trace_to_host(self, camera: CameraDevice, sun: SunDevice, out: np.ndarray | None = None, out_ptr: int | None = None) -> np.ndarray | None:
        nbytes = self.width * self.height * 4
        if out_ptr is not None:
            self._check(self._l.vrt_trace_to_host(self.handle, C.byref(camera), C.byref(sun), C.c_void_p(out_ptr), nbytes))
            return None
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(self._l.vrt_trace_to_host(self.handle, C.byref(camera), C.byref(sun), _ptr(out), nbytes))
        return out

    def trace_to_host_async(self, camera: CameraDevice, sun: SunDevice, out_ptr: int | None):
        """Pipelined frame: returns after enqueueing; `out_ptr` (pinned host memory, width*height*4 bytes) is valid after sync()."""
        self._check(self._l.vrt_trace_to_host_async(self.handle, C.byref(camera), C.byref(sun), C.c_void_p(out_ptr) if out_ptr else None,
                                                    self.width * self.height * 4))

    def trace_rays(self, origins: np.ndarray, directions: np.ndarray) -> np.ndarray:
        """Explicit-ray mode (vrt_trace_rays_host): n x 3 origins and directions in, n hit records out."""
        n = len(origins)
        rays = np.zeros(n, dtype=RAY_DTYPE)
        rays["origin"], rays["direction"] = origins, directions
        hits = np.zeros(n, dtype=RAY_HIT_DTYPE)
        self._check(self._l.vrt_trace_rays_host(self.handle, _ptr(rays), _ptr(hits), n))
        return hits

    def trace_rays_device(self, rays_ptr: int, hits_ptr: int, count: int):
        self._check(self._l.vrt_trace_rays(self.handle, C.c_void_p(rays_ptr), C.c_void_p(hits_ptr), count))

    def read_aov(self) -> np.ndarray:
        out = np.empty(self.height * self.width, dtype=AOV_DTYPE)
        self._check(self._l.vrt_read_aov(self.handle, _ptr(out), out.shape[0]))
        return out.reshape(self.height, self.width)

    def counters(self) -> dict:
        c = Counters()
        self._check(self._l.vrt_get_counters(self.handle, C.byref(c)))
        return c.as_dict()

    def last_trace_ms(self) -> float:
        ms = C.c_float()
        self._check(self._l.vrt_last_trace_ms(self.handle, C.byref(ms)))
        return ms.value

    def debug_dist_planes(self) -> np.ndarray:
        """The derived distance planes (8 octants x padded grid bytes), up to date with the uploads so far.  Tests only."""
        n = C.c_size_t(0)
        # capacity is 8 * plane bytes: ask for everything by probing with the grid's padded size
        g = self._grid_dims
        lx = 1
        while (1 << lx) < g[0] + 2:
            lx += 1
        lz = 1
        while (1 << lz) < g[2] + 2:
            lz += 1
        count = 8 * ((g[1] + 2) << (lx + lz))
        out = np.empty(count, dtype=np.uint8)
        self._check(self._l.vrt_download_buffer(self.handle, 100, 0, _ptr(out), count))
        return out.reshape(8, g[1] + 2, 1 << lz, 1 << lx)

    def debug_force_accel_rebuild(self):
        self._check(self._l.vrt_debug_force_accel_rebuild(self.handle))

    def last_trace_kernel_ms(self) -> float:
        """The part of last_trace_ms() before the exchange: (rebuild +) trace kernel."""
        ms = C.c_float()
        self._check(self._l.vrt_last_trace_kernel_ms(self.handle, C.byref(ms)))
        return ms.value

    def set_schedule(self, mode: int, interval: int = 0):
        self._check(self._l.vrt_set_schedule(self.handle, mode, interval))

    @property
    def n_tiles(self) -> int:
        return ((self.width + 7) // 8) * ((self.height + 3) // 4)

    def sched_costs(self, count: int | None = None) -> np.ndarray:
        out = np.zeros(count or self.n_tiles, dtype=np.uint16)
        self._check(self._l.vrt_sched_get_costs(self.handle, _ptr(out), out.shape[0]))
        return out

    def sched_set_costs(self, costs: np.ndarray):
        a = np.ascontiguousarray(costs, dtype=np.uint16)
        self._check(self._l.vrt_sched_set_costs(self.handle, _ptr(a), a.shape[0]))

    def last_trace_launches(self) -> int:
        n = C.c_uint32()
        self._check(self._l.vrt_last_trace_launches(self.handle, C.byref(n)))
        return n.value

    def set_stream(self, cuda_stream: int | None):
        self._check(self._l.vrt_set_stream(self.handle, C.c_void_p(cuda_stream) if cuda_stream else None))

    def attach_framebuffer(self, device_ptr: int | None, nbytes: int = 0):
        self._check(self._l.vrt_attach_framebuffer(self.handle, C.c_void_p(device_ptr) if device_ptr else None, nbytes))

    def framebuffer_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._l.vrt_framebuffer_device_ptr(self.handle, C.byref(p)))
        return p.value

    # multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * VRT_NCCL_ID_BYTES)()
        rc = lib().vrt_comm_get_unique_id(buf)
        if rc != 0:
            raise VrtError(rc, lib().vrt_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, rank, world, unique_id: bytes):
        buf = (C.c_uint8 * VRT_NCCL_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._l.vrt_comm_init(self.handle, rank, world, buf))

    def comm_ipc_handle(self) -> bytes:
        buf = (C.c_uint8 * VRT_IPC_HANDLE_BYTES)()
        self._check(self._l.vrt_comm_get_ipc_handle(self.handle, buf))
        return bytes(buf)

    def comm_open_peers(self, rank, world, handles: bytes):
        buf = (C.c_uint8 * (VRT_IPC_HANDLE_BYTES * world)).from_buffer_copy(handles)
        self._check(self._l.vrt_comm_open_peers(self.handle, rank, world, buf))

    def comm_set_exchange(self, mode: int):
        self._check(self._l.vrt_comm_set_exchange(self.handle, mode))


class Renderer:
    """VoxelRT facade (src/modules/VoxelRT.zig): init / pushMaterials / updateGridDelta / updateSun / draw."""

    def __init__(self, grid: Grid, width=1280, height=720, samples_per_pixel=2, max_bounce=2, origin=(0.0, 0.0, 0.0),
                 sun_enabled=True, sun_radius=5.0, sun_animate=True, device=0, flags=0, rows=(0, 0)):
        h = host_lib()
        self._h = h
        self.grid = grid
        self.width, self.height = width, height
        cfg = RendererConfig()
        h.vrt_renderer_default_config(C.byref(cfg))
        cfg.internal_resolution_width, cfg.internal_resolution_height = width, height
        cfg.camera.samples_per_pixel, cfg.camera.max_bounce = samples_per_pixel, max_bounce
        cfg.camera.origin[:] = [float(v) for v in origin]
        cfg.sun.enabled = 1 if sun_enabled else 0
        cfg.sun.radius = sun_radius
        cfg.sun.animate = 1 if sun_animate else 0
        cfg.device, cfg.flags = device, flags
        cfg.row_begin, cfg.row_end = rows
        handle = C.c_void_p()
        rc = h.vrt_renderer_create(C.byref(handle), grid.handle, C.byref(cfg))
        if rc != 0:
            raise VrtError(rc, h.vrt_renderer_last_error(None).decode())
        self.handle = handle
        self.camera = HostCamera(0, 0, 0, handle=h.vrt_renderer_camera(handle))
        self.sun = HostSun(handle=h.vrt_renderer_sun(handle))
        self.ctx = Context(width, height, 0, handle=C.c_void_p(h.vrt_renderer_ctx(handle)))

    def close(self):
        if self.handle:
            self.ctx.handle = None
            self._h.vrt_renderer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VrtError(rc, self._h.vrt_renderer_last_error(self.handle).decode())

    def push_materials(self, materials: np.ndarray):
        a = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        self._check(self._h.vrt_renderer_push_materials(self.handle, _ptr(a), a.shape[0]))

    def update_grid_delta(self):
        self._check(self._h.vrt_renderer_update_grid_delta(self.handle))

    def update_sun(self, dt):
        self._h.vrt_renderer_update_sun(self.handle, dt)

    def draw(self):
        self._check(self._h.vrt_renderer_draw(self.handle))

    def draw_to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(self._h.vrt_renderer_draw_to_host(self.handle, _ptr(out), out.nbytes))
        return out

    def run_benchmark(self, duration_s: float = 60.0, extent_scale: float = 1.0) -> BenchmarkReport:
        """The reference's benchmark mode: frames along the fly-through until `duration_s` of frame time has accumulated."""
        rep = BenchmarkReport()
        self._check(self._h.vrt_renderer_run_benchmark(self.handle, duration_s, extent_scale, C.byref(rep)))
        return rep

    def present_to_host(self, out_width: int | None = None, out_height: int | None = None, params: "DenoiseParams | None" = None, flags: int = 0) -> np.ndarray:
        """Pipeline.draw's graphics half: draw, then image.frag into an (out_height, out_width, 4) image."""
        ow, oh = out_width or self.width, out_height or self.height
        out = np.empty((oh, ow, 4), dtype=np.uint8)
        self._check(self._h.vrt_renderer_present_to_host(self.handle, C.byref(params) if params is not None else None, ow, oh, flags, _ptr(out), out.nbytes))
        return out
