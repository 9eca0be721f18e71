"""ctypes binding of libtrt_b200.so — exactly the symbols include/trt_b200.h declares.

There is no Python or CPU implementation behind this module: if the shared library has not been
built (python -m terminalraytracer_b200.build) loading fails loudly, and trt_init() exits the process
when no CUDA device is usable."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# TRT_B200_LIB selects an experiment build of the same library (scripts/ only); the product is libtrt_b200.so
LIB_PATH = os.path.join(_HERE, os.environ.get("TRT_B200_LIB", "libtrt_b200.so"))

# trt_frame_sink (include/trt_b200.h)
FRAME_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)

# trt_frame_acquire (include/trt_b200.h)
FRAME_ACQUIRE = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_size_t, C.c_void_p)

# name -> (restype, argtypes); mirrors include/trt_b200.h one to one
SIGNATURES = {
    "trt_init": (C.c_int, [C.c_int]),
    "trt_shutdown": (None, []),
    "trt_is_initialized": (C.c_int, []),
    "trt_stream": (C.c_void_p, []),
    "trt_set_stream": (C.c_int, [C.c_void_p]),
    "trt_use_own_stream": (C.c_int, []),
    "trt_set_cull": (C.c_int, [C.c_int]),
    "trt_upload_skybox": (C.c_int, [C.POINTER(abi.Skybox)]),
    "trt_project_scene": (None, [C.POINTER(abi.Scene), C.POINTER(abi.Screen)]),
    "trt_draw_screen": (C.c_size_t, [C.POINTER(abi.Screen), C.c_void_p]),
    "trt_buffered_draw_screen": (None, [C.POINTER(abi.Screen)]),
    "trt_render_ansi": (C.c_size_t, [C.POINTER(abi.Scene), C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "trt_render_orbit": (C.c_int, [C.POINTER(abi.Scene), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "trt_render_orbit_to": (C.c_int, [C.POINTER(abi.Scene), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "trt_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "trt_ipc_import": (C.c_void_p, [C.c_void_p]),
    "trt_ipc_close": (C.c_int, [C.c_void_p]),
    "trt_push_to_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "trt_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "trt_host_unregister": (C.c_int, [C.c_void_p]),
    "trt_peer_copies_wait": (C.c_int, []),
    "trt_signal_step": (C.c_int, [C.c_void_p, C.c_uint, C.c_int]),
    "trt_wait_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.c_int]),
    "trt_stream_wait_copies": (C.c_int, []),
    "trt_set_scene": (C.c_int, [C.POINTER(abi.Scene)]),
    "trt_set_scene_async": (C.c_int, [C.POINTER(abi.Scene)]),
    "trt_render_rows_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "trt_encode_rows_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "trt_render_rows_quant_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "trt_encode_rows_quant_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "trt_render_rows_ansi_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "trt_stream_frame_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "trt_estimate_row_costs": (C.c_int, [C.POINTER(abi.Scene), C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "trt_count_rows_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_longlong)]),
    "trt_model_flops": (C.c_double, [C.POINTER(C.c_longlong)]),
    "trt_probe_trace_ray": (C.c_int, [C.POINTER(abi.Scene), C.c_void_p, C.c_int, C.c_void_p]),
    "trt_probe_skybox": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "trt_probe_sphere": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "trt_probe_plane": (C.c_int, [C.POINTER(abi.Scene), C.c_void_p, C.c_int, C.c_void_p]),
    "trt_probe_lighting": (C.c_int, [C.POINTER(abi.Scene), C.c_void_p, C.c_int, C.c_void_p]),
    "trt_selftest_division": (C.c_longlong, [C.c_ulonglong, C.c_longlong]),
    "trt_device_alloc": (C.c_void_p, [C.c_size_t]),
    "trt_device_free": (None, [C.c_void_p]),
    "trt_host_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "trt_host_free_pinned": (None, [C.c_void_p]),
    "trt_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "trt_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "trt_synchronize": (C.c_int, []),
    "trt_debug_bounds": (C.c_int, [C.c_void_p]),
    "trt_last_render_ms": (C.c_float, []),
    "trt_last_encode_ms": (C.c_float, []),
    "trt_measure_fp32_tflops": (C.c_double, []),
    "trt_measure_fp64_tflops": (C.c_double, []),
    "trt_init_camera": (None, [C.POINTER(abi.Camera), C.c_int, C.c_int]),
    "trt_orbit_camera": (None, [C.POINTER(abi.Camera), C.c_double]),
    "trt_pose_camera": (None, [C.POINTER(abi.Camera), C.c_double, C.c_double, C.c_double]),
    "trt_subpixel_offsets": (None, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "trt_demo_scene": (None, [C.POINTER(abi.Scene), C.POINTER(abi.Sphere), C.POINTER(abi.DirectionalLight),
                              C.POINTER(abi.PointLight), C.c_int, C.c_int]),
    "trt_stress_scene": (C.c_int, [C.POINTER(abi.Scene), C.POINTER(abi.Sphere), C.c_int, C.POINTER(abi.DirectionalLight),
                                   C.POINTER(abi.PointLight), C.c_int, C.c_int]),
    "trt_read_ppm": (None, [C.c_char_p, C.POINTER(C.POINTER(abi.Color)), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "trt_load_skybox": (None, [C.POINTER(abi.Skybox), C.c_char_p]),
    "trt_load_skybox_dir": (None, [C.POINTER(abi.Skybox), C.c_char_p]),
    "trt_free_skybox": (None, [C.POINTER(abi.Skybox)]),
}

_lib = None


def load():
    """dlopen libtrt_b200.so and type every entry point.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m terminalraytracer_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the render path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
