"""Python-side mirror of the reference's per-frame calls, bound to libtrt_b200.so.

    reference (TRT.c)                     here
    project_scene(&scene,&screen) :1339   Renderer.project_scene(scene)      -> (H,W,3) float64
    buffered_draw_screen(&screen) :1342   Renderer.draw_screen(pixels)       -> bytes of the terminal stream
    both, fused                           Renderer.render_ansi(scene)        -> uint8 array (pinned)
    row band of a frame (multi-GPU)       Renderer.render_band_bytes(...)    -> torch uint8 tensor on the GPU

torch is used for device memory and for torch.distributed only; every kernel is ours."""
import ctypes as C

import numpy as np

from . import abi, lib as _lib


class Renderer:
    def __init__(self, device=0, skybox=None):
        self.L = _lib.load()
        self.L.trt_init(int(device))  # exits the process if there is no CUDA device: no fallback
        self.device = int(device)
        self._pinned = None
        self._pinned_cap = 0
        self._skybox = None
        if skybox is not None:
            self.upload_skybox(skybox)

    # ---- skybox -----------------------------------------------------------------------------
    def upload_skybox(self, skybox):
        self.L.trt_upload_skybox(C.byref(skybox.c))
        self._skybox = skybox

    # ---- drop-ins (host buffers in, host buffers out) ------------------------------------------
    def project_scene(self, scene, out=None):
        h, w = scene.height, scene.width
        if out is None:
            out = np.empty((h, w, 3), dtype=np.float64)
        screen = abi.Screen(out.ctypes.data_as(C.POINTER(abi.Vector)), w, h)
        self.L.trt_project_scene(C.byref(scene.c), C.byref(screen))
        return out

    def draw_screen(self, pixels):
        h, w, _ = pixels.shape
        pixels = np.ascontiguousarray(pixels, dtype=np.float64)
        out = np.empty(abi.stream_bytes(w, h), dtype=np.uint8)
        screen = abi.Screen(pixels.ctypes.data_as(C.POINTER(abi.Vector)), w, h)
        n = self.L.trt_draw_screen(C.byref(screen), out.ctypes.data)
        assert n == out.size
        return out

    def _pinned_buffer(self, nbytes):
        if self._pinned_cap < nbytes:
            if self._pinned:
                self.L.trt_host_free_pinned(self._pinned)
            self._pinned = self.L.trt_host_alloc_pinned(nbytes)
            self._pinned_cap = nbytes
        return self._pinned

    def render_ansi(self, scene):
        """One call = upload scene, K1, K2, one D2H of the byte stream into pinned host memory."""
        n = abi.stream_bytes(scene.width, scene.height)
        buf = self._pinned_buffer(n)
        got = self.L.trt_render_ansi(C.byref(scene.c), scene.width, scene.height, buf, n)
        assert got == n
        return np.ctypeslib.as_array((C.c_ubyte * n).from_address(buf))

    def render_orbit(self, scene, times, sink, first=0, stride=1):
        """trt_render_orbit: frames first, first+stride, ... of the camera path `times`, each handed to
        sink(frame_index, uint8 array view — valid only during the call) in order while the next frame renders.
        The scene's camera must be un-posed (as SceneData builds it); returns the number of frames delivered."""
        n = len(times)
        arr = (C.c_double * n)(*[float(t) for t in times])

        array_type = C.c_ubyte * abi.stream_bytes(scene.width, scene.height)

        def _cb(ptr, nbytes, frame, _user):
            view = np.frombuffer(array_type.from_address(ptr), dtype=np.uint8)    # no copy: the library's pinned buffer
            return 1 if sink(frame, view) else 0

        cb = _lib.FRAME_SINK(_cb)
        return self.L.trt_render_orbit(C.byref(scene.c), scene.width, scene.height, arr, n, int(first), int(stride),
                                       C.cast(cb, C.c_void_p), None)

    # ---- device-resident band API ---------------------------------------------------------------
    def set_scene(self, scene):
        self.L.trt_set_scene(C.byref(scene.c))

    def set_scene_async(self, scene):
        """trt_set_scene without the host wait (the upload is ordered on the stream like the kernels that follow it)"""
        self.L.trt_set_scene_async(C.byref(scene.c))

    def render_rows(self, width, height, row0, row1, d_pixels_ptr):
        self.L.trt_render_rows_device(width, height, row0, row1, d_pixels_ptr)

    def render_rows_quant(self, width, height, row0, row1, d_quant_ptr):
        self.L.trt_render_rows_quant_device(width, height, row0, row1, d_quant_ptr)

    def render_rows_ansi(self, width, height, row0, row1, stream_ptr):
        """K1 with the encoder fused in: the rows' terminal bytes are stored at their place in the stream at stream_ptr
        (device, peer or page-locked host memory)."""
        self.L.trt_render_rows_ansi_device(width, height, row0, row1, C.c_void_p(int(stream_ptr)))

    def encode_rows(self, d_pixels_ptr, width, rows, d_bytes_ptr, byte_offset):
        self.L.trt_encode_rows_device(d_pixels_ptr, width, rows, d_bytes_ptr, byte_offset)

    def encode_rows_quant(self, d_quant_ptr, width, rows, d_bytes_ptr, byte_offset):
        self.L.trt_encode_rows_quant_device(d_quant_ptr, width, rows, d_bytes_ptr, byte_offset)

    def stream_frame(self, d_stream_ptr, width, height):
        self.L.trt_stream_frame_device(d_stream_ptr, width, height)

    def estimate_row_costs(self, scene):
        """Per-row cost estimate of scene.width x scene.height (1/8-resolution pre-pass on the GPU)."""
        out = (C.c_double * scene.height)()
        self.L.trt_estimate_row_costs(C.byref(scene.c), scene.width, scene.height, out)
        return list(out)

    def count_rows(self, width, height, row0, row1, d_pixels_ptr=None):
        ctr = (C.c_longlong * abi.NUM_COUNTERS)()
        self.L.trt_count_rows_device(width, height, row0, row1, d_pixels_ptr, ctr)
        return list(ctr), self.L.trt_model_flops(ctr)

    def use_stream(self, cuda_stream_ptr):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream, where 0 is the
        legacy default stream); None = back to the library's own stream."""
        if cuda_stream_ptr is None:
            self.L.trt_use_own_stream()
        else:
            self.L.trt_set_stream(C.c_void_p(int(cuda_stream_ptr)))

    def synchronize(self):
        self.L.trt_synchronize()

    def last_ms(self):
        return float(self.L.trt_last_render_ms()), float(self.L.trt_last_encode_ms())

    def measure_peaks(self):
        return {"fp32_tflops": float(self.L.trt_measure_fp32_tflops()),
                "fp64_tflops": float(self.L.trt_measure_fp64_tflops())}

    def close(self):
        if self._pinned:
            self.L.trt_host_free_pinned(self._pinned)
            self._pinned = None
            self._pinned_cap = 0
        self.L.trt_shutdown()
