/*
 * trt_b200.h — C ABI of libtrt_b200.so, the B200 (sm_100a) implementation of TerminalRayTracer's
 * per-pixel render path.  Plain C: pointers, ints and sizes only; no CUDA or torch types.
 *
 * The reference has no plugin / FFI layer; its "operator interface" for this path is the set of
 * plain C functions its main() calls (TRT.c = /root/reference/TerminalRayTracer.c):
 *
 *      reference call site                      replacement exported here
 *      ---------------------------------------  -------------------------------------------------
 *      initialize_screenbuffer()   TRT.c:1241   trt_init()               (device context + staging)
 *      load_skybox(&sky, name)     TRT.c:1244   trt_load_skybox() then trt_upload_skybox()
 *      project_scene(&scene,&scr)  TRT.c:1339   trt_project_scene()      (same signature, same pixels)
 *      buffered_draw_screen(&scr)  TRT.c:1342   trt_buffered_draw_screen() / trt_draw_screen()
 *      1339 + 1342 together                     trt_render_ansi()        (fused: one D2H, one fwrite)
 *      free_skybox(&sky)           TRT.c:1369   trt_free_skybox(), trt_shutdown()
 *
 * Error behaviour follows the reference (TRT.c:318-322 …): the drop-in calls cannot fail from
 * the caller's point of view; a CUDA failure prints "file:line: message" to stderr and exit(1)s.
 * There is NO CPU fallback: without a usable CUDA device trt_init() exits.
 * Threading follows the reference too: one caller thread, not re-entrant.
 *
 * INTEGRATION.md shows the edit a maintainer makes to the reference's main() to bind these.
 */
#ifndef TRT_B200_H
#define TRT_B200_H

#include <stddef.h>
#include "trt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TRT_DEMO_SPHERES 6
/* LIMIT (differs from the reference, which takes any count): at most TRT_MAX_LIGHTS directional and TRT_MAX_LIGHTS point
 * lights per scene — they live in the kernel's constant block.  A scene with more makes every entry point that takes a
 * scene print "unsupported scene" to stderr and exit(1), in the reference's die-on-error style.  INTEGRATION.md repeats this
 * next to the swap instructions. */
#define TRT_MAX_LIGHTS 16        /* per kind; the demo scene uses 1 + 1 (TRT.c:1278-1287) */

/* ---- lifecycle ---------------------------------------------------------------------------- */
/* Bind this process to CUDA device `device` (one process per GPU; under torchrun pass LOCAL_RANK).
 * Creates the library's stream and staging buffers.  Exits if no CUDA device is usable. */
int trt_init(int device);
void trt_shutdown(void);
/* 1 if trt_init() has succeeded in this process */
int trt_is_initialized(void);
/* the library's CUDA stream as an opaque pointer (cudaStream_t), so plumbing code (torch) can order against it */
void *trt_stream(void);
/* Run all subsequent work on the caller's CUDA stream (a cudaStream_t; e.g. torch's current stream, so that
 * the caller's events and collectives are ordered with the kernels).  NULL means the legacy default stream,
 * which is what torch uses unless told otherwise.  trt_use_own_stream() goes back to the library's stream. */
int trt_set_stream(void *cuda_stream);
int trt_use_own_stream(void);

/* The render kernel skips the FP64 sphere test whenever a conservative FP32 test certifies the
 * reference's `discriminant < 0` outcome (DESIGN.md "FP32 cull"; results are bit-identical either way).
 * trt_set_cull(0) forces the all-FP64 path — for A/B measurements and for the parity tests of that path.
 * Takes effect at the next scene upload. */
int trt_set_cull(int enabled);

/* ---- skybox ingest (replaces the pointer chase through Scene.skybox, TRT.c:782-788) ----------- */
/* Copies the six dim*dim RGB planes to the device (each padded with dim+1 black texels, see
 * trt_host.c).  Must be called before the first render and whenever the skybox changes.
 * The caller's planes are not retained. */
int trt_upload_skybox(const trt_Skybox *skybox);

/* ---- drop-ins ------------------------------------------------------------------------------ */
/* project_scene (TRT.c:966): fills screen->pixels[row*width+col] with the same FP64 RGB the
 * reference computes, bit for bit.  scene->skybox is ignored in favour of the uploaded skybox
 * only as far as the texel storage goes (dim must match). */
void trt_project_scene(const trt_Scene *scene, trt_Screen *screen);

/* buffered_draw_screen (TRT.c:1142) without the fwrite: writes the 9+(25W+1)H bytes of the
 * terminal stream for `screen` to `out` and returns the byte count. */
size_t trt_draw_screen(const trt_Screen *screen, char *out);

/* exact drop-in for buffered_draw_screen: encodes on the GPU and fwrite()s the stream to stdout */
void trt_buffered_draw_screen(const trt_Screen *screen);

/* fused path: render w x h and encode on the device, copy only the byte stream back.
 * `out` needs TRT_STREAM_BYTES(w,h) bytes; returns the byte count (0 if cap is too small).
 * PRECONDITION of every path that carries QUANTISED cells between the kernels (this call, trt_render_orbit*, the *_quant_ and
 * *_ansi_ device calls): pixels in [0,1], which every scene with finite lights, materials in [0,1] and a loaded skybox gives
 * (TRT.c:960 clamps the lit colour, 1061 normalises by the weight sum).  A NaN/Inf light or material makes (int)(c*255) leave
 * 0..255; the reference feeds that int to byte_to_digits (TRT.c:1134-1139) and so does trt_project_scene +
 * trt_draw_screen, whereas the quantised paths keep its low byte: different digits for such pixels only. */
size_t trt_render_ansi(const trt_Scene *scene, int width, int height, char *out, size_t cap);

/* Streaming sink for camera animations — the reference's frame loop (TRT.c:1317-1367) with both hot calls on the GPU.
 * For every frame k in {first, first+stride, ...} below n_frames the camera is posed as the reference poses it at
 * wall-clock time times[k] (TRT.c:1327-1336, trt_orbit_camera), the frame is rendered and encoded on the device, and
 * `sink` receives the finished terminal stream (the bytes buffered_draw_screen would fwrite, TRT.c:1171) in frame
 * order.  Two device and two pinned host buffers: the device-to-host copy of frame k and the sink's work on it run
 * while frame k+1 renders.  `bytes` is only valid during the call; a non-zero return from `sink` stops the loop.
 * first/stride give the frame-index sharding of BASELINE config 4 (rank r of N: first = r, stride = N).
 * Returns the number of frames delivered.  scene->camera is used as the un-posed camera (trt_demo_scene's). */
typedef int (*trt_frame_sink)(const char *bytes, size_t n_bytes, int frame, void *user);
int trt_render_orbit(const trt_Scene *scene, int width, int height, const double *times, int n_frames, int first, int stride,
                     trt_frame_sink sink, void *user);

/* The same loop with caller-owned destinations — the cross-process form of the streaming sink: `acquire(frame, n_bytes, user)`
 * returns the PAGE-LOCKED address (trt_host_alloc_pinned / trt_host_register) the frame's bytes are copied to straight from
 * the device; it is called after the frame's kernels have been enqueued and may block until a destination is free (NULL stops
 * the loop).  `sink` (may be NULL) is called with that address once the bytes have landed.  With the destinations in a
 * shared-memory ring that every rank of a node has registered (pipeline.OrderedFrameRing), rank r renders frames r, r+N, ...
 * into the ring while one consumer writes the frames out strictly in order — frame-index sharding of the reference's loop
 * (TRT.c:1317-1367) with ONE ordered output, no gather at the end. */
typedef char *(*trt_frame_acquire)(int frame, size_t n_bytes, void *user);
int trt_render_orbit_to(const trt_Scene *scene, int width, int height, const double *times, int n_frames, int first, int stride,
                        trt_frame_acquire acquire, trt_frame_sink sink, void *user);

/* ---- device-resident pieces (row bands; used by the multi-GPU plumbing and by bench.py) ------- */
/* Upload scene (spheres, ground, lights, camera) for subsequent *_device calls. */
int trt_set_scene(const trt_Scene *scene);
/* The same without waiting for the upload: the copies are enqueued on trt_stream() from page-locked staging (two arenas,
 * an event behind each upload), so a host loop that re-poses the camera every frame never stalls on the device.  The
 * scene is not retained: the caller's structures may change as soon as the call returns. */
int trt_set_scene_async(const trt_Scene *scene);

/* Render rows [row0,row1) of a width x height frame into d_pixels (device pointer to
 * (row1-row0)*width*3 doubles, band-local row-major).  Asynchronous on trt_stream(). */
int trt_render_rows_device(int width, int height, int row0, int row1, double *d_pixels);

/* Encode `rows` rows of `width` pixels from d_pixels into d_bytes: rows*(25*width+1) bytes,
 * starting at d_bytes + byte_offset (any alignment).  Asynchronous on trt_stream(). */
int trt_encode_rows_device(const double *d_pixels, int width, int rows, char *d_bytes, size_t byte_offset);

/* Fast-path variants: the render kernel writes one quantised cell (r,g,b,0) = (int)(c*255) per pixel
 * (4 bytes instead of 24) and the encoder reads those; the bytes produced are identical. */
int trt_render_rows_quant_device(int width, int height, int row0, int row1, unsigned char *d_quant);
int trt_encode_rows_quant_device(const unsigned char *d_quant, int width, int rows, char *d_bytes, size_t byte_offset);
/* K1 with the encoder and the transfer fused in: renders rows [row0,row1) and stores every finished tile's cells as
 * terminal bytes at their place in the stream that starts at `stream` (the rows' bytes only: home sequence and trailing
 * NULs are trt_stream_frame_device's).  `stream` may be memory of this GPU, of a peer (trt_ipc_import: the bytes cross
 * NVLink as the tiles finish, no separate copy) or page-locked host memory (trt_host_alloc_pinned / trt_host_register:
 * zero-copy stores over PCIe).  Same bytes as buffered_draw_screen (TRT.c:1142-1172) produces for these rows. */
int trt_render_rows_ansi_device(int width, int height, int row0, int row1, char *stream);

/* ---- multi-GPU gather over NVLink peer memory (one process per GPU) ------------------------------------------------
 * The only exchange step of the path: every rank's encoded row band is written straight into rank 0's stream buffer.
 * Rank 0 exports its buffer (allocated with trt_device_alloc) as a 64-byte CUDA IPC handle, the other ranks import it
 * once and then push each finished piece with a copy-engine transfer that runs while their next piece renders; no SM
 * and no receiver-side kernel is involved.  trt_peer_copies_wait() blocks until this rank's pushes have landed. */
#define TRT_IPC_HANDLE_BYTES 64
int trt_ipc_export(const void *d_ptr, unsigned char handle[TRT_IPC_HANDLE_BYTES]);
void *trt_ipc_import(const unsigned char handle[TRT_IPC_HANDLE_BYTES]);
int trt_ipc_close(void *d_peer_ptr);
/* after everything enqueued so far on trt_stream(): copy `bytes` from d_src (this GPU) to d_peer_dst (any GPU) */
int trt_push_to_peer(void *d_peer_dst, const void *d_src, size_t bytes);
int trt_peer_copies_wait(void);
/* Step completion without the host.  A frame is complete on rank 0 when every rank's bytes have landed; instead of a host-side
 * collective per frame, every rank ends its step with trt_signal_step(flag of this rank inside rank 0's allocation, step number,
 * after_copies) — a one-thread kernel behind the step's kernels (after_copies = 0, fused gather) or behind its copy-engine pushes
 * (after_copies = 1) that stores the step number with system-scope fences — and rank 0 enqueues trt_wait_steps(flags, n, step): a
 * kernel on ITS stream that waits (bounded: ~4 s, then flags[32] is set) until all n flags have reached the step.  The hosts never
 * synchronise: every rank can enqueue step after step.  Back-pressure works the same way in the other direction: rank 0 signals
 * "frame k consumed" into one more flag word once it has taken the frame, and the other ranks put trt_wait_steps(that word, 1, k)
 * in front of their first write of frame k+1 (on_copy_stream: in front of the pushes instead of the kernels).
 * The flag array (128 words, zeroed) lives behind rank 0's stream buffer. */
int trt_signal_step(void *d_flag, unsigned int value, int after_copies);
int trt_wait_steps(const void *d_flags, int n_flags, unsigned int value, int on_copy_stream);
/* orders trt_stream() behind everything enqueued on the library's copy stream so far (the previous step's pushes) */
int trt_stream_wait_copies(void);
/* When the caller wants the stream in HOST memory (the buffer it fwrite()s, TRT.c:1171) the device-side gather is not
 * needed at all: every rank maps the same host buffer (POSIX shared memory), page-locks its mapping with
 * trt_host_register, and pushes its own bands there with trt_push_to_peer (the destination may be any address the
 * device can reach) — N PCIe links carry the device-to-host transfer instead of rank 0's alone. */
int trt_host_register(void *host_ptr, size_t bytes);
int trt_host_unregister(void *host_ptr);

/* Write the 6-byte home sequence at d_stream[0..5] and the 3 NUL bytes after the last row of a
 * width x height stream (TRT.c:1102, 1104, 1130). */
int trt_stream_frame_device(char *d_stream, int width, int height);

/* Load-balancing pre-pass for row-band sharding: renders the scene at 1/8 resolution with the same camera and
 * returns, for every row of the width x height frame, an estimate of its cost (closest-hit queries).  Deterministic,
 * so every rank that calls it derives the same bands (terminalraytracer_b200/sharding.py, row_bands(weights=...)). */
int trt_estimate_row_costs(const trt_Scene *scene, int width, int height, double *cost_per_row);

/* Same render as trt_render_rows_device, additionally accumulating the work counters of the
 * algorithmic flop model into counters[TRT_NUM_COUNTERS] (host array). Synchronous. */
#define TRT_NUM_COUNTERS 32
int trt_count_rows_device(int width, int height, int row0, int row1, double *d_pixels, long long *counters);
/* F(frame) of SURVEY.md §8(d) from such a counter array */
double trt_model_flops(const long long *counters);

/* Unit-level probes for parity tests of single queries (host arrays in and out, synchronous):
 * trace_ray (TRT.c:793) for n rays (6 doubles each: origin, direction) -> 11 doubles each:
 * kind (0 none,1 sphere,2 ground), pushed-back point[3], unit normal[3], material colour[3], reflectivity;
 * get_skybox_color (TRT.c:700) for n directions (3 doubles each) -> 5 ints each: face, texel index, r, g, b. */
int trt_probe_trace_ray(const trt_Scene *scene, const double *rays, int n, double *out);
int trt_probe_skybox(const double *dirs, int n, int *out);
/* single-function probes against the reference's own known answers (tests/golden/units.npz):
 * ray_intersects_sphere TRT.c:638-672 — rays n x 6 (origin, direction), spheres n x 4 (centre, radius) -> out n x 4 (hit, point);
 * ray_intersects_plane  TRT.c:677-695 against scene->ground — rays n x 6 -> out n x 4 (hit, point);
 * apply_lighting        TRT.c:894-963 — surface n x 9 (point, unit normal, material colour) -> out n x 3 (lit colour, clamped) */
int trt_probe_sphere(const double *rays, const double *spheres, int n, double *out);
int trt_probe_plane(const trt_Scene *scene, const double *rays, int n, double *out);
int trt_probe_lighting(const trt_Scene *scene, const double *surface, int n, double *out);

/* Self-test of the shared-reciprocal division used by the normalisations (csrc/trt_device.cuh): evaluates
 * about `quotients` random and adversarial a/b on the GPU both ways and returns how many differ from the
 * IEEE-754 division in any bit (must be 0). */
long long trt_selftest_division(unsigned long long seed, long long quotients);

/* device memory helpers for plain-C callers (thin wrappers over cudaMalloc / cudaMemcpy) */
void *trt_device_alloc(size_t bytes);
void trt_device_free(void *p);
void *trt_host_alloc_pinned(size_t bytes);
void trt_host_free_pinned(void *p);
int trt_copy_to_host(void *dst, const void *d_src, size_t bytes);
int trt_copy_to_device(void *d_dst, const void *src, size_t bytes);
int trt_synchronize(void);

/* Self-checking build (libtrt_b200 compiled with -DTRT_BOUNDS_CHECK; scripts/bounds_check.py): every computed index of the kernels
 * is checked on the device; out32[0..15] = violations per check site of the render kernels, [16..31] of the encode kernel, read and
 * cleared.  Returns 1 when the checks are compiled in, 0 in the product build (all counters 0). */
int trt_debug_bounds(unsigned int *out32);

/* timing of the last trt_project_scene / trt_render_ansi call, CUDA events on trt_stream(), ms */
float trt_last_render_ms(void);
float trt_last_encode_ms(void);

/* ---- measured ALU peaks (dependent-free FMA loops on every SM; TFLOP/s) ---------------------- */
double trt_measure_fp32_tflops(void);
double trt_measure_fp64_tflops(void);

/* ---- host-side helpers (trt_host.c; no CUDA) -------------------------------------------------- */
void trt_init_camera(trt_Camera *camera, int width, int height);                     /* TRT.c:299 */
void trt_orbit_camera(trt_Camera *camera, double t);                                 /* TRT.c:1327-1336 */
/* the same pose recipe from explicit angles (radians) and distance: keyboard-driven cameras (README TODO of the reference) */
void trt_pose_camera(trt_Camera *camera, double pitch, double yaw, double radius);
void trt_subpixel_offsets(double dx[TRT_RAYS_PER_PIXEL], double dy[TRT_RAYS_PER_PIXEL]); /* TRT.c:992-993 */
void trt_demo_scene(trt_Scene *scene, trt_Sphere spheres[TRT_DEMO_SPHERES], trt_DirectionalLight *dl, trt_PointLight *pl,
                    int width, int height);                                          /* TRT.c:1256-1306 */
int trt_stress_scene(trt_Scene *scene, trt_Sphere *spheres, int count, trt_DirectionalLight *dl, trt_PointLight *pl,
                     int width, int height);                                         /* SURVEY §8d config 3 */
void trt_read_ppm(const char *filename, trt_Color **colors_ptr, int *width, int *height); /* TRT.c:309 */
void trt_load_skybox(trt_Skybox *skybox, const char *skybox_name);                   /* TRT.c:388 (cwd/skybox/<name>) */
void trt_load_skybox_dir(trt_Skybox *skybox, const char *dir);
void trt_free_skybox(trt_Skybox *skybox);                                            /* TRT.c:430 */

#ifdef __cplusplus
}
#endif
#endif /* TRT_B200_H */
