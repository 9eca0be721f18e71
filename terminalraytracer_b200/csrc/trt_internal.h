// trt_internal.h — C++ glue between the translation units of libtrt_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "trt_device.cuh"

namespace trt {

// trt_render.cu (owns the __constant__ scene)
void upload_scene_constants(const DevScene &scene, cudaStream_t stream);
// one_plus_one: the scene has exactly one directional and one point light (specialised kernel flavour)
void launch_render(const RenderParams &p, bool count, int cull, bool one_plus_one, int num_sms, cudaStream_t stream);
int render_ctas_per_sm();
int render_bounds_read(unsigned int *out16);   // self-checking build (-DTRT_BOUNDS_CHECK): violation counters of trt_render.cu
int encode_bounds_read(unsigned int *out16);   // ... of trt_encode.cu
size_t render_scratch_bytes(int num_sms);   // RenderParams::sample_scratch must be at least this big
size_t render_tile_info_bytes(int width, int rows);   // RenderParams::tile_info for a launch of `rows` rows (k_tile_certs)
unsigned long long run_selftest_division(unsigned long long seed, int ctas, int iters, unsigned long long *d_scratch, cudaStream_t stream);
void launch_probe_trace(const RenderParams &p, const double *d_rays, int n, double *d_out, cudaStream_t stream);
void launch_probe_sky(const RenderParams &p, const double *d_dirs, int n, int *d_out, cudaStream_t stream);
void launch_probe_sphere(const double *d_rays, const double *d_geom, int n, double *d_out, cudaStream_t stream);
void launch_probe_plane(const double *d_rays, int n, double *d_out, cudaStream_t stream);
void launch_probe_lighting(const RenderParams &p, const double *d_in, int n, double *d_out, cudaStream_t stream);

// trt_encode.cu
// Encode `rows` rows of `width` cells into out_base[byte_offset ...); source is either the FP64
// framebuffer (3 doubles per pixel) or the quantised cells written by the render kernel.
void launch_encode_f64(const double *pixels, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream);
void launch_encode_quant(const uchar4 *quant, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream);
void launch_stream_frame(char *stream_base, int width, int height, cudaStream_t stream);
void launch_signal(unsigned int *flag, unsigned int value, cudaStream_t stream);
void launch_wait_flags(const unsigned int *flags, int n, unsigned int value, unsigned int *timed_out, cudaStream_t stream);

// trt_peak.cu
double measure_fp32_tflops(cudaStream_t stream);
double measure_fp64_tflops(cudaStream_t stream);

} // namespace trt
