// trt_encode.cu — K2: framebuffer -> 24-bit ANSI escape stream.
//
// Replaces initialize_screenbuffer (TRT.c:1107-1131), byte_to_digits (1134-1139) and the patching
// loop of buffered_draw_screen (1142-1168); TRT.c = /root/reference/TerminalRayTracer.c.
// Output layout (TRT.c:1102-1104):  "\033[0;0H"  +  H x ( W x "\033[48;2;RRR;GGG;BBBm  \033[0m" + "\n" )
// + 3 NUL bytes.  A cell is 25 bytes, a row 25W+1 bytes -> rows start at odd, unaligned addresses.
//
// This is pure byte traffic (24 B read per FP64 pixel or 4 B per quantised cell, 25 B written), so the
// kernel is built around store efficiency: each CTA owns a 16-byte-aligned 8 KB window of the
// destination, threads format whole cells into shared memory (byte-granular, any alignment), and the
// window then leaves the SM as full 16-byte vector stores; only the first/last window of a band can be
// partial and falls back to byte stores.
#include <cstdio>
#include <cstdlib>
#include "trt_internal.h"

namespace trt {

constexpr int ENC_THREADS = 256;
constexpr int ENC_WINDOW = ENC_THREADS * 16 * 2; // bytes of destination per CTA

// (int)(c*255) as the x86-64 reference build does it, then byte_to_digits' integer arithmetic
__device__ __forceinline__ int quantise(double c)
{
    const double v = c * 255;
    return (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
}

__device__ __forceinline__ void put(unsigned char *win, long long at, long long win_len, unsigned char b)
{
    if (at >= 0 && at < win_len) win[at] = b;
}

template <typename Src> struct Load;
template <> struct Load<double> {
    static __device__ __forceinline__ void rgb(const double *src, size_t pixel, int &r, int &g, int &b)
    {
        const double *p = src + pixel * 3;
        r = quantise(__ldg(p + 0));
        g = quantise(__ldg(p + 1));
        b = quantise(__ldg(p + 2));
    }
};
template <> struct Load<uchar4> {
    static __device__ __forceinline__ void rgb(const uchar4 *src, size_t pixel, int &r, int &g, int &b)
    {
        const uchar4 q = __ldg(src + pixel);
        r = q.x;
        g = q.y;
        b = q.z;
    }
};

template <typename Src>
__global__ void __launch_bounds__(ENC_THREADS) k_encode(const Src *__restrict__ src, int width, int rows,
                                                        unsigned char *__restrict__ out_base, unsigned long long byte_offset)
{
    __shared__ __align__(16) unsigned char win[ENC_WINDOW];
    const unsigned long long row_bytes = (unsigned long long)TRT_CELL_BYTES * (unsigned long long)width + 1ull;
    const unsigned long long region0 = byte_offset;
    const unsigned long long region1 = byte_offset + row_bytes * (unsigned long long)rows;
    const unsigned long long aligned0 = region0 & ~15ull;
    const unsigned long long w0 = aligned0 + (unsigned long long)blockIdx.x * ENC_WINDOW; // window start (16B aligned)
    unsigned long long w1 = w0 + ENC_WINDOW;
    const unsigned long long lo = w0 > region0 ? w0 : region0; // valid bytes of this window: [lo, hi)
    const unsigned long long hi = w1 < region1 ? w1 : region1;
    if (lo >= hi) return;
    const long long win_len = (long long)(hi - w0);

    // cells (row-major, the last cell of a row also owns the '\n') that intersect [lo, hi)
    const unsigned long long rel_lo = lo - region0, rel_hi = hi - 1 - region0;
    const unsigned long long r_lo = rel_lo / row_bytes, r_hi = rel_hi / row_bytes;
    unsigned long long c_lo = (rel_lo - r_lo * row_bytes) / TRT_CELL_BYTES;
    unsigned long long c_hi = (rel_hi - r_hi * row_bytes) / TRT_CELL_BYTES;
    if (c_lo >= (unsigned long long)width) c_lo = width - 1;
    if (c_hi >= (unsigned long long)width) c_hi = width - 1;
    const unsigned long long cell0 = r_lo * width + c_lo, cell1 = r_hi * width + c_hi;
    const int ncells = (int)(cell1 - cell0 + 1);

    for (int i = threadIdx.x; i < ncells; i += ENC_THREADS) {
        const unsigned long long cid = cell0 + (unsigned long long)i;
        const unsigned long long row = cid / (unsigned long long)width;
        const unsigned int cell = (unsigned int)(cid - row * (unsigned long long)width);
        int v[3];
        Load<Src>::rgb(src, (size_t)cid, v[0], v[1], v[2]);
        // position of the cell's first byte relative to the window start
        const long long at = (long long)(region0 + row * row_bytes + (unsigned long long)cell * TRT_CELL_BYTES) - (long long)w0;
        unsigned char b[TRT_CELL_BYTES + 1];
        b[0] = 0x1b; b[1] = '['; b[2] = '4'; b[3] = '8'; b[4] = ';'; b[5] = '2'; b[6] = ';';
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            b[7 + ch * 4 + 0] = (unsigned char)(v[ch] / 100 + '0');        // TRT.c:1136
            b[7 + ch * 4 + 1] = (unsigned char)((v[ch] / 10) % 10 + '0');  // TRT.c:1137
            b[7 + ch * 4 + 2] = (unsigned char)(v[ch] % 10 + '0');         // TRT.c:1138
        }
        b[10] = ';'; b[14] = ';'; b[18] = 'm'; b[19] = ' '; b[20] = ' ';
        b[21] = 0x1b; b[22] = '['; b[23] = '0'; b[24] = 'm'; b[25] = '\n';
        const int nbytes = (cell == (unsigned int)(width - 1)) ? TRT_CELL_BYTES + 1 : TRT_CELL_BYTES;
        if (at >= 0 && at + TRT_CELL_BYTES + 1 <= win_len) {
#pragma unroll
            for (int j = 0; j < TRT_CELL_BYTES; j++) win[at + j] = b[j];
            if (nbytes > TRT_CELL_BYTES) win[at + TRT_CELL_BYTES] = b[TRT_CELL_BYTES];
        } else {
#pragma unroll
            for (int j = 0; j < TRT_CELL_BYTES + 1; j++)
                if (j < nbytes) put(win, at + j, win_len, b[j]);
        }
    }
    __syncthreads();

    // drain the window: full 16-byte stores wherever the whole chunk is valid
    unsigned char *dst = out_base + w0;
    for (int chunk = threadIdx.x; chunk < ENC_WINDOW / 16; chunk += ENC_THREADS) {
        const unsigned long long g0 = w0 + (unsigned long long)chunk * 16, g1 = g0 + 16;
        if (g0 >= lo && g1 <= hi) {
            *reinterpret_cast<uint4 *>(dst + chunk * 16) = *reinterpret_cast<const uint4 *>(win + chunk * 16);
        } else if (g1 > lo && g0 < hi) {
            for (int j = 0; j < 16; j++)
                if (g0 + j >= lo && g0 + j < hi) dst[chunk * 16 + j] = win[chunk * 16 + j];
        }
    }
}

__global__ void k_stream_frame(unsigned char *base, unsigned long long tail_at)
{
    const unsigned char home[TRT_HOME_BYTES] = {0x1b, '[', '0', ';', '0', 'H'}; // TRT.c:1102
    const int t = threadIdx.x;
    if (t < TRT_HOME_BYTES) base[t] = home[t];
    else if (t < TRT_HOME_BYTES + TRT_TAIL_NULS) base[tail_at + (t - TRT_HOME_BYTES)] = 0; // TRT.c:1104, 1130
}

static void ck(cudaError_t e, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "%s:%d: CUDA error: %s\n", __FILE__, line, cudaGetErrorString(e));
        exit(1);
    }
}

template <typename Src>
static void launch_encode(const Src *src, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    if (rows <= 0 || width <= 0) return;
    if ((reinterpret_cast<uintptr_t>(out_base) & 15) != 0) {
        fprintf(stderr, "%s:%d: encode destination must be 16-byte aligned\n", __FILE__, __LINE__);
        exit(1);
    }
    const unsigned long long row_bytes = (unsigned long long)TRT_CELL_BYTES * width + 1ull;
    const unsigned long long region0 = byte_offset, region1 = byte_offset + row_bytes * rows;
    const unsigned long long aligned0 = region0 & ~15ull;
    const unsigned long long windows = (region1 - aligned0 + ENC_WINDOW - 1) / ENC_WINDOW;
    k_encode<Src><<<(unsigned)windows, ENC_THREADS, 0, stream>>>(src, width, rows, reinterpret_cast<unsigned char *>(out_base),
                                                               (unsigned long long)byte_offset);
    ck(cudaGetLastError(), __LINE__);
}

void launch_encode_f64(const double *pixels, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    launch_encode<double>(pixels, width, rows, out_base, byte_offset, stream);
}
void launch_encode_quant(const uchar4 *quant, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    launch_encode<uchar4>(quant, width, rows, out_base, byte_offset, stream);
}
void launch_stream_frame(char *stream_base, int width, int height, cudaStream_t stream)
{
    const unsigned long long tail_at = TRT_HOME_BYTES + ((unsigned long long)TRT_CELL_BYTES * width + 1ull) * height;
    k_stream_frame<<<1, 32, 0, stream>>>(reinterpret_cast<unsigned char *>(stream_base), tail_at);
    ck(cudaGetLastError(), __LINE__);
}

} // namespace trt
