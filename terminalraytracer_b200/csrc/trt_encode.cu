// trt_encode.cu — K2: framebuffer -> 24-bit ANSI escape stream.
//
// Replaces initialize_screenbuffer (TRT.c:1107-1131), byte_to_digits (1134-1139) and the patching
// loop of buffered_draw_screen (1142-1168); TRT.c = /root/reference/TerminalRayTracer.c.
// Output layout (TRT.c:1102-1104):  "\033[0;0H"  +  H x ( W x "\033[48;2;RRR;GGG;BBBm  \033[0m" + "\n" )
// + 3 NUL bytes.  A cell is 25 bytes, a row 25W+1 bytes -> rows start at odd, unaligned addresses.
//
// This is pure byte traffic (24 B read per FP64 pixel or 4 B per quantised cell, 25 B written), so the kernel is
// built around instruction economy and store efficiency: whole words are composed in registers (no per-byte
// stores, no divisions), staged in shared memory, and leave the SM as 16-byte stores aligned to the destination.
#include <cstdio>
#include <cstdlib>
#include "trt_internal.h"

namespace trt {

constexpr int ENC_THREADS = 256;
constexpr int ENC_CELLS_PER_THREAD = 4;                             // 4 cells = 100 bytes = 25 words: word aligned
constexpr int ENC_CELLS = ENC_THREADS * ENC_CELLS_PER_THREAD;       // cells of one row per CTA
constexpr int ENC_WORDS = ENC_CELLS * TRT_CELL_BYTES / 4;           // 6400 words staged per CTA
static_assert((ENC_CELLS_PER_THREAD * TRT_CELL_BYTES) % 4 == 0, "a thread's cells must fill whole words");

// (int)(c*255) as the x86-64 reference build does it, then byte_to_digits' integer arithmetic
__device__ __forceinline__ int quantise(double c)
{
    const double v = c * 255;
    return (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;
}

// byte_to_digits (TRT.c:1134-1139) on an int: value/100, (value/10)%10, value%10, each + '0', truncated to a char.
// For 0..255 — everything a [0,1] framebuffer produces — the three characters come from a table; other ints
// (pixels outside [0,1], NaN -> INT_MIN) take the arithmetic, whose wrap-around is what the reference prints.
__device__ __forceinline__ unsigned int digits_of(int v, const unsigned int *s_lut)
{
    if ((unsigned)v < 256u) return s_lut[v];
    return (unsigned)(unsigned char)(v / 100 + '0') | ((unsigned)(unsigned char)((v / 10) % 10 + '0') << 8) |
           ((unsigned)(unsigned char)(v % 10 + '0') << 16);
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

// OR the 32-bit little-endian `word` into out[] at byte offset OFF (compile-time): at most two registers
template <int OFF>
__device__ __forceinline__ void put_word(unsigned int (&out)[ENC_CELLS_PER_THREAD * TRT_CELL_BYTES / 4 + 1], unsigned int word)
{
    out[OFF >> 2] |= word << (8 * (OFF & 3));
    if constexpr ((OFF & 3) != 0) out[(OFF >> 2) + 1] |= word >> (32 - 8 * (OFF & 3));
}

// the 25 bytes of one cell, "\033[48;2;RRR;GGG;BBBm  \033[0m" (TRT.c:1103), as seven words at byte offset OFF
template <int OFF>
__device__ __forceinline__ void put_cell(unsigned int (&out)[ENC_CELLS_PER_THREAD * TRT_CELL_BYTES / 4 + 1], unsigned int r3, unsigned int g3,
                                         unsigned int b3)
{
    put_word<OFF + 0>(out, 0x38345b1bu);                                     // ESC [ 4 8
    put_word<OFF + 4>(out, 0x003b323bu | (r3 << 24));                        // ; 2 ; R
    put_word<OFF + 8>(out, (r3 >> 8) | 0x003b0000u | (g3 << 24));            // R R ; G
    put_word<OFF + 12>(out, (g3 >> 8) | 0x003b0000u | (b3 << 24));           // G G ; B
    put_word<OFF + 16>(out, (b3 >> 8) | 0x206d0000u);                        // B B m ' '
    put_word<OFF + 20>(out, 0x305b1b20u);                                    // ' ' ESC [ 0
    out[(OFF + 24) >> 2] |= 0x6du << (8 * ((OFF + 24) & 3));                 // m
}

// One CTA = up to 1024 consecutive cells of ONE row (no divisions to find rows).  Each thread formats 4 cells — 100 bytes,
// 25 whole words built in registers with compile-time byte positions.  Rows start at odd offsets (6 + r(25W+1)), so the CTA's
// window of the stream begins `a` bytes into a 16-byte chunk of the destination; the window is STAGED AT THAT SAME PHASE in
// shared memory (staging byte a + j = window byte j): a thread shifts its 100 bytes by a & 3 bytes while they are still in
// registers (the low bytes of its first word are the tail of its left neighbour's last cell — always "[0m" — so every word
// is written whole, by one thread, conflict-free: stride 25 words).  The staged chunks then are the destination's chunks, and
// all complete ones leave with ONE bulk copy (cp.async.bulk shared -> global, the TMA engine: no per-thread loads, shifts or
// stores); only the first and last chunk of a CTA can straddle a neighbour's bytes and are written bytewise.
template <typename Src>
__global__ void __launch_bounds__(ENC_THREADS) k_encode(const Src *__restrict__ src, int width, int rows,
                                                        unsigned char *__restrict__ out_base, unsigned long long byte_offset)
{
    __shared__ __align__(128) unsigned int s_win[ENC_WORDS + 8];
    __shared__ unsigned int s_lut[256];
    {
        const int v = threadIdx.x;   // ENC_THREADS == 256
        s_lut[v] = (unsigned)(v / 100 + '0') | ((unsigned)((v / 10) % 10 + '0') << 8) | ((unsigned)(v % 10 + '0') << 16);
    }
    const int row = blockIdx.y;
    const int c0 = blockIdx.x * ENC_CELLS;
    const int ncells = min(ENC_CELLS, width - c0);
    const unsigned long long row_bytes = (unsigned long long)TRT_CELL_BYTES * (unsigned long long)width + 1ull;
    // first byte of this CTA in the stream, and how many it owns (the last CTA of a row also owns the '\n')
    const unsigned long long g0 = byte_offset + (unsigned long long)row * row_bytes + (unsigned long long)c0 * TRT_CELL_BYTES;
    const int nbytes = ncells * TRT_CELL_BYTES + ((c0 + ncells == width) ? 1 : 0);
    const unsigned int a = (unsigned int)(g0 & 15ull);           // window byte 0 sits `a` bytes into its 16-byte chunk
    __syncthreads();

    // ---- format: 4 cells per thread, staged at the destination's phase ---------------------------------------------------
    const int first = threadIdx.x * ENC_CELLS_PER_THREAD;
    if (first < ncells) {
        constexpr int NW = ENC_CELLS_PER_THREAD * TRT_CELL_BYTES / 4;   // 25
        unsigned int out[NW + 1];
#pragma unroll
        for (int k = 0; k < NW + 1; k++) out[k] = 0u;
        unsigned int d[ENC_CELLS_PER_THREAD][3];
#pragma unroll
        for (int k = 0; k < ENC_CELLS_PER_THREAD; k++) {
            int v[3] = {0, 0, 0};
            if (first + k < ncells) Load<Src>::rgb(src, (size_t)row * (size_t)width + (size_t)(c0 + first + k), v[0], v[1], v[2]);
            d[k][0] = digits_of(v[0], s_lut);
            d[k][1] = digits_of(v[1], s_lut);
            d[k][2] = digits_of(v[2], s_lut);
        }
        put_cell<0 * TRT_CELL_BYTES>(out, d[0][0], d[0][1], d[0][2]);
        put_cell<1 * TRT_CELL_BYTES>(out, d[1][0], d[1][1], d[1][2]);
        put_cell<2 * TRT_CELL_BYTES>(out, d[2][0], d[2][1], d[2][2]);
        put_cell<3 * TRT_CELL_BYTES>(out, d[3][0], d[3][1], d[3][2]);
        // shift by a & 3 bytes: staging word (a >> 2) + 25 t + k takes the high bytes of out[k-1] and the low bytes of out[k];
        // out[-1] is the left neighbour's last word, and a cell always ends in ESC [ 0 m
        const unsigned int sh = 8u * (a & 3u);                   // uniform over the CTA
        unsigned int *const dst = s_win + (a >> 2) + threadIdx.x * NW;
        TRT_BOUND((a >> 2) + threadIdx.x * NW + NW < ENC_WORDS + 8, 0);
        unsigned int prev = 0x6d305b1bu;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            dst[k] = __funnelshift_l(prev, out[k], sh);           // (out[k] << sh) | (prev >> (32 - sh)); sh == 0: out[k]
            prev = out[k];
        }
        // the word behind a thread's 25 belongs to its right neighbour — unless there is none (end of the row or of the CTA)
        if (first + ENC_CELLS_PER_THREAD >= ncells) dst[NW] = __funnelshift_l(prev, 0u, sh);
        // end of the row inside this thread's cells: the byte after the last cell is '\n' (TRT.c:1103, 1125)
        const int mine = min(ENC_CELLS_PER_THREAD, ncells - first);
        if (c0 + first + mine == width) reinterpret_cast<unsigned char *>(s_win)[a + threadIdx.x * 100 + mine * TRT_CELL_BYTES] = 0x0a;
    }
    // the staged bytes are read by the bulk-copy engine (async proxy): make the generic-proxy writes visible to it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // ---- drain ----------------------------------------------------------------------------------------------------------
    unsigned char *const chunk0 = out_base + (g0 - a);            // 16-byte aligned (out_base is)
    const int nchunks = (int)((a + (unsigned)nbytes + 15u) >> 4);
    const int q_first = a ? 1 : 0;                                // chunk 0 starts before the window unless a == 0
    const int q_last = (int)((a + (unsigned)nbytes) >> 4);        // first chunk that is not complete
    if (threadIdx.x == 0 && q_last > q_first) {
        const unsigned int bytes = (unsigned int)(q_last - q_first) * 16u;
        const unsigned int s_addr = (unsigned int)__cvta_generic_to_shared(s_win) + (unsigned int)q_first * 16u;
        unsigned char *const g_addr = chunk0 + (size_t)q_first * 16;
        TRT_BOUND(g0 >= byte_offset && g0 + (unsigned long long)nbytes <= byte_offset + (unsigned long long)rows * row_bytes &&
                  (unsigned long long)q_first * 16ull >= a && (unsigned long long)q_last * 16ull <= a + (unsigned long long)nbytes &&
                  (reinterpret_cast<unsigned long long>(g_addr) & 15ull) == 0ull && (s_addr & 15u) == 0u && q_last * 16 <= (ENC_WORDS + 8) * 4, 1);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_addr), "r"(s_addr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // the at most two partial chunks, bytewise (lanes of the second warp: the first one's lane 0 is busy with the bulk copy)
    const unsigned char *const s_bytes = reinterpret_cast<const unsigned char *>(s_win);
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        const int j = (int)threadIdx.x - 32;                      // 0..15: head chunk, 16..31: tail chunk
        const int q = j < 16 ? 0 : q_last;
        const int b = q * 16 + (j & 15);                          // staging byte
        const bool partial = j < 16 ? (q_first == 1) : (q_last < nchunks && q_last >= q_first);
        if (partial && b >= (int)a && b < (int)a + nbytes) {
            TRT_BOUND(b < (ENC_WORDS + 8) * 4, 2);
            chunk0[b] = s_bytes[b];
        }
    }
    if (threadIdx.x == 0 && q_last > q_first) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the read
}

__global__ void k_stream_frame(unsigned char *base, unsigned long long tail_at)
{
    const unsigned char home[TRT_HOME_BYTES] = {0x1b, '[', '0', ';', '0', 'H'}; // TRT.c:1102
    const int t = threadIdx.x;
    if (t < TRT_HOME_BYTES) base[t] = home[t];
    else if (t < TRT_HOME_BYTES + TRT_TAIL_NULS) base[tail_at + (t - TRT_HOME_BYTES)] = 0; // TRT.c:1104, 1130
}

// ---- step completion without the host (multi-GPU) ---------------------------------------------------------------------
// k_signal: "everything this rank enqueued before me on this stream has been performed" -> a step number stored into a flag word
// that lives in rank 0's memory (peer mapping) or in its own.  The kernels and copy-engine transfers that wrote the band's bytes
// precede it in stream order; the system-scope fence orders the flag behind them for an observer on another GPU.
__global__ void k_signal(unsigned int *flag, unsigned int value)
{
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int *>(flag) = value;
    __threadfence_system();
}

// k_wait_flags: rank 0's stream waits until every rank's flag has reached `value` (lane = rank).  Bounded: gives up after
// ~4 s of GPU clock and raises *timed_out instead of hanging the device if a rank died.
__global__ void k_wait_flags(const unsigned int *flags, int n, unsigned int value, unsigned int *timed_out)
{
    const int r = threadIdx.x;
    if (r >= n) return;
    const volatile unsigned int *f = flags + r;
    const long long t0 = clock64();
    // (values only grow; the subtraction keeps the comparison right across a wrap of the 32-bit step counter)
    while ((int)(*f - value) < 0) {
        if (clock64() - t0 > 8000000000ll) {
            *timed_out = 1u;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

static void ck(cudaError_t e, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "%s:%d: CUDA error: %s\n", __FILE__, line, cudaGetErrorString(e));
        exit(1);
    }
}

int encode_bounds_read(unsigned int *out16)
{
#ifdef TRT_BOUNDS_CHECK
    ck(cudaDeviceSynchronize(), __LINE__);
    ck(cudaMemcpyFromSymbol(out16, g_trt_bounds, sizeof(unsigned int) * 16), __LINE__);
    const unsigned int zero[16] = {0};
    ck(cudaMemcpyToSymbol(g_trt_bounds, zero, sizeof zero), __LINE__);
    return 1;
#else
    for (int i = 0; i < 16; i++) out16[i] = 0u;
    return 0;
#endif
}

template <typename Src>
static void launch_encode(const Src *src, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    if (rows <= 0 || width <= 0) return;
    if ((reinterpret_cast<uintptr_t>(out_base) & 15) != 0) {
        fprintf(stderr, "%s:%d: encode destination must be 16-byte aligned\n", __FILE__, __LINE__);
        exit(1);
    }
    const int blocks_x = (width + ENC_CELLS - 1) / ENC_CELLS;
    // gridDim.y is limited to 65535: taller bands go in slices
    for (int r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        const unsigned long long row_bytes = (unsigned long long)TRT_CELL_BYTES * width + 1ull;
        k_encode<Src><<<dim3((unsigned)blocks_x, (unsigned)nr), ENC_THREADS, 0, stream>>>(
            src + (size_t)r0 * (size_t)width * (sizeof(Src) == sizeof(double) ? 3 : 1), width, nr, reinterpret_cast<unsigned char *>(out_base),
            (unsigned long long)byte_offset + (unsigned long long)r0 * row_bytes);
        ck(cudaGetLastError(), __LINE__);
    }
}

void launch_encode_f64(const double *pixels, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    launch_encode<double>(pixels, width, rows, out_base, byte_offset, stream);
}
void launch_encode_quant(const uchar4 *quant, int width, int rows, char *out_base, size_t byte_offset, cudaStream_t stream)
{
    launch_encode<uchar4>(quant, width, rows, out_base, byte_offset, stream);
}
void launch_signal(unsigned int *flag, unsigned int value, cudaStream_t stream)
{
    k_signal<<<1, 1, 0, stream>>>(flag, value);
    ck(cudaGetLastError(), __LINE__);
}
void launch_wait_flags(const unsigned int *flags, int n, unsigned int value, unsigned int *timed_out, cudaStream_t stream)
{
    k_wait_flags<<<1, 32, 0, stream>>>(flags, n, value, timed_out);
    ck(cudaGetLastError(), __LINE__);
}
void launch_stream_frame(char *stream_base, int width, int height, cudaStream_t stream)
{
    const unsigned long long tail_at = TRT_HOME_BYTES + ((unsigned long long)TRT_CELL_BYTES * width + 1ull) * height;
    k_stream_frame<<<1, 32, 0, stream>>>(reinterpret_cast<unsigned char *>(stream_base), tail_at);
    ck(cudaGetLastError(), __LINE__);
}

} // namespace trt
