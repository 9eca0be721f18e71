// trt_render.cu — K1: the per-pixel render loop of the reference as one persistent sm_100a kernel.
//
// Replaces project_scene (TRT.c:966-1069) and everything it calls: trace_ray (793-889),
// ray_intersects_sphere (638-672), ray_intersects_plane (677-695), get_skybox_color (700-789),
// apply_lighting (894-963), reflect_vector (627-633).  TRT.c = /root/reference/TerminalRayTracer.c.
//
// Exactness contract: every value that reaches a pixel is computed with the reference's IEEE-double
// operations in the reference's order (this TU is compiled with -fmad=false), so the framebuffer is
// bit-identical to the reference's.  What is NOT reproduced is work that cannot influence a pixel:
//   * the skybox lookup the reference performs when a SHADOW ray misses (TRT.c:907, 937 -> 858-867),
//   * the reflect/normalise the reference performs after a sample has already hit the sky (1054-1055),
//   * the `view` vector (1029) and the specularity field (118), both never read.
//
// Execution model (why it looks nothing like the reference's call tree):
//   * persistent CTAs, one per SM slot; each WARP pulls 8x4-pixel tiles from a global atomic counter,
//     lane = pixel, so the 32 rays of a warp are spatially coherent;
//   * every lane runs ONE loop whose body starts with the closest-hit query (the sphere loop is ~half
//     of all work) and then advances a small per-lane state machine: primary/bounce ray -> one shadow
//     ray per light -> shade/reflect -> next bounce or next sample.  Primary, bounce and shadow rays of
//     different lanes therefore share the same sphere-loop instructions instead of serialising three
//     inlined copies of trace_ray, and a lane that finishes a sample immediately starts its next one;
//   * the scene sits in __constant__ memory (warp-uniform index in the sphere loop -> one broadcast
//     LDC per operand); the skybox is an RGBA8 repack read through the read-only path and stays
//     L2-resident; the k/255.0 colour table is in shared memory.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "trt_device.cuh"
#include "trt_internal.h"

namespace trt {

__constant__ DevScene c_scene;
__constant__ float4 c_sphere_cull[TRT_MAX_CONST_SPHERES + 2]; // (cx, cy, cz, r_pad) in float, see sphere_cull()

constexpr int TILE_W = 8;
constexpr int TILE_H = 4;
#ifndef TRT_WARPS_PER_CTA
#define TRT_WARPS_PER_CTA 4
#endif
#ifndef TRT_MIN_CTAS_PER_SM
#define TRT_MIN_CTAS_PER_SM 6
#endif
constexpr int WARPS_PER_CTA = TRT_WARPS_PER_CTA;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;

enum Phase : int { PH_MAIN = 0, PH_SHADOW = 1 };

// (cx,cy,cz,r*r) of sphere i through the read-only path (scenes too large for __constant__)
__device__ __forceinline__ double4 ldg_geom(const double4 *geom, int i)
{
    const double2 *p = reinterpret_cast<const double2 *>(geom + i);
    const double2 lo = __ldg(p), hi = __ldg(p + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}

template <bool COUNT>
struct Tally {
    unsigned long long *g;
    __device__ __forceinline__ void add(int id) const
    {
        if (COUNT) atomicAdd(&g[id], 1ull);
    }
};

// ---- get_skybox_color, TRT.c:700-789 -----------------------------------------------------------------
// Returns the linear texel index inside the chosen face and the face itself.  The dot products with
// the axis table reduce exactly (x*1.0 + y*0.0 + z*0.0 == x for finite inputs), so the face argmax,
// the projection and the (u,v) extraction below are the reference's values, not approximations.
// `dir` is the already normalised direction (normalize_vector_copy, TRT.c:702, is done by the caller).
__device__ __forceinline__ int sky_texel_index(const d3 &dir, int dim, int &face)
{
    // argmax over (+x,-x,+y,-y,+z,-z), strict >, first wins, start at -1.0 (TRT.c:703-713)
    double best_t = -1.0;
    int best = -1;
    const double cand[6] = {dir.x, -dir.x, dir.y, -dir.y, dir.z, -dir.z};
#pragma unroll
    for (int f = 0; f < 6; f++) {
        if (cand[f] > best_t) {
            best_t = cand[f];
            best = f;
        }
    }
    // scale_by == best_t (sum of dir (*) axis), TRT.c:717-719
    const double s = ieee_div(1.0, best_t);
    const double px = dir.x * s, py = dir.y * s, pz = dir.z * s;
    // orthogonal component * 0.5, then dots with axes (best+2)%6 and (best+4)%6 (TRT.c:720-727)
    double u, v;
    switch (best) {
    case 0: u = py * 0.5; v = pz * 0.5; break;
    case 1: u = -(py * 0.5); v = -(pz * 0.5); break;
    case 2: u = pz * 0.5; v = px * 0.5; break;
    case 3: u = -(pz * 0.5); v = -(px * 0.5); break;
    case 4: u = px * 0.5; v = py * 0.5; break;
    default: u = -(px * 0.5); v = -(py * 0.5); break;
    }
    if (best & 1) u *= -1.0;                     // :730
    if (best <= 1) { double t = u; u = v; v = -t; }          // :735
    else if (best <= 3) { double t = u; u = -v; v = t; }     // :742-755
    else if (best == 4) { u *= -1.0; v *= -1.0; }            // :756
    u = clampd(u, -0.5, 0.5);
    v = clampd(v, -0.5, 0.5);
    const int ui = x86_int((u + 0.5) * dim);
    const int vi = x86_int((v + 0.5) * dim);
    face = best;
    return ui + vi * dim;
}

// ---- conservative FP32 miss test ---------------------------------------------------------------------
// ~90% of all ray/sphere tests of this path end at `discriminant < 0` (TRT.c:651).  That outcome is a
// geometric fact — the line passes the centre at more than r — which single precision can CERTIFY for
// all but the rays that graze the silhouette.  A sphere is culled only when
//        |oc_f x d_f|^2  >  ( |d_f| * r_pad  +  |d_f| * CULL_EPS * (|o|_1 + max_i |c_i|_1) )^2
// evaluated in float, where r_pad >= r*(1+2^-20) (host, rounded up) and CULL_EPS = 64 * 2^-24.
// Error budget (DESIGN.md "FP32 cull"): rounding o, c, d to float and the float evaluation of the
// cross product move |oc x d| by at most 9*2^-24*|d|*(|o|_1+|c|_1); the reference's own FP64 rounding
// of the discriminant is below 2^-50 of the same scale.  With the 64*2^-24 margin the inequality above
// implies the reference's COMPUTED discriminant is negative, i.e. the reference returns "miss" for this
// sphere; everything else (and any NaN/inf, which makes the comparison false) goes to the exact FP64
// test below, so results stay bit-identical.  The counting build re-checks every culled sphere exactly
// (CTR_CULL_VIOLATIONS must be 0; tests/test_gpu_parity.py).
constexpr float CULL_EPS = 64.0f * 5.9604644775390625e-08f;

struct RayF32 {
    float ox, oy, oz, dx, dy, dz;
    float norm_d;     // |d_f|
    float slack;      // |d_f| * CULL_EPS * (|o|_1 + max_i |c_i|_1)
    bool usable;      // magnitudes inside the range the bound was derived for
};

__device__ __forceinline__ RayF32 ray_to_f32(const d3 &o, const d3 &d)
{
    RayF32 r;
    r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
    r.dx = (float)d.x; r.dy = (float)d.y; r.dz = (float)d.z;
    r.norm_d = __fsqrt_rn(__fmaf_rn(r.dz, r.dz, __fmaf_rn(r.dy, r.dy, r.dx * r.dx)));
    const float l1 = (fabsf(r.ox) + fabsf(r.oy)) + (fabsf(r.oz) + c_scene.filter_centre_l1);
    r.slack = r.norm_d * (CULL_EPS * l1);
    r.usable = c_scene.filter_enabled && (l1 < 1e15f) && (r.norm_d > 1e-15f) && (r.norm_d < 1e15f);
    return r;
}

// true = the reference certainly computes discriminant < 0 for this sphere
__device__ __forceinline__ bool sphere_cull(const RayF32 &r, const float4 g)
{
    const float ocx = r.ox - g.x, ocy = r.oy - g.y, ocz = r.oz - g.z;
    const float cx = __fmaf_rn(ocy, r.dz, -(ocz * r.dy));
    const float cy = __fmaf_rn(ocz, r.dx, -(ocx * r.dz));
    const float cz = __fmaf_rn(ocx, r.dy, -(ocy * r.dx));
    const float q = __fmaf_rn(cz, cz, __fmaf_rn(cy, cy, cx * cx));
    const float t = __fmaf_rn(r.norm_d, g.w, r.slack);
    return q > t * t;
}

// Conservative FP32 miss test for the ground plane (TRT.c:677-695): the reference reports a hit only if
// t = ((point - origin) . n) / (direction . n) > 1e-5.  When numerator and denominator have opposite signs
// the quotient is negative (or -0), so the test is a miss; single precision certifies the two signs unless
// either value is within its rounding error of zero.  Error bounds: every product of the two 3-term dot
// products is computed from float-rounded inputs (relative error 2^-24 each) and accumulated in float, so
// |num_f - num| <= 8*2^-24 * sum_k (|p_k| + |o_k|)|n_k| and |den_f - den| <= 8*2^-24 * sum_k |d_k||n_k|;
// the margins below are twice that, and the reference's own FP64 rounding is ~2^-29 of them.
// NaN/inf make the comparisons false.
__device__ __forceinline__ bool plane_cull(const RayF32 &r)
{
    const float nx = c_scene.ground_normal_f[0], ny = c_scene.ground_normal_f[1], nz = c_scene.ground_normal_f[2];
    const float px = c_scene.ground_point_f[0], py = c_scene.ground_point_f[1], pz = c_scene.ground_point_f[2];
    const float anx = fabsf(nx), any = fabsf(ny), anz = fabsf(nz);
    const float num = __fmaf_rn(pz - r.oz, nz, __fmaf_rn(py - r.oy, ny, (px - r.ox) * nx));
    const float den = __fmaf_rn(r.dz, nz, __fmaf_rn(r.dy, ny, r.dx * nx));
    const float num_scale = __fmaf_rn(fabsf(pz) + fabsf(r.oz), anz, __fmaf_rn(fabsf(py) + fabsf(r.oy), any, (fabsf(px) + fabsf(r.ox)) * anx));
    const float den_scale = __fmaf_rn(fabsf(r.dz), anz, __fmaf_rn(fabsf(r.dy), any, fabsf(r.dx) * anx));
    const float e_num = (16.0f * 5.9604644775390625e-08f) * num_scale;
    const float e_den = (16.0f * 5.9604644775390625e-08f) * den_scale;
    const bool finite = r.usable && (num_scale < 1e30f) && (den_scale < 1e30f);
    return finite && ((num < -e_num && den > e_den) || (num > e_num && den < -e_den));
}

// ---- ray_intersects_sphere (TRT.c:638-672) + the closest-so-far update of trace_ray (TRT.c:807-827) ----
template <bool COUNT>
__device__ __forceinline__ void sphere_exact(const double4 g, int i, const d3 &o, const d3 &d, double two_a, double four_a,
                                             double &closest, int &obj, int &index, d3 &hit, const Tally<COUNT> &tally)
{
    tally.add(CTR_EXACT_SPHERE_TESTS);
    const d3 oc = mk3(o.x - g.x, o.y - g.y, o.z - g.z);
    const double b = 2.0 * dot(oc, d);
    const double c = dot(oc, oc) - g.w;      // g.w = radius*radius, evaluated on the host in double
    const double disc = b * b - four_a * c;  // (4.0*a)*c, scaling by 4 is exact
    if (!(disc < 0.0)) {                     // TRT.c:651
        tally.add(CTR_SPHERE_DISC_OK);
        const double t0 = ieee_div(-b - sqrt(disc), two_a);
        if (t0 > 0.0) {
            tally.add(CTR_SPHERE_T0_POS);
            const d3 p = mk3(o.x + t0 * d.x, o.y + t0 * d.y, o.z + t0 * d.z);
            const d3 back = o - p;
            const double d2 = dot(back, back);
            if (d2 < closest) {
                tally.add(CTR_SPHERE_CLOSEST);
                closest = d2;
                obj = 1;
                index = i;
                hit = p;
            }
        }
    }
}

// exact discriminant sign only — used by the counting build to audit the cull
__device__ __forceinline__ bool exact_disc_negative(const double4 g, const d3 &o, const d3 &d, double four_a)
{
    const d3 oc = mk3(o.x - g.x, o.y - g.y, o.z - g.z);
    const double b = 2.0 * dot(oc, d);
    const double c = dot(oc, oc) - g.w;
    return (b * b - four_a * c) < 0.0;
}

// ---- closest-hit query, the geometric half of trace_ray (TRT.c:805-853) -------------------------------
// obj: 0 none, 1 sphere, 2 ground.  hit = un-pushed intersection point of the closest object.
// Spheres are visited in index order (ties keep the lowest index, strict <), the ground last.
// CULL: 0 = every test in FP64; 1 = FP32 cull records in __constant__; 2 = cull records in global memory.
template <bool COUNT, int CULL>
__device__ __forceinline__ void closest_hit(const RenderParams &P, const d3 &o, const d3 &d, int &obj, int &index,
                                            d3 &hit, const Tally<COUNT> &tally)
{
    double closest = INFINITY;
    obj = 0;
    index = -1;
    const double a = dot(d, d);                  // TRT.c:646, loop-invariant
    const double two_a = 2.0 * a;
    const double four_a = 4.0 * a;
    const int n = c_scene.num_spheres;
    if (COUNT) atomicAdd(&P.counters[CTR_SPHERE_TESTS], (unsigned long long)n);
    bool ground_culled = false;
    if (CULL) {
        const RayF32 rf = ray_to_f32(o, d);
        ground_culled = !COUNT && plane_cull(rf);   // the counting build runs the exact plane test: it counts its branches
        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            // pass 1 (FP32, branch-free, warp-uniform operands): which spheres of this chunk survive.
            // The records are padded to an even count (pad records are masked off below), so the loop
            // needs no remainder handling: two independent tests per trip.
            unsigned int survivors = 0;
#pragma unroll 1
            for (int j = 0; j < cnt; j += 2) {
                const float4 g0 = CULL == 1 ? c_sphere_cull[base + j] : __ldg(&P.sphere_cull[base + j]);
                const float4 g1 = CULL == 1 ? c_sphere_cull[base + j + 1] : __ldg(&P.sphere_cull[base + j + 1]);
                const unsigned int s0 = sphere_cull(rf, g0) ? 0u : 1u;
                const unsigned int s1 = sphere_cull(rf, g1) ? 0u : 2u;
                survivors |= (s0 | s1) << j;
            }
            const unsigned int valid = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
            survivors = rf.usable ? (survivors & valid) : valid;
            if (COUNT) {
                for (int j = 0; j < cnt; j++)
                    if (!((survivors >> j) & 1u) && !exact_disc_negative(ldg_geom(P.sphere_geom, base + j), o, d, four_a))
                        atomicAdd(&P.counters[CTR_CULL_VIOLATIONS], 1ull);
            }
            // pass 2 (FP64, exact): each lane walks its own survivors in index order
            while (survivors) {
                const int j = __ffs(survivors) - 1;
                survivors &= survivors - 1;
                sphere_exact<COUNT>(ldg_geom(P.sphere_geom, base + j), base + j, o, d, two_a, four_a, closest, obj, index, hit, tally);
            }
        }
    } else {
        for (int i = 0; i < n; i++)
            sphere_exact<COUNT>(ldg_geom(P.sphere_geom, i), i, o, d, two_a, four_a, closest, obj, index, hit, tally);
    }
    // ground, TRT.c:677-695 and 831-853
    tally.add(CTR_PLANE_TESTS);
    if (ground_culled) return;
    const d3 gn = mk3(c_scene.ground_normal[0], c_scene.ground_normal[1], c_scene.ground_normal[2]);
    const double denom = dot(d, gn);
    if (fabs(denom) > 0.00001) {
        tally.add(CTR_PLANE_DENOM_OK);
        const d3 to_plane = mk3(c_scene.ground_point[0] - o.x, c_scene.ground_point[1] - o.y, c_scene.ground_point[2] - o.z);
        const double t = ieee_div(dot(to_plane, gn), denom);
        if (t > 0.00001) {
            tally.add(CTR_PLANE_T_POS);
            const d3 p = mk3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
            const d3 back = o - p;
            const double d2 = dot(back, back);
            if (d2 < closest) {
                tally.add(CTR_PLANE_CLOSEST);
                obj = 2;
                // checker parity, TRT.c:850
                index = x86_int(floor(p.x) + floor(p.z)) & 1;
                hit = p;
            }
        }
    }
}

// hit point pushed back toward the ray origin by EPSILON, TRT.c:871-874
__device__ __forceinline__ d3 push_back(const d3 &o, const d3 &hit)
{
    d3 back = unit(o - hit);
    back = back * TRT_EPSILON;
    return hit + back;
}

__device__ __forceinline__ const DevMaterial *surface_material(const RenderParams &P, int obj, int index)
{
    return (obj == 1) ? &P.sphere_mat[index] : (index ? &c_scene.ground_odd : &c_scene.ground_even);
}

// ---- K1 ------------------------------------------------------------------------------------------------
// One loop, one ray per trip.  Every step of the trip appears ONCE in the instruction stream (one
// normalisation of the ray direction, one closest-hit query, one push-back, one normal/sky normalisation,
// one accumulate, one reflect, one shadow-ray setup) and is predicated by the lane's state, so primary,
// bounce and shadow rays of different lanes execute the same instructions, and the loop body stays small
// enough for the instruction caches (an earlier version that inlined trace_ray's callees per call site was
// 41 KB of SASS and stalled on instruction fetch as soon as occupancy was raised; profiles/).
template <bool COUNT, int CULL>
__global__ void __launch_bounds__(CTA_THREADS, TRT_MIN_CTAS_PER_SM) k_render(const RenderParams P)
{
    __shared__ double s_byte_to_unit[256]; // k/255.0 (TRT.c:866), evaluated on the host in double
    __shared__ unsigned int s_tile[WARPS_PER_CTA];
    for (int k = threadIdx.x; k < 256; k += CTA_THREADS) s_byte_to_unit[k] = P.byte_to_unit[k];
    __syncthreads();

    const Tally<COUNT> tally{P.counters};
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int band_rows = P.row1 - P.row0;
    const int tiles_x = (P.width + TILE_W - 1) / TILE_W;
    const int tiles_y = (band_rows + TILE_H - 1) / TILE_H;
    const unsigned int num_tiles = (unsigned int)(tiles_x * tiles_y);

    const double sw = c_scene.screen_width, sh = c_scene.screen_height;
    const double pixel_w = sw / P.width;   // TRT.c:981
    const double pixel_h = sh / P.height;  // TRT.c:982
    const int num_dir = c_scene.num_dir;
    const int num_lights = num_dir + c_scene.num_point;
    // (double)col / (double)width and (double)row / (double)height (TRT.c:987-988) through shared reciprocals:
    // numerators are integers >= 1 (0 is special-cased), far inside the fast path of the IEEE division
    const Reciprocal inv_w = reciprocal_of((double)P.width), inv_h = reciprocal_of((double)P.height);

    for (;;) {
        if (lane == 0) s_tile[warp] = atomicAdd(P.tile_counter, 1u);
        __syncwarp();
        const unsigned int tile = s_tile[warp];
        __syncwarp();
        if (tile >= num_tiles) break;
        const int ty = (int)(tile / (unsigned)tiles_x), tx = (int)(tile % (unsigned)tiles_x);
        const int col = tx * TILE_W + (lane & (TILE_W - 1));
        const int brow = ty * TILE_H + (lane >> 3); // band-local row
        const int row = P.row0 + brow;
        if (col >= P.width || brow >= band_rows) continue; // lane idles for this tile
        tally.add(CTR_PIXELS);

        // per-pixel part of the primary ray, TRT.c:987-988
        const double fx = col ? div_by_unchecked((double)col, inv_w) : 0.0;
        const double fy = row ? div_by_unchecked((double)row, inv_h) : 0.0;
        const double sx0 = (fx * sw - sw / 2.0);
        const double sy0 = -(fy * sh - sh / 2.0);

        d3 average = mk3(0.0, 0.0, 0.0);
        // ---- per-lane ray state ------------------------------------------------------------------
        int k = 0;                         // sample index (ray_num)
        int phase = PH_MAIN;
        int bounces = 0;
        int light = 0;                     // index of the light whose shadow ray is in flight
        int surf_obj = 0, surf_index = 0;  // what the main ray hit
        double weight = 1.0, weight_sum = 0.0, light_d2 = 0.0, intensity = 0.0;
        d3 o = mk3(0, 0, 0), d = mk3(0, 0, 0), d_main = d, nrm = d, lit = d, sample = d;
        bool fresh = true;                 // start the next sample of this pixel
        bool raw_dir = false;              // d still has to be normalised
        unsigned int trips = 0;            // closest-hit queries of this pixel (cost estimate for band balancing)

        for (;;) {
            if (fresh) {
                if (k == TRT_RAYS_PER_PIXEL) break;
                // primary ray of sample k, TRT.c:987-1016
                tally.add(CTR_SAMPLES);
                const double sx = sx0 + c_scene.sub_dx[k] * pixel_w;
                const double sy = sy0 + c_scene.sub_dy[k] * pixel_h;
                const double sz = -c_scene.screen_distance;
                d3 sp = mk3(0.0, 0.0, 0.0);
                sp = sp + mk3(c_scene.bx[0] * sx, c_scene.bx[1] * sx, c_scene.bx[2] * sx);
                sp = sp + mk3(c_scene.by[0] * sy, c_scene.by[1] * sy, c_scene.by[2] * sy);
                sp = sp + mk3(c_scene.bz[0] * sz, c_scene.bz[1] * sz, c_scene.bz[2] * sz);
                o = mk3(c_scene.eye[0], c_scene.eye[1], c_scene.eye[2]);
                d = sp - o;                 // TRT.c:1005 (origin subtracted from an untranslated vector)
                raw_dir = true;
                sample = mk3(0.0, 0.0, 0.0);
                bounces = 0;
                weight = 1.0;
                weight_sum = 0.0;
                phase = PH_MAIN;
                fresh = false;
            }
            // (1) the one normalisation of a ray direction: primary (TRT.c:1008), reflected (1055), point-light (933)
            if (raw_dir) d = unit(d);

            // (2) the one closest-hit query
            trips++;
            int obj, index;
            d3 hit;
            const bool is_main = phase == PH_MAIN;
            tally.add(CTR_TRACE_CALLS);
            if (is_main) tally.add(CTR_BOUNCE_ITERS);
            closest_hit<COUNT, CULL>(P, o, d, obj, index, hit, tally);
            if (obj != 0) tally.add(CTR_TRACE_HITS);
            else { tally.add(CTR_SKY_LOOKUPS); if (!is_main) tally.add(CTR_SKY_SKIPPED); }

            // (3) the one push-back (TRT.c:871-874); a directional light's shadow ray only needs hit / no hit
            const bool point_shadow = !is_main && light >= num_dir;
            d3 at = o;
            if (obj != 0 && (is_main || point_shadow)) at = push_back(o, hit);

            bool shade_done = false, sample_done = false, have_colour = false;
            d3 colour = mk3(0.0, 0.0, 0.0);
            if (is_main) {
                // (4) the one normal / sky normalisation: unit(hit - centre) or unit(ground normal) on a hit
                //     (TRT.c:878), unit(direction) once more for the cubemap lookup on a miss (TRT.c:702)
                d3 v = d;
                if (obj == 1) {
                    const double4 g = ldg_geom(P.sphere_geom, index);
                    v = mk3(hit.x - g.x, hit.y - g.y, hit.z - g.z);   // TRT.c:824
                } else if (obj == 2) {
                    v = mk3(c_scene.ground_normal[0], c_scene.ground_normal[1], c_scene.ground_normal[2]);
                }
                const d3 u = unit(v);
                if (obj == 0) {
                    // sky: TRT.c:858-867; the sample ends here
                    int face;
                    const int texel = sky_texel_index(u, c_scene.sky_dim, face);
                    const uchar4 t = __ldg(&P.sky[(size_t)face * (size_t)c_scene.sky_face_stride + (size_t)texel]);
                    colour = mk3(s_byte_to_unit[t.x], s_byte_to_unit[t.y], s_byte_to_unit[t.z]);
                    have_colour = true;
                    sample_done = true;
                } else {
                    // surface hit: remember it, start the light loop (apply_lighting, TRT.c:894-963)
                    tally.add(CTR_LIGHTING_CALLS);
                    surf_obj = obj;
                    surf_index = index;
                    nrm = u;
                    d_main = d;
                    o = at;
                    lit = mk3(0.0, 0.0, 0.0);
                    light = 0;
                    if (num_lights == 0) shade_done = true;
                    else phase = PH_SHADOW;
                }
            } else {
                // ---- a shadow ray came back: add this light's contribution ---------------------------
                bool open = (obj == 0);
                double f = 1.0;
                const double *lc;
                if (light < num_dir) {                                    // TRT.c:908-921
                    lc = c_scene.dir[light].color;
                } else {                                                  // TRT.c:939-954
                    if (!open) {
                        const d3 to_blocker = at - o;
                        open = light_d2 < dot(to_blocker, to_blocker);
                    }
                    lc = c_scene.point[light - num_dir].color;
                    f = intensity;
                }
                if (open) {
                    const DevMaterial *m = surface_material(P, surf_obj, surf_index);
                    const double lambert = fmin(dot(nrm, d), 1.0);
                    if (light >= num_dir) f = f * lambert; else f = lambert;
                    d3 diffuse = mk3(lc[0] * f, lc[1] * f, lc[2] * f);
                    diffuse = hadamard(diffuse, mk3(m->color[0], m->color[1], m->color[2]));
                    lit = lit + diffuse;
                }
                light++;
                if (light == num_lights) shade_done = true;
            }

            if (shade_done) {
                colour.x = clampd(lit.x, 0.0, 1.0);                       // TRT.c:960
                colour.y = clampd(lit.y, 0.0, 1.0);
                colour.z = clampd(lit.z, 0.0, 1.0);
                have_colour = true;
            }
            // (5) the one accumulate, TRT.c:1034-1035, 1051
            if (have_colour) {
                weight_sum += weight;
                colour = colour * weight;
                sample = sample + colour;
            }
            if (shade_done) {
                weight *= surface_material(P, surf_obj, surf_index)->reflectivity;   // TRT.c:1041-1042
                bounces++;
                if (bounces < TRT_BOUNCE_LIMIT && weight > 0.00001) {      // loop condition, TRT.c:1018
                    // (6) the one reflect, TRT.c:627-633; normalised at the top of the next trip
                    const double dn = dot(d_main, nrm);
                    d = mk3(d_main.x - 2.0 * dn * nrm.x, d_main.y - 2.0 * dn * nrm.y, d_main.z - 2.0 * dn * nrm.z);
                    raw_dir = true;
                    phase = PH_MAIN;                                      // the origin o already is the surface point
                } else {
                    sample_done = true;
                }
            } else if (!sample_done) {
                // (7) the one shadow-ray setup: aim at light `light` from the surface point o
                if (light < num_dir) {
                    const DevLightDir &Ld = c_scene.dir[light];
                    d = mk3(Ld.L[0], Ld.L[1], Ld.L[2]);                    // unit(-direction), host-evaluated (TRT.c:903-904)
                    raw_dir = false;
                } else {
                    const DevLightPoint &Lp = c_scene.point[light - num_dir];
                    d = mk3(Lp.pos[0] - o.x, Lp.pos[1] - o.y, Lp.pos[2] - o.z);   // TRT.c:929
                    light_d2 = dot(d, d);
                    intensity = clampd(ieee_div(Lp.intensity, light_d2), 0.0, 1.0);         // TRT.c:931
                    raw_dir = true;
                }
            }
            if (sample_done) {
                if (COUNT) atomicAdd(&P.counters[CTR_BOUNCE_HIST0 + bounces], 1ull);
                sample = sample * ieee_div(1.0, weight_sum);                     // TRT.c:1061
                average = average + sample;                               // TRT.c:1063
                k++;
                fresh = true;
            }
        }

        if (P.row_cost) atomicAdd(&P.row_cost[brow], trips);
        average = average * (1.0 / TRT_RAYS_PER_PIXEL);                   // TRT.c:1065
        const size_t pix = (size_t)brow * (size_t)P.width + (size_t)col;
        if (P.pixels) {
            P.pixels[pix * 3 + 0] = average.x;
            P.pixels[pix * 3 + 1] = average.y;
            P.pixels[pix * 3 + 2] = average.z;
        }
        if (P.quant) {
            // the quantisation of buffered_draw_screen, TRT.c:1157-1163: truncation toward zero
            uchar4 q;
            q.x = (unsigned char)x86_int(average.x * 255);
            q.y = (unsigned char)x86_int(average.y * 255);
            q.z = (unsigned char)x86_int(average.z * 255);
            q.w = 0;
            P.quant[pix] = q;
        }
    }
}

// ---- unit-level probe: trace_ray (TRT.c:793-889) for an array of rays, all out-params ---------------
// Used by the parity tests to compare single queries (hit kind, pushed-back point, unit normal,
// material incl. the skybox colour on a miss) with the reference, not only whole frames.
// out: 11 doubles per ray = kind, point[3], normal[3], colour[3], reflectivity
__global__ void k_probe_trace(const RenderParams P, const double *__restrict__ rays, int n, double *__restrict__ out)
{
    __shared__ double s_byte_to_unit[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_byte_to_unit[k] = P.byte_to_unit[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 o = mk3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]);
    const d3 d = mk3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
    const Tally<false> tally{nullptr};
    int obj, index;
    d3 hit;
    if (c_scene.filter_in_const) closest_hit<false, 1>(P, o, d, obj, index, hit, tally);
    else closest_hit<false, 2>(P, o, d, obj, index, hit, tally);
    d3 point, normal, colour;
    double reflectivity = 0.0;
    if (obj == 0) {
        point = o;
        normal = d;
        int face;
        const int texel = sky_texel_index(unit(d), c_scene.sky_dim, face);
        const uchar4 t = __ldg(&P.sky[(size_t)face * (size_t)c_scene.sky_face_stride + (size_t)texel]);
        colour = mk3(s_byte_to_unit[t.x], s_byte_to_unit[t.y], s_byte_to_unit[t.z]);
    } else {
        const DevMaterial *m;
        if (obj == 1) {
            const double4 g = ldg_geom(P.sphere_geom, index);
            normal = mk3(hit.x - g.x, hit.y - g.y, hit.z - g.z);
            m = &P.sphere_mat[index];
        } else {
            normal = mk3(c_scene.ground_normal[0], c_scene.ground_normal[1], c_scene.ground_normal[2]);
            m = index ? &c_scene.ground_odd : &c_scene.ground_even;
        }
        colour = mk3(m->color[0], m->color[1], m->color[2]);
        reflectivity = m->reflectivity;
        point = push_back(o, hit);
    }
    normal = unit(normal);
    double *r = out + (size_t)i * 11;
    r[0] = (double)obj;
    r[1] = point.x; r[2] = point.y; r[3] = point.z;
    r[4] = normal.x; r[5] = normal.y; r[6] = normal.z;
    r[7] = colour.x; r[8] = colour.y; r[9] = colour.z;
    r[10] = reflectivity;
}

// get_skybox_color (TRT.c:700-789) for an array of directions: face, texel index, r, g, b per entry
__global__ void k_probe_sky(const RenderParams P, const double *__restrict__ dirs, int n, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 d = mk3(dirs[i * 3 + 0], dirs[i * 3 + 1], dirs[i * 3 + 2]);
    int face;
    const int texel = sky_texel_index(unit(d), c_scene.sky_dim, face);
    const uchar4 t = __ldg(&P.sky[(size_t)face * (size_t)c_scene.sky_face_stride + (size_t)texel]);
    out[i * 5 + 0] = face;
    out[i * 5 + 1] = texel;
    out[i * 5 + 2] = t.x;
    out[i * 5 + 3] = t.y;
    out[i * 5 + 4] = t.z;
}

// ---- self-test: shared-reciprocal division (trt_device.cuh) against the IEEE division --------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long &state)
{
    unsigned long long z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double random_double(unsigned long long &state, int exp_lo, int exp_hi)
{
    const unsigned long long bits = mix64(state);
    const unsigned long long mant = bits & 0x000FFFFFFFFFFFFFull;
    const int e = exp_lo + (int)((bits >> 52) % (unsigned long long)(exp_hi - exp_lo + 1));
    const unsigned long long sign = (mix64(state) & 1ull) << 63;
    return __longlong_as_double((long long)(sign | ((unsigned long long)(e + 1023) << 52) | mant));
}

__global__ void k_selftest_division(unsigned long long seed, int iters, unsigned long long *mismatches)
{
    unsigned long long state = seed + 0x1000003ull * (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad = 0;
    for (int it = 0; it < iters; it++) {
        double a[3], b;
        const int mode = it & 3;
        if (mode == 0) {            // what unit() does: components over their own length, moderate scale
            const int e = -20 + (int)(mix64(state) % 41);
            a[0] = random_double(state, e - 3, e); a[1] = random_double(state, e - 3, e); a[2] = random_double(state, e - 30, e);
            b = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        } else if (mode == 1) {     // unrelated operands, wide exponent range
            a[0] = random_double(state, -500, 500); a[1] = random_double(state, -500, 500); a[2] = random_double(state, -1022, 1023);
            b = random_double(state, -500, 500);
        } else if (mode == 2) {     // quotients near the overflow / underflow guards and zeros
            a[0] = random_double(state, -1022, -960); a[1] = 0.0; a[2] = random_double(state, 900, 1023);
            b = random_double(state, -60, 60);
        } else {                    // divisors with extreme significands (all ones / all zeros)
            a[0] = random_double(state, -4, 4); a[1] = random_double(state, -4, 4); a[2] = 1.0;
            const int e = -8 + (int)(mix64(state) % 17);
            const unsigned long long m = (mix64(state) & 1ull) ? 0x000FFFFFFFFFFFFFull : (mix64(state) & 0xFull);
            b = __longlong_as_double((long long)(((unsigned long long)(e + 1023) << 52) | m));
        }
        if (mode == 0 || mode == 2) {
            // the vector form used by the renderer: unit() against sqrt + three IEEE divisions
            const d3 v = mode == 0 ? mk3(a[0], a[1], a[2]) : mk3(a[0] * 1e-30, a[2] * 1e-250, a[1]);
            const d3 u = unit(v);
            const double len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
            d3 w = v;
            if (len > 0.0001) { w.x = __ddiv_rn(v.x, len); w.y = __ddiv_rn(v.y, len); w.z = __ddiv_rn(v.z, len); }
            if (__double_as_longlong(u.x) != __double_as_longlong(w.x) && !(u.x != u.x && w.x != w.x)) bad++;
            if (__double_as_longlong(u.y) != __double_as_longlong(w.y) && !(u.y != u.y && w.y != w.y)) bad++;
            if (__double_as_longlong(u.z) != __double_as_longlong(w.z) && !(u.z != u.z && w.z != w.z)) bad++;
        }
        const Reciprocal inv = reciprocal_of(b);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double fast = div_by(a[k], inv);
            const double ieee = __ddiv_rn(a[k], b);
            if (__double_as_longlong(fast) != __double_as_longlong(ieee) && !(fast != fast && ieee != ieee)) bad++;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------------------
// host side of this TU: scene upload (constant memory lives here) and the launcher

static void die(cudaError_t e, const char *file, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "%s:%d: CUDA error: %s\n", file, line, cudaGetErrorString(e));
        exit(1);
    }
}
#define CK(x) die((x), __FILE__, __LINE__)

void upload_scene_constants(const DevScene &scene, const float4 *cull, int count, cudaStream_t stream)
{
    CK(cudaMemcpyToSymbolAsync(c_scene, &scene, sizeof(DevScene), 0, cudaMemcpyHostToDevice, stream));
    if (cull && count > 0 && count <= TRT_MAX_CONST_SPHERES + 2)
        CK(cudaMemcpyToSymbolAsync(c_sphere_cull, cull, sizeof(float4) * (size_t)count, 0, cudaMemcpyHostToDevice, stream));
}

int render_ctas_per_sm()
{
    static int cached = 0;
    if (!cached) {
        int n = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render<false, 1>, CTA_THREADS, 0));
        cached = n > 0 ? n : 1;
    }
    return cached;
}

void launch_render(const RenderParams &p, bool count, int cull, int num_sms, cudaStream_t stream)
{
    CK(cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), stream));
    const int band_rows = p.row1 - p.row0;
    if (band_rows <= 0 || p.width <= 0) return;
    const long long tiles = (long long)((p.width + TILE_W - 1) / TILE_W) * ((band_rows + TILE_H - 1) / TILE_H);
    long long want = (tiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    long long grid = (long long)num_sms * render_ctas_per_sm();   // persistent: every CTA slot of the chip, once
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    dim3 g((unsigned)grid), b(CTA_THREADS);
    if (count) {
        if (cull == 1) k_render<true, 1><<<g, b, 0, stream>>>(p);
        else if (cull == 2) k_render<true, 2><<<g, b, 0, stream>>>(p);
        else k_render<true, 0><<<g, b, 0, stream>>>(p);
    } else {
        if (cull == 1) k_render<false, 1><<<g, b, 0, stream>>>(p);
        else if (cull == 2) k_render<false, 2><<<g, b, 0, stream>>>(p);
        else k_render<false, 0><<<g, b, 0, stream>>>(p);
    }
    CK(cudaGetLastError());
}

unsigned long long run_selftest_division(unsigned long long seed, int ctas, int iters, unsigned long long *d_scratch, cudaStream_t stream)
{
    CK(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long), stream));
    k_selftest_division<<<ctas, 256, 0, stream>>>(seed, iters, d_scratch);
    CK(cudaGetLastError());
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, d_scratch, sizeof bad, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return bad;
}

void launch_probe_trace(const RenderParams &p, const double *d_rays, int n, double *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_trace<<<(n + 127) / 128, 128, 0, stream>>>(p, d_rays, n, d_out);
    CK(cudaGetLastError());
}

void launch_probe_sky(const RenderParams &p, const double *d_dirs, int n, int *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_sky<<<(n + 127) / 128, 128, 0, stream>>>(p, d_dirs, n, d_out);
    CK(cudaGetLastError());
}

} // namespace trt
