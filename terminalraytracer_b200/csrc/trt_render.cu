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
//   * the `view` vector (1029) and the specularity field (118), both never read,
//   * intersection tests whose outcome a single-precision certificate (trt_cert.h) proves in advance.
//
// Execution model — a per-warp WAVEFRONT, so that every instruction stream is uniform across the warp:
//   * persistent CTAs; each WARP pulls 8x4-pixel tiles (x 10 samples = 320 primary rays) from an atomic counter;
//   * PRODUCE step (one per sample index k, lane = pixel): primary ray -> closest hit.  Spheres are tested
//     exactly, but only those the TILE certificate could not rule out for the whole tile (usually none or one),
//     with the eye-relative terms of the quadratic precomputed on the host.  A sample that sees the sky is
//     finished on the spot; a sample that hits a surface becomes a 40-byte record (direction, hit parameter, ids)
//     in warp-private shared-memory ring A;
//   * CONSUME step (whenever a ring holds 32 records, lane = record): push-back, normal, then ONE loop over the
//     record's queries — a shadow ray per light (float certificates decide open / blocked for most of them; the rest
//     walk their few surviving spheres exactly), then the reflected ray — shading, accumulation.  The reflected ray's
//     hit becomes a 104-byte record in ring B; its miss finishes the sample with a sky lookup.  First-generation
//     hits have their own ring because a tile's ground hits share tile-level ("patch") certificates;
//   * finished samples are parked in an L2-resident per-warp slice of global memory and summed per pixel in sample
//     order at the end of the tile (TRT.c:1063 adds them in that order; floating-point addition is not associative);
//   * scenes of more than 32 spheres are ordered along a k-d tree with bounding balls per 32 and per 8 spheres
//     (CULL == 2): whole chunks of the query loop are skipped when no lane's ray can reach their ball, and inside a
//     chunk the float classification is dealt across the warp per (ray, group of 8) — query_clustered, classify_chunk.
// Compared with one-ray-per-trip state machines, no lane ever executes another lane's phase under predication:
// the rings keep full warps of like work together however the bounce counts diverge.  DESIGN.md 4.2 has the
// measurements and the list of variants that were tried.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "trt_device.cuh"
#include "trt_internal.h"

namespace trt {

__constant__ DevScene c_scene;
// (The certificate records — two spheres per record so that the classification runs on the packed FP32 pipe, FFMA2/FADD2/FMUL2
// of sm_100: one issue slot, two spheres; .x = sphere 2p, .y = sphere 2p+1, a pad sphere has r = 0 — live in global memory,
// RenderParams::cull_pairs; small scenes copy theirs into shared memory at kernel start.)

#ifndef TRT_TILES_BOTTOM_UP
#define TRT_TILES_BOTTOM_UP 1
#endif
#ifndef TRT_FUSED_SHADOWS
#define TRT_FUSED_SHADOWS 1     // small scenes, 1 + 1 lights: both shadow rays of a record classified in one loop (0: A/B runs)
#endif
#ifndef TRT_SUM_UNROLL
#define TRT_SUM_UNROLL 10       // per-pixel sum of the ten samples at the end of a tile: unroll factor (code size vs loop overhead)
#endif
#ifndef TRT_TILE_PREPASS
#define TRT_TILE_PREPASS 1      // 0: every flavour takes its tile certificates inside k_render (A/B runs)
#endif
constexpr int SUM_UNROLL = TRT_SUM_UNROLL;
constexpr int TILE_W = 8;
constexpr int TILE_H = 4;
#ifndef TRT_WARPS_PER_CTA
#define TRT_WARPS_PER_CTA 4
#endif
#ifndef TRT_MIN_CTAS_PER_SM
#define TRT_MIN_CTAS_PER_SM 4
#endif
constexpr int WARPS_PER_CTA = TRT_WARPS_PER_CTA;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr int QCAP = 64;                                  // ring capacity: at most 31 queued + 32 pushed by one step
constexpr int TILE_SAMPLES = 32 * TRT_RAYS_PER_PIXEL;
constexpr int TMASK_WORDS = 128;                          // tile certificate masks: scenes of up to 4096 spheres

// what a closest-hit query is for
enum QueryMode : int {
    Q_CLOSEST = 0, // bounce rays: the hit itself is needed
    Q_DIR = 1,     // directional-light shadow ray: any hit blocks (TRT.c:907-908)
    Q_POINT = 2    // point-light shadow ray: only a hit closer than the light blocks (TRT.c:936-941)
};

// warp-private shared memory: two hit-record rings (structure of arrays: conflict-free for lane = slot) and the
// tile / patch certificate masks.  Ring A holds FIRST-generation hits (primary rays: origin = eye, nothing
// accumulated yet, so a record is just direction + hit parameter), ring B the hits of bounce rays.  They are
// consumed separately because a tile's first-generation hits share tile-level (patch) certificates.
constexpr int PATCH_MAX_SPHERES = 32;                     // patch certificates: one mask word per query
constexpr int PATCH_QUERIES = 2 * TRT_MAX_LIGHTS + 1;     // every directional light, every point light, the bounce
constexpr int ANSI_STAGE_STRIDE = 208;                      // fused encode: 8 cells x 25 bytes + '\n', rounded up to 16
struct WarpShared {
    // ring A
    double a_dx[QCAP], a_dy[QCAP], a_dz[QCAP];   // unit direction of the primary ray
    double a_t[QCAP];                            // hit parameter: hit point = eye + t d (TRT.c:663-665 / 690-692)
    unsigned int a_meta[QCAP];                   // pixel lane | sample k << 5 | object kind << 13
    int a_index[QCAP];                           // sphere index of the hit
    // ring B
    double ox[QCAP], oy[QCAP], oz[QCAP];   // origin of the ray that hit
    double dx[QCAP], dy[QCAP], dz[QCAP];   // its unit direction
    double t[QCAP];                        // hit parameter
    double sr[QCAP], sg[QCAP], sb[QCAP];   // colour accumulated by the sample so far (TRT.c:1051)
    double w[QCAP], ws[QCAP];              // weight, weight_sum (TRT.c:1017, 1034)
    unsigned int meta[QCAP];               // pixel lane | sample k << 5 | bounces << 9 | object kind << 13
    int index[QCAP];
    unsigned int tmask[TMASK_WORDS];       // spheres the tile certificate could not rule out for primary rays
    unsigned int pmask[PATCH_QUERIES + 1]; // patch certificates: spheres each query of a first-generation hit can reach
    unsigned int ansi_stage[TILE_H * ANSI_STAGE_STRIDE / 4];   // fused encode: one staging row of cells per tile row
};
// k/255 table, then (small scenes) the certificate records of the single chunk: 17 pairs x 32 B.  From shared memory a pair
// is two 16-byte loads; from __constant__ memory with a run-time index it is four 8-byte ones.
constexpr size_t SMEM_PAIRS_OFFSET = 256 * sizeof(double);
constexpr int SMEM_PAIRS = TRT_CLUSTER_MIN_SPHERES / 2 + 1;
constexpr size_t SMEM_TABLE_BYTES = SMEM_PAIRS_OFFSET + ((SMEM_PAIRS * sizeof(CullPair) + 127) & ~(size_t)127);
constexpr size_t SMEM_BYTES = SMEM_TABLE_BYTES + WARPS_PER_CTA * sizeof(WarpShared);
// Many-sphere flavours (CULL == 2) add a per-warp scratch behind the rings (classify_chunk): the float records of the current
// query's rays, the work list of the chunk being classified and the survivor words.  The other flavours' layout does not change.
struct ClusterScratch {
    float4 ray[32][2];            // per ray (lane): (ox, oy, oz, slack), (dx, dy, dz, far limit)
    float near_limit[32];
    unsigned int surv[32];        // per ray: spheres of the current chunk its certificate could not rule out
    unsigned int blk[32];         // per ray: a sphere of the current chunk certainly blocks the light
    unsigned char item[128];      // work list of the current chunk: ray << 2 | group of 8
};
constexpr size_t SMEM_BYTES_CLUSTERED = SMEM_BYTES + WARPS_PER_CTA * sizeof(ClusterScratch);
template <int CULL> constexpr size_t smem_bytes_of() { return CULL == 2 ? SMEM_BYTES_CLUSTERED : SMEM_BYTES; }
static_assert(SMEM_BYTES % 16 == 0, "the clustered scratch starts 16-byte aligned");
static_assert(TRT_MIN_CTAS_PER_SM * (SMEM_BYTES_CLUSTERED + 1024) <= 228 * 1024, "the resident CTAs' shared memory (plus 1 KB each the system reserves) must fit an SM's 228 KB");

// (cx,cy,cz,r*r) of sphere i through the read-only path
__device__ __forceinline__ double4 ldg4(const double4 *geom, int i)
{
    const double2 *p = reinterpret_cast<const double2 *>(geom + i);
    const double2 lo = __ldg(p), hi = __ldg(p + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ CullPair ldg_pair(const CullPair *pairs, int p)
{
    const float4 *q = reinterpret_cast<const float4 *>(pairs + p);
    const float4 lo = __ldg(q), hi = __ldg(q + 1);
    CullPair g;
    g.cx = make_float2(lo.x, lo.y); g.cy = make_float2(lo.z, lo.w);
    g.cz = make_float2(hi.x, hi.y); g.r = make_float2(hi.z, hi.w);
    return g;
}

template <bool COUNT>
struct Tally {
    unsigned long long *g;
    __device__ __forceinline__ void add(int id, unsigned long long n = 1ull) const
    {
        if (COUNT) atomicAdd(&g[id], n);
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

// colour of the sky in direction d (a unit vector; the reference normalises it once more, TRT.c:702)
#ifndef TRT_SKY_INLINE
#define TRT_SKY_INLINE __noinline__
#endif
static __device__ TRT_SKY_INLINE d3 sky_colour_of(const uchar4 *sky, const double *s_byte_to_unit, d3 d)
{
    // face and texel are a decision: certified in float for ~99.6 % of the directions (trt_cert_sky_texel), else
    // the reference's own double arithmetic
    int face, texel;
    if (TRT_UNLIKELY(!(c_scene.filter_enabled && trt_cert_sky_texel((float)d.x, (float)d.y, (float)d.z, c_scene.sky_dim, &face, &texel))))
        texel = sky_texel_index(unit(d), c_scene.sky_dim, face);
    TRT_BOUND(face >= 0 && face < 6 && texel >= 0 && texel < c_scene.sky_face_stride, 5);
    const uchar4 t = __ldg(&sky[(size_t)face * (size_t)c_scene.sky_face_stride + (size_t)texel]);
    return mk3(s_byte_to_unit[t.x], s_byte_to_unit[t.y], s_byte_to_unit[t.z]);   // TRT.c:866
}
// counting build: the texel certificate against the exact lookup
__device__ __noinline__ bool sky_certificate_disagrees(d3 d)
{
    int face, texel, face2;
    if (!(c_scene.filter_enabled && trt_cert_sky_texel((float)d.x, (float)d.y, (float)d.z, c_scene.sky_dim, &face, &texel))) return false;
    const int texel2 = sky_texel_index(unit(d), c_scene.sky_dim, face2);
    return face != face2 || texel != texel2;
}
__device__ __forceinline__ d3 sky_colour(const RenderParams &P, const double *s_byte_to_unit, const d3 &d)
{
    return sky_colour_of(P.sky, s_byte_to_unit, d);
}

// ---- exact tests: the reference's operations in the reference's order -------------------------------------

// ray_intersects_sphere (TRT.c:638-672) with oc = origin - centre and c = oc.oc - r*r given, plus the
// closest-so-far update of trace_ray (TRT.c:807-827).  Keeps the hit PARAMETER; the point is o + t d.
// Spheres may be visited in any order (k-d-sorted scenes, per-lane survivor walks): `oi` is the sphere's index in
// the reference's array, and a hit replaces the closest one if it is closer or equally close with a lower reference
// index — exactly what the reference's index-ordered scan with strict < keeps (TRT.c:810).
template <bool COUNT, bool ANY_ORDER = true>
__device__ __forceinline__ void sphere_exact_oc(const d3 &oc, double c, int i, int oi, const d3 &o, const d3 &d, double two_a, double four_a,
                                                double &closest, int &obj, int &index, int &best_oi, double &t_hit, const Tally<COUNT> &tally)
{
    const double b = 2.0 * dot(oc, d);
    const double disc = b * b - four_a * c;  // (4.0*a)*c, scaling by 4 is exact
    if (!(disc < 0.0)) {                     // TRT.c:651
        tally.add(CTR_SPHERE_DISC_OK);
        const double t0 = ieee_div(-b - sqrt(disc), two_a);
        if (t0 > 0.0) {
            tally.add(CTR_SPHERE_T0_POS);
            const d3 p = mk3(o.x + t0 * d.x, o.y + t0 * d.y, o.z + t0 * d.z);
            const d3 back = o - p;
            const double d2 = dot(back, back);
            if (d2 < closest || (ANY_ORDER && d2 == closest && obj == 1 && oi < best_oi)) {
                tally.add(CTR_SPHERE_CLOSEST);
                closest = d2;
                obj = 1;
                index = i;
                best_oi = oi;
                t_hit = t0;
            }
        }
    }
}

template <bool COUNT, bool ANY_ORDER = true>
__device__ __forceinline__ void sphere_exact(const double4 g, int i, int oi, const d3 &o, const d3 &d, double two_a, double four_a,
                                             double &closest, int &obj, int &index, int &best_oi, double &t_hit, const Tally<COUNT> &tally)
{
    const d3 oc = mk3(o.x - g.x, o.y - g.y, o.z - g.z);
    const double c = dot(oc, oc) - g.w;      // g.w = radius*radius, evaluated on the host in double
    sphere_exact_oc<COUNT, ANY_ORDER>(oc, c, i, oi, o, d, two_a, four_a, closest, obj, index, best_oi, t_hit, tally);
}

// ray_intersects_plane (TRT.c:677-695) + the ground branch of trace_ray (TRT.c:831-853); `num` is
// dot(ground point - origin, normal) (TRT.c:684-685), passed in because several callers share it
template <bool COUNT>
__device__ __forceinline__ void plane_exact_num(double num, const d3 &o, const d3 &d, double &closest, int &obj, double &t_hit,
                                                const Tally<COUNT> &tally)
{
    const d3 gn = mk3(c_scene.ground_normal[0], c_scene.ground_normal[1], c_scene.ground_normal[2]);
    const double denom = dot(d, gn);
    if (fabs(denom) > 0.00001) {
        tally.add(CTR_PLANE_DENOM_OK);
        const double t = ieee_div(num, denom);
        if (t > 0.00001) {
            tally.add(CTR_PLANE_T_POS);
            const d3 p = mk3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
            const d3 back = o - p;
            const double d2 = dot(back, back);
            if (d2 < closest) {
                tally.add(CTR_PLANE_CLOSEST);
                closest = d2;
                obj = 2;
                t_hit = t;
            }
        }
    }
}

__device__ __forceinline__ double plane_numerator(const d3 &o)
{
    const d3 gn = mk3(c_scene.ground_normal[0], c_scene.ground_normal[1], c_scene.ground_normal[2]);
    const d3 to_plane = mk3(c_scene.ground_point[0] - o.x, c_scene.ground_point[1] - o.y, c_scene.ground_point[2] - o.z);
    return dot(to_plane, gn);
}

// trace_ray's geometric half (TRT.c:805-853) over ALL objects in the reference's order: spheres by index
// (strict <, so ties keep the lowest index), the ground last.  obj: 0 none, 1 sphere, 2 ground.
// This is the all-FP64 path: the counting build, trt_set_cull(0) and the audit of the certificates use it.
template <bool COUNT>
__device__ __noinline__ void query_reference(const RenderParams &P, const d3 &o, const d3 &d, int &obj, int &index, double &t_hit,
                                             const Tally<COUNT> &tally)
{
    double closest = INFINITY;
    obj = 0;
    index = -1;
    t_hit = 0.0;
    const double a = dot(d, d);                  // TRT.c:646, loop-invariant
    const double two_a = 2.0 * a, four_a = 4.0 * a;
    const int n = c_scene.num_spheres;
    tally.add(CTR_SPHERE_TESTS, (unsigned long long)n);
    int best_oi = -1;
    for (int oi = 0; oi < n; oi++) {
        const int i = c_scene.clustered ? __ldg(&P.sphere_pos[oi]) : oi;   // where reference sphere oi lives
        sphere_exact<COUNT>(ldg4(P.sphere_geom, i), i, oi, o, d, two_a, four_a, closest, obj, index, best_oi, t_hit, tally);
    }
    tally.add(CTR_PLANE_TESTS);
    plane_exact_num<COUNT>(plane_numerator(o), o, d, closest, obj, t_hit, tally);
}

// exact tests of a lane's surviving spheres of one chunk, in index order (TRT.c:805-828 restricted to the survivors)
struct ClosestHit { double closest, t_hit; int obj, index, best_oi; };
template <bool CLUSTERED>
static __device__ __noinline__ ClosestHit walk_survivors(const double4 *geom, const int *orig, unsigned int survivors, int base, d3 o, d3 d, ClosestHit h)
{
    const Tally<false> no_tally{nullptr};
    const double a = dot(d, d);                  // TRT.c:646
    const double two_a = 2.0 * a, four_a = 4.0 * a;
    while (survivors) {
        const int j = __ffs(survivors) - 1;
        survivors &= survivors - 1;
        const int oi = CLUSTERED ? __ldg(&orig[base + j]) : base + j;
        sphere_exact<false, CLUSTERED>(ldg4(geom, base + j), base + j, oi, o, d, two_a, four_a, h.closest, h.obj, h.index, h.best_oi, h.t_hit, no_tally);
    }
    return h;
}

// ---- certificate-guided query ---------------------------------------------------------------------------
// Float certificates (trt_cert.h) classify every sphere against the ray; only the spheres they cannot decide
// are evaluated exactly, each lane walking its own survivors in index order (tie-breaking preserved).
//   Q_CLOSEST: returns false; (obj, index, t_hit) = the closest hit over all objects, as trace_ray finds it.
//   Q_DIR / Q_POINT: returns true when a certificate proves the light BLOCKED.  Otherwise (obj, t_hit) is the
//     closest hit among the objects that can still matter — every other object is proven either not hit or,
//     for a point light, hit only beyond the light (farther than any hit that could block) — so the caller's
//     decision on it equals the reference's decision on the closest hit over all objects.
// `mode` is warp-uniform at run time: the consume step runs all of a record's queries through ONE copy of this
// code (three inlined copies cost more in instruction-cache misses than the uniform branches cost in issue slots).
struct Query {
    d3 d;                  // unit direction, exact
    trt_cert_ray rf;       // the same ray in float, with its error slack
    float near_limit;      // Q_POINT: a hit closer than this certainly blocks; +inf otherwise
    float far_limit;       // Q_POINT: a sphere entirely beyond this cannot block; +inf otherwise
    double plane_denom;    // Q_DIR: dot(direction, ground normal), a per-light constant
    bool ground_candidate; // Q_POINT: the ground was not ruled out by trt_cert_ground_cannot_block
    int mode;
};

// CONST_RECORDS: small scene (at most TRT_CLUSTER_MIN_SPHERES spheres): records in shared memory (s_pairs, copied at kernel
// start), reference order, no clusters.  Otherwise the scene is k-d-sorted with bounding balls and its records are read
// from global memory.
// raw outcome of pass 1 (the float classification of a small scene's single chunk) when it was taken elsewhere: see classify_two_shadows
struct Classified { unsigned int survivors; bool blocked; };

// the ground's part of a query, after the spheres (TRT.c:831-853 through 907 / 936-941): shared by both query forms
__device__ __forceinline__ void finish_query(const Query &qy, const d3 &o, double num_g, bool blocked, double &closest, int &obj, double &t_hit)
{
    const Tally<false> no_tally{nullptr};
    const bool usable = qy.rf.usable != 0;
    if (qy.mode == Q_CLOSEST) {
        // (num_g is the reference's own numerator for this origin: only the denominator's sign is left to the float certificate)
        if (!(usable && trt_cert_plane_miss_num(&qy.rf, num_g, c_scene.ground_normal_f[0], c_scene.ground_normal_f[1], c_scene.ground_normal_f[2])))
            plane_exact_num<false>(num_g, o, qy.d, closest, obj, t_hit, no_tally);
    } else if (qy.mode == Q_DIR) {
        // any hit blocks.  The ground (TRT.c:677-695): numerator and denominator are the reference's own doubles
        // (the denominator is a per-light constant); opposite signs or a zero numerator give t <= 0, a miss,
        // without the division.
        if (TRT_UNLIKELY(!blocked && obj == 0 && fabs(qy.plane_denom) > 0.00001 && num_g != 0.0 && ((num_g < 0.0) == (qy.plane_denom < 0.0)))) {
            const double t = ieee_div(num_g, qy.plane_denom);
            if (t > 0.00001) obj = 2;
        }
    } else {
        if (TRT_UNLIKELY(!blocked && qy.ground_candidate)) plane_exact_num<false>(num_g, o, qy.d, closest, obj, t_hit, no_tally);
    }
}

// One chunk of a many-sphere query, float classification (pass 1) with the WORK compacted across the warp.  With lane = ray and a loop
// over the chunk's pairs the warp classified the UNION of what its rays could reach: ncu showed 101 pair trips per warp and query where
// one ray needs about 30 — bounce and shadow rays of a warp diverge — and that loop was 56 % of the kernel's instructions.  Here every
// (ray, group of 8 spheres the ray's certificate could not skip) is one work item, items are dealt to QUADS of lanes (lane = one pair of
// the group's spheres, packed FP32 as before: the 4 pair records are one 128-byte line), and only the items that exist are processed: a
// round of the loop classifies eight items, 64 spheres.  A ray's float record is read from shared memory (written once per query), the few spheres it cannot rule out are
// OR-ed into its survivor word there.  The arithmetic per (ray, sphere) is trt_cert_sphere's.
// `groups`: bit g set = this lane's ray needs group g of the chunk.  Returns (the lane's survivor mask, a sphere certainly blocks its light).
static __device__ __noinline__ uint2 classify_chunk(const CullPair *__restrict__ cull_pairs, ClusterScratch *cs, int base, int cnt, unsigned int groups,
                                                   bool shadow, const unsigned int lanes)
{
    const int lane = threadIdx.x & 31;
    const unsigned int below = lanes & ((1u << lane) - 1u);
    const int my_rank = __popc(below), n_act = __popc(lanes);
    // the work list: items ordered by group, then by ray
    int total = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const bool mine = (groups >> g) & 1u;
        const unsigned int b = __ballot_sync(lanes, mine);
        if (mine) cs->item[total + __popc(b & below)] = (unsigned char)((lane << 2) | g);
        total += __popc(b);
    }
    cs->surv[lane] = 0u;
    cs->blk[lane] = 0u;
    __syncwarp(lanes);
    // the lanes of a QUAD share an item, each takes one pair of its 8 spheres (the 4 pair records are one 128-byte line); with
    // fewer than 4 active lanes one "quad" of n_act lanes takes the pairs in turns
    const bool full = n_act >= 4;
    const int width = full ? 4 : n_act, quads = full ? n_act >> 2 : 1;
    const int quad = full ? my_rank >> 2 : 0, first = full ? my_rank & 3 : my_rank;
    const float2 neg1 = make_float2(-1.0f, -1.0f), shrink2 = make_float2(0.99999237060546875f, 0.99999237060546875f);
    if (quad < quads) {
        for (int it = quad; it < total; it += quads) {
            const unsigned int item = cs->item[it];
            const int r = (int)(item >> 2), g = (int)(item & 3u);
            const float4 ro = cs->ray[r][0], rd = cs->ray[r][1];
            const float slack = ro.w, far_limit = rd.w;
            const float2 nox = make_float2(-ro.x, -ro.x), noy = make_float2(-ro.y, -ro.y), noz = make_float2(-ro.z, -ro.z);
            const float2 dx = make_float2(rd.x, rd.x), dy = make_float2(rd.y, rd.y), dz = make_float2(rd.z, rd.z);
            const float2 slack2 = make_float2(slack, slack);
            for (int pp = first; pp < 4; pp += width) {
                const int j = 8 * g + 2 * pp;               // spheres base + j and base + j + 1
                if (j >= cnt) break;                         // (the scene's last chunk may be short; an odd count ends in a pad record)
                const CullPair sp = ldg_pair(cull_pairs, (base + j) >> 1);
                // the arithmetic of trt_cert_sphere2 (same operations, same rounding) on float pairs
                const float2 ocx = __fadd2_rn(sp.cx, nox), ocy = __fadd2_rn(sp.cy, noy), ocz = __fadd2_rn(sp.cz, noz);
                const float2 tc = __ffma2_rn(ocz, dz, __ffma2_rn(ocy, dy, __fmul2_rn(ocx, dx)));
                const float2 ntc = __fmul2_rn(tc, neg1);
                const float2 wx = __ffma2_rn(ntc, dx, ocx), wy = __ffma2_rn(ntc, dy, ocy), wz = __ffma2_rn(ntc, dz, ocz);
                const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
                const float2 outer = __fadd2_rn(sp.r, slack2);
                const float2 outer_sq = __fmul2_rn(outer, outer);
                const float2 front = __ffma2_rn(sp.r, neg1, tc);
                const bool miss0 = (h2.x > outer_sq.x) || (tc.x < -slack) || (front.x > far_limit);
                const bool miss1 = (h2.y > outer_sq.y) || (tc.y < -slack) || (front.y > far_limit);
                const unsigned int keep = (miss0 ? 0u : 1u) | (miss1 ? 0u : 2u);
                if (keep) atomicOr(&cs->surv[r], keep << j);
                if (shadow) {
                    const float near_limit = cs->near_limit[r];
                    const float2 inner = __ffma2_rn(sp.r, shrink2, make_float2(-slack, -slack));
                    const float2 inner_sq = __fmul2_rn(inner, inner);
                    const bool blocks0 = (inner.x > 0.0f) && (h2.x < inner_sq.x) && (front.x > slack) && (tc.x < near_limit);
                    const bool blocks1 = (inner.y > 0.0f) && (h2.y < inner_sq.y) && (front.y > slack) && (tc.y < near_limit);
                    // (a pad sphere, r = 0, has inner < 0 and cannot block)
                    if (blocks0 || blocks1) cs->blk[r] = 1u;
                }
            }
        }
    }
    __syncwarp(lanes);
    return make_uint2(cs->surv[lane], cs->blk[lane]);
}

// Many-sphere scenes (CULL == 2): the spheres are in k-d order, every 32 consecutive ones — one chunk of this loop — under a bounding
// ball and every 8 under a ball inside it (trt_cert_cluster_miss).  The WARP walks the chunks together (ball records at warp-uniform
// addresses, two chunk balls per trip): a chunk no lane's ray can reach is skipped by all, every lane tests the chunk's four group
// balls for its own ray, and the spheres of the groups it still needs are classified by classify_chunk — per ray, dealt across the
// warp.  (A per-lane stackless walk of a bounding-ball tree over the same order, chunks taken nearest first with a shrinking reach,
// and a lane = sphere form were measured and lost; DESIGN.md 4.4 has the numbers.)
__device__ __forceinline__ bool query_clustered(const RenderParams &P, ClusterScratch *cs, const unsigned int lanes, const Query &qy, const d3 &o, double num_g,
                                                int &obj, int &index, double &t_hit, unsigned int *exact_tests)
{
    // `lanes`: the lanes that run this query together — named by the caller, not taken from __activemask(): they share the warp's
    // scratch (classify_chunk), so they must really be here together
    __syncwarp(lanes);
    const d3 d = qy.d;
    double closest = INFINITY;
    obj = 0;
    index = -1;
    t_hit = 0.0;
    const int n = c_scene.num_spheres;
    const bool usable = qy.rf.usable != 0;
    const bool shadow = qy.mode != Q_CLOSEST;        // warp-uniform: a step runs the same query of all its records
    bool blocked = false;
    // the certificate ray in packed form: both halves of every pair carry the same value
    struct { float2 nox, noy, noz, dx, dy, dz; } rp;
    rp.nox = make_float2(-qy.rf.ox, -qy.rf.ox); rp.noy = make_float2(-qy.rf.oy, -qy.rf.oy); rp.noz = make_float2(-qy.rf.oz, -qy.rf.oz);
    rp.dx = make_float2(qy.rf.dx, qy.rf.dx); rp.dy = make_float2(qy.rf.dy, qy.rf.dy); rp.dz = make_float2(qy.rf.dz, qy.rf.dz);
    const float slack = qy.rf.slack_t;
    const float2 slack2 = make_float2(slack, slack);
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    int best_oi = -1;
    // this ray's float record where the lanes that classify for it find it (classify_chunk)
    {
        const int lane = threadIdx.x & 31;
        cs->ray[lane][0] = make_float4(qy.rf.ox, qy.rf.oy, qy.rf.oz, slack);
        cs->ray[lane][1] = make_float4(qy.rf.dx, qy.rf.dy, qy.rf.dz, qy.far_limit);
        cs->near_limit[lane] = qy.near_limit;
    }
    const int nchunks = (n + 31) >> 5;
    const float2 far2 = make_float2(qy.far_limit, qy.far_limit);
    bool all_finished = false;
    for (int c0 = 0; c0 < nchunks && !all_finished; c0 += 2) {
        // two chunk balls per trip (trt_cert_cluster_miss's arithmetic on float pairs; the array ends in a spare record)
        bool miss0, miss1;
        {
            const float4 b0 = __ldg(&P.clusters[c0]), b1 = __ldg(&P.clusters[c0 + 1]);
            const float2 R = make_float2(b0.w, b1.w);
            const float2 ocx = __fadd2_rn(make_float2(b0.x, b1.x), rp.nox), ocy = __fadd2_rn(make_float2(b0.y, b1.y), rp.noy),
                         ocz = __fadd2_rn(make_float2(b0.z, b1.z), rp.noz);
            const float2 tc = __ffma2_rn(ocz, rp.dz, __ffma2_rn(ocy, rp.dy, __fmul2_rn(ocx, rp.dx)));
            const float2 ntc = __fmul2_rn(tc, neg1);
            const float2 wx = __ffma2_rn(ntc, rp.dx, ocx), wy = __ffma2_rn(ntc, rp.dy, ocy), wz = __ffma2_rn(ntc, rp.dz, ocz);
            const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
            const float2 outer = __fadd2_rn(R, slack2);
            const float2 outer_sq = __fmul2_rn(outer, outer);
            const float2 back = __fadd2_rn(tc, R), front = __ffma2_rn(R, neg1, tc);
            miss0 = (h2.x > outer_sq.x) || (back.x < -slack) || (front.x > far2.x);
            miss1 = (h2.y > outer_sq.y) || (back.y < -slack) || (front.y > far2.y);
        }
#pragma unroll 1
        for (int half_c = 0; half_c < 2; half_c++) {
        if (c0 + half_c >= nchunks) break;
        const int base = (c0 + half_c) << 5;
        const int cnt = min(32, n - base);
        // the chunk is a cluster of the k-d order with a bounding ball, and four balls of 8 inside it: skip what no lane's ray can
        // reach, and let every lane drop what its own ray cannot.  A shadow ray whose answer is known — certainly blocked, or
        // (directional light: any hit blocks, TRT.c:907-908) an exact hit found — needs nothing more.
        const bool finished = shadow && usable && (blocked || (qy.mode == Q_DIR && obj != 0));
        const bool ball_missed = finished || (usable && (half_c ? miss1 : miss0));
        if (__all_sync(lanes, ball_missed)) {
            // once every lane's answer is known the remaining chunks cannot change anything
            if (__all_sync(lanes, finished)) {
                all_finished = true;
                break;
            }
            continue;
        }
        unsigned int groups = 0u;      // this lane's: groups of 8 its ray may still hit
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const CullPair g = ldg_pair(P.subballs, 2 * (base >> 5) + half);
            const float2 ocx = __fadd2_rn(g.cx, rp.nox), ocy = __fadd2_rn(g.cy, rp.noy), ocz = __fadd2_rn(g.cz, rp.noz);
            const float2 tc = __ffma2_rn(ocz, rp.dz, __ffma2_rn(ocy, rp.dy, __fmul2_rn(ocx, rp.dx)));
            const float2 ntc = __fmul2_rn(tc, neg1);
            const float2 wx = __ffma2_rn(ntc, rp.dx, ocx), wy = __ffma2_rn(ntc, rp.dy, ocy), wz = __ffma2_rn(ntc, rp.dz, ocz);
            const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
            const float2 outer = __fadd2_rn(g.r, slack2);
            const float2 outer_sq = __fmul2_rn(outer, outer);
            const float2 front = __ffma2_rn(g.r, neg1, tc), back = __fadd2_rn(g.r, tc);
            const bool m0 = usable && ((h2.x > outer_sq.x) || (back.x < -slack) || (front.x > qy.far_limit));
            const bool m1 = usable && ((h2.y > outer_sq.y) || (back.y < -slack) || (front.y > qy.far_limit));
            if (!m0 && !ball_missed) groups |= 1u << (2 * half);
            if (!m1 && !ball_missed) groups |= 2u << (2 * half);
        }
        if (__all_sync(lanes, groups == 0u)) continue;
        // pass 1 (float): the spheres of the groups this ray still needs, classified by whichever lanes are free (classify_chunk);
        // a ray without a usable certificate keeps everything
        const unsigned int valid = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
        const uint2 cls = classify_chunk(P.cull_pairs, cs, base, cnt, usable ? groups : 0u, shadow, lanes);
        unsigned int survivors = cls.x & valid;
        if (cls.y) blocked = true;
        if (!usable) survivors = groups ? valid : 0u;
        if (shadow && usable && blocked) survivors = 0;
        if (exact_tests) *exact_tests += (unsigned int)__popc(survivors);
        // pass 2 (double, exact): each lane walks its own survivors (out of line: the kernel's hot code has to fit the instruction cache)
        if (TRT_UNLIKELY(survivors != 0)) {
            const ClosestHit h = walk_survivors<true>(P.sphere_geom, P.sphere_orig, survivors, base, o, d, ClosestHit{closest, t_hit, obj, index, best_oi});
            closest = h.closest;
            t_hit = h.t_hit;
            obj = h.obj;
            index = h.index;
            best_oi = h.best_oi;
        }
        }
    }
    blocked = blocked && usable && shadow;
    finish_query(qy, o, num_g, blocked, closest, obj, t_hit);
    return blocked;
}

// Small scenes (CONST_RECORDS: at most TRT_CLUSTER_MIN_SPHERES = 32 spheres, one chunk, records in shared memory) classify their
// candidates in one warp-uniform loop; many-sphere scenes go to query_clustered.
template <bool CONST_RECORDS, bool PRE = false>
__device__ __forceinline__ bool query_certified(const RenderParams &P, const float4 *s_pairs, const Query &qy, const d3 &o, double num_g, bool use_patch,
                                                unsigned int patch_mask, int &obj, int &index, double &t_hit, unsigned int *exact_tests,
                                                const Classified pre = Classified{0u, false}, ClusterScratch *cs = nullptr, unsigned int lanes = 0u)
{
    static_assert(!PRE || CONST_RECORDS, "a precomputed classification covers the single chunk of a small scene");
    if (!CONST_RECORDS) return query_clustered(P, cs, lanes, qy, o, num_g, obj, index, t_hit, exact_tests);
    const d3 d = qy.d;
    double closest = INFINITY;
    obj = 0;
    index = -1;
    t_hit = 0.0;
    const bool usable = qy.rf.usable != 0;
    const bool shadow = qy.mode != Q_CLOSEST;
    bool blocked = false;
    // the certificate ray in packed form: both halves of every pair carry the same value
    struct { float2 nox, noy, noz, dx, dy, dz; } rp;
    rp.nox = make_float2(-qy.rf.ox, -qy.rf.ox); rp.noy = make_float2(-qy.rf.oy, -qy.rf.oy); rp.noz = make_float2(-qy.rf.oz, -qy.rf.oz);
    rp.dx = make_float2(qy.rf.dx, qy.rf.dx); rp.dy = make_float2(qy.rf.dy, qy.rf.dy); rp.dz = make_float2(qy.rf.dz, qy.rf.dz);
    const float slack = qy.rf.slack_t;
    const float2 slack2 = make_float2(slack, slack), nslack2 = make_float2(-slack, -slack);
    const float2 neg1 = make_float2(-1.0f, -1.0f), shrink2 = make_float2(0.99999237060546875f, 0.99999237060546875f);
    int best_oi = -1;
    {
        // pass 1 (float, warp-uniform record addresses): classify the candidate spheres — all of them, or, for the
        // first-generation hits of a patch tile, the few the patch certificate left (use_patch)
        // (the mask of existing spheres is a per-scene constant, host-evaluated)
        const unsigned int valid = c_scene.sphere_mask;
        const unsigned int candidates = use_patch ? (patch_mask & valid) : valid;
        unsigned int survivors = PRE ? pre.survivors : 0u;
        if (PRE) blocked = pre.blocked;
        // two spheres per trip: the arithmetic of trt_cert_sphere2 (same operations, same rounding) on float pairs
#pragma unroll 1
        for (unsigned int m = PRE ? 0u : ((candidates | (candidates >> 1)) & 0x55555555u); m; m &= m - 1) {
            const int j = __ffs(m) - 1;             // even: spheres j and j + 1
            const float4 lo = s_pairs[j], hi = s_pairs[j + 1];      // pair j / 2 = float4 j and j + 1 (j is even)
            CullPair g;
            g.cx = make_float2(lo.x, lo.y); g.cy = make_float2(lo.z, lo.w);
            g.cz = make_float2(hi.x, hi.y); g.r = make_float2(hi.z, hi.w);
            const float2 ocx = __fadd2_rn(g.cx, rp.nox), ocy = __fadd2_rn(g.cy, rp.noy), ocz = __fadd2_rn(g.cz, rp.noz);
            const float2 tc = __ffma2_rn(ocz, rp.dz, __ffma2_rn(ocy, rp.dy, __fmul2_rn(ocx, rp.dx)));
            const float2 ntc = __fmul2_rn(tc, neg1);
            const float2 wx = __ffma2_rn(ntc, rp.dx, ocx), wy = __ffma2_rn(ntc, rp.dy, ocy), wz = __ffma2_rn(ntc, rp.dz, ocz);
            const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
            const float2 outer = __fadd2_rn(g.r, slack2);
            const float2 outer_sq = __fmul2_rn(outer, outer);
            const float2 front = __ffma2_rn(g.r, neg1, tc);
            const bool miss0 = (h2.x > outer_sq.x) || (tc.x < -slack) || (front.x > qy.far_limit);
            const bool miss1 = (h2.y > outer_sq.y) || (tc.y < -slack) || (front.y > qy.far_limit);
            if (!miss0) survivors |= 1u << j;
            if (!miss1) survivors |= 2u << j;
            if (shadow) {
                const float2 inner = __ffma2_rn(g.r, shrink2, nslack2);
                const float2 inner_sq = __fmul2_rn(inner, inner);
                const bool blocks0 = (inner.x > 0.0f) && (h2.x < inner_sq.x) && (front.x > slack) && (tc.x < qy.near_limit);
                const bool blocks1 = (inner.y > 0.0f) && (h2.y < inner_sq.y) && (front.y > slack) && (tc.y < qy.near_limit);
                // a pad sphere (r = 0) has inner < 0 and cannot block; a non-candidate of a patch tile that "blocks" is
                // impossible as well: the patch certificate proved that none of the tile's rays can reach it
                blocked = blocked || blocks0 || blocks1;
            }
        }
        survivors &= candidates;
        if (!usable) survivors = candidates;
        if (shadow && usable && blocked) survivors = 0;
        if (exact_tests) *exact_tests += (unsigned int)__popc(survivors);
        // pass 2 (double, exact): each lane walks its own survivors in index order
        if (TRT_UNLIKELY(survivors != 0)) {
            // out of line: three unrolled queries would each carry a copy of the exact test, and the kernel's hot code
            // has to fit the instruction cache (no_instruction was 20 % of the stall samples with the copies inline)
            const ClosestHit h = walk_survivors<false>(P.sphere_geom, P.sphere_orig, survivors, 0, o, d, ClosestHit{closest, t_hit, obj, index, best_oi});
            closest = h.closest;
            t_hit = h.t_hit;
            obj = h.obj;
            index = h.index;
            best_oi = h.best_oi;
        }
    }
    blocked = blocked && usable && shadow;
    finish_query(qy, o, num_g, blocked, closest, obj, t_hit);
    return blocked;
}

// Pass 1 of BOTH shadow queries of a record in one loop (small scenes with 1 + 1 lights): the two rays leave the same point, so a
// pair's records are loaded once and its centre offsets formed once; the two classifications are independent chains that
// interleave, and the loop overhead is paid once.  Per ray exactly the operations of query_certified's own pass 1 (same rounding,
// same comparisons): the results are the ones the separate loops would give.  Candidates: the union of the two queries' candidate
// masks; each query keeps its own afterwards (query_certified<.., PRE>).  (The bounce ray in the same loop as well — it needs the
// reflected direction before the shadow queries — was measured: 23.2 vs 22.8 ms.)
__device__ __forceinline__ void classify_two_shadows(const float4 *s_pairs, const Query &qa, const Query &qb, unsigned int candidates,
                                                     Classified &ca, Classified &cb)
{
    const float2 nox = make_float2(-qa.rf.ox, -qa.rf.ox), noy = make_float2(-qa.rf.oy, -qa.rf.oy), noz = make_float2(-qa.rf.oz, -qa.rf.oz);
    const float2 adx = make_float2(qa.rf.dx, qa.rf.dx), ady = make_float2(qa.rf.dy, qa.rf.dy), adz = make_float2(qa.rf.dz, qa.rf.dz);
    const float2 bdx = make_float2(qb.rf.dx, qb.rf.dx), bdy = make_float2(qb.rf.dy, qb.rf.dy), bdz = make_float2(qb.rf.dz, qb.rf.dz);
    const float sa = qa.rf.slack_t, sb = qb.rf.slack_t;
    const float2 sa2 = make_float2(sa, sa), nsa2 = make_float2(-sa, -sa), sb2 = make_float2(sb, sb), nsb2 = make_float2(-sb, -sb);
    const float2 neg1 = make_float2(-1.0f, -1.0f), shrink2 = make_float2(0.99999237060546875f, 0.99999237060546875f);
    unsigned int surv_a = 0u, surv_b = 0u;
    bool blk_a = false, blk_b = false;
#pragma unroll 1
    for (unsigned int m = (candidates | (candidates >> 1)) & 0x55555555u; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const float4 lo = s_pairs[j], hi = s_pairs[j + 1];
        const float2 gcx = make_float2(lo.x, lo.y), gcy = make_float2(lo.z, lo.w), gcz = make_float2(hi.x, hi.y), gr = make_float2(hi.z, hi.w);
        const float2 ocx = __fadd2_rn(gcx, nox), ocy = __fadd2_rn(gcy, noy), ocz = __fadd2_rn(gcz, noz);
        // ray a (directional light: no length)
        {
            const float2 tc = __ffma2_rn(ocz, adz, __ffma2_rn(ocy, ady, __fmul2_rn(ocx, adx)));
            const float2 ntc = __fmul2_rn(tc, neg1);
            const float2 wx = __ffma2_rn(ntc, adx, ocx), wy = __ffma2_rn(ntc, ady, ocy), wz = __ffma2_rn(ntc, adz, ocz);
            const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
            const float2 outer = __fadd2_rn(gr, sa2);
            const float2 outer_sq = __fmul2_rn(outer, outer);
            const float2 front = __ffma2_rn(gr, neg1, tc);
            const bool miss0 = (h2.x > outer_sq.x) || (tc.x < -sa) || (front.x > qa.far_limit);
            const bool miss1 = (h2.y > outer_sq.y) || (tc.y < -sa) || (front.y > qa.far_limit);
            if (!miss0) surv_a |= 1u << j;
            if (!miss1) surv_a |= 2u << j;
            const float2 inner = __ffma2_rn(gr, shrink2, nsa2);
            const float2 inner_sq = __fmul2_rn(inner, inner);
            const bool blocks0 = (inner.x > 0.0f) && (h2.x < inner_sq.x) && (front.x > sa) && (tc.x < qa.near_limit);
            const bool blocks1 = (inner.y > 0.0f) && (h2.y < inner_sq.y) && (front.y > sa) && (tc.y < qa.near_limit);
            blk_a = blk_a || blocks0 || blocks1;
        }
        // ray b (point light: near and far limits)
        {
            const float2 tc = __ffma2_rn(ocz, bdz, __ffma2_rn(ocy, bdy, __fmul2_rn(ocx, bdx)));
            const float2 ntc = __fmul2_rn(tc, neg1);
            const float2 wx = __ffma2_rn(ntc, bdx, ocx), wy = __ffma2_rn(ntc, bdy, ocy), wz = __ffma2_rn(ntc, bdz, ocz);
            const float2 h2 = __ffma2_rn(wz, wz, __ffma2_rn(wy, wy, __fmul2_rn(wx, wx)));
            const float2 outer = __fadd2_rn(gr, sb2);
            const float2 outer_sq = __fmul2_rn(outer, outer);
            const float2 front = __ffma2_rn(gr, neg1, tc);
            const bool miss0 = (h2.x > outer_sq.x) || (tc.x < -sb) || (front.x > qb.far_limit);
            const bool miss1 = (h2.y > outer_sq.y) || (tc.y < -sb) || (front.y > qb.far_limit);
            if (!miss0) surv_b |= 1u << j;
            if (!miss1) surv_b |= 2u << j;
            const float2 inner = __ffma2_rn(gr, shrink2, nsb2);
            const float2 inner_sq = __fmul2_rn(inner, inner);
            const bool blocks0 = (inner.x > 0.0f) && (h2.x < inner_sq.x) && (front.x > sb) && (tc.x < qb.near_limit);
            const bool blocks1 = (inner.y > 0.0f) && (h2.y < inner_sq.y) && (front.y > sb) && (tc.y < qb.near_limit);
            blk_b = blk_b || blocks0 || blocks1;
        }
    }
    ca.survivors = surv_a; ca.blocked = blk_a;
    cb.survivors = surv_b; cb.blocked = blk_b;
}

// hit point pushed back toward the ray origin by EPSILON, TRT.c:871-874
__device__ __forceinline__ d3 push_back(const d3 &o, const d3 &hit)
{
    d3 back = unit(o - hit);
    back = back * TRT_EPSILON;
    return hit + back;
}

// ---- fused encode (RenderParams::ansi): a finished tile's cells as terminal bytes, stored where they belong -----------
// The destination may be another GPU's memory (NVLink peer mapping) or page-locked host memory: every store instruction
// should carry whole words to consecutive addresses.  Each warp keeps one staging row per tile row in shared memory,
// pre-filled once with the constant part of the cells (pixel_str, TRT.c:1103) and the newline; a finished pixel only drops its
// nine digits in (byte_to_digits, TRT.c:1134-1139), and the warp copies the rows out, 128 consecutive bytes per store
// instruction, the byte shift between the word-aligned staging row and the destination done with a funnel shift.
__constant__ unsigned char c_cell_template[TRT_CELL_BYTES + 3] = {0x1b, '[', '4', '8', ';', '2', ';', '0', '0', '0', ';', '0', '0', '0', ';',
                                                                  '0', '0', '0', 'm', ' ', ' ', 0x1b, '[', '0', 'm', 0, 0, 0};

__device__ __forceinline__ void stage_fill(unsigned char *stage, int lane)
{
#pragma unroll 1
    for (int i = lane; i < TILE_H * ANSI_STAGE_STRIDE; i += 32) {
        const int pos = i % ANSI_STAGE_STRIDE;
        stage[i] = pos < TILE_W * TRT_CELL_BYTES ? c_cell_template[pos % TRT_CELL_BYTES] : (unsigned char)(pos == TILE_W * TRT_CELL_BYTES ? '\n' : 0);
    }
}

__device__ __forceinline__ void stage_digits(unsigned char *cell, unsigned int r, unsigned int g, unsigned int b)
{
    const unsigned int v[3] = {r, g, b};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const unsigned int h = v[c] / 100u, rest = v[c] - 100u * h, t = rest / 10u;
        cell[7 + 4 * c] = (unsigned char)('0' + h);
        cell[8 + 4 * c] = (unsigned char)('0' + t);
        cell[9 + 4 * c] = (unsigned char)('0' + rest - 10u * t);
    }
}

// `rows` staging rows of n bytes each to dst0, dst0 + row_bytes, ... (any alignment).  One compact loop (the code runs once
// per tile and must stay small): 64 word slots per row, two trips of the warp per row.
__device__ __forceinline__ void copy_tile_rows(unsigned char *dst0, size_t row_bytes, const unsigned char *stage, int n, int rows, int lane)
{
#pragma unroll 1
    for (int idx = lane; idx < rows * 64; idx += 32) {
        const int r = idx >> 6, j = idx & 63;
        unsigned char *const dst = dst0 + (size_t)r * row_bytes;
        const unsigned int s = (unsigned int)(reinterpret_cast<unsigned long long>(dst) & 3ull);
        // destination word j (of the word-aligned address below dst) = staging bytes [4j - s, 4j - s + 4):
        // the upper s bytes of staging word j-1 and the lower 4-s of word j
        const int first = 4 * j - (int)s;
        if (first < n && j < ANSI_STAGE_STRIDE / 4) {
            const unsigned int *const src = reinterpret_cast<const unsigned int *>(stage + r * ANSI_STAGE_STRIDE);
            const unsigned int v = __funnelshift_rc(j > 0 ? src[j - 1] : 0u, src[j], 8u * (4u - s));
            unsigned char *const out = dst + first;
            if (first >= 0 && first + 4 <= n) {
                *reinterpret_cast<unsigned int *>(out) = v;
            } else {
                // the at most three bytes before the first whole word and after the last
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (first + b >= 0 && first + b < n) out[b] = (unsigned char)(v >> (8 * b));
            }
        }
    }
}

// Q_CLOSEST query set-up for a ray with a double unit direction; S0 = |o|_1 + max centre |.|_1 (inf: unusable)
__device__ __forceinline__ void setup_closest_query(Query &qy, const d3 &o, const d3 &d, float S0)
{
    qy.d = d;
    trt_cert_set_origin(&qy.rf, o.x, o.y, o.z);
    trt_cert_set_unit_dir(&qy.rf, d.x, d.y, d.z, S0);
    qy.near_limit = INFINITY;
    qy.far_limit = INFINITY;
    qy.plane_denom = 0.0;
    qy.ground_candidate = false;
    qy.mode = Q_CLOSEST;
}

// ---- K1 ------------------------------------------------------------------------------------------------
// Build flavours:  CULL == 0 or COUNT: every query is the all-FP64 reference-order query (COUNT tallies the
// reference's work counters); otherwise the certificate-guided one.  COUNT with CULL != 0 runs BOTH and reports any
// disagreement of the final answers in CTR_CULL_VIOLATIONS (the on-device audit of the certificates, of the
// survivor logic and of the host-precomputed primary-ray terms).
template <bool COUNT, int CULL, int LIGHTS>
#ifdef TRT_MAXNREG
__global__ void __maxnreg__(TRT_MAXNREG) k_render(const RenderParams P)
#else
__global__ void __launch_bounds__(CTA_THREADS, TRT_MIN_CTAS_PER_SM) k_render(const RenderParams P)
#endif
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_byte_to_unit = reinterpret_cast<double *>(smem_raw);   // k/255.0 (TRT.c:866), evaluated on the host in double
    for (int k = threadIdx.x; k < 256; k += CTA_THREADS) s_byte_to_unit[k] = P.byte_to_unit[k];
    float4 *const s_pairs = reinterpret_cast<float4 *>(smem_raw + SMEM_PAIRS_OFFSET);
    if (CULL == 1 && threadIdx.x < 2 * SMEM_PAIRS && threadIdx.x < 2 * (c_scene.num_spheres / 2 + 1))
        s_pairs[threadIdx.x] = __ldg(reinterpret_cast<const float4 *>(P.cull_pairs) + threadIdx.x);
    __syncthreads();

    const Tally<COUNT> tally{P.counters};
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpShared &W = reinterpret_cast<WarpShared *>(smem_raw + SMEM_TABLE_BYTES)[warp];
    ClusterScratch *const cs = CULL == 2 ? reinterpret_cast<ClusterScratch *>(smem_raw + SMEM_BYTES) + warp : nullptr;
    if (P.ansi) stage_fill(reinterpret_cast<unsigned char *>(W.ansi_stage), lane);
    // finished samples of the current tile, [channel][k * 32 + pixel lane]: written once, read once at the end of
    // the tile by the pixel's lane -> parked in an L2-resident per-warp slice of global memory, not in shared memory
    double *const res = P.sample_scratch + (size_t)(blockIdx.x * WARPS_PER_CTA + warp) * (size_t)(3 * TILE_SAMPLES);
    const int band_rows = P.row1 - P.row0;
    const int tiles_x = (P.width + TILE_W - 1) / TILE_W;
    const int tiles_y = (band_rows + TILE_H - 1) / TILE_H;
    const unsigned int num_tiles = (unsigned int)(tiles_x * tiles_y);

    // LIGHTS == 1: exactly one directional and one point light — the reference's own scene shape (TRT.c:1278-1287) —
    // known at compile time: the query loop below unrolls into three specialised queries with no mode branches
    const int num_dir = LIGHTS == 1 ? 1 : c_scene.num_dir, num_point = LIGHTS == 1 ? 1 : c_scene.num_point;
    const int num_spheres = c_scene.num_spheres;
    const d3 eye = mk3(c_scene.eye[0], c_scene.eye[1], c_scene.eye[2]);
    // tile certificates need the masks to fit; bigger scenes test every sphere exactly for primary rays
    const bool tile_certs = CULL == 1 || (CULL == 2 && num_spheres <= 32 * TMASK_WORDS);   // compile-time true for small scenes
    const int mask_words = CULL == 1 ? 1 : (num_spheres + 31) >> 5;   // small scenes: one word, loops over it fold away
    // certificates on (CULL != 0: the host only picks these flavours when the scene's magnitudes are inside the range the error
    // bounds hold for, DevScene::filter_enabled) or off (inf: every certificate ray unusable) is a compile-time property here
    const float S_max = CULL != 0 ? c_scene.filter_centre_l1 : INFINITY;
    // small scenes with exactly 1 + 1 lights (the reference's own shape): tile and patch certificates come from the prepass
    constexpr bool PREPASS = TRT_TILE_PREPASS && CULL == 1 && LIGHTS == 1;

    for (;;) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= num_tiles) break;
#if TRT_TILES_BOTTOM_UP
        // bottom rows first: ground and reflections are the expensive tiles, sky rows the cheap ones — the persistent kernel's
        // tail (warps waiting for the last tiles) is then made of cheap tiles
        tile = num_tiles - 1u - tile;
#endif
        const int ty = (int)(tile / (unsigned)tiles_x), tx = (int)(tile % (unsigned)tiles_x);
        const int col = tx * TILE_W + (lane & (TILE_W - 1));
        const int brow = ty * TILE_H + (lane >> 3); // band-local row
        const int row = P.row0 + brow;
        const bool valid = col < P.width && brow < band_rows;
        if (valid) tally.add(CTR_PIXELS);

        // per-pixel part of the primary ray, TRT.c:987-988 (two IEEE divisions per tile and lane; nothing of this is
        // kept in registers across tiles: registers are what limits the resident warps)
        const double sx0 = ((ieee_div((double)col, (double)P.width)) * c_scene.screen_width - c_scene.screen_width / 2.0);
        const double sy0 = -((ieee_div((double)row, (double)P.height)) * c_scene.screen_height - c_scene.screen_height / 2.0);

        // ---- tile certificates (float): which spheres can any primary ray of this tile hit at all? the ground?
        bool tile_ground_miss = false;
        bool patch = false;       // patch certificates available for this tile's first-generation hits
        if (PREPASS) {
            // the tile's certificates were taken by k_tile_certs (same functions, same inputs, one thread per tile): the
            // per-tile code of this kernel is a 32-byte load instead of 8 KB of instructions that would evict the hot loop
            TRT_BOUND(tile < num_tiles, 4);
            const uint4 m = __ldg(P.tile_info + 2 * (size_t)tile), f = __ldg(P.tile_info + 2 * (size_t)tile + 1);
            if (lane == 0) {
                W.tmask[0] = m.x;
                W.pmask[0] = m.y;
                W.pmask[1] = m.z;
                W.pmask[2] = m.w;
            }
            tile_ground_miss = (f.x & 1u) != 0u;
            patch = (f.x & 2u) != 0u;
        } else if (tile_certs) {
            // (read in place: the fields become constant-bank operands; a local copy would be twenty loads per tile)
            const trt_cert_camera &cam = c_scene.cam_f;
            float Dx, Dy, Dz, hx, hy;
            trt_cert_tile_cone(&cam, P.pixel_w_f, P.pixel_h_f, tx * TILE_W, P.row0 + ty * TILE_H, TILE_W, TILE_H, &Dx, &Dy, &Dz, &hx, &hy);
            const float h = fmaf(hx, cam.nbx, hy * cam.nby);
            const float S = c_scene.eye_l1 + c_scene.filter_centre_l1;
            unsigned int any_sphere = 0;
            for (int wd = 0; wd < mask_words; wd++) {
                const int i = wd * 32 + lane;
                bool keep = false;
                if (i < num_spheres) {
                    const float4 g = __ldg(&P.sphere_cull[i]);
                    keep = !(CULL != 0 && trt_cert_tile_sphere_miss(cam.ex, cam.ey, cam.ez, Dx, Dy, Dz, h, g.x, g.y, g.z, g.w, S));
                }
                const unsigned int m = __ballot_sync(0xffffffffu, keep);
                any_sphere |= m;
                if (lane == 0) W.tmask[wd] = m;
            }
            const float *gn = c_scene.ground_normal_f;
            const float dn = fmaf(Dz, gn[2], fmaf(Dy, gn[1], Dx * gn[0]));
            const float bxn = fmaf(cam.bx[2], gn[2], fmaf(cam.bx[1], gn[1], cam.bx[0] * gn[0]));
            const float byn = fmaf(cam.by[2], gn[2], fmaf(cam.by[1], gn[1], cam.by[0] * gn[0]));
            const float scale = (fabsf(Dx) + fabsf(Dy) + fabsf(Dz) + 2.0f * h) * c_scene.ground_normal_l1;
            const int sgn = CULL != 0 ? trt_cert_tile_plane_sign(dn, bxn, byn, hx, hy, scale) : 0;
            // numerator < 0 with every denominator > 0 (or the mirror image): t < 0 for every primary ray of the tile
            tile_ground_miss = (c_scene.prim_num_sign < 0 && sgn > 0) || (c_scene.prim_num_sign > 0 && sgn < 0);

            // ---- patch certificates: no sphere in reach of the primary rays => every hit of this tile's primary rays
            // is a ground hit inside a ball (trt_cert_patch_ball); decide once which spheres its shadow and bounce
            // rays can reach (lane = sphere), instead of classifying every sphere for every ray
            if (any_sphere == 0 && !tile_ground_miss && (CULL == 1 || num_spheres <= PATCH_MAX_SPHERES) && c_scene.prim_num_sign != 0) {
                trt_cert_ball ball;
                trt_cert_patch_ball(&cam, Dx, Dy, Dz, hx, hy, c_scene.prim_num_f, gn[0], gn[1], gn[2], S, &ball);
#ifdef TRT_NO_PATCH
                patch = false;
#else
                patch = ball.ok != 0;
#endif
                if (patch) {
                    const float S_ball = fabsf(ball.cx) + fabsf(ball.cy) + fabsf(ball.cz) + ball.r + c_scene.filter_centre_l1;
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane < num_spheres) g = __ldg(&P.sphere_cull[lane]);
                    for (int l = 0; l < num_dir; l++) {
                        const float *Lf = c_scene.dir[l].Lf;
                        const bool cand = lane < num_spheres && trt_cert_patch_dir_candidate(&ball, Lf[0], Lf[1], Lf[2], g.x, g.y, g.z, g.w, S_ball);
                        const unsigned int m = __ballot_sync(0xffffffffu, cand);
                        if (lane == 0) W.pmask[l] = m;
                    }
                    for (int l = 0; l < num_point; l++) {
                        const DevLightPoint &Lp = c_scene.point[l];
                        const bool cand = lane < num_spheres &&
                                          trt_cert_patch_point_candidate(&ball, Lp.pos_f[0], Lp.pos_f[1], Lp.pos_f[2], g.x, g.y, g.z, g.w, S_ball + Lp.pos_l1);
                        const unsigned int m = __ballot_sync(0xffffffffu, cand);
                        if (lane == 0) W.pmask[num_dir + l] = m;
                    }
                    {
                        // reflection of the tile's central direction about the plane (unit normal in float)
                        const float *un = c_scene.ground_unit_normal_f;
                        const float dnn = 2.0f * fmaf(Dz, un[2], fmaf(Dy, un[1], Dx * un[0]));
                        const float Rx = fmaf(-dnn, un[0], Dx), Ry = fmaf(-dnn, un[1], Dy), Rz = fmaf(-dnn, un[2], Dz);
                        const float hb = h * 1.0001f + (32.0f * TRT_CERT_U) * (fabsf(Dx) + fabsf(Dy) + fabsf(Dz));
                        const bool cand = lane < num_spheres && trt_cert_patch_bounce_candidate(&ball, Rx, Ry, Rz, hb, g.x, g.y, g.z, g.w, S_ball);
                        const unsigned int m = __ballot_sync(0xffffffffu, cand);
                        if (lane == 0) W.pmask[num_dir + num_point] = m;
                    }
                }
            }
        }
        __syncwarp();

        int round = 0;            // next sample index to produce
        int a_head = 0, a_count = 0, qhead = 0, qcount = 0;   // ring A (first-generation hits), ring B (bounce hits)

        // Full warps first: a ring is consumed as soon as it holds 32 records (so neither can exceed 63), primary
        // rays are produced while both are short, and the remainders are drained at the end of the tile.
        while (round < TRT_RAYS_PER_PIXEL || a_count > 0 || qcount > 0) {
            const bool consume_b = qcount >= 32 || (a_count == 0 && round == TRT_RAYS_PER_PIXEL);
            const bool consume_a = !consume_b && (a_count >= 32 || round == TRT_RAYS_PER_PIXEL);
            if (!consume_a && !consume_b) {
                // =========================== PRODUCE: primary rays of sample `round` ==============================
                const int k = round++;
                bool hit_surface = false;
                d3 d = mk3(0.0, 0.0, 0.0);
                double t_hit = 0.0;
                int obj = 0, index = -1;
                if (valid) {
                    tally.add(CTR_SAMPLES);
                    tally.add(CTR_BOUNCE_ITERS);
                    tally.add(CTR_TRACE_CALLS);
                    const double sx = sx0 + c_scene.sub_dx[k] * P.pixel_w;    // TRT.c:992 (pixel_w = screen_width / width, host-evaluated)
                    const double sy = sy0 + c_scene.sub_dy[k] * P.pixel_h;    // TRT.c:993
                    const double sz = -c_scene.screen_distance;
                    d3 sp = mk3(0.0, 0.0, 0.0);
                    sp = sp + mk3(c_scene.bx[0] * sx, c_scene.bx[1] * sx, c_scene.bx[2] * sx);
                    sp = sp + mk3(c_scene.by[0] * sy, c_scene.by[1] * sy, c_scene.by[2] * sy);
                    sp = sp + mk3(c_scene.bz[0] * sz, c_scene.bz[1] * sz, c_scene.bz[2] * sz);
                    d = unit(sp - eye);         // TRT.c:1005 (origin subtracted from an untranslated vector), 1008
                    if (COUNT || !tile_certs) query_reference<COUNT>(P, eye, d, obj, index, t_hit, tally);
                    if (tile_certs) {
                        // exact tests of the tile's candidate spheres; oc and c of TRT.c:640-648 are per-frame
                        // constants for rays leaving the eye (host-evaluated with the reference's operations)
                        int obj2 = 0, index2 = -1;
                        double t2 = 0.0;
                        unsigned int exact = 0;
                        const Tally<false> no_tally{nullptr};
                        double closest = INFINITY;
                        int best_oi = -1;
                        const double a = dot(d, d);
                        const double two_a = 2.0 * a, four_a = 4.0 * a;
                        for (int wd = 0; wd < mask_words; wd++) {
                            const unsigned int m = W.tmask[wd];
                            exact += (unsigned int)__popc(m);
                            if (m) {
                                const ClosestHit hh = walk_survivors<CULL == 2>(P.sphere_geom, P.sphere_orig, m, wd * 32, eye, d, ClosestHit{closest, t2, obj2, index2, best_oi});
                                closest = hh.closest; t2 = hh.t_hit; obj2 = hh.obj; index2 = hh.index; best_oi = hh.best_oi;
                                // (the shared out-of-line walk, not a second inlined copy of the exact test with its sqrt and
                                // division: 1.5 KB less hot code)
                            }
                        }
                        if (!tile_ground_miss) plane_exact_num<false>(c_scene.prim_num, eye, d, closest, obj2, t2, no_tally);
                        if (COUNT) {
                            tally.add(CTR_EXACT_SPHERE_TESTS, exact);
                            if (obj2 != obj || (obj && (__double_as_longlong(t2) != __double_as_longlong(t_hit) || (obj == 1 && index2 != index))))
                                tally.add(CTR_CULL_VIOLATIONS);
                        } else {
                            obj = obj2;
                            index = index2;
                            t_hit = t2;
                        }
                    }
                    if (obj == 0) {
                        // the sample sees the sky: weight 1, weight_sum 1, so the sample IS the texel colour
                        // (c*1.0, 0.0 + c, c*(1.0/1.0) are exact), TRT.c:1034-1035, 1051, 1061
                        tally.add(CTR_SKY_LOOKUPS);
                        if (COUNT) atomicAdd(&P.counters[CTR_BOUNCE_HIST0], 1ull);
                        const d3 c = sky_colour(P, s_byte_to_unit, d);
                        if (COUNT && CULL != 0 && sky_certificate_disagrees(d)) tally.add(CTR_CULL_VIOLATIONS);
                        TRT_BOUND(k >= 0 && k < TRT_RAYS_PER_PIXEL && (size_t)(blockIdx.x * WARPS_PER_CTA + warp + 1) * (3 * TILE_SAMPLES) * sizeof(double) <= P.scratch_bytes, 2);
                        __stcg(&res[0 * TILE_SAMPLES + k * 32 + lane], c.x);
                        __stcg(&res[1 * TILE_SAMPLES + k * 32 + lane], c.y);
                        __stcg(&res[2 * TILE_SAMPLES + k * 32 + lane], c.z);
                    } else {
                        tally.add(CTR_TRACE_HITS);
                        hit_surface = true;
                    }
                }
                const unsigned int pushers = __ballot_sync(0xffffffffu, hit_surface);
                if (hit_surface) {
                    const int slot = (a_head + a_count + __popc(pushers & ((1u << lane) - 1u))) & (QCAP - 1);
                    TRT_BOUND(a_count + __popc(pushers) <= QCAP && slot >= 0 && slot < QCAP, 0);
                    W.a_dx[slot] = d.x; W.a_dy[slot] = d.y; W.a_dz[slot] = d.z;
                    W.a_t[slot] = t_hit;
                    W.a_meta[slot] = (unsigned)lane | ((unsigned)k << 5) | ((unsigned)obj << 13);
                    W.a_index[slot] = index;
                }
                a_count += __popc(pushers);
                __syncwarp();
            } else {
                // =========================== CONSUME: up to 32 queued surface hits ================================
                const int n_rec = min(32, consume_a ? a_count : qcount);
                const bool active = lane < n_rec;
                // Only what the first steps need is read now; the accumulated colour and weights stay in the ring slot
                // until the shading step (registers are what limits the number of resident warps).  The slot cannot
                // be overwritten before that: pushes happen after the warp-wide ballot at the end of this step.
                const int slot = ((consume_a ? a_head : qhead) + lane) & (QCAP - 1);
                d3 o = eye, d = mk3(0, 0, 0);
                double t_hit = 0.0;
                unsigned int meta = 0;
                int index = 0;
                if (consume_a) {
                    // first-generation hit: the ray left the eye, nothing accumulated yet (TRT.c:1014-1017)
                    if (active) {
                        d = mk3(W.a_dx[slot], W.a_dy[slot], W.a_dz[slot]);
                        t_hit = W.a_t[slot];
                        meta = W.a_meta[slot];
                        index = W.a_index[slot];
                    }
                    a_head = (a_head + n_rec) & (QCAP - 1);
                    a_count -= n_rec;
                } else {
                    if (active) {
                        o = mk3(W.ox[slot], W.oy[slot], W.oz[slot]);
                        d = mk3(W.dx[slot], W.dy[slot], W.dz[slot]);
                        t_hit = W.t[slot];
                        meta = W.meta[slot];
                        index = W.index[slot];
                    }
                    qhead = (qhead + n_rec) & (QCAP - 1);
                    qcount -= n_rec;
                }
                const bool use_patch = consume_a && patch;   // warp-uniform
                d3 sample = mk3(0.0, 0.0, 0.0);
                double weight = 1.0, weight_sum = 0.0;

                bool push = false;
                if (active) {
                    const int pix = (int)(meta & 31u), k = (int)((meta >> 5) & 15u);
                    int bounces = (int)((meta >> 9) & 15u);
                    const int obj = (int)((meta >> 13) & 3u);
                    if (COUNT && P.row_cost) atomicAdd(&P.row_cost[ty * TILE_H + (pix >> 3)], 25u);   // (counting flavour: the cost pre-pass)
                    tally.add(CTR_LIGHTING_CALLS);

                    // the surface point (TRT.c:663-665 / 690-692), pushed back by EPSILON (871-874), and its normal (878)
                    const d3 hit = mk3(o.x + t_hit * d.x, o.y + t_hit * d.y, o.z + t_hit * d.z);
                    const d3 at = push_back(o, hit);
                    d3 nrm;
                    const DevMaterial *mat;
                    if (obj == 1) {
                        const double4 g = ldg4(P.sphere_geom, index);
                        nrm = unit(mk3(hit.x - g.x, hit.y - g.y, hit.z - g.z));   // TRT.c:824, 878
                        mat = &P.sphere_mat[index];
                    } else {
                        nrm = mk3(c_scene.ground_unit_normal[0], c_scene.ground_unit_normal[1], c_scene.ground_unit_normal[2]);
                        const int odd = x86_int(floor(hit.x) + floor(hit.z)) & 1;  // checker parity, TRT.c:850
                        mat = odd ? &c_scene.ground_odd : &c_scene.ground_even;
                    }

                    // Every query of this record leaves the surface point `at`: one shadow query per light
                    // (apply_lighting, TRT.c:894-963), then the bounce ray (TRT.c:1054-1057, next trip of 1018).
                    Query qy;
                    const float S0 = trt_cert_set_origin(&qy.rf, at.x, at.y, at.z) + S_max;
                    const double num_g = plane_numerator(at);
                    d3 lit = mk3(0.0, 0.0, 0.0);
                    bool done = false;
                    const int nq = num_dir + num_point + 1;
                    // small scenes with 1 + 1 lights: the float classification of BOTH shadow rays runs in one loop at q == 0
                    // (classify_two_shadows); the point-light query, set up there, and its raw classification wait for q == 1
                    constexpr bool FUSE = TRT_FUSED_SHADOWS && CULL == 1 && LIGHTS == 1;
                    Query qy_point;
                    double light_d2_point = 0.0;
                    Classified pre_dir{0u, false}, pre_point{0u, false};
                    auto setup_point_query = [&](Query &qp, const DevLightPoint &Lp, double &light_d2) {
                        d3 ld = mk3(Lp.pos[0] - at.x, Lp.pos[1] - at.y, Lp.pos[2] - at.z);            // TRT.c:929
                        light_d2 = dot(ld, ld);
                        qp.mode = Q_POINT;
                        qp.plane_denom = 0.0;
                        qp.d = unit(ld);                                                              // TRT.c:933
                        const float dist = trt_cert_set_dir_toward(&qp.rf, Lp.pos_f[0], Lp.pos_f[1], Lp.pos_f[2], S0 + Lp.pos_l1);
                        const float guard = fmaf(2.0f, qp.rf.slack_t, 1e-5f);
                        qp.near_limit = dist - guard;
                        qp.far_limit = dist + guard;
                        qp.ground_candidate = !trt_cert_ground_cannot_block(num_g, Lp.height, c_scene.ground_margin);
                    };
                    auto one_query = [&](const int q) {
                        // ---- set-up: warp-uniform branch on the kind of query ------------------------------------
                        bool run = true;
                        double light_d2 = 0.0;
                        qy.near_limit = INFINITY;
                        qy.far_limit = INFINITY;
                        qy.ground_candidate = false;
                        qy.plane_denom = 0.0;
                        if (FUSE && q == 1) {
                            qy = qy_point;
                            light_d2 = light_d2_point;
                        } else if (q < num_dir) {
                            const DevLightDir &Ld = c_scene.dir[q];
                            qy.mode = Q_DIR;
                            qy.d = mk3(Ld.L[0], Ld.L[1], Ld.L[2]);                  // unit(-direction), TRT.c:903-904
                            qy.rf.dx = Ld.Lf[0]; qy.rf.dy = Ld.Lf[1]; qy.rf.dz = Ld.Lf[2];
                            qy.rf.slack_t = (32.0f * TRT_CERT_U) * S0;
                            qy.rf.usable = (S0 < 1e15f) && Ld.lf_unit;     // |Lf| = 1 within 1e-5: checked once on the host
                            qy.plane_denom = Ld.plane_denom;
                            if (FUSE) {
                                qy_point.rf = qy.rf;                    // same origin
                                setup_point_query(qy_point, c_scene.point[0], light_d2_point);
                                const unsigned int cand = use_patch ? ((W.pmask[0] | W.pmask[1]) & c_scene.sphere_mask) : c_scene.sphere_mask;
                                classify_two_shadows(s_pairs, qy, qy_point, cand, pre_dir, pre_point);
                            }
                        } else if (q < num_dir + num_point) {
                            setup_point_query(qy, c_scene.point[q - num_dir], light_d2);
                        } else {
                            // all lights done: finish apply_lighting (TRT.c:960) and accumulate (TRT.c:1034-1051) ...
                            qy.mode = Q_CLOSEST;
                            if (!consume_a) {
                                sample = mk3(W.sr[slot], W.sg[slot], W.sb[slot]);
                                weight = W.w[slot];
                                weight_sum = W.ws[slot];
                            }
                            d3 colour = mk3(clampd(lit.x, 0.0, 1.0), clampd(lit.y, 0.0, 1.0), clampd(lit.z, 0.0, 1.0));
                            weight_sum += weight;
                            colour = colour * weight;
                            sample = sample + colour;
                            weight *= mat->reflectivity;                                     // TRT.c:1041-1042
                            bounces++;
                            run = bounces < TRT_BOUNCE_LIMIT && weight > 0.00001;            // loop condition, TRT.c:1018
                            if (run) {
                                // ... reflect (TRT.c:627-633, 1054-1055): the bounce ray leaves the surface point
                                if (consume_a) d = mk3(W.a_dx[slot], W.a_dy[slot], W.a_dz[slot]);
                                else d = mk3(W.dx[slot], W.dy[slot], W.dz[slot]);
                                const double dn = dot(d, nrm);
                                qy.d = unit(mk3(d.x - 2.0 * dn * nrm.x, d.y - 2.0 * dn * nrm.y, d.z - 2.0 * dn * nrm.z));
                                trt_cert_set_unit_dir(&qy.rf, qy.d.x, qy.d.y, qy.d.z, S0);
                                tally.add(CTR_BOUNCE_ITERS);
                            } else {
                                done = true;
                            }
                        }
                        // ---- the query: one copy of the code for all kinds ----------------------------------------
                        int obj2 = 0, index2 = -1;
                        double t2 = 0.0;
                        bool blocked = false;
                        // many-sphere scenes: the lanes of this step that run the query, named explicitly (they share per-warp scratch)
                        unsigned int qlanes = 0u;
                        if (CULL == 2) qlanes = __ballot_sync(n_rec >= 32 ? 0xffffffffu : ((1u << n_rec) - 1u), run);
                        if (run) {
                            if (COUNT || CULL == 0) {
                                tally.add(CTR_TRACE_CALLS);
                                query_reference<COUNT>(P, at, qy.d, obj2, index2, t2, tally);
                            }
                            if (CULL != 0) {
                                int obj3, index3;
                                double t3;
                                unsigned int exact = 0;
                                bool blocked3;
                                if (FUSE && q < 2)
                                    blocked3 = query_certified<true, true>(P, s_pairs, qy, at, num_g, use_patch, use_patch ? W.pmask[q] : 0u, obj3, index3, t3,
                                                                           COUNT ? &exact : nullptr, q == 0 ? pre_dir : pre_point);
                                else
                                    blocked3 = query_certified<CULL == 1>(P, s_pairs, qy, at, num_g, use_patch, use_patch ? W.pmask[q] : 0u, obj3, index3, t3,
                                                                          COUNT ? &exact : nullptr, Classified{0u, false}, cs, qlanes);
                                if (COUNT) {
                                    // the audit: both paths must lead to the same decision / the same hit
                                    tally.add(CTR_EXACT_SPHERE_TESTS, exact);
                                    bool same;
                                    if (qy.mode == Q_CLOSEST) {
                                        same = obj3 == obj2 && (!obj2 || (__double_as_longlong(t3) == __double_as_longlong(t2) && (obj2 != 1 || index3 == index2)));
                                    } else {
                                        bool open_ref = obj2 == 0, open_cert = !blocked3 && obj3 == 0;
                                        if (qy.mode == Q_POINT) {
                                            if (!open_ref) {
                                                const d3 h2 = mk3(at.x + t2 * qy.d.x, at.y + t2 * qy.d.y, at.z + t2 * qy.d.z);
                                                const d3 tb = push_back(at, h2) - at;
                                                open_ref = light_d2 < dot(tb, tb);
                                            }
                                            if (!blocked3 && obj3 != 0) {
                                                const d3 h3 = mk3(at.x + t3 * qy.d.x, at.y + t3 * qy.d.y, at.z + t3 * qy.d.z);
                                                const d3 tb = push_back(at, h3) - at;
                                                open_cert = light_d2 < dot(tb, tb);
                                            }
                                        }
                                        same = open_ref == open_cert;
                                    }
                                    if (!same) tally.add(CTR_CULL_VIOLATIONS);
                                } else {
                                    obj2 = obj3;
                                    index2 = index3;
                                    t2 = t3;
                                    blocked = blocked3;
                                }
                            }
                            if (obj2) tally.add(CTR_TRACE_HITS);
                            else {
                                tally.add(CTR_SKY_LOOKUPS);
                                if (qy.mode != Q_CLOSEST) tally.add(CTR_SKY_SKIPPED);
                            }
                        }
                        // ---- what the answer means ----------------------------------------------------------------
                        if (qy.mode == Q_CLOSEST) {
                            if (run) {
                                if (obj2 == 0) {
                                    // the bounce ray leaves the scene: sky colour, TRT.c:858-867; the sample ends
                                    d3 c = sky_colour(P, s_byte_to_unit, qy.d);
                                    if (COUNT && CULL != 0 && sky_certificate_disagrees(qy.d)) tally.add(CTR_CULL_VIOLATIONS);
                                    weight_sum += weight;
                                    c = c * weight;
                                    sample = sample + c;
                                    done = true;
                                } else {
                                    push = true;
                                    o = at;
                                    d = qy.d;
                                    t_hit = t2;
                                    index = index2;
                                    meta = (unsigned)pix | ((unsigned)k << 5) | ((unsigned)bounces << 9) | ((unsigned)obj2 << 13);
                                }
                            }
                        } else {
                            bool open = !blocked && obj2 == 0;
                            double f;
                            const double *lc;
                            if (qy.mode == Q_DIR) {
                                lc = c_scene.dir[q].color;
                                f = 1.0;
                            } else {
                                const DevLightPoint &Lp = c_scene.point[q - num_dir];
                                if (TRT_UNLIKELY(!blocked && obj2 != 0)) {
                                    // a blocker that is farther than the light does not block, TRT.c:936-941
                                    const d3 bh = mk3(at.x + t2 * qy.d.x, at.y + t2 * qy.d.y, at.z + t2 * qy.d.z);
                                    const d3 to_blocker = push_back(at, bh) - at;
                                    open = light_d2 < dot(to_blocker, to_blocker);
                                }
                                lc = Lp.color;
                                f = open ? clampd(ieee_div(Lp.intensity, light_d2), 0.0, 1.0) : 0.0;       // TRT.c:931
                            }
                            if (open) {
                                const double lambert = fmin(dot(nrm, qy.d), 1.0);                          // TRT.c:910, 943
                                f = qy.mode == Q_DIR ? lambert : f * lambert;
                                d3 diffuse = mk3(lc[0] * f, lc[1] * f, lc[2] * f);
                                diffuse = hadamard(diffuse, mk3(mat->color[0], mat->color[1], mat->color[2]));
                                lit = lit + diffuse;
                            }
                        }
                    };
                    // (1 + 1 lights: three specialised copies.  One rolled copy for both shadow queries was measured: no smaller —
                    // it carries both set-ups and both answers — and 26.1 vs 23.7 ms.)
#pragma unroll(LIGHTS == 1 ? 3 : 1)
                    for (int q = 0; q < nq; q++) one_query(q);
                    if (done) {
                        if (COUNT) atomicAdd(&P.counters[CTR_BOUNCE_HIST0 + bounces], 1ull);
                        sample = sample * ieee_div(1.0, weight_sum);                 // TRT.c:1061
                        TRT_BOUND(k >= 0 && k < TRT_RAYS_PER_PIXEL && pix >= 0 && pix < 32, 2);
                    __stcg(&res[0 * TILE_SAMPLES + k * 32 + pix], sample.x);
                        __stcg(&res[1 * TILE_SAMPLES + k * 32 + pix], sample.y);
                        __stcg(&res[2 * TILE_SAMPLES + k * 32 + pix], sample.z);
                    }
                }
                const unsigned int pushers = __ballot_sync(0xffffffffu, push);
                if (push) {
                    const int ps = (qhead + qcount + __popc(pushers & ((1u << lane) - 1u))) & (QCAP - 1);
                    TRT_BOUND(qcount + __popc(pushers) <= QCAP && ps >= 0 && ps < QCAP, 1);
                    W.ox[ps] = o.x; W.oy[ps] = o.y; W.oz[ps] = o.z;
                    W.dx[ps] = d.x; W.dy[ps] = d.y; W.dz[ps] = d.z;
                    W.t[ps] = t_hit;
                    W.sr[ps] = sample.x; W.sg[ps] = sample.y; W.sb[ps] = sample.z;
                    W.w[ps] = weight; W.ws[ps] = weight_sum;
                    W.meta[ps] = meta;
                    W.index[ps] = index;
                }
                qcount += __popc(pushers);
                __syncwarp();
            }
        }

        // ---- per pixel: add the samples in order, average, store (TRT.c:1063-1066) ---------------------------
        __threadfence_block();    // the samples were written by other lanes of this warp
        __syncwarp();
        d3 average = mk3(0.0, 0.0, 0.0);
        if (valid) {
#pragma unroll SUM_UNROLL
            for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++)
                average = average + mk3(__ldcg(&res[0 * TILE_SAMPLES + k * 32 + lane]), __ldcg(&res[1 * TILE_SAMPLES + k * 32 + lane]),
                                        __ldcg(&res[2 * TILE_SAMPLES + k * 32 + lane]));
            average = average * (1.0 / TRT_RAYS_PER_PIXEL);
            if (COUNT && P.row_cost) atomicAdd(&P.row_cost[brow], 50u);
            const size_t pix = (size_t)brow * (size_t)P.width + (size_t)col;
            TRT_BOUND(brow >= 0 && brow < band_rows && col >= 0 && col < P.width, 3);
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
        if (P.ansi) {
            unsigned char *const stage = reinterpret_cast<unsigned char *>(W.ansi_stage);
            static_assert(TILE_W * TRT_CELL_BYTES + 1 <= ANSI_STAGE_STRIDE - 4, "a tile row, its newline and the word read past it");
            const int cols = min(TILE_W, P.width - tx * TILE_W);
            const bool row_end = tx == tiles_x - 1, narrow = row_end && cols < TILE_W;
            if (valid) {
                // the quantisation of buffered_draw_screen, TRT.c:1157-1163: truncation toward zero
                unsigned char *const cell = stage + (lane >> 3) * ANSI_STAGE_STRIDE + (lane & (TILE_W - 1)) * TRT_CELL_BYTES;
                stage_digits(cell, (unsigned char)x86_int(average.x * 255), (unsigned char)x86_int(average.y * 255), (unsigned char)x86_int(average.z * 255));
                if (narrow && (lane & (TILE_W - 1)) == cols - 1) cell[TRT_CELL_BYTES] = '\n';   // the row ends inside the tile (TRT.c:1125)
            }
            __syncwarp();
            const size_t row_bytes = (size_t)TRT_CELL_BYTES * (size_t)P.width + 1;
            const int n = cols * TRT_CELL_BYTES + (row_end ? 1 : 0);
            TRT_BOUND(n > 0 && n <= TILE_W * TRT_CELL_BYTES + 1 && P.row0 + ty * TILE_H + min(TILE_H, band_rows - ty * TILE_H) <= P.row1 && tx * TILE_W + cols <= P.width, 6);
            copy_tile_rows(P.ansi + TRT_HOME_BYTES + (size_t)(P.row0 + ty * TILE_H) * row_bytes + (size_t)(tx * TILE_W) * TRT_CELL_BYTES, row_bytes,
                           stage, n, min(TILE_H, band_rows - ty * TILE_H), lane);
            if (narrow) {
                __syncwarp();
                if (lane < TILE_H) stage[lane * ANSI_STAGE_STRIDE + cols * TRT_CELL_BYTES] = 0x1b;   // back to the template
            }
        }
        __syncwarp();   // the sample slice and tmask are rewritten by the next tile
    }
}

// ---- tile certificates as a prepass (small scenes, 1 + 1 lights) --------------------------------------------------
// K1 is bound by instruction supply: the code it executes regularly is larger than the SM's instruction cache, and the 8 KB of
// per-tile certificate code evicted the hot loop once per tile and warp (ncu: gcc__cache_requests_type_instruction at 83 % of
// its peak rate).  This kernel takes the same certificates — same trt_cert.h functions on the same inputs — with one THREAD
// per tile instead of one warp per tile (lane = sphere), ahead of k_render, and leaves 32 bytes per tile:
//   word 0: spheres the tile's primary rays can reach   words 1-3: patch masks (directional light, point light, bounce)
//   word 4: bit 0 ground certainly missed by the tile's primary rays, bit 1 patch certificates valid
__global__ void __launch_bounds__(128) k_tile_certs(const RenderParams P, uint4 *__restrict__ out)
{
    const int band_rows = P.row1 - P.row0;
    const int tiles_x = (P.width + TILE_W - 1) / TILE_W;
    const int tiles_y = (band_rows + TILE_H - 1) / TILE_H;
    const unsigned int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= (unsigned int)(tiles_x * tiles_y)) return;
    const int ty = (int)(tile / (unsigned)tiles_x), tx = (int)(tile % (unsigned)tiles_x);
    TRT_BOUND((size_t)(tile + 1) * 2 * sizeof(uint4) <= P.tile_info_bytes && c_scene.num_spheres <= 32, 8);
    const int num_spheres = c_scene.num_spheres;          // <= 32 here
    const trt_cert_camera &cam = c_scene.cam_f;
    float Dx, Dy, Dz, hx, hy;
    trt_cert_tile_cone(&cam, P.pixel_w_f, P.pixel_h_f, tx * TILE_W, P.row0 + ty * TILE_H, TILE_W, TILE_H, &Dx, &Dy, &Dz, &hx, &hy);
    const float h = fmaf(hx, cam.nbx, hy * cam.nby);
    const float S = c_scene.eye_l1 + c_scene.filter_centre_l1;
    unsigned int tmask = 0u;
    for (int i = 0; i < num_spheres; i++) {
        const float4 g = __ldg(&P.sphere_cull[i]);
        if (!trt_cert_tile_sphere_miss(cam.ex, cam.ey, cam.ez, Dx, Dy, Dz, h, g.x, g.y, g.z, g.w, S)) tmask |= 1u << i;
    }
    const float *gn = c_scene.ground_normal_f;
    const float dn = fmaf(Dz, gn[2], fmaf(Dy, gn[1], Dx * gn[0]));
    const float bxn = fmaf(cam.bx[2], gn[2], fmaf(cam.bx[1], gn[1], cam.bx[0] * gn[0]));
    const float byn = fmaf(cam.by[2], gn[2], fmaf(cam.by[1], gn[1], cam.by[0] * gn[0]));
    const float scale = (fabsf(Dx) + fabsf(Dy) + fabsf(Dz) + 2.0f * h) * c_scene.ground_normal_l1;
    const int sgn = trt_cert_tile_plane_sign(dn, bxn, byn, hx, hy, scale);
    const bool tile_ground_miss = (c_scene.prim_num_sign < 0 && sgn > 0) || (c_scene.prim_num_sign > 0 && sgn < 0);
    bool patch = false;
    unsigned int pm_dir = 0u, pm_point = 0u, pm_bounce = 0u;
    if (tmask == 0u && !tile_ground_miss && c_scene.prim_num_sign != 0) {
        trt_cert_ball ball;
        trt_cert_patch_ball(&cam, Dx, Dy, Dz, hx, hy, c_scene.prim_num_f, gn[0], gn[1], gn[2], S, &ball);
#ifndef TRT_NO_PATCH
        patch = ball.ok != 0;
#endif
        if (patch) {
            const float S_ball = fabsf(ball.cx) + fabsf(ball.cy) + fabsf(ball.cz) + ball.r + c_scene.filter_centre_l1;
            const float *Lf = c_scene.dir[0].Lf;
            const DevLightPoint &Lp = c_scene.point[0];
            // reflection of the tile's central direction about the plane (unit normal in float)
            const float *un = c_scene.ground_unit_normal_f;
            const float dnn = 2.0f * fmaf(Dz, un[2], fmaf(Dy, un[1], Dx * un[0]));
            const float Rx = fmaf(-dnn, un[0], Dx), Ry = fmaf(-dnn, un[1], Dy), Rz = fmaf(-dnn, un[2], Dz);
            const float hb = h * 1.0001f + (32.0f * TRT_CERT_U) * (fabsf(Dx) + fabsf(Dy) + fabsf(Dz));
            for (int i = 0; i < num_spheres; i++) {
                const float4 g = __ldg(&P.sphere_cull[i]);
                if (trt_cert_patch_dir_candidate(&ball, Lf[0], Lf[1], Lf[2], g.x, g.y, g.z, g.w, S_ball)) pm_dir |= 1u << i;
                if (trt_cert_patch_point_candidate(&ball, Lp.pos_f[0], Lp.pos_f[1], Lp.pos_f[2], g.x, g.y, g.z, g.w, S_ball + Lp.pos_l1)) pm_point |= 1u << i;
                if (trt_cert_patch_bounce_candidate(&ball, Rx, Ry, Rz, hb, g.x, g.y, g.z, g.w, S_ball)) pm_bounce |= 1u << i;
            }
        }
    }
    out[2 * (size_t)tile] = make_uint4(tmask, pm_dir, pm_point, pm_bounce);
    out[2 * (size_t)tile + 1] = make_uint4((tile_ground_miss ? 1u : 0u) | (patch ? 2u : 0u), 0u, 0u, 0u);
}

// ---- unit-level probe: trace_ray (TRT.c:793-889) for an array of rays, all out-params ---------------
// Used by the parity tests to compare single queries (hit kind, pushed-back point, unit normal,
// material incl. the skybox colour on a miss) with the reference, not only whole frames.  Goes through the
// certificate-guided query when the scene allows it.
// out: 11 doubles per ray = kind, point[3], normal[3], colour[3], reflectivity
__global__ void k_probe_trace(const RenderParams P, const double *__restrict__ rays, int n, double *__restrict__ out)
{
    __shared__ ClusterScratch s_cs[4];           // (128 threads per CTA)
    __shared__ double s_byte_to_unit[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_byte_to_unit[k] = P.byte_to_unit[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int probe_lanes = __ballot_sync(0xffffffffu, i < n);   // the lanes that run a query together (query_clustered)
    if (i >= n) return;
    const d3 o = mk3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]);
    const d3 d = mk3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
    const Tally<false> tally{nullptr};
    int obj, index;
    double t_hit;
    if (c_scene.filter_enabled) {
        Query qy;
        setup_closest_query(qy, o, d, fabsf((float)o.x) + fabsf((float)o.y) + fabsf((float)o.z) + c_scene.filter_centre_l1);
        if (!c_scene.clustered) query_certified<true>(P, reinterpret_cast<const float4 *>(P.cull_pairs), qy, o, plane_numerator(o), false, 0u, obj, index, t_hit, nullptr);
        else query_certified<false>(P, nullptr, qy, o, plane_numerator(o), false, 0u, obj, index, t_hit, nullptr, Classified{0u, false}, &s_cs[threadIdx.x >> 5], probe_lanes);
    } else {
        query_reference<false>(P, o, d, obj, index, t_hit, tally);
    }
    d3 point, normal, colour;
    double reflectivity = 0.0;
    if (obj == 0) {
        point = o;
        normal = unit(d);
        colour = sky_colour(P, s_byte_to_unit, d);
    } else {
        const d3 hit = mk3(o.x + t_hit * d.x, o.y + t_hit * d.y, o.z + t_hit * d.z);
        const DevMaterial *m;
        if (obj == 1) {
            const double4 g = ldg4(P.sphere_geom, index);
            normal = unit(mk3(hit.x - g.x, hit.y - g.y, hit.z - g.z));
            m = &P.sphere_mat[index];
        } else {
            normal = mk3(c_scene.ground_unit_normal[0], c_scene.ground_unit_normal[1], c_scene.ground_unit_normal[2]);
            m = (x86_int(floor(hit.x) + floor(hit.z)) & 1) ? &c_scene.ground_odd : &c_scene.ground_even;
        }
        colour = mk3(m->color[0], m->color[1], m->color[2]);
        reflectivity = m->reflectivity;
        point = push_back(o, hit);
    }
    double *r = out + (size_t)i * 11;
    r[0] = (double)obj;
    r[1] = point.x; r[2] = point.y; r[3] = point.z;
    r[4] = normal.x; r[5] = normal.y; r[6] = normal.z;
    r[7] = colour.x; r[8] = colour.y; r[9] = colour.z;
    r[10] = reflectivity;
}

// ray_intersects_sphere (TRT.c:638-672) for an array of (ray, sphere) pairs: out = hit flag, intersection point
__global__ void k_probe_sphere(const double *__restrict__ rays, const double *__restrict__ geom, int n, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 o = mk3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]);
    const d3 d = mk3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
    const double radius = geom[i * 4 + 3];
    const double4 g = make_double4(geom[i * 4 + 0], geom[i * 4 + 1], geom[i * 4 + 2], radius * radius);   // TRT.c:648: one rounded product
    const Tally<false> no_tally{nullptr};
    const double a = dot(d, d);                                                                           // TRT.c:646
    double closest = INFINITY, t_hit = 0.0;
    int obj = 0, index = -1, best_oi = -1;
    sphere_exact<false>(g, 0, 0, o, d, 2.0 * a, 4.0 * a, closest, obj, index, best_oi, t_hit, no_tally);
    double *r = out + (size_t)i * 4;
    r[0] = (double)obj;
    r[1] = obj ? o.x + t_hit * d.x : 0.0;                                                                 // TRT.c:663-665
    r[2] = obj ? o.y + t_hit * d.y : 0.0;
    r[3] = obj ? o.z + t_hit * d.z : 0.0;
}

// ray_intersects_plane (TRT.c:677-695) against the scene's ground: out = hit flag, intersection point
__global__ void k_probe_plane(const double *__restrict__ rays, int n, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 o = mk3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]);
    const d3 d = mk3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
    const Tally<false> no_tally{nullptr};
    double closest = INFINITY, t_hit = 0.0;
    int obj = 0;
    plane_exact_num<false>(plane_numerator(o), o, d, closest, obj, t_hit, no_tally);
    double *r = out + (size_t)i * 4;
    r[0] = obj ? 1.0 : 0.0;
    r[1] = obj ? o.x + t_hit * d.x : 0.0;                                                                 // TRT.c:690-692
    r[2] = obj ? o.y + t_hit * d.y : 0.0;
    r[3] = obj ? o.z + t_hit * d.z : 0.0;
}

// apply_lighting (TRT.c:894-963) for an array of surface points: in = point, normal, material colour (9 doubles), out = the
// lit, clamped colour.  Mirrors the shading of k_render's consume step statement by statement on the same building blocks
// (certificate-guided shadow queries when the scene allows them, push_back, unit, the shared-reciprocal division).
__global__ void k_probe_lighting(const RenderParams P, const double *__restrict__ in, int n, double *__restrict__ out)
{
    __shared__ ClusterScratch s_cs[4];           // (128 threads per CTA)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int probe_lanes = __ballot_sync(0xffffffffu, i < n);   // the lanes that run a query together (query_clustered)
    if (i >= n) return;
    const d3 at = mk3(in[i * 9 + 0], in[i * 9 + 1], in[i * 9 + 2]);
    const d3 nrm = mk3(in[i * 9 + 3], in[i * 9 + 4], in[i * 9 + 5]);
    const d3 colour = mk3(in[i * 9 + 6], in[i * 9 + 7], in[i * 9 + 8]);
    const Tally<false> no_tally{nullptr};
    const bool cert = c_scene.filter_enabled != 0;
    const float S_max = cert ? c_scene.filter_centre_l1 : INFINITY;
    Query qy;
    const float S0 = trt_cert_set_origin(&qy.rf, at.x, at.y, at.z) + S_max;
    const double num_g = plane_numerator(at);
    d3 lit = mk3(0.0, 0.0, 0.0);
    for (int q = 0; q < c_scene.num_dir + c_scene.num_point; q++) {
        double light_d2 = 0.0;
        qy.near_limit = INFINITY;
        qy.far_limit = INFINITY;
        qy.ground_candidate = false;
        qy.plane_denom = 0.0;
        if (q < c_scene.num_dir) {
            const DevLightDir &Ld = c_scene.dir[q];
            qy.mode = Q_DIR;
            qy.d = mk3(Ld.L[0], Ld.L[1], Ld.L[2]);
            qy.rf.dx = Ld.Lf[0]; qy.rf.dy = Ld.Lf[1]; qy.rf.dz = Ld.Lf[2];
            qy.rf.slack_t = (32.0f * TRT_CERT_U) * S0;
            qy.rf.usable = (S0 < 1e15f) && Ld.lf_unit;
            qy.plane_denom = Ld.plane_denom;
        } else {
            const DevLightPoint &Lp = c_scene.point[q - c_scene.num_dir];
            qy.mode = Q_POINT;
            const d3 ld = mk3(Lp.pos[0] - at.x, Lp.pos[1] - at.y, Lp.pos[2] - at.z);
            light_d2 = dot(ld, ld);
            qy.d = unit(ld);
            const float dist = trt_cert_set_dir_toward(&qy.rf, Lp.pos_f[0], Lp.pos_f[1], Lp.pos_f[2], S0 + Lp.pos_l1);
            const float guard = fmaf(2.0f, qy.rf.slack_t, 1e-5f);
            qy.near_limit = dist - guard;
            qy.far_limit = dist + guard;
            qy.ground_candidate = !trt_cert_ground_cannot_block(num_g, Lp.height, c_scene.ground_margin);
        }
        int obj2 = 0, index2 = -1;
        double t2 = 0.0;
        bool blocked = false;
        if (!cert) query_reference<false>(P, at, qy.d, obj2, index2, t2, no_tally);
        else if (!c_scene.clustered) blocked = query_certified<true>(P, reinterpret_cast<const float4 *>(P.cull_pairs), qy, at, num_g, false, 0u, obj2, index2, t2, nullptr);
        else blocked = query_certified<false>(P, nullptr, qy, at, num_g, false, 0u, obj2, index2, t2, nullptr, Classified{0u, false}, &s_cs[threadIdx.x >> 5], probe_lanes);
        bool open = !blocked && obj2 == 0;
        double f = 1.0;
        const double *lc;
        if (qy.mode == Q_DIR) {
            lc = c_scene.dir[q].color;
        } else {
            const DevLightPoint &Lp = c_scene.point[q - c_scene.num_dir];
            if (!blocked && obj2 != 0) {
                const d3 bh = mk3(at.x + t2 * qy.d.x, at.y + t2 * qy.d.y, at.z + t2 * qy.d.z);
                const d3 to_blocker = push_back(at, bh) - at;
                open = light_d2 < dot(to_blocker, to_blocker);                                      // TRT.c:936-941
            }
            lc = Lp.color;
            f = open ? clampd(ieee_div(Lp.intensity, light_d2), 0.0, 1.0) : 0.0;                    // TRT.c:931
        }
        if (open) {
            const double lambert = fmin(dot(nrm, qy.d), 1.0);                                       // TRT.c:910, 943
            f = qy.mode == Q_DIR ? lambert : f * lambert;
            d3 diffuse = mk3(lc[0] * f, lc[1] * f, lc[2] * f);
            diffuse = hadamard(diffuse, colour);
            lit = lit + diffuse;
        }
    }
    out[i * 3 + 0] = clampd(lit.x, 0.0, 1.0);                                                       // TRT.c:960
    out[i * 3 + 1] = clampd(lit.y, 0.0, 1.0);
    out[i * 3 + 2] = clampd(lit.z, 0.0, 1.0);
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
            // the guarded square root alone, wherever it claims its fast path
            bool ok = true;
            const double ss = v.x * v.x + v.y * v.y + v.z * v.z;
            const double root = sqrt_guarded(ss, ok);
            if (ok && __double_as_longlong(root) != __double_as_longlong(__dsqrt_rn(ss))) bad++;
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

// self-checking build: the violation counters of this translation unit (16 words), read and cleared; returns 1 when compiled in
int render_bounds_read(unsigned int *out16)
{
#ifdef TRT_BOUNDS_CHECK
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out16, g_trt_bounds, sizeof(unsigned int) * 16));
    const unsigned int zero[16] = {0};
    CK(cudaMemcpyToSymbol(g_trt_bounds, zero, sizeof zero));
    return 1;
#else
    for (int i = 0; i < 16; i++) out16[i] = 0u;
    return 0;
#endif
}

void upload_scene_constants(const DevScene &scene, cudaStream_t stream)
{
    CK(cudaMemcpyToSymbolAsync(c_scene, &scene, sizeof(DevScene), 0, cudaMemcpyHostToDevice, stream));
}

template <bool COUNT, int CULL>
static void prepare_kernel()
{
    // dynamic shared memory per CTA (the rings of 4 warps) may exceed the 48 KB default
    CK(cudaFuncSetAttribute(k_render<COUNT, CULL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of<CULL>()));
    CK(cudaFuncSetAttribute(k_render<COUNT, CULL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_of<CULL>()));
}

int render_ctas_per_sm()
{
    static int cached = 0;
    if (!cached) {
        prepare_kernel<false, 0>(); prepare_kernel<false, 1>(); prepare_kernel<false, 2>();
        prepare_kernel<true, 0>(); prepare_kernel<true, 1>(); prepare_kernel<true, 2>();
        int n = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_render<false, 1, 0>, CTA_THREADS, SMEM_BYTES));
        cached = n > 0 ? n : 1;
    }
    return cached;
}

size_t render_tile_info_bytes(int width, int rows)
{
    const size_t tiles = (size_t)((width + TILE_W - 1) / TILE_W) * (size_t)((rows + TILE_H - 1) / TILE_H);
    return (tiles ? tiles : 1) * 2 * sizeof(uint4);
}

size_t render_scratch_bytes(int num_sms)
{
    // one slice of finished samples (3 channels x 320 samples) per warp of the persistent grid
    return (size_t)num_sms * (size_t)render_ctas_per_sm() * WARPS_PER_CTA * (size_t)(3 * TILE_SAMPLES) * sizeof(double);
}

void launch_render(const RenderParams &p, bool count, int cull, bool one_plus_one, int num_sms, cudaStream_t stream)
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
    if (TRT_TILE_PREPASS && cull == 1 && one_plus_one) {
        if (!p.tile_info) {
            fprintf(stderr, "libtrt_b200: launch_render without a tile_info buffer\n");
            exit(1);
        }
        k_tile_certs<<<(unsigned)((tiles + 127) / 128), 128, 0, stream>>>(p, const_cast<uint4 *>(p.tile_info));
    }
    // 12 flavours: counting or not, certificates off / small scene / clustered scene, generic lights or exactly 1 + 1
#define TRT_LAUNCH(COUNT, CULL) \
    do { \
        if (one_plus_one) k_render<COUNT, CULL, 1><<<g, b, smem_bytes_of<CULL>(), stream>>>(p); \
        else k_render<COUNT, CULL, 0><<<g, b, smem_bytes_of<CULL>(), stream>>>(p); \
    } while (0)
    if (count) {
        if (cull == 1) TRT_LAUNCH(true, 1);
        else if (cull == 2) TRT_LAUNCH(true, 2);
        else TRT_LAUNCH(true, 0);
    } else {
        if (cull == 1) TRT_LAUNCH(false, 1);
        else if (cull == 2) TRT_LAUNCH(false, 2);
        else TRT_LAUNCH(false, 0);
    }
#undef TRT_LAUNCH
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

void launch_probe_sphere(const double *d_rays, const double *d_geom, int n, double *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_sphere<<<(n + 127) / 128, 128, 0, stream>>>(d_rays, d_geom, n, d_out);
    CK(cudaGetLastError());
}

void launch_probe_plane(const double *d_rays, int n, double *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_plane<<<(n + 127) / 128, 128, 0, stream>>>(d_rays, n, d_out);
    CK(cudaGetLastError());
}

void launch_probe_lighting(const RenderParams &p, const double *d_in, int n, double *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_lighting<<<(n + 127) / 128, 128, 0, stream>>>(p, d_in, n, d_out);
    CK(cudaGetLastError());
}

void launch_probe_sky(const RenderParams &p, const double *d_dirs, int n, int *d_out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_probe_sky<<<(n + 127) / 128, 128, 0, stream>>>(p, d_dirs, n, d_out);
    CK(cudaGetLastError());
}

} // namespace trt
