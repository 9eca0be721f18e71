/*
 * trt_cert.h — single-precision CERTIFICATES for outcomes of the reference's FP64 intersection tests.
 *
 * The render kernel (trt_render.cu) must reproduce the reference's IEEE-double results bit for bit
 * (TRT.c = /root/reference/TerminalRayTracer.c).  Most ray/sphere and ray/plane tests, however, only
 * feed a DECISION (hit / no hit, blocked / open), and most of those decisions are geometrically
 * clear-cut.  The functions below evaluate a test in float together with a bound on everything float
 * got wrong, and answer only when the reference's own double evaluation provably reaches the same
 * decision; every other case is reported as "unknown" and is evaluated exactly in FP64 by the caller.
 * They never produce a value that reaches a pixel.
 *
 * The header is plain C99 / C++ with no CUDA dependency except the TRT_HD qualifier, so that the very
 * same code is (a) compiled into the kernel and (b) run on the CPU against the oracle on millions of
 * rays (tests/test_certificates.py via oracle/cert_check.c; 0 contradictions required), in addition to
 * the on-device audit of the counting build (CTR_CULL_VIOLATIONS).
 *
 * Notation: u = 2^-24 (float unit round-off).  S = an upper bound on the L1 norm of every point
 * that enters a difference (ray origin, sphere centres, light position).  A float dot product of
 * differences of such points with a unit vector, and the distance of such a point from the ray's line,
 * are within 16uS of the real value; the slack used is 2x that.  The reference's own
 * double rounding is ~2^-29 of the slacks, which is why a decision certified here is also the
 * decision of the reference's COMPUTED discriminant / quotient and not only of the real-valued one.
 */
#ifndef TRT_CERT_H
#define TRT_CERT_H

#include <math.h>

#ifdef __CUDACC__
#define TRT_HD __host__ __device__ __forceinline__
#else
#define TRT_HD static inline
#endif

#define TRT_CERT_U 5.9604644775390625e-08f /* 2^-24 */

/* a ray as the certificates see it */
typedef struct {
    float ox, oy, oz;  /* origin, rounded to nearest                                             */
    float dx, dy, dz;  /* direction, |d| = 1 within `dir_err`                                    */
    float slack_t;     /* 2x the bound on the float error of a coordinate along or across the ray */
    int usable;        /* 0: magnitudes outside the range the bounds hold for -> everything unknown */
} trt_cert_ray;

/* origin of a certificate ray from the double origin; returns |o|_1 (float) */
TRT_HD float trt_cert_set_origin(trt_cert_ray *r, double ox, double oy, double oz)
{
    r->ox = (float)ox;
    r->oy = (float)oy;
    r->oz = (float)oz;
    return fabsf(r->ox) + fabsf(r->oy) + fabsf(r->oz);
}

/* direction rounded from a double UNIT vector (primary, bounce and directional-light rays).
 * S >= |o|_1 + max_i |c_i|_1.  A direction that is not unit to float accuracy (normalize_vector's
 * len <= 1e-4 escape, TRT.c:441) makes the ray unusable: every answer becomes "unknown". */
TRT_HD void trt_cert_set_unit_dir(trt_cert_ray *r, double dx, double dy, double dz, float S)
{
    r->dx = (float)dx;
    r->dy = (float)dy;
    r->dz = (float)dz;
    const float dd = fmaf(r->dz, r->dz, fmaf(r->dy, r->dy, r->dx * r->dx));
    r->slack_t = (32.0f * TRT_CERT_U) * S;
    r->usable = (S < 1e15f) && (dd > 0.99999f) && (dd < 1.00001f);
}

/* On the device the square roots and quotients of the certificates use the approximate units (a few ulp off,
 * 2 instructions instead of ~25 with IEEE fix-up paths); every use pads its result by >= 16 ulp.  sqrt(0) becomes
 * NaN there, which fails every comparison, i.e. "no certificate" — the safe answer. */
#if defined(__CUDA_ARCH__)
#define TRT_CERT_RSQRT(x) rsqrtf(x)
#define TRT_CERT_SQRT(x) ((x) * rsqrtf(x))
#define TRT_CERT_DIV(a, b) __fdividef((a), (b))
#else
#define TRT_CERT_RSQRT(x) (1.0f / sqrtf(x))
#define TRT_CERT_SQRT(x) sqrtf(x)
#define TRT_CERT_DIV(a, b) ((a) / (b))
#endif

/* direction from the ray's origin toward the point (lx,ly,lz), formed and normalised in float
 * (point-light shadow rays, TRT.c:929-934).  S >= |o|_1 + max_i |c_i|_1 + |l|_1.  Returns the float
 * distance to the point.  The float direction is off by up to ~2u * S / distance, which enters the slacks. */
TRT_HD float trt_cert_set_dir_toward(trt_cert_ray *r, float lx, float ly, float lz, float S)
{
    const float vx = lx - r->ox, vy = ly - r->oy, vz = lz - r->oz;
    const float dd = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
    const float inv = TRT_CERT_RSQRT(dd);
    r->dx = vx * inv;
    r->dy = vy * inv;
    r->dz = vz * inv;
    const float k = fmaf(2.0f * S, inv, 1.0f);
    r->slack_t = (32.0f * TRT_CERT_U) * S * k;
    r->usable = (S < 1e15f) && (k < 1e4f) && (dd > 1e-30f) && (dd < 1e30f);
    return dd * inv;
}

/* outcome bits of one ray/sphere classification */
#define TRT_CERT_MISS 1   /* the reference's ray_intersects_sphere (TRT.c:638-672) returns false, or its hit
                             lies beyond `far_limit` and therefore cannot decide a point-light shadow test  */
#define TRT_CERT_BLOCKS 2 /* the reference returns true with a hit closer than `near_limit`                 */

/*
 * Classify sphere (cx,cy,cz, r_pad) against ray r.  r_pad >= r * (1 + 2^-20), rounded up (host).
 *   tc = (c - o) . d             position of the centre along the ray
 *   w  = (c - o) - tc d          offset of the centre from the ray's line;  h = |w|
 * (h is formed from w and not as |c-o|^2 - tc^2: the latter cancels catastrophically for far origins, w does
 * not — its float error is ~u |c - o| per component, i.e. |h_f - h| <= 16uS <= slack_t / 2.)
 * MISS   if  h > r_pad + slack_t          the line passes the sphere: discriminant < 0 (TRT.c:651)
 *        or  tc < -slack_t                centre behind the origin: b = -2 tc > 0, so the near root
 *                                         t0 = (-b - sqrt(disc)) / 2a is negative whatever disc is (TRT.c:657-659)
 *        or  tc - r_pad > far_limit       every point of the sphere is farther than far_limit along the ray
 * BLOCKS if  h < r_pad (1 - 2^-17) - slack_t   the line passes well inside the sphere (disc > 0)
 *        and tc - r_pad > slack_t               the whole sphere is in front: both roots positive, t0 > 0
 *        and tc < near_limit                    t0 <= tc: the hit is closer than near_limit
 * Pass far_limit = +inf, near_limit = +inf for rays without a length (bounce rays, directional lights).
 * The caller must ignore the answer when !r->usable.
 */
TRT_HD void trt_cert_sphere2(const trt_cert_ray *r, float cx, float cy, float cz, float r_pad, float near_limit, float far_limit,
                             int *miss, int *blocks)
{
    const float ocx = cx - r->ox, ocy = cy - r->oy, ocz = cz - r->oz;
    const float tc = fmaf(ocz, r->dz, fmaf(ocy, r->dy, ocx * r->dx));
    const float wx = fmaf(-tc, r->dx, ocx), wy = fmaf(-tc, r->dy, ocy), wz = fmaf(-tc, r->dz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    const float outer = r_pad + r->slack_t;
    const float inner = fmaf(r_pad, 0.99999237060546875f, -r->slack_t);
    const float front = tc - r_pad;
    *miss = (h2 > outer * outer) || (tc < -r->slack_t) || (front > far_limit);
    *blocks = (inner > 0.0f) && (h2 < inner * inner) && (front > r->slack_t) && (tc < near_limit);
}

TRT_HD int trt_cert_sphere(const trt_cert_ray *r, float cx, float cy, float cz, float r_pad, float near_limit, float far_limit)
{
    const float ocx = cx - r->ox, ocy = cy - r->oy, ocz = cz - r->oz;
    const float tc = fmaf(ocz, r->dz, fmaf(ocy, r->dy, ocx * r->dx));
    const float wx = fmaf(-tc, r->dx, ocx), wy = fmaf(-tc, r->dy, ocy), wz = fmaf(-tc, r->dz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    const float outer = r_pad + r->slack_t;
    const float inner = fmaf(r_pad, 0.99999237060546875f, -r->slack_t);
    const float front = tc - r_pad;
    int out = 0;
    if ((h2 > outer * outer) || (tc < -r->slack_t) || (front > far_limit)) out |= TRT_CERT_MISS;
    if ((inner > 0.0f) && (h2 < inner * inner) && (front > r->slack_t) && (tc < near_limit)) out |= TRT_CERT_BLOCKS;
    return out;
}

/* miss-only variant for rays whose hit POINT is needed (bounce rays): 16 float operations */
TRT_HD int trt_cert_sphere_miss(const trt_cert_ray *r, float cx, float cy, float cz, float r_pad)
{
    const float ocx = cx - r->ox, ocy = cy - r->oy, ocz = cz - r->oz;
    const float tc = fmaf(ocz, r->dz, fmaf(ocy, r->dy, ocx * r->dx));
    const float wx = fmaf(-tc, r->dx, ocx), wy = fmaf(-tc, r->dy, ocy), wz = fmaf(-tc, r->dz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    const float outer = r_pad + r->slack_t;
    return (h2 > outer * outer) || (tc < -r->slack_t);
}

/*
 * Ground plane, TRT.c:677-695: hit iff |denom| > 1e-5 and t = ((p - o) . n) / (d . n) > 1e-5.
 * When numerator and denominator certainly have opposite signs the quotient is negative: miss.
 * (px,py,pz) plane point, (nx,ny,nz) plane normal (any length), both rounded to nearest.
 * Error bounds: |num_f - num| <= 8u * sum_k (|p_k| + |o_k|)|n_k|, |den_f - den| <= 8u * sum_k |d_k||n_k|;
 * the margins are twice that.
 */
TRT_HD int trt_cert_plane_miss(const trt_cert_ray *r, float px, float py, float pz, float nx, float ny, float nz)
{
    const float anx = fabsf(nx), any_ = fabsf(ny), anz = fabsf(nz);
    const float num = fmaf(pz - r->oz, nz, fmaf(py - r->oy, ny, (px - r->ox) * nx));
    const float den = fmaf(r->dz, nz, fmaf(r->dy, ny, r->dx * nx));
    const float num_scale = fmaf(fabsf(pz) + fabsf(r->oz), anz, fmaf(fabsf(py) + fabsf(r->oy), any_, (fabsf(px) + fabsf(r->ox)) * anx));
    const float den_scale = fmaf(fabsf(r->dz), anz, fmaf(fabsf(r->dy), any_, fabsf(r->dx) * anx));
    const float e_num = (16.0f * TRT_CERT_U) * num_scale;
    const float e_den = (16.0f * TRT_CERT_U) * den_scale;
    const int finite = r->usable && (num_scale < 1e30f) && (den_scale < 1e30f);
    return finite && ((num < -e_num && den > e_den) || (num > e_num && den < -e_den));
}

/* The same when the caller holds the numerator the reference itself computes, (p - o) . n in double with the reference's
 * operations (TRT.c:684-685): its sign is exact, only the denominator's has to be certified.  A zero numerator gives t = 0,
 * which is not > 1e-5: a miss as well. */
TRT_HD int trt_cert_plane_miss_num(const trt_cert_ray *r, double num, float nx, float ny, float nz)
{
    const float den = fmaf(r->dz, nz, fmaf(r->dy, ny, r->dx * nx));
    const float den_scale = fmaf(fabsf(r->dz), fabsf(nz), fmaf(fabsf(r->dy), fabsf(ny), fabsf(r->dx) * fabsf(nx)));
    const float e_den = (16.0f * TRT_CERT_U) * den_scale;
    const int finite = r->usable && (den_scale < 1e30f);
    return finite && ((num < 0.0 && den > e_den) || (num > 0.0 && den < -e_den) || num == 0.0);
}

/*
 * Tile-level certificate for PRIMARY rays.  All sample rays of a pixel tile leave the eye with directions
 * D = Dc + e, |e| <= h (Dc: direction through the tile centre, un-normalised; h: half extent of the tile on the
 * camera plane times the basis lengths).  For the sphere with centre c (oc = eye - c):
 *     |oc x D| >= |oc x Dc| - |oc| h      and      |D| <= |Dc| + h
 * so if  |oc x Dc| - |oc| h  >  r_pad (|Dc| + h) + slack  every such line passes the sphere at more than r_pad:
 * the discriminant of TRT.c:651 is negative for every sample of the tile.  S = |eye|_1 + max_i |c_i|_1.
 */
TRT_HD int trt_cert_tile_sphere_miss(float ex, float ey, float ez, float Dx, float Dy, float Dz, float h,
                                     float cx, float cy, float cz, float r_pad, float S)
{
    const float ocx = ex - cx, ocy = ey - cy, ocz = ez - cz;
    const float X = fmaf(ocy, Dz, -(ocz * Dy));
    const float Y = fmaf(ocz, Dx, -(ocx * Dz));
    const float Z = fmaf(ocx, Dy, -(ocy * Dx));
    const float nX = TRT_CERT_SQRT(fmaf(Z, Z, fmaf(Y, Y, X * X))) * (1.0f - 32.0f * TRT_CERT_U);
    const float nOC = TRT_CERT_SQRT(fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx))) * (1.0f + 32.0f * TRT_CERT_U);
    const float nD = TRT_CERT_SQRT(fmaf(Dz, Dz, fmaf(Dy, Dy, Dx * Dx))) * (1.0f + 32.0f * TRT_CERT_U) + h;
    const float lhs = nX - nOC * h;
    const float rhs = fmaf(r_pad, nD, (64.0f * TRT_CERT_U) * S * nD);
    return (S < 1e15f) && (nD < 1e15f) && (lhs > rhs);
}

/*
 * Tile-level ground certificate for primary rays: every D = Dc + ex*bx + ey*by (|ex| <= hx, |ey| <= hy) has
 * D . n of one certain sign.  Returns +1 (all positive), -1 (all negative) or 0 (unknown).
 * dn = Dc . n, bxn = bx . n, byn = by . n (float), scale = (|Dc|_1 + hx |bx|_1 + hy |by|_1) * |n|_1.
 */
TRT_HD int trt_cert_tile_plane_sign(float dn, float bxn, float byn, float hx, float hy, float scale)
{
    const float spread = fmaf(hx, fabsf(bxn), hy * fabsf(byn));
    const float margin = (64.0f * TRT_CERT_U) * scale;
    if (!(scale < 1e30f)) return 0;
    if (dn - spread > margin) return 1;
    if (dn + spread < -margin) return -1;
    return 0;
}

/* ---- host-side preparation shared by the library (trt_api.cu) and the CPU checker ----------------------- */

/* smallest float >= v */
TRT_HD float trt_cert_round_up(double v)
{
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}

/* padded radius of a cull record: >= r (1 + 2^-20), rounded up; r = sqrt(fl(radius*radius)) covers negative radii */
TRT_HD float trt_cert_pad_radius(double radius)
{
    const double r = sqrt(radius * radius);
    return trt_cert_round_up(r * (1.0 + 1.0 / 1048576.0));
}

/* camera as the tile certificates see it (all float, rounded to nearest unless stated) */
typedef struct {
    float ex, ey, ez;             /* eye                                                            */
    float bx[3], by[3], bz[3];    /* camera basis                                                   */
    float nbx, nby;               /* |bx|_2, |by|_2 rounded up                                      */
    float sw, sh, dist;           /* screen width, height, distance                                 */
    float pw, ph;                 /* sw / W, sh / H for the W x H screen being rendered (rounded to nearest): passed to trt_cert_tile_cone */
    float off_x, off_y;           /* largest sub-pixel offset in x and y, in pixels (TRT.c:992-993) */
} trt_cert_camera;

/* Direction through the centre of the tile [col0,col0+tw) x [row0,row0+th) of a W x H screen, and the half
 * extents (hx, hy, in camera-plane units) of the tile's sample positions around it, as the primary-ray
 * construction of TRT.c:987-1005 places them (top-left pixel corner + offset in [0, off] pixels; screen_y negated
 * before the offset is added).  Every sample direction of the tile is Dc + ex*bx + ey*by, |ex| <= hx, |ey| <= hy;
 * the sphere certificate takes h = hx*|bx| + hy*|by|. */
TRT_HD void trt_cert_tile_cone(const trt_cert_camera *c, float pw, float ph, int col0, int row0, int tw, int th,
                               float *Dx, float *Dy, float *Dz, float *hx, float *hy)
{
    const float half_w = 0.5f * ((float)(tw - 1) + c->off_x), half_h = 0.5f * ((float)(th - 1) + c->off_y);
    const float scx = fmaf((float)col0 + half_w, pw, -0.5f * c->sw);
    const float syc = fmaf(-((float)(row0 + th - 1) - half_h), ph, 0.5f * c->sh);
    *hx = fmaf(half_w + 0.01f, pw, (16.0f * TRT_CERT_U) * c->sw);
    *hy = fmaf(half_h + 0.01f, ph, (16.0f * TRT_CERT_U) * c->sh);
    *Dx = fmaf(c->bz[0], -c->dist, fmaf(c->by[0], syc, c->bx[0] * scx)) - c->ex;
    *Dy = fmaf(c->bz[1], -c->dist, fmaf(c->by[1], syc, c->bx[1] * scx)) - c->ey;
    *Dz = fmaf(c->bz[2], -c->dist, fmaf(c->by[2], syc, c->bx[2] * scx)) - c->ez;
}

/* Point-light shadow ray vs the ground plane (TRT.c:677-695 through 936-941).  num = (plane point - origin) . n
 * in double, as the reference evaluates it; light_height = (light - plane point) . n; margin > 0 scales with |n|.
 * When origin and light are on the same side of the plane the ray cannot reach the plane before it reaches the
 * light, and any hit behind the light is farther than the light by at least margin/|n|: it cannot block. */
TRT_HD int trt_cert_ground_cannot_block(double num, double light_height, double margin)
{
    return (num < 0.0 && light_height > margin) || (num > 0.0 && light_height < -margin);
}

/* ---- patch certificates: all first-generation ground hits of a pixel tile at once ----------------------------
 * When no sphere can be hit by a tile's primary rays (tile sphere certificate) every surface hit of the tile is a
 * GROUND hit, and all of them lie in the quadrilateral the tile's four corner directions cut out of the plane
 * (the map direction -> plane point is projective, and keeps convexity while direction . n keeps its sign).
 * trt_cert_patch_ball bounds that quadrilateral, and with it the origins `at` of the tile's shadow and bounce rays,
 * by a ball; the three functions after it decide once per tile which spheres those rays can reach at all.  Spheres
 * they rule out are certain misses for every such ray (or, for a point light, certain not to be hit closer than
 * the light); the rest go through the per-ray certificates as usual. */

typedef struct {
    float cx, cy, cz;  /* centre                                                                      */
    float r;           /* radius, padded for float error, for the EPSILON push-back and by a safety term */
    int ok;            /* 0: the tile reaches the horizon / numbers out of range: no patch certificate  */
} trt_cert_ball;

/* (Dx,Dy,Dz), hx, hy: trt_cert_tile_cone; num: (plane point - eye) . n as a float (its sign must be robust, the
 * caller checks that); (nx,ny,nz): plane normal; S: |eye|_1 + scene scale */
TRT_HD void trt_cert_patch_ball(const trt_cert_camera *c, float Dx, float Dy, float Dz, float hx, float hy, float num,
                                float nx, float ny, float nz, float S, trt_cert_ball *out)
{
    float px[4], py[4], pz[4], reach = 0.0f;
    int ok = 1;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int k = 0; k < 4; k++) {
        const float sxk = (k & 1) ? hx : -hx, syk = (k & 2) ? hy : -hy;
        const float dx = fmaf(c->by[0], syk, fmaf(c->bx[0], sxk, Dx));
        const float dy = fmaf(c->by[1], syk, fmaf(c->bx[1], sxk, Dy));
        const float dz = fmaf(c->by[2], syk, fmaf(c->bx[2], sxk, Dz));
        const float dn = fmaf(dz, nz, fmaf(dy, ny, dx * nx));
        const float dscale = (fabsf(dx) + fabsf(dy) + fabsf(dz)) * (fabsf(nx) + fabsf(ny) + fabsf(nz));
        /* the corner ray must meet the plane in front of the eye, at a well-conditioned angle */
        ok = ok && (fabsf(dn) > 1e-3f * dscale) && ((dn < 0.0f) == (num < 0.0f));
        const float t = TRT_CERT_DIV(num, dn);
        px[k] = fmaf(t, dx, c->ex);
        py[k] = fmaf(t, dy, c->ey);
        pz[k] = fmaf(t, dz, c->ez);
        const float ax = t * dx, ay = t * dy, az = t * dz;
        reach = fmaxf(reach, fabsf(ax) + fabsf(ay) + fabsf(az));
    }
    out->cx = 0.25f * ((px[0] + px[1]) + (px[2] + px[3]));
    out->cy = 0.25f * ((py[0] + py[1]) + (py[2] + py[3]));
    out->cz = 0.25f * ((pz[0] + pz[1]) + (pz[2] + pz[3]));
    float r2 = 0.0f;
    for (int k = 0; k < 4; k++) {
        const float ax = px[k] - out->cx, ay = py[k] - out->cy, az = pz[k] - out->cz;
        r2 = fmaxf(r2, fmaf(az, az, fmaf(ay, ay, ax * ax)));
    }
    /* conditioning 1e-3 => the float quotient t is within ~2^-10 * 1e-2 of the real one; 4e-3 * reach covers it */
    out->r = TRT_CERT_SQRT(r2) * 1.001f + 4e-3f * reach + 1e-4f * (1.0f + S) * 1e-2f + 1e-5f;
    out->ok = ok && (S < 1e15f) && (reach < 1e15f) && (out->r < 1e15f);
}

/* directional light with unit direction (lx,ly,lz): can a ray from some point of the ball hit the sphere? */
TRT_HD int trt_cert_patch_dir_candidate(const trt_cert_ball *b, float lx, float ly, float lz, float cx, float cy, float cz,
                                        float r_pad, float S)
{
    const float ocx = cx - b->cx, ocy = cy - b->cy, ocz = cz - b->cz;
    const float tc = fmaf(ocz, lz, fmaf(ocy, ly, ocx * lx));
    const float wx = fmaf(-tc, lx, ocx), wy = fmaf(-tc, ly, ocy), wz = fmaf(-tc, lz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    const float slack = (64.0f * TRT_CERT_U) * S;
    const float reach = r_pad + b->r + slack;
    /* every line passes the centre at more than r, or the centre is behind every origin (per-ray: tc < -slack) */
    return !((h2 > reach * reach) || (tc + b->r < -slack));
}

/* point light at (lx,ly,lz): can the sphere come between some point of the ball and the light? */
TRT_HD int trt_cert_patch_point_candidate(const trt_cert_ball *b, float lx, float ly, float lz, float cx, float cy, float cz,
                                          float r_pad, float S)
{
    const float vx = lx - b->cx, vy = ly - b->cy, vz = lz - b->cz;
    const float dd = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
    const float slack = (64.0f * TRT_CERT_U) * S + 2e-5f;
    const float reach = r_pad + b->r + slack;
    if (!(dd > 1e-20f)) return 1;
    /* closest point of the segment [ball centre, light] (stretched by `reach` at both ends) to the sphere centre */
    const float ocx = cx - b->cx, ocy = cy - b->cy, ocz = cz - b->cz;
    const float s = TRT_CERT_DIV(fmaf(ocz, vz, fmaf(ocy, vy, ocx * vx)), dd);
    const float sc = fminf(fmaxf(s, 0.0f), 1.0f);
    const float wx = fmaf(-sc, vx, ocx), wy = fmaf(-sc, vy, ocy), wz = fmaf(-sc, vz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    return !(h2 > reach * reach);
}

/* bounce rays off the ground: origins in the ball, directions R(D) for the tile's primary directions D = Dc + e,
 * |e| <= h, R = reflection about the plane (linear, length-preserving): (Rx,Ry,Rz) = R(Dc). */
TRT_HD int trt_cert_patch_bounce_candidate(const trt_cert_ball *b, float Rx, float Ry, float Rz, float h, float cx, float cy, float cz,
                                           float r_pad, float S)
{
    return !trt_cert_tile_sphere_miss(b->cx, b->cy, b->cz, Rx, Ry, Rz, h, cx, cy, cz, r_pad + b->r, S);
}

/* ---- skybox texel certificate ------------------------------------------------------------------------------------
 * get_skybox_color (TRT.c:700-789) turns a direction into a cube face and a texel index: a DECISION.  Evaluated in
 * float, face and index are certainly the reference's when (a) the largest |component| beats the runner-up by more
 * than the float error (face argmax, TRT.c:703-713) and (b) both scaled coordinates (u + 0.5) * dim lie farther from
 * an integer than their float error (truncation, TRT.c:782-783; this also excludes the clamped edges u, v = +-0.5,
 * where the reference's index runs into the next row / past the plane).  (dx,dy,dz): the ray's unit direction rounded
 * to float (the reference normalises once more, TRT.c:702: a change in the last double bit, far below float).
 * Float error of the scaled coordinate (u + 0.5) * dim: inputs rounded to float (u), the approximate quotient 0.5 / m (2 ulp =
 * 4u), two products and a sum (3u), all on values <= 1, then the scaling: <= 8u * dim; margin 12u * dim + 2e-5.  (The margin
 * decides how often the exact double path runs: every failing lane costs its whole warp ~150 instructions, and with
 * 16u * dim + 1e-4 at dim = 1024 that was 0.43 % of the lanes, 13 % of the warp-level lookups, 5 % of K1's stall samples.)
 * Returns 1 and sets *face, *texel when certain, else 0. */
TRT_HD int trt_cert_sky_texel(float dx, float dy, float dz, int dim, int *face, int *texel)
{
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    /* candidates in the reference's order +x,-x,+y,-y,+z,-z; with clear margins ties cannot occur */
    int best;
    float m, second;
    if (ax >= ay && ax >= az) { best = dx > 0.0f ? 0 : 1; m = ax; second = fmaxf(ay, az); }
    else if (ay >= az) { best = dy > 0.0f ? 2 : 3; m = ay; second = fmaxf(ax, az); }
    else { best = dz > 0.0f ? 4 : 5; m = az; second = fmaxf(ax, ay); }
    if (!(m - second > 8.0f * TRT_CERT_U) || !(m > 0.5f) || !(m < 1.0001f)) return 0;
    /* project onto the face (TRT.c:717-727): dir / major component, halved; the two face coordinates are dots with
     * axes (best+2)%6 and (best+4)%6 — negative axes for the odd faces */
    const float s = TRT_CERT_DIV(0.5f, m);
    const float px = dx * s, py = dy * s, pz = dz * s;
    float u, v;
    if (best <= 1) { u = py; v = pz; }
    else if (best <= 3) { u = pz; v = px; }
    else { u = px; v = py; }
    if (best & 1) { u = -u; v = -v; }
    if (best & 1) u = -u;                                         /* :730 */
    if (best <= 1) { const float t = u; u = v; v = -t; }          /* :735 */
    else if (best <= 3) { const float t = u; u = -v; v = t; }     /* :742-755 */
    else if (best == 4) { u = -u; v = -v; }                       /* :756 */
    const float fu = (u + 0.5f) * (float)dim, fv = (v + 0.5f) * (float)dim;
    const float margin = fmaf(12.0f * TRT_CERT_U, (float)dim, 2e-5f);
    const float iu = floorf(fu), iv = floorf(fv);
    if (!(fu - iu > margin) || !(iu + 1.0f - fu > margin) || !(fv - iv > margin) || !(iv + 1.0f - fv > margin)) return 0;
    if (!(iu >= 0.0f) || !(iu < (float)dim) || !(iv >= 0.0f) || !(iv < (float)dim) || dim > 16384) return 0;
    *face = best;
    *texel = (int)iu + (int)iv * dim;
    return 1;
}

/* ---- clusters: many-sphere scenes ---------------------------------------------------------------------------------
 * Spheres are sorted along a k-d tree (trt_cert_kd_order, host, once per upload) and every 32 consecutive ones get a bounding ball
 * (C, R): |c_i - C| + r_pad_i <= R for every member.  A ray that certainly passes the ball, or has the whole ball
 * behind its origin, or (point-light shadow rays) has the whole ball beyond the light, certainly misses every member
 * in the sense of trt_cert_sphere2's MISS: with tc = (C - o).d and h = distance of C from the ray's line, member i has
 * h_i >= h - (R - r_pad_i), tc_i <= tc + (R - r_pad_i) and tc_i - r_pad_i >= tc - R. */
TRT_HD int trt_cert_cluster_miss(const trt_cert_ray *r, float cx, float cy, float cz, float R, float far_limit)
{
    const float ocx = cx - r->ox, ocy = cy - r->oy, ocz = cz - r->oz;
    const float tc = fmaf(ocz, r->dz, fmaf(ocy, r->dy, ocx * r->dx));
    const float wx = fmaf(-tc, r->dx, ocx), wy = fmaf(-tc, r->dy, ocy), wz = fmaf(-tc, r->dz, ocz);
    const float h2 = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
    const float outer = R + r->slack_t;
    return (h2 > outer * outer) || (tc + R < -r->slack_t) || (tc - R > far_limit);
}

#include <stdlib.h>
/* host: bounding ball of `count` certificate records (4 floats each: centre, r_pad), out = (C, R rounded up) */
static inline void trt_cert_cluster_bound(const float *cull4, int count, float out[4])
{
    double c[3] = {0.0, 0.0, 0.0}, R = 0.0;
    for (int i = 0; i < count; i++)
        for (int k = 0; k < 3; k++) c[k] += cull4[4 * i + k];
    for (int k = 0; k < 3; k++) out[k] = (float)(c[k] / (count > 0 ? count : 1));
    for (int i = 0; i < count; i++) {
        const double dx = (double)cull4[4 * i] - out[0], dy = (double)cull4[4 * i + 1] - out[1], dz = (double)cull4[4 * i + 2] - out[2];
        const double reach = sqrt(dx * dx + dy * dy + dz * dz) * (1.0 + 1e-6) + cull4[4 * i + 3];
        if (reach > R) R = reach;
    }
    out[3] = trt_cert_round_up(R * (1.0 + 1e-6));
}

/* host: order[] along a k-d tree — recursive median split along the longest axis of the centres' bounding box, until
 * groups of 8 remain — so that every aligned group of 8 (and of 32) consecutive spheres is spatially compact.  (On the 1024-sphere
 * stress scene: 85 % of the balls of 32 are certainly missed by a ray, against 55 % with a Morton order.) */
typedef struct { float key; int index; } trt_cert_kd_item;
static inline int trt_cert_kd_cmp(const void *a, const void *b)
{
    const trt_cert_kd_item *x = (const trt_cert_kd_item *)a, *y = (const trt_cert_kd_item *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0);
}
static inline void trt_cert_kd_split(const float *cull4, int *order, int lo, int hi, trt_cert_kd_item *tmp)
{
    const int n = hi - lo;
    if (n <= 8) return;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = lo; j < hi; j++)
        for (int k = 0; k < 3; k++) {
            const float v = cull4[4 * order[j] + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
    int axis = 0;
    if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
    if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
    for (int j = lo; j < hi; j++) {
        tmp[j - lo].key = cull4[4 * order[j] + axis];
        tmp[j - lo].index = order[j];
    }
    qsort(tmp, (size_t)n, sizeof(trt_cert_kd_item), trt_cert_kd_cmp);
    for (int j = lo; j < hi; j++) order[j] = tmp[j - lo].index;
    /* split at a multiple of 8 nearest the middle, so that the leaves line up with the groups of 8 and 32 */
    int half = ((n / 2 + 4) / 8) * 8;
    if (half <= 0) half = 8;
    if (half >= n) half = n - (n % 8 ? n % 8 : 8);
    trt_cert_kd_split(cull4, order, lo, lo + half, tmp);
    trt_cert_kd_split(cull4, order, lo + half, hi, tmp);
}
static inline void trt_cert_kd_order(const float *cull4, int n, int *order)
{
    trt_cert_kd_item *tmp = (trt_cert_kd_item *)malloc(sizeof(trt_cert_kd_item) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) order[i] = i;
    trt_cert_kd_split(cull4, order, 0, n, tmp);
    free(tmp);
}

#endif /* TRT_CERT_H */
