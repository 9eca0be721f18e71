/*
 * oracle/cert_check.c — CPU audit of the product's single-precision certificates.  TEST INFRASTRUCTURE ONLY.
 *
 * Includes the product header terminalraytracer_b200/csrc/trt_cert.h (the same code the render kernel compiles)
 * and walks the reference's render loop (through the oracle restatement, oracle/trt_oracle.c) over a frame.
 * For every closest-hit query the loop performs — primary rays, bounce rays, directional-light and point-light
 * shadow rays (TRT.c:966-1069, 894-963) — it evaluates the certificates exactly as trt_render.cu uses them and
 * compares each decision (certified in float, or taken from exact tests on the certificates' survivors only) with
 * the oracle's exact one over all objects.  Returns the number of contradictions (must be
 * 0) and fills work statistics (how many exact tests survive, how many shadow queries stay "unknown").
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TRT_NO_REFERENCE_NAMES
#include "trt_types.h"
#include "../terminalraytracer_b200/csrc/trt_cert.h"

typedef trt_Vector v3;

int orc_hit_sphere(const trt_Ray *ray, const trt_Sphere *s, trt_Point *hit, void *ctr);
int orc_hit_plane(const trt_Ray *ray, const trt_Plane *p, trt_Point *hit, void *ctr);
trt_ObjectType orc_closest_hit(const trt_Scene *scene, const trt_Ray *ray, trt_Point *hit_out, trt_Vector *normal_out,
                               trt_Material *material_out, void *ctr);
void orc_subpixel_offsets(double *dx, double *dy);
void orc_sky_texel(const trt_Skybox *sky, const trt_Vector *direction, trt_Color *color, int *face_out, long *index_out);

static inline double dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 sub3(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline v3 add3(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline v3 scale3(v3 a, double s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline v3 unit3(v3 a)
{
    double l = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (l > 0.0001) { a.x /= l; a.y /= l; a.z /= l; }
    return a;
}

enum {
    ST_PRIMARY = 0, ST_PRIMARY_TILE_SURVIVORS, ST_PRIMARY_GROUND_CULLED,
    ST_BOUNCE, ST_BOUNCE_SURVIVORS, ST_BOUNCE_GROUND_CULLED, ST_BOUNCE_EXACT_HITS,
    ST_DIR, ST_DIR_OPEN, ST_DIR_BLOCKED, ST_DIR_UNKNOWN,
    ST_POINT, ST_POINT_OPEN, ST_POINT_BLOCKED, ST_POINT_UNKNOWN,
    ST_SHADOW_EXACT_TESTS,
    ST_SKY, ST_SKY_CERTIFIED, ST_CLUSTER_TESTS, ST_CLUSTERS_MISSED, ST_SUBCLUSTER_TESTS, ST_SUBCLUSTERS_MISSED,
    ST_TILES, ST_TILES_PATCH, ST_TILES_PATCH_EMPTY, ST_PATCH_DIR_CANDIDATES, ST_PATCH_POINT_CANDIDATES, ST_PATCH_BOUNCE_CANDIDATES, ST_PATCH_RECORDS,
    ST_COUNT
};

#define TILE_W 8
#define TILE_H 4

typedef struct {
    const trt_Scene *scene;
    float *cull; /* 4 floats per sphere: centre, r_pad */
    float centre_l1;
    float gp[3], gn[3];
    long long *st;
    long long bad;
    unsigned char *scratch; int num_clusters; int *order; float *cluster4; float *sub4; /* sub4: balls of 8 inside each cluster */ /* k-d order of the spheres and bounding balls of 32 (many-sphere scenes) */
    const unsigned char *cand; /* patch certificate: spheres still possible for the query at hand; NULL = all */
} ctx_t;

/* exact decision of a shadow query as apply_lighting takes it (TRT.c:907, 936-941) */
static int exact_dir_open(const trt_Scene *scene, const trt_Ray *ray)
{
    return orc_closest_hit(scene, ray, NULL, NULL, NULL, NULL) == TRT_NONE;
}
static int exact_point_open(const trt_Scene *scene, const trt_Ray *ray, v3 P, double light_d2)
{
    trt_Point blocker;
    trt_ObjectType what = orc_closest_hit(scene, ray, &blocker, NULL, NULL, NULL);
    v3 bv = {blocker.x, blocker.y, blocker.z};
    v3 to_blocker = sub3(bv, P);
    return what == TRT_NONE || light_d2 < dot3(to_blocker, to_blocker);
}

/* members of a cluster whose bounding ball the ray certainly misses are certain misses: flags per sphere (reference index) */
static void cluster_flags(const ctx_t *c, const trt_cert_ray *r, float far_limit, unsigned char *missed)
{
    const int n = c->scene->num_spheres;
    memset(missed, 0, (size_t)(n > 0 ? n : 1));
    if (!c->num_clusters || !r->usable) return;
    for (int k = 0; k < c->num_clusters; k++)
        if (trt_cert_cluster_miss(r, c->cluster4[4 * k], c->cluster4[4 * k + 1], c->cluster4[4 * k + 2], c->cluster4[4 * k + 3], far_limit)) {
            c->st[ST_CLUSTERS_MISSED]++;
            for (int j = 32 * k; j < 32 * k + 32 && j < n; j++) missed[c->order[j]] = 1;
        } else {
            /* second level: four balls of 8 */
            for (int q = 0; q < 4 && 32 * k + 8 * q < n; q++) {
                const float *b = c->sub4 + 4 * (4 * k + q);
                c->st[ST_SUBCLUSTER_TESTS]++;
                if (trt_cert_cluster_miss(r, b[0], b[1], b[2], b[3], far_limit)) {
                    c->st[ST_SUBCLUSTERS_MISSED]++;
                    for (int j = 32 * k + 8 * q; j < 32 * k + 8 * q + 8 && j < n; j++) missed[c->order[j]] = 1;
                }
            }
        }
    c->st[ST_CLUSTER_TESTS] += c->num_clusters;
}

static void check_bounce(ctx_t *c, const trt_Ray *ray)
{
    const trt_Scene *s = c->scene;
    trt_cert_ray r;
    const float S = trt_cert_set_origin(&r, ray->origin.x, ray->origin.y, ray->origin.z) + c->centre_l1;
    trt_cert_set_unit_dir(&r, ray->direction.x, ray->direction.y, ray->direction.z, S);
    c->st[ST_BOUNCE]++;
    unsigned char *ball_missed = c->scratch;
    cluster_flags(c, &r, INFINITY, ball_missed);
    for (int i = 0; i < s->num_spheres; i++) {
        trt_Point p;
        const int hit = orc_hit_sphere(ray, &s->spheres[i], &p, NULL);
        const int miss = (c->cand && !c->cand[i]) || ball_missed[i] ||
                         (r.usable && trt_cert_sphere_miss(&r, c->cull[4 * i], c->cull[4 * i + 1], c->cull[4 * i + 2], c->cull[4 * i + 3]));
        if (!miss) c->st[ST_BOUNCE_SURVIVORS]++;
        if (hit) c->st[ST_BOUNCE_EXACT_HITS]++;
        if (miss && hit) c->bad++;
    }
    trt_Point p;
    const int ghit = orc_hit_plane(ray, &s->ground, &p, NULL);
    const int gmiss = trt_cert_plane_miss(&r, c->gp[0], c->gp[1], c->gp[2], c->gn[0], c->gn[1], c->gn[2]);
    if (gmiss) c->st[ST_BOUNCE_GROUND_CULLED]++;
    if (gmiss && ghit) c->bad++;
    {
        /* the form the kernel's bounce queries use: the reference's own numerator (TRT.c:684-685), float denominator */
        const double tx = s->ground.point.x - ray->origin.x, ty = s->ground.point.y - ray->origin.y, tz = s->ground.point.z - ray->origin.z;
        const double num = tx * s->ground.normal.x + ty * s->ground.normal.y + tz * s->ground.normal.z;
        const int gmiss_num = trt_cert_plane_miss_num(&r, num, c->gn[0], c->gn[1], c->gn[2]);
        if (gmiss_num && ghit) c->bad++;
        if (gmiss && !gmiss_num) c->bad++;          /* it must not be weaker than the all-float form */
    }
}

/* closest hit among the SURVIVORS of the certificates (spheres in index order, strict <, then the ground), as
 * trace_ray orders them (TRT.c:805-853); returns 0 when none of them is hit, else 1 and the pushed-back point */
static int closest_survivor(const trt_Scene *s, const trt_Ray *ray, const unsigned char *survivor, int ground_survivor, v3 *blocker)
{
    double closest = INFINITY;
    int any = 0;
    v3 o = {ray->origin.x, ray->origin.y, ray->origin.z}, best = o;
    trt_Point p;
    for (int i = 0; i < s->num_spheres; i++)
        if (survivor[i] && orc_hit_sphere(ray, &s->spheres[i], &p, NULL)) {
            v3 pv = {p.x, p.y, p.z};
            v3 back = sub3(o, pv);
            const double d2 = dot3(back, back);
            if (d2 < closest) { closest = d2; best = pv; any = 1; }
        }
    if (ground_survivor && orc_hit_plane(ray, &s->ground, &p, NULL)) {
        v3 pv = {p.x, p.y, p.z};
        v3 back = sub3(o, pv);
        const double d2 = dot3(back, back);
        if (d2 < closest) { closest = d2; best = pv; any = 1; }
    }
    if (any) {
        v3 back = scale3(unit3(sub3(o, best)), TRT_EPSILON);
        *blocker = add3(best, back);
    }
    return any;
}

/* the kernel's decision procedure for a directional light: 1 open, 0 blocked.  *certified = decided in float */
static int cert_dir(ctx_t *c, v3 at, v3 L, int *certified, unsigned char *survivor)
{
    const trt_Scene *s = c->scene;
    trt_cert_ray r;
    const float S = trt_cert_set_origin(&r, at.x, at.y, at.z) + c->centre_l1;
    trt_cert_set_unit_dir(&r, L.x, L.y, L.z, S);
    int any_blocks = 0, survivors = 0;
    cluster_flags(c, &r, INFINITY, c->scratch);
    for (int i = 0; i < s->num_spheres; i++) {
        const int k = ((c->cand && !c->cand[i]) || c->scratch[i]) ? TRT_CERT_MISS :
                      (r.usable ? trt_cert_sphere(&r, c->cull[4 * i], c->cull[4 * i + 1], c->cull[4 * i + 2], c->cull[4 * i + 3], INFINITY, INFINITY) : 0);
        any_blocks |= (k & TRT_CERT_BLOCKS) != 0;
        survivor[i] = !(k & TRT_CERT_MISS);
        survivors += survivor[i];
    }
    *certified = 1;
    if (any_blocks) return 0;
    /* ground: sign logic on the reference's own numerator and denominator, the division only when they agree */
    v3 gp = {s->ground.point.x, s->ground.point.y, s->ground.point.z};
    const double denom = dot3(L, s->ground.normal);
    if (fabs(denom) > 0.00001) {
        const double num = dot3(sub3(gp, at), s->ground.normal);
        if (num != 0.0 && ((num < 0.0) == (denom < 0.0))) {
            const double t = num / denom;
            if (t > 0.00001) return 0;
        }
    }
    if (!survivors) return 1;
    *certified = 0;
    c->st[ST_SHADOW_EXACT_TESTS] += survivors;
    trt_Ray ray = {{at.x, at.y, at.z}, L};
    v3 blocker;
    return !closest_survivor(s, &ray, survivor, 0, &blocker);
}

static int cert_point(ctx_t *c, v3 at, const trt_PointLight *pl, int *certified, unsigned char *survivor)
{
    const trt_Scene *s = c->scene;
    trt_cert_ray r;
    const float lx = (float)pl->position.x, ly = (float)pl->position.y, lz = (float)pl->position.z;
    const float S = trt_cert_set_origin(&r, at.x, at.y, at.z) + c->centre_l1 + (fabsf(lx) + fabsf(ly) + fabsf(lz));
    const float dist = trt_cert_set_dir_toward(&r, lx, ly, lz, S);
    const float guard = fmaf(2.0f, r.slack_t, 1e-5f);
    const float near_limit = dist - guard, far_limit = dist + guard;
    int any_blocks = 0, survivors = 0;
    cluster_flags(c, &r, far_limit, c->scratch);
    for (int i = 0; i < s->num_spheres; i++) {
        const int k = ((c->cand && !c->cand[i]) || c->scratch[i]) ? TRT_CERT_MISS :
                      (r.usable ? trt_cert_sphere(&r, c->cull[4 * i], c->cull[4 * i + 1], c->cull[4 * i + 2], c->cull[4 * i + 3], near_limit, far_limit) : 0);
        any_blocks |= (k & TRT_CERT_BLOCKS) != 0;
        survivor[i] = !(k & TRT_CERT_MISS);
        survivors += survivor[i];
    }
    *certified = 1;
    if (any_blocks) return 0;
    v3 gp = {s->ground.point.x, s->ground.point.y, s->ground.point.z};
    v3 lp = {pl->position.x, pl->position.y, pl->position.z};
    const double num = dot3(sub3(gp, at), s->ground.normal);
    const double height = dot3(sub3(lp, gp), s->ground.normal);
    const double nl = sqrt(dot3(s->ground.normal, s->ground.normal));
    const double l1 = fabs(lp.x) + fabs(lp.y) + fabs(lp.z) + fabs(gp.x) + fabs(gp.y) + fabs(gp.z);
    const double margin = nl * (1e-4 + 1e-9 * l1);
    const int ground_survivor = !trt_cert_ground_cannot_block(num, height, margin);
    if (!survivors && !ground_survivor) return 1;
    *certified = 0;
    c->st[ST_SHADOW_EXACT_TESTS] += survivors;
    v3 Lr = sub3(lp, at);
    const double light_d2 = dot3(Lr, Lr);
    trt_Ray ray = {{at.x, at.y, at.z}, unit3(Lr)};
    v3 blocker;
    if (!closest_survivor(s, &ray, survivor, ground_survivor, &blocker)) return 1;
    v3 to_blocker = sub3(blocker, at);
    return light_d2 < dot3(to_blocker, to_blocker);
}

long long cert_check_rows(const trt_Scene *scene, int W, int H, int row0, int row1, long long *stats)
{
    ctx_t c;
    memset(&c, 0, sizeof c);
    c.scene = scene;
    c.st = stats;
    memset(stats, 0, sizeof(long long) * ST_COUNT);
    const int n = scene->num_spheres;
    c.cull = (float *)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
    double centre_l1 = 0.0;
    for (int i = 0; i < n; i++) {
        const trt_Sphere *sp = &scene->spheres[i];
        c.cull[4 * i] = (float)sp->center.x;
        c.cull[4 * i + 1] = (float)sp->center.y;
        c.cull[4 * i + 2] = (float)sp->center.z;
        c.cull[4 * i + 3] = trt_cert_pad_radius(sp->radius);
        const double l1 = fabs(sp->center.x) + fabs(sp->center.y) + fabs(sp->center.z);
        if (l1 > centre_l1) centre_l1 = l1;
    }
    c.centre_l1 = trt_cert_round_up(centre_l1 * (1.0 + 1.0 / 1048576.0));
    c.scratch = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    c.order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    c.num_clusters = n > 32 ? (n + 31) / 32 : 0;          /* TRT_CLUSTER_MIN_SPHERES of the library */
    c.cluster4 = (float *)malloc(sizeof(float) * 4 * (size_t)(c.num_clusters > 0 ? c.num_clusters : 1));
    c.sub4 = (float *)malloc(sizeof(float) * 16 * (size_t)(c.num_clusters > 0 ? c.num_clusters : 1));
    if (c.num_clusters) {
        trt_cert_kd_order(c.cull, n, c.order);
        float *sorted = (float *)malloc(sizeof(float) * 4 * (size_t)n);
        for (int j = 0; j < n; j++) memcpy(sorted + 4 * j, c.cull + 4 * c.order[j], sizeof(float) * 4);
        for (int k = 0; k < c.num_clusters; k++) {
            trt_cert_cluster_bound(sorted + 4 * 32 * k, n - 32 * k < 32 ? n - 32 * k : 32, c.cluster4 + 4 * k);
            for (int q = 0; q < 4; q++) {
                const int j0 = 32 * k + 8 * q, cnt = n - j0 < 8 ? n - j0 : 8;
                if (cnt > 0) trt_cert_cluster_bound(sorted + 4 * j0, cnt, c.sub4 + 4 * (4 * k + q));
                else memset(c.sub4 + 4 * (4 * k + q), 0, sizeof(float) * 4);
            }
        }
        free(sorted);
    }
    c.gp[0] = (float)scene->ground.point.x; c.gp[1] = (float)scene->ground.point.y; c.gp[2] = (float)scene->ground.point.z;
    c.gn[0] = (float)scene->ground.normal.x; c.gn[1] = (float)scene->ground.normal.y; c.gn[2] = (float)scene->ground.normal.z;

    const trt_Camera *cam = &scene->camera;
    double sdx[TRT_RAYS_PER_PIXEL], sdy[TRT_RAYS_PER_PIXEL];
    orc_subpixel_offsets(sdx, sdy);
    trt_cert_camera cc;
    cc.ex = (float)cam->frame.origin.x; cc.ey = (float)cam->frame.origin.y; cc.ez = (float)cam->frame.origin.z;
    const v3 *B[3] = {&cam->frame.basis.x, &cam->frame.basis.y, &cam->frame.basis.z};
    float *F[3] = {cc.bx, cc.by, cc.bz};
    for (int k = 0; k < 3; k++) { F[k][0] = (float)B[k]->x; F[k][1] = (float)B[k]->y; F[k][2] = (float)B[k]->z; }
    cc.nbx = trt_cert_round_up(sqrt(dot3(*B[0], *B[0])) * (1.0 + 1e-6));
    cc.nby = trt_cert_round_up(sqrt(dot3(*B[1], *B[1])) * (1.0 + 1e-6));
    cc.sw = (float)cam->screen_width; cc.sh = (float)cam->screen_height; cc.dist = (float)cam->screen_distance;
    cc.pw = (float)(cam->screen_width / W); cc.ph = (float)(cam->screen_height / H);
    double mx = 0, my = 0;
    for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++) { if (sdx[k] > mx) mx = sdx[k]; if (sdy[k] > my) my = sdy[k]; }
    cc.off_x = trt_cert_round_up(mx); cc.off_y = trt_cert_round_up(my);
    const float S_eye = fabsf(cc.ex) + fabsf(cc.ey) + fabsf(cc.ez) + c.centre_l1;
    v3 gp = {scene->ground.point.x, scene->ground.point.y, scene->ground.point.z};
    v3 eye = {cam->frame.origin.x, cam->frame.origin.y, cam->frame.origin.z};
    const double prim_num = dot3(sub3(gp, eye), scene->ground.normal);
    const double gl1 = fabs(gp.x) + fabs(gp.y) + fabs(gp.z) + fabs(eye.x) + fabs(eye.y) + fabs(eye.z);
    const double gnl1 = fabs(scene->ground.normal.x) + fabs(scene->ground.normal.y) + fabs(scene->ground.normal.z);
    const int prim_sign = prim_num < -1e-9 * gl1 * gnl1 ? -1 : (prim_num > 1e-9 * gl1 * gnl1 ? 1 : 0);

    unsigned char *tile_miss = (unsigned char *)malloc(2 * (size_t)(n > 0 ? n : 1)); /* + survivor flags of the shadow queries */
    unsigned char *patch_cand = (unsigned char *)malloc(33 * 32); /* [16 dir lights | 16 point lights | bounce][32 spheres] */
    for (int ty = row0; ty < row1; ty += TILE_H)
        for (int tx = 0; tx < W; tx += TILE_W) {
            /* tile certificates, as the kernel evaluates them once per tile */
            float Dx, Dy, Dz, hx, hy;
            trt_cert_tile_cone(&cc, cc.pw, cc.ph, tx, ty, TILE_W, TILE_H, &Dx, &Dy, &Dz, &hx, &hy);
            for (int i = 0; i < n; i++)
                tile_miss[i] = (unsigned char)trt_cert_tile_sphere_miss(cc.ex, cc.ey, cc.ez, Dx, Dy, Dz, fmaf(hx, cc.nbx, hy * cc.nby), c.cull[4 * i], c.cull[4 * i + 1],
                                                                        c.cull[4 * i + 2], c.cull[4 * i + 3], S_eye);
            const float dn = fmaf(Dz, c.gn[2], fmaf(Dy, c.gn[1], Dx * c.gn[0]));
            const float bxn = fmaf(cc.bx[2], c.gn[2], fmaf(cc.bx[1], c.gn[1], cc.bx[0] * c.gn[0]));
            const float byn = fmaf(cc.by[2], c.gn[2], fmaf(cc.by[1], c.gn[1], cc.by[0] * c.gn[0]));
            const float scale = (fabsf(Dx) + fabsf(Dy) + fabsf(Dz) + 2.0f * (hx * cc.nbx + hy * cc.nby)) * (fabsf(c.gn[0]) + fabsf(c.gn[1]) + fabsf(c.gn[2]));
            const int sgn = trt_cert_tile_plane_sign(dn, bxn, byn, hx, hy, scale);
            const int tile_ground_miss = (prim_sign < 0 && sgn > 0) || (prim_sign > 0 && sgn < 0);
            /* patch certificates: only when no sphere can be hit by the tile's primary rays */
            int patch_ok = n <= 32 && prim_sign != 0;
            for (int i = 0; i < n; i++) patch_ok = patch_ok && tile_miss[i];
            trt_cert_ball ball;
            ball.ok = 0;
            stats[ST_TILES]++;
            if (patch_ok) {
                trt_cert_patch_ball(&cc, Dx, Dy, Dz, hx, hy, (float)prim_num, c.gn[0], c.gn[1], c.gn[2], S_eye, &ball);
                patch_ok = ball.ok;
            }
            if (patch_ok) {
                stats[ST_TILES_PATCH]++;
                const float S_ball = fabsf(ball.cx) + fabsf(ball.cy) + fabsf(ball.cz) + ball.r + c.centre_l1;
                for (int l = 0; l < scene->num_directional_lights && l < 16; l++) {
                    v3 L = unit3(scale3(scene->directional_lights[l].direction, -1.0));
                    for (int i = 0; i < n; i++) {
                        patch_cand[(size_t)l * 32 + i] = (unsigned char)trt_cert_patch_dir_candidate(&ball, (float)L.x, (float)L.y, (float)L.z, c.cull[4 * i],
                                                                                                      c.cull[4 * i + 1], c.cull[4 * i + 2], c.cull[4 * i + 3], S_ball);
                        stats[ST_PATCH_DIR_CANDIDATES] += patch_cand[(size_t)l * 32 + i];
                    }
                }
                for (int l = 0; l < scene->num_point_lights && l < 16; l++) {
                    const trt_PointLight *pl = &scene->point_lights[l];
                    const float lx = (float)pl->position.x, ly = (float)pl->position.y, lz = (float)pl->position.z;
                    for (int i = 0; i < n; i++) {
                        patch_cand[(size_t)(16 + l) * 32 + i] = (unsigned char)trt_cert_patch_point_candidate(
                            &ball, lx, ly, lz, c.cull[4 * i], c.cull[4 * i + 1], c.cull[4 * i + 2], c.cull[4 * i + 3], S_ball + fabsf(lx) + fabsf(ly) + fabsf(lz));
                        stats[ST_PATCH_POINT_CANDIDATES] += patch_cand[(size_t)(16 + l) * 32 + i];
                    }
                }
                {
                    /* reflection of the tile's central direction about the plane, unit normal in float */
                    const float nl = sqrtf(fmaf(c.gn[2], c.gn[2], fmaf(c.gn[1], c.gn[1], c.gn[0] * c.gn[0])));
                    const float ux = c.gn[0] / nl, uy = c.gn[1] / nl, uz = c.gn[2] / nl;
                    const float dnn = 2.0f * fmaf(Dz, uz, fmaf(Dy, uy, Dx * ux));
                    const float Rx = fmaf(-dnn, ux, Dx), Ry = fmaf(-dnn, uy, Dy), Rz = fmaf(-dnn, uz, Dz);
                    const float h = fmaf(hx, cc.nbx, hy * cc.nby) * 1.0001f + (32.0f * TRT_CERT_U) * (fabsf(Dx) + fabsf(Dy) + fabsf(Dz));
                    for (int i = 0; i < n; i++) {
                        patch_cand[(size_t)32 * 32 + i] = (unsigned char)trt_cert_patch_bounce_candidate(&ball, Rx, Ry, Rz, h, c.cull[4 * i], c.cull[4 * i + 1],
                                                                                                          c.cull[4 * i + 2], c.cull[4 * i + 3], S_ball);
                        stats[ST_PATCH_BOUNCE_CANDIDATES] += patch_cand[(size_t)32 * 32 + i];
                    }
                }
                {
                    long long any = 0;
                    for (int l = 0; l < scene->num_directional_lights && l < 16; l++) for (int i = 0; i < n; i++) any += patch_cand[(size_t)l * 32 + i];
                    for (int l = 0; l < scene->num_point_lights && l < 16; l++) for (int i = 0; i < n; i++) any += patch_cand[(size_t)(16 + l) * 32 + i];
                    for (int i = 0; i < n; i++) any += patch_cand[(size_t)32 * 32 + i];
                    if (!any) stats[ST_TILES_PATCH_EMPTY]++;
                }
            }

            for (int row = ty; row < ty + TILE_H && row < row1; row++)
                for (int col = tx; col < tx + TILE_W && col < W; col++)
                    for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++) {
                        double pw = cam->screen_width / W, ph = cam->screen_height / H;
                        double sx = (((double)col / (double)W) * cam->screen_width - cam->screen_width / 2.0);
                        double sy = -(((double)row / (double)H) * cam->screen_height - cam->screen_height / 2.0);
                        double sz = -cam->screen_distance;
                        sx += sdx[k] * pw;
                        sy += sdy[k] * ph;
                        v3 dir = {0, 0, 0};
                        dir = add3(dir, scale3(cam->frame.basis.x, sx));
                        dir = add3(dir, scale3(cam->frame.basis.y, sy));
                        dir = add3(dir, scale3(cam->frame.basis.z, sz));
                        dir = unit3(sub3(dir, eye));
                        trt_Ray ray = {cam->frame.origin, dir};
                        int bounces = 0, going = 1;
                        double weight = 1.0;
                        while (going && bounces < TRT_BOUNCE_LIMIT && weight > 0.00001) {
                            if (bounces == 0 && ray.origin.x == eye.x && ray.origin.y == eye.y && ray.origin.z == eye.z) {
                                stats[ST_PRIMARY]++;
                                trt_Point p;
                                for (int i = 0; i < n; i++) {
                                    const int hit = orc_hit_sphere(&ray, &scene->spheres[i], &p, NULL);
                                    if (!tile_miss[i]) stats[ST_PRIMARY_TILE_SURVIVORS]++;
                                    if (tile_miss[i] && hit) c.bad++;
                                }
                                if (tile_ground_miss) {
                                    stats[ST_PRIMARY_GROUND_CULLED]++;
                                    if (orc_hit_plane(&ray, &scene->ground, &p, NULL)) c.bad++;
                                }
                            } else {
                                /* the bounce ray of a first-generation ground hit of a patch tile uses the tile's candidates */
                                c.cand = (patch_ok && bounces == 1) ? patch_cand + (size_t)32 * 32 : NULL;
                                check_bounce(&c, &ray);
                                c.cand = NULL;
                            }
                            trt_Point at;
                            v3 nrm;
                            trt_Material m;
                            trt_ObjectType what = orc_closest_hit(scene, &ray, &at, &nrm, &m, NULL);
                            if (what != TRT_NONE) {
                                v3 P = {at.x, at.y, at.z};
                                const int use_patch = patch_ok && bounces == 0;
                                if (use_patch) {
                                    stats[ST_PATCH_RECORDS]++;
                                    const double ddx = at.x - ball.cx, ddy = at.y - ball.cy, ddz = at.z - ball.cz;
                                    if (what != TRT_GROUND || sqrt(ddx * ddx + ddy * ddy + ddz * ddz) > ball.r) c.bad++; /* every such hit is a ground hit inside the ball */
                                }
                                for (int i = 0; i < scene->num_directional_lights; i++) {
                                    v3 L = unit3(scale3(scene->directional_lights[i].direction, -1.0));
                                    trt_Ray sh = {at, L};
                                    const int want = exact_dir_open(scene, &sh);
                                    int certified;
                                    c.cand = (use_patch && i < 16) ? patch_cand + (size_t)i * 32 : NULL;
                                    const int got = cert_dir(&c, P, L, &certified, tile_miss + n);
                                    c.cand = NULL;
                                    stats[ST_DIR]++;
                                    stats[!certified ? ST_DIR_UNKNOWN : (got ? ST_DIR_OPEN : ST_DIR_BLOCKED)]++;
                                    if (got != want) c.bad++;
                                }
                                for (int i = 0; i < scene->num_point_lights; i++) {
                                    const trt_PointLight *pl = &scene->point_lights[i];
                                    v3 lp = {pl->position.x, pl->position.y, pl->position.z};
                                    v3 Lr = sub3(lp, P);
                                    const double light_d2 = dot3(Lr, Lr);
                                    trt_Ray sh = {at, unit3(Lr)};
                                    const int want = exact_point_open(scene, &sh, P, light_d2);
                                    int certified;
                                    c.cand = (use_patch && i < 16) ? patch_cand + (size_t)(16 + i) * 32 : NULL;
                                    const int got = cert_point(&c, P, pl, &certified, tile_miss + n);
                                    c.cand = NULL;
                                    stats[ST_POINT]++;
                                    stats[!certified ? ST_POINT_UNKNOWN : (got ? ST_POINT_OPEN : ST_POINT_BLOCKED)]++;
                                    if (got != want) c.bad++;
                                }
                                weight *= m.reflectivity;
                                bounces++;
                            } else {
                                /* the sample ends in the sky: face and texel of the lookup, certificate vs reference */
                                int face = -1, texel = -1, rface = -1;
                                long rindex = -1;
                                trt_Color col;
                                stats[ST_SKY]++;
                                orc_sky_texel(&scene->skybox, &ray.direction, &col, &rface, &rindex);
                                if (trt_cert_sky_texel((float)ray.direction.x, (float)ray.direction.y, (float)ray.direction.z, scene->skybox.dim, &face, &texel)) {
                                    stats[ST_SKY_CERTIFIED]++;
                                    if (face != rface || (long)texel != rindex) c.bad++;
                                }
                                weight = 0.0;
                                going = 0;
                            }
                            double dn2 = dot3(ray.direction, nrm);
                            ray.direction.x = ray.direction.x - 2.0 * dn2 * nrm.x;
                            ray.direction.y = ray.direction.y - 2.0 * dn2 * nrm.y;
                            ray.direction.z = ray.direction.z - 2.0 * dn2 * nrm.z;
                            ray.direction = unit3(ray.direction);
                            ray.origin = at;
                        }
                    }
        }
    free(tile_miss);
    free(patch_cand);
    free(c.cull);
    free(c.scratch);
    free(c.order);
    free(c.cluster4);
    free(c.sub4);
    return c.bad;
}

int cert_check_num_stats(void) { return ST_COUNT; }
