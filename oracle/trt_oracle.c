/*
 * oracle/trt_oracle.c — CPU restatement of the reference render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (libtrt_b200.so) never links, loads or calls it.
 *
 * What it restates (TRT.c = /root/reference/TerminalRayTracer.c):
 *   orc_hit_sphere      ray_intersects_sphere   TRT.c:638-672
 *   orc_hit_plane       ray_intersects_plane    TRT.c:677-695
 *   orc_sky_texel       get_skybox_color        TRT.c:700-789
 *   orc_closest_hit     trace_ray               TRT.c:793-889
 *   orc_light_surface   apply_lighting          TRT.c:894-963
 *   orc_render_rows     project_scene           TRT.c:966-1069   (any row range of the same loop)
 *   orc_encode_stream   initialize_screenbuffer + buffered_draw_screen   TRT.c:1102-1172
 *   orc_subpixel_offsets  triangle_wave use at  TRT.c:225-228, 992-993
 *
 * Every arithmetic expression keeps the reference's operand order and grouping, in IEEE double,
 * unfused (build: -O3 -ffp-contract=off, no -march; see oracle/Makefile), so the output is meant
 * to be BIT-IDENTICAL to the reference's.  Parity status: PINNED — tests/test_oracle.py checks this
 * file function by function and frame by frame against oracle/_ref/libtrt_ref.so (the unmodified
 * reference TU compiled here) and against tests/golden/ (vectors written by the reference build,
 * generator: tests/golden/make_golden.py).  The reference ships no tests or golden vectors of its
 * own, so executing the reference is the only pin there is.
 *
 * One documented divergence from the letter of the reference (not from its observable behaviour):
 * the cubemap index can reach `dim` when u or v clamps to exactly +0.5 (TRT.c:778-788).  The
 * reference then reads one texel past the row, or up to dim+1 texels past the malloc'd plane
 * (glibc mmap slack: zeros).  Here, and in the CUDA path, every plane is REQUIRED to carry
 * dim+1 readable texels after its dim*dim payload; loaders zero-fill them ("black past the end").
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define TRT_NO_REFERENCE_NAMES
#include "trt_types.h"

typedef trt_Vector v3;

/* work counters for the algorithmic flop model of SURVEY.md §8(d) */
typedef struct
{
    long long sphere_tests;    /* entries of ray_intersects_sphere            */
    long long sphere_disc_ok;  /* discriminant >= 0                            */
    long long sphere_t0_pos;   /* t0 > 0 (a hit point is produced)             */
    long long sphere_closest;  /* hit became the new closest                   */
    long long plane_tests;     /* entries of ray_intersects_plane              */
    long long plane_denom_ok;  /* |denom| > 1e-5                               */
    long long plane_t_pos;     /* t > 1e-5                                     */
    long long plane_closest;   /* ground became closest                        */
    long long sky_lookups;     /* get_skybox_color calls                       */
    long long trace_calls;     /* trace_ray calls                              */
    long long trace_hits;      /* trace_ray calls that hit something           */
    long long lighting_calls;  /* apply_lighting calls                         */
    long long bounce_iters;    /* bodies of the bounce loop                    */
    long long samples;         /* (pixel, ray_num) pairs                       */
    long long pixels;
    long long bounce_hist[TRT_BOUNCE_LIMIT + 1];
} orc_counters;

#define COUNT(c, f) do { if (c) (c)->f++; } while (0)

/* ---- leaf math (TRT.c:439-546, 627-633): same grouping as the reference ------------------ */
static inline double dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          /* :463 */
static inline v3 sub3(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }      /* :501 */
static inline v3 add3(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }      /* :485 */
static inline v3 mul3(v3 a, v3 b) { v3 r = {a.x * b.x, a.y * b.y, a.z * b.z}; return r; }      /* :517 */
static inline v3 scale3(v3 a, double s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }      /* :469 */
static inline v3 unit3(v3 a)                                                                   /* :439-450 */
{
    double len = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 0.0001)
    {
        a.x /= len;
        a.y /= len;
        a.z /= len;
    }
    return a;
}
static inline double clampd(double v, double lo, double hi)                                   /* :523-530 */
{
    if (v < lo)
        return lo;
    if (v > hi)
        return hi;
    return v;
}

/* ---- TRT.c:638-672 --------------------------------------------------------------------- */
int orc_hit_sphere(const trt_Ray *ray, const trt_Sphere *s, trt_Point *hit, orc_counters *ctr)
{
    COUNT(ctr, sphere_tests);
    v3 o = {ray->origin.x, ray->origin.y, ray->origin.z};
    v3 c = {s->center.x, s->center.y, s->center.z};
    v3 oc = sub3(o, c);
    double a = dot3(ray->direction, ray->direction);
    double b = 2.0 * dot3(oc, ray->direction);
    double cc = dot3(oc, oc) - s->radius * s->radius;
    double disc = b * b - 4.0 * a * cc;
    if (disc < 0.0)
        return 0;
    COUNT(ctr, sphere_disc_ok);
    double t0 = (-b - sqrt(disc)) / (2.0 * a);
    if (t0 > 0.0)
    {
        COUNT(ctr, sphere_t0_pos);
        hit->x = ray->origin.x + t0 * ray->direction.x;
        hit->y = ray->origin.y + t0 * ray->direction.y;
        hit->z = ray->origin.z + t0 * ray->direction.z;
        return 1;
    }
    return 0;
}

/* ---- TRT.c:677-695 --------------------------------------------------------------------- */
int orc_hit_plane(const trt_Ray *ray, const trt_Plane *p, trt_Point *hit, orc_counters *ctr)
{
    COUNT(ctr, plane_tests);
    double denom = dot3(ray->direction, p->normal);
    if (fabs(denom) > 0.00001)
    {
        COUNT(ctr, plane_denom_ok);
        v3 pp = {p->point.x, p->point.y, p->point.z};
        v3 o = {ray->origin.x, ray->origin.y, ray->origin.z};
        v3 to_plane = sub3(pp, o);
        double t = dot3(to_plane, p->normal) / denom;
        if (t > 0.00001)
        {
            COUNT(ctr, plane_t_pos);
            hit->x = ray->origin.x + t * ray->direction.x;
            hit->y = ray->origin.y + t * ray->direction.y;
            hit->z = ray->origin.z + t * ray->direction.z;
            return 1;
        }
    }
    return 0;
}

/* ---- TRT.c:700-789 --------------------------------------------------------------------- */
static const v3 FACE_AXIS[6] = { /* TRT.c:137-143: +X,-X,+Y,-Y,+Z,-Z */
    {1.0, 0.0, 0.0}, {-1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, -1.0, 0.0}, {0.0, 0.0, 1.0}, {0.0, 0.0, -1.0}};

/* returns face and linear texel index (for unit tests) and the colour itself */
void orc_sky_texel(const trt_Skybox *sky, const trt_Vector *direction, trt_Color *color, int *face_out, long *index_out)
{
    v3 dir = unit3(*direction);
    int best = -1;
    double best_t = -1.0;
    for (int f = 0; f < 6; f++)
    {
        double t = dot3(dir, FACE_AXIS[f]);
        if (t > best_t)
        {
            best_t = t;
            best = f;
        }
    }
    v3 touching = mul3(dir, FACE_AXIS[best]);
    double scale_by = touching.x + touching.y + touching.z;
    dir = scale3(dir, 1.0 / scale_by);
    double t = dot3(dir, FACE_AXIS[best]);
    v3 along = scale3(FACE_AXIS[best], t);
    v3 across = sub3(dir, along);
    across = scale3(across, 0.5);
    double u = dot3(across, FACE_AXIS[(best + 2) % 6]);
    double v = dot3(across, FACE_AXIS[(best + 4) % 6]);
    if (best % 2 == 1)          /* :730 mirror */
        u *= -1.0;
    if (best == 0 || best == 1) /* :735 rotate -90 */
    {
        double tmp = u;
        u = v;
        v = -tmp;
    }
    else if (best == 2 || best == 3) /* :742-755 both branches are the same +90 rotation */
    {
        double tmp = u;
        u = -v;
        v = tmp;
    }
    else if (best == 4) /* :756 rotate 180 */
    {
        u *= -1.0;
        v *= -1.0;
    }
    u = clampd(u, -0.5, 0.5);
    v = clampd(v, -0.5, 0.5);
    int ui = (int)((u + 0.5) * sky->dim);
    int vi = (int)((v + 0.5) * sky->dim);
    long idx = (long)(ui + vi * sky->dim);
    *color = sky->colors[best][idx]; /* may touch the dim+1 pad texels, see header */
    if (face_out)
        *face_out = best;
    if (index_out)
        *index_out = idx;
}

/* ---- TRT.c:793-889 --------------------------------------------------------------------- */
trt_ObjectType orc_closest_hit(const trt_Scene *scene, const trt_Ray *ray, trt_Point *hit_out, trt_Vector *normal_out,
                               trt_Material *material_out, orc_counters *ctr)
{
    COUNT(ctr, trace_calls);
    double closest = INFINITY;
    trt_ObjectType what = TRT_NONE;
    trt_Point p;
    v3 best_p = {0, 0, 0}, best_n = {0, 0, 0};
    trt_Material best_m;
    memset(&best_m, 0, sizeof best_m);
    v3 o = {ray->origin.x, ray->origin.y, ray->origin.z};

    for (int i = 0; i < scene->num_spheres; i++)
    {
        if (orc_hit_sphere(ray, &scene->spheres[i], &p, ctr))
        {
            v3 pv = {p.x, p.y, p.z};
            v3 back = sub3(o, pv);
            double d2 = dot3(back, back);
            if (d2 < closest)
            {
                COUNT(ctr, sphere_closest);
                what = TRT_SPHERE;
                closest = d2;
                best_p = pv;
                v3 c = {scene->spheres[i].center.x, scene->spheres[i].center.y, scene->spheres[i].center.z};
                best_n = sub3(pv, c);
                best_m = scene->spheres[i].material;
            }
        }
    }
    if (orc_hit_plane(ray, &scene->ground, &p, ctr))
    {
        v3 pv = {p.x, p.y, p.z};
        v3 back = sub3(o, pv);
        double d2 = dot3(back, back);
        if (d2 < closest)
        {
            COUNT(ctr, plane_closest);
            what = TRT_GROUND;
            closest = d2;
            best_p = pv;
            best_n = scene->ground.normal;
            int odd = (int)(floor(p.x) + floor(p.z)) & 1; /* :850 */
            best_m = odd ? scene->ground.odd_material : scene->ground.even_material;
        }
    }
    if (what == TRT_NONE)
    {
        best_p = o;
        best_n = ray->direction;
        trt_Color texel;
        COUNT(ctr, sky_lookups);
        orc_sky_texel(&scene->skybox, &ray->direction, &texel, NULL, NULL);
        memset(&best_m, 0, sizeof best_m); /* compound literal zero-fills reflectivity/specularity, :866 */
        best_m.color.x = texel.r / 255.0;
        best_m.color.y = texel.g / 255.0;
        best_m.color.z = texel.b / 255.0;
    }
    else
    {
        COUNT(ctr, trace_hits);
        v3 back = unit3(sub3(o, best_p)); /* :871-874 */
        back = scale3(back, TRT_EPSILON);
        best_p = add3(best_p, back);
    }
    best_n = unit3(best_n); /* :878 */
    if (hit_out)
    {
        hit_out->x = best_p.x;
        hit_out->y = best_p.y;
        hit_out->z = best_p.z;
    }
    if (normal_out)
        *normal_out = best_n;
    if (material_out)
        *material_out = best_m;
    return what;
}

/* ---- TRT.c:894-963 (the unused `view` argument is dropped) -------------------------------- */
void orc_light_surface(const trt_Scene *scene, const trt_Point *at, const trt_Vector *normal, trt_Material *material,
                       orc_counters *ctr)
{
    COUNT(ctr, lighting_calls);
    v3 out = {0.0, 0.0, 0.0};
    v3 P = {at->x, at->y, at->z};
    for (int i = 0; i < scene->num_directional_lights; i++)
    {
        v3 L = unit3(scale3(scene->directional_lights[i].direction, -1.0));
        trt_Ray shadow = {*at, L};
        if (orc_closest_hit(scene, &shadow, NULL, NULL, NULL, ctr) == TRT_NONE)
        {
            v3 diffuse = scale3(scene->directional_lights[i].color, fmin(dot3(*normal, L), 1.0));
            diffuse = mul3(diffuse, material->color);
            out = add3(out, diffuse);
        }
    }
    for (int i = 0; i < scene->num_point_lights; i++)
    {
        v3 lp = {scene->point_lights[i].position.x, scene->point_lights[i].position.y, scene->point_lights[i].position.z};
        v3 L = sub3(lp, P);
        double light_d2 = dot3(L, L);
        double intensity = clampd(scene->point_lights[i].intensity / light_d2, 0.0, 1.0);
        L = unit3(L);
        trt_Ray shadow = {*at, L};
        trt_Point blocker;
        trt_ObjectType what = orc_closest_hit(scene, &shadow, &blocker, NULL, NULL, ctr);
        v3 bv = {blocker.x, blocker.y, blocker.z};
        v3 to_blocker = sub3(bv, P);
        double blocker_d2 = dot3(to_blocker, to_blocker);
        if (what == TRT_NONE || light_d2 < blocker_d2)
        {
            v3 diffuse = scale3(scene->point_lights[i].color, intensity * fmin(dot3(*normal, L), 1.0));
            diffuse = mul3(diffuse, material->color);
            out = add3(out, diffuse);
        }
    }
    out.x = clampd(out.x, 0.0, 1.0);
    out.y = clampd(out.y, 0.0, 1.0);
    out.z = clampd(out.z, 0.0, 1.0);
    material->color = out;
}

/* ---- TRT.c:225-228 and its two call sites :992-993 --------------------------------------- */
static double tri_wave(double t)
{
    return (fmod(t, 2 * TRT_PI) < TRT_PI) ? (fmod(t, 2 * TRT_PI) / TRT_PI) : (2 - (fmod(t, 2 * TRT_PI) / TRT_PI));
}
void orc_subpixel_offsets(double *dx, double *dy)
{
    for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++)
    {
        dx[k] = tri_wave(2 * TRT_PI * k / TRT_RAYS_PER_PIXEL) / 2;
        dy[k] = tri_wave(TRT_PI * k / TRT_RAYS_PER_PIXEL) / 2;
    }
}

/* ---- TRT.c:966-1069, rows [row0,row1) of the same loop nest --------------------------------- */
void orc_render_rows(const trt_Scene *scene, trt_Screen *screen, int row0, int row1, orc_counters *ctr)
{
    const trt_Camera *cam = &scene->camera;
    for (int row = row0; row < row1; row++)
    {
        for (int col = 0; col < screen->width; col++)
        {
            COUNT(ctr, pixels);
            v3 average = {0.0, 0.0, 0.0};
            for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++)
            {
                COUNT(ctr, samples);
                double pixel_w = cam->screen_width / screen->width;
                double pixel_h = cam->screen_height / screen->height;
                double sx = (((double)col / (double)screen->width) * cam->screen_width - cam->screen_width / 2.0);
                double sy = -(((double)row / (double)screen->height) * cam->screen_height - cam->screen_height / 2.0);
                double sz = -cam->screen_distance;
                sx += tri_wave(2 * TRT_PI * k / TRT_RAYS_PER_PIXEL) / 2 * pixel_w;
                sy += tri_wave(TRT_PI * k / TRT_RAYS_PER_PIXEL) / 2 * pixel_h;

                v3 wx = scale3(cam->frame.basis.x, sx);
                v3 wy = scale3(cam->frame.basis.y, sy);
                v3 wz = scale3(cam->frame.basis.z, sz);
                v3 dir = {0.0, 0.0, 0.0};
                dir = add3(dir, wx);
                dir = add3(dir, wy);
                dir = add3(dir, wz);
                v3 eye = {cam->frame.origin.x, cam->frame.origin.y, cam->frame.origin.z};
                dir = sub3(dir, eye); /* :1005 the origin-subtraction quirk */
                dir = unit3(dir);

                trt_Ray ray = {cam->frame.origin, dir};
                v3 sample = {0.0, 0.0, 0.0};
                int bounces = 0;
                double weight = 1.0, weight_sum = 0.0;
                int going = 1;
                while (going && bounces < TRT_BOUNCE_LIMIT && weight > 0.00001)
                {
                    COUNT(ctr, bounce_iters);
                    trt_Point at;
                    trt_Vector n;
                    trt_Material m;
                    trt_ObjectType what = orc_closest_hit(scene, &ray, &at, &n, &m, ctr);
                    if (what != TRT_NONE)
                        orc_light_surface(scene, &at, &n, &m, ctr);
                    weight_sum += weight;
                    m.color = scale3(m.color, weight);
                    if (what != TRT_NONE)
                    {
                        weight *= m.reflectivity;
                        bounces++;
                    }
                    else
                    {
                        weight = 0.0;
                        going = 0;
                    }
                    sample = add3(sample, m.color);
                    double dn = dot3(ray.direction, n); /* :627-633 */
                    ray.direction.x = ray.direction.x - 2.0 * dn * n.x;
                    ray.direction.y = ray.direction.y - 2.0 * dn * n.y;
                    ray.direction.z = ray.direction.z - 2.0 * dn * n.z;
                    ray.direction = unit3(ray.direction);
                    ray.origin = at;
                }
                if (ctr)
                    ctr->bounce_hist[bounces]++;
                sample = scale3(sample, 1.0 / weight_sum);
                average = add3(average, sample);
            }
            average = scale3(average, 1.0 / TRT_RAYS_PER_PIXEL);
            screen->pixels[row * screen->width + col] = average;
        }
    }
}

void orc_project_scene(const trt_Scene *scene, trt_Screen *screen)
{
    orc_render_rows(scene, screen, 0, screen->height, NULL);
}

/* ---- TRT.c:1102-1172: home + H x (W cells + '\n') + NUL + 2 more zero bytes ---------------- */
static void digits3(int value, char out[3]) /* :1134-1139 */
{
    out[0] = value / 100 + '0';
    out[1] = (value / 10) % 10 + '0';
    out[2] = value % 10 + '0';
}

size_t orc_stream_bytes(int w, int h) { return TRT_STREAM_BYTES(w, h); }

/* encode rows [row0,row1) into `out` (which receives exactly (row1-row0)*(25W+1) bytes) */
size_t orc_encode_rows(const trt_Screen *screen, int row0, int row1, char *out)
{
    static const char cell[] = "\033[48;2;000;000;000m  \033[0m"; /* :1103 */
    char *p = out;
    for (int i = row0; i < row1; i++)
    {
        for (int j = 0; j < screen->width; j++)
        {
            memcpy(p, cell, TRT_CELL_BYTES);
            trt_Vector px = screen->pixels[i * screen->width + j];
            digits3((int)(px.x * 255), p + 7);
            digits3((int)(px.y * 255), p + 11);
            digits3((int)(px.z * 255), p + 15);
            p += TRT_CELL_BYTES;
        }
        *p++ = '\n';
    }
    return (size_t)(p - out);
}

size_t orc_encode_stream(const trt_Screen *screen, char *out)
{
    memcpy(out, "\033[0;0H", TRT_HOME_BYTES); /* :1102, :1112 */
    size_t n = TRT_HOME_BYTES;
    n += orc_encode_rows(screen, 0, screen->height, out + n);
    memset(out + n, 0, TRT_TAIL_NULS); /* terminator + the two unused bytes of the static array, :1104, :1130, :1171 */
    return n + TRT_TAIL_NULS;
}

/* ---- algorithmic flop model, SURVEY.md §8(d) ----------------------------------------------- */
double orc_model_flops(const orc_counters *c)
{
    return 25.0 * c->sphere_tests + 5.0 * c->sphere_disc_ok + 14.0 * c->sphere_t0_pos + 3.0 * c->sphere_closest +
           5.0 * c->plane_tests + 9.0 * c->plane_denom_ok + 14.0 * c->plane_t_pos + 1.0 * c->plane_closest +
           90.0 * c->sky_lookups + 27.0 * c->trace_hits + 55.0 * c->lighting_calls + 34.0 * c->bounce_iters +
           68.0 * c->samples + 4.0 * c->pixels;
}

size_t orc_sizeof_counters(void) { return sizeof(orc_counters); }
