/*
 * trt_host.c — host-side (plain C) half of libtrt_b200: everything the reference does on the
 * CPU *around* the hot path and that a caller needs in order to feed it.  No CUDA here.
 *
 *   trt_init_camera        init_camera                       TRT.c:299-305 (aspect from w,h instead of macros)
 *   trt_orbit_camera       the per-frame pose recipe of main  TRT.c:1327-1336 (rotate_basis_x/_y :576-593,
 *                          rotate_basis :558-573, transform_frame :607-624)
 *   trt_demo_scene         the scene literals of main         TRT.c:1256-1306
 *   trt_stress_scene       SURVEY.md §8(d) config 3 (1024 random spheres; ranges after TRT.c:245-248)
 *   trt_read_ppm / trt_load_skybox / trt_free_skybox          TRT.c:309-436, same parsing quirks and
 *                          the same "printf to stdout + exit(1)" error behaviour
 *   trt_subpixel_offsets   triangle_wave at its call sites    TRT.c:225-228, 992-993
 *
 * Must be compiled with -ffp-contract=off and without -march flags: the camera basis it produces is
 * an input of the bit-exact render path, and the parity tests compare it bit for bit with the
 * reference's own functions.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "trt_b200.h"

/* ------------------------------------------------------------------------------------------ */
/* 3x3 helpers in the reference's row convention                                               */

static void frame_identity(trt_Frame *f) /* init_frame, TRT.c:290-296 */
{
    memset(f, 0, sizeof *f);
    f->basis.x.x = 1.0;
    f->basis.y.y = 1.0;
    f->basis.z.z = 1.0;
}

/* rows of `b` dotted with rows of `r` (B * R^T), TRT.c:558-573 */
static void basis_times_rows(trt_Basis *b, const trt_Basis *r)
{
    const trt_Vector *rows[3] = {&b->x, &b->y, &b->z};
    trt_Vector out[3];
    for (int i = 0; i < 3; i++)
    {
        const trt_Vector *v = rows[i];
        out[i].x = v->x * r->x.x + v->y * r->x.y + v->z * r->x.z;
        out[i].y = v->x * r->y.x + v->y * r->y.y + v->z * r->y.z;
        out[i].z = v->x * r->z.x + v->y * r->z.y + v->z * r->z.z;
    }
    b->x = out[0];
    b->y = out[1];
    b->z = out[2];
}

static void spin_about_x(trt_Basis *b, double angle) /* TRT.c:576-583 */
{
    trt_Basis r = {{1.0, 0.0, 0.0}, {0.0, cos(angle), -sin(angle)}, {0.0, sin(angle), cos(angle)}};
    basis_times_rows(b, &r);
}

static void spin_about_y(trt_Basis *b, double angle) /* TRT.c:586-593 */
{
    trt_Basis r = {{cos(angle), 0.0, sin(angle)}, {0.0, 1.0, 0.0}, {-sin(angle), 0.0, cos(angle)}};
    basis_times_rows(b, &r);
}

/* row-vector times homogeneous matrix, TRT.c:607-624 */
static void frame_apply(trt_Frame *f, const trt_Frame *tf)
{
    const trt_Vector *rows[3] = {&f->basis.x, &f->basis.y, &f->basis.z};
    trt_Vector out[3];
    const trt_Basis *m = &tf->basis;
    for (int i = 0; i < 3; i++)
    {
        const trt_Vector *v = rows[i];
        out[i].x = v->x * m->x.x + v->y * m->y.x + v->z * m->z.x;
        out[i].y = v->x * m->x.y + v->y * m->y.y + v->z * m->z.y;
        out[i].z = v->x * m->x.z + v->y * m->y.z + v->z * m->z.z;
    }
    trt_Point o;
    o.x = f->origin.x * m->x.x + f->origin.y * m->y.x + f->origin.z * m->z.x + tf->origin.x;
    o.y = f->origin.x * m->x.y + f->origin.y * m->y.y + f->origin.z * m->z.y + tf->origin.y;
    o.z = f->origin.x * m->x.z + f->origin.y * m->y.z + f->origin.z * m->z.z + tf->origin.z;
    f->basis.x = out[0];
    f->basis.y = out[1];
    f->basis.z = out[2];
    f->origin = o;
}

/* ------------------------------------------------------------------------------------------ */

void trt_init_camera(trt_Camera *camera, int width, int height)
{
    frame_identity(&camera->frame);
    camera->screen_distance = 1.0;
    camera->screen_width = 5 * (double)width / (double)height; /* TRT.c:303 with run-time w,h */
    camera->screen_height = 5 * 1.0;
}

void trt_orbit_camera(trt_Camera *camera, double t)
{
    trt_Frame spin, lift;
    frame_identity(&spin);
    frame_identity(&lift);
    frame_identity(&camera->frame);
    spin_about_x(&spin.basis, 2.0 * TRT_PI * t * -0.03); /* :1331 */
    spin_about_y(&spin.basis, 2.0 * TRT_PI * t * 0.05);  /* :1332 */
    lift.origin.x += 0.0;                                /* :1333-1334 root_to_camera = (0,0,1.99) */
    lift.origin.y += 0.0;
    lift.origin.z += 1.99;
    frame_apply(&camera->frame, &lift); /* :1335 */
    frame_apply(&camera->frame, &spin); /* :1336 */
}

/* The same recipe with the angles and the distance given instead of derived from the clock: pitch about x, then yaw about y (both in
 * radians), the camera `radius` away from the scene root — what a keyboard-driven camera needs (the reference README's open TODO
 * "camera controls from keyboard input"; host/trt_demo.c --keys).  trt_orbit_camera(c, t) == trt_pose_camera(c, 2 pi t * -0.03,
 * 2 pi t * 0.05, 1.99) bit for bit (tests/test_host.py). */
void trt_pose_camera(trt_Camera *camera, double pitch, double yaw, double radius)
{
    trt_Frame spin, lift;
    frame_identity(&spin);
    frame_identity(&lift);
    frame_identity(&camera->frame);
    spin_about_x(&spin.basis, pitch);
    spin_about_y(&spin.basis, yaw);
    lift.origin.x += 0.0;
    lift.origin.y += 0.0;
    lift.origin.z += radius;
    frame_apply(&camera->frame, &lift);
    frame_apply(&camera->frame, &spin);
}

void trt_subpixel_offsets(double dx[TRT_RAYS_PER_PIXEL], double dy[TRT_RAYS_PER_PIXEL])
{
    for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++)
    {
        double tx = 2 * TRT_PI * k / TRT_RAYS_PER_PIXEL; /* :992 */
        double ty = TRT_PI * k / TRT_RAYS_PER_PIXEL;     /* :993 */
        double wx = (fmod(tx, 2 * TRT_PI) < TRT_PI) ? (fmod(tx, 2 * TRT_PI) / TRT_PI) : (2 - (fmod(tx, 2 * TRT_PI) / TRT_PI));
        double wy = (fmod(ty, 2 * TRT_PI) < TRT_PI) ? (fmod(ty, 2 * TRT_PI) / TRT_PI) : (2 - (fmod(ty, 2 * TRT_PI) / TRT_PI));
        dx[k] = wx / 2;
        dy[k] = wy / 2;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* scene fixtures                                                                              */

static trt_Material mat(double r, double g, double b, double reflectivity)
{
    trt_Material m = {{r, g, b}, reflectivity, 100.0};
    return m;
}

static void ground_and_lights(trt_Scene *scene, trt_DirectionalLight *dl, trt_PointLight *pl)
{
    memset(&scene->ground, 0, sizeof scene->ground);
    scene->ground.normal.y = 1.0;                         /* TRT.c:1270 */
    scene->ground.point.y = -2.0;                         /* :1271 */
    scene->ground.even_material = mat(1.0, 1.0, 1.0, 0.2); /* :88, :1272 */
    scene->ground.odd_material = mat(1.0, 0.0, 0.0, 0.2);  /* :89, :1273 */
    memset(dl, 0, sizeof *dl);
    dl->direction.x = dl->direction.y = dl->direction.z = -1.0; /* :1279 */
    dl->color.x = dl->color.y = dl->color.z = 1.0;              /* :1280 */
    memset(pl, 0, sizeof *pl);
    pl->color.x = pl->color.y = pl->color.z = 1.0; /* :1284 (position = origin) */
    pl->intensity = 10.0;
    scene->directional_lights = dl;
    scene->num_directional_lights = 1;
    scene->point_lights = pl;
    scene->num_point_lights = 1;
}

void trt_demo_scene(trt_Scene *scene, trt_Sphere spheres[TRT_DEMO_SPHERES], trt_DirectionalLight *dl, trt_PointLight *pl,
                    int width, int height)
{
    /* TRT.c:1256-1263: one sphere per axis direction, radius 0.5 */
    static const double centre[TRT_DEMO_SPHERES][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
    static const double colour[TRT_DEMO_SPHERES][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
    static const double reflect[TRT_DEMO_SPHERES] = {1.0, 0.8, 0.8, 0.8, 0.8, 0.8};
    trt_Skybox keep = scene->skybox;
    memset(scene, 0, sizeof *scene);
    scene->skybox = keep;
    for (int i = 0; i < TRT_DEMO_SPHERES; i++)
    {
        spheres[i].center.x = centre[i][0];
        spheres[i].center.y = centre[i][1];
        spheres[i].center.z = centre[i][2];
        spheres[i].radius = 0.5;
        spheres[i].material = mat(colour[i][0], colour[i][1], colour[i][2], reflect[i]);
    }
    scene->spheres = spheres;
    scene->num_spheres = TRT_DEMO_SPHERES;
    ground_and_lights(scene, dl, pl);
    trt_init_camera(&scene->camera, width, height);
}

/* splitmix64 -> uniform double in [0,1) */
static uint64_t sm64(uint64_t *state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double sm64_unit(uint64_t *state) { return (double)(sm64(state) >> 11) * (1.0 / 9007199254740992.0); }
static double sm64_range(uint64_t *state, double lo, double hi) { return lo + sm64_unit(state) * (hi - lo); }

int trt_stress_scene(trt_Scene *scene, trt_Sphere *spheres, int count, trt_DirectionalLight *dl, trt_PointLight *pl,
                     int width, int height)
{
    static const double reflect[4] = {0.0, 0.2, 0.8, 1.0};
    uint64_t state = 0x5EED1024ull;
    trt_Skybox keep = scene->skybox;
    memset(scene, 0, sizeof *scene);
    scene->skybox = keep;
    int made = 0;
    while (made < count)
    {
        trt_Sphere s;
        memset(&s, 0, sizeof s);
        s.center.x = sm64_range(&state, -8.0, 8.0);
        s.center.y = sm64_range(&state, -1.5, 6.0);
        s.center.z = sm64_range(&state, -8.0, 8.0);
        s.radius = sm64_range(&state, 0.1, 0.5);                 /* TRT.c:245 */
        double r = sm64_range(&state, 0.0, 1.0);                 /* TRT.c:246-248 */
        double g = sm64_range(&state, 0.0, 1.0);
        double b = sm64_range(&state, 0.0, 1.0);
        /* keep the camera's orbit shell (radius 1.99) clear of geometry */
        double d = sqrt(s.center.x * s.center.x + s.center.y * s.center.y + s.center.z * s.center.z);
        if (d > 1.99 - s.radius - 0.05 && d < 1.99 + s.radius + 0.05)
            continue;
        s.material = mat(r, g, b, reflect[made & 3]);
        spheres[made++] = s;
    }
    scene->spheres = spheres;
    scene->num_spheres = count;
    ground_and_lights(scene, dl, pl);
    trt_init_camera(&scene->camera, width, height);
    return made;
}

/* ------------------------------------------------------------------------------------------ */
/* PPM / cubemap ingest (TRT.c:309-436).  Same accepted grammar as the reference:              */
/* "P6", one whitespace, any number of '#' comment lines, "W H", one whitespace, maxval, one   */
/* whitespace, then W*H RGB bytes.  maxval must be 255.  Errors: message on stdout, exit(1).   */

void trt_read_ppm(const char *filename, trt_Color **colors_ptr, int *width, int *height)
{
    FILE *fp = fopen(filename, "r");
    if (fp == NULL)
    {
        printf("Error opening file %s\n", filename);
        exit(1);
    }
    char magic[3] = {0, 0, 0};
    if (fgets(magic, 3, fp) == NULL || strncmp(magic, "P6", 2) != 0)
    {
        printf("Error: file is not ppm\n");
        fclose(fp);
        exit(1);
    }
    fgetc(fp);
    while (fgetc(fp) == '#')
        while (fgetc(fp) != '\n')
            ;
    fseek(fp, -1, SEEK_CUR);
    int maxval = 0;
    if (fscanf(fp, "%d %d", width, height) != 2)
        *width = *height = 0;
    fgetc(fp);
    if (fscanf(fp, "%d", &maxval) != 1)
        maxval = 0;
    fgetc(fp);
    if (maxval != 255)
    {
        printf("Error: max color value is not 255\n");
        fclose(fp);
        exit(1);
    }
    /* payload + (width+1) zeroed pad texels: the sampler's index can reach one row past the
     * end when u or v clamps to +0.5 (TRT.c:778-788); the reference reads allocator slack there. */
    size_t n = (size_t)(*width) * (size_t)(*height);
    trt_Color *colors = (trt_Color *)calloc(n + (size_t)(*width) + 1, sizeof(trt_Color));
    if (colors == NULL)
    {
        printf("Error allocating memory for colors\n");
        fclose(fp);
        exit(1);
    }
    for (size_t i = 0; i < n; i++)
    {
        colors[i].r = (unsigned char)fgetc(fp);
        colors[i].g = (unsigned char)fgetc(fp);
        colors[i].b = (unsigned char)fgetc(fp);
    }
    fclose(fp);
    *colors_ptr = colors;
}

void trt_load_skybox_dir(trt_Skybox *skybox, const char *dir)
{
    static const char *face_file[6] = {"+X.ppm", "-X.ppm", "+Y.ppm", "-Y.ppm", "+Z.ppm", "-Z.ppm"}; /* TRT.c:390 */
    char *path = (char *)malloc(strlen(dir) + 16);
    int dim = -1;
    for (int i = 0; i < 6; i++)
    {
        sprintf(path, "%s/%s", dir, face_file[i]);
        trt_Color *colors;
        int w, h;
        trt_read_ppm(path, &colors, &w, &h);
        if (dim == -1)
            dim = w;
        if (dim != w || dim != h)
        {
            printf("Error: all faces of the skybox must be the same size\n");
            exit(1);
        }
        skybox->colors[i] = colors;
    }
    skybox->dim = dim;
    free(path);
}

void trt_load_skybox(trt_Skybox *skybox, const char *skybox_name)
{
    /* the reference resolves "skybox/<name>/" relative to the working directory, TRT.c:403 */
    char *dir = (char *)malloc(strlen(skybox_name) + 16);
    sprintf(dir, "skybox/%s", skybox_name);
    trt_load_skybox_dir(skybox, dir);
    free(dir);
}

void trt_free_skybox(trt_Skybox *skybox)
{
    if (skybox->dim < 0)
        return;
    for (int i = 0; i < 6; i++)
    {
        free(skybox->colors[i]);
        skybox->colors[i] = NULL;
    }
    skybox->dim = -1;
}
