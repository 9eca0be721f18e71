/*
 * oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Builds the UNMODIFIED reference translation unit, where it lies under
 * /root/reference, into a shared object so that tests and the CPU-baseline leg of
 * bench.py can call the reference's own functions:
 *     project_scene (TRT.c:966)      trace_ray (TRT.c:793)     apply_lighting (TRT.c:894)
 *     ray_intersects_sphere (638)    ray_intersects_plane (677) get_skybox_color (700)
 *     byte_to_digits (1134)          initialize_screenbuffer (1107)  buffered_draw_screen (1142)
 *     read_ppm (309) load_skybox (388) init_camera (299) rotate_basis_* (576-603) transform_frame (607)
 * All of those are external-linkage symbols of the reference TU and are therefore
 * exported by the .so as they are.  The reference's `main` is renamed by -Dmain=ref_main
 * on the command line (see oracle/Makefile); nothing else is changed.
 *
 * The few helpers below only *call* reference functions; the single piece of
 * restated logic is the camera-orbit call sequence of TRT.c:1327-1336, which lives
 * inside the reference's main() and cannot be called.
 *
 * Build recipe: oracle/Makefile (outputs only into oracle/_ref/, which is git-ignored).
 */
#ifndef TRT_HARNESS_TU_ALREADY_PRESENT
#include TRT_REFERENCE_TU /* "/root/reference/TerminalRayTracer.c", passed by the Makefile */
#endif

#include <unistd.h>
#include <fcntl.h>
#include <stddef.h>

/* ---- ABI probes: tests compare these with include/trt_types.h ------------------- */
size_t ref_sizeof_scene(void) { return sizeof(Scene); }
size_t ref_sizeof_sphere(void) { return sizeof(Sphere); }
size_t ref_sizeof_plane(void) { return sizeof(Plane); }
size_t ref_sizeof_camera(void) { return sizeof(Camera); }
size_t ref_sizeof_skybox(void) { return sizeof(Skybox); }
size_t ref_sizeof_screen(void) { return sizeof(Screen); }
size_t ref_offsetof_scene_camera(void) { return offsetof(Scene, camera); }
size_t ref_offsetof_scene_skybox(void) { return offsetof(Scene, skybox); }
size_t ref_offsetof_scene_ground(void) { return offsetof(Scene, ground); }
size_t ref_offsetof_scene_point_lights(void) { return offsetof(Scene, point_lights); }
int ref_screen_width(void) { return SCREEN_WIDTH; }
int ref_screen_height(void) { return SCREEN_HEIGHT; }
size_t ref_screenbuffer_bytes(void) { return sizeof(screenbuffer); }
int ref_rays_per_pixel(void) { return RAYS_PER_PIXEL; }
int ref_bounce_limit(void) { return BOUNCE_LIMIT; }

/* ---- camera pose for time t: the call sequence of TRT.c:1327-1336 ---------------- */
void ref_orbit_camera(Camera *camera, double t)
{
    Frame tf0, tf1;
    init_frame(&tf0);
    init_frame(&tf1);
    init_frame(&(camera->frame));
    rotate_basis_x(&tf0.basis, 2.0 * PI * t * -0.03);
    rotate_basis_y(&tf0.basis, 2.0 * PI * t * 0.05);
    Vector root_to_camera = {.x = 0.0, .y = 0.0, .z = 1.99};
    add_vectors((Vector *)&tf1.origin, &root_to_camera);
    transform_frame(&camera->frame, &tf1);
    transform_frame(&camera->frame, &tf0);
}

/* ---- the demo scene of main(): literals of TRT.c:1256-1288, camera by the reference's own init_camera
 * (TRT.c:299-305) with the screen width of line 303 re-evaluated for a W x H screen.  The literals live inside
 * main() and cannot be called; this restates them so that the reference arm of bench.py builds its input without
 * loading the product library. `spheres` must hold 6 entries. */
void ref_demo_scene(Scene *scene, Sphere *spheres, DirectionalLight *dl, PointLight *pl, int width, int height)
{
    static const double centre[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
    static const double colour[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
    Skybox keep = scene->skybox;
    memset(scene, 0, sizeof *scene);
    scene->skybox = keep;
    for (int i = 0; i < 6; i++)
    {
        Sphere s = {.center = {.x = centre[i][0], .y = centre[i][1], .z = centre[i][2]},
                    .material = {.color = {.x = colour[i][0], .y = colour[i][1], .z = colour[i][2]}, .reflectivity = i == 0 ? 1.0 : 0.8, .specularity = 100.0},
                    .radius = 0.5};
        spheres[i] = s;
    }
    Plane ground = {
        .normal = {.x = 0.0, .y = 1.0, .z = 0.0},
        .point = {.x = 0.0, .y = -2.0, .z = 0.0},
        .even_material = {.color = GROUND_EVEN_COLOR, .reflectivity = 0.2, .specularity = 100.0},
        .odd_material = {.color = GROUND_ODD_COLOR, .reflectivity = 0.2, .specularity = 100.0},
    };
    DirectionalLight d = {.direction = {.x = -1.0, .y = -1.0, .z = -1.0}, .color = {.x = 1.0, .y = 1.0, .z = 1.0}};
    PointLight p = {.position = {.x = 0.0, .y = 0.0, .z = 0.0}, .color = {.x = 1.0, .y = 1.0, .z = 1.0}, .intensity = 10.0};
    *dl = d;
    *pl = p;
    scene->spheres = spheres;
    scene->num_spheres = 6;
    scene->ground = ground;
    scene->directional_lights = dl;
    scene->num_directional_lights = 1;
    scene->point_lights = pl;
    scene->num_point_lights = 1;
    init_camera(&scene->camera);
    scene->camera.screen_width = 5 * (double)width / (double)height; /* TRT.c:303 for a W x H screen */
}

/* ---- sub-pixel offsets exactly as TRT.c:992-993 evaluates them ------------------- */
void ref_subpixel_offsets(double *dx, double *dy)
{
    for (int ray_num = 0; ray_num < RAYS_PER_PIXEL; ray_num++)
    {
        dx[ray_num] = triangle_wave(2 * PI * ray_num / RAYS_PER_PIXEL) / 2;
        dy[ray_num] = triangle_wave(PI * ray_num / RAYS_PER_PIXEL) / 2;
    }
}

/* ---- capture what buffered_draw_screen fwrite()s to stdout ------------------------
 * Only valid for screens of the compiled SCREEN_WIDTH x SCREEN_HEIGHT (the reference's
 * buffer is a macro-sized static array, TRT.c:1104).  Returns the byte count. */
long ref_draw_screen_bytes(Screen *screen, char *out, long cap)
{
    if (screen->width != SCREEN_WIDTH || screen->height != SCREEN_HEIGHT)
        return -1;
    char path[] = "/tmp/trt_ref_stdout_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0)
        return -2;
    unlink(path);
    fflush(stdout);
    int saved = dup(1);
    dup2(fd, 1);
    initialize_screenbuffer();
    buffered_draw_screen(screen);
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    long n = (long)lseek(fd, 0, SEEK_END);
    lseek(fd, 0, SEEK_SET);
    long got = 0;
    while (got < n && got < cap)
    {
        long r = (long)read(fd, out + got, (size_t)((n < cap ? n : cap) - got));
        if (r <= 0)
            break;
        got += r;
    }
    close(fd);
    return n;
}

/* ---- wall-clock of the two hot calls, for bench.py's cpu_baseline leg -------------- */
double ref_time_project_scene(Scene *scene, Screen *screen)
{
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    project_scene(scene, screen);
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

#ifdef TRT_REF_ROW_RANGE
/* Row-range build (oracle/Makefile pipes the reference TU through a one-line sed that
 * turns the row loop of TRT.c:973 into `for (int row = ref_row0; row < ref_row1; ...`).
 * Used ONLY to spread golden generation over host cores; validated against the
 * unpatched build in tests/test_oracle.py. */
void ref_project_rows(Scene *scene, Screen *screen, int row0, int row1)
{
    ref_row0 = row0;
    ref_row1 = row1;
    project_scene(scene, screen);
}
#endif
