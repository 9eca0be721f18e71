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
