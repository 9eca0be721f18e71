/*
 * host/trt_demo.c — the reference's frame loop (TRT.c:1235-1370) as a plain-C host program with the two hot
 * calls redirected to libtrt_b200 through its C ABI.  Everything else (scene literals, orbit camera, frame
 * pacing, fps footer, SIGINT polling) stays host C, as in the reference.
 *
 *   build:  gcc -O2 -Iinclude host/trt_demo.c -Lterminalraytracer_b200 -ltrt_b200 -Wl,-rpath,'$ORIGIN/../terminalraytracer_b200' -lm -o host/trt_demo
 *   run  :  host/trt_demo [skybox-name [width height [frames]]]        (cwd must contain skybox/<name>/, TRT.c:403)
 *           host/trt_demo --orbit N [skybox-name [width height]]       N frames of one full camera turn, streamed to
 *                                                                      stdout through trt_render_orbit (BASELINE config 4)
 *           host/trt_demo --keys [skybox-name [width height]]          camera from the keyboard (the reference README's open
 *                                                                      TODO): a/d yaw, w/s pitch, +/- distance, q quits
 *
 * `TRT.c` = /root/reference/TerminalRayTracer.c
 */
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <termios.h>
#include <time.h>
#include <unistd.h>

#include "trt_b200.h"

#define FRAME_RATE 60                                              /* TRT.c:50 */
#define FRAME_DURATION_NS (long long)((1.0 / FRAME_RATE) * 1000000000) /* TRT.c:51-52 */

static volatile sig_atomic_t sigint_received = 0; /* TRT.c:1224 */
static void sigint_handler(int sig) { (void)sig; sigint_received = 1; }

/* sink of trt_render_orbit: the reference's single fwrite per frame (TRT.c:1171) */
static int write_frame(const char *bytes, size_t n, int frame, void *user)
{
    (void)frame;
    (void)user;
    fwrite(bytes, sizeof(char), n, stdout);
    return sigint_received;
}

static int orbit_mode(int argc, char **argv)
{
    const int n_frames = atoi(argv[2]);
    const char *skybox_name = argc > 3 ? argv[3] : "milky_way";
    const int width = argc > 5 ? atoi(argv[4]) : 1920, height = argc > 5 ? atoi(argv[5]) : 1080;
    if (n_frames <= 0) return 1;
    trt_init(0);
    Skybox skybox;
    trt_load_skybox(&skybox, (char *)skybox_name);
    trt_upload_skybox(&skybox);
    signal(SIGINT, sigint_handler);
    Sphere spheres[TRT_DEMO_SPHERES];
    DirectionalLight directional_light;
    PointLight point_light;
    Scene scene;
    scene.skybox = skybox;
    trt_demo_scene(&scene, spheres, &directional_light, &point_light, width, height);
    double *times = (double *)malloc(sizeof(double) * (size_t)n_frames);
    for (int k = 0; k < n_frames; k++) times[k] = k * (20.0 / n_frames);   /* one yaw turn at 0.05 rev/s, TRT.c:1332 */
    struct timespec t0, t1;
    timespec_get(&t0, TIME_UTC);
    const int done = trt_render_orbit(&scene, width, height, times, n_frames, 0, 1, write_frame, NULL);
    timespec_get(&t1, TIME_UTC);
    fprintf(stderr, "%d frames of %dx%d in %.3f s\n", done, width, height, (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
    free(times);
    trt_free_skybox(&skybox);
    trt_shutdown();
    return 0;
}

/* camera controls from keyboard input: the terminal in non-canonical, non-blocking mode; every key moves the pose that
 * trt_pose_camera turns into the camera frame (the reference's own recipe, TRT.c:1327-1336, with the angles from the keys) */
static int keys_mode(int argc, char **argv)
{
    const char *skybox_name = argc > 2 ? argv[2] : "milky_way";
    const int width = argc > 4 ? atoi(argv[3]) : TRT_DEFAULT_WIDTH, height = argc > 4 ? atoi(argv[4]) : TRT_DEFAULT_HEIGHT;
    struct termios saved, raw;
    const int tty = isatty(STDIN_FILENO);
    if (tty) {
        tcgetattr(STDIN_FILENO, &saved);
        raw = saved;
        raw.c_lflag &= (tcflag_t) ~(ICANON | ECHO);
        raw.c_cc[VMIN] = 0;
        raw.c_cc[VTIME] = 0;
        tcsetattr(STDIN_FILENO, TCSANOW, &raw);
    }
    trt_init(0);
    Skybox skybox;
    trt_load_skybox(&skybox, (char *)skybox_name);
    trt_upload_skybox(&skybox);
    signal(SIGINT, sigint_handler);
    Sphere spheres[TRT_DEMO_SPHERES];
    DirectionalLight directional_light;
    PointLight point_light;
    Scene scene;
    scene.skybox = skybox;
    trt_demo_scene(&scene, spheres, &directional_light, &point_light, width, height);
    char *stream = (char *)trt_host_alloc_pinned(TRT_STREAM_BYTES(width, height));
    double pitch = -0.3, yaw = 0.6, radius = 1.99;
    const double step = 0.05;
    int quit = 0;
    while (!quit && !sigint_received) {
        char key;
        int got = 0;
        while (read(STDIN_FILENO, &key, 1) == 1) {
            got = 1;
            if (key == 'a') yaw -= step;
            else if (key == 'd') yaw += step;
            else if (key == 'w') pitch -= step;
            else if (key == 's') pitch += step;
            else if (key == '+' || key == '=') radius = radius > 0.3 ? radius - 0.1 : radius;
            else if (key == '-') radius += 0.1;
            else if (key == 'q') quit = 1;
        }
        if (!tty && !got) quit = 1;                                /* piped input: leave when it is used up */
        trt_pose_camera(&scene.camera, pitch, yaw, radius);
        size_t n = trt_render_ansi(&scene, width, height, stream, TRT_STREAM_BYTES(width, height));
        fwrite(stream, sizeof(char), n, stdout);
        fflush(stdout);
        const struct timespec delay = {.tv_sec = 0, .tv_nsec = FRAME_DURATION_NS};
        nanosleep(&delay, NULL);
    }
    if (tty) tcsetattr(STDIN_FILENO, TCSANOW, &saved);
    trt_host_free_pinned(stream);
    trt_free_skybox(&skybox);
    trt_shutdown();
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 2 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 'o') return orbit_mode(argc, argv);
    if (argc > 1 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 'k') return keys_mode(argc, argv);
    const char *skybox_name = argc > 1 ? argv[1] : "milky_way"; /* TRT.c:1244 */
    const int width = argc > 3 ? atoi(argv[2]) : TRT_DEFAULT_WIDTH;
    const int height = argc > 3 ? atoi(argv[3]) : TRT_DEFAULT_HEIGHT;
    const long max_frames = argc > 4 ? atol(argv[4]) : -1;

    trt_init(0);                                   /* replaces initialize_screenbuffer(), TRT.c:1241 */
    Skybox skybox;
    trt_load_skybox(&skybox, (char *)skybox_name); /* load_skybox(&global_skybox, ...), TRT.c:1244 */
    trt_upload_skybox(&skybox);
    signal(SIGINT, sigint_handler);                /* TRT.c:1247 */

    struct timespec start, ts;
    timespec_get(&start, TIME_UTC);                /* TRT.c:1250-1251 */

    Sphere spheres[TRT_DEMO_SPHERES];
    DirectionalLight directional_light;
    PointLight point_light;
    Scene scene;
    scene.skybox = skybox;
    trt_demo_scene(&scene, spheres, &directional_light, &point_light, width, height); /* TRT.c:1256-1306 */

    char *stream = (char *)trt_host_alloc_pinned(TRT_STREAM_BYTES(width, height));

    for (long frame = 0; !sigint_received && (max_frames < 0 || frame < max_frames); frame++) /* TRT.c:1317 */
    {
        timespec_get(&ts, TIME_UTC);
        long long start_nanos = (ts.tv_sec - start.tv_sec) * 1000000000LL + ts.tv_nsec - start.tv_nsec;
        double t = (double)start_nanos / 1000000000.0;             /* TRT.c:1324 */

        trt_orbit_camera(&scene.camera, t);                        /* TRT.c:1327-1336 */

        /* project_scene(&scene,&screen); buffered_draw_screen(&screen);   TRT.c:1339, 1342 */
        size_t n = trt_render_ansi(&scene, width, height, stream, TRT_STREAM_BYTES(width, height));
        fwrite(stream, sizeof(char), n, stdout);                   /* TRT.c:1171 */

        timespec_get(&ts, TIME_UTC);                               /* TRT.c:1345-1355 */
        long long end_nanos = (ts.tv_sec - start.tv_sec) * 1000000000LL + ts.tv_nsec - start.tv_nsec;
        long long frame_time_nanos = end_nanos - start_nanos;
        if (FRAME_DURATION_NS > frame_time_nanos)
        {
            long long nanos_to_sleep = FRAME_DURATION_NS - frame_time_nanos;
            const struct timespec delay = {.tv_sec = nanos_to_sleep / 1000000000, .tv_nsec = nanos_to_sleep % 1000000000};
            nanosleep(&delay, NULL);
        }
        timespec_get(&ts, TIME_UTC);                               /* TRT.c:1358-1365 */
        long long end_nanos_2 = (ts.tv_sec - start.tv_sec) * 1000000000LL + ts.tv_nsec - start.tv_nsec;
        fputs("\033[0;0H", stdout);
        printf("%.02f fps\n", 1.0 / ((double)(end_nanos_2 - start_nanos) / 1000000000.0));
        fputs("\033[0;0H", stdout);
    }

    trt_host_free_pinned(stream);
    trt_free_skybox(&skybox);                                      /* TRT.c:1369 */
    trt_shutdown();
    return 0;
}
