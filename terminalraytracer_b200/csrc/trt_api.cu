// trt_api.cu — the C ABI of libtrt_b200.so (include/trt_b200.h): device context, scene/skybox upload,
// the drop-in entry points and the device-resident band API.  Host C calls land here; nothing in this
// file computes pixels on the CPU — if CUDA is unusable the library exits (no CPU fallback by design,
// mirroring the reference's printf+exit(1) error style, TRT.c:318-322).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "trt_internal.h"

using namespace trt;

namespace {

void die_if(cudaError_t e, const char *file, int line)
{
    if (e != cudaSuccess) {
        fprintf(stderr, "%s:%d: CUDA error: %s\n", file, line, cudaGetErrorString(e));
        exit(1);
    }
}
#define CK(x) die_if((x), __FILE__, __LINE__)

struct Buffer {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuffer {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) CK(cudaFreeHost(p));
        CK(cudaMallocHost(&p, bytes));
        cap = bytes;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Context {
    bool ready = false;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device-to-host copies of finished row chunks (trt_render_ansi)
    static constexpr int MAX_CHUNKS = 8;
    cudaEvent_t chunk_ev[MAX_CHUNKS][3] = {};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_render_ms = 0.f, last_encode_ms = 0.f;
    // scene
    DevScene scene;
    bool have_scene = false;
    bool cull_allowed = true;   // trt_set_cull(); the FP32 miss test can be switched off for A/B runs
    bool cull = true;           // cull_allowed && this scene's magnitudes are inside the bound's range
    // every per-scene array (sphere geometry, certificate records, materials, k-d order, ball tree) lives in ONE device blob with
    // the layout of the staging arena, so that an upload is one host-to-device copy plus the constant block
    Buffer scene_blob;
    size_t off_geom = 0, off_cull = 0, off_mat = 0, off_pairs = 0, off_orig = 0, off_pos = 0, off_ball = 0, off_link = 0;   // off_ball / off_link: cluster balls / their sub-balls
    // skybox
    Buffer sky;
    int sky_dim = -1, sky_face_stride = 0;
    // work
    Buffer tile_counter, counters, byte_to_unit, scratch, tile_info;
    Buffer pixels, quant, bytes;
    PinnedBuffer stage;
    // the streaming sink's double buffers and events: kept across calls (page-locking and freeing 100 MB per call took longer
    // than rendering 360 frames)
    Buffer orbit_dev[2];
    PinnedBuffer orbit_host[2];
    cudaEvent_t orbit_encoded[2] = {nullptr, nullptr}, orbit_copied[2] = {nullptr, nullptr};
    // pinned staging for scene uploads: copies from pageable memory make the runtime wait for the stream first, which
    // would serialise the frames of trt_render_orbit; two arenas so that a frame's upload never overwrites the previous one's
    PinnedBuffer arena[2];
    cudaEvent_t arena_done[2] = {nullptr, nullptr};   // recorded behind the copies of the upload that last used the arena
    bool arena_busy[2] = {false, false};
    size_t arena_used = 0;
    int arena_which = 0;
} g;

// copy `bytes` from (pageable) src into the current staging arena; returns the pinned address (16-byte aligned)
const void *staged(const void *src, size_t bytes)
{
    const size_t at = (g.arena_used + 15) & ~(size_t)15;
    if (at + bytes > g.arena[g.arena_which].cap) {
        fprintf(stderr, "libtrt_b200: scene staging arena too small\n");
        exit(1);
    }
    memcpy((char *)g.arena[g.arena_which].p + at, src, bytes);
    g.arena_used = at + bytes;
    return (const char *)g.arena[g.arena_which].p + at;
}

// the same, returning the offset inside the arena (= the offset inside the device blob)
size_t staged_at(const void *src, size_t bytes)
{
    const char *p = (const char *)staged(src, bytes);
    return (size_t)(p - (const char *)g.arena[g.arena_which].p);
}

void require_init(const char *who)
{
    if (!g.ready) {
        fprintf(stderr, "libtrt_b200: %s called before trt_init()\n", who);
        exit(1);
    }
}

// unit(-direction) exactly as apply_lighting does it (TRT.c:903-904, 439-450); this TU's host code is
// compiled without FMA contraction (x86-64 baseline), so the bits equal the reference's.
void negated_unit(const trt_Vector &dir, double out[3])
{
    volatile double x = dir.x * -1.0, y = dir.y * -1.0, z = dir.z * -1.0;
    volatile double xx = x * x, yy = y * y, zz = z * z;
    volatile double s = xx + yy;
    s = s + zz;
    double len = sqrt(s);
    if (len > 0.0001) {
        out[0] = x / len;
        out[1] = y / len;
        out[2] = z / len;
    } else {
        out[0] = x;
        out[1] = y;
        out[2] = z;
    }
}

// smallest float >= v
float float_round_up(double v) { return trt_cert_round_up(v); }

// dot_product with the reference's grouping (TRT.c:461-464), every product and sum rounded once
double dot3_ref(const double a[3], const double b[3])
{
    volatile double xx = a[0] * b[0], yy = a[1] * b[1], zz = a[2] * b[2];
    volatile double s = xx + yy;
    s = s + zz;
    return s;
}

// normalize_vector, TRT.c:439-450
void unit3_ref(const double v[3], double out[3])
{
    const double len = sqrt(dot3_ref(v, v));
    for (int k = 0; k < 3; k++) {
        volatile double q = len > 0.0001 ? v[k] / len : v[k];
        out[k] = q;
    }
}

void set_material(DevMaterial &m, const trt_Material &src)
{
    m.color[0] = src.color.x;
    m.color[1] = src.color.y;
    m.color[2] = src.color.z;
    m.reflectivity = src.reflectivity;
}

void upload_scene(const trt_Scene *scene, bool wait = true)
{
    if (scene->num_directional_lights > TRT_MAX_LIGHTS || scene->num_point_lights > TRT_MAX_LIGHTS ||
        scene->num_directional_lights < 0 || scene->num_point_lights < 0 || scene->num_spheres < 0) {
        fprintf(stderr, "libtrt_b200: unsupported scene (at most %d lights of each kind)\n", TRT_MAX_LIGHTS);
        exit(1);
    }
    DevScene &s = g.scene;
    {
        const size_t per_sphere = sizeof(double4) + sizeof(DevMaterial) + sizeof(float4) * 2 + sizeof(CullPair) + sizeof(int) * 2 + 32;
        const size_t need = sizeof(DevScene) + 4096 + per_sphere * ((size_t)scene->num_spheres + 64);
        g.arena_which = wait ? 0 : (g.arena_which ^ 1);
        // an arena may be rewritten only after the copies of the upload that last used it have run
        if (g.arena_busy[g.arena_which]) {
            CK(cudaEventSynchronize(g.arena_done[g.arena_which]));
            g.arena_busy[g.arena_which] = false;
        }
        if (g.arena[g.arena_which].cap < need) {
            CK(cudaStreamSynchronize(g.stream));       // nothing may still be reading the arena that is replaced
            g.arena[g.arena_which].reserve(need);
        }
        g.arena_used = 0;
    }
    bool ground_in_range = true;
    const trt_Camera &c = scene->camera;
    s.bx[0] = c.frame.basis.x.x; s.bx[1] = c.frame.basis.x.y; s.bx[2] = c.frame.basis.x.z;
    s.by[0] = c.frame.basis.y.x; s.by[1] = c.frame.basis.y.y; s.by[2] = c.frame.basis.y.z;
    s.bz[0] = c.frame.basis.z.x; s.bz[1] = c.frame.basis.z.y; s.bz[2] = c.frame.basis.z.z;
    s.eye[0] = c.frame.origin.x; s.eye[1] = c.frame.origin.y; s.eye[2] = c.frame.origin.z;
    s.screen_distance = c.screen_distance;
    s.screen_width = c.screen_width;
    s.screen_height = c.screen_height;
    s.ground_point[0] = scene->ground.point.x; s.ground_point[1] = scene->ground.point.y; s.ground_point[2] = scene->ground.point.z;
    s.ground_normal[0] = scene->ground.normal.x; s.ground_normal[1] = scene->ground.normal.y; s.ground_normal[2] = scene->ground.normal.z;
    for (int k = 0; k < 3; k++) {
        s.ground_point_f[k] = (float)s.ground_point[k];
        s.ground_normal_f[k] = (float)s.ground_normal[k];
    }
    if (!(fabs(s.ground_point[0]) + fabs(s.ground_point[1]) + fabs(s.ground_point[2]) < 1e12)) ground_in_range = false;
    unit3_ref(s.ground_normal, s.ground_unit_normal);
    {
        // rays leaving the eye: numerator of TRT.c:685 and its robust sign
        double to_plane[3];
        for (int k = 0; k < 3; k++) {
            volatile double d = s.ground_point[k] - s.eye[k];
            to_plane[k] = d;
        }
        s.prim_num = dot3_ref(to_plane, s.ground_normal);
        double scale = 0.0, nl1 = 0.0;
        for (int k = 0; k < 3; k++) {
            scale += fabs(s.ground_point[k]) + fabs(s.eye[k]);
            nl1 += fabs(s.ground_normal[k]);
        }
        const double tol = 1e-9 * scale * nl1;
        s.prim_num_sign = s.prim_num < -tol ? -1 : (s.prim_num > tol ? 1 : 0);
        s.ground_normal_l1 = float_round_up(nl1);
        s.prim_num_f = (float)s.prim_num;
        for (int k = 0; k < 3; k++) s.ground_unit_normal_f[k] = (float)s.ground_unit_normal[k];
        // camera in float for the tile certificates (trt_cert.h)
        trt_cert_camera &cf = s.cam_f;
        cf.ex = (float)s.eye[0]; cf.ey = (float)s.eye[1]; cf.ez = (float)s.eye[2];
        for (int k = 0; k < 3; k++) {
            cf.bx[k] = (float)s.bx[k];
            cf.by[k] = (float)s.by[k];
            cf.bz[k] = (float)s.bz[k];
        }
        cf.nbx = float_round_up(sqrt(dot3_ref(s.bx, s.bx)) * (1.0 + 1e-6));
        cf.nby = float_round_up(sqrt(dot3_ref(s.by, s.by)) * (1.0 + 1e-6));
        cf.sw = (float)s.screen_width;
        cf.sh = (float)s.screen_height;
        cf.dist = (float)s.screen_distance;
        cf.pw = cf.ph = 0.f;    // per launch: RenderParams::pixel_w_f / pixel_h_f
        s.eye_l1 = float_round_up((fabs(s.eye[0]) + fabs(s.eye[1]) + fabs(s.eye[2])) * (1.0 + 1e-6));
        if (!(fabs(s.eye[0]) + fabs(s.eye[1]) + fabs(s.eye[2]) < 1e12) || !(fabs(s.screen_width) + fabs(s.screen_height) + fabs(s.screen_distance) < 1e12))
            ground_in_range = false;
    }
    set_material(s.ground_even, scene->ground.even_material);
    set_material(s.ground_odd, scene->ground.odd_material);
    s.num_dir = scene->num_directional_lights;
    s.num_point = scene->num_point_lights;
    for (int i = 0; i < s.num_dir; i++) {
        negated_unit(scene->directional_lights[i].direction, s.dir[i].L);
        s.dir[i].color[0] = scene->directional_lights[i].color.x;
        s.dir[i].color[1] = scene->directional_lights[i].color.y;
        s.dir[i].color[2] = scene->directional_lights[i].color.z;
        s.dir[i].plane_denom = dot3_ref(s.dir[i].L, s.ground_normal);      // TRT.c:681 for this light's shadow rays
        s.dir[i].plane_possible = fabs(s.dir[i].plane_denom) > 0.00001 ? 1 : 0;
        for (int k = 0; k < 3; k++) s.dir[i].Lf[k] = (float)s.dir[i].L[k];
        {
            const float *f = s.dir[i].Lf;
            const float dd = fmaf(f[2], f[2], fmaf(f[1], f[1], f[0] * f[0]));     // the float expression the certificates' bound assumes
            s.dir[i].lf_unit = (dd > 0.99999f && dd < 1.00001f) ? 1 : 0;
        }
    }
    double light_l1 = 0.0;
    for (int i = 0; i < s.num_point; i++) {
        const trt_PointLight &p = scene->point_lights[i];
        s.point[i].pos[0] = p.position.x; s.point[i].pos[1] = p.position.y; s.point[i].pos[2] = p.position.z;
        s.point[i].color[0] = p.color.x; s.point[i].color[1] = p.color.y; s.point[i].color[2] = p.color.z;
        s.point[i].intensity = p.intensity;
        double rel[3], l1 = 0.0;
        for (int k = 0; k < 3; k++) {
            rel[k] = s.point[i].pos[k] - s.ground_point[k];
            s.point[i].pos_f[k] = (float)s.point[i].pos[k];
            l1 += fabs(s.point[i].pos[k]);
        }
        s.point[i].height = dot3_ref(rel, s.ground_normal);
        s.point[i].pos_l1 = float_round_up(l1 * (1.0 + 1e-6));
        if (!(l1 < 1e12)) ground_in_range = false;
        if (l1 > light_l1) light_l1 = l1;
    }
    {
        double gl1 = 0.0;
        for (int k = 0; k < 3; k++) gl1 += fabs(s.ground_point[k]);
        s.ground_margin = sqrt(dot3_ref(s.ground_normal, s.ground_normal)) * (1e-4 + 1e-9 * (gl1 + light_l1));
    }
    {
        double mx = 0.0, my = 0.0, dx[TRT_RAYS_PER_PIXEL], dy[TRT_RAYS_PER_PIXEL];
        trt_subpixel_offsets(dx, dy);
        for (int k = 0; k < TRT_RAYS_PER_PIXEL; k++) {
            if (dx[k] > mx) mx = dx[k];
            if (dy[k] > my) my = dy[k];
        }
        s.cam_f.off_x = float_round_up(mx);
        s.cam_f.off_y = float_round_up(my);
    }
    s.num_spheres = scene->num_spheres;
    s.sphere_mask = scene->num_spheres >= 32 ? 0xffffffffu : ((1u << scene->num_spheres) - 1u);
    s.sky_dim = g.sky_dim;
    s.sky_face_stride = g.sky_face_stride;
    trt_subpixel_offsets(s.sub_dx, s.sub_dy);

    const int n = s.num_spheres;
    std::vector<double4> geom((size_t)(n > 0 ? n : 1));
    std::vector<DevMaterial> mats((size_t)(n > 0 ? n : 1));
    std::vector<float4> cull((size_t)n + 2, make_float4(0.f, 0.f, 0.f, 0.f));   // padded to an even count (+1 spare)
    bool in_range = true;       // magnitudes for which the FP32 cull's error bound was derived
    double centre_l1 = 0.0;
    for (int i = 0; i < n; i++) {
        const trt_Sphere &sp = scene->spheres[i];
        volatile double r2 = sp.radius * sp.radius;   // TRT.c:648, a single rounded product
        geom[i] = make_double4(sp.center.x, sp.center.y, sp.center.z, r2);
        set_material(mats[i], sp.material);
        // FP32 cull record: centre rounded to nearest, radius padded and rounded UP (trt_render.cu, sphere_cull)
        const double l1 = fabs(sp.center.x) + fabs(sp.center.y) + fabs(sp.center.z);
        const double r = sqrt((double)r2);
        if (!(l1 < 1e12) || !(r < 1e12) || (r != 0.0 && !(r > 1e-12))) in_range = false;
        if (l1 > centre_l1) centre_l1 = l1;
        cull[i] = make_float4((float)sp.center.x, (float)sp.center.y, (float)sp.center.z, trt_cert_pad_radius(sp.radius));
        (void)r;
    }
    // Many-sphere scenes: the device sees the spheres in k-d order, 32 consecutive ones share a bounding ball
    // (trt_cert_cluster_miss).  `orig` keeps the reference's index of every sorted sphere for its tie-breaking rule
    // (TRT.c:810, strict <: the lowest index wins), `pos` is the inverse (the all-FP64 query scans in the reference's order).
    std::vector<int> orig((size_t)(n > 0 ? n : 1)), pos((size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) orig[i] = pos[i] = i;
    const int num_clusters = (n + 31) / 32;
    std::vector<float4> clusters((size_t)(num_clusters > 0 ? num_clusters : 1) + 1, make_float4(0.f, 0.f, 0.f, 0.f));   // (+1: read in pairs)
    std::vector<CullPair> subballs((size_t)(num_clusters > 0 ? 2 * num_clusters : 2));
    memset(subballs.data(), 0, sizeof(CullPair) * subballs.size());
    s.clustered = n > TRT_CLUSTER_MIN_SPHERES ? 1 : 0;
    if (s.clustered) {
        trt_cert_kd_order(reinterpret_cast<const float *>(cull.data()), n, orig.data());
        std::vector<double4> geom2(geom);
        std::vector<DevMaterial> mats2(mats);
        std::vector<float4> cull2(cull);
        for (int j = 0; j < n; j++) {
            const int i = orig[j];
            pos[i] = j;
            geom[j] = geom2[i];
            mats[j] = mats2[i];
            cull[j] = cull2[i];
        }
        for (int c = 0; c < num_clusters; c++) {
            float b[4];
            const int count = n - 32 * c < 32 ? n - 32 * c : 32;
            trt_cert_cluster_bound(reinterpret_cast<const float *>(cull.data() + 32 * c), count, b);
            clusters[c] = make_float4(b[0], b[1], b[2], b[3]);
            float q[4][4];
            for (int k = 0; k < 4; k++) {
                const int j0 = 32 * c + 8 * k, cnt = n - j0 < 8 ? n - j0 : 8;
                q[k][0] = q[k][1] = q[k][2] = q[k][3] = 0.f;
                if (cnt > 0) trt_cert_cluster_bound(reinterpret_cast<const float *>(cull.data() + j0), cnt, q[k]);
            }
            for (int half = 0; half < 2; half++) {
                CullPair &sp = subballs[(size_t)(2 * c + half)];
                sp.cx = make_float2(q[2 * half][0], q[2 * half + 1][0]);
                sp.cy = make_float2(q[2 * half][1], q[2 * half + 1][1]);
                sp.cz = make_float2(q[2 * half][2], q[2 * half + 1][2]);
                sp.r = make_float2(q[2 * half][3], q[2 * half + 1][3]);
            }
        }
    }
    g.off_ball = staged_at(clusters.data(), sizeof(float4) * clusters.size());
    g.off_link = staged_at(subballs.data(), sizeof(CullPair) * subballs.size());
    g.off_orig = staged_at(orig.data(), sizeof(int) * orig.size());
    g.off_pos = staged_at(pos.data(), sizeof(int) * pos.size());
    g.cull = g.cull_allowed && in_range && ground_in_range;
    s.filter_enabled = g.cull ? 1 : 0;
    s.filter_centre_l1 = float_round_up(centre_l1 * (1.0 + 1.0 / 1048576.0));
    g.off_cull = staged_at(cull.data(), sizeof(float4) * cull.size());
    g.off_geom = staged_at(geom.data(), sizeof(double4) * geom.size());
    g.off_mat = staged_at(mats.data(), sizeof(DevMaterial) * mats.size());
    // the same records two by two for the packed classification; the odd one out is paired with a sphere of radius 0
    std::vector<CullPair> pairs((size_t)(n / 2 + 1));
    for (size_t p = 0; p < pairs.size(); p++) {
        const float4 a = cull[2 * p], b = cull[2 * p + 1];   // cull has n + 2 entries, zero-filled past n
        pairs[p].cx = make_float2(a.x, b.x);
        pairs[p].cy = make_float2(a.y, b.y);
        pairs[p].cz = make_float2(a.z, b.z);
        pairs[p].r = make_float2(a.w, b.w);
    }
    g.off_pairs = staged_at(pairs.data(), sizeof(CullPair) * pairs.size());
    // Every source went through staged(): a memcpy into the page-locked arena g.arena[g.arena_which], so the vectors may die when
    // this function returns and the copy is truly asynchronous.  Invariant: an arena may be rewritten only after the copies of the
    // upload that last used it have run — arena_done[] is recorded behind them and waited for at the top of this function;
    // wait == false alternates the two arenas, so the host runs at most two uploads ahead of the device.  The device blob itself is
    // rewritten in stream order, behind the kernels of the previous frame.
    const size_t blob_bytes = g.arena_used;
    if (g.scene_blob.cap < blob_bytes) {
        CK(cudaStreamSynchronize(g.stream));       // nothing may still be reading the blob that is replaced
        g.scene_blob.reserve(blob_bytes + 4096);
    }
    CK(cudaMemcpyAsync(g.scene_blob.p, g.arena[g.arena_which].p, blob_bytes, cudaMemcpyHostToDevice, g.stream));
    upload_scene_constants(*(const DevScene *)staged(&s, sizeof s), g.stream);
    CK(cudaEventRecord(g.arena_done[g.arena_which], g.stream));
    g.arena_busy[g.arena_which] = true;
    // wait: callers that time expect a resident scene on return
    if (wait) {
        CK(cudaStreamSynchronize(g.stream));
        g.arena_busy[g.arena_which] = false;
    }
    g.have_scene = true;
}

// 0: all FP64; 1: small scene (one chunk), certificate records copied to shared memory at kernel start; 2: k-d-sorted scene with cluster balls, records in global memory
int cull_mode() { return !g.cull ? 0 : (g.scene.clustered ? 2 : 1); }

bool one_plus_one() { return g.scene.num_dir == 1 && g.scene.num_point == 1; }

RenderParams make_params(int width, int height, int row0, int row1, double *d_pixels, uchar4 *d_quant, bool count)
{
    if (!g.have_scene) {
        fprintf(stderr, "libtrt_b200: render requested before a scene was set\n");
        exit(1);
    }
    if (g.sky_dim <= 0) {
        fprintf(stderr, "libtrt_b200: render requested before trt_upload_skybox()\n");
        exit(1);
    }
    RenderParams p;
    p.width = width;
    p.height = height;
    p.row0 = row0;
    p.row1 = row1;
    {
        volatile double pw = width > 0 ? g.scene.screen_width / (double)width : 0.0;     // one IEEE division each, as TRT.c:981-982
        volatile double ph = height > 0 ? g.scene.screen_height / (double)height : 0.0;
        p.pixel_w = pw;
        p.pixel_h = ph;
    }
    p.pixel_w_f = width > 0 ? (float)(g.scene.screen_width / width) : 0.f;
    p.pixel_h_f = height > 0 ? (float)(g.scene.screen_height / height) : 0.f;
    p.pixels = d_pixels;
    p.quant = d_quant;
    p.ansi = nullptr;
    const char *blob = (const char *)g.scene_blob.p;
    p.sphere_geom = (const double4 *)(blob + g.off_geom);
    p.sphere_cull = (const float4 *)(blob + g.off_cull);
    p.cull_pairs = (const CullPair *)(blob + g.off_pairs);
    p.sphere_orig = (const int *)(blob + g.off_orig);
    p.sphere_pos = (const int *)(blob + g.off_pos);
    p.clusters = (const float4 *)(blob + g.off_ball);
    p.subballs = (const CullPair *)(blob + g.off_link);
    p.sphere_mat = (const DevMaterial *)(blob + g.off_mat);
    p.byte_to_unit = (const double *)g.byte_to_unit.p;
    p.sky = (const uchar4 *)g.sky.p;
    // sized for the whole frame, so that bands of any height (adaptive bands, row chunks) never reallocate
    g.tile_info.reserve(render_tile_info_bytes(width, height > row1 - row0 ? height : row1 - row0));
    p.tile_info = (const uint4 *)g.tile_info.p;
    p.tile_counter = (unsigned int *)g.tile_counter.p;
    g.scratch.reserve(render_scratch_bytes(g.num_sms));
    p.sample_scratch = (double *)g.scratch.p;
    p.scratch_bytes = g.scratch.cap;
    p.tile_info_bytes = g.tile_info.cap;
    p.counters = count ? (unsigned long long *)g.counters.p : nullptr;
    p.row_cost = nullptr;
    return p;
}

// common part of the array probes: inputs up, kernel, outputs down
template <typename Launch>
void run_probe(const double *in, size_t in_doubles, double *out, size_t out_doubles, const double *in2, size_t in2_doubles, Launch launch)
{
    Buffer d_in, d_in2, d_out;
    d_in.reserve(sizeof(double) * (in_doubles ? in_doubles : 1));
    d_in2.reserve(sizeof(double) * (in2_doubles ? in2_doubles : 1));
    d_out.reserve(sizeof(double) * (out_doubles ? out_doubles : 1));
    CK(cudaMemcpyAsync(d_in.p, in, sizeof(double) * in_doubles, cudaMemcpyHostToDevice, g.stream));
    if (in2_doubles) CK(cudaMemcpyAsync(d_in2.p, in2, sizeof(double) * in2_doubles, cudaMemcpyHostToDevice, g.stream));
    launch((const double *)d_in.p, (const double *)d_in2.p, (double *)d_out.p);
    CK(cudaMemcpyAsync(out, d_out.p, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    d_in.release();
    d_in2.release();
    d_out.release();
}

} // namespace

extern "C" {

int trt_init(int device)
{
    if (g.ready) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "libtrt_b200: no usable CUDA device (%s); this library has no CPU fallback\n",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        exit(1);
    }
    if (device < 0 || device >= n) {
        fprintf(stderr, "libtrt_b200: device %d out of range (0..%d)\n", device, n - 1);
        exit(1);
    }
    CK(cudaSetDevice(device));
    g.device = device;
    CK(cudaDeviceGetAttribute(&g.num_sms, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    for (auto &ev : g.ev) CK(cudaEventCreate(&ev));
    for (auto &ev : g.arena_done) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    g.arena_busy[0] = g.arena_busy[1] = false;
    for (auto &row : g.chunk_ev)
        for (auto &ev : row) CK(cudaEventCreate(&ev));
    g.tile_counter.reserve(256);
    g.counters.reserve(sizeof(unsigned long long) * TRT_NUM_COUNTERS);
    {
        // k/255.0 for every byte value: the three divisions of TRT.c:866, evaluated once, in host double
        double table[256];
        for (int k = 0; k < 256; k++) {
            volatile double num = (double)k;
            table[k] = num / 255.0;
        }
        g.byte_to_unit.reserve(sizeof table);
        CK(cudaMemcpy(g.byte_to_unit.p, table, sizeof table, cudaMemcpyHostToDevice));
    }
    g.ready = true;
    return 0;
}

void trt_shutdown(void)
{
    if (!g.ready) return;
    cudaStreamSynchronize(g.stream);
    g.scene_blob.release();
    g.sky.release();
    g.tile_counter.release();
    g.scratch.release();
    g.tile_info.release();
    g.byte_to_unit.release();
    g.counters.release();
    g.pixels.release();
    g.quant.release();
    g.bytes.release();
    g.stage.release();
    for (int b = 0; b < 2; b++) {
        g.orbit_dev[b].release();
        g.orbit_host[b].release();
        if (g.orbit_encoded[b]) cudaEventDestroy(g.orbit_encoded[b]);
        if (g.orbit_copied[b]) cudaEventDestroy(g.orbit_copied[b]);
        g.orbit_encoded[b] = g.orbit_copied[b] = nullptr;
    }
    g.arena[0].release();
    g.arena[1].release();
    for (auto &ev : g.ev) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
    }
    for (auto &ev : g.arena_done) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
    }
    for (auto &row : g.chunk_ev)
        for (auto &ev : row) {
            if (ev) cudaEventDestroy(ev);
            ev = nullptr;
        }
    cudaStreamDestroy(g.copy_stream);
    g.copy_stream = nullptr;
    cudaStreamDestroy(g.own_stream);
    g.stream = g.own_stream = nullptr;
    g.have_scene = false;
    g.sky_dim = -1;
    g.ready = false;
}

int trt_is_initialized(void) { return g.ready ? 1 : 0; }
void *trt_stream(void) { return (void *)g.stream; }

int trt_set_stream(void *cuda_stream)
{
    require_init("trt_set_stream");
    CK(cudaStreamSynchronize(g.stream));
    g.stream = (cudaStream_t)cuda_stream;   // NULL is a valid handle: the legacy default stream
    return 0;
}

int trt_set_cull(int enabled)
{
    g.cull_allowed = enabled != 0;
    return 0;
}

int trt_use_own_stream(void)
{
    require_init("trt_use_own_stream");
    CK(cudaStreamSynchronize(g.stream));
    g.stream = g.own_stream;
    return 0;
}

int trt_upload_skybox(const trt_Skybox *skybox)
{
    require_init("trt_upload_skybox");
    const int dim = skybox->dim;
    if (dim <= 0) {
        fprintf(stderr, "libtrt_b200: skybox has no faces loaded (dim=%d)\n", dim);
        exit(1);
    }
    // RGBA8 repack: one aligned 4-byte load per lookup instead of three byte loads.  Each face keeps
    // dim+1 black texels after its payload: the reference's index can run that far (TRT.c:778-788).
    const size_t payload = (size_t)dim * (size_t)dim;
    const size_t stride = payload + (size_t)dim + 1;
    std::vector<uchar4> packed(stride * 6, make_uchar4(0, 0, 0, 0));
    for (int f = 0; f < 6; f++) {
        const trt_Color *src = skybox->colors[f];
        uchar4 *dst = packed.data() + stride * (size_t)f;
        for (size_t i = 0; i < payload; i++) dst[i] = make_uchar4(src[i].r, src[i].g, src[i].b, 0);
    }
    g.sky.reserve(sizeof(uchar4) * packed.size());
    CK(cudaMemcpyAsync(g.sky.p, packed.data(), sizeof(uchar4) * packed.size(), cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.sky_dim = dim;
    g.sky_face_stride = (int)stride;
    if (g.have_scene) {
        // the band API (trt_render_rows_*) renders with the constants of the last trt_set_scene: they carry the skybox
        // geometry, so a skybox uploaded after the scene has to refresh them (set_scene -> upload_skybox -> render is legal)
        g.scene.sky_dim = dim;
        g.scene.sky_face_stride = (int)stride;
        upload_scene_constants(g.scene, g.stream);
        CK(cudaStreamSynchronize(g.stream));
    }
    return 0;
}

int trt_set_scene(const trt_Scene *scene)
{
    require_init("trt_set_scene");
    upload_scene(scene);
    return 0;
}

int trt_set_scene_async(const trt_Scene *scene)
{
    require_init("trt_set_scene_async");
    upload_scene(scene, false);
    return 0;
}

int trt_render_rows_device(int width, int height, int row0, int row1, double *d_pixels)
{
    require_init("trt_render_rows_device");
    RenderParams p = make_params(width, height, row0, row1, d_pixels, nullptr, false);
    launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    return 0;
}

int trt_render_rows_quant_device(int width, int height, int row0, int row1, unsigned char *d_quant)
{
    require_init("trt_render_rows_quant_device");
    RenderParams p = make_params(width, height, row0, row1, nullptr, (uchar4 *)d_quant, false);
    launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    return 0;
}

int trt_render_rows_ansi_device(int width, int height, int row0, int row1, char *stream)
{
    require_init("trt_render_rows_ansi_device");
    RenderParams p = make_params(width, height, row0, row1, nullptr, nullptr, false);
    p.ansi = (unsigned char *)stream;
    launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    return 0;
}

int trt_encode_rows_device(const double *d_pixels, int width, int rows, char *d_bytes, size_t byte_offset)
{
    require_init("trt_encode_rows_device");
    launch_encode_f64(d_pixels, width, rows, d_bytes, byte_offset, g.stream);
    return 0;
}

int trt_encode_rows_quant_device(const unsigned char *d_quant, int width, int rows, char *d_bytes, size_t byte_offset)
{
    require_init("trt_encode_rows_quant_device");
    launch_encode_quant((const uchar4 *)d_quant, width, rows, d_bytes, byte_offset, g.stream);
    return 0;
}

int trt_stream_frame_device(char *d_stream, int width, int height)
{
    require_init("trt_stream_frame_device");
    launch_stream_frame(d_stream, width, height, g.stream);
    return 0;
}

int trt_count_rows_device(int width, int height, int row0, int row1, double *d_pixels, long long *counters)
{
    require_init("trt_count_rows_device");
    CK(cudaMemsetAsync(g.counters.p, 0, sizeof(unsigned long long) * TRT_NUM_COUNTERS, g.stream));
    RenderParams p = make_params(width, height, row0, row1, d_pixels, nullptr, true);
    launch_render(p, true, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    CK(cudaMemcpyAsync(counters, g.counters.p, sizeof(unsigned long long) * TRT_NUM_COUNTERS, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int trt_estimate_row_costs(const trt_Scene *scene, int width, int height, double *cost_per_row)
{
    require_init("trt_estimate_row_costs");
    if (width <= 0 || height <= 0) return 0;
    // Same camera, 1/8 of the rows and columns (at least 64 x 32): per-row closest-hit counts of the small frame
    // are spread over the rows of the big one.  Sky rows cost ~1 query per sample, sphere/ground rows 5+.
    int sw = width / 8, sh = height / 8;
    if (sw < 64) sw = width < 64 ? width : 64;
    if (sh < 32) sh = height < 32 ? height : 32;
    upload_scene(scene);
    Buffer d_cost, d_quant;
    d_cost.reserve(sizeof(unsigned int) * (size_t)sh);
    d_quant.reserve(sizeof(uchar4) * (size_t)sw * (size_t)sh);
    CK(cudaMemsetAsync(d_cost.p, 0, sizeof(unsigned int) * (size_t)sh, g.stream));
    // the counting flavour of K1 carries the per-row cost accounting (the production flavour has no branch for it)
    g.counters.reserve(sizeof(unsigned long long) * TRT_NUM_COUNTERS);
    CK(cudaMemsetAsync(g.counters.p, 0, sizeof(unsigned long long) * TRT_NUM_COUNTERS, g.stream));
    RenderParams p = make_params(sw, sh, 0, sh, nullptr, (uchar4 *)d_quant.p, true);
    p.row_cost = (unsigned int *)d_cost.p;
    launch_render(p, true, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    std::vector<unsigned int> cost((size_t)sh);
    CK(cudaMemcpyAsync(cost.data(), d_cost.p, sizeof(unsigned int) * (size_t)sh, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    d_cost.release();
    d_quant.release();
    for (int r = 0; r < height; r++) {
        const int sr = (int)(((long long)r * sh) / height);
        cost_per_row[r] = (double)cost[(size_t)sr] + 1.0;
    }
    return 0;
}

double trt_model_flops(const long long *c)
{
    // as-written adds/muls/divs/sqrts of the reference per counted event
    return 25.0 * c[CTR_SPHERE_TESTS] + 5.0 * c[CTR_SPHERE_DISC_OK] + 14.0 * c[CTR_SPHERE_T0_POS] + 3.0 * c[CTR_SPHERE_CLOSEST] +
           5.0 * c[CTR_PLANE_TESTS] + 9.0 * c[CTR_PLANE_DENOM_OK] + 14.0 * c[CTR_PLANE_T_POS] + 1.0 * c[CTR_PLANE_CLOSEST] +
           90.0 * c[CTR_SKY_LOOKUPS] + 27.0 * c[CTR_TRACE_HITS] + 55.0 * c[CTR_LIGHTING_CALLS] + 34.0 * c[CTR_BOUNCE_ITERS] +
           68.0 * c[CTR_SAMPLES] + 4.0 * c[CTR_PIXELS];
}

// ---- unit-level probes (parity tests of single queries) -------------------------------------------------

int trt_probe_trace_ray(const trt_Scene *scene, const double *rays, int n, double *out)
{
    require_init("trt_probe_trace_ray");
    upload_scene(scene);
    if (n <= 0) return 0;
    Buffer d_in, d_out;
    d_in.reserve(sizeof(double) * 6 * (size_t)n);
    d_out.reserve(sizeof(double) * 11 * (size_t)n);
    CK(cudaMemcpyAsync(d_in.p, rays, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, g.stream));
    RenderParams p = make_params(1, 1, 0, 1, nullptr, nullptr, false);
    launch_probe_trace(p, (const double *)d_in.p, n, (double *)d_out.p, g.stream);
    CK(cudaMemcpyAsync(out, d_out.p, sizeof(double) * 11 * (size_t)n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    d_in.release();
    d_out.release();
    return 0;
}

int trt_probe_sphere(const double *rays, const double *spheres, int n, double *out)
{
    require_init("trt_probe_sphere");
    if (n <= 0) return 0;
    run_probe(rays, 6 * (size_t)n, out, 4 * (size_t)n, spheres, 4 * (size_t)n,
              [&](const double *a, const double *b, double *o) { launch_probe_sphere(a, b, n, o, g.stream); });
    return 0;
}

int trt_probe_plane(const trt_Scene *scene, const double *rays, int n, double *out)
{
    require_init("trt_probe_plane");
    upload_scene(scene);
    if (n <= 0) return 0;
    run_probe(rays, 6 * (size_t)n, out, 4 * (size_t)n, nullptr, 0,
              [&](const double *a, const double *, double *o) { launch_probe_plane(a, n, o, g.stream); });
    return 0;
}

int trt_probe_lighting(const trt_Scene *scene, const double *surface, int n, double *out)
{
    require_init("trt_probe_lighting");
    upload_scene(scene);
    if (n <= 0) return 0;
    RenderParams p = make_params(1, 1, 0, 1, nullptr, nullptr, false);
    run_probe(surface, 9 * (size_t)n, out, 3 * (size_t)n, nullptr, 0,
              [&](const double *a, const double *, double *o) { launch_probe_lighting(p, a, n, o, g.stream); });
    return 0;
}

int trt_probe_skybox(const double *dirs, int n, int *out)
{
    require_init("trt_probe_skybox");
    if (n <= 0) return 0;
    if (!g.have_scene) {
        // the sampler only needs the skybox fields of the constant block
        memset(&g.scene, 0, sizeof g.scene);
        g.scene.sky_dim = g.sky_dim;
        g.scene.sky_face_stride = g.sky_face_stride;
        upload_scene_constants(g.scene, g.stream);
        g.scene_blob.reserve(4096);
        g.off_geom = g.off_cull = g.off_mat = g.off_pairs = g.off_orig = g.off_pos = g.off_ball = g.off_link = 0;
        g.have_scene = true;
    } else {
        g.scene.sky_dim = g.sky_dim;
        g.scene.sky_face_stride = g.sky_face_stride;
        upload_scene_constants(g.scene, g.stream);
    }
    Buffer d_in, d_out;
    d_in.reserve(sizeof(double) * 3 * (size_t)n);
    d_out.reserve(sizeof(int) * 5 * (size_t)n);
    CK(cudaMemcpyAsync(d_in.p, dirs, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, g.stream));
    RenderParams p = make_params(1, 1, 0, 1, nullptr, nullptr, false);
    launch_probe_sky(p, (const double *)d_in.p, n, (int *)d_out.p, g.stream);
    CK(cudaMemcpyAsync(out, d_out.p, sizeof(int) * 5 * (size_t)n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    d_in.release();
    d_out.release();
    return 0;
}

long long trt_selftest_division(unsigned long long seed, long long quotients)
{
    require_init("trt_selftest_division");
    // 3 quotients per iteration, 256 threads per CTA
    const int iters = 4096;
    long long ctas = quotients / (3LL * 256 * iters);
    if (ctas < 1) ctas = 1;
    if (ctas > 1 << 20) ctas = 1 << 20;
    return (long long)run_selftest_division(seed, (int)ctas, iters, (unsigned long long *)g.counters.p, g.stream);
}

// ---- drop-ins ---------------------------------------------------------------------------------------

void trt_project_scene(const trt_Scene *scene, trt_Screen *screen)
{
    require_init("trt_project_scene");
    const int w = screen->width, h = screen->height;
    if (w <= 0 || h <= 0) return;
    upload_scene(scene);
    const size_t bytes = sizeof(double) * 3 * (size_t)w * (size_t)h;
    g.pixels.reserve(bytes);
    RenderParams p = make_params(w, h, 0, h, (double *)g.pixels.p, nullptr, false);
    CK(cudaEventRecord(g.ev[0], g.stream));
    launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
    CK(cudaEventRecord(g.ev[1], g.stream));
    CK(cudaMemcpyAsync(screen->pixels, g.pixels.p, bytes, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaEventElapsedTime(&g.last_render_ms, g.ev[0], g.ev[1]));
}

size_t trt_draw_screen(const trt_Screen *screen, char *out)
{
    require_init("trt_draw_screen");
    const int w = screen->width, h = screen->height;
    const size_t total = TRT_STREAM_BYTES(w, h);
    const size_t px_bytes = sizeof(double) * 3 * (size_t)w * (size_t)h;
    g.pixels.reserve(px_bytes ? px_bytes : 8);
    g.bytes.reserve(total + 16);
    if (px_bytes) CK(cudaMemcpyAsync(g.pixels.p, screen->pixels, px_bytes, cudaMemcpyHostToDevice, g.stream));
    CK(cudaEventRecord(g.ev[2], g.stream));
    launch_stream_frame((char *)g.bytes.p, w, h, g.stream);
    launch_encode_f64((const double *)g.pixels.p, w, h, (char *)g.bytes.p, TRT_HOME_BYTES, g.stream);
    CK(cudaEventRecord(g.ev[3], g.stream));
    CK(cudaMemcpyAsync(out, g.bytes.p, total, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaEventElapsedTime(&g.last_encode_ms, g.ev[2], g.ev[3]));
    return total;
}

void trt_buffered_draw_screen(const trt_Screen *screen)
{
    require_init("trt_buffered_draw_screen");
    const size_t total = TRT_STREAM_BYTES(screen->width, screen->height);
    g.stage.reserve(total);
    trt_draw_screen(screen, (char *)g.stage.p);
    fwrite(g.stage.p, sizeof(char), total, stdout); // TRT.c:1171
}

size_t trt_render_ansi(const trt_Scene *scene, int width, int height, char *out, size_t cap)
{
    require_init("trt_render_ansi");
    const size_t total = TRT_STREAM_BYTES(width, height);
    if (cap < total || width <= 0 || height <= 0) return 0;
    upload_scene(scene, false);            // ordered on the stream in front of the kernels: no host wait
    g.quant.reserve(sizeof(uchar4) * (size_t)width * (size_t)height);
    g.bytes.reserve(total + 16);
    // Big frames are rendered as up to MAX_CHUNKS row chunks so that the device-to-host copy of a chunk's
    // bytes (829 MB per 7680x4320 frame: PCIe time comparable to K1's) overlaps the rendering of the next one.
    // All kernels are enqueued first; the copies run on a second stream, each behind its chunk's event, so the
    // overlap also happens when `out` is pageable memory (where cudaMemcpyAsync blocks the host).
    // Chunks shrink geometrically (each ~60 % of the one before): only the LAST chunk's copy is exposed after the last
    // kernel, and every extra launch costs a kernel tail, so few chunks with a small last one beat many equal ones.
    const size_t pixels = (size_t)width * (size_t)height;
    int chunks = 1;
    for (size_t px = pixels >> 20; px > 1 && chunks < Context::MAX_CHUNKS; px >>= 1) chunks++;   // 1 Mpixel frames: 1 chunk ... 32 Mpixel: 6
    if (chunks > height) chunks = height;
    int bounds[Context::MAX_CHUNKS + 1];
    {
        double weight[Context::MAX_CHUNKS], total = 0.0, acc = 0.0;
        for (int c = 0; c < chunks; c++) total += (weight[c] = pow(0.6, c));
        bounds[0] = 0;
        for (int c = 0; c < chunks; c++) {
            acc += weight[c];
            int r = (int)(height * (acc / total) + 0.5);
            if (r <= bounds[c]) r = bounds[c] + 1;              // never empty
            if (r > height - (chunks - 1 - c)) r = height - (chunks - 1 - c);
            bounds[c + 1] = c == chunks - 1 ? height : r;
        }
    }
    const size_t row_bytes = TRT_ROW_BYTES(width);
    launch_stream_frame((char *)g.bytes.p, width, height, g.stream);
    for (int c = 0; c < chunks; c++) {
        const int r0 = bounds[c], r1 = bounds[c + 1];
        RenderParams p = make_params(width, height, r0, r1, nullptr, (uchar4 *)g.quant.p + (size_t)r0 * (size_t)width, false);
        CK(cudaEventRecord(g.chunk_ev[c][0], g.stream));
        launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
        CK(cudaEventRecord(g.chunk_ev[c][1], g.stream));
        launch_encode_quant((const uchar4 *)g.quant.p + (size_t)r0 * (size_t)width, width, r1 - r0, (char *)g.bytes.p,
                            TRT_HOME_BYTES + (size_t)r0 * row_bytes, g.stream);
        CK(cudaEventRecord(g.chunk_ev[c][2], g.stream));
    }
    for (int c = 0; c < chunks; c++) {
        const int r0 = bounds[c], r1 = bounds[c + 1];
        const size_t b0 = c == 0 ? 0 : TRT_HOME_BYTES + (size_t)r0 * row_bytes;
        const size_t b1 = c == chunks - 1 ? total : TRT_HOME_BYTES + (size_t)r1 * row_bytes;
        CK(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[c][2], 0));
        CK(cudaMemcpyAsync(out + b0, (const char *)g.bytes.p + b0, b1 - b0, cudaMemcpyDeviceToHost, g.copy_stream));
    }
    CK(cudaStreamSynchronize(g.copy_stream));
    CK(cudaStreamSynchronize(g.stream));
    g.last_render_ms = g.last_encode_ms = 0.f;
    for (int c = 0; c < chunks; c++) {
        float k1 = 0.f, k2 = 0.f;
        CK(cudaEventElapsedTime(&k1, g.chunk_ev[c][0], g.chunk_ev[c][1]));
        CK(cudaEventElapsedTime(&k2, g.chunk_ev[c][1], g.chunk_ev[c][2]));
        g.last_render_ms += k1;
        g.last_encode_ms += k2;
    }
    return total;
}

// The frame loop behind trt_render_orbit / trt_render_orbit_to.  `acquire` names the page-locked destination of a frame's
// bytes (it may block until one is free: called after the frame's kernels have been enqueued, so the GPU works meanwhile);
// `sink` is called once the bytes have landed.  Two device buffers: the copy of frame k runs while frame k+1 renders.
static int orbit_loop(const trt_Scene *scene, int width, int height, const double *times, int n_frames, int first, int stride,
                      trt_frame_acquire acquire, trt_frame_sink sink, void *user)
{
    const size_t total = TRT_STREAM_BYTES(width, height);
    g.quant.reserve(sizeof(uchar4) * (size_t)width * (size_t)height);
    Buffer (&d_bytes)[2] = g.orbit_dev;
    cudaEvent_t (&encoded)[2] = g.orbit_encoded, (&copied)[2] = g.orbit_copied;
    for (int b = 0; b < 2; b++) {
        d_bytes[b].reserve(total + 16);
        if (!encoded[b]) CK(cudaEventCreateWithFlags(&encoded[b], cudaEventDisableTiming));
        if (!copied[b]) CK(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
    }
    trt_Scene posed = *scene;
    int launched = 0, delivered = 0, pending_frame[2] = {-1, -1};
    char *pending_dst[2] = {nullptr, nullptr};
    bool stop = false;
    auto deliver = [&](int b) {
        // frame pending_frame[b] is on its way to pending_dst[b]: wait for the copy, hand it to the sink
        CK(cudaEventSynchronize(copied[b]));
        if (!stop) {
            if (sink && sink(pending_dst[b], total, pending_frame[b], user) != 0) stop = true;
            delivered++;
        }
        pending_frame[b] = -1;
    };
    for (int k = first; k < n_frames && !stop; k += stride) {
        const int b = launched & 1;
        if (pending_frame[b] >= 0) deliver(b);          // frame k-2*stride: its device buffer and upload arena are reused now
        if (stop) break;
        posed.camera = scene->camera;
        trt_orbit_camera(&posed.camera, times[k]);
        upload_scene(&posed, false);
        RenderParams p = make_params(width, height, 0, height, nullptr, (uchar4 *)g.quant.p, false);
        launch_render(p, false, cull_mode(), one_plus_one(), g.num_sms, g.stream);
        launch_stream_frame((char *)d_bytes[b].p, width, height, g.stream);
        launch_encode_quant((const uchar4 *)g.quant.p, width, height, (char *)d_bytes[b].p, TRT_HOME_BYTES, g.stream);
        CK(cudaEventRecord(encoded[b], g.stream));
        char *dst = acquire(k, total, user);            // may block: the kernels above are already running
        if (!dst) {
            stop = true;
            break;
        }
        CK(cudaStreamWaitEvent(g.copy_stream, encoded[b], 0));
        CK(cudaMemcpyAsync(dst, d_bytes[b].p, total, cudaMemcpyDeviceToHost, g.copy_stream));
        CK(cudaEventRecord(copied[b], g.copy_stream));
        pending_frame[b] = k;
        pending_dst[b] = dst;
        launched++;
        // while this frame renders, deliver the previous one
        if (pending_frame[b ^ 1] >= 0) deliver(b ^ 1);
    }
    for (int i = 0; i < 2; i++) {
        const int b = (launched + i) & 1;               // older buffer first
        if (pending_frame[b] >= 0) deliver(b);
    }
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaStreamSynchronize(g.copy_stream));
    return delivered;
}

namespace {
struct OwnBuffers {
    int next = 0;
    trt_frame_sink sink = nullptr;
    void *user = nullptr;
};
char *own_acquire(int, size_t, void *u)
{
    OwnBuffers *o = (OwnBuffers *)u;
    return (char *)g.orbit_host[(o->next++) & 1].p;     // frame k-2's buffer: orbit_loop delivered that frame before asking again
}
int own_sink(const char *bytes, size_t n, int frame, void *u)
{
    OwnBuffers *o = (OwnBuffers *)u;
    return o->sink(bytes, n, frame, o->user);
}
} // namespace

int trt_render_orbit(const trt_Scene *scene, int width, int height, const double *times, int n_frames, int first, int stride,
                     trt_frame_sink sink, void *user)
{
    require_init("trt_render_orbit");
    if (width <= 0 || height <= 0 || n_frames <= 0 || stride <= 0 || first < 0 || !sink) return 0;
    OwnBuffers own;
    own.sink = sink;
    own.user = user;
    for (int b = 0; b < 2; b++) g.orbit_host[b].reserve(TRT_STREAM_BYTES(width, height));
    return orbit_loop(scene, width, height, times, n_frames, first, stride, own_acquire, own_sink, &own);
}

int trt_render_orbit_to(const trt_Scene *scene, int width, int height, const double *times, int n_frames, int first, int stride,
                        trt_frame_acquire acquire, trt_frame_sink sink, void *user)
{
    require_init("trt_render_orbit_to");
    if (width <= 0 || height <= 0 || n_frames <= 0 || stride <= 0 || first < 0 || !acquire) return 0;
    return orbit_loop(scene, width, height, times, n_frames, first, stride, acquire, sink, user);
}

// ---- gather over NVLink peer memory ------------------------------------------------------------------------

int trt_ipc_export(const void *d_ptr, unsigned char handle[TRT_IPC_HANDLE_BYTES])
{
    require_init("trt_ipc_export");
    static_assert(sizeof(cudaIpcMemHandle_t) == TRT_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return 0;
}

void *trt_ipc_import(const unsigned char handle[TRT_IPC_HANDLE_BYTES])
{
    require_init("trt_ipc_import");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
}

int trt_ipc_close(void *d_peer_ptr)
{
    if (d_peer_ptr) CK(cudaIpcCloseMemHandle(d_peer_ptr));
    return 0;
}

int trt_push_to_peer(void *d_peer_dst, const void *d_src, size_t bytes)
{
    require_init("trt_push_to_peer");
    if (!bytes) return 0;
    // order the transfer after the encode that produced the bytes, run it on the copy stream (copy engine, NVLink)
    CK(cudaEventRecord(g.ev[3], g.stream));
    CK(cudaStreamWaitEvent(g.copy_stream, g.ev[3], 0));
    CK(cudaMemcpyAsync(d_peer_dst, d_src, bytes, cudaMemcpyDefault, g.copy_stream));
    return 0;
}

int trt_signal_step(void *d_flag, unsigned int value, int after_copies)
{
    require_init("trt_signal_step");
    // behind the kernels of this step (trt_stream()) or, after_copies, behind the copy-engine pushes on the copy stream
    launch_signal((unsigned int *)d_flag, value, after_copies ? g.copy_stream : g.stream);
    return 0;
}

int trt_wait_steps(const void *d_flags, int n_flags, unsigned int value, int on_copy_stream)
{
    require_init("trt_wait_steps");
    if (n_flags > 32 || n_flags <= 0) return -1;
    // word 32 behind the first flag waited for records a timeout (a rank that never signalled)
    launch_wait_flags((const unsigned int *)d_flags, n_flags, value, (unsigned int *)d_flags + 32, on_copy_stream ? g.copy_stream : g.stream);
    return 0;
}

int trt_stream_wait_copies(void)
{
    require_init("trt_stream_wait_copies");
    // work enqueued on trt_stream() from here on starts after everything enqueued on the copy stream so far (the previous
    // step's pushes still read the buffers the next encode overwrites)
    CK(cudaEventRecord(g.ev[2], g.copy_stream));
    CK(cudaStreamWaitEvent(g.stream, g.ev[2], 0));
    return 0;
}

int trt_peer_copies_wait(void)
{
    require_init("trt_peer_copies_wait");
    CK(cudaStreamSynchronize(g.copy_stream));
    return 0;
}

int trt_host_register(void *host_ptr, size_t bytes)
{
    require_init("trt_host_register");
    CK(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int trt_host_unregister(void *host_ptr)
{
    if (host_ptr) CK(cudaHostUnregister(host_ptr));
    return 0;
}

// ---- small helpers for plain-C callers -----------------------------------------------------------------

void *trt_device_alloc(size_t bytes)
{
    require_init("trt_device_alloc");
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes ? bytes : 1));
    return p;
}
void trt_device_free(void *p)
{
    if (p) CK(cudaFree(p));
}
void *trt_host_alloc_pinned(size_t bytes)
{
    require_init("trt_host_alloc_pinned");
    void *p = nullptr;
    CK(cudaMallocHost(&p, bytes ? bytes : 1));
    return p;
}
void trt_host_free_pinned(void *p)
{
    if (p) CK(cudaFreeHost(p));
}
int trt_copy_to_host(void *dst, const void *d_src, size_t bytes)
{
    require_init("trt_copy_to_host");
    CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int trt_copy_to_device(void *d_dst, const void *src, size_t bytes)
{
    require_init("trt_copy_to_device");
    CK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int trt_synchronize(void)
{
    require_init("trt_synchronize");
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}
int trt_debug_bounds(unsigned int *out32)
{
    require_init("trt_debug_bounds");
    const int a = render_bounds_read(out32), b = encode_bounds_read(out32 + 16);
    return a && b;
}
float trt_last_render_ms(void) { return g.last_render_ms; }
float trt_last_encode_ms(void) { return g.last_encode_ms; }

double trt_measure_fp32_tflops(void)
{
    require_init("trt_measure_fp32_tflops");
    return measure_fp32_tflops(g.stream);
}
double trt_measure_fp64_tflops(void)
{
    require_init("trt_measure_fp64_tflops");
    return measure_fp64_tflops(g.stream);
}

} // extern "C"
