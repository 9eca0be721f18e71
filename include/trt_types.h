/*
 * trt_types.h — plain-C data layout that crosses the host <-> libtrt_b200 boundary.
 *
 * These are layout-compatible restatements of the reference's own structs
 * (/root/reference/TerminalRayTracer.c:61-208, "TRT.c" below): same member order,
 * same member types, all `double` based, so a `Scene*` / `Screen*` / `Skybox*`
 * built by the reference's host code can be handed to the library unchanged.
 *
 * Every type is declared as trt_<Name>; the reference's bare names (Scene, Screen …)
 * are provided as aliases unless TRT_NO_REFERENCE_NAMES is defined (the oracle
 * harness that #includes the reference translation unit defines it, because the
 * reference declares those names itself).
 */
#ifndef TRT_TYPES_H
#define TRT_TYPES_H

#ifdef __cplusplus
extern "C" {
#endif

/* compile-time constants of the reference that are part of the rendering contract */
#define TRT_EPSILON 0.000001          /* TRT.c:44  hit-point push-back            */
#define TRT_BOUNCE_LIMIT 10           /* TRT.c:54  max mirror bounces per sample   */
#define TRT_RAYS_PER_PIXEL 10         /* TRT.c:58  samples per pixel               */
#define TRT_PI 3.14159265358979323846 /* TRT.c:43                                  */
#define TRT_DEFAULT_WIDTH 480         /* TRT.c:47                                  */
#define TRT_DEFAULT_HEIGHT 280        /* TRT.c:48                                  */
#define TRT_CELL_BYTES 25             /* TRT.c:1103 "\033[48;2;RRR;GGG;BBBm  \033[0m" */
#define TRT_HOME_BYTES 6              /* TRT.c:1102 "\033[0;0H"                     */
#define TRT_TAIL_NULS 3               /* TRT.c:1104 array is 3 bytes longer than its text */

/* TRT.c:61-67 */
typedef enum { TRT_NONE = 0, TRT_SPHERE = 1, TRT_GROUND = 2 } trt_ObjectType;

/* TRT.c:70-83 — Point and Vector are the same three doubles */
typedef struct { double x, y, z; } trt_Point;
typedef struct { double x, y, z; } trt_Vector;

/* TRT.c:92-104 */
typedef struct { trt_Vector x, y, z; } trt_Basis;
typedef struct { trt_Basis basis; trt_Point origin; } trt_Frame;

/* TRT.c:107-111 */
typedef struct { trt_Point origin; trt_Vector direction; } trt_Ray;

/* TRT.c:114-119 — specularity is carried but never read by the render path */
typedef struct { trt_Vector color; double reflectivity; double specularity; } trt_Material;

/* TRT.c:122-127 */
typedef struct { unsigned char r, g, b; } trt_Color;

/* TRT.c:130-134 — six separately allocated dim*dim planes, order +X,-X,+Y,-Y,+Z,-Z */
typedef struct { trt_Color *colors[6]; int dim; } trt_Skybox;

/* TRT.c:146-158 */
typedef struct { trt_Vector direction; trt_Vector color; } trt_DirectionalLight;
typedef struct { trt_Point position; trt_Vector color; double intensity; } trt_PointLight;

/* TRT.c:161-175 */
typedef struct { trt_Point center; double radius; trt_Material material; } trt_Sphere;
typedef struct { trt_Point point; trt_Vector normal; trt_Material even_material; trt_Material odd_material; } trt_Plane;

/* TRT.c:178-184 */
typedef struct { trt_Frame frame; double screen_distance; double screen_width; double screen_height; } trt_Camera;

/* TRT.c:188-193 — pixels[row*width+column], caller-owned */
typedef struct { trt_Vector *pixels; int width; int height; } trt_Screen;

/* TRT.c:196-208 */
typedef struct {
    trt_Sphere *spheres;
    int num_spheres;
    trt_Plane ground;
    trt_DirectionalLight *directional_lights;
    int num_directional_lights;
    trt_PointLight *point_lights;
    int num_point_lights;
    trt_Camera camera;
    trt_Skybox skybox;
} trt_Scene;

#ifndef TRT_NO_REFERENCE_NAMES
typedef trt_ObjectType ObjectType;
typedef trt_Point Point;
typedef trt_Vector Vector;
typedef trt_Basis Basis;
typedef trt_Frame Frame;
typedef trt_Ray Ray;
typedef trt_Material Material;
typedef trt_Color Color;
typedef trt_Skybox Skybox;
typedef trt_DirectionalLight DirectionalLight;
typedef trt_PointLight PointLight;
typedef trt_Sphere Sphere;
typedef trt_Plane Plane;
typedef trt_Camera Camera;
typedef trt_Screen Screen;
typedef trt_Scene Scene;
#endif

/* bytes in the terminal stream for a w x h screen: TRT.c:1104
 * (sizeof(reset_str)+1) + ((sizeof(pixel_str)-1)*W + 1)*H + 1  ==  9 + (25W+1)H  */
#define TRT_STREAM_BYTES(w, h) ((size_t)9 + ((size_t)TRT_CELL_BYTES * (size_t)(w) + 1) * (size_t)(h))
#define TRT_ROW_BYTES(w) ((size_t)TRT_CELL_BYTES * (size_t)(w) + 1)

#ifdef __cplusplus
}
#endif
#endif /* TRT_TYPES_H */
