"""ctypes mirror of include/trt_types.h (itself a layout restatement of
/root/reference/TerminalRayTracer.c:61-208).  Used by the product binding (lib.py) and by the
tests, which hand the very same structures to the reference build and to the oracle."""
import ctypes as C

EPSILON = 0.000001
BOUNCE_LIMIT = 10
RAYS_PER_PIXEL = 10
CELL_BYTES = 25
HOME_BYTES = 6
HOME = b"\x1b[0;0H"        # reset_str, TRT.c:1102
TAIL_NULS = 3
DEMO_SPHERES = 6
NUM_COUNTERS = 32


class Vector(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def tup(self):
        return (self.x, self.y, self.z)


Point = Vector


class Basis(C.Structure):
    _fields_ = [("x", Vector), ("y", Vector), ("z", Vector)]


class Frame(C.Structure):
    _fields_ = [("basis", Basis), ("origin", Vector)]


class Ray(C.Structure):
    _fields_ = [("origin", Vector), ("direction", Vector)]


class Material(C.Structure):
    _fields_ = [("color", Vector), ("reflectivity", C.c_double), ("specularity", C.c_double)]


class Color(C.Structure):
    _fields_ = [("r", C.c_ubyte), ("g", C.c_ubyte), ("b", C.c_ubyte)]


class Skybox(C.Structure):
    _fields_ = [("colors", C.POINTER(Color) * 6), ("dim", C.c_int)]


class DirectionalLight(C.Structure):
    _fields_ = [("direction", Vector), ("color", Vector)]


class PointLight(C.Structure):
    _fields_ = [("position", Vector), ("color", Vector), ("intensity", C.c_double)]


class Sphere(C.Structure):
    _fields_ = [("center", Vector), ("radius", C.c_double), ("material", Material)]


class Plane(C.Structure):
    _fields_ = [("point", Vector), ("normal", Vector), ("even_material", Material), ("odd_material", Material)]


class Camera(C.Structure):
    _fields_ = [("frame", Frame), ("screen_distance", C.c_double), ("screen_width", C.c_double),
                ("screen_height", C.c_double)]


class Screen(C.Structure):
    _fields_ = [("pixels", C.POINTER(Vector)), ("width", C.c_int), ("height", C.c_int)]


class Scene(C.Structure):
    _fields_ = [
        ("spheres", C.POINTER(Sphere)),
        ("num_spheres", C.c_int),
        ("ground", Plane),
        ("directional_lights", C.POINTER(DirectionalLight)),
        ("num_directional_lights", C.c_int),
        ("point_lights", C.POINTER(PointLight)),
        ("num_point_lights", C.c_int),
        ("camera", Camera),
        ("skybox", Skybox),
    ]


def stream_bytes(width: int, height: int) -> int:
    """sizeof(screenbuffer) for a width x height screen, TRT.c:1104."""
    return 9 + (CELL_BYTES * width + 1) * height


def row_bytes(width: int) -> int:
    return CELL_BYTES * width + 1
