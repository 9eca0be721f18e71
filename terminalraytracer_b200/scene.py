"""Host-side mirror of the reference's caller code for the render path: the demo scene literals
(TRT.c:1256-1306), the orbit camera (TRT.c:1327-1336), skybox ingest (TRT.c:388-427) and the
synthetic stand-in for the skybox the reference ships without (skybox/milky_way, see
/root/reference/.MISSING_LARGE_BLOBS).  All numerics are done by the C helpers inside libtrt_b200
(csrc/trt_host.c); this module only owns the buffers and keeps them alive."""
import ctypes as C
import os

import numpy as np

from . import abi, lib as _lib

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACE_FILES = ("+X.ppm", "-X.ppm", "+Y.ppm", "-Y.ppm", "+Z.ppm", "-Z.ppm")  # TRT.c:390


class SkyboxData:
    """Six dim*dim RGB planes, each followed by dim+1 zero texels (see csrc/trt_host.c), plus the
    ctypes trt_Skybox that points at them."""

    def __init__(self, planes):
        dim = planes[0].shape[0]
        self.dim = dim
        self.planes = []
        self.c = abi.Skybox()
        for f, p in enumerate(planes):
            assert p.shape == (dim, dim, 3) and p.dtype == np.uint8
            buf = np.zeros((dim * dim + dim + 1, 3), dtype=np.uint8)
            buf[: dim * dim] = p.reshape(-1, 3)
            self.planes.append(buf)
            self.c.colors[f] = buf.ctypes.data_as(C.POINTER(abi.Color))
        self.c.dim = dim

    def face(self, f):
        return self.planes[f][: self.dim * self.dim].reshape(self.dim, self.dim, 3)


def read_ppm(path):
    """P6 reader with the grammar the reference accepts (TRT.c:309-380): magic, one whitespace,
    '#' comment lines, W H, maxval == 255, one whitespace, payload."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] != b"P6":
        raise ValueError("file is not ppm")
    pos = 3
    while data[pos:pos + 1] == b"#":
        pos = data.index(b"\n", pos) + 1
    fields = []
    while len(fields) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        fields.append(int(data[start:pos]))
    pos += 1
    w, h, maxval = fields
    if maxval != 255:
        raise ValueError("max color value is not 255")
    return np.frombuffer(data, dtype=np.uint8, count=w * h * 3, offset=pos).reshape(h, w, 3).copy()


def write_ppm(path, img):
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


def load_skybox_dir(directory):
    planes = [read_ppm(os.path.join(directory, name)) for name in FACE_FILES]
    dim = planes[0].shape[1]
    for p in planes:
        if p.shape[0] != dim or p.shape[1] != dim:
            raise ValueError("all faces of the skybox must be the same size")
    return SkyboxData(planes)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_cubemap(name="milky_way", dim=1024, seed=0x5EED):
    """Deterministic stand-ins for skybox folders.

    milky_way : black sky, ~0.3% star texels of random brightness and a soft bright band;
                integer hashing only for the stars, so the bytes do not depend on the numpy version.
    colors    : the reference's face-selection test card (one flat colour per face,
                +X red, -X cyan, +Y green, -Y magenta, +Z blue, -Z yellow).
    uv_gradient : r = column, g = row ramps plus a 32-texel checker in b — an orientation test card in
                the spirit of skybox/uv_checker, six identical faces."""
    if name == "colors":
        flat = [(255, 0, 0), (0, 255, 255), (0, 255, 0), (255, 0, 255), (0, 0, 255), (255, 255, 0)]
        return SkyboxData([np.tile(np.array(c, dtype=np.uint8), (dim, dim, 1)) for c in flat])
    if name == "uv_gradient":
        col = (np.arange(dim) * 256 // dim).astype(np.uint8)
        img = np.zeros((dim, dim, 3), dtype=np.uint8)
        img[:, :, 0] = col[None, :]
        img[:, :, 1] = col[:, None]
        img[:, :, 2] = (((np.arange(dim)[None, :] // 32) + (np.arange(dim)[:, None] // 32)) & 1) * 255
        return SkyboxData([img.copy() for _ in range(6)])
    if name != "milky_way":
        raise ValueError(name)
    planes = []
    with np.errstate(over="ignore"):
        for f in range(6):
            idx = np.arange(dim * dim, dtype=np.uint64) + np.uint64(f) * np.uint64(dim * dim) + (np.uint64(seed) << np.uint64(40))
            h = _splitmix64(idx)
            star = (h % np.uint64(1000)) < np.uint64(3)
            bright = ((h >> np.uint64(16)) % np.uint64(200) + np.uint64(56)).astype(np.uint8)
            tint = ((h >> np.uint64(32)) % np.uint64(3)).astype(np.uint8)
            img = np.zeros((dim * dim, 3), dtype=np.uint8)
            for ch in range(3):
                img[:, ch] = np.where(star, np.where(tint == ch, bright, (bright // np.uint8(4)) * np.uint8(3)), 0)
            # band: brightness falls off with |v - 0.5|, integer arithmetic only
            v = (np.arange(dim * dim, dtype=np.int64) // dim) * 1024 // dim - 512
            glow = np.clip(40 - (v * v) // 1500, 0, 40).astype(np.uint8)
            if f in (0, 1, 4, 5):
                img = np.minimum(img.astype(np.int32) + glow[:, None].astype(np.int32), 255).astype(np.uint8)
            planes.append(img.reshape(dim, dim, 3))
    return SkyboxData(planes)


def get_skybox(name, dim=None):
    """Resolve a skybox by name: <repo>/skybox/<name>/ if it exists on disk, else the synthetic one."""
    d = os.path.join(REPO_ROOT, "skybox", name)
    if os.path.isdir(d) and all(os.path.exists(os.path.join(d, n)) for n in FACE_FILES):
        return load_skybox_dir(d)
    return synthetic_cubemap(name, dim or (1024 if name == "milky_way" else 256))


class SceneData:
    """A trt_Scene plus the arrays it points to (kept alive here)."""

    def __init__(self, width, height, skybox, kind="demo", num_spheres=1024):
        L = _lib.load()
        self.width, self.height = width, height
        self.skybox = skybox
        self.dl = abi.DirectionalLight()
        self.pl = abi.PointLight()
        self.c = abi.Scene()
        self.c.skybox = skybox.c
        if kind == "demo":
            self.spheres = (abi.Sphere * abi.DEMO_SPHERES)()
            L.trt_demo_scene(C.byref(self.c), self.spheres, C.byref(self.dl), C.byref(self.pl), width, height)
        elif kind == "stress":
            self.spheres = (abi.Sphere * num_spheres)()
            L.trt_stress_scene(C.byref(self.c), self.spheres, num_spheres, C.byref(self.dl), C.byref(self.pl), width, height)
        else:
            raise ValueError(kind)

    def set_time(self, t):
        """Camera pose of the reference's frame loop at wall-clock time t (TRT.c:1327-1336)."""
        _lib.load().trt_orbit_camera(C.byref(self.c.camera), float(t))
        return self
