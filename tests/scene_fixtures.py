"""Small hand-made scenes for tests (built through ctl_scene_create_from_mesh)."""
import numpy as np

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import Material


def _mat(bsdf=0, refl=(0.7, 0.7, 0.7), two_sided=False, distr=0, alpha=0.2, eta=(1.5, 1.5, 1.5), k=(0, 0, 0)):
    m = Material(); m.bsdf_type = bsdf; m.flags = 1 if two_sided else 0; m.node_light_index = 0xffffffff; m.distr_type = distr
    m.reflectance[:] = refl; m.alpha_u = m.alpha_v = alpha; m.eta[:] = eta; m.k[:] = k; m.transmittance = 1.0
    return m


TWO_LIGHT_CAMERA = ((0, 0, -0.95), (0, -0.1, 0.5), (0, 1, 0), 80.0)


def two_light_room(w, h):
    """Closed box [-1,1]^3 with TWO area lights of different radiance and size (exercises the light-selection CDF,
    pdfEmitter and per-light area CDFs), a two-sided diffuse quad floating in the middle, a GGX conductor block."""
    V, I, M, mats, emissive = two_light_room_arrays()
    return ctl.Scene.from_mesh(V, I, M, mats, emissive, *TWO_LIGHT_CAMERA, w, h)


def two_light_room_arrays():
    """(verts (nv, 3) f32, indices (3 nt) u32, material index per triangle u8, materials, emissive (nm, 3))."""
    V, I, M = [], [], []

    def quad(a, b, c, d, mat):
        i0 = len(V); V.extend([a, b, c, d]); I.extend([(i0, i0 + 1, i0 + 2), (i0, i0 + 2, i0 + 3)]); M.extend([mat, mat])
    s = 1.0
    quad((-s, -s, -s), (s, -s, -s), (s, -s, s), (-s, -s, s), 0)          # floor  (normal +y)
    quad((-s, s, -s), (-s, s, s), (s, s, s), (s, s, -s), 0)              # ceiling (normal -y)
    quad((-s, -s, s), (s, -s, s), (s, s, s), (-s, s, s), 0)              # back
    quad((-s, -s, -s), (-s, -s, s), (-s, s, s), (-s, s, -s), 1)          # left (red)
    quad((s, -s, -s), (s, s, -s), (s, s, s), (s, -s, s), 2)              # right (green)
    quad((-0.6, 0.99, -0.2), (-0.6, 0.99, 0.2), (-0.2, 0.99, 0.2), (-0.2, 0.99, -0.2), 3)   # light A (small, bright), facing down
    quad((0.1, 0.99, -0.5), (0.1, 0.99, 0.5), (0.8, 0.99, 0.5), (0.8, 0.99, -0.5), 4)       # light B (large, dim)
    quad((-0.5, -0.2, -0.3), (0.1, -0.2, -0.3), (0.1, -0.2, 0.4), (-0.5, -0.2, 0.4), 5)     # two-sided diffuse card
    # conductor block (5 faces)
    b0, b1 = (0.3, -1.0, -0.2), (0.8, -0.4, 0.5)
    x0, y0, z0 = b0; x1, y1, z1 = b1
    quad((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0), 6)
    quad((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0), 6)
    quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1), 6)
    quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0), 6)
    quad((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1), 6)
    mats = [_mat(refl=(0.73, 0.73, 0.73)), _mat(refl=(0.63, 0.065, 0.05)), _mat(refl=(0.14, 0.45, 0.091)), _mat(refl=(0.78, 0.78, 0.78)), _mat(refl=(0.78, 0.78, 0.78)),
            _mat(refl=(0.3, 0.4, 0.8), two_sided=True), _mat(bsdf=1, distr=1, alpha=0.25, refl=(1, 1, 1), eta=(0.2, 0.924, 1.102), k=(3.912, 2.452, 2.142))]
    emissive = np.zeros((7, 3), np.float32); emissive[3] = (30, 26, 20); emissive[4] = (2, 3, 5)
    return np.array(V, np.float32), np.array(I, np.uint32).ravel(), np.array(M, np.uint8), mats, emissive
