"""Seeded random scenes for parity tests: triangle soups, spheres (also enclosing ones, so rays
start inside), and materials that exercise every shading branch (fixed reflectivity, Fresnel
with ior != 1, glossy cones, emitters, zero-albedo).  The camera block is built like the
reference's Camera constructor (src/math/Camera.h:40-51) so the scene can also be handed to
oracle/_ref/ref_tool as a PTSCENE2 file."""
import numpy as np

from pt_three_ways_b200 import scenefile


def camera18(eye, look_at, up, width, height, fov_degrees, focus=None, aperture=0.0):
    eye, look_at, up = (np.asarray(v, dtype=np.float64) for v in (eye, look_at, up))
    z = look_at - eye
    z = z / np.sqrt(z @ z)
    x = np.cross(up / np.sqrt(up @ up), z)
    x = x / np.sqrt(x @ x)
    y = np.cross(z, x)
    cam = np.zeros(18)
    cam[0:3], cam[3:6], cam[6:9], cam[9:12] = eye, x, y, z
    cam[12] = float(width) / height
    cam[13] = 1.0 / np.tan(fov_degrees * np.pi / 360.0)
    cam[14], cam[15] = 1.0 / height, 1.0 / width
    if focus is not None:
        d = np.asarray(focus, dtype=np.float64) - eye
        cam[16], cam[17] = aperture, np.sqrt(d @ d)
    return cam


def random_scene(seed, num_triangles=30, num_spheres=4, enclose=True):
    rng = np.random.default_rng(seed)
    materials = [
        [0, 0, 0, 0.7, 0.7, 0.7, 1.0, -1.0, 0.0],            # plain diffuse (Fresnel with ior 1)
        [0, 0, 0, 0.6, 0.2, 0.2, 1.3, -1.0, 0.6],            # Fresnel, wide cone
        [0, 0, 0, 0.9, 0.9, 0.9, 1.0, 0.8, 0.05],            # fixed reflectivity, tight cone
        [0, 0, 0, 0.9, 0.9, 0.9, 1.5, 0.5, 0.0],             # mirror (cone < Epsilon)
        [6, 5, 4, 0, 0, 0, 1.0, -1.0, 0.0],                  # light, zero albedo
        [0.5, 0.5, 0.5, 0.4, 0.5, 0.6, 1.1, -1.0, 2.5],      # emissive and diffuse, huge cone
    ]
    triangles = []
    for _ in range(num_triangles):
        centre = rng.uniform(-2, 2, 3)
        a, b, c = (centre + rng.normal(scale=0.8, size=3) for _ in range(3))
        triangles.append((a, b, c, int(rng.integers(0, len(materials)))))
    # a quad made of two coplanar triangles sharing an edge (the Cornell situation)
    triangles.append(((-3, -2.5, -3), (3, -2.5, -3), (3, -2.5, 3), 0))
    triangles.append(((-3, -2.5, -3), (3, -2.5, 3), (-3, -2.5, 3), 0))
    spheres = [(rng.uniform(-1.5, 1.5, 3), float(rng.uniform(0.2, 0.8)), int(rng.integers(0, len(materials))))
               for _ in range(num_spheres)]
    if enclose:
        spheres.append(((0.0, 0.0, 0.0), 9.0, 0))  # everything happens inside this one
    cam = camera18((0.3, 0.4, -5.0), (0, 0, 0), (0, 1, 0), 64, 48, 45.0, focus=(0, 0, 0), aperture=0.05)
    return scenefile.make(triangles=triangles, spheres=spheres, materials=materials,
                          environment=(0.05, 0.06, 0.08), camera64x48=cam)


def rays_for(scene, n, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-3, 3, size=(n, 3))
    k = n // 3
    tri = scene.triangle_vertices[rng.integers(0, scene.num_triangles, k)].reshape(k, 3, 3)
    a, b = rng.uniform(size=(2, k))
    flip = a + b > 1
    a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
    o[:k] = tri[:, 0] + a[:, None] * (tri[:, 1] - tri[:, 0]) + b[:, None] * (tri[:, 2] - tri[:, 0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)
