"""PTSCENE2 scene fixtures: flat SoA arrays + the recipe camera.

The fixtures under tests/golden/scenes/ were produced by oracle/_ref/ref_tool (the reference's
own OBJ/MTL loader and scene recipes, src/main/main.cpp:69-309) and hold exactly what
dod::Scene::addTriangle/addSphere/setEnvironmentColour received, so the GPU box (which has no
/root/reference) can render the reference's scenes.  Layout (little-endian):

    char[8]  "PTSCENE2"
    u32      numTriangles, numSpheres, numMaterials, 0
    f64[3]   environment
    f64[T*9] triangle vertices v0 v1 v2
    u32[T]   triangle material index
    f64[S*4] sphere centre + radius
    u32[S]   sphere material index
    f64[M*9] materials {emission, diffuse, ior, reflectivity, coneAngle}
    f64[18]  camera state (src/math/Camera.h:11-18) for a 64x48 image
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

MAGIC = b"PTSCENE2"


@dataclasses.dataclass
class SceneArrays:
    triangle_vertices: np.ndarray  # (T, 9) f64
    triangle_material: np.ndarray  # (T,) u32
    sphere_centre_radius: np.ndarray  # (S, 4) f64
    sphere_material: np.ndarray  # (S,) u32
    materials: np.ndarray  # (M, 9) f64
    environment: np.ndarray  # (3,) f64
    camera64x48: np.ndarray | None = None  # (18,) f64

    @property
    def num_triangles(self) -> int:
        return int(self.triangle_material.shape[0])

    @property
    def num_spheres(self) -> int:
        return int(self.sphere_material.shape[0])

    def camera(self, width: int, height: int) -> np.ndarray:
        """The recipe camera for another image size: only aspectRatio_, reciprocalHeight_ and
        reciprocalWidth_ depend on it (Camera.h:43,46)."""
        if self.camera64x48 is None:
            raise ValueError("scene has no camera block")
        cam = self.camera64x48.copy()
        cam[12] = float(width) / height
        cam[14] = 1.0 / height
        cam[15] = 1.0 / width
        return cam

    def sweep_bytes(self) -> int:
        """Algorithmic bytes one ray cast sweeps: 72 B/triangle + 32 B/sphere."""
        return 72 * self.num_triangles + 32 * self.num_spheres

    def sweep_flops(self) -> int:
        """Unconditional fp64 flops of one ray cast: 46/triangle + 16/sphere."""
        return 46 * self.num_triangles + 16 * self.num_spheres


def load(path: str | os.PathLike) -> SceneArrays:
    data = open(path, "rb").read()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not a PTSCENE2 file")
    t, s, m, _ = np.frombuffer(data, dtype="<u4", count=4, offset=8)
    off = 24
    def take(dtype, count):
        nonlocal off
        arr = np.frombuffer(data, dtype=dtype, count=count, offset=off).copy()
        off += arr.nbytes
        return arr
    env = take("<f8", 3)
    tri = take("<f8", int(t) * 9).reshape(int(t), 9)
    tri_mat = take("<u4", int(t))
    sph = take("<f8", int(s) * 4).reshape(int(s), 4)
    sph_mat = take("<u4", int(s))
    mats = take("<f8", int(m) * 9).reshape(int(m), 9)
    cam = take("<f8", 18) if len(data) - off >= 144 else None
    return SceneArrays(tri, tri_mat, sph, sph_mat, mats, env, cam)


def save(scene: SceneArrays, path: str | os.PathLike) -> None:
    with open(path, "wb") as out:
        out.write(MAGIC)
        out.write(np.array([scene.num_triangles, scene.num_spheres, scene.materials.shape[0], 0],
                           dtype="<u4").tobytes())
        out.write(np.asarray(scene.environment, dtype="<f8").tobytes())
        out.write(np.ascontiguousarray(scene.triangle_vertices, dtype="<f8").tobytes())
        out.write(np.ascontiguousarray(scene.triangle_material, dtype="<u4").tobytes())
        out.write(np.ascontiguousarray(scene.sphere_centre_radius, dtype="<f8").tobytes())
        out.write(np.ascontiguousarray(scene.sphere_material, dtype="<u4").tobytes())
        out.write(np.ascontiguousarray(scene.materials, dtype="<f8").tobytes())
        cam = scene.camera64x48 if scene.camera64x48 is not None else np.zeros(18)
        out.write(np.asarray(cam, dtype="<f8").tobytes())


def make(triangles=(), spheres=(), materials=None, environment=(0.0, 0.0, 0.0),
         camera64x48=None) -> SceneArrays:
    """Builds a scene from python lists.  triangles: [(v0, v1, v2, materialIndex)],
    spheres: [(centre, radius, materialIndex)], materials: rows of 9 doubles."""
    if materials is None:
        materials = [[0, 0, 0, 0, 0, 0, 1.0, -1.0, 0.0]]  # MaterialSpec{} defaults
    tri = np.array([list(a) + list(b) + list(c) for a, b, c, _ in triangles],
                   dtype=np.float64).reshape(-1, 9)
    tri_mat = np.array([m for *_, m in triangles], dtype=np.uint32)
    sph = np.array([list(c) + [r] for c, r, _ in spheres], dtype=np.float64).reshape(-1, 4)
    sph_mat = np.array([m for *_, m in spheres], dtype=np.uint32)
    return SceneArrays(tri, tri_mat, sph, sph_mat,
                       np.array(materials, dtype=np.float64).reshape(-1, 9),
                       np.array(environment, dtype=np.float64),
                       None if camera64x48 is None else np.array(camera64x48, dtype=np.float64))
