"""Synthetic cube mesh of tetrahedra with the reference's connectivity conventions (host logic, numpy).

BASELINE.json configs[4] asks for a "synthetic CubeGenerator mesh".  The reference's own cube generator is
disabled at this revision (src/Initializer/InitProcedure/InitMesh.cpp:331-341 logs "not supported"), so the mesh
is built here: nx x ny x nz unit cubes, each split into 5 tetrahedra with alternating parity so that the face
diagonals of adjacent cubes match (conforming mesh).  What the element-update path needs from a mesh is exactly
what `MeshReader`/`PUMLReader` hand to the LTS layout (src/Initializer/InitProcedure/InitLayout.cpp:230-310):

  neighbors[e][f]      element behind face f (or -1 at the domain boundary),
  face_rel[e][f][0]    the neighbor's local face id (Element::neighborSides),
  face_rel[e][f][1]    the side orientation (Element::sideOrientations),

with the reference's local numbering: face f has the vertices FACE2NODES[f] (src/Geometry/MeshTools.cpp:21) and
the orientation is FaceVertexToOrientation[neighborSide][position of my face's first vertex among the
neighbor's vertices] (src/Geometry/PUMLReader.cpp:150-163, 416-426).
"""
import numpy as np

FACE2NODES = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]])
FIRST_FACE_VERTEX = np.array([0, 0, 0, 1])
FACE_VERTEX_TO_ORIENTATION = np.array([[0, 2, 1, -1], [0, 1, -1, 2], [0, -1, 2, 1], [-1, 0, 1, 2]])

# 5-tetrahedron split of the unit cube (corner = bit pattern x | y<<1 | z<<2): four corner tetrahedra and the
# central one.  Odd-parity cubes use the mirrored split (x -> 1-x), which makes shared cube faces conforming.
def _positive(tets):
    """Swap the last two vertices of every tetrahedron with negative volume."""
    c = np.stack([tets & 1, (tets >> 1) & 1, (tets >> 2) & 1], -1).astype(np.float64)
    neg = np.linalg.det(c[:, 1:] - c[:, :1]) < 0
    out = tets.copy()
    out[neg] = tets[neg][:, [0, 1, 3, 2]]
    return out


_EVEN = _positive(np.array([[0, 1, 2, 4], [3, 1, 2, 7], [5, 1, 4, 7], [6, 2, 4, 7], [1, 2, 4, 7]]))
_ODD = _positive(_EVEN ^ 1)
# the 12 even permutations of 4 vertices: a per-element choice (hash of the global id) varies the local
# numbering, so that all 4 x 3 (neighbor face, orientation) relations occur as in an unstructured mesh
_A4 = np.array([p for p in __import__("itertools").permutations(range(4))
                if sum(p[i] > p[j] for i in range(4) for j in range(i + 1, 4)) % 2 == 0])


class CubeMesh:
    """Tetrahedra of the cubes with global x index in [x0, x1) (a slab of the global nx x ny x nz cube grid).

    Periodic in x (ranks form a ring); y and z end at the domain boundary.  Element ids are global:
    gid = 5 * ((ix * ny + iy) * nz + iz) + t, so that every rank derives the same data for the same element."""

    def __init__(self, nx, ny, nz, x0=0, x1=None, periodic_yz=False):
        if nx % 2:
            raise ValueError("nx must be even (parity split + periodicity in x)")
        if periodic_yz and (ny % 2 or nz % 2 or min(nx, ny, nz) < 4):
            raise ValueError("a fully periodic mesh needs even ny, nz and at least 4 cubes per direction")
        self.nx, self.ny, self.nz = nx, ny, nz
        self.periodic_yz = periodic_yz
        x1 = nx if x1 is None else x1
        self.x0, self.x1 = x0, x1
        ix = np.arange(x0, x1)
        gx, gy, gz = np.meshgrid(ix % nx, np.arange(ny), np.arange(nz), indexing="ij")
        gx, gy, gz = gx.ravel(), gy.ravel(), gz.ravel()
        parity = (gx + gy + gz) % 2
        corners = np.where(parity[:, None, None] == 0, _EVEN[None], _ODD[None])        # [cubes, 5, 4]
        ux = gx[:, None, None] + (corners & 1)
        uy = gy[:, None, None] + ((corners >> 1) & 1)
        uz = gz[:, None, None] + ((corners >> 2) & 1)
        cx = ux % nx
        py, pz = (ny, nz) if periodic_yz else (ny + 1, nz + 1)
        cy, cz = uy % py, uz % pz
        vertices = ((cx * py + cy) * pz + cz).reshape(-1, 4).astype(np.int64)
        coords = np.stack([ux, uy, uz], -1).reshape(-1, 4, 3).astype(np.float64)      # unwrapped vertex coordinates
        cube = (gx * ny + gy) * nz + gz
        self.gid = (5 * cube[:, None] + np.arange(5)[None]).reshape(-1).astype(np.int64)
        perm = _A4[(self.gid * 2654435761 >> 7) % 12]
        self.vertices = np.take_along_axis(vertices, perm, axis=1)
        self.coords = np.take_along_axis(coords, perm[:, :, None], axis=1)   # [n, 4, 3] in the element's local numbering
        self.cube_x = np.repeat(gx, 5)
        self.cube_z = np.repeat(gz, 5)
        self.n = self.vertices.shape[0]
        self._connect()

    def vertex_coords(self):
        """[n, 4, 3] vertex coordinates (unit cubes); x of wrapped vertices is not unwrapped."""
        v = self.vertices
        z = v % (self.nz + 1)
        y = (v // (self.nz + 1)) % (self.ny + 1)
        x = v // ((self.nz + 1) * (self.ny + 1))
        return np.stack([x, y, z], -1).astype(np.float64)

    def _connect(self):
        n = self.n
        fv = self.vertices[:, FACE2NODES]                       # [n, 4, 3] face vertices, reference order
        key = np.sort(fv.reshape(-1, 3), axis=1)
        order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
        ks = key[order]
        same = np.all(ks[1:] == ks[:-1], axis=1)
        # a face is shared by at most two tetrahedra: matched pairs are adjacent after sorting
        first = np.nonzero(same)[0]
        a, b = order[first], order[first + 1]
        nb = np.full(4 * n, -1, dtype=np.int64)
        nb_side = np.zeros(4 * n, dtype=np.int64)
        nb[a], nb[b] = b // 4, a // 4
        nb_side[a], nb_side[b] = b % 4, a % 4
        self.neighbors = nb.reshape(n, 4)                       # local element index or -1
        nb_side = nb_side.reshape(n, 4)
        # orientation: position of my face's first vertex in the neighbor's vertex list
        first_vertex = self.vertices[:, FIRST_FACE_VERTEX]      # [n, 4]
        nbv = self.vertices[np.maximum(self.neighbors, 0)]      # [n, 4, 4]
        pos = np.argmax(nbv == first_vertex[:, :, None], axis=2)
        orient = FACE_VERTEX_TO_ORIENTATION[nb_side, pos]
        has = self.neighbors >= 0
        assert np.all(orient[has] >= 0)
        rel = np.zeros((n, 4, 2), dtype=np.uint8)
        rel[:, :, 0] = np.where(has, nb_side, 0)
        rel[:, :, 1] = np.where(has, orient, 0)
        self.face_rel = rel


def element_values(gid, count, stream, scale=1.0, dtype=np.float32):
    """Deterministic bounded pseudo-random values in [-scale, scale) per (global element id, entry index):
    a 64-bit mix hash, so that any partition of the mesh sees identical element data."""
    with np.errstate(over="ignore"):
        h = (gid.astype(np.uint64)[:, None] * np.uint64(0x9E3779B97F4A7C15)
             + (np.arange(count, dtype=np.uint64)[None, :] + np.uint64(stream) * np.uint64(1 << 20))
             * np.uint64(0xC2B2AE3D27D4EB4F))
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xFF51AFD7ED558CCD)
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xC4CEB9FE1A85EC53)
        h ^= h >> np.uint64(33)
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 24)     # [0, 1)
    return ((2.0 * u - 1.0) * scale).astype(dtype)
