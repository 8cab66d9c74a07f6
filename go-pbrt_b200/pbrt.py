"""Host-side mirror of the reference's Go interfaces for the hot path (the Go toolchain is absent in this image,
SURVEY.md §0.1, so this Python layer is the testable host; the cgo shim a maintainer would add is in go/ and
INTEGRATION.md).  Same names, argument meaning and error behaviour as the Go constructors it mirrors; every
function cites the reference file:line.  Host-side quirks of the reference (Matrix4x4.Mul, Transform.Mul, the
shutter) are reproduced because they define the matrices the library receives.

Nothing here computes an intersection or a radiance value: all of that happens in libgopbrt_cuda.so.
"""
import ctypes as C
import math

import numpy as np

from . import abi
from . import gomath

# --------------------------------------------------------------------------- pkg/pbrt/transform.go


class Matrix4x4:
    def __init__(self, rows=None):
        self.m = [list(map(float, r)) for r in rows] if rows is not None else [
            [1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]]

    def Transpose(self):  # transform.go:53-60
        m = self.m
        return Matrix4x4([[m[j][i] for j in range(4)] for i in range(4)])

    def Mul(self, other):  # transform.go:62-70 — last term uses m[3][j], not other[3][j] (Q7)
        m, o = self.m, other.m
        r = [[0.0] * 4 for _ in range(4)]
        for i in range(4):
            for j in range(4):
                r[i][j] = m[i][0] * o[0][j] + m[i][1] * o[1][j] + m[i][2] * o[2][j] + m[i][3] * m[3][j]
        return Matrix4x4(r)

    def Inverse(self):  # transform.go:72-142 (Gauss-Jordan with full pivoting)
        indxc, indxr, ipiv = [0] * 4, [0] * 4, [0] * 4
        minv = [row[:] for row in self.m]
        for i in range(4):
            irow = icol = 0
            big = 0.0
            for j in range(4):
                if ipiv[j] != 1:
                    for k in range(4):
                        if ipiv[k] == 0:
                            if abs(minv[j][k]) >= big:
                                big = abs(minv[j][k])
                                irow, icol = j, k
                        elif ipiv[k] > 1:
                            raise ValueError("singular Matrix in Matrix invert")
            ipiv[icol] += 1
            if irow != icol:
                minv[irow], minv[icol] = minv[icol], minv[irow]
            indxr[i], indxc[i] = irow, icol
            if minv[icol][icol] == 0.0:
                raise ValueError("singular Matrix in Matrix invert")
            pivinv = 1.0 / minv[icol][icol]
            minv[icol][icol] = 1.0
            for j in range(4):
                minv[icol][j] *= pivinv
            for j in range(4):
                if j != icol:
                    save = minv[j][icol]
                    minv[j][icol] = 0.0
                    for k in range(4):
                        minv[j][k] -= minv[icol][k] * save
        for j in range(3, -1, -1):
            if indxr[j] != indxc[j]:
                for k in range(4):
                    minv[k][indxr[j]], minv[k][indxc[j]] = minv[k][indxc[j]], minv[k][indxr[j]]
        return Matrix4x4(minv)

    def flat(self):
        return [v for r in self.m for v in r]


class Transform:
    def __init__(self, Matrix, MatrixInverse):
        self.Matrix, self.MatrixInverse = Matrix, MatrixInverse

    def Inverse(self):  # transform.go:175-177
        return Transform(self.MatrixInverse, self.Matrix)

    def Mul(self, other):  # transform.go:179-184 — inverses multiplied in the SAME order (Q7b)
        return Transform(self.Matrix.Mul(other.Matrix), self.MatrixInverse.Mul(other.MatrixInverse))

    def IsIdentity(self):
        return self.Matrix.m == Matrix4x4().m

    def TransformPoint(self, p):  # transform.go:227-247 (value only)
        m = self.Matrix.m
        x, y, z = p
        xp = m[0][0] * x + m[0][1] * y + m[0][2] * z + m[0][3]
        yp = m[1][0] * x + m[1][1] * y + m[1][2] * z + m[1][3]
        zp = m[2][0] * x + m[2][1] * y + m[2][2] * z + m[2][3]
        wp = m[3][0] * x + m[3][1] * y + m[3][2] * z + m[3][3]
        if wp == 1.0:
            return (xp, yp, zp)
        return (xp / wp, yp / wp, zp / wp)

    def TransformVector(self, v):  # transform.go:249-255
        m = self.Matrix.m
        x, y, z = v
        return (m[0][0] * x + m[0][1] * y + m[0][2] * z, m[1][0] * x + m[1][1] * y + m[1][2] * z,
                m[2][0] * x + m[2][1] * y + m[2][2] * z)

    def abi(self):
        t = abi.Transform()
        t.m[:] = self.Matrix.flat()
        t.minv[:] = self.MatrixInverse.flat()
        return t


def NewTransform(m):  # transform.go:148-157
    return Transform(m, m.Inverse())


def Translate(delta):  # transform.go:347-362
    x, y, z = delta
    return Transform(Matrix4x4([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]]),
                     Matrix4x4([[1, 0, 0, -x], [0, 1, 0, -y], [0, 0, 1, -z], [0, 0, 0, 1]]))


def Scale(x, y, z):  # transform.go:364-379
    return Transform(Matrix4x4([[x, 0, 0, 0], [0, y, 0, 0], [0, 0, z, 0], [0, 0, 0, 1]]),
                     Matrix4x4([[1.0 / x, 0, 0, 0], [0, 1.0 / y, 0, 0], [0, 0, 1.0 / z, 0], [0, 0, 0, 1]]))


def RotateX(degrees):  # transform.go:381-394 — Go's own sin/cos (gomath), not libm
    s, c = gomath.Sin(gomath.Radians(degrees)), gomath.Cos(gomath.Radians(degrees))
    m = Matrix4x4([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]])
    return Transform(m, m.Transpose())


def RotateY(degrees):  # transform.go:396-409
    s, c = gomath.Sin(gomath.Radians(degrees)), gomath.Cos(gomath.Radians(degrees))
    m = Matrix4x4([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]])
    return Transform(m, m.Transpose())


def RotateZ(degrees):  # transform.go:411-424
    s, c = gomath.Sin(gomath.Radians(degrees)), gomath.Cos(gomath.Radians(degrees))
    m = Matrix4x4([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    return Transform(m, m.Transpose())


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _cross(a, b):  # xyz.go:583-585
    return ((a[1] * b[2]) - (a[2] * b[1]), (a[2] * b[0]) - (a[0] * b[2]), (a[0] * b[1]) - (a[1] * b[0]))


def _normalized(a):  # xyz.go:597-606
    n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2]
    if n2 > 0:
        inv = 1.0 / math.sqrt(n2)
        return (a[0] * inv, a[1] * inv, a[2] * inv)
    return tuple(a)


def LookAt(pos, look, up):  # transform.go:453-486
    c2w = [[0.0] * 4 for _ in range(4)]
    c2w[0][3], c2w[1][3], c2w[2][3], c2w[3][3] = pos[0], pos[1], pos[2], 1.0
    d = _normalized(_sub(look, pos))
    cr = _cross(_normalized(up), d)
    if math.sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]) == 0:
        raise ValueError("same direction in Matrix invert")
    right = _normalized(cr)
    new_up = _cross(d, right)
    for i in range(3):
        c2w[i][0], c2w[i][1], c2w[i][2] = right[i], new_up[i], d[i]
    m = Matrix4x4(c2w)
    return Transform(m, m.Inverse())


def Perspective(fov, n, f):  # transform.go:492-502
    persp = Matrix4x4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, f / (f - n), -f * n / (f - n)], [0, 0, 1, 0]])
    inv_tan = 1.0 / gomath.Tan(gomath.Radians(fov) / 2)
    return Scale(inv_tan, inv_tan, 1).Mul(NewTransform(persp))


class AnimatedTransform:  # transform.go:512-548 — only the static case exists in the reference (Decompose is a TODO)
    def __init__(self, start, end, startTime, endTime):
        if start is not end:
            raise NotImplementedError("actuallyAnimated transforms nil-deref in the reference (transform.go:536-541)")
        self.startTransform, self.startTime, self.endTime = start, startTime, endTime


def NewAnimatedTransform(start, end, startTime, endTime):
    return AnimatedTransform(start, end, startTime, endTime)


# --------------------------------------------------------------------------- spectra, textures, materials
def NewSpectrum(v):  # spectrum.go:45-54
    return [float(v)] * 3


def NewRGBSpectrum(r, g, b):
    return [float(r), float(g), float(b)]


class ConstantSpectrumTexture:  # texture.go:56-68
    def __init__(self, value):
        self.value = list(value)


class ConstantFloatTexture:  # texture.go:70-82
    def __init__(self, value):
        self.value = float(value)


NewConstantSpectrumTexture = ConstantSpectrumTexture
NewConstantFloatTexture = ConstantFloatTexture


class UVMapping2D:  # texture.go:9-26
    def __init__(self, su=1.0, sv=1.0, du=0.0, dv=0.0):
        self.su, self.sv, self.du, self.dv = su, sv, du, dv


class PlanarMapping2D:  # texture.go:28-46
    def __init__(self, vs, vt, ds, dt):
        self.vs, self.vt, self.ds, self.dt = tuple(vs), tuple(vt), ds, dt


NewUvMapping2D = UVMapping2D
NewPlanarMapping2D = PlanarMapping2D


class Checkerboard2D:  # checkerboard.go:14-28
    def __init__(self, mapping, tex1, tex2):
        self.mapping, self.tex1, self.tex2 = mapping, tex1, tex2


NewCheckerboard2D = Checkerboard2D


class MatteMaterial:  # matte.go:8-19
    def __init__(self, Kd, sigma, bumpMap=None):
        assert bumpMap is None, "Bump is a no-op stub in the reference (material.go:18-34)"
        self.Kd, self.sigma = Kd, sigma


class Mirror:  # mirror.go:9-19
    def __init__(self):
        self.Kr = ConstantSpectrumTexture(NewSpectrum(0.9))


class Glass:  # glass.go:7-25
    def __init__(self, Kr, Kt, uRoughness, vRoughness, index, bumpMap=None):
        assert bumpMap is None
        self.Kr, self.Kt, self.uRoughness, self.vRoughness, self.index = Kr, Kt, uRoughness, vRoughness, index


NewMatteMaterial = MatteMaterial
NewMirror = Mirror
NewGlass = Glass

# --------------------------------------------------------------------------- shapes and primitives


class Sphere:  # sphere.go:8-36
    kind = abi.SHAPE_SPHERE

    def __init__(self, name, objectToWorld, reverseOrientation, radius, zMin=None, zMax=None, phiMax=360.0):
        self.name, self.objectToWorld, self.reverseOrientation, self.radius = name, objectToWorld, reverseOrientation, radius
        self.zMin = -radius if zMin is None else zMin
        self.zMax = radius if zMax is None else zMax
        self.phiMax = phiMax


def NewSphereShape(name, o2w, reverseOrientation, radius):  # sphere.go:34-36
    return Sphere(name, o2w, reverseOrientation, radius)


class Disk:  # disk.go:14-35 (reverseOrientation hard-wired false)
    kind = abi.SHAPE_DISK

    def __init__(self, objectToWorld, height, radius, innerRadius, phiMax):
        self.objectToWorld, self.height, self.radius, self.innerRadius, self.phiMax = objectToWorld, height, radius, innerRadius, phiMax
        self.reverseOrientation = False


NewDisk = Disk


class GeometricPrimitive:  # primitive.go:22-36
    def __init__(self, shape, material):
        self.Shape, self.material = shape, material


class TransformedPrimitive:  # primitive.go:82-92
    def __init__(self, primitive, primitiveToWorld):
        self.primitive, self.primitiveToWorld = primitive, primitiveToWorld


NewGeometricPrimitive = GeometricPrimitive
NewTransformedPrimitive = TransformedPrimitive


class TriangleMesh:
    """New shape (not in the reference): world-space float64 vertices (N,3), int32 indices (M,3).
    Stands for M GeometricPrimitives sharing one material."""

    def __init__(self, vertices, indices, material, reverseOrientation=False):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        self.material, self.reverseOrientation = material, reverseOrientation


SplitSAH, SplitHLBVH, SplitMiddle, SplitEqualCounts = 1, 2, 3, 4  # bvh.go:16-21


class BVH:
    """accelerator.NewBVH(primitives, maxPrimsInNode, splitMethod) (bvh.go:223): on this backend the tree is built by the
    library inside gopbrt_scene_create (the reference's own builder is degenerate); splitMethod is recorded
    and ignored."""

    def __init__(self, primitives, maxPrimsInNode, splitMethod=SplitSAH):
        self.primitives, self.maxPrimsInNode, self.splitMethod = list(primitives), int(min(255, maxPrimsInNode)), splitMethod


NewBVH = BVH

# --------------------------------------------------------------------------- lights


class Distant:  # distant.go:8-31
    def __init__(self, lightToWorld, L, wLight):
        self.L = list(L)
        self.wLight = _normalized(lightToWorld.TransformVector(wLight))


class Point:  # point.go:8-31
    def __init__(self, lightToWorld, mediumAccessor, I):
        self.pLight = lightToWorld.TransformPoint((0.0, 0.0, 0.0))
        self.I = list(I)


class DiffuseAreaLight:  # diffuse.go:8-24
    def __init__(self, lightToWorld, mediumAccessor, LEmit, nSamples, shape, twoSided):
        self.LEmit, self.nSamples, self.shape, self.twoSided = list(LEmit), nSamples, shape, twoSided


NewDistant = Distant
NewPoint = Point
NewDiffuseAreaLight = DiffuseAreaLight

# --------------------------------------------------------------------------- scene flattening


class Scene:
    """pbrt.NewScene(aggregate, lights) (scene.go:16-36) → gopbrt_scene_desc."""

    def __init__(self, aggregate, lights):
        self.aggregate, self.lights = aggregate, list(lights)
        self._desc = None
        self._keep = None

    def desc(self):
        if self._desc is not None:
            return self._desc
        xf, xf_ids = [], {}
        spheres, disks, sh_ids = [], [], {}
        mats, mat_ids, texs, tex_ids = [], {}, [], {}
        prim_rows = []
        meshes = []

        def xf_id(t):
            if id(t) not in xf_ids:
                xf_ids[id(t)] = len(xf)
                xf.append(t.abi())
            return xf_ids[id(t)]

        def tex_id(t):
            if id(t) in tex_ids:
                return tex_ids[id(t)]
            r = abi.Texture()
            if isinstance(t, ConstantSpectrumTexture):
                r.kind = abi.TEX_CONSTANT
                r.rgb[:] = t.value
                r.tex1 = r.tex2 = -1
            elif isinstance(t, Checkerboard2D):
                r.kind = abi.TEX_CHECKERBOARD
                r.tex1, r.tex2 = tex_id(t.tex1), tex_id(t.tex2)
                if isinstance(t.mapping, PlanarMapping2D):
                    r.mapping = abi.MAP_PLANAR
                    r.vs[:], r.vt[:], r.ds, r.dt = t.mapping.vs, t.mapping.vt, t.mapping.ds, t.mapping.dt
                else:
                    r.mapping = abi.MAP_UV
                    r.su, r.sv, r.du, r.dv = t.mapping.su, t.mapping.sv, t.mapping.du, t.mapping.dv
            else:
                raise TypeError(f"unsupported texture {t!r}")
            tex_ids[id(t)] = len(texs)
            texs.append(r)
            return tex_ids[id(t)]

        def mat_id(m):
            if m is None:
                return -1
            if id(m) in mat_ids:
                return mat_ids[id(m)]
            r = abi.Material()
            r.tex_a = r.tex_b = -1
            if isinstance(m, MatteMaterial):
                r.kind, r.tex_a, r.sigma = abi.MAT_MATTE, tex_id(m.Kd), m.sigma.value
            elif isinstance(m, Mirror):
                r.kind, r.tex_a = abi.MAT_MIRROR, tex_id(m.Kr)
            elif isinstance(m, Glass):
                r.kind, r.tex_a, r.tex_b = abi.MAT_GLASS, tex_id(m.Kr), tex_id(m.Kt)
                r.eta, r.u_rough, r.v_rough = m.index.value, m.uRoughness.value, m.vRoughness.value
            else:
                raise TypeError(f"unsupported material {m!r}")
            mat_ids[id(m)] = len(mats)
            mats.append(r)
            return mat_ids[id(m)]

        def shape_id(s):
            if id(s) in sh_ids:
                return sh_ids[id(s)]
            if isinstance(s, Sphere):
                r = abi.Sphere(xf_id(s.objectToWorld), int(bool(s.reverseOrientation)), s.radius, s.zMin, s.zMax, s.phiMax)
                sh_ids[id(s)] = (abi.SHAPE_SPHERE, len(spheres))
                spheres.append(r)
            elif isinstance(s, Disk):
                r = abi.Disk(xf_id(s.objectToWorld), int(bool(s.reverseOrientation)), s.height, s.radius, s.innerRadius, s.phiMax)
                sh_ids[id(s)] = (abi.SHAPE_DISK, len(disks))
                disks.append(r)
            else:
                raise TypeError(f"unsupported shape {s!r}")
            return sh_ids[id(s)]

        n_tris = 0
        n_verts = 0
        for p in self.aggregate.primitives:
            if isinstance(p, TriangleMesh):
                meshes.append((p, len(prim_rows), n_tris, n_verts, mat_id(p.material)))
                prim_rows.append(("mesh", len(meshes) - 1))
                n_tris += len(p.indices)
                n_verts += len(p.vertices)
                continue
            p2w = -1
            if isinstance(p, TransformedPrimitive):
                p2w = xf_id(p.primitiveToWorld.startTransform)
                p = p.primitive
            k, i = shape_id(p.Shape)
            prim_rows.append((k, i, mat_id(p.material), p2w))

        # primitives array (meshes expand in place, preserving primitive order)
        n_prims = sum(len(meshes[r[1]][0].indices) if r[0] == "mesh" else 1 for r in prim_rows)
        prims = np.zeros((n_prims, 4), dtype=np.int32)
        verts = np.zeros((max(n_verts, 1), 3), dtype=np.float64)
        tris = np.zeros((max(n_tris, 1), 4), dtype=np.int32)
        o = 0
        for r in prim_rows:
            if r[0] == "mesh":
                mesh, _, t0, v0, mid = meshes[r[1]]
                nt = len(mesh.indices)
                verts[v0:v0 + len(mesh.vertices)] = mesh.vertices
                tris[t0:t0 + nt, :3] = mesh.indices + v0
                tris[t0:t0 + nt, 3] = int(bool(mesh.reverseOrientation))
                prims[o:o + nt, 0] = abi.SHAPE_TRIANGLE
                prims[o:o + nt, 1] = np.arange(t0, t0 + nt, dtype=np.int32)
                prims[o:o + nt, 2] = mid
                prims[o:o + nt, 3] = -1
                o += nt
            else:
                prims[o] = r
                o += 1

        lights = []
        for l in self.lights:
            r = abi.Light()
            r.shape_kind = r.shape_index = -1
            if isinstance(l, Distant):
                r.kind = abi.LIGHT_DISTANT
                r.rgb[:], r.v[:] = l.L, l.wLight
            elif isinstance(l, Point):
                r.kind = abi.LIGHT_POINT
                r.rgb[:], r.v[:] = l.I, l.pLight
            elif isinstance(l, DiffuseAreaLight):
                r.kind = abi.LIGHT_DIFFUSE_AREA
                r.rgb[:] = l.LEmit
                r.shape_kind, r.shape_index = shape_id(l.shape)
                r.two_sided = int(bool(l.twoSided))
            else:
                raise TypeError(f"unsupported light {l!r}")
            lights.append(r)

        def arr(T, items):
            a = (T * max(len(items), 1))()
            for i, it in enumerate(items):
                a[i] = it
            return a

        keep = dict(xf=arr(abi.Transform, xf), spheres=arr(abi.Sphere, spheres), disks=arr(abi.Disk, disks),
                    mats=arr(abi.Material, mats), texs=arr(abi.Texture, texs), lights=arr(abi.Light, lights),
                    prims=prims, verts=verts, tris=tris)
        d = abi.SceneDesc()
        d.n_transforms, d.transforms = len(xf), keep["xf"]
        d.n_spheres, d.spheres = len(spheres), keep["spheres"]
        d.n_disks, d.disks = len(disks), keep["disks"]
        d.n_vertices, d.vertices = n_verts, verts.ctypes.data_as(C.POINTER(C.c_double))
        d.n_triangles, d.triangles = n_tris, C.cast(tris.ctypes.data, C.POINTER(abi.Triangle))
        d.n_primitives, d.primitives = n_prims, C.cast(prims.ctypes.data, C.POINTER(abi.Primitive))
        d.n_materials, d.materials = len(mats), keep["mats"]
        d.n_textures, d.textures = len(texs), keep["texs"]
        d.n_lights, d.lights = len(lights), keep["lights"]
        d.max_prims_in_node = self.aggregate.maxPrimsInNode
        d.flags = 0
        self._desc, self._keep = d, keep
        return d


NewScene = Scene

# --------------------------------------------------------------------------- film / sampler / camera / integrator


class BoxFilter:  # filter.go:20-32
    def __init__(self, radius):
        self.radius = tuple(radius)


NewBoxFilter = BoxFilter


class Film:  # film.go:27-76
    def __init__(self, filename, resolution, cropWindow, filter, diagonal=100.0, scale=1.0, maxSampleLuminance=1.0):
        self.Filename, self.FullResolution, self.cropWindow, self.Filter = filename, tuple(resolution), tuple(cropWindow), filter
        W, H = self.FullResolution
        c = self.cropWindow
        self.CroppedPixelBounds = (int(math.ceil(W * c[0])), int(math.ceil(H * c[1])), int(math.ceil(W * c[2])), int(math.ceil(H * c[3])))
        self.pixels = None  # (H', W', 4) float64 after Render

    def abi(self):
        f = abi.Film()
        f.width, f.height = self.FullResolution
        f.crop[:] = self.cropWindow
        f.filter_radius[:] = self.Filter.radius
        return f

    def shape(self):
        x0, y0, x1, y1 = self.CroppedPixelBounds
        return (y1 - y0, x1 - x0, 4)

    def WriteImage(self):
        """Film.WriteImage (film.go:142-163) minus the PNG encode: uint8(clamp(sum XYZ, 0, 1) * 255), no division,
        no XYZ->RGB, no gamma.  Returns (H, W, 4) uint8 NRGBA."""
        v = np.clip(self.pixels[..., :3], 0.0, 1.0) * 255.0
        out = np.empty(self.pixels.shape[:2] + (4,), dtype=np.uint8)
        out[..., :3] = v.astype(np.uint8)  # Go uint8(float) truncates
        out[..., 3] = 255
        return out


NewFilm = Film


class Stratified:  # stratified.go:5-20
    def __init__(self, xSamples, ySamples, jitterSamples, nSampledDimensions):
        self.xSamples, self.ySamples, self.jitter, self.nDims = xSamples, ySamples, jitterSamples, nSampledDimensions

    def abi(self, mode):
        return abi.Sampler(abi.SAMPLER_STRATIFIED, self.xSamples, self.ySamples, int(bool(self.jitter)), self.nDims, mode)

    def GetSamplesPerPixel(self):
        return self.xSamples * self.ySamples


class RandomSampler:  # random.go:6-19
    def __init__(self, ns, seed=0):
        self.ns = ns

    def abi(self, mode):
        return abi.Sampler(abi.SAMPLER_RANDOM, self.ns, 1, 0, 0, mode)

    def GetSamplesPerPixel(self):
        return self.ns


NewStratified = Stratified
NewRandomSampler = RandomSampler


class PerspectiveCamera:  # camera.go:98-165
    def __init__(self, cameraToWorld, screenWindow, shutterOpen, shutterClose, lensRadius, focalDistance, fov, film, medium=None):
        W, H = film.FullResolution
        sw = screenWindow  # (min.x, min.y, max.x, max.y)
        cameraToScreen = Perspective(fov, 1e-2, 1000.0)
        # NewProjectiveCamera (camera.go:106-124)
        s2r = Scale(float(W), float(H), 1.0)
        s2r = s2r.Mul(Scale(1.0 / (sw[2] - sw[0]), 1.0 / (sw[1] - sw[3]), 1.0))
        s2r = s2r.Mul(Translate((-sw[0], -sw[3], 0.0)))
        r2s = s2r.Inverse()
        self.RasterToCamera = cameraToScreen.Inverse().Mul(r2s)
        self.cameraToWorld = cameraToWorld
        self.shutterOpen = shutterOpen
        self.shutterClose = shutterOpen  # camera.go:116 passes shutterOpen twice (Q21)
        self.lensRadius, self.focalDistance, self.Film = lensRadius, focalDistance, film

    def abi(self):
        c = abi.Camera()
        c.raster_to_camera[:] = self.RasterToCamera.Matrix.flat()
        c.camera_to_world[:] = self.cameraToWorld.startTransform.Matrix.flat()
        c.lens_radius, c.focal_distance = self.lensRadius, self.focalDistance
        c.shutter_open, c.shutter_close = self.shutterOpen, self.shutterClose
        return c

    def GetFilm(self):
        return self.Film


NewPerspectiveCamera = PerspectiveCamera

Uniform, Power, Spatial = 1, 2, 4  # lightdistribution.go:5-9


class Path:  # path.go:10-26
    def __init__(self, maxDepth, camera, sampler, pixelBounds, rrThreshold, lightSampleStrategy):
        # Uniform, Power (reproduced bugs included: no light is ever sampled, lightdistribution.go:44-68) or Spatial (nil
        # distribution in the reference: the library answers GOPBRT_ERR_UNSUPPORTED where path.go:80 panics)
        self.maxDepth, self.camera, self.sampler, self.rrThreshold = maxDepth, camera, sampler, rrThreshold
        self.lightSampleStrategy = lightSampleStrategy

    def GetCamera(self):
        return self.camera

    def GetSampler(self):
        return self.sampler

    def abi(self, tileSize):
        return abi.Integrator(0, self.maxDepth, self.rrThreshold, self.lightSampleStrategy, 0, tileSize)


NewPath = Path

UniformSampleAll, UniformSampleOne = 1, 2  # directlighting.go:12-15


class DirectLighting:  # directlighting.go:17-36
    """integrator.NewDirectLighting(strategy, maxDepth, camera, sampler, pixelBounds).  Preprocess (directlighting.go:46-60)
    requests sample arrays on the prototype sampler only; the per-tile clones never see them (pixel.go:34-42), so nothing
    of it reaches the hot path."""

    def __init__(self, strategy, maxDepth, camera, sampler, pixelBounds=None):
        if strategy not in (UniformSampleAll, UniformSampleOne):
            raise ValueError("unknown lighting strategy (directlighting.go:92-94 panics)")
        self.strategy, self.maxDepth, self.camera, self.sampler = strategy, maxDepth, camera, sampler

    def GetCamera(self):
        return self.camera

    def GetSampler(self):
        return self.sampler

    def abi(self, tileSize):
        return abi.Integrator(1, self.maxDepth, 0.0, self.strategy, 0, tileSize)


NewDirectLighting = DirectLighting


class Device:
    """gopbrt_ctx: one per process and GPU."""

    def __init__(self, device=0):
        self.lib = abi.load()
        self.h = C.c_void_p()
        rc = self.lib.gopbrt_init(device, C.byref(self.h))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_init({device}) failed rc={rc} (no CPU fallback)")

    def error(self):
        e = self.lib.gopbrt_last_error(self.h)
        return e.decode() if e else ""

    def launches(self):
        return int(self.lib.gopbrt_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.gopbrt_shutdown(self.h)
            self.h = C.c_void_p()

    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id made on rank 0 and handed to every rank by any host channel (gopbrt_comm_unique_id)."""
        buf = C.create_string_buffer(abi.COMM_ID_BYTES)
        rc = abi.load().gopbrt_comm_unique_id(buf)
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_comm_unique_id rc={rc} (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, comm_id, rank, world):
        """joins this context to the film-reduce communicator (gopbrt_comm_init_rank); collective over the ranks"""
        rc = self.lib.gopbrt_comm_init_rank(self.h, bytes(comm_id), rank, world)
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_comm_init_rank rc={rc}: {self.error()}")


class GpuScene:
    """gopbrt_scene: the uploaded scene (BVH built and flattened once)."""

    def __init__(self, dev, scene):
        self.dev, self.scene = dev, scene
        self.h = C.c_void_p()
        rc = dev.lib.gopbrt_scene_create(dev.h, C.byref(scene.desc()), C.byref(self.h))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_scene_create failed rc={rc}: {dev.error()}")

    def close(self):
        if self.h:
            self.dev.lib.gopbrt_scene_destroy(self.h)
            self.h = C.c_void_p()

    def WorldBound(self):
        out = (C.c_double * 6)()
        self.dev.lib.gopbrt_scene_world_bound(self.h, out)
        return list(out)

    @staticmethod
    def _rays(o, d, tmax):
        o = np.ascontiguousarray(o, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1, 3)
        n = len(o)
        cols = [np.ascontiguousarray(o[:, i]) for i in range(3)] + [np.ascontiguousarray(d[:, i]) for i in range(3)]
        tm = np.full(n, np.inf) if tmax is None else np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
        cols.append(tm)
        return n, cols

    def Intersect(self, o, d, tmax=None):
        """Aggregate.Intersect over a batch: returns (prim int32[n], t f64[n], p f64[n,3], n f64[n,3])."""
        n, cols = self._rays(o, d, tmax)
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float64)
        p = np.empty((n, 3), dtype=np.float64)
        nr = np.empty((n, 3), dtype=np.float64)
        ptr = lambda a: a.ctypes.data_as(abi.dp)
        rc = self.dev.lib.gopbrt_trace_closest(self.h, n, *[ptr(c) for c in cols], prim.ctypes.data_as(C.POINTER(C.c_int32)),
                                               ptr(t), ptr(p), ptr(nr))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_trace_closest rc={rc}: {self.dev.error()}")
        return prim, t, p, nr

    def IntersectP(self, o, d, tmax=None):
        n, cols = self._rays(o, d, tmax)
        hit = np.empty(n, dtype=np.uint8)
        ptr = lambda a: a.ctypes.data_as(abi.dp)
        rc = self.dev.lib.gopbrt_trace_any(self.h, n, *[ptr(c) for c in cols], hit.ctypes.data_as(C.POINTER(C.c_uint8)))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_trace_any rc={rc}: {self.dev.error()}")
        return hit.astype(bool)


def Render(gpu_scene, integrator, tileSize, mode=abi.MODE_STRICT, rank=0, world=1, flags=0, max_lanes=0, device_film=None, out=None, groups=0):
    """pbrt.Render(ctx, integrator, scene, tileSize) (integrator.go:291-350) on the GPU.  Fills camera.Film.pixels with
    the (H', W', 4) float64 film {X, Y, Z sums, filterWeightSum} and returns gopbrt_stats as a dict.
    device_film: optional CUDA device pointer (int) of H'*W'*4 doubles — the film then stays on the device.
    out: optional preallocated float64 array of the film's shape (e.g. a view of pinned host memory) to receive the film.
    groups: FAST mode only — lane groups per pixel tile (GOPBRT_FLAG_GROUPS_*), 0 = automatic."""
    cam = integrator.GetCamera()
    film = cam.GetFilm()
    lib = gpu_scene.dev.lib
    c, s, i, f = cam.abi(), integrator.GetSampler().abi(mode), integrator.abi(tileSize), film.abi()
    o = abi.RenderOptions(rank, world, flags | ((int(groups) & 0xff) << 8), max_lanes)
    st = abi.Stats()
    if device_film is not None:
        rc = lib.gopbrt_render_device(gpu_scene.h, C.byref(c), C.byref(s), C.byref(i), C.byref(f), C.byref(o),
                                      C.c_void_p(device_film), C.byref(st))
    else:
        host = True
        if out is not None:
            assert out.dtype == np.float64 and out.size == int(np.prod(film.shape())) and out.flags["C_CONTIGUOUS"]
            film.pixels = out.reshape(film.shape())
        elif (flags & abi.FLAG_REDUCE_FILM) and rank != 0:
            host = False  # the summed film exists on rank 0 only
        else:
            film.pixels = np.empty(film.shape(), dtype=np.float64)
        rc = lib.gopbrt_render(gpu_scene.h, C.byref(c), C.byref(s), C.byref(i), C.byref(f), C.byref(o),
                               film.pixels.ctypes.data_as(abi.dp) if host else None, C.byref(st))
    if rc != abi.OK:
        raise RuntimeError(f"gopbrt_render rc={rc}: {gpu_scene.dev.error()}")
    return st.as_dict()


class MultiDevice:
    """gopbrt_multi: one process, N GPUs (contexts + NCCL communicators)."""

    def __init__(self, n_gpus, devices=None):
        self.lib = abi.load()
        self.h = C.c_void_p()
        arr = (C.c_int * n_gpus)(*devices) if devices is not None else None
        rc = self.lib.gopbrt_multi_init(n_gpus, arr, C.byref(self.h))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_multi_init({n_gpus}) failed rc={rc} (no CPU fallback)")
        self.n = n_gpus

    def error(self):
        e = self.lib.gopbrt_multi_last_error(self.h)
        return e.decode() if e else ""

    def launches(self):
        return int(self.lib.gopbrt_multi_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.gopbrt_multi_shutdown(self.h)
            self.h = C.c_void_p()


class MultiGpuScene:
    """gopbrt_multi_scene: BVH built once on the host, uploaded to every device."""

    def __init__(self, mdev, scene):
        self.dev, self.scene = mdev, scene
        self.h = C.c_void_p()
        rc = mdev.lib.gopbrt_multi_scene_create(mdev.h, C.byref(scene.desc()), C.byref(self.h))
        if rc != abi.OK:
            raise RuntimeError(f"gopbrt_multi_scene_create failed rc={rc}: {mdev.error()}")

    def close(self):
        if self.h:
            self.dev.lib.gopbrt_multi_scene_destroy(self.h)
            self.h = C.c_void_p()


def RenderMulti(multi_scene, integrator, tileSize, mode=abi.MODE_FAST, flags=0, out=None, groups=0):
    """pbrt.Render over every GPU of the process (gopbrt_multi_render): samples (FAST) or tiles (STRICT) split by device, one NCCL
    film reduce, host film from device 0.  Returns the stats dict (counters summed over devices, times = slowest device)."""
    cam = integrator.GetCamera()
    film = cam.GetFilm()
    lib = multi_scene.dev.lib
    c, s, i, f = cam.abi(), integrator.GetSampler().abi(mode), integrator.abi(tileSize), film.abi()
    st = abi.Stats()
    if out is not None:
        assert out.dtype == np.float64 and out.size == int(np.prod(film.shape())) and out.flags["C_CONTIGUOUS"]
        film.pixels = out.reshape(film.shape())
    else:
        film.pixels = np.empty(film.shape(), dtype=np.float64)
    rc = lib.gopbrt_multi_render(multi_scene.h, C.byref(c), C.byref(s), C.byref(i), C.byref(f), flags | ((int(groups) & 0xff) << 8),
                                 film.pixels.ctypes.data_as(abi.dp), C.byref(st))
    if rc != abi.OK:
        raise RuntimeError(f"gopbrt_multi_render rc={rc}: {multi_scene.dev.error()}")
    return st.as_dict()
