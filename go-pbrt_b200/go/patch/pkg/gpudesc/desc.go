// Package gpudesc holds the plain-Go mirror of the descriptors in include/gopbrt_cuda.h.  It has no cgo and no
// dependency on pkg/pbrt, so every package of the renderer (pbrt, shapes, lights, materials, textures, sampler,
// integrator, accelerator) can export its unexported fields into it without an import cycle; the cgo package
// (gopbrt) converts these structs field by field into the C ones.
//
// SOURCE ONLY: no Go toolchain exists in the image this backend was built in, so these files are neither compiled nor
// tested here.  The tested flattening is go-pbrt_b200/pbrt.py (Scene.desc), which these files follow table by table.
package gpudesc

// Transform is pbrt.Transform{Matrix, MatrixInverse} (pkg/pbrt/transform.go:144-146), row-major, verbatim: the
// backend never recomputes an inverse (Transform.Mul's is "wrong" for non-commuting factors, transform.go:179-184,
// and parity needs it as it is).
type Transform struct{ M, MInv [16]float64 }

type ShapeKind int32

const (
	ShapeSphere ShapeKind = iota
	ShapeDisk
	ShapeTriangle
)

// Sphere: the arguments NewSphere was called with (sphere.go:19-32).  PhiMaxDeg is in DEGREES: the struct keeps
// radians (sphere.go:30), the exporter converts back with the constructor's own clamp so that the backend redoes
// math.Radians(math.Clamp(phiMax, 0, 360)) on the same number.
type Sphere struct {
	ObjectToWorld      int32 // index into SceneDesc.Transforms
	ReverseOrientation bool
	Radius, ZMin, ZMax float64
	PhiMaxDeg          float64
}

type Disk struct {
	ObjectToWorld               int32
	Height, Radius, InnerRadius float64
	PhiMaxDeg                   float64
}

type Primitive struct {
	Kind        ShapeKind
	ShapeIndex  int32
	Material    int32 // -1: nil material (GeometricPrimitive.ComputeScatteringFunctions panics, primitive.go:72-75)
	PrimToWorld int32 // -1 for a bare GeometricPrimitive; else the TransformedPrimitive's static transform
}

type MaterialKind int32

const (
	Matte MaterialKind = iota
	Mirror
	Glass
)

type Material struct {
	Kind                   MaterialKind
	TexA, TexB             int32 // matte: Kd; mirror: Kr; glass: Kr, Kt
	Sigma, Eta             float64
	URoughness, VRoughness float64
}

type TextureKind int32

const (
	TexConstant TextureKind = iota
	TexCheckerboard
)

type Texture struct {
	Kind           TextureKind
	Planar         bool
	Tex1, Tex2     int32
	RGB            [3]float64
	Vs, Vt         [3]float64
	Ds, Dt         float64
	Su, Sv, Du, Dv float64
}

type LightKind int32

const (
	Distant LightKind = iota
	Point
	DiffuseArea
)

type Light struct {
	Kind       LightKind
	ShapeKind  ShapeKind
	ShapeIndex int32
	TwoSided   bool
	RGB        [3]float64 // L / I / LEmit
	V          [3]float64 // Distant.wLight (already transformed + normalised) / Point.pLight
}

type Camera struct {
	RasterToCamera, CameraToWorld [16]float64
	LensRadius, FocalDistance     float64
	ShutterOpen, ShutterClose     float64
}

type Sampler struct {
	Random                     bool
	XSamples, YSamples         int32
	Jitter                     bool
	NSampledDimensions         int32
}

type Integrator struct {
	DirectLighting bool
	MaxDepth       int32
	RRThreshold    float64
	LightStrategy  int32
}

type Film struct {
	Width, Height int32
	Crop          [4]float64
	FilterRadius  [2]float64
}
