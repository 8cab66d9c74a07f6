// export_gpu.go — in-package accessors for the fields the GPU backend needs and pkg/pbrt does not export
// (Sphere, GeometricPrimitive, TransformedPrimitive, textures and mappings, PerspectiveCamera, Film).
// Added to pkg/pbrt by the integration patch (INTEGRATION.md §3); nothing here changes the CPU renderer.
// SOURCE ONLY (no Go toolchain in the image this was written in).
package pbrt

import (
	"errors"
	gomath "math"

	"github.com/ssttuu/go-pbrt/pkg/gpudesc"
)

func flat(m *Matrix4x4) (out [16]float64) {
	for i := 0; i < 4; i++ {
		for j := 0; j < 4; j++ {
			out[4*i+j] = m[i][j]
		}
	}
	return
}

// ExportGPU passes Matrix and MatrixInverse exactly as the struct holds them (transform.go:144-146).
func (t *Transform) ExportGPU() gpudesc.Transform {
	return gpudesc.Transform{M: flat(t.Matrix), MInv: flat(t.MatrixInverse)}
}

// StartTransform: the only transform a static AnimatedTransform ever applies (Interpolate returns it when
// !actuallyAnimated, transform.go:550-560).  ok is false for a moving transform, which the backend does not take.
func (a *AnimatedTransform) StartTransform() (t *Transform, ok bool) {
	return a.startTransform, !a.actuallyAnimated
}

// ExportGPU returns NewSphere's arguments (sphere.go:19-32); objectToWorld is returned separately so that the
// caller can deduplicate transforms by pointer.
func (s *Sphere) ExportGPU() (gpudesc.Sphere, *Transform) {
	return gpudesc.Sphere{
		ReverseOrientation: s.reverseOrientation,
		Radius:             s.radius,
		ZMin:               s.zMin,
		ZMax:               s.zMax,
		PhiMaxDeg:          s.phiMax * 180.0 / gomath.Pi,
	}, s.objectToWorld
}

// ExportGPU of a GeometricPrimitive: its shape and material (primitive.go:22-27).
func (p *GeometricPrimitive) ExportGPU() (Shape, Material) { return p.Shape, p.material }

// ExportGPU of a TransformedPrimitive: the wrapped primitive and its primitiveToWorld (primitive.go:89-92).
func (p *TransformedPrimitive) ExportGPU() (Primitive, *AnimatedTransform) {
	return p.primitive, p.primitiveToWorld
}

func (t *ConstantSpectrumTexture) Value() Spectrum { return t.value } // texture.go:56-58
func (t *ConstantFloatTexture) Value() float64     { return t.value } // texture.go:70-72

func (t *UVMapping2D) ExportGPU() (su, sv, du, dv float64) { return t.su, t.sv, t.du, t.dv } // texture.go:18-20
func (m *PlanarMapping2D) ExportGPU() (vs, vt *Vector3f, ds, dt float64) { // texture.go:37-40
	return m.vs, m.vt, m.ds, m.dt
}

// ExportGPU: the two matrices GenerateRayDifferential reads plus lens and shutter as the struct holds them
// (camera.go:106-165; note camera.go:116 stores shutterClose = shutterOpen).
func (c *PerspectiveCamera) ExportGPU() (gpudesc.Camera, error) {
	c2w, ok := c.cameraToWorld.StartTransform()
	if !ok {
		return gpudesc.Camera{}, errors.New("gpu backend: moving cameras are not supported")
	}
	return gpudesc.Camera{
		RasterToCamera: flat(c.RasterToCamera.Matrix),
		CameraToWorld:  flat(c2w.Matrix),
		LensRadius:     c.lensRadius,
		FocalDistance:  c.focalDistance,
		ShutterOpen:    c.shutterOpen,
		ShutterClose:   c.shutterClose,
	}, nil
}

// ExportGPU: NewFilm's resolution, crop window and box-filter radius (film.go:43-76).  The crop window is recovered
// from CroppedPixelBounds, which is what the hot path reads (ceil(res * crop) reproduces the same integers).
func (f *Film) ExportGPU() gpudesc.Film {
	w, h := float64(f.FullResolution.X), float64(f.FullResolution.Y)
	r := f.Filter.GetRadius()
	return gpudesc.Film{
		Width: int32(f.FullResolution.X), Height: int32(f.FullResolution.Y),
		Crop: [4]float64{float64(f.CroppedPixelBounds.Min.X) / w, float64(f.CroppedPixelBounds.Min.Y) / h,
			float64(f.CroppedPixelBounds.Max.X) / w, float64(f.CroppedPixelBounds.Max.Y) / h},
		FilterRadius: [2]float64{r.X, r.Y},
	}
}

// SetPixels hands the backend's film back: px holds {X, Y, Z sums, filterWeightSum} per pixel, row-major over
// CroppedPixelBounds — the Pixel fields MergeFilmTile accumulates (film.go:115-132).  WriteImage (film.go:142-179)
// then runs unchanged.
func (f *Film) SetPixels(px []float64) {
	f.mutex.Lock()
	defer f.mutex.Unlock()
	for i := range f.pixels {
		f.pixels[i].value = [3]float64{px[4*i], px[4*i+1], px[4*i+2]}
		f.pixels[i].filterWeightSum = px[4*i+3]
	}
}
