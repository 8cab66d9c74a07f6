// export_gpu.go — accessors for lights.Point / Distant / DiffuseAreaLight (point.go:8-31, distant.go:8-31,
// diffuse.go:8-24).  SOURCE ONLY, see pkg/gpudesc/desc.go.
package lights

import (
	"github.com/ssttuu/go-pbrt/pkg/gpudesc"
	"github.com/ssttuu/go-pbrt/pkg/pbrt"
)

func rgb(s pbrt.Spectrum) [3]float64 { return [3]float64{s[0], s[1], s[2]} }

// ExportGPU: pLight as NewPoint computed it (lightToWorld applied to the origin with Go's own arithmetic) and I.
func (l *Point) ExportGPU() gpudesc.Light {
	return gpudesc.Light{Kind: gpudesc.Point, RGB: rgb(l.I), V: [3]float64{l.pLight.X, l.pLight.Y, l.pLight.Z}}
}

// ExportGPU: wLight as the struct holds it — already transformed and normalised (distant.go:23-29) — and L.
// worldCenter / worldRadius are set by Preprocess from the scene bound; the backend derives them the same way.
func (d *Distant) ExportGPU() gpudesc.Light {
	return gpudesc.Light{Kind: gpudesc.Distant, RGB: rgb(d.L), V: [3]float64{d.wLight.X, d.wLight.Y, d.wLight.Z}}
}

// ExportGPU: LEmit, twoSided and the emitting shape (the caller resolves it to a sphere / disk table index).
func (l *DiffuseAreaLight) ExportGPU() (gpudesc.Light, pbrt.Shape) {
	return gpudesc.Light{Kind: gpudesc.DiffuseArea, RGB: rgb(l.LEmit), TwoSided: l.twoSided}, l.shape
}
