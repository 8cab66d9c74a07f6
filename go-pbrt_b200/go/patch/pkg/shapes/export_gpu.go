// export_gpu.go — accessor for shapes.Disk (pkg/shapes/disk.go:14-35).  SOURCE ONLY, see pkg/gpudesc/desc.go.
package shapes

import (
	gomath "math"

	"github.com/ssttuu/go-pbrt/pkg/gpudesc"
	"github.com/ssttuu/go-pbrt/pkg/pbrt"
)

// ExportGPU returns NewDisk's arguments; phiMax goes back to degrees (disk.go:33 stores radians of the clamped value).
func (d *Disk) ExportGPU() (gpudesc.Disk, *pbrt.Transform) {
	return gpudesc.Disk{
		Height:      d.height,
		Radius:      d.radius,
		InnerRadius: d.innerRadius,
		PhiMaxDeg:   d.phiMax * 180.0 / gomath.Pi,
	}, d.shape.objectToWorld
}
