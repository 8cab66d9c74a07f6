// export_gpu.go — accessors for integrator.Path / DirectLighting (path.go:10-26, directlighting.go:17-36).
// SOURCE ONLY, see pkg/gpudesc/desc.go.
package integrator

import "github.com/ssttuu/go-pbrt/pkg/gpudesc"

func (p *Path) ExportGPU() gpudesc.Integrator {
	return gpudesc.Integrator{MaxDepth: p.maxDepth, RRThreshold: p.rrThreshold, LightStrategy: int32(p.lightSampleStrategy)}
}

func (d *DirectLighting) ExportGPU() gpudesc.Integrator {
	return gpudesc.Integrator{DirectLighting: true, MaxDepth: int32(d.maxDepth), LightStrategy: int32(d.strategy)}
}
