// export_gpu.go — accessors for sampler.Stratified / RandomSampler (stratified.go:5-20, random.go:6-19).
// SOURCE ONLY, see pkg/gpudesc/desc.go.
package sampler

import "github.com/ssttuu/go-pbrt/pkg/gpudesc"

// ExportGPU: NewStratified's arguments; nSampledDimensions is the length of the 1-D sample tables (pixel.go:16-32).
func (s *Stratified) ExportGPU() gpudesc.Sampler {
	return gpudesc.Sampler{XSamples: s.xPixelSamples, YSamples: s.yPixelSamples, Jitter: s.jitterSamples,
		NSampledDimensions: int32(len(s.PixelSampler.samples1D))}
}

// ExportGPU: ns samples per pixel; the seed is overwritten per tile by Clone (integrator.go:318-328).
func (s *RandomSampler) ExportGPU() gpudesc.Sampler {
	return gpudesc.Sampler{Random: true, XSamples: s.SamplesPerPixel, YSamples: 1}
}
