// export_gpu.go — accessor for textures.Checkerboard2D (checkerboard.go:14-28).  SOURCE ONLY, see pkg/gpudesc/desc.go.
package textures

import "github.com/ssttuu/go-pbrt/pkg/pbrt"

// ExportGPU: the mapping and the two child textures (antiAliasingMethod is always `none`, checkerboard.go:19).
func (c *Checkerboard2D) ExportGPU() (pbrt.TextureMapping2D, pbrt.SpectrumTexture, pbrt.SpectrumTexture) {
	return c.mapping, c.tex1, c.tex2
}
