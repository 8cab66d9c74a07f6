// export_gpu.go — accessors for materials.MatteMaterial / Mirror / Glass (matte.go:8-19, mirror.go:9-24,
// glass.go:7-25).  Float textures are constants in the reference (texture.go:70-82), so they travel by value.
// SOURCE ONLY, see pkg/gpudesc/desc.go.
package materials

import (
	"errors"

	"github.com/ssttuu/go-pbrt/pkg/gpudesc"
	"github.com/ssttuu/go-pbrt/pkg/pbrt"
)

func constFloat(t pbrt.FloatTexture) (float64, error) {
	if t == nil {
		return 0, nil
	}
	c, ok := t.(*pbrt.ConstantFloatTexture)
	if !ok {
		return 0, errors.New("gpu backend: only ConstantFloatTexture exists in the reference")
	}
	return c.Value(), nil
}

// ExportGPU: Kd (a spectrum texture the caller interns) and sigma.  A bump map is not supported (Bump is unreached
// by every scene the reference builds).
func (m *MatteMaterial) ExportGPU() (gpudesc.Material, pbrt.SpectrumTexture, error) {
	if m.bumpMap != nil {
		return gpudesc.Material{}, nil, errors.New("gpu backend: bump maps are not supported")
	}
	sigma, err := constFloat(m.sigma)
	return gpudesc.Material{Kind: gpudesc.Matte, TexB: -1, Sigma: sigma}, m.Kd, err
}

func (m *Mirror) ExportGPU() (gpudesc.Material, pbrt.SpectrumTexture, error) {
	if m.bumpMap != nil {
		return gpudesc.Material{}, nil, errors.New("gpu backend: bump maps are not supported")
	}
	return gpudesc.Material{Kind: gpudesc.Mirror, TexB: -1}, m.Kr, nil
}

// ExportGPU: Kr, Kt, the index and the two roughnesses (non-zero roughness selects the microfacet branch, on which
// the reference panics, glass.go:41-55; the backend counts it the same way).
func (g *Glass) ExportGPU() (gpudesc.Material, pbrt.SpectrumTexture, pbrt.SpectrumTexture, error) {
	if g.bumpMap != nil {
		return gpudesc.Material{}, nil, nil, errors.New("gpu backend: bump maps are not supported")
	}
	eta, err := constFloat(g.index)
	if err != nil {
		return gpudesc.Material{}, nil, nil, err
	}
	ur, err := constFloat(g.uRoughness)
	if err != nil {
		return gpudesc.Material{}, nil, nil, err
	}
	vr, err := constFloat(g.vRoughness)
	return gpudesc.Material{Kind: gpudesc.Glass, Eta: eta, URoughness: ur, VRoughness: vr}, g.Kr, g.Kt, err
}
