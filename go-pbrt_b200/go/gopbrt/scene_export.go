// scene_export.go — flattens what internal/render/server.go assembles (a BVH over pbrt.Primitives + the light list,
// server.go:29-132) into the tables of gopbrt_scene_desc, through the in-package accessors of go/patch/pkg/*/export_gpu.go.
// Follows go-pbrt_b200/pbrt.py Scene.desc() (the tested flattening) table by table: transforms, textures and materials are
// interned by pointer; a checkerboard's children are interned before the checkerboard itself, so child indices are always
// smaller than the parent's (the library requires a DAG in table order).  SOURCE ONLY (no Go toolchain in this image).
package gopbrt

/*
#include "gopbrt_cuda.h"
*/
import "C"

import (
	"github.com/pkg/errors"

	"github.com/ssttuu/go-pbrt/pkg/accelerator"
	"github.com/ssttuu/go-pbrt/pkg/gpudesc"
	"github.com/ssttuu/go-pbrt/pkg/integrator"
	"github.com/ssttuu/go-pbrt/pkg/lights"
	"github.com/ssttuu/go-pbrt/pkg/materials"
	"github.com/ssttuu/go-pbrt/pkg/pbrt"
	"github.com/ssttuu/go-pbrt/pkg/sampler"
	"github.com/ssttuu/go-pbrt/pkg/shapes"
	"github.com/ssttuu/go-pbrt/pkg/textures"
)

type exporter struct {
	d       SceneDesc
	xf      map[*pbrt.Transform]int32
	tex     map[pbrt.SpectrumTexture]int32
	mat     map[pbrt.Material]int32
	spheres map[*pbrt.Sphere]int32
	disks   map[*shapes.Disk]int32
}

func d16(a [16]float64) (out [16]C.double) {
	for i, v := range a {
		out[i] = C.double(v)
	}
	return
}
func d3(a [3]float64) [3]C.double { return [3]C.double{C.double(a[0]), C.double(a[1]), C.double(a[2])} }
func b32(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

func (e *exporter) transform(t *pbrt.Transform) int32 {
	if i, ok := e.xf[t]; ok {
		return i
	}
	g := t.ExportGPU()
	i := int32(len(e.d.Transforms))
	e.d.Transforms = append(e.d.Transforms, C.gopbrt_transform{m: d16(g.M), minv: d16(g.MInv)})
	e.xf[t] = i
	return i
}

func (e *exporter) texture(t pbrt.SpectrumTexture) (int32, error) {
	if i, ok := e.tex[t]; ok {
		return i, nil
	}
	var r C.gopbrt_texture
	switch v := t.(type) {
	case *pbrt.ConstantSpectrumTexture:
		s := v.Value()
		r.kind, r.tex1, r.tex2 = C.GOPBRT_TEX_CONSTANT, -1, -1
		r.rgb = d3([3]float64{s[0], s[1], s[2]})
	case *textures.Checkerboard2D:
		m, t1, t2 := v.ExportGPU()
		i1, err := e.texture(t1)
		if err != nil {
			return 0, err
		}
		i2, err := e.texture(t2)
		if err != nil {
			return 0, err
		}
		r.kind, r.tex1, r.tex2 = C.GOPBRT_TEX_CHECKERBOARD, C.int32_t(i1), C.int32_t(i2)
		switch mp := m.(type) {
		case *pbrt.PlanarMapping2D:
			vs, vt, ds, dt := mp.ExportGPU()
			r.mapping = C.GOPBRT_MAP_PLANAR
			r.vs, r.vt = d3([3]float64{vs.X, vs.Y, vs.Z}), d3([3]float64{vt.X, vt.Y, vt.Z})
			r.ds, r.dt = C.double(ds), C.double(dt)
		case *pbrt.UVMapping2D:
			su, sv, du, dv := mp.ExportGPU()
			r.mapping = C.GOPBRT_MAP_UV
			r.su, r.sv, r.du, r.dv = C.double(su), C.double(sv), C.double(du), C.double(dv)
		default:
			return 0, errors.Errorf("gpu backend: unsupported texture mapping %T", m)
		}
	default:
		return 0, errors.Errorf("gpu backend: unsupported spectrum texture %T", t)
	}
	i := int32(len(e.d.Textures))
	e.d.Textures = append(e.d.Textures, r)
	e.tex[t] = i
	return i, nil
}

func (e *exporter) material(m pbrt.Material) (int32, error) {
	if m == nil {
		return -1, nil // the reference panics when such a primitive is shaded (primitive.go:72-75); the backend counts it
	}
	if i, ok := e.mat[m]; ok {
		return i, nil
	}
	var g gpudesc.Material
	var ta, tb pbrt.SpectrumTexture
	var err error
	switch v := m.(type) {
	case *materials.MatteMaterial:
		g, ta, err = v.ExportGPU()
	case *materials.Mirror:
		g, ta, err = v.ExportGPU()
	case *materials.Glass:
		g, ta, tb, err = v.ExportGPU()
	default:
		err = errors.Errorf("gpu backend: unsupported material %T", m)
	}
	if err != nil {
		return 0, err
	}
	r := C.gopbrt_material{kind: C.int32_t(g.Kind), tex_b: -1, sigma: C.double(g.Sigma), eta: C.double(g.Eta),
		u_rough: C.double(g.URoughness), v_rough: C.double(g.VRoughness)}
	ia, err := e.texture(ta)
	if err != nil {
		return 0, err
	}
	r.tex_a = C.int32_t(ia)
	if tb != nil {
		ib, err := e.texture(tb)
		if err != nil {
			return 0, err
		}
		r.tex_b = C.int32_t(ib)
	}
	i := int32(len(e.d.Materials))
	e.d.Materials = append(e.d.Materials, r)
	e.mat[m] = i
	return i, nil
}

// shape interns a sphere or a disk and returns (kind, index).
func (e *exporter) shape(s pbrt.Shape) (C.int32_t, int32, error) {
	switch v := s.(type) {
	case *pbrt.Sphere:
		if i, ok := e.spheres[v]; ok {
			return C.GOPBRT_SHAPE_SPHERE, i, nil
		}
		g, o2w := v.ExportGPU()
		i := int32(len(e.d.Spheres))
		e.d.Spheres = append(e.d.Spheres, C.gopbrt_sphere{object_to_world: C.int32_t(e.transform(o2w)),
			reverse_orientation: b32(g.ReverseOrientation), radius: C.double(g.Radius), z_min: C.double(g.ZMin),
			z_max: C.double(g.ZMax), phi_max_deg: C.double(g.PhiMaxDeg)})
		e.spheres[v] = i
		return C.GOPBRT_SHAPE_SPHERE, i, nil
	case *shapes.Disk:
		if i, ok := e.disks[v]; ok {
			return C.GOPBRT_SHAPE_DISK, i, nil
		}
		g, o2w := v.ExportGPU()
		i := int32(len(e.d.Disks))
		e.d.Disks = append(e.d.Disks, C.gopbrt_disk{object_to_world: C.int32_t(e.transform(o2w)), height: C.double(g.Height),
			radius: C.double(g.Radius), inner_radius: C.double(g.InnerRadius), phi_max_deg: C.double(g.PhiMaxDeg)})
		e.disks[v] = i
		return C.GOPBRT_SHAPE_DISK, i, nil
	}
	return 0, 0, errors.Errorf("gpu backend: unsupported shape %T", s)
}

func (e *exporter) primitive(p pbrt.Primitive, primToWorld int32) error {
	switch v := p.(type) {
	case *pbrt.GeometricPrimitive:
		sh, m := v.ExportGPU()
		kind, idx, err := e.shape(sh)
		if err != nil {
			return err
		}
		mi, err := e.material(m)
		if err != nil {
			return err
		}
		e.d.Primitives = append(e.d.Primitives, C.gopbrt_primitive{shape_kind: kind, shape_index: C.int32_t(idx),
			material: C.int32_t(mi), prim_to_world: C.int32_t(primToWorld)})
		return nil
	case *pbrt.TransformedPrimitive:
		if primToWorld >= 0 {
			return errors.New("gpu backend: nested TransformedPrimitives are not supported")
		}
		inner, at := v.ExportGPU()
		t, static := at.StartTransform()
		if !static {
			return errors.New("gpu backend: moving AnimatedTransforms are not supported (Decompose is a TODO in the reference too)")
		}
		return e.primitive(inner, e.transform(t))
	case *accelerator.BVH:
		for _, q := range v.Primitives() {
			if err := e.primitive(q, primToWorld); err != nil {
				return err
			}
		}
		return nil
	}
	return errors.Errorf("gpu backend: unsupported primitive %T", p)
}

// ExportScene flattens NewBVH's primitives and NewScene's lights (server.go:104,132) into a SceneDesc.
func ExportScene(agg *accelerator.BVH, ls []pbrt.Light) (*SceneDesc, error) {
	e := &exporter{xf: map[*pbrt.Transform]int32{}, tex: map[pbrt.SpectrumTexture]int32{}, mat: map[pbrt.Material]int32{},
		spheres: map[*pbrt.Sphere]int32{}, disks: map[*shapes.Disk]int32{}}
	e.d.MaxPrimsInNode = agg.MaxPrimsInNode()
	if err := e.primitive(agg, -1); err != nil {
		return nil, err
	}
	for _, l := range ls {
		var r C.gopbrt_light
		switch v := l.(type) {
		case *lights.Point:
			g := v.ExportGPU()
			r.kind, r.rgb, r.v = C.GOPBRT_LIGHT_POINT, d3(g.RGB), d3(g.V)
		case *lights.Distant:
			g := v.ExportGPU()
			r.kind, r.rgb, r.v = C.GOPBRT_LIGHT_DISTANT, d3(g.RGB), d3(g.V)
		case *lights.DiffuseAreaLight:
			g, sh := v.ExportGPU()
			kind, idx, err := e.shape(sh)
			if err != nil {
				return nil, err
			}
			r.kind, r.rgb, r.shape_kind, r.shape_index, r.two_sided = C.GOPBRT_LIGHT_DIFFUSE_AREA, d3(g.RGB), kind, C.int32_t(idx), b32(g.TwoSided)
		default:
			return nil, errors.Errorf("gpu backend: unsupported light %T", l)
		}
		e.d.Lights = append(e.d.Lights, r)
	}
	return &e.d, nil
}

// ExportCamera / ExportSampler / ExportIntegrator / ExportFilm: the per-frame descriptors of gopbrt_render.
func ExportCamera(c *pbrt.PerspectiveCamera) (C.gopbrt_camera, error) {
	g, err := c.ExportGPU()
	return C.gopbrt_camera{raster_to_camera: d16(g.RasterToCamera), camera_to_world: d16(g.CameraToWorld),
		lens_radius: C.double(g.LensRadius), focal_distance: C.double(g.FocalDistance),
		shutter_open: C.double(g.ShutterOpen), shutter_close: C.double(g.ShutterClose)}, err
}

func ExportSampler(s pbrt.Sampler, mode int) (C.gopbrt_sampler, error) {
	var g gpudesc.Sampler
	switch v := s.(type) {
	case *sampler.Stratified:
		g = v.ExportGPU()
	case *sampler.RandomSampler:
		g = v.ExportGPU()
	default:
		return C.gopbrt_sampler{}, errors.Errorf("gpu backend: unsupported sampler %T", s)
	}
	kind := C.int32_t(C.GOPBRT_SAMPLER_STRATIFIED)
	if g.Random {
		kind = C.GOPBRT_SAMPLER_RANDOM
	}
	return C.gopbrt_sampler{kind: kind, x_samples: C.int32_t(g.XSamples), y_samples: C.int32_t(g.YSamples), jitter: b32(g.Jitter),
		n_sampled_dimensions: C.int32_t(g.NSampledDimensions), mode: C.int32_t(mode)}, nil
}

func ExportIntegrator(i pbrt.Integrator, tileSize int64) (C.gopbrt_integrator, error) {
	var g gpudesc.Integrator
	switch v := i.(type) {
	case *integrator.Path:
		g = v.ExportGPU()
	case *integrator.DirectLighting:
		g = v.ExportGPU()
	default:
		return C.gopbrt_integrator{}, errors.Errorf("gpu backend: unsupported integrator %T", i)
	}
	kind := C.int32_t(C.GOPBRT_INTEGRATOR_PATH)
	if g.DirectLighting {
		kind = C.GOPBRT_INTEGRATOR_DIRECT_LIGHTING
	}
	return C.gopbrt_integrator{kind: kind, max_depth: C.int32_t(g.MaxDepth), rr_threshold: C.double(g.RRThreshold),
		light_strategy: C.int32_t(g.LightStrategy), tile_size: C.int64_t(tileSize)}, nil
}

func ExportFilm(f *pbrt.Film) C.gopbrt_film {
	g := f.ExportGPU()
	return C.gopbrt_film{width: C.int32_t(g.Width), height: C.int32_t(g.Height),
		crop:          [4]C.double{C.double(g.Crop[0]), C.double(g.Crop[1]), C.double(g.Crop[2]), C.double(g.Crop[3])},
		filter_radius: [2]C.double{C.double(g.FilterRadius[0]), C.double(g.FilterRadius[1])}}
}
