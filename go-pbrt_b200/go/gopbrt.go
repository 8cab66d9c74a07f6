// Package gopbrt is the cgo shim a go-pbrt maintainer adds to run the ray-intersection + path-integrator hot path on
// libgopbrt_cuda.so (include/gopbrt_cuda.h).  SOURCE ONLY: this image has no Go toolchain (SURVEY.md §0.1), so the file
// is neither compiled nor tested here; the tested host is go-pbrt_b200/pbrt.py, which flattens scenes the same way.
//
// Integrator kinds: integ.kind = GOPBRT_INTEGRATOR_PATH (integrator.NewPath) or GOPBRT_INTEGRATOR_DIRECT_LIGHTING
// (integrator.NewDirectLighting; light_strategy = GOPBRT_DL_SAMPLE_ALL / GOPBRT_DL_SAMPLE_ONE).  Sampler modes: STRICT reproduces
// the reference's per-tile RNG streams; FAST (counter-based, needs the matching Go Sampler of INTEGRATION.md §3) splits a
// pixel's samples across GPUs (opt.rank / opt.world) and across lane groups inside a GPU (opt.flags bits 8..15, 0 = automatic).
//
// It replaces two call sites of internal/render/server.go:
//   agg := accelerator.NewBVH(primitives, 2, accelerator.SplitSAH); scene := pbrt.NewScene(agg, ls)   (:104,:132)
//       -> scene, err := gopbrt.NewScene(dev, desc)
//   err = pbrt.Render(ctx, dli, scene, 16)                                                             (:164)
//       -> err = gopbrt.Render(ctx, scene, camera, sampler, integrator, film, tileSize)
// Most fields needed to fill the descriptors are unexported in pkg/pbrt (Sphere.radius, GeometricPrimitive.material,
// Point.pLight, ... SURVEY §8b), so the exporter functions belong in-package (pkg/pbrt/export_gpu.go etc.).
package gopbrt

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../csrc -lgopbrt_cuda
#include <stdlib.h>
#include "gopbrt_cuda.h"
*/
import "C"

import (
	"context"
	"unsafe"

	"github.com/pkg/errors"
)

// Device is one gopbrt_ctx: one per process and GPU.
type Device struct{ h *C.gopbrt_ctx }

func NewDevice(ordinal int) (*Device, error) {
	var h *C.gopbrt_ctx
	if rc := C.gopbrt_init(C.int(ordinal), &h); rc != C.GOPBRT_OK {
		return nil, errors.Errorf("gopbrt_init(%d): status %d (no sm_100 device; there is no CPU fallback)", ordinal, int(rc))
	}
	return &Device{h}, nil
}

func (d *Device) Close() { C.gopbrt_shutdown(d.h) }

// SceneDesc mirrors gopbrt_scene_desc with Go slices; the slices are only read during NewScene (no Go pointer is
// retained by C after the call returns).
type SceneDesc struct {
	Transforms     []C.gopbrt_transform
	Spheres        []C.gopbrt_sphere
	Disks          []C.gopbrt_disk
	Vertices       []float64
	Triangles      []C.gopbrt_triangle
	Primitives     []C.gopbrt_primitive
	Materials      []C.gopbrt_material
	Textures       []C.gopbrt_texture
	Lights         []C.gopbrt_light
	MaxPrimsInNode int
}

type Scene struct {
	dev *Device
	h   *C.gopbrt_scene
}

func ptr[T any](s []T) *T {
	if len(s) == 0 {
		return nil
	}
	return &s[0]
}

// NewScene == accelerator.NewBVH + pbrt.NewScene: copies everything, builds and flattens the BVH, uploads once.
func NewScene(dev *Device, d *SceneDesc) (*Scene, error) {
	// the descriptor itself lives in C memory so that it may hold Go pointers only for the duration of the call
	cd := (*C.gopbrt_scene_desc)(C.calloc(1, C.size_t(unsafe.Sizeof(C.gopbrt_scene_desc{}))))
	defer C.free(unsafe.Pointer(cd))
	cd.n_transforms, cd.transforms = C.int32_t(len(d.Transforms)), ptr(d.Transforms)
	cd.n_spheres, cd.spheres = C.int32_t(len(d.Spheres)), ptr(d.Spheres)
	cd.n_disks, cd.disks = C.int32_t(len(d.Disks)), ptr(d.Disks)
	cd.n_vertices, cd.vertices = C.int64_t(len(d.Vertices)/3), (*C.double)(ptr(d.Vertices))
	cd.n_triangles, cd.triangles = C.int64_t(len(d.Triangles)), ptr(d.Triangles)
	cd.n_primitives, cd.primitives = C.int64_t(len(d.Primitives)), ptr(d.Primitives)
	cd.n_materials, cd.materials = C.int32_t(len(d.Materials)), ptr(d.Materials)
	cd.n_textures, cd.textures = C.int32_t(len(d.Textures)), ptr(d.Textures)
	cd.n_lights, cd.lights = C.int32_t(len(d.Lights)), ptr(d.Lights)
	cd.max_prims_in_node = C.int32_t(d.MaxPrimsInNode)
	var h *C.gopbrt_scene
	// NOTE: cgo's pointer-passing rules forbid C memory holding Go pointers across the call boundary unless pinned;
	// with Go >= 1.21 wrap the slices in a runtime.Pinner, with older toolchains copy them to C.malloc'ed arrays.
	if rc := C.gopbrt_scene_create(dev.h, cd, &h); rc != C.GOPBRT_OK {
		return nil, errors.Errorf("gopbrt_scene_create: status %d: %s", int(rc), C.GoString(C.gopbrt_last_error(dev.h)))
	}
	return &Scene{dev, h}, nil
}

func (s *Scene) Close() { C.gopbrt_scene_destroy(s.h) }

// Intersect == Aggregate.Intersect over a batch (SoA float64), used by parity tests on the Go side.
func (s *Scene) Intersect(ox, oy, oz, dx, dy, dz, tmax []float64) (prim []int32, t []float64, err error) {
	n := len(ox)
	prim, t = make([]int32, n), make([]float64, n)
	if n == 0 {
		return
	}
	rc := C.gopbrt_trace_closest(s.h, C.int64_t(n), (*C.double)(&ox[0]), (*C.double)(&oy[0]), (*C.double)(&oz[0]),
		(*C.double)(&dx[0]), (*C.double)(&dy[0]), (*C.double)(&dz[0]), (*C.double)(&tmax[0]),
		(*C.int32_t)(&prim[0]), (*C.double)(&t[0]), nil, nil)
	if rc != C.GOPBRT_OK {
		err = errors.Errorf("gopbrt_trace_closest: status %d: %s", int(rc), C.GoString(C.gopbrt_last_error(s.dev.h)))
	}
	return
}

// Render has pbrt.Render's meaning (pkg/pbrt/integrator.go:291): it fills film (W'*H'*4 float64: XYZ sums +
// filterWeightSum, row-major over CroppedPixelBounds — the layout of Film.pixels, film.go:20-25) which the caller hands
// to Film.WriteImage.  ctx cancellation is forwarded with gopbrt_cancel from a watcher goroutine (the reference only
// observes ctx in the tile producer, integrator.go:332-336).
func Render(ctx context.Context, s *Scene, cam *C.gopbrt_camera, smp *C.gopbrt_sampler, integ *C.gopbrt_integrator,
	film *C.gopbrt_film, opt *C.gopbrt_render_options, out []float64) (C.gopbrt_stats, error) {
	var st C.gopbrt_stats
	done := make(chan struct{})
	go func() {
		select {
		case <-ctx.Done():
			C.gopbrt_cancel(s.h)
		case <-done:
		}
	}()
	rc := C.gopbrt_render(s.h, cam, smp, integ, film, opt, (*C.double)(&out[0]), &st)
	close(done)
	switch rc {
	case C.GOPBRT_OK:
		return st, nil
	case C.GOPBRT_ERR_CANCELLED:
		return st, errors.Wrap(ctx.Err(), "waiting for render group") // same wrap as integrator.go:344
	default:
		return st, errors.Errorf("gopbrt_render: status %d: %s", int(rc), C.GoString(C.gopbrt_last_error(s.dev.h)))
	}
}
