// Package gopbrt is the cgo shim a go-pbrt maintainer adds to run the ray-intersection + path-integrator hot path on
// libgopbrt_cuda.so (include/gopbrt_cuda.h).  SOURCE ONLY: this image has no Go toolchain (SURVEY.md §0.1), so the file
// is neither compiled nor tested here; the tested host is go-pbrt_b200/pbrt.py, which flattens scenes the same way, and
// tests/cpp/abi_smoke.c drives the same C calls from plain C.
//
// It replaces two call sites of internal/render/server.go:
//   agg := accelerator.NewBVH(primitives, 2, accelerator.SplitSAH); scene := pbrt.NewScene(agg, ls)   (:104,:132)
//       -> desc, err := gopbrt.ExportScene(agg, ls);  scene, err := gopbrt.NewScene(dev, desc)          (scene_export.go)
//   err = pbrt.Render(ctx, dli, scene, 16)                                                             (:164)
//       -> err = gopbrt.Render(ctx, scene, camera, sampler, integrator, tileSize, gopbrt.ModeStrict)
// and, for a daemon that owns several GPUs, NewMulti / NewMultiScene / RenderMulti (one process, N devices, one NCCL film
// reduce — gopbrt_multi_* in the header).
package gopbrt

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../csrc -lgopbrt_cuda
#include <stdlib.h>
#include <string.h>
#include "gopbrt_cuda.h"
*/
import "C"

import (
	"context"
	"unsafe"

	"github.com/pkg/errors"
)

const (
	ModeStrict = C.GOPBRT_MODE_STRICT // the reference's RNG stream per tile: identical per-pixel sample sequences
	ModeFast   = C.GOPBRT_MODE_FAST   // counter-based stream per (pixel, sample): pairs with FastStratified (fast_sampler.go)
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

// cArray copies a Go slice into C.malloc'ed memory (freed by the returned func).  The descriptor handed to
// gopbrt_scene_create lives in C memory and must not hold Go pointers (cgo pointer-passing rules), so every table is
// copied once here; the library copies them again into its own storage and keeps no pointer after the call.
func cArray[T any](s []T) (unsafe.Pointer, func()) {
	if len(s) == 0 {
		return nil, func() {}
	}
	n := C.size_t(len(s)) * C.size_t(unsafe.Sizeof(s[0]))
	p := C.malloc(n)
	C.memcpy(p, unsafe.Pointer(&s[0]), n)
	return p, func() { C.free(p) }
}

// SceneDesc mirrors gopbrt_scene_desc with Go slices (filled by ExportScene).
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

// fill builds the C descriptor in C memory; the returned func frees everything.
func (d *SceneDesc) fill() (*C.gopbrt_scene_desc, func()) {
	cd := (*C.gopbrt_scene_desc)(C.calloc(1, C.size_t(unsafe.Sizeof(C.gopbrt_scene_desc{}))))
	frees := []func(){func() { C.free(unsafe.Pointer(cd)) }}
	keep := func(p unsafe.Pointer, f func()) unsafe.Pointer { frees = append(frees, f); return p }
	cd.n_transforms, cd.transforms = C.int32_t(len(d.Transforms)), (*C.gopbrt_transform)(keep(cArray(d.Transforms)))
	cd.n_spheres, cd.spheres = C.int32_t(len(d.Spheres)), (*C.gopbrt_sphere)(keep(cArray(d.Spheres)))
	cd.n_disks, cd.disks = C.int32_t(len(d.Disks)), (*C.gopbrt_disk)(keep(cArray(d.Disks)))
	cd.n_vertices, cd.vertices = C.int64_t(len(d.Vertices)/3), (*C.double)(keep(cArray(d.Vertices)))
	cd.n_triangles, cd.triangles = C.int64_t(len(d.Triangles)), (*C.gopbrt_triangle)(keep(cArray(d.Triangles)))
	cd.n_primitives, cd.primitives = C.int64_t(len(d.Primitives)), (*C.gopbrt_primitive)(keep(cArray(d.Primitives)))
	cd.n_materials, cd.materials = C.int32_t(len(d.Materials)), (*C.gopbrt_material)(keep(cArray(d.Materials)))
	cd.n_textures, cd.textures = C.int32_t(len(d.Textures)), (*C.gopbrt_texture)(keep(cArray(d.Textures)))
	cd.n_lights, cd.lights = C.int32_t(len(d.Lights)), (*C.gopbrt_light)(keep(cArray(d.Lights)))
	cd.max_prims_in_node = C.int32_t(d.MaxPrimsInNode)
	return cd, func() {
		for _, f := range frees {
			f()
		}
	}
}

type Scene struct {
	dev *Device
	h   *C.gopbrt_scene
}

// NewScene == accelerator.NewBVH + pbrt.NewScene: copies everything, builds and flattens the BVH, uploads once.
func NewScene(dev *Device, d *SceneDesc) (*Scene, error) {
	cd, free := d.fill()
	defer free()
	var h *C.gopbrt_scene
	if rc := C.gopbrt_scene_create(dev.h, cd, &h); rc != C.GOPBRT_OK {
		return nil, errors.Errorf("gopbrt_scene_create: status %d: %s", int(rc), C.GoString(C.gopbrt_last_error(dev.h)))
	}
	return &Scene{dev, h}, nil
}

func (s *Scene) Close() { C.gopbrt_scene_destroy(s.h) }

// Intersect == Aggregate.Intersect over a batch (SoA float64), used by parity tests on the Go side.  The slices are
// plain []float64 / []int32 (no pointers inside), which cgo may pass directly for the duration of the call.
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

// watch forwards ctx cancellation to cancel() while a render call runs.  The returned stop func JOINS the watcher, so
// no cancel can be issued after the render call has returned (a late one would hit the scene's next frame, or a
// closed scene).  A context that is already done never starts the frame.
func watch(ctx context.Context, cancel func()) (stop func(), err error) {
	if err = ctx.Err(); err != nil {
		return func() {}, err
	}
	done, joined := make(chan struct{}), make(chan struct{})
	go func() {
		defer close(joined)
		select {
		case <-ctx.Done():
			cancel()
		case <-done:
		}
	}()
	return func() { close(done); <-joined }, nil
}

// Render has pbrt.Render's meaning (pkg/pbrt/integrator.go:291): it fills out (W'*H'*4 float64: XYZ sums +
// filterWeightSum, row-major over CroppedPixelBounds — the layout of Film.pixels, film.go:20-25), which the caller hands
// to Film.SetPixels + Film.WriteImage.  ctx cancellation is forwarded with gopbrt_cancel (the reference only observes
// ctx in the tile producer, integrator.go:332-336).
func Render(ctx context.Context, s *Scene, cam *C.gopbrt_camera, smp *C.gopbrt_sampler, integ *C.gopbrt_integrator,
	film *C.gopbrt_film, opt *C.gopbrt_render_options, out []float64) (C.gopbrt_stats, error) {
	var st C.gopbrt_stats
	stop, err := watch(ctx, func() { C.gopbrt_cancel(s.h) })
	if err != nil {
		return st, errors.Wrap(err, "waiting for render group")
	}
	rc := C.gopbrt_render(s.h, cam, smp, integ, film, opt, (*C.double)(&out[0]), &st)
	stop()
	switch rc {
	case C.GOPBRT_OK:
		return st, nil
	case C.GOPBRT_ERR_CANCELLED:
		return st, errors.Wrap(ctx.Err(), "waiting for render group") // same wrap as integrator.go:344
	default:
		return st, errors.Errorf("gopbrt_render: status %d: %s", int(rc), C.GoString(C.gopbrt_last_error(s.dev.h)))
	}
}

// ---- one process, N GPUs ----

// Multi is gopbrt_multi: one context per device plus their NCCL communicators (ncclCommInitAll).
type Multi struct{ h *C.gopbrt_multi }

func NewMulti(nGPUs int) (*Multi, error) {
	var h *C.gopbrt_multi
	if rc := C.gopbrt_multi_init(C.int(nGPUs), nil, &h); rc != C.GOPBRT_OK {
		return nil, errors.Errorf("gopbrt_multi_init(%d): status %d", nGPUs, int(rc))
	}
	return &Multi{h}, nil
}

func (m *Multi) Close() { C.gopbrt_multi_shutdown(m.h) }

type MultiScene struct {
	m *Multi
	h *C.gopbrt_multi_scene
}

// NewMultiScene builds the BVH once on the host and uploads the scene to every device.
func NewMultiScene(m *Multi, d *SceneDesc) (*MultiScene, error) {
	cd, free := d.fill()
	defer free()
	var h *C.gopbrt_multi_scene
	if rc := C.gopbrt_multi_scene_create(m.h, cd, &h); rc != C.GOPBRT_OK {
		return nil, errors.Errorf("gopbrt_multi_scene_create: status %d: %s", int(rc), C.GoString(C.gopbrt_multi_last_error(m.h)))
	}
	return &MultiScene{m, h}, nil
}

func (s *MultiScene) Close() { C.gopbrt_multi_scene_destroy(s.h) }

// RenderMulti is pbrt.Render over every GPU of the process: samples (ModeFast) or tiles (ModeStrict) split by device,
// one ncclReduce of the films onto device 0, out read from there.
func RenderMulti(ctx context.Context, s *MultiScene, cam *C.gopbrt_camera, smp *C.gopbrt_sampler, integ *C.gopbrt_integrator,
	film *C.gopbrt_film, flags int, out []float64) (C.gopbrt_stats, error) {
	var st C.gopbrt_stats
	stop, err := watch(ctx, func() { C.gopbrt_multi_cancel(s.h) })
	if err != nil {
		return st, errors.Wrap(err, "waiting for render group")
	}
	rc := C.gopbrt_multi_render(s.h, cam, smp, integ, film, C.int(flags), (*C.double)(&out[0]), &st)
	stop()
	switch rc {
	case C.GOPBRT_OK:
		return st, nil
	case C.GOPBRT_ERR_CANCELLED:
		return st, errors.Wrap(ctx.Err(), "waiting for render group")
	default:
		return st, errors.Errorf("gopbrt_multi_render: status %d: %s", int(rc), C.GoString(C.gopbrt_multi_last_error(s.m.h)))
	}
}
