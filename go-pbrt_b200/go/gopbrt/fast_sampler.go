// FastStratified is the Go half of GOPBRT_MODE_FAST: a pbrt.Sampler whose every (pixel, sample) is an independent stream,
// so that the GPU may split a pixel's samples across GPUs and lane groups (include/gopbrt_cuda.h) and the CPU renderer, run
// with this sampler, draws the very same numbers.  SOURCE ONLY (no Go toolchain in this image, SURVEY.md §0.1); the
// arithmetic below is what go-pbrt_b200/csrc/gp_render.cuh (get1d / get2d / generate_lane) implements; the repository's CPU
// checker restates the same sampler and the two are tested against each other bit for bit (tests/test_gpu_parity.py).
//
// Differences to sampler.Stratified (pkg/sampler/stratified.go), all confined to the sampler:
//   - StartNextSample reseeds the generator: rng.SetSequence(pixelIndex*spp + sampleIndex) (pixelIndex is row-major over
//     the film's CroppedPixelBounds);
//   - the i-th sampled 1-D dimension of sample s is stratum kensler(s, spp, hash(pixelIndex, i)) — a stateless
//     permutation instead of a shuffled table;
//   - the first nSampledDimensions 2-D requests return (0, 0), as Stratified's never-written 2-D tables do
//     (sampling.go:112-127 assigns into a copy);
//   - everything past the sampled dimensions comes from the reseeded generator, as in PixelSampler (pixel.go:60-80).
// The reference's sample-count quirk is kept: StartNextSample increments first, so samples 1..spp-1 run (pixel.go:48-52).
package gopbrt

import (
	"math"

	"github.com/ssttuu/go-pbrt/pkg/pbrt"
)

type FastStratified struct {
	xSamples, ySamples int32
	jitter             bool
	nSampledDimensions int
	filmX0, filmY0     int64 // CroppedPixelBounds.Min
	filmWidth          int64 // CroppedPixelBounds extent in x

	rng        *pbrt.RandomNumberGenerator
	pixelIndex uint64
	sampleIdx  int32
	cur1, cur2 int
}

func NewFastStratified(xSamples, ySamples int32, jitter bool, nSampledDimensions int, cropped *pbrt.Bounds2i) *FastStratified {
	return &FastStratified{xSamples: xSamples, ySamples: ySamples, jitter: jitter, nSampledDimensions: nSampledDimensions,
		filmX0: cropped.Min.X, filmY0: cropped.Min.Y, filmWidth: cropped.Max.X - cropped.Min.X,
		rng: pbrt.NewRandomNumberGenerator()}
}

func (s *FastStratified) GetSamplesPerPixel() int32 { return s.xSamples * s.ySamples }
func (s *FastStratified) RoundCount(n int32) int32   { return n }
func (s *FastStratified) Request2DArray(n int32)     {}
func (s *FastStratified) Get1DArray(n int32) []float64      { return nil }
func (s *FastStratified) Get2DArray(n int32) []pbrt.Point2f { return nil }

// Clone ignores the tile seed: streams are keyed by (pixel, sample), not by tile (integrator.go:318-328).
func (s *FastStratified) Clone(seed uint64) pbrt.Sampler {
	c := *s
	c.rng = pbrt.NewRandomNumberGenerator()
	return &c
}

func (s *FastStratified) StartPixel(p *pbrt.Point2i) {
	s.pixelIndex = uint64((p.Y-s.filmY0)*s.filmWidth + (p.X - s.filmX0))
	s.sampleIdx = 0
}

func (s *FastStratified) StartNextSample() bool {
	s.cur1, s.cur2 = 0, 0
	s.sampleIdx++
	spp := s.GetSamplesPerPixel()
	if s.sampleIdx < spp {
		s.rng.SetSequence(s.pixelIndex*uint64(spp) + uint64(s.sampleIdx))
	}
	return s.sampleIdx < spp
}

func (s *FastStratified) Get1D() float64 {
	if s.cur1 < s.nSampledDimensions {
		spp := uint32(s.GetSamplesPerPixel())
		j := kenslerPermute(uint32(s.sampleIdx), spp, hashU32(s.pixelIndex, uint32(s.cur1)))
		s.cur1++
		delta := 0.5
		if s.jitter {
			delta = s.rng.UniformFloat()
		}
		return math.Min((float64(j)+delta)*(1.0/float64(spp)), oneMinusEpsilon)
	}
	return s.rng.UniformFloat()
}

func (s *FastStratified) Get2D() *pbrt.Point2f {
	if s.cur2 < s.nSampledDimensions {
		s.cur2++
		return &pbrt.Point2f{}
	}
	x := s.rng.UniformFloat()
	y := s.rng.UniformFloat()
	return &pbrt.Point2f{X: x, Y: y}
}

// GetCameraSample in the reference's order: Get2D pFilm, Get2D pLens, Get1D time (pkg/sampler/sampler.go:75-80).
func (s *FastStratified) GetCameraSample(pRaster *pbrt.Point2i) *pbrt.CameraSample {
	pFilm := pbrt.NewPoint2fFromPoint2i(pRaster).Add(s.Get2D())
	pLens := s.Get2D()
	return pbrt.NewCameraSample(pFilm, pLens, s.Get1D())
}

const oneMinusEpsilon = 0.99999999999999989 // math.OneMinusEpsilon of pkg/math (0x3fefffffffffffff)

// Kensler's stateless permutation of [0, l): position i under key p ("Correlated Multi-Jittered Sampling", 2013).
func kenslerPermute(i, l, p uint32) uint32 {
	w := l - 1
	w |= w >> 1
	w |= w >> 2
	w |= w >> 4
	w |= w >> 8
	w |= w >> 16
	for {
		i ^= p
		i *= 0xe170893d
		i ^= p >> 16
		i ^= (i & w) >> 4
		i ^= p >> 8
		i *= 0x0929eb3f
		i ^= p >> 23
		i ^= (i & w) >> 1
		i *= 1 | p>>27
		i *= 0x6935fa69
		i ^= (i & w) >> 11
		i *= 0x74dcb303
		i ^= (i & w) >> 2
		i *= 0x9e501cc3
		i ^= (i & w) >> 2
		i *= 0xc860a3df
		i &= w
		i ^= i >> 5
		if i < l {
			break
		}
	}
	return (i + p) % l
}

func hashU32(a uint64, b uint32) uint32 {
	x := a*0x9E3779B97F4A7C15 + uint64(b)*0xD1B54A32D192ED03 + 0x632BE59BD9B4E019
	x ^= x >> 32
	x *= 0xD6E8FEB86659FD93
	x ^= x >> 32
	x *= 0xD6E8FEB86659FD93
	x ^= x >> 32
	return uint32(x)
}
