// export_gpu.go — accessors for accelerator.BVH (bvh.go:216-221): the primitive list NewBVH was given and its
// maxPrimsInNode.  The backend builds its own tree from them (any tree that never culls a primitive whose own
// bound passes reproduces BVH.Intersect / IntersectP, bvh.go:659-765).  SOURCE ONLY, see pkg/gpudesc/desc.go.
package accelerator

import "github.com/ssttuu/go-pbrt/pkg/pbrt"

func (b *BVH) Primitives() []pbrt.Primitive { return b.primitives }
func (b *BVH) MaxPrimsInNode() int         { return int(b.maxPrimsInNode) }
