// ufo_index.cuh -- coordinate <-> Key <-> Code indexing, BBX clipping and the
// Amanatides-Woo voxel walk, shared by host (C-ABI helpers, computeRay) and
// device (integration kernels).
//
// Bit-exact counterparts of the reference (paths under /root/reference/ufomap/include/ufo):
//   toKey      map/octree.h:317-324      toCoord       map/octree.h:372-383
//   Code(Key)  map/code.h:183-192        Code::toKey   map/code.h:205-230
//   moveLineInside / inBBX / getIntersection           map/octree.h:1240-1332
//   computeRayInit / computeRayTakeStep                map/octree.h:1192-1233
//   Vector3::norm / minElementIndex / min              math/vector3.h:203,241-251
//
// All geometry is IEEE double with round-to-nearest and NO fused multiply-add:
// the reference is built without FMA contraction and accumulates t_max by
// repeated addition, so a contracted or closed-form variant can flip a tie
// (SURVEY.md section 7 "hard parts").  Every double operation below therefore goes
// through the dop:: helpers, which map to __dadd_rn/__dmul_rn/... on the device
// (never contracted by nvcc) and to plain operators on the host (this file is
// compiled with -fmad=false / -ffp-contract=off on both sides).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define UFO_HD __host__ __device__ __forceinline__
#else
#define UFO_HD inline
#endif

namespace ufo_b200
{
namespace dop
{
UFO_HD double add(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
UFO_HD double sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}
UFO_HD double mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
UFO_HD double div(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __ddiv_rn(a, b);
#else
	return a / b;
#endif
}
UFO_HD double sqrt(double a)
{
#if defined(__CUDA_ARCH__)
	return __dsqrt_rn(a);
#else
	return ::sqrt(a);
#endif
}
}  // namespace dop

struct Vec3 {
	double x, y, z;
	UFO_HD double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
	UFO_HD double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

UFO_HD Vec3 vsub(Vec3 a, Vec3 b) { return {dop::sub(a.x, b.x), dop::sub(a.y, b.y), dop::sub(a.z, b.z)}; }
UFO_HD Vec3 vadd(Vec3 a, Vec3 b) { return {dop::add(a.x, b.x), dop::add(a.y, b.y), dop::add(a.z, b.z)}; }
UFO_HD Vec3 vscale(Vec3 a, double s) { return {dop::mul(a.x, s), dop::mul(a.y, s), dop::mul(a.z, s)}; }
UFO_HD Vec3 vdiv(Vec3 a, double s) { return {dop::div(a.x, s), dop::div(a.y, s), dop::div(a.z, s)}; }
// (x*x + y*y) + z*z, the association the reference's squaredNorm uses
UFO_HD double vsqnorm(Vec3 a)
{
	return dop::add(dop::add(dop::mul(a.x, a.x), dop::mul(a.y, a.y)), dop::mul(a.z, a.z));
}
UFO_HD double vnorm(Vec3 a) { return dop::sqrt(vsqnorm(a)); }

// Rigid frame of a cloud: unit quaternion (w, x, y, z) + translation.
struct Frame {
	double qw, qx, qy, qz;
	double tx, ty, tz;
};

// Hamilton product a*b with the term order of the reference's Quaternion::operator*
// (math/quaternion.h:253-259); every sum is evaluated left to right.
struct Quat {
	double w, x, y, z;
};
UFO_HD Quat quat_mul(Quat a, Quat b)
{
	using namespace dop;
	Quat r;
	r.w = sub(sub(sub(mul(a.w, b.w), mul(a.x, b.x)), mul(a.y, b.y)), mul(a.z, b.z));
	r.x = add(add(sub(mul(a.y, b.z), mul(b.y, a.z)), mul(a.w, b.x)), mul(b.w, a.x));
	r.y = add(add(sub(mul(a.z, b.x), mul(b.z, a.x)), mul(a.w, b.y)), mul(b.w, a.y));
	r.z = add(add(sub(mul(a.x, b.y), mul(b.x, a.y)), mul(a.w, b.z)), mul(b.w, a.z));
	return r;
}

// Pose6::transform (math/pose6.h:115-125): the sandwich product q (0,v) conj(q) of
// Quaternion::rotate (math/quaternion.h:277-286), then += translation.  Not simplified:
// the zero-weight terms stay so that signed zeros and rounding match the reference.
UFO_HD Vec3 frame_transform(const Frame& f, Vec3 v)
{
	Quat q{f.qw, f.qx, f.qy, f.qz};
	Quat r = quat_mul(quat_mul(q, Quat{0.0, v.x, v.y, v.z}), Quat{f.qw, -f.qx, -f.qy, -f.qz});
	return {dop::add(r.x, f.tx), dop::add(r.y, f.ty), dop::add(r.z, f.tz)};
}

struct Key3 {
	uint32_t x, y, z;
};
UFO_HD bool operator==(Key3 a, Key3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
UFO_HD bool operator!=(Key3 a, Key3 b) { return !(a == b); }

// Geometry constants of one map (octree.h:922-943).
struct Geometry {
	double resolution;
	double resolution_factor;  // 1.0 / resolution
	double half_size[24];      // half_size[d]: half edge of a depth-d node; size(d) = half_size[d+1]
	uint32_t depth_levels;     // root depth L
	uint32_t max_value;        // 2^(L-1)
	uint32_t key_mask;         // (1<<L)-1: the tree consumes only L key bits per axis
};

inline Geometry make_geometry(double resolution, uint32_t depth_levels)
{
	Geometry g{};
	g.resolution = resolution;
	g.resolution_factor = 1.0 / resolution;
	g.depth_levels = depth_levels;
	g.max_value = 1u << (depth_levels - 1);
	g.key_mask = (depth_levels >= 32) ? 0xffffffffu : ((1u << depth_levels) - 1u);
	g.half_size[0] = resolution / 2.0;
	g.half_size[1] = resolution;
	for (uint32_t i = 2; i < 24; ++i) g.half_size[i] = g.half_size[i - 1] * 2.0;
	return g;
}

UFO_HD double node_half(const Geometry& g, uint32_t d) { return g.half_size[d]; }
UFO_HD double node_size(const Geometry& g, uint32_t d) { return g.half_size[d + 1]; }

// ---------------------------------------------------------------------------
// Morton helpers.  21 bits per axis, x -> bit 0, y -> bit 1, z -> bit 2.
// ---------------------------------------------------------------------------
UFO_HD uint64_t spread3(uint32_t v)
{
	uint64_t x = v & 0x1fffffu;
	x = (x | (x << 32)) & 0x001f00000000ffffull;
	x = (x | (x << 16)) & 0x001f0000ff0000ffull;
	x = (x | (x << 8)) & 0x100f00f00f00f00full;
	x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
	x = (x | (x << 2)) & 0x1249249249249249ull;
	return x;
}

UFO_HD uint32_t compact3(uint64_t c)
{
	uint64_t x = c & 0x1249249249249249ull;
	x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
	x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
	x = (x ^ (x >> 8)) & 0x001f0000ff0000ffull;
	x = (x ^ (x >> 16)) & 0x001f00000000ffffull;
	x = (x ^ (x >> 32)) & 0x1fffffull;
	return (uint32_t)x;
}

UFO_HD uint64_t key_to_code(Key3 k) { return spread3(k.x) | (spread3(k.y) << 1) | (spread3(k.z) << 2); }
UFO_HD Key3 code_to_key(uint64_t c) { return {compact3(c), compact3(c >> 1), compact3(c >> 2)}; }

// Morton index of the low two bits of a key (position of a voxel inside its 4^3
// block, or of a block inside its 16^3 brick): 6 bits.
UFO_HD uint32_t morton2(uint32_t x, uint32_t y, uint32_t z)
{
	uint32_t a = (x & 1u) | ((x & 2u) << 2);
	uint32_t b = (y & 1u) | ((y & 2u) << 2);
	uint32_t c = (z & 1u) | ((z & 2u) << 2);
	return a | (b << 1) | (c << 2);
}

// Per-scan hit/miss masks of a 4^3 block use the cheap linear bit order
// x + 4y + 16z (three ops in the ray walk); the leaf array itself is in Morton order.
UFO_HD uint32_t linear2(uint32_t x, uint32_t y, uint32_t z)
{
	return (x & 3u) | ((y & 3u) << 2) | ((z & 3u) << 4);
}

// the 8 mask bits of octet o (Morton child index of the depth-1 node inside the
// block), returned in Morton order of the voxels inside the octet
UFO_HD uint32_t octet_bits8(unsigned long long mask, uint32_t o)
{
	uint32_t base = ((o & 1u) << 1) | ((o & 2u) << 2) | ((o & 4u) << 3);
	uint32_t s = (uint32_t)(mask >> base);
	return (s & 3u) | ((s >> 2) & 0xcu) | ((s >> 12) & 0x30u) | ((s >> 14) & 0xc0u);
}

// all 8 voxels of the octet that contains key (x, y, z), in mask bit order
UFO_HD unsigned long long octet_mask_of(uint32_t x, uint32_t y, uint32_t z)
{
	uint32_t base = (x & 2u) | ((y & 2u) << 2) | ((z & 2u) << 4);
	return 0x330033ull << base;
}

// ---------------------------------------------------------------------------
// coordinate <-> key
// ---------------------------------------------------------------------------
UFO_HD uint32_t coord_to_key(const Geometry& g, double c, uint32_t depth)
{
	int kv = (int)floor(dop::mul(g.resolution_factor, c));
	if (0 == depth) return (uint32_t)kv + g.max_value;
	int snapped = (int)(((uint32_t)(kv >> depth)) << depth);
	return (uint32_t)(snapped + (1 << (depth - 1))) + g.max_value;
}

UFO_HD Key3 point_to_key(const Geometry& g, Vec3 p, uint32_t depth)
{
	return {coord_to_key(g, p.x, depth), coord_to_key(g, p.y, depth), coord_to_key(g, p.z, depth)};
}

UFO_HD double key_to_coord1(const Geometry& g, uint32_t key, uint32_t depth)
{
	if (g.depth_levels == depth) return 0.0;
	double divider = (double)(1u << depth);
	double q = floor(dop::div(dop::sub((double)key, (double)g.max_value), divider));
	return dop::mul(dop::add(q, 0.5), node_size(g, depth));
}

UFO_HD Vec3 key_to_coord(const Geometry& g, Key3 k, uint32_t depth)
{
	return {key_to_coord1(g, k.x, depth), key_to_coord1(g, k.y, depth), key_to_coord1(g, k.z, depth)};
}

// ---------------------------------------------------------------------------
// BBX clipping
// ---------------------------------------------------------------------------
UFO_HD bool in_bbx(Vec3 p, double lo, double hi)
{
	return lo <= p.x && hi >= p.x && lo <= p.y && hi >= p.y && lo <= p.z && hi >= p.z;
}

UFO_HD bool in_bbx_face(Vec3 p, int axis, double lo, double hi)
{
	double a = p[(axis + 1) % 3], b = p[(axis + 2) % 3];
	return a > lo && a < hi && b > lo && b < hi;
}

UFO_HD bool plane_crossing(double d1, double d2, Vec3 p1, Vec3 p2, Vec3& hit)
{
	if (0 <= dop::mul(d1, d2)) return false;
	double f = dop::div(-d1, dop::sub(d2, d1));
	hit = vadd(p1, vscale(vsub(p2, p1), f));
	return true;
}

// Returns false when the segment lies outside the map; otherwise clips both
// ends onto the map cube [-half(L), +half(L)]^3.
UFO_HD bool move_line_inside(const Geometry& g, Vec3& o, Vec3& e)
{
	const double hi = node_half(g, g.depth_levels), lo = -hi;
	for (int i = 0; i < 3; ++i) {
		if ((o[i] < lo && e[i] < lo) || (o[i] > hi && e[i] > hi)) return false;
	}
	if (in_bbx(o, lo, hi) && in_bbx(e, lo, hi)) return true;

	int hits = 0;
	Vec3 hit[2];
	for (int i = 0; i < 3 && hits < 2; ++i) {
		if (plane_crossing(dop::sub(o[i], lo), dop::sub(e[i], lo), o, e, hit[hits]) &&
		    in_bbx_face(hit[hits], i, lo, hi))
			++hits;
	}
	for (int i = 0; i < 3 && hits < 2; ++i) {
		if (plane_crossing(dop::sub(o[i], hi), dop::sub(e[i], hi), o, e, hit[hits]) &&
		    in_bbx_face(hit[hits], i, lo, hi))
			++hits;
	}
	if (1 == hits) {
		if (in_bbx(o, lo, hi)) e = hit[0];
		else o = hit[0];
	} else if (2 == hits) {
		double a = dop::add(vsqnorm(vsub(o, hit[0])), vsqnorm(vsub(e, hit[1])));
		double b = dop::add(vsqnorm(vsub(o, hit[1])), vsqnorm(vsub(e, hit[0])));
		if (a <= b) {
			o = hit[0];
			e = hit[1];
		} else {
			o = hit[1];
			e = hit[0];
		}
	}
	return true;
}

// ---------------------------------------------------------------------------
// Voxel walk state
// ---------------------------------------------------------------------------
struct Walk {
	Key3 cur, end;
	int sx, sy, sz;       // key step per axis (+-(1<<depth) or 0)
	double tx, ty, tz;    // t_max
	double dx, dy, dz;    // t_delta
	bool same;            // start voxel == end voxel (walk state not initialised)
};

UFO_HD void walk_axis(double dir, double centre_minus_origin, double half, double size, uint32_t depth,
                      int& step, double& t_delta, double& t_max)
{
	if (0 < dir) {
		step = (int)(1u << depth);
		t_delta = dop::div(size, fabs(dir));
		t_max = dop::div(dop::add(centre_minus_origin, half), dir);
	} else if (0 > dir) {
		step = -(int)(1u << depth);
		t_delta = dop::div(size, fabs(dir));
		t_max = dop::div(dop::sub(centre_minus_origin, half), dir);
	} else {
		step = 0;
		t_delta = DBL_MAX;
		t_max = DBL_MAX;
	}
}

// from -> to with unit direction dir (supplied by the caller exactly as the
// reference computes it).
UFO_HD void walk_init(const Geometry& g, Vec3 from, Vec3 to, Vec3 dir, uint32_t depth, Walk& w)
{
	w.cur = point_to_key(g, from, depth);
	w.end = point_to_key(g, to, depth);
	w.same = (w.cur == w.end);
	if (w.same) return;
	const double size = node_size(g, depth), half = node_half(g, depth);
	Vec3 border = vsub(key_to_coord(g, w.cur, depth), from);
	walk_axis(dir.x, border.x, half, size, depth, w.sx, w.dx, w.tx);
	walk_axis(dir.y, border.y, half, size, depth, w.sy, w.dy, w.ty);
	walk_axis(dir.z, border.z, half, size, depth, w.sz, w.dz, w.tz);
}

// argmin with the reference's tie-break (<=, x before y before z)
UFO_HD void walk_step(Walk& w)
{
	if (w.tx <= w.ty) {
		if (w.tx <= w.tz) {
			w.cur.x += (uint32_t)w.sx;
			w.tx = dop::add(w.tx, w.dx);
		} else {
			w.cur.z += (uint32_t)w.sz;
			w.tz = dop::add(w.tz, w.dz);
		}
	} else {
		if (w.ty <= w.tz) {
			w.cur.y += (uint32_t)w.sy;
			w.ty = dop::add(w.ty, w.dy);
		} else {
			w.cur.z += (uint32_t)w.sz;
			w.tz = dop::add(w.tz, w.dz);
		}
	}
}

UFO_HD double walk_tmin(const Walk& w)
{
	double a = w.ty < w.tx ? w.ty : w.tx;  // std::min(std::min(x, y), z)
	return w.tz < a ? w.tz : a;
}

}  // namespace ufo_b200
