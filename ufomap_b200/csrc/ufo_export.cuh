// ufo_export.cuh -- the map in the reference's file/wire format.
//
//   Octree::write / writeData     map/octree.h:833-917   (text header + node stream)
//   writeNodes / writeNodesRecurs map/occupancy_map_base.h:1457-1533
//   node payloads                 map/occupancy_map_node.h:71-75 (float), :106-110 (+3 colour bytes)
//
// The stream is a pre-order walk: one byte per inner node that has children (bit i = child i
// has children), then the children in order -- a recursive record for a child with children, a
// payload for a leaf child; the eight children of a depth-1 node are written as eight payloads
// with no mask byte.  Pre-order over an octree is Morton order, so the stream of a 16^3 brick
// (depth 4 downward) is self-contained: k_brick_stream builds it in shared memory, one CTA per
// brick, one thread per 4^3 block; the host strings the bricks together with the few bytes of
// the levels above (ufo_map.cu, export_image).
//
// Tree shape.  pruned != 0 (the default of the C ABI): the canonical minimal tree of the value
// field -- a node whose eight children are leaves with equal payload is a leaf; untouched space
// is a leaf with the default payload (0.0, no colour).  The reference collapses such nodes itself
// whenever an update passes through them (updateNode, occupancy_map_base.h:1210-1212; with
// automatic pruning off it still marks them is_leaf, octree.h:1060-1066), so its file is this one
// unless its update order left a collapsible node behind (updateParents stops at the first
// unchanged aggregate, :1126-1133).  pruned == 0: every octet that was ever touched ("initialised"
// bits of the block meta word) is written as eight voxels -- value-equivalent, larger.
#pragma once

#include "ufo_device.cuh"

namespace ufo_b200
{
constexpr int kBrickStreamMax = 1 + 8 * (1 + 8 * (1 + 64 * 7));  // colour payloads: 28 745 bytes

// Partial / truncated export (Octree::writeData(stream, bounding_volume, false, min_depth),
// octree.h:885-917): children at min_depth are written as leaves carrying their aggregate
// payload; a child whose cube misses the box is skipped.  lo/hi are getMin()/getMax() of the
// reference's AABB(min, max) (geometry/aabb.h:62-69); the test is intersects(AABB, AABB) of
// collision_checks.cpp:256-264; node centres are accumulated top-down like getChildCenter
// (octree.h:625-633) so that the comparisons see the same doubles.
struct ExportBox {
	int on;
	double lo[3], hi[3];
};

UFO_HD bool box_hits(const ExportBox& bx, const double c[3], double hs)
{
	if (!bx.on) return true;
	for (int k = 0; k < 3; ++k) {
		const double mn = dop::sub(c[k], hs), mx = dop::add(c[k], hs);
		if (!(mn <= bx.hi[k])) return false;
		if (!(bx.lo[k] <= mx)) return false;
	}
	return true;
}

UFO_HD void child_center(const double c[3], double hs, uint32_t i, double out[3])
{
	out[0] = dop::add(c[0], (i & 1u) ? hs : -hs);
	out[1] = dop::add(c[1], (i & 2u) ? hs : -hs);
	out[2] = dop::add(c[2], (i & 4u) ? hs : -hs);
}

struct BrickInfo {
	uint32_t size;   // bytes the brick contributes to the stream (0: skipped by the box)
	uint32_t flags;  // bit 0: has children (in the map); bit 1: not part of the map (alias collector);
	                 // bit 2: written as a record (size bytes at its offset), else as a payload
	float occ;       // payload when bit 2 is clear: the leaf value, or the depth-4 aggregate
	uint32_t rgb;
};

template <bool COLOR>
__device__ __forceinline__ void put_payload(uint8_t* p, float occ, uint32_t rgb)
{
	const uint32_t u = __float_as_uint(occ);
	p[0] = (uint8_t)u;
	p[1] = (uint8_t)(u >> 8);
	p[2] = (uint8_t)(u >> 16);
	p[3] = (uint8_t)(u >> 24);
	if (COLOR) {
		p[4] = (uint8_t)rgb;
		p[5] = (uint8_t)(rgb >> 8);
		p[6] = (uint8_t)(rgb >> 16);
	}
}

// One octet of block b: returns its bytes in the stream and writes them when dst != nullptr.
template <bool COLOR>
__device__ __forceinline__ uint32_t octet_record(const DeviceMap& M, size_t b, uint32_t o, bool has,
                                                 float leaf_occ, uint32_t leaf_rgb, uint32_t min_depth,
                                                 const ExportBox& bx, const double oc[3], uint8_t* dst)
{
	constexpr uint32_t P = COLOR ? 7u : 4u;
	if (!box_hits(bx, oc, M.g.half_size[1])) return 0;
	if (has && min_depth == 0) {
		uint32_t n = 0;
		for (uint32_t j = 0; j < 8; ++j) {
			double vc[3];
			child_center(oc, M.g.half_size[0], j, vc);
			if (!box_hits(bx, vc, M.g.half_size[0])) continue;
			if (dst) put_payload<COLOR>(dst + n, M.leaf[b * 64 + 8 * o + j], COLOR ? M.leaf_rgb[b * 64 + 8 * o + j] : 0u);
			n += P;
		}
		return n;
	}
	if (dst) {
		if (has) put_payload<COLOR>(dst, M.sum1[b * 8 + o], COLOR ? M.sum1_rgb[b * 8 + o] : 0u);
		else put_payload<COLOR>(dst, leaf_occ, leaf_rgb);
	}
	return P;
}

// offsets == nullptr: size pass (fills info); otherwise the record of every brick whose info has
// bit 2 set is written to out + offsets[brick].
template <bool COLOR>
__global__ void __launch_bounds__(64) k_brick_stream(DeviceMap M, uint32_t n_bricks, int pruned, uint32_t min_depth,
                                                     ExportBox bx, BrickInfo* info,
                                                     const unsigned long long* offsets, uint8_t* out)
{
	extern __shared__ uint8_t s_buf[];
	__shared__ uint32_t s_blk_size[64], s_blk_rgb[64], s_d3_size[8], s_d3_rgb[8], s_d3_off[8];
	__shared__ float s_blk_occ[64], s_d3_occ[8];
	__shared__ uint8_t s_blk_has[64], s_d3_has[8], s_d3_mask[8], s_d3_hit[8];
	__shared__ uint32_t s_brick_size, s_brick_record;
	__shared__ double s_center[3];
	constexpr uint32_t P = COLOR ? 7u : 4u;
	const uint32_t brick = blockIdx.x, t = threadIdx.x;
	if (brick >= n_bricks) return;
	uint32_t kx, ky, kz;
	unpack_key(M.brick_key[brick], kx, ky, kz);
	if ((kx | ky | kz) & ~(M.g.key_mask >> 4)) {  // collects out-of-tree marks, owns no voxels
		if (!offsets && t == 0) info[brick] = {0u, 2u, 0.0f, 0u};
		return;
	}
	if (t == 0) {
		// centre of the brick, accumulated from the root like the reference's recursion
		double c[3] = {0.0, 0.0, 0.0};
		for (uint32_t d = M.g.depth_levels; d > 4; --d) {
			const uint32_t bit = d - 5;  // brick-key bit that selects the child of the depth-d node
			const uint32_t i = ((kx >> bit) & 1u) | (((ky >> bit) & 1u) << 1) | (((kz >> bit) & 1u) << 2);
			double n[3];
			child_center(c, M.g.half_size[d - 1], i, n);
			c[0] = n[0];
			c[1] = n[1];
			c[2] = n[2];
		}
		s_center[0] = c[0];
		s_center[1] = c[1];
		s_center[2] = c[2];
	}
	__syncthreads();
	const uint32_t j3 = t >> 3, k3 = t & 7u;
	double bc[3], d3c[3], blkc[3];
	bc[0] = s_center[0];
	bc[1] = s_center[1];
	bc[2] = s_center[2];
	child_center(bc, M.g.half_size[3], j3, d3c);
	child_center(d3c, M.g.half_size[2], k3, blkc);
	const bool brick_hit = box_hits(bx, bc, M.g.half_size[4]);
	const bool d3_hit = box_hits(bx, d3c, M.g.half_size[3]);
	const bool blk_hit = box_hits(bx, blkc, M.g.half_size[2]);

	const size_t b = (size_t)brick * 64 + t;
	const uint32_t init8 = (M.meta[b] >> 16) & 0xffu;
	// octets in the map: has children, or a leaf with one payload
	uint32_t oct_has = 0;
	float oocc[8];
	uint32_t orgb[8];
#pragma unroll
	for (uint32_t o = 0; o < 8; ++o) {
		oocc[o] = 0.0f;
		orgb[o] = 0u;
		if (!((init8 >> o) & 1u)) continue;
		if (!pruned) {
			oct_has |= 1u << o;
			continue;
		}
		const float* lp = M.leaf + b * 64 + 8 * o;
		const uint32_t* cp = COLOR ? M.leaf_rgb + b * 64 + 8 * o : nullptr;
		const float v0 = lp[0];
		const uint32_t c0 = COLOR ? cp[0] : 0u;
		bool same = true;
		for (int j = 1; j < 8; ++j) same = same && lp[j] == v0 && (!COLOR || cp[j] == c0);
		if (same) {
			oocc[o] = v0;
			orgb[o] = c0;
		} else {
			oct_has |= 1u << o;
		}
	}
	bool blk_has = oct_has != 0;
	if (pruned && !blk_has) {
#pragma unroll
		for (int o = 1; o < 8; ++o) blk_has = blk_has || oocc[o] != oocc[0] || orgb[o] != orgb[0];
	}
	// the block's contribution to the stream
	uint32_t blk_size = 0;
	if (blk_hit && d3_hit && brick_hit) {
		if (blk_has && min_depth < 2) {
			blk_size = 1;
#pragma unroll
			for (uint32_t o = 0; o < 8; ++o) {
				double oc[3];
				child_center(blkc, M.g.half_size[1], o, oc);
				blk_size += octet_record<COLOR>(M, b, o, (oct_has >> o) & 1u, oocc[o], orgb[o], min_depth, bx, oc, nullptr);
			}
		} else {
			blk_size = P;
		}
	}
	s_blk_has[t] = blk_has ? 1 : 0;
	s_blk_size[t] = blk_size;
	// payload when the block is written as a leaf: its value, or its depth-2 aggregate
	s_blk_occ[t] = blk_has ? M.agg2[b].occ : oocc[0];
	s_blk_rgb[t] = blk_has ? (COLOR ? M.rgb2[b] : 0u) : orgb[0];
	__syncthreads();
	if (t < 8) {
		double c3[3];
		child_center(bc, M.g.half_size[3], t, c3);
		const bool hit = brick_hit && box_hits(bx, c3, M.g.half_size[3]);
		bool has = false;
		uint32_t size = 1, mask = 0;
		for (int k = 0; k < 8; ++k) {
			const int c = 8 * t + k;
			has = has || s_blk_has[c];
			if (min_depth < 2) mask |= (uint32_t)s_blk_has[c] << k;
			size += s_blk_size[c];
		}
		if (pruned && !has) {
			// leaf blocks only: s_blk_occ / s_blk_rgb hold their leaf payloads
			for (int k = 1; k < 8; ++k) has = has || s_blk_occ[8 * t + k] != s_blk_occ[8 * t] || s_blk_rgb[8 * t + k] != s_blk_rgb[8 * t];
		}
		s_d3_has[t] = has ? 1 : 0;
		s_d3_hit[t] = hit ? 1 : 0;
		s_d3_mask[t] = (uint8_t)mask;
		s_d3_size[t] = !hit ? 0u : ((has && min_depth < 3) ? size : P);
		s_d3_occ[t] = has ? M.brick_sum3[(size_t)brick * 8 + t].occ : s_blk_occ[8 * t];
		s_d3_rgb[t] = has ? (COLOR ? M.brick_rgb3[(size_t)brick * 8 + t] : 0u) : s_blk_rgb[8 * t];
	}
	__syncthreads();
	if (t == 0) {
		bool has = false;
		uint32_t size = 1;
		for (int j = 0; j < 8; ++j) {
			has = has || s_d3_has[j];
			s_d3_off[j] = size;
			size += s_d3_size[j];
		}
		if (pruned && !has) {
			for (int j = 1; j < 8; ++j) has = has || s_d3_occ[j] != s_d3_occ[0] || s_d3_rgb[j] != s_d3_rgb[0];
		}
		const bool record = has && min_depth < 4 && brick_hit;
		s_brick_record = record ? 1u : 0u;
		s_brick_size = !brick_hit ? 0u : (record ? size : P);
		if (!offsets) {
			const float occ = has ? M.brick_sum4[brick].occ : s_d3_occ[0];
			const uint32_t rgb = has ? (COLOR ? M.brick_rgb4[brick] : 0u) : s_d3_rgb[0];
			info[brick] = {s_brick_size, (has ? 1u : 0u) | (record ? 4u : 0u), occ, rgb};
		}
	}
	__syncthreads();
	if (!offsets || !s_brick_record) return;

	// ---- build the record in shared memory ----
	if (t == 0) {
		uint32_t mask4 = 0;
		if (min_depth < 3)
			for (int q = 0; q < 8; ++q) mask4 |= (uint32_t)s_d3_has[q] << q;
		s_buf[0] = (uint8_t)mask4;
	}
	if (s_d3_hit[j3]) {
		if (s_d3_has[j3] && min_depth < 3) {
			uint32_t at = s_d3_off[j3];
			if (k3 == 0) s_buf[at] = s_d3_mask[j3];
			at += 1;
			for (uint32_t q = 0; q < k3; ++q) at += s_blk_size[8 * j3 + q];
			if (blk_size) {
				if (blk_has && min_depth < 2) {
					s_buf[at++] = (uint8_t)(min_depth < 1 ? oct_has : 0u);
#pragma unroll
					for (uint32_t o = 0; o < 8; ++o) {
						double oc[3];
						child_center(blkc, M.g.half_size[1], o, oc);
						at += octet_record<COLOR>(M, b, o, (oct_has >> o) & 1u, oocc[o], orgb[o], min_depth, bx, oc, s_buf + at);
					}
				} else {
					put_payload<COLOR>(s_buf + at, s_blk_occ[t], s_blk_rgb[t]);
				}
			}
		} else if (k3 == 0) {
			put_payload<COLOR>(s_buf + s_d3_off[j3], s_d3_occ[j3], s_d3_rgb[j3]);
		}
	}
	__syncthreads();
	uint8_t* dst = out + offsets[brick];
	for (uint32_t i = t; i < s_brick_size; i += 64) dst[i] = s_buf[i];
}

// ---------------------------------------------------------------------------
// Filtered node read-out: the leaf iteration of the reference with state and
// bounding-volume filters (beginLeaves(occupied, free, unknown, contains = false, min_depth),
// occupancy_map_base.h:130-216; validity rule iterator/occupancy_map.h:168-207), on the value
// field: every node of depth `depth` inside the known space (existing bricks / upper nodes) whose
// own state -- occupied: value > occupied threshold, free: value < free threshold, unknown: in
// between, the value of an inner node being the maximum of its subtree -- passes the filter and
// whose cube intersects the box.  The reference returns a collapsed leaf above `depth` as one
// node; here it appears as its depth-`depth` cells (same space, same state).
// One thread per (brick, block) for depth 0..4; k_export_upper for depth >= 5.
struct NodeFilter {
	uint32_t depth;
	int occupied, free_, unknown;
	ExportBox box;
};

__device__ __forceinline__ bool node_passes(const DeviceMap& M, const NodeFilter& f, float v)
{
	const bool occ = M.occ_thr < (double)v, fre = M.free_thr > (double)v;
	return (f.occupied && occ) || (f.free_ && fre) || (f.unknown && !occ && !fre);
}

__device__ __forceinline__ void node_emit(const DeviceMap& M, const NodeFilter& f, uint32_t x, uint32_t y, uint32_t z,
                                          float v, uint32_t rgb, unsigned long long* codes, float* occ, uint32_t* rgbs,
                                          unsigned long long cap, unsigned long long* count)
{
	if (!node_passes(M, f, v)) return;
	const uint32_t d = f.depth;
	const uint32_t centre = d ? (1u << (d - 1)) : 0u, snap = ~((1u << d) - 1u);
	const Key3 k = {(x & snap) | centre, (y & snap) | centre, (z & snap) | centre};
	if (f.box.on) {
		const Vec3 c = key_to_coord(M.g, k, d);
		const double cc[3] = {c.x, c.y, c.z};
		if (!box_hits(f.box, cc, M.g.half_size[d])) return;
	}
	const unsigned long long at = atomicAdd(count, 1ull);
	if (codes && at < cap) {
		codes[at] = key_to_code(k);
		occ[at] = v;
		if (rgbs) rgbs[at] = rgb;
	}
}

__global__ void __launch_bounds__(256) k_export_nodes(DeviceMap M, uint32_t n_bricks, NodeFilter f, unsigned long long* codes,
                                                      float* occ, uint32_t* rgbs, unsigned long long cap,
                                                      unsigned long long* count)
{
	const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (blk >= (size_t)n_bricks * 64) return;
	const uint32_t brick = (uint32_t)(blk >> 6), ch = (uint32_t)(blk & 63);
	uint32_t bx, by, bz;
	unpack_key(M.brick_key[brick], bx, by, bz);
	if ((bx | by | bz) & ~(M.g.key_mask >> 4)) return;  // alias collector brick
	const uint32_t cx = (ch & 1u) | ((ch >> 2) & 2u), cy = ((ch >> 1) & 1u) | ((ch >> 3) & 2u),
	               cz = ((ch >> 2) & 1u) | ((ch >> 4) & 2u);
	const uint32_t x0 = (bx << 4) | (cx << 2), y0 = (by << 4) | (cy << 2), z0 = (bz << 4) | (cz << 2);
	const uint32_t d = f.depth;
	if (d == 4) {
		if (ch == 0) node_emit(M, f, x0, y0, z0, M.brick_sum4[brick].occ, M.color ? M.brick_rgb4[brick] : 0u, codes, occ, rgbs, cap, count);
		return;
	}
	if (d == 3) {
		if ((ch & 7u) == 0) {
			const uint32_t j = ch >> 3;
			node_emit(M, f, x0, y0, z0, M.brick_sum3[(size_t)brick * 8 + j].occ, M.color ? M.brick_rgb3[(size_t)brick * 8 + j] : 0u, codes,
			          occ, rgbs, cap, count);
		}
		return;
	}
	const uint32_t meta = M.meta[blk];
	const bool init = (meta >> 16) != 0;
	if (d == 2) {
		node_emit(M, f, x0, y0, z0, init ? M.agg2[blk].occ : 0.0f, (init && M.color) ? M.rgb2[blk] : 0u, codes, occ, rgbs, cap, count);
		return;
	}
	for (uint32_t o = 0; o < 8; ++o) {
		const bool oi = init && ((meta >> (16 + o)) & 1u);
		const uint32_t ox = x0 | ((o & 1u) << 1), oy = y0 | (o & 2u), oz = z0 | ((o & 4u) >> 1);
		if (d == 1) {
			node_emit(M, f, ox, oy, oz, oi ? M.sum1[blk * 8 + o] : 0.0f, (oi && M.color) ? M.sum1_rgb[blk * 8 + o] : 0u, codes, occ, rgbs,
			          cap, count);
			continue;
		}
		for (uint32_t j = 0; j < 8; ++j) {
			const float v = oi ? M.leaf[blk * 64 + 8 * o + j] : 0.0f;
			const uint32_t c = (oi && M.color) ? M.leaf_rgb[blk * 64 + 8 * o + j] : 0u;
			node_emit(M, f, ox | (j & 1u), oy | ((j >> 1) & 1u), oz | (j >> 2), v, c, codes, occ, rgbs, cap, count);
		}
	}
}

__global__ void __launch_bounds__(256) k_export_upper(DeviceMap M, uint32_t n_upper, NodeFilter f, unsigned long long* codes,
                                                      float* occ, uint32_t* rgbs, unsigned long long cap,
                                                      unsigned long long* count)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_upper) return;
	uint32_t x, y, z;
	const unsigned long long key = M.up_key[i];
	unpack_key(key, x, y, z);
	const uint32_t d = x >> 16;  // depth tag (upper_key)
	if (d != f.depth) return;
	x &= 0xffffu;
	node_emit(M, f, x << d, y << d, z << d, M.up_agg[i].occ, M.color ? M.up_rgb[i] : 0u, codes, occ, rgbs, cap, count);
}

// ---------------------------------------------------------------------------
// setValueVolume(AABB, value, min_depth)   occupancy_map_base.h:492-518, :986-1031
// ---------------------------------------------------------------------------
// The reference descends from the root and, at every level down to min_depth, keeps the children
// whose cube intersects the box; the nodes reached at min_depth are collapsed and set.  On the
// value field: a voxel is set iff every ancestor cube from depth L-1 down to min_depth (itself at
// min_depth 0) intersects the box.  k_volume_mark evaluates exactly that chain -- centres
// accumulated top-down like getChildCenter -- for one candidate brick per CTA, one 4^3 block per
// thread, and leaves the result as a bit mask in the block's miss mask; the set itself is the K3
// update kernel in SET mode, followed by the usual aggregate passes.
struct VolumeArgs {
	ExportBox box;
	uint32_t min_depth;       // 0..4
	uint32_t bx0, by0, bz0;   // first candidate brick (brick coordinates)
	uint32_t nx, ny, nz;      // candidate bricks per axis
};

__global__ void __launch_bounds__(64) k_volume_mark(DeviceMap M, VolumeArgs v)
{
	__shared__ double s_center[3];
	__shared__ uint32_t s_slot;
	const uint32_t t = threadIdx.x;
	const uint32_t ix = blockIdx.x % v.nx, iy = (blockIdx.x / v.nx) % v.ny, iz = blockIdx.x / (v.nx * v.ny);
	const uint32_t kx = v.bx0 + ix, ky = v.by0 + iy, kz = v.bz0 + iz;
	if (t == 0) {
		double c[3] = {0.0, 0.0, 0.0};
		bool ok = true;
		for (uint32_t d = M.g.depth_levels; d > 4; --d) {
			const uint32_t bit = d - 5;
			const uint32_t i = ((kx >> bit) & 1u) | (((ky >> bit) & 1u) << 1) | (((kz >> bit) & 1u) << 2);
			double n[3];
			child_center(c, M.g.half_size[d - 1], i, n);
			c[0] = n[0];
			c[1] = n[1];
			c[2] = n[2];
			ok = ok && box_hits(v.box, c, M.g.half_size[d - 1]);  // cube of the depth d-1 ancestor
		}
		uint32_t slot = kNone;
		if (ok) {
			slot = brick_find_or_create(M, pack_key(kx, ky, kz));
			if (slot != kNone) touch_brick(M, slot);
		}
		s_slot = slot;
		s_center[0] = c[0];
		s_center[1] = c[1];
		s_center[2] = c[2];
	}
	__syncthreads();
	const uint32_t slot = s_slot;
	if (slot == kNone) return;
	unsigned long long mask = 0ull;
	if (v.min_depth >= 4) {
		mask = ~0ull;
	} else {
		double bc[3] = {s_center[0], s_center[1], s_center[2]}, d3c[3], blkc[3];
		child_center(bc, M.g.half_size[3], t >> 3, d3c);
		child_center(d3c, M.g.half_size[2], t & 7u, blkc);
		if (box_hits(v.box, d3c, M.g.half_size[3])) {
			if (v.min_depth == 3) {
				mask = ~0ull;
			} else if (box_hits(v.box, blkc, M.g.half_size[2])) {
				if (v.min_depth == 2) {
					mask = ~0ull;
				} else {
					for (uint32_t o = 0; o < 8; ++o) {
						double oc[3];
						child_center(blkc, M.g.half_size[1], o, oc);
						if (!box_hits(v.box, oc, M.g.half_size[1])) continue;
						const uint32_t x0 = (o & 1u) << 1, y0 = o & 2u, z0 = (o & 4u) >> 1;
						if (v.min_depth == 1) {
							mask |= octet_mask_of(x0, y0, z0);
							continue;
						}
						for (uint32_t j = 0; j < 8; ++j) {
							double vc[3];
							child_center(oc, M.g.half_size[0], j, vc);
							if (box_hits(v.box, vc, M.g.half_size[0]))
								mask |= 1ull << linear2(x0 + (j & 1u), y0 + ((j >> 1) & 1u), z0 + (j >> 2));
						}
					}
				}
			}
		}
	}
	// the masks of a block belong to one thread here, and no scan is in flight
	if (mask) M.miss_mask[(size_t)slot * 64 + t] |= mask;
}
}  // namespace ufo_b200
