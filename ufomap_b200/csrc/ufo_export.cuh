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

struct BrickInfo {
	uint32_t size;   // bytes of the brick's stream (payload size when the brick is a leaf)
	uint32_t flags;  // bit 0: has children; bit 1: not part of the map (alias collector)
	float occ;       // leaf payload when bit 0 is clear
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

// offsets == nullptr: size pass (fills info); otherwise the stream of every brick that has
// children is written to out + offsets[brick].
template <bool COLOR>
__global__ void __launch_bounds__(64) k_brick_stream(DeviceMap M, uint32_t n_bricks, int pruned, BrickInfo* info,
                                                     const unsigned long long* offsets, uint8_t* out)
{
	extern __shared__ uint8_t s_buf[];
	__shared__ uint32_t s_blk_size[64], s_blk_rgb[64], s_d3_size[8], s_d3_rgb[8], s_d3_off[8];
	__shared__ float s_blk_occ[64], s_d3_occ[8];
	__shared__ uint8_t s_blk_has[64], s_d3_has[8], s_d3_mask[8];
	__shared__ uint32_t s_brick_size, s_brick_has;
	constexpr uint32_t P = COLOR ? 7u : 4u;
	const uint32_t brick = blockIdx.x, t = threadIdx.x;
	if (brick >= n_bricks) return;
	{
		uint32_t x, y, z;
		unpack_key(M.brick_key[brick], x, y, z);
		if ((x | y | z) & ~(M.g.key_mask >> 4)) {  // collects out-of-tree marks, owns no voxels
			if (!offsets && t == 0) info[brick] = {0u, 2u, 0.0f, 0u};
			return;
		}
	}
	const size_t b = (size_t)brick * 64 + t;
	const uint32_t init8 = (M.meta[b] >> 16) & 0xffu;
	// octets: has children (8 voxel payloads follow) or leaf with one payload
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
		for (int o = 1; o < 8; ++o) blk_has = blk_has || oocc[o] != oocc[0] || orgb[o] != orgb[0];
	}
	const uint32_t n_oct = __popc(oct_has);
	s_blk_has[t] = blk_has ? 1 : 0;
	s_blk_size[t] = blk_has ? 1u + n_oct * 8u * P + (8u - n_oct) * P : P;
	s_blk_occ[t] = oocc[0];  // the block's payload when it is a leaf (default payload if unpruned)
	s_blk_rgb[t] = orgb[0];
	__syncthreads();
	if (t < 8) {
		bool has = false;
		uint32_t size = 1, mask = 0;
		for (int k = 0; k < 8; ++k) {
			const int c = 8 * t + k;
			has = has || s_blk_has[c];
			mask |= (uint32_t)s_blk_has[c] << k;
			size += s_blk_size[c];
			if (pruned && (s_blk_occ[c] != s_blk_occ[8 * t] || s_blk_rgb[c] != s_blk_rgb[8 * t])) has = true;
		}
		s_d3_has[t] = has ? 1 : 0;
		s_d3_mask[t] = (uint8_t)mask;
		s_d3_size[t] = has ? size : P;
		s_d3_occ[t] = s_blk_occ[8 * t];
		s_d3_rgb[t] = s_blk_rgb[8 * t];
	}
	__syncthreads();
	if (t == 0) {
		bool has = false;
		uint32_t size = 1;
		for (int j = 0; j < 8; ++j) {
			has = has || s_d3_has[j];
			s_d3_off[j] = size;
			size += s_d3_size[j];
			if (pruned && (s_d3_occ[j] != s_d3_occ[0] || s_d3_rgb[j] != s_d3_rgb[0])) has = true;
		}
		s_brick_has = has ? 1u : 0u;
		s_brick_size = has ? size : P;
		if (!offsets) info[brick] = {s_brick_size, s_brick_has, s_d3_occ[0], s_d3_rgb[0]};
	}
	__syncthreads();
	if (!offsets || !s_brick_has) return;

	// ---- build the stream in shared memory ----
	const uint32_t j = t >> 3, k = t & 7u;
	if (t == 0) {
		uint32_t mask4 = 0;
		for (int q = 0; q < 8; ++q) mask4 |= (uint32_t)s_d3_has[q] << q;
		s_buf[0] = (uint8_t)mask4;
	}
	if (s_d3_has[j]) {
		uint32_t at = s_d3_off[j];
		if (k == 0) s_buf[at] = s_d3_mask[j];
		at += 1;
		for (uint32_t q = 0; q < k; ++q) at += s_blk_size[8 * j + q];
		if (blk_has) {
			s_buf[at++] = (uint8_t)oct_has;
			for (uint32_t o = 0; o < 8; ++o) {
				if ((oct_has >> o) & 1u) {
					const float* lp = M.leaf + b * 64 + 8 * o;
					const uint32_t* cp = COLOR ? M.leaf_rgb + b * 64 + 8 * o : nullptr;
					for (int v = 0; v < 8; ++v) {
						put_payload<COLOR>(s_buf + at, lp[v], COLOR ? cp[v] : 0u);
						at += P;
					}
				} else {
					put_payload<COLOR>(s_buf + at, oocc[o], orgb[o]);
					at += P;
				}
			}
		} else {
			put_payload<COLOR>(s_buf + at, s_blk_occ[t], s_blk_rgb[t]);
		}
	} else if (k == 0) {
		put_payload<COLOR>(s_buf + s_d3_off[j], s_d3_occ[j], s_d3_rgb[j]);
	}
	__syncthreads();
	uint8_t* dst = out + offsets[brick];
	for (uint32_t i = t; i < s_brick_size; i += 64) dst[i] = s_buf[i];
}
}  // namespace ufo_b200
