// ufo_route.cuh -- routed multi-GPU mode (SURVEY.md 8(e) variant 2; BASELINE config #4 "octant
// shard" and config #5 "one GPU per sensor with boundary merge").
//
// State is owned by space: rank = brick_owner(brick key).  Marking is parallel over rays: every
// rank walks ITS rays (a slice of one scan, or its own sensor's scan) into its own map with the
// ordinary fused kernels.  What crosses GPUs is the per-scan SET UNION only:
//   k_outbox   one warp per touched brick: the 64 miss masks of a brick another rank owns are
//              written, behind the brick key, straight into that rank's inbox (peer memory over
//              NVLink -- plain coalesced stores on a mapped peer pointer; no staging, no NCCL on
//              the data path) and cleared locally;  hit voxels were forwarded by K1 (mark_hit)
//   k_outbox_publish   record counts -> the owners' inbox headers
//   [barrier across the ranks: the one collective, a stream-ordered NCCL all-reduce of 4 bytes]
//   k_inbox    one warp per received brick record: find-or-create the brick, OR the 64 masks in;
//              one thread per received hit voxel
// followed by the ordinary K3 / K4 over the touched list.  With route_self the rank's own marks
// take the same way, so that the owner can apply the sensors ONE AFTER THE OTHER in sensor order
// (clamping makes the order matter): k_inbox(source s) + K3 for s = 0..G-1.
// Inbox regions are double-buffered by scan parity; the barrier of scan k+1 orders the reads of
// scan k before the writes of scan k+2.
#pragma once

#include "ufo_kernels.cuh"

namespace ufo_b200
{
__global__ void __launch_bounds__(256) k_outbox(DeviceMap M)
{
	const uint32_t lane = threadIdx.x & 31;
	constexpr uint32_t FULL = 0xffffffffu;
	RouteTable* R = M.route;
	if (__ldg(&M.ctr->overflow) & ~4u) return;
	const uint32_t n = __ldg(&M.ctr->n_touched);
	for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += (gridDim.x * blockDim.x) >> 5) {
		const uint32_t brick = M.touched[e];
		const unsigned long long key = M.brick_key[brick];
		const uint32_t owner = brick_owner(key, M.route_world);
		unsigned long long* mp = M.mask_base + (size_t)M.touched_mi[e] * 64 + 2 * lane;
		const ulonglong2 mk = *reinterpret_cast<const ulonglong2*>(mp);
		if (owner == M.route_rank && !M.route_self) {
			// own brick, marks stay: it opens the touched list of the apply phase (scan id + 1),
			// which reads every mask from miss_mask -- masks marked into the scan volume move there
			if (M.dense) {
				*reinterpret_cast<ulonglong2*>(M.miss_mask + (size_t)brick * 64 + 2 * lane) = mk;
				*reinterpret_cast<ulonglong2*>(mp) = make_ulonglong2(0ull, 0ull);
			}
			if (lane == 0) {
				M.brick_stamp[brick] = M.scan_id + 1u;
				const uint32_t i = atomicAdd(&M.ctr->n_touched_alt, 1u);
				M.touched_alt[i] = brick;
				M.touched_alt_mi[i] = brick;
			}
			continue;
		}
		if (!__any_sync(FULL, (mk.x | mk.y) != 0ull)) continue;
		uint32_t slot = 0;
		if (lane == 0) slot = atomicAdd(&R->out_miss[owner], 1u);
		slot = __shfl_sync(FULL, slot, 0);
		if (slot >= R->cap_m) {
			if (lane == 0) atomicOr(&R->overflow, 2u);
			continue;
		}
		unsigned long long* dst = R->out[owner].miss + (size_t)slot * kMissRecWords;
		if (lane == 0) dst[0] = key;
		dst[1 + 2 * lane] = mk.x;
		dst[2 + 2 * lane] = mk.y;
		*reinterpret_cast<ulonglong2*>(mp) = make_ulonglong2(0ull, 0ull);
	}
}

__global__ void k_outbox_publish(DeviceMap M)
{
	RouteTable* R = M.route;
	const uint32_t d = threadIdx.x;
	if (d < M.route_world) {
		const uint32_t nm = min(R->out_miss[d], R->cap_m), nh = min(R->out_hit[d], R->cap_h);
		R->out[d].hdr[0] = nm;
		R->out[d].hdr[1] = nh;
		R->out_miss[d] = 0;
		R->out_hit[d] = 0;
	}
	if (d == 0) {
		// the apply phase starts from the own bricks that kept local marks
		M.ctr->n_touched = M.ctr->n_touched_alt;
		M.ctr->n_touched_alt = 0;
	}
	__threadfence_system();
}

// applies the records of sources [first, first + count)
__global__ void __launch_bounds__(256) k_inbox(DeviceMap M, uint32_t first, uint32_t count)
{
	const uint32_t lane = threadIdx.x & 31;
	constexpr uint32_t FULL = 0xffffffffu;
	const RouteTable* R = M.route;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = first; s < first + count; ++s) {
		const RouteBox box = R->in[s];
		const uint32_t nm = ld_volatile_u32(&box.hdr[0]), nh = ld_volatile_u32(&box.hdr[1]);
		for (uint32_t i = warp; i < nm; i += warps) {
			const unsigned long long* src = box.miss + (size_t)i * kMissRecWords;
			uint32_t slot = kNone;
			if (lane == 0) {
				slot = brick_find_or_create(M, src[0]);
				if (slot != kNone) touch_brick(M, slot);
			}
			slot = __shfl_sync(FULL, slot, 0);
			if (slot == kNone) continue;
			const unsigned long long m0 = src[1 + 2 * lane], m1 = src[2 + 2 * lane];
			unsigned long long* mp = M.miss_mask + (size_t)slot * 64 + 2 * lane;
			if (m0) atomicOr(mp, m0);
			if (m1) atomicOr(mp + 1, m1);
		}
		for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += gridDim.x * blockDim.x) {
			uint32_t x, y, z;
			unpack_key(box.hit[i], x, y, z);
			const uint32_t brick = brick_find_or_create(M, pack_key(x >> 4, y >> 4, z >> 4));
			if (brick == kNone) continue;
			touch_brick(M, brick);
			atomicOr(&M.hit_mask[(size_t)brick * 64 + morton2(x >> 2, y >> 2, z >> 2)], 1ull << linear2(x, y, z));
		}
	}
}

}  // namespace ufo_b200
