// ufo_update.cuh -- K3: hit-then-miss float log-odds update of the marked voxels
// (updateOccupancy, occupancy_map_base.h:1139-1145) and the depth 1-4 aggregates
// (updateNode, :1179-1224), over the bricks the scan touched.
#pragma once

#include "ufo_device.cuh"

namespace ufo_b200
{
constexpr int kStatSlots = 64;  // per-scan counters are spread over slots to avoid same-address atomics

__device__ __forceinline__ uint32_t rms_rgb(const uint32_t* c, int n)
{
	// getAverageColor (occupancy_map_color.cpp:200-222) over the set colours
	double s[3] = {0, 0, 0};
	int cnt = 0;
	for (int i = 0; i < n; ++i) {
		if (!c[i]) continue;
		for (int k = 0; k < 3; ++k) {
			double v = (double)((c[i] >> (8 * k)) & 0xffu);
			s[k] = dop::add(s[k], dop::mul(v, v));
		}
		++cnt;
	}
	if (!cnt) return 0;
	uint32_t out = 0;
	for (int k = 0; k < 3; ++k)
		out |= ((uint32_t)(int)dop::sqrt(dop::div(s[k], (double)cnt)) & 0xffu) << (8 * k);
	return out;
}

// SET = true: the marked voxels are set to `miss` (already clamped) instead of updated --
// setValueVolume, occupancy_map_base.h:492-518, :1151-1157.
// Change detection (enableChangeDetection, occupancy_map_base.h:779-790): when M.chg_mask is
// allocated, the voxels whose value actually changed (updateOccupancy returned true, :1139-1145)
// are OR-ed into the block's change mask (linear bit order, like the miss / hit masks).
__device__ __forceinline__ unsigned long long octet_unbits8(uint32_t c8, uint32_t o)
{
	const uint32_t base = ((o & 1u) << 1) | ((o & 2u) << 2) | ((o & 4u) << 3);
	const unsigned long long s = (c8 & 3u) | ((c8 & 0xcu) << 2) | ((c8 & 0x30u) << 12) | ((c8 & 0xc0u) << 14);
	return s << base;
}

template <bool SET = false>
__device__ __forceinline__ void update_octet(const DeviceMap& M, float miss, float* lp, uint32_t m8,
                                             uint32_t h8, float4 a0, float4 a1, float& omax,
                                             uint32_t& oflags, size_t blk = 0, uint32_t oct = 0)
{
	float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
	const float v0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
	omax = -3.402823466e+38f;
	float omin = 3.402823466e+38f;
	bool unk = false;
	// all hits of a scan are applied before its misses (occupancy_map_base.h:1351-1365);
	// hits are rare (one voxel per ray), so their arithmetic is skipped for octets without one
	if (!SET && h8) {
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			float hv = apply_update(M, v[j], M.hit);
			v[j] = ((h8 >> j) & 1u) ? hv : v[j];
		}
	}
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		float mv = SET ? miss : apply_update(M, v[j], miss);
		v[j] = ((m8 >> j) & 1u) ? mv : v[j];
		omax = fmaxf(omax, v[j]);
		omin = fminf(omin, v[j]);
		unk = unk || (v[j] >= M.free_ceil && v[j] <= M.occ_floor);
	}
	// contains_free = any voxel below the free threshold, contains_unknown = any in between
	oflags = (omin < M.free_ceil ? 1u : 0u) | (unk ? 2u : 0u);
	if (!SET && M.chg_mask) {
		uint32_t c8 = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) c8 |= (__float_as_uint(v[j]) != __float_as_uint(v0[j]) ? 1u : 0u) << j;
		if (c8) atomicOr(&M.chg_mask[blk], octet_unbits8(c8, oct));
	}
	reinterpret_cast<float4*>(lp)[0] = make_float4(v[0], v[1], v[2], v[3]);
	reinterpret_cast<float4*>(lp)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- mbarrier + 1-D bulk copy (TMA engine, cp.async.bulk) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "UFO_MBAR_WAIT_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra UFO_MBAR_DONE_%=;\n\t"
	    "bra UFO_MBAR_WAIT_%=;\n\t"
	    "UFO_MBAR_DONE_%=:\n\t"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is
// signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(smem_dst)),
	             "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// A brick whose key lies outside the tree only collects out-of-tree marks (see k_alias_*):
// returns true for such a brick and the wrapped brick coordinates its marks land in.
__device__ __forceinline__ bool alias_source(const DeviceMap& M, uint32_t brick, uint32_t& tx, uint32_t& ty,
                                             uint32_t& tz)
{
	uint32_t x, y, z;
	unpack_key(M.brick_key[brick], x, y, z);
	const uint32_t km = M.g.key_mask >> 4;
	tx = x & km;
	ty = y & km;
	tz = z & km;
	return ((x | y | z) & ~km) != 0;
}

struct __align__(128) WarpSlab {
	unsigned long long mm[64];
	unsigned long long hm[64];
	uint32_t meta[64];
};

// K3: one WARP per brick of the scan's TOUCHED LIST (the cost of a scan does not depend on the size
// of the map), no CTA-wide barrier at all: a warp waiting for its leaf sectors never holds others at
// a __syncthreads (a CTA-cooperative version spent 31 % of its stall samples there, profiles/r02a_*).
//  * lane 0 stages the header slab of the warp's NEXT brick -- miss masks 512 B, hit masks 512 B,
//    meta 256 B -- with three bulk copies (cp.async.bulk) on a per-warp mbarrier, two stages;
//  * lane = block (twice: 64 blocks): masks from shared memory, touched-octet list of the WHOLE
//    brick by warp prefix sums -- built one brick EARLY, right after the previous brick's leaf loop,
//    so that the brick's leaf sectors can be pulled into the L2 (prefetch.global.L2, no registers)
//    while the previous brick's aggregates are still being written;
//  * the old depth-1 sectors / depth-2 aggregates the block lanes need at the end travel by
//    cp.async into shared memory behind the leaf loop (no registers held across it);
//  * lanes then own list entries = 32 B leaf sectors, UFO_UB_INFLIGHT requested per lane before the
//    first is used: v = clamp(v + hit), then v = clamp(v + miss) (float, order fixed), written back
//    in place; octet maxima / flags return through warp-private shared memory (__syncwarp only);
//  * block lanes write the depth-1 sector, the depth-2 aggregate, the meta word, clear the masks;
//    depth-3 / depth-4 aggregates by shuffles.
// Measured on the bench workload (profiles/README.md): 0.70 ms.  A list per half brick: 0.88 (latency
// exposed twice per brick); 2 sectors in flight at 6 CTAs of 80 registers: 0.82; without the L2
// prefetch: 0.75; 228 KB instead of 196 KB of shared memory per SM (28 KB of L1 left): 0.85; a
// cp.async leaf pipeline through shared memory: 0.91-0.95 (fewer warps, more instructions).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* g)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* g)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* g)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];" ::"l"(g)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef UFO_UB_MINBLOCKS
#define UFO_UB_MINBLOCKS 5
#endif
#ifndef UFO_UB_PREFETCH
#define UFO_UB_PREFETCH 1  // pull the next brick's leaf sectors into the L2 one brick early
#endif
#ifndef UFO_UB_INFLIGHT_COLOR
#define UFO_UB_INFLIGHT_COLOR 1  // colour maps (each entry is two sectors: log-odds + colour): 1 / 2 / 3 in flight = 5.4 / 6.9 / 8.6 ms on config #3
#endif
#ifndef UFO_UB_INFLIGHT
#define UFO_UB_INFLIGHT 3  // leaf sectors requested per lane before the first is used
#endif
#ifndef UFO_UB_MINBLOCKS_COLOR
#define UFO_UB_MINBLOCKS_COLOR 7
#endif
// One warp's shared memory.  Kept small on purpose: with 5 CTAs per SM it has to fit the 196 KB
// carve-out -- one step more (228 KB) leaves 28 KB of L1, too little for the leaf sectors the warps keep
// in flight, and costs 15 % of the kernel (profiles/README.md).
template <bool COLOR>
struct __align__(128) BrickWork {
	WarpSlab slab[2];
	// old depth-1 sector of a marked block; for an unmarked block the slot holds its old depth-2
	// aggregate (first 8 bytes of s1lo) and colour (first 4 bytes of s1hi) instead
	float4 s1lo[64], s1hi[64];
	uint4 c1lo[COLOR ? 64 : 1], c1hi[COLOR ? 64 : 1];
	float omax[512];
	uint32_t orgb[COLOR ? 512 : 2];
	// touched octets of a brick: (block << 3) | octet; after use the slot carries the octet's flags.
	// [it & 1]: the brick being processed, the other one: the warp's next brick (built one brick early)
	uint16_t list[2][512];
	unsigned long long bar[2];
};
template <bool COLOR>
struct UbShape {
	static constexpr int kWarps = COLOR ? 2 : 4;
	static constexpr int kMinBlocks = COLOR ? UFO_UB_MINBLOCKS_COLOR : UFO_UB_MINBLOCKS;
};

template <bool COLOR, bool SET = false>
__global__ void __launch_bounds__(UbShape<COLOR>::kWarps * 32, UbShape<COLOR>::kMinBlocks) k_update_brick(DeviceMap M, float miss)
{
	constexpr int WARPS = UbShape<COLOR>::kWarps;
	__shared__ BrickWork<COLOR> s_work[WARPS];
	constexpr uint32_t FULL = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	BrickWork<COLOR>& S = s_work[w];
	if (__ldg(&M.ctr->overflow) & ~4u) return;  // a pool overflowed while marking: the host grows it and repeats the scan
	const uint32_t n_touched = __ldg(&M.ctr->n_touched);
	const uint32_t gw = blockIdx.x * WARPS + w, n_warps = gridDim.x * WARPS;
	if (gw >= n_touched) return;
	const bool fold = M.alias_miss == nullptr;

	uint32_t st_brick0 = kNone, st_brick1 = kNone, st_mi0 = 0, st_mi1 = 0;  // lane 0: (brick, mask group) of the two stages
	auto issue = [&](uint32_t stage, uint32_t e) {
		const uint32_t brick = M.touched[e], mi = M.touched_mi[e];
		if (stage == 0) {
			st_brick0 = brick;
			st_mi0 = mi;
		} else {
			st_brick1 = brick;
			st_mi1 = mi;
		}
		fence_proxy_async();
		mbar_expect_tx(&S.bar[stage], 512u + 512u + 256u);
		bulk_load(S.slab[stage].mm, M.mask_base + (size_t)mi * 64, 512u, &S.bar[stage]);
		bulk_load(S.slab[stage].hm, M.hit_mask + (size_t)brick * 64, 512u, &S.bar[stage]);
		bulk_load(S.slab[stage].meta, M.meta + (size_t)brick * 64, 256u, &S.bar[stage]);
	};
	if (lane == 0) {
		mbar_init(&S.bar[0], 1);
		mbar_init(&S.bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		issue(0, gw);
	}
	__syncwarp();

	unsigned int s_vox = 0, s_hit = 0, s_oct = 0, s_blk = 0, s_new = 0, s_lin = 0;

	// lane = block lane of either half: touched octets of the brick in `slab`, list by warp prefix sums
	auto build_list = [&](const WarpSlab& slab, uint16_t* list, uint32_t (&t8)[2], uint32_t (&excl)[2], uint32_t& total) {
		total = 0;
#pragma unroll
		for (uint32_t h = 0; h < 2; ++h) {
			const uint32_t blk = h * 32 + lane;
			const unsigned long long u = slab.mm[blk] | slab.hm[blk];
			uint32_t t = 0;
			if (u) {
#pragma unroll
				for (uint32_t o = 0; o < 8; ++o) {
					const uint32_t base = ((o & 1u) << 1) | ((o & 2u) << 2) | ((o & 4u) << 3);
					t |= (((u >> base) & 0x330033ull) ? 1u : 0u) << o;
				}
			}
			uint32_t incl = __popc(t);
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(FULL, incl, o);
				if (lane >= (uint32_t)o) incl += v;
			}
			excl[h] = total + incl - __popc(t);
			t8[h] = t;
			{
				uint32_t bits = t, at = excl[h];
				while (bits) {
					const uint32_t o = __ffs(bits) - 1;
					bits &= bits - 1;
					list[at++] = (uint16_t)((blk << 3) | o);
				}
			}
			total += __shfl_sync(FULL, incl, 31);
		}
	};

	uint32_t t8[2], excl[2], total = 0;     // the brick being processed
	uint32_t t8n[2], excln[2], totaln = 0;  // the warp's next brick (its list is built one brick early)
	mbar_wait(&S.bar[0], 0u);
	build_list(S.slab[0], S.list[0], t8, excl, total);
	__syncwarp();
	uint32_t it = 0;
	for (uint32_t e = gw; e < n_touched; e += n_warps, ++it) {
		const uint32_t st = it & 1u;
		const bool has_next = e + n_warps < n_touched;
		if (lane == 0 && has_next) issue(st ^ 1u, e + n_warps);  // stage st^1 was released by the __syncwarp at the end of the loop
		const uint32_t brick = __shfl_sync(FULL, st == 0 ? st_brick0 : st_brick1, 0);
		const uint32_t mi = __shfl_sync(FULL, st == 0 ? st_mi0 : st_mi1, 0);
		const WarpSlab& slab = S.slab[st];
		const uint16_t* list = S.list[st];
		const size_t b0 = (size_t)brick * 64;

		// what the block lanes need after the leaf loop: requested now, into shared memory
#pragma unroll
		for (uint32_t h = 0; h < 2; ++h) {
			const uint32_t blk = h * 32 + lane;
			if (t8[h]) {
				if (t8[h] != 0xffu) {
					cp_async16(&S.s1lo[blk], M.sum1 + (b0 + blk) * 8);
					cp_async16(&S.s1hi[blk], M.sum1 + (b0 + blk) * 8 + 4);
					if (COLOR) {
						cp_async16(&S.c1lo[COLOR ? blk : 0], M.sum1_rgb + (b0 + blk) * 8);
						cp_async16(&S.c1hi[COLOR ? blk : 0], M.sum1_rgb + (b0 + blk) * 8 + 4);
					}
				}
			} else if (fold && (slab.meta[blk] & 0xff0000u)) {
				cp_async8(&S.s1lo[blk], &M.agg2[b0 + blk]);
				if (COLOR) cp_async4(&S.s1hi[blk], &M.rgb2[b0 + blk]);
			}
		}
		cp_async_commit();
		__syncwarp();
		// lanes own list entries = 32 B leaf sectors (+ the colour sector of the same octet), NF of them
		// requested per lane before the first is used
		constexpr int NF = COLOR ? UFO_UB_INFLIGHT_COLOR : UFO_UB_INFLIGHT;
		for (uint32_t i = lane; i < total; i += 32 * NF) {
			float4 a0[NF], a1[NF];
			uint4 x0[COLOR ? NF : 1], x1[COLOR ? NF : 1];
			uint32_t en[NF];
#pragma unroll
			for (int f = 0; f < NF; ++f) {
				const uint32_t idx = i + 32 * f;
				en[f] = idx < total ? (uint32_t)list[idx] : 0xffffffffu;
				if (en[f] != 0xffffffffu) {
					const size_t off = (b0 + (en[f] >> 3)) * 64 + 8 * (en[f] & 7u);
					const float4* lp = reinterpret_cast<const float4*>(M.leaf + off);
					a0[f] = lp[0];
					a1[f] = lp[1];
					if (COLOR) {
						const uint4* cp = reinterpret_cast<const uint4*>(M.leaf_rgb + off);
						x0[COLOR ? f : 0] = cp[0];
						x1[COLOR ? f : 0] = cp[1];
					}
				}
			}
#pragma unroll
			for (int f = 0; f < NF; ++f) {
				if (en[f] == 0xffffffffu) continue;
				const uint32_t idx = i + 32 * f, t = en[f] >> 3, o = en[f] & 7u;
				const uint32_t m8 = octet_bits8(slab.mm[t], o), h8 = octet_bits8(slab.hm[t], o);
				float omax;
				uint32_t ofl;
				update_octet<SET>(M, miss, M.leaf + (b0 + t) * 64 + 8 * o, m8, h8, a0[f], a1[f], omax, ofl, b0 + t, o);
				S.omax[idx] = omax;
				S.list[st][idx] = (uint16_t)ofl;  // the entry has been consumed: the slot carries the flags back
				if (COLOR) {
					const uint4 c0 = x0[COLOR ? f : 0], c1 = x1[COLOR ? f : 0];
					const uint32_t cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
					S.orgb[COLOR ? idx : 0] = rms_rgb(cc, 8);
				}
				s_vox += __popc(m8 | h8);
				s_hit += __popc(h8);
				s_oct += 1;
			}
		}
		if (has_next) {
			// the next brick of this warp: its slab was requested at the top of this iteration; build its
			// list now and pull its leaf sectors into the L2 while this brick's aggregates are written
			mbar_wait(&S.bar[st ^ 1u], ((it + 1u) >> 1) & 1u);
			build_list(S.slab[st ^ 1u], S.list[st ^ 1u], t8n, excln, totaln);
#if UFO_UB_PREFETCH
			__syncwarp();
			const uint32_t brick_n = __shfl_sync(FULL, st == 0 ? st_brick1 : st_brick0, 0);
			const float* leaf_n = M.leaf + (size_t)brick_n * 64 * 64;
			for (uint32_t i = lane; i < totaln; i += 32) {
				const uint32_t en = S.list[st ^ 1u][i];
				prefetch_l2(leaf_n + (en >> 3) * 64 + 8 * (en & 7u));
				if (COLOR) prefetch_l2(M.leaf_rgb + (size_t)brick_n * 64 * 64 + (en >> 3) * 64 + 8 * (en & 7u));
			}
#pragma unroll
			for (uint32_t h = 0; h < 2; ++h)
				if (t8n[h] && t8n[h] != 0xffu) prefetch_l2(M.sum1 + ((size_t)brick_n * 64 + h * 32 + lane) * 8);
#endif
		}
		cp_async_wait_all();
		__syncwarp();
		// block lanes: depth-1 sector, depth-2 aggregate, meta, mask clearing
		float agg_occ[2];
		uint32_t agg_fl[2], agg_rgb[2];
#pragma unroll
		for (uint32_t h = 0; h < 2; ++h) {
			const uint32_t blk = h * 32 + lane;
			const size_t b = b0 + blk;
			const unsigned long long hm = slab.hm[blk];
			const uint32_t mt = slab.meta[blk];
			const bool marked = (slab.mm[blk] | hm) != 0ull;
			float my_occ = 0.0f;
			uint32_t my_fl = M.default_flags, my_rgb = 0;
			if (marked) {
				float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
				uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
				if (t8[h] != 0xffu) {
					p0 = S.s1lo[blk];
					p1 = S.s1hi[blk];
					if (COLOR) {
						q0 = S.c1lo[COLOR ? blk : 0];
						q1 = S.c1hi[COLOR ? blk : 0];
					}
				}
				const float old1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
				const uint32_t oldc[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
				float new1[8];
				uint32_t newc[8];
				float bmax = -3.402823466e+38f;
				uint32_t bfl = 0, newmeta = 0, at = excl[h];
#pragma unroll
				for (uint32_t o = 0; o < 8; ++o) {
					float om = 0.0f;
					uint32_t fl = M.default_flags, touched = 0, oc = 0;
					if ((t8[h] >> o) & 1u) {
						om = S.omax[at];
						fl = S.list[st][at];
						if (COLOR) oc = S.orgb[COLOR ? at : 0];
						++at;
						touched = 1;
					} else if ((mt >> (16 + o)) & 1u) {
						om = old1[o];
						fl = (mt >> (2 * o)) & 3u;
						if (COLOR) oc = oldc[o];
					}
					new1[o] = om;
					newc[o] = oc;
					bmax = fmaxf(bmax, om);
					bfl |= fl;
					newmeta |= (fl << (2 * o)) | (touched << (16 + o));
				}
				float4* sp = reinterpret_cast<float4*>(M.sum1 + b * 8);
				sp[0] = make_float4(new1[0], new1[1], new1[2], new1[3]);
				sp[1] = make_float4(new1[4], new1[5], new1[6], new1[7]);
				M.agg2[b] = {bmax, bfl};
				if (COLOR) {
					uint4* cp = reinterpret_cast<uint4*>(M.sum1_rgb + b * 8);
					cp[0] = make_uint4(newc[0], newc[1], newc[2], newc[3]);
					cp[1] = make_uint4(newc[4], newc[5], newc[6], newc[7]);
					my_rgb = rms_rgb(newc, 8);
					M.rgb2[b] = my_rgb;
				}
				M.meta[b] = (newmeta & 0xffffffu) | (mt & 0xff0000u) | (M.scan_id << 24);
				M.mask_base[(size_t)mi * 64 + blk] = 0ull;  // masks cleared for the next scan
				if (hm) M.hit_mask[b] = 0ull;
				s_blk += 1;
				s_lin += ((t8[h] & 0x0fu) ? 1u : 0u) + ((t8[h] & 0xf0u) ? 1u : 0u);
				s_new += (mt & 0xff0000u) ? 0u : 1u;
				my_occ = bmax;
				my_fl = bfl | 0x100u;
			} else if (fold && (mt & 0xff0000u)) {
				const float4 o2 = S.s1lo[blk];
				my_occ = o2.x;
				my_fl = __float_as_uint(o2.y);
				if (COLOR) my_rgb = __float_as_uint(S.s1hi[blk].x);
			}
			agg_occ[h] = my_occ;
			agg_fl[h] = my_fl;
			agg_rgb[h] = my_rgb;
		}
		// depth-3 / depth-4 aggregates: half h holds depth-3 nodes 4h..4h+3, eight lanes each
		if (fold) {
			float m3[2], m4 = -3.402823466e+38f;
			uint32_t f3[2], u3[2], f4 = 0, rgb3[2] = {0, 0};
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				m3[h] = agg_occ[h];
				f3[h] = agg_fl[h] & 3u;
				u3[h] = (agg_fl[h] >> 8) & 1u;
#pragma unroll
				for (int o = 1; o < 8; o <<= 1) {
					m3[h] = fmaxf(m3[h], __shfl_xor_sync(FULL, m3[h], o));
					f3[h] |= __shfl_xor_sync(FULL, f3[h], o);
					u3[h] |= __shfl_xor_sync(FULL, u3[h], o);
				}
				if (COLOR) {
					uint32_t cc[8];
#pragma unroll
					for (int q = 0; q < 8; ++q) cc[q] = __shfl_sync(FULL, agg_rgb[h], (lane & 24) + q);
					rgb3[h] = rms_rgb(cc, 8);
				}
				float t = m3[h];
				uint32_t tf = f3[h];
#pragma unroll
				for (int o = 8; o < 32; o <<= 1) {
					t = fmaxf(t, __shfl_xor_sync(FULL, t, o));
					tf |= __shfl_xor_sync(FULL, tf, o);
				}
				m4 = fmaxf(m4, t);
				f4 |= tf;
			}
			uint32_t rgb4 = 0;
			if (COLOR) {
				uint32_t cc[8];
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					cc[q] = __shfl_sync(FULL, rgb3[0], 8 * q);
					cc[4 + q] = __shfl_sync(FULL, rgb3[1], 8 * q);
				}
				rgb4 = rms_rgb(cc, 8);
			}
			const uint32_t ub0 = __ballot_sync(FULL, u3[0] != 0 && (lane & 7) == 0), ub1 = __ballot_sync(FULL, u3[1] != 0 && (lane & 7) == 0);
			if ((lane & 7) == 0) {
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					M.brick_sum3[(size_t)brick * 8 + 4 * h + (lane >> 3)] = {m3[h], f3[h]};
					if (COLOR) M.brick_rgb3[(size_t)brick * 8 + 4 * h + (lane >> 3)] = rgb3[h];
				}
			}
			if (lane == 0) {
				M.brick_sum4[brick] = {m4, f4};
				if (COLOR) M.brick_rgb4[brick] = rgb4;
				unsigned long long* slot = M.ctr->stat[brick % kStatSlots];
				atomicAdd(&slot[5], 1ull);
				atomicAdd(&slot[6], (unsigned long long)(__popc(ub0) + __popc(ub1)));
			}
		}
		t8[0] = t8n[0];
		t8[1] = t8n[1];
		excl[0] = excln[0];
		excl[1] = excln[1];
		total = totaln;
		__syncwarp();  // the slab stage, the list and the result arrays are reused
	}
	for (int o = 16; o > 0; o >>= 1) {
		s_vox += __shfl_xor_sync(FULL, s_vox, o);
		s_hit += __shfl_xor_sync(FULL, s_hit, o);
		s_oct += __shfl_xor_sync(FULL, s_oct, o);
		s_blk += __shfl_xor_sync(FULL, s_blk, o);
		s_new += __shfl_xor_sync(FULL, s_new, o);
		s_lin += __shfl_xor_sync(FULL, s_lin, o);
	}
	if (lane == 0) {
		unsigned long long* slot = M.ctr->stat[gw % kStatSlots];
		if (s_vox) atomicAdd(&slot[0], (unsigned long long)s_vox);
		if (s_hit) atomicAdd(&slot[1], (unsigned long long)s_hit);
		if (s_oct) atomicAdd(&slot[2], (unsigned long long)s_oct);
		if (s_blk) atomicAdd(&slot[3], (unsigned long long)s_blk);
		if (s_new) atomicAdd(&slot[4], (unsigned long long)s_new);
		if (s_lin) atomicAdd(&slot[7], (unsigned long long)s_lin);
	}
}

// Stand-alone depth-3 / depth-4 pass over the touched list: only used when out-of-tree marks
// exist (they are applied after k_update_brick, which then skips its folded reduction).
// One warp per brick, lane owns children 2*lane and 2*lane+1 (both under depth-3 node lane/4).
template <bool COLOR>
__global__ void __launch_bounds__(256) k_brick_agg(DeviceMap M)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t n_touched = __ldg(&M.ctr->n_touched);
	if (__ldg(&M.ctr->overflow) & ~4u) return;
	constexpr uint32_t FULL = 0xffffffffu;
	for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n_touched; e += (gridDim.x * blockDim.x) >> 5) {
		const uint32_t brick = M.touched[e];
		if (M.alias_miss) {  // bricks that only collect out-of-tree marks own no voxels
			uint32_t ax, ay, az;
			if (alias_source(M, brick, ax, ay, az)) continue;
		}
		const size_t b = (size_t)brick * 64 + 2 * lane;
		const uint2 mt = *reinterpret_cast<const uint2*>(&M.meta[b]);
		const uint4 ag = *reinterpret_cast<const uint4*>(&M.agg2[b]);
		Agg c0 = {__uint_as_float(ag.x), ag.y}, c1 = {__uint_as_float(ag.z), ag.w};
		if (!(mt.x & 0xff0000u)) c0 = {0.0f, M.default_flags};  // never written: unknown space
		if (!(mt.y & 0xff0000u)) c1 = {0.0f, M.default_flags};
		const uint32_t tag = M.scan_id & 0xffu;
		const bool upd = ((mt.x & 0xff0000u) && (mt.x >> 24) == tag) || ((mt.y & 0xff0000u) && (mt.y >> 24) == tag);
		float m3 = fmaxf(c0.occ, c1.occ);
		uint32_t f3 = (c0.flags | c1.flags) & 3u;
#pragma unroll
		for (int o = 1; o < 4; o <<= 1) {
			m3 = fmaxf(m3, __shfl_xor_sync(FULL, m3, o));
			f3 |= __shfl_xor_sync(FULL, f3, o);
		}
		float m4 = m3;
		uint32_t f4 = f3;
#pragma unroll
		for (int o = 4; o < 32; o <<= 1) {
			m4 = fmaxf(m4, __shfl_xor_sync(FULL, m4, o));
			f4 |= __shfl_xor_sync(FULL, f4, o);
		}
		if ((lane & 3) == 0) M.brick_sum3[(size_t)brick * 8 + (lane >> 2)] = {m3, f3};
		if (lane == 0) M.brick_sum4[brick] = {m4, f4};
		if (COLOR) {
			const uint2 cr = *reinterpret_cast<const uint2*>(&M.rgb2[b]);
			const uint32_t r0 = (mt.x & 0xff0000u) ? cr.x : 0u, r1 = (mt.y & 0xff0000u) ? cr.y : 0u;
			uint32_t cc[8];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				cc[2 * j] = __shfl_sync(FULL, r0, (lane & 28) + j);
				cc[2 * j + 1] = __shfl_sync(FULL, r1, (lane & 28) + j);
			}
			const uint32_t rgb3 = rms_rgb(cc, 8);
#pragma unroll
			for (int j = 0; j < 8; ++j) cc[j] = __shfl_sync(FULL, rgb3, 4 * j);
			const uint32_t rgb4 = rms_rgb(cc, 8);
			if ((lane & 3) == 0) M.brick_rgb3[(size_t)brick * 8 + (lane >> 2)] = rgb3;
			if (lane == 0) M.brick_rgb4[brick] = rgb4;
		}
		const uint32_t ub = __ballot_sync(FULL, upd);
		if (lane == 0) {
			uint32_t d3 = 0;
			for (int k = 0; k < 8; ++k) d3 += ((ub >> (4 * k)) & 0xfu) ? 1u : 0u;
			unsigned long long* slot = M.ctr->stat[brick % kStatSlots];
			atomicAdd(&slot[5], 1ull);
			atomicAdd(&slot[6], (unsigned long long)d3);
		}
	}
}

}  // namespace ufo_b200
