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

// K3.  Persistent CTAs over chunks of four bricks (256 blocks) taken from the scan's TOUCHED
// LIST, so the cost of a scan does not depend on the size of the map.
//  * One elected thread stages the per-brick header slabs of the CTA's NEXT chunk -- miss masks
//    (512 B), hit masks (512 B), block meta words (256 B) -- into shared memory with 1-D bulk
//    copies (cp.async.bulk + mbarrier transaction count, two stages), a whole chunk ahead of
//    their use and at no register cost.
//  * Block threads (one per block) take their masks from shared memory, request the block data
//    that is needed last first (old depth-1 maxima of octets that stay untouched, old depth-2
//    aggregate of unmarked blocks), and build by prefix sum a DENSE LIST of the touched octets.
//  * Every thread then owns list entries = 32 B leaf sectors: v = clamp(v + hit), then
//    v = clamp(v + miss) (float, order fixed), written back in place; octet maxima / flags go
//    through shared memory to the block threads, which write the depth-1 sector, the depth-2
//    aggregate, the meta word and clear the masks.
//  * The depth-3 / depth-4 aggregates of a chunk's bricks are reduced from shared memory by 32
//    threads during the next iteration (no extra barrier, no second kernel).
#ifndef UFO_UC_MINBLOCKS
#define UFO_UC_MINBLOCKS 4
#endif
#ifndef UFO_UC_MINBLOCKS_COLOR
#define UFO_UC_MINBLOCKS_COLOR 3
#endif
constexpr int kUcBricks = 4;
constexpr int kUcBlocks = kUcBricks * 64;  // blocks per chunk
constexpr int kUcThreads = kUcBlocks;
#ifndef UFO_UC_GRID_PER_SM
#define UFO_UC_GRID_PER_SM UFO_UC_MINBLOCKS
#endif

// depth-3 / depth-4 aggregates of the four bricks of a chunk from the chunk's 256 depth-2
// aggregates in shared memory: called by threads 0..31 (warp 0), thread t owns depth-3 node t & 7
// of brick t >> 3.
template <bool COLOR>
__device__ __forceinline__ void chunk_brick_reduce(const DeviceMap& M, const float* aocc, const uint32_t* afl,
                                                   const uint32_t* argb, uint32_t rb, uint32_t tid)
{
	constexpr uint32_t FULL = 0xffffffffu;
	const uint32_t base = tid * 8;
	float m3 = -3.402823466e+38f;
	uint32_t f3 = 0, upd = 0;
	uint32_t cc[8];
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		m3 = fmaxf(m3, aocc[base + j]);
		const uint32_t f = afl[base + j];
		f3 |= f & 3u;
		upd |= (f >> 8) & 1u;
		if (COLOR) cc[j] = argb[base + j];
	}
	float m4 = m3;
	uint32_t f4 = f3;
#pragma unroll
	for (int o = 1; o < 8; o <<= 1) {
		m4 = fmaxf(m4, __shfl_xor_sync(FULL, m4, o));
		f4 |= __shfl_xor_sync(FULL, f4, o);
	}
	uint32_t rgb3 = 0, rgb4 = 0;
	if (COLOR) {
		rgb3 = rms_rgb(cc, 8);
#pragma unroll
		for (int j = 0; j < 8; ++j) cc[j] = __shfl_sync(FULL, rgb3, (tid & 24) + j);
		rgb4 = rms_rgb(cc, 8);
	}
	const uint32_t ub = __ballot_sync(FULL, upd != 0);
	if (rb == kNone) return;
	M.brick_sum3[(size_t)rb * 8 + (tid & 7)] = {m3, f3};
	if (COLOR) M.brick_rgb3[(size_t)rb * 8 + (tid & 7)] = rgb3;
	if ((tid & 7) == 0) {
		M.brick_sum4[rb] = {m4, f4};
		if (COLOR) M.brick_rgb4[rb] = rgb4;
		// D_3 / D_4 counters: a depth-3 node is touched iff one of its 8 blocks was updated this scan
		const uint32_t d3 = __popc((ub >> (tid & 24)) & 0xffu);
		unsigned long long* slot = M.ctr->stat[rb % kStatSlots];
		atomicAdd(&slot[5], 1ull);
		atomicAdd(&slot[6], (unsigned long long)d3);
	}
}

template <bool COLOR, bool SET = false>
__global__ void __launch_bounds__(kUcThreads, COLOR ? UFO_UC_MINBLOCKS_COLOR : UFO_UC_MINBLOCKS) k_update_compact(DeviceMap M, float miss)
{
	__shared__ __align__(128) unsigned long long s_mm[2][kUcBlocks];
	__shared__ __align__(128) unsigned long long s_hm[2][kUcBlocks];
	__shared__ __align__(128) uint32_t s_meta[2][kUcBlocks];
	__shared__ __align__(8) unsigned long long s_bar[2];
	__shared__ uint32_t s_brick[2][kUcBricks];
	__shared__ uint32_t s_mi[2][kUcBricks];  // where each brick's free-space masks are (DeviceMap::touched_mi)
	__shared__ uint16_t s_list[kUcBlocks * 8];
	__shared__ float s_omax[8 * kUcBlocks];  // [octet][block]: conflict-free for the block threads
	__shared__ unsigned char s_ofl[8 * kUcBlocks];
	__shared__ uint32_t s_orgb[COLOR ? 8 * kUcBlocks : 1];  // depth-1 colours of the touched octets
	__shared__ uint32_t s_wtot[2][kUcBlocks / 32];
	__shared__ float s_aocc[2][kUcBlocks];    // depth-2 aggregates of the chunk's blocks
	__shared__ uint32_t s_afl[2][kUcBlocks];  // bits 0..1 flags, bit 8 updated this scan
	__shared__ uint32_t s_argb[COLOR ? 2 : 1][COLOR ? kUcBlocks : 1];
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	// Launched without the host having seen this scan's counters.  If a pool overflowed while
	// marking, the host regrows and repeats the scan, so nothing may be consumed now (bit 2, the
	// upper-node pool, belongs to the propagation pass that runs after this kernel).
	if (__ldg(&M.ctr->overflow) & ~4u) return;
	const uint32_t n_touched = __ldg(&M.ctr->n_touched);
	const uint32_t n_chunks = (n_touched + kUcBricks - 1) / kUcBricks;
	if (blockIdx.x >= n_chunks) return;
	// the brick-level reduction is folded in unless out-of-tree marks exist: those are applied
	// after this kernel and k_brick_agg follows them
	const bool fold = M.alias_miss == nullptr;

	uint32_t ids_next[kUcBricks] = {kNone, kNone, kNone, kNone}, mis_next[kUcBricks] = {0, 0, 0, 0};
	auto load_ids = [&](uint32_t c, uint32_t* ids, uint32_t* mis) {
#pragma unroll
		for (uint32_t q = 0; q < (uint32_t)kUcBricks; ++q) {
			const uint32_t e = c * kUcBricks + q;
			const bool ok = c < n_chunks && e < n_touched;
			ids[q] = ok ? M.touched[e] : kNone;
			mis[q] = ok ? M.touched_mi[e] : 0u;
		}
	};
	auto issue = [&](uint32_t stage, const uint32_t* ids, const uint32_t* mis) {
		uint32_t bytes = 0;
#pragma unroll
		for (uint32_t q = 0; q < (uint32_t)kUcBricks; ++q) {
			s_brick[stage][q] = ids[q];
			s_mi[stage][q] = mis[q];
			if (ids[q] != kNone) bytes += 512u + 512u + 256u;
		}
		if (!bytes) return;
		fence_proxy_async();  // earlier generic-proxy reads of this stage are ordered before the async writes
		mbar_expect_tx(&s_bar[stage], bytes);
#pragma unroll
		for (uint32_t q = 0; q < (uint32_t)kUcBricks; ++q) {
			if (ids[q] == kNone) continue;
			const size_t b0 = (size_t)ids[q] * 64;
			bulk_load(&s_mm[stage][q * 64], M.mask_base + (size_t)mis[q] * 64, 512u, &s_bar[stage]);
			bulk_load(&s_hm[stage][q * 64], M.hit_mask + b0, 512u, &s_bar[stage]);
			bulk_load(&s_meta[stage][q * 64], M.meta + b0, 256u, &s_bar[stage]);
		}
	};
	if (tid == 0) {
		mbar_init(&s_bar[0], 1);
		mbar_init(&s_bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		uint32_t ids[kUcBricks], mis[kUcBricks];
		load_ids(blockIdx.x, ids, mis);
		issue(0, ids, mis);
		load_ids(blockIdx.x + gridDim.x, ids_next, mis_next);
	}
	__syncthreads();

	unsigned int s_vox = 0, s_hit = 0, s_oct = 0, s_blk = 0, s_new = 0;
	uint32_t rb_cur = kNone, rb_prev = kNone;  // threads 0..31: brick (tid >> 3) of this / the previous chunk
	bool have_prev = false;                    // uniform: the previous chunk left aggregates to reduce
	uint32_t it = 0;
	for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
		const uint32_t st = it & 1u;
		if (tid == 0 && c + gridDim.x < n_chunks) {
			// stage st^1 was last read before the previous iteration's final barrier
			issue(st ^ 1u, ids_next, mis_next);
			load_ids(c + 2 * gridDim.x, ids_next, mis_next);
		}
		mbar_wait(&s_bar[st], (it >> 1) & 1u);
		// ---- block threads: masks, touched-octet bitmap, list offsets ----
		const uint32_t brick = s_brick[st][tid >> 6];
		const bool vb = brick != kNone;
		const size_t b = (size_t)brick * 64 + (tid & 63);
		unsigned long long mm = 0ull, hm = 0ull;
		uint32_t mt = 0;
		if (vb) {
			mm = s_mm[st][tid];
			hm = s_hm[st][tid];
			mt = s_meta[st][tid];
		}
		rb_prev = rb_cur;
		rb_cur = tid < 32 ? s_brick[st][tid >> 3] : kNone;
		uint32_t t8 = 0, excl = 0;
		float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
		uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
		Agg old2 = {0.0f, M.default_flags};
		uint32_t old2rgb = 0;
		const bool marked = (mm | hm) != 0ull;
		if (marked) {
			const unsigned long long u = mm | hm;
#pragma unroll
			for (uint32_t o = 0; o < 8; ++o) {
				const uint32_t base = ((o & 1u) << 1) | ((o & 2u) << 2) | ((o & 4u) << 3);
				t8 |= (((u >> base) & 0x330033ull) ? 1u : 0u) << o;
			}
			// used last, requested first: the depth-1 maxima of the octets that stay untouched
			// (garbage for a never-written block; masked by meta below)
			if (t8 != 0xffu) {
				const float4* sp = reinterpret_cast<const float4*>(M.sum1 + b * 8);
				p0 = sp[0];
				p1 = sp[1];
				if (COLOR) {
					const uint4* cp = reinterpret_cast<const uint4*>(M.sum1_rgb + b * 8);
					q0 = cp[0];
					q1 = cp[1];
				}
			}
		} else if (fold && vb && (mt & 0xff0000u)) {
			// unmarked block of a touched brick: its aggregate enters the brick reduction
			old2 = M.agg2[b];
			if (COLOR) old2rgb = M.rgb2[b];
		}
		{
			uint32_t incl = __popc(t8);
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= (uint32_t)o) incl += v;
			}
			excl = incl - __popc(t8);
			if (lane == 31) s_wtot[st][tid >> 5] = incl;
		}
		__syncthreads();
		// depth-3/4 aggregates of the previous chunk (its depth-2 values were written before the barrier)
		if (have_prev && tid < 32)
			chunk_brick_reduce<COLOR>(M, s_aocc[st ^ 1u], s_afl[st ^ 1u], s_argb[COLOR ? (st ^ 1u) : 0], rb_prev, tid);
		have_prev = false;
		uint32_t total = 0;
#pragma unroll
		for (int w = 0; w < kUcBlocks / 32; ++w) {
			const uint32_t wt = s_wtot[st][w];
			if ((uint32_t)w < (tid >> 5)) excl += wt;
			total += wt;
		}
		if (total == 0) continue;  // nothing marked in this chunk (uniform; s_wtot is double-buffered)
		{
			uint32_t bits = t8, at = excl;
			while (bits) {
				const uint32_t o = __ffs(bits) - 1;
				bits &= bits - 1;
				s_list[at++] = (uint16_t)((tid << 3) | o);
			}
		}
		__syncthreads();

		// ---- all threads: one touched octet (= one 32 B sector) each ----
		for (uint32_t i = tid; i < total; i += kUcThreads) {
			const uint32_t e = s_list[i], t = e >> 3, oct = e & 7u;
			const uint32_t m8 = octet_bits8(s_mm[st][t], oct), h8 = octet_bits8(s_hm[st][t], oct);
			const size_t blk = (size_t)s_brick[st][t >> 6] * 64 + (t & 63u);
			float* lp = M.leaf + blk * 64 + 8 * oct;
			const float4 a0 = reinterpret_cast<const float4*>(lp)[0], a1 = reinterpret_cast<const float4*>(lp)[1];
			float omax;
			uint32_t ofl;
			update_octet<SET>(M, miss, lp, m8, h8, a0, a1, omax, ofl, blk, oct);
			s_omax[oct * kUcBlocks + t] = omax;
			s_ofl[oct * kUcBlocks + t] = (unsigned char)ofl;
			if (COLOR) {
				// depth-1 colour of the octet (getAverageChildColor, occupancy_map_color.cpp:177-194)
				const uint4* cp = reinterpret_cast<const uint4*>(M.leaf_rgb + blk * 64 + 8 * oct);
				const uint4 c0 = cp[0], c1 = cp[1];
				const uint32_t cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
				s_orgb[oct * kUcBlocks + t] = rms_rgb(cc, 8);
			}
			s_vox += __popc(m8 | h8);
			s_hit += __popc(h8);
			s_oct += 1;
		}
		__syncthreads();

		// ---- block threads: depth-1 sector, depth-2 aggregate, meta, mask clearing ----
		float my_occ = old2.occ;
		uint32_t my_fl = old2.flags, my_rgb = old2rgb;
		if (marked) {
			const float old1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
			const uint32_t oldc[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
			float new1[8];
			uint32_t newc[8];
			float bmax = -3.402823466e+38f;
			uint32_t bfl = 0, newmeta = 0;
#pragma unroll
			for (uint32_t o = 0; o < 8; ++o) {
				float om = 0.0f;
				uint32_t fl = M.default_flags, touched = 0, oc = 0;
				if ((t8 >> o) & 1u) {
					om = s_omax[o * kUcBlocks + tid];
					fl = s_ofl[o * kUcBlocks + tid];
					if (COLOR) oc = s_orgb[o * kUcBlocks + tid];
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
			// bits 24..31: low byte of the scan that last updated the block
			M.meta[b] = (newmeta & 0xffffffu) | (mt & 0xff0000u) | (M.scan_id << 24);
			M.mask_base[(size_t)s_mi[st][tid >> 6] * 64 + (tid & 63)] = 0ull;  // masks cleared for the next scan
			if (hm) M.hit_mask[b] = 0ull;
			s_blk += 1;
			s_new += (mt & 0xff0000u) ? 0u : 1u;
			my_occ = bmax;
			my_fl = bfl | 0x100u;
		}
		if (fold) {
			s_aocc[st][tid] = my_occ;
			s_afl[st][tid] = my_fl;
			if (COLOR) s_argb[st][tid] = my_rgb;
			have_prev = true;
		}
	}
	if (have_prev) {  // uniform
		__syncthreads();
		if (tid < 32)
			chunk_brick_reduce<COLOR>(M, s_aocc[(it - 1u) & 1u], s_afl[(it - 1u) & 1u], s_argb[COLOR ? ((it - 1u) & 1u) : 0], rb_cur,
			                          tid);
	}
	// counters: per-thread sums over the CTA's chunks, one warp reduction and one set of atomics
	for (int o = 16; o > 0; o >>= 1) {
		s_vox += __shfl_xor_sync(0xffffffffu, s_vox, o);
		s_hit += __shfl_xor_sync(0xffffffffu, s_hit, o);
		s_oct += __shfl_xor_sync(0xffffffffu, s_oct, o);
		s_blk += __shfl_xor_sync(0xffffffffu, s_blk, o);
		s_new += __shfl_xor_sync(0xffffffffu, s_new, o);
	}
	if (lane == 0) {
		unsigned long long* slot = M.ctr->stat[(blockIdx.x * (kUcThreads / 32) + (tid >> 5)) % kStatSlots];
		if (s_vox) atomicAdd(&slot[0], (unsigned long long)s_vox);
		if (s_hit) atomicAdd(&slot[1], (unsigned long long)s_hit);
		if (s_oct) atomicAdd(&slot[2], (unsigned long long)s_oct);
		if (s_blk) atomicAdd(&slot[3], (unsigned long long)s_blk);
		if (s_new) atomicAdd(&slot[4], (unsigned long long)s_new);
	}
}

// K3, warp-autonomous variant (the default): one WARP per touched brick, no CTA-wide barrier at
// all, so a warp waiting for its leaf sectors never holds seven others at a __syncthreads (the
// CTA-cooperative kernel above spends 31 % of its stall samples there, profiles/r02a_*).  Same
// arithmetic, same arrays, same staging idea at warp scope:
//  * lane 0 stages the brick's header slab (miss masks 512 B, hit masks 512 B, meta 256 B) of the
//    warp's NEXT brick with three bulk copies on a per-warp mbarrier (two stages);
//  * the brick is processed as two halves of 32 blocks, lane = block: masks from shared memory,
//    late-use data requested first, touched-octet list by warp prefix sum;
//  * lanes then own list entries (32 B leaf sectors), two in flight per lane; results go back to
//    the block lanes through warp-private shared memory (__syncwarp only);
//  * depth-3 / depth-4 aggregates by shuffles inside the warp.
#ifndef UFO_UW_WARPS
#define UFO_UW_WARPS 4
#endif
#ifndef UFO_UW_MINBLOCKS
#define UFO_UW_MINBLOCKS 6
#endif
#ifndef UFO_UW_MINBLOCKS_COLOR
#define UFO_UW_MINBLOCKS_COLOR 5
#endif
struct __align__(128) WarpSlab {
	unsigned long long mm[64];
	unsigned long long hm[64];
	uint32_t meta[64];
};

template <bool COLOR, bool SET = false>
__global__ void __launch_bounds__(UFO_UW_WARPS * 32, COLOR ? UFO_UW_MINBLOCKS_COLOR : UFO_UW_MINBLOCKS) k_update_warp(DeviceMap M, float miss)
{
	__shared__ WarpSlab s_slab[UFO_UW_WARPS][2];
	__shared__ __align__(8) unsigned long long s_bar[UFO_UW_WARPS][2];
	__shared__ uint16_t s_list[UFO_UW_WARPS][256];
	__shared__ float s_omax[UFO_UW_WARPS][256];
	__shared__ unsigned char s_ofl[UFO_UW_WARPS][256];
	__shared__ uint32_t s_orgb[COLOR ? UFO_UW_WARPS : 1][COLOR ? 256 : 1];
	constexpr uint32_t FULL = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (__ldg(&M.ctr->overflow) & ~4u) return;  // see k_update_compact
	const uint32_t n_touched = __ldg(&M.ctr->n_touched);
	const uint32_t gw = blockIdx.x * UFO_UW_WARPS + w, n_warps = gridDim.x * UFO_UW_WARPS;
	if (gw >= n_touched) return;
	const bool fold = M.alias_miss == nullptr;
	const uint32_t lt_mask = (1u << lane) - 1u;
	(void)lt_mask;

	// lane 0: (brick, mask group) of the two stages
	uint32_t st_brick0 = kNone, st_brick1 = kNone, st_mi0 = 0, st_mi1 = 0;
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
		mbar_expect_tx(&s_bar[w][stage], 512u + 512u + 256u);
		bulk_load(s_slab[w][stage].mm, M.mask_base + (size_t)mi * 64, 512u, &s_bar[w][stage]);
		bulk_load(s_slab[w][stage].hm, M.hit_mask + (size_t)brick * 64, 512u, &s_bar[w][stage]);
		bulk_load(s_slab[w][stage].meta, M.meta + (size_t)brick * 64, 256u, &s_bar[w][stage]);
	};
	if (lane == 0) {
		mbar_init(&s_bar[w][0], 1);
		mbar_init(&s_bar[w][1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		issue(0, gw);
	}
	__syncwarp();

	unsigned int s_vox = 0, s_hit = 0, s_oct = 0, s_blk = 0, s_new = 0;
	uint32_t it = 0;
	for (uint32_t e = gw; e < n_touched; e += n_warps, ++it) {
		const uint32_t st = it & 1u;
		if (lane == 0 && e + n_warps < n_touched) issue(st ^ 1u, e + n_warps);  // stage st^1 was released by the __syncwarp below
		mbar_wait(&s_bar[w][st], (it >> 1) & 1u);
		const uint32_t brick = __shfl_sync(FULL, st == 0 ? st_brick0 : st_brick1, 0);
		const uint32_t mi = __shfl_sync(FULL, st == 0 ? st_mi0 : st_mi1, 0);
		const WarpSlab& slab = s_slab[w][st];
		float agg_occ[2];
		uint32_t agg_fl[2], agg_rgb[2];
#pragma unroll
		for (uint32_t half = 0; half < 2; ++half) {
			const uint32_t blk = half * 32 + lane;
			const size_t b = (size_t)brick * 64 + blk;
			const unsigned long long mm = slab.mm[blk], hm = slab.hm[blk];
			const uint32_t mt = slab.meta[blk];
			const bool marked = (mm | hm) != 0ull;
			uint32_t t8 = 0;
			float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
			uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
			Agg old2 = {0.0f, M.default_flags};
			uint32_t old2rgb = 0;
			if (marked) {
				const unsigned long long u = mm | hm;
#pragma unroll
				for (uint32_t o = 0; o < 8; ++o) {
					const uint32_t base = ((o & 1u) << 1) | ((o & 2u) << 2) | ((o & 4u) << 3);
					t8 |= (((u >> base) & 0x330033ull) ? 1u : 0u) << o;
				}
				// used last, requested first
				if (t8 != 0xffu) {
					const float4* sp = reinterpret_cast<const float4*>(M.sum1 + b * 8);
					p0 = sp[0];
					p1 = sp[1];
					if (COLOR) {
						const uint4* cp = reinterpret_cast<const uint4*>(M.sum1_rgb + b * 8);
						q0 = cp[0];
						q1 = cp[1];
					}
				}
			} else if (fold && (mt & 0xff0000u)) {
				old2 = M.agg2[b];
				if (COLOR) old2rgb = M.rgb2[b];
			}
			// touched-octet list of this half: warp prefix sum over the per-block counts
			uint32_t incl = __popc(t8);
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(FULL, incl, o);
				if (lane >= (uint32_t)o) incl += v;
			}
			const uint32_t excl = incl - __popc(t8);
			const uint32_t total = __shfl_sync(FULL, incl, 31);
			if (total) {
				{
					uint32_t bits = t8, at = excl;
					while (bits) {
						const uint32_t o = __ffs(bits) - 1;
						bits &= bits - 1;
						s_list[w][at++] = (uint16_t)((lane << 3) | o);
					}
				}
				__syncwarp();
				// lanes own list entries = 32 B leaf sectors, two in flight per lane
				for (uint32_t i = lane; i < total; i += 64) {
					const uint32_t i2 = i + 32;
					const bool two = i2 < total;
					const uint32_t e0 = s_list[w][i], e1 = two ? s_list[w][i2] : e0;
					const uint32_t t0 = e0 >> 3, o0 = e0 & 7u, t1 = e1 >> 3, o1 = e1 & 7u;
					float* lp0 = M.leaf + ((size_t)brick * 64 + half * 32 + t0) * 64 + 8 * o0;
					float* lp1 = M.leaf + ((size_t)brick * 64 + half * 32 + t1) * 64 + 8 * o1;
					const float4 a0 = reinterpret_cast<const float4*>(lp0)[0], a1 = reinterpret_cast<const float4*>(lp0)[1];
					float4 c0 = a0, c1 = a1;
					if (two) {
						c0 = reinterpret_cast<const float4*>(lp1)[0];
						c1 = reinterpret_cast<const float4*>(lp1)[1];
					}
					{
						const uint32_t m8 = octet_bits8(slab.mm[half * 32 + t0], o0), h8 = octet_bits8(slab.hm[half * 32 + t0], o0);
						float omax;
						uint32_t ofl;
						update_octet<SET>(M, miss, lp0, m8, h8, a0, a1, omax, ofl, (size_t)brick * 64 + half * 32 + t0, o0);
						s_omax[w][i] = omax;
						s_ofl[w][i] = (unsigned char)ofl;
						if (COLOR) {
							const uint4* cp = reinterpret_cast<const uint4*>(M.leaf_rgb + ((size_t)brick * 64 + half * 32 + t0) * 64 + 8 * o0);
							const uint4 x0 = cp[0], x1 = cp[1];
							const uint32_t cc[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
							s_orgb[COLOR ? w : 0][COLOR ? i : 0] = rms_rgb(cc, 8);
						}
						s_vox += __popc(m8 | h8);
						s_hit += __popc(h8);
						s_oct += 1;
					}
					if (two) {
						const uint32_t m8 = octet_bits8(slab.mm[half * 32 + t1], o1), h8 = octet_bits8(slab.hm[half * 32 + t1], o1);
						float omax;
						uint32_t ofl;
						update_octet<SET>(M, miss, lp1, m8, h8, c0, c1, omax, ofl, (size_t)brick * 64 + half * 32 + t1, o1);
						s_omax[w][i2] = omax;
						s_ofl[w][i2] = (unsigned char)ofl;
						if (COLOR) {
							const uint4* cp = reinterpret_cast<const uint4*>(M.leaf_rgb + ((size_t)brick * 64 + half * 32 + t1) * 64 + 8 * o1);
							const uint4 x0 = cp[0], x1 = cp[1];
							const uint32_t cc[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
							s_orgb[COLOR ? w : 0][COLOR ? i2 : 0] = rms_rgb(cc, 8);
						}
						s_vox += __popc(m8 | h8);
						s_hit += __popc(h8);
						s_oct += 1;
					}
				}
				__syncwarp();
			}
			// block lanes: depth-1 sector, depth-2 aggregate, meta, mask clearing
			float my_occ = old2.occ;
			uint32_t my_fl = old2.flags, my_rgb = old2rgb;
			if (marked) {
				const float old1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
				const uint32_t oldc[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
				float new1[8];
				uint32_t newc[8];
				float bmax = -3.402823466e+38f;
				uint32_t bfl = 0, newmeta = 0, at = excl;
#pragma unroll
				for (uint32_t o = 0; o < 8; ++o) {
					float om = 0.0f;
					uint32_t fl = M.default_flags, touched = 0, oc = 0;
					if ((t8 >> o) & 1u) {
						om = s_omax[w][at];
						fl = s_ofl[w][at];
						if (COLOR) oc = s_orgb[COLOR ? w : 0][COLOR ? at : 0];
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
				s_new += (mt & 0xff0000u) ? 0u : 1u;
				my_occ = bmax;
				my_fl = bfl | 0x100u;
			}
			agg_occ[half] = my_occ;
			agg_fl[half] = my_fl;
			agg_rgb[half] = my_rgb;
			__syncwarp();  // the list / result arrays are reused by the next half
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
					for (int j = 0; j < 8; ++j) cc[j] = __shfl_sync(FULL, agg_rgb[h], (lane & 24) + j);
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
				for (int j = 0; j < 4; ++j) {
					cc[j] = __shfl_sync(FULL, rgb3[0], 8 * j);
					cc[4 + j] = __shfl_sync(FULL, rgb3[1], 8 * j);
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
	}
	// counters
	for (int o = 16; o > 0; o >>= 1) {
		s_vox += __shfl_xor_sync(FULL, s_vox, o);
		s_hit += __shfl_xor_sync(FULL, s_hit, o);
		s_oct += __shfl_xor_sync(FULL, s_oct, o);
		s_blk += __shfl_xor_sync(FULL, s_blk, o);
		s_new += __shfl_xor_sync(FULL, s_new, o);
	}
	if (lane == 0) {
		unsigned long long* slot = M.ctr->stat[gw % kStatSlots];
		if (s_vox) atomicAdd(&slot[0], (unsigned long long)s_vox);
		if (s_hit) atomicAdd(&slot[1], (unsigned long long)s_hit);
		if (s_oct) atomicAdd(&slot[2], (unsigned long long)s_oct);
		if (s_blk) atomicAdd(&slot[3], (unsigned long long)s_blk);
		if (s_new) atomicAdd(&slot[4], (unsigned long long)s_new);
	}
}

// Stand-alone depth-3 / depth-4 pass over the touched list: only used when out-of-tree marks
// exist (they are applied after k_update_compact, which then skips its folded reduction).
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
