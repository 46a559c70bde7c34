// atomic_ceiling.cu -- micro-benchmark of the L2 atomic throughput the marking stage
// (K2b: one hash probe + one 64-bit atomicOr per ray-walk record) is bounded by.
// SURVEY.md section 8(d)(ii) asks for the marks/s of the raycast stage to be reported
// against this ceiling.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomic_ceiling tools/atomic_ceiling.cu
//   tools/atomic_ceiling            (prints one JSON line per pattern)
//
// Patterns (N operations, one per thread, all CTAs resident-sized grids):
//   red64_uniform   red.global.or.b64 to a uniformly random word of a footprint of F bytes
//   atom64_uniform  same with the old value returned (atom instead of red)
//   red32_uniform   32-bit variant
//   red64_brick     a warp's lanes fall into few 512 B groups (64 masks of one brick), like
//                   records of neighbouring rays
//   probe_red64     dependent chain of K2b: 16 B load from a 4 MB table, then the atomic
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
	do {                                                                             \
		cudaError_t e = (x);                                                           \
		if (e != cudaSuccess) {                                                        \
			fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));    \
			exit(1);                                                                     \
		}                                                                              \
	} while (0)

__device__ __forceinline__ uint64_t mix(uint64_t k)
{
	k ^= k >> 33;
	k *= 0xff51afd7ed558ccdull;
	k ^= k >> 33;
	k *= 0xc4ceb9fe1a85ec53ull;
	k ^= k >> 33;
	return k;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_atomics(unsigned long long* buf, uint64_t words, const ulonglong2* table,
                                                 uint32_t table_mask, uint64_t n, unsigned long long* sink)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint64_t h = mix(i * 0x9e3779b97f4a7c15ull + 12345);
		uint64_t w;
		if (MODE == 3) {
			// warp-coherent: 4 groups of 8 lanes, each group inside one 512 B brick row
			const uint64_t g = mix((i >> 3) * 0x9e3779b97f4a7c15ull + 777);
			w = ((g % (words / 64)) * 64) + (h & 63);
		} else {
			w = h % words;
		}
		const unsigned long long bits = 1ull << (h >> 58);
		if (MODE == 0 || MODE == 3) {
			asm volatile("red.global.or.b64 [%0], %1;" ::"l"(buf + w), "l"(bits) : "memory");
		} else if (MODE == 1) {
			acc += atomicOr(buf + w, bits);
		} else if (MODE == 2) {
			asm volatile("red.global.or.b32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(buf) + 2 * w), "r"((uint32_t)bits | 1u)
			             : "memory");
		} else if (MODE == 4) {
			const ulonglong2 e = table[(uint32_t)(h >> 20) & table_mask];
			w = (w + (e.y & 1)) % words;
			asm volatile("red.global.or.b64 [%0], %1;" ::"l"(buf + w), "l"(bits) : "memory");
		}
	}
	if (acc == 0x1234567ull) *sink = acc;
}

template <int MODE>
double run(const char* name, unsigned long long* buf, uint64_t footprint, const ulonglong2* table, uint32_t table_mask,
           uint64_t n, unsigned long long* sink, int grid)
{
	const uint64_t words = footprint / 8;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		CK(cudaMemsetAsync(buf, 0, footprint));
		CK(cudaEventRecord(e0));
		k_atomics<MODE><<<grid, 256>>>(buf, words, table, table_mask, n, sink);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (rep && ms < best) best = ms;
	}
	const double gops = (double)n / (best * 1e-3) / 1e9;
	printf("{\"pattern\": \"%s\", \"footprint_mb\": %.1f, \"ops\": %llu, \"ms\": %.4f, \"gops_per_s\": %.2f}\n", name,
	       footprint / 1048576.0, (unsigned long long)n, best, gops);
	fflush(stdout);
	return gops;
}

int main()
{
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const int grid = sms * 8;
	const uint64_t n = 26ull << 20;  // records of one bench scan
	const uint64_t max_fp = 2048ull << 20;
	unsigned long long *buf, *sink;
	ulonglong2* table;
	const uint32_t table_entries = 1u << 18;  // 4 MB of 16 B entries
	CK(cudaMalloc(&buf, max_fp));
	CK(cudaMalloc(&sink, 8));
	CK(cudaMalloc(&table, (size_t)table_entries * 16));
	CK(cudaMemset(table, 0x5a, (size_t)table_entries * 16));
	const uint64_t fps[] = {4ull << 20, 32ull << 20, 64ull << 20, 128ull << 20, 512ull << 20, 2048ull << 20};
	for (uint64_t fp : fps) {
		run<0>("red64_uniform", buf, fp, table, table_entries - 1, n, sink, grid);
		run<1>("atom64_uniform", buf, fp, table, table_entries - 1, n, sink, grid);
		run<2>("red32_uniform", buf, fp, table, table_entries - 1, n, sink, grid);
		run<3>("red64_brick", buf, fp, table, table_entries - 1, n, sink, grid);
		run<4>("probe_red64", buf, fp, table, table_entries - 1, n, sink, grid);
	}
	return 0;
}
