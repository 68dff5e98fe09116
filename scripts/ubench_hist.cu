// micro-benchmark: private per-lane histogram updates in shared memory, bank = lane (conflict free)
//   mode 0: LDS.U8 + IADD + STS.U8     (byte counters, what k_kmer round 1 uses)
//   mode 1: ATOMS.ADD u32 without return (word counters)
//   mode 2: ATOMS.ADD u32 of 1<<(8*(x&3)) on packed byte counters
//   mode 3: LDS + IADD + STS 32-bit words
#include <cstdio>
#include <cuda_runtime.h>
__host__ __device__ inline int words_per_warp(int mode)
{
	switch(mode) {
	case 1: case 3: return 256 * 32;
	case 6: return 256 * 16;
	case 7: return 128 * 8;
	case 8: return 128 * 4;
	case 9: return 512 * 4;
	case 10: return 512 * 8;
	default: return 64 * 32;
	}
}
__global__ void k(int mode, int iters, unsigned* out, unsigned seed)
{
	extern __shared__ unsigned sm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int words = words_per_warp(mode);
	unsigned* h = sm + warp * words;
	for(int i = lane; i < words; i += 32) h[i] = 0;
	__syncwarp();
	unsigned x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
	for(int it = 0; it < iters; it++) {
#pragma unroll
		for(int t = 0; t < 16; t++) {
			x = x * 1664525u + 1013904223u;
			unsigned c = (x >> 13) & 0xFFu;
			if(mode == 0) {
				unsigned char* b = reinterpret_cast<unsigned char*>(h + (c & 63u) * 32 + lane) + (c >> 6);
				*b = (unsigned char)(*b + 1);
			}
			else if(mode == 1)
				atomicAdd(h + c * 32 + lane, 1u);
			else if(mode == 2)
				atomicAdd(h + (c >> 2) * 32 + lane, 1u << ((c & 3u) * 8));
			else if(mode == 3) {
				unsigned* p = h + c * 32 + lane;
				*p = *p + 1;
			}
			else if(mode == 4)
				atomicAdd(h + c * 8 + (lane >> 2), 1u);          // 8 sub-histograms per warp (4 lanes each), bin-major: only lanes of one group can collide
			else if(mode == 5)
				atomicAdd(h + c * 4 + (lane >> 3), 1u);          // 4 sub-histograms (8 lanes each)
			else if(mode == 6)
				atomicAdd(h + c * 16 + (lane >> 1), 1u);         // 16 sub-histograms (2 lanes each)
			else if(mode == 7)
				atomicAdd(h + (c >> 1) * 8 + (lane >> 2), 1u << ((c & 1u) * 16));   // 16-bit counters, two bins per word, 8 sub-histograms
			else if(mode == 8)
				atomicAdd(h + (c >> 1) * 4 + (lane >> 3), 1u << ((c & 1u) * 16));   // 16-bit counters, 4 sub-histograms
			else if(mode == 9) {
				unsigned c10 = (x >> 13) & 0x3FFu;               // 5-mer bins (one update covers two 4-mer starts), 16-bit counters, 4 sub-histograms
				atomicAdd(h + (c10 >> 1) * 4 + (lane >> 3), 1u << ((c10 & 1u) * 16));
			}
			else {
				unsigned c10 = (x >> 13) & 0x3FFu;               // 5-mer bins, 16-bit counters, 8 sub-histograms
				atomicAdd(h + (c10 >> 1) * 8 + (lane >> 2), 1u << ((c10 & 1u) * 16));
			}
		}
	}
	__syncwarp();
	unsigned s = 0;
	for(int i = lane; i < words; i += 32) s += h[i];
	if(s == 0xdeadbeef) out[0] = s;
}
int main()
{
	unsigned* out; cudaMalloc(&out, 4);
	const int iters = 2000;
	for(int mode = 0; mode < 11; mode++) {
		const int warps = (mode == 1 || mode == 3) ? 4 : 8;            // 32 KB vs 8 KB per warp
		const size_t smem = (size_t)warps * words_per_warp(mode) * 4;
		cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		int blocks_per_sm = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k, warps * 32, smem);
		const int grid = 148 * blocks_per_sm;
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		k<<<grid, warps * 32, smem>>>(mode, 10, out, 1);
		cudaEventRecord(a);
		k<<<grid, warps * 32, smem>>>(mode, iters, out, 2);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		double updates = (double)grid * warps * 32 * iters * 16;
		printf("mode %d: warps/SM %d  %.3f ms  %.1f G updates/s  = %.2f updates/clk/SM @1.9GHz  (%s)\n", mode, blocks_per_sm * warps, ms, updates / ms / 1e6,
		       updates / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
