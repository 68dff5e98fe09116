// Device-wide exclusive scan and a stable LSD radix sort (8-bit digits), hand-written for sm_100a.
//
// The sort is used once per split-search problem (every Dimension is sorted once, not once per
// cluster as std::sort in ClusterSeparatorBySensitivitySpecificity.cpp:114 does) and once per sample in
// the coverage stage (reads grouped by segment while keeping SAM order).  Stability matters for the
// latter: coverage is an order-dependent fp64 sum.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v)
{
	int lane = threadIdx.x & 31;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
		if(lane >= o)
			v += t;
	}
	return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t* total, uint64_t* smem /* [33] */)
{
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint64_t incl = warp_incl_scan_u64(v);
	if(lane == 31)
		smem[warp] = incl;
	__syncthreads();
	if(warp == 0) {
		uint64_t w = (lane < nw)? smem[lane] : 0;
		uint64_t wi = warp_incl_scan_u64(w);
		smem[lane] = wi - w;
		if(lane == 31)
			smem[32] = wi;
	}
	__syncthreads();
	uint64_t r = smem[warp] + incl - v;
	*total = smem[32];
	__syncthreads();
	return r;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const Tin* __restrict__ in, uint64_t n, uint64_t* __restrict__ block_sums)
{
	__shared__ uint64_t sm[33];
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint64_t s = 0;
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++)
		if(base + i < n)
			s += (uint64_t)in[base + i];
	uint64_t total;
	block_excl_scan_u64(s, &total, sm);
	if(threadIdx.x == 0)
		block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(uint64_t* __restrict__ block_sums, uint64_t nblocks, uint64_t* __restrict__ total_out)
{
	__shared__ uint64_t sm[33];
	uint64_t carry = 0;
	for(uint64_t base = 0; base < nblocks; base += blockDim.x) {
		uint64_t i = base + threadIdx.x;
		uint64_t v = (i < nblocks)? block_sums[i] : 0;
		uint64_t total;
		uint64_t ex = block_excl_scan_u64(v, &total, sm);
		if(i < nblocks)
			block_sums[i] = carry + ex;
		carry += total;
	}
	if(threadIdx.x == 0 && total_out != nullptr)
		*total_out = carry;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const Tin* __restrict__ in, uint64_t* __restrict__ out, uint64_t n, const uint64_t* __restrict__ block_sums)
{
	__shared__ uint64_t sm[33];
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint64_t v[SCAN_ITEMS];
	uint64_t s = 0;
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++) {
		v[i] = (base + i < n)? (uint64_t)in[base + i] : 0;
		s += v[i];
	}
	uint64_t total;
	uint64_t ex = block_excl_scan_u64(s, &total, sm) + block_sums[blockIdx.x];
#pragma unroll
	for(int i = 0; i < SCAN_ITEMS; i++) {
		if(base + i < n)
			out[base + i] = ex;
		ex += v[i];
	}
}

template <typename Tin>
int scan_impl(abw_ctx* ctx, const Tin* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total)
{
	if(n == 0) {
		if(d_total)
			ABW_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint64_t), ctx->stream));
		return ABW_OK;
	}
	unsigned int nblocks = abw_div_up(n, SCAN_TILE);
	DevBuf<uint64_t> sums;
	ABW_CUDA(ctx, sums.alloc(nblocks));
	ABW_LAUNCH(ctx, k_scan_reduce<Tin>, nblocks, SCAN_THREADS, 0, d_in, n, sums.p);
	ABW_LAUNCH(ctx, k_scan_block_sums, 1, 1024, 0, sums.p, (uint64_t)nblocks, d_total);
	ABW_LAUNCH(ctx, k_scan_apply<Tin>, nblocks, SCAN_THREADS, 0, d_in, d_out, n, sums.p);
	return ABW_OK;                                       // sums is freed in stream order
}

// ---------------------------------------------------------------------------------------------------
// radix sort
// ---------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                       // keys per lane
constexpr int RS_WARP_TILE = 32 * RS_ITEMS;        // 512
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 4096

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_or_and(const K* __restrict__ keys, uint64_t n, uint64_t stride, unsigned long long* __restrict__ or_and)
{
	const K* a = keys + (uint64_t)blockIdx.y * stride;
	unsigned long long o = 0, an = ~0ull;
	for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long k = (unsigned long long)a[i];
		o |= k;
		an &= k;
	}
#pragma unroll
	for(int s = 16; s > 0; s >>= 1) {
		o |= __shfl_xor_sync(0xffffffffu, o, s);
		an &= __shfl_xor_sync(0xffffffffu, an, s);
	}
	if((threadIdx.x & 31) == 0) {
		atomicOr(&or_and[0], o);
		atomicAnd(&or_and[1], an);
	}
}

// per-block digit histogram, layout hist[((array * 256) + digit) * nblocks + block]
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K* __restrict__ keys, uint64_t n, uint64_t stride, int shift, uint32_t* __restrict__ hist)
{
	__shared__ uint32_t sh[256];
	const K* a = keys + (uint64_t)blockIdx.y * stride;
	sh[threadIdx.x] = 0;
	__syncthreads();
	uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
	for(int r = 0; r < RS_ITEMS; r++) {
		uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		if(i < n)
			atomicAdd(&sh[(uint32_t)(a[i] >> shift) & 0xFFu], 1u);
	}
	__syncthreads();
	hist[((uint64_t)blockIdx.y * 256 + threadIdx.x) * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
                                                          uint32_t* __restrict__ vals_out, uint64_t n, uint64_t stride, int shift, const uint64_t* __restrict__ offsets)
{
	__shared__ uint32_t cnt[RS_WARPS][257];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t arr = blockIdx.y;
	const K* ki = keys_in + arr * stride;
	const uint32_t* vi = vals_in + arr * stride;
	K* ko = keys_out + arr * stride;
	uint32_t* vo = vals_out + arr * stride;
	for(int d = lane; d < 257; d += 32)
		cnt[warp][d] = 0;
	__syncwarp();
	K key[RS_ITEMS];
	uint32_t lrank[RS_ITEMS];
	const uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * RS_WARP_TILE;
	const uint32_t lt = (1u << lane) - 1u;
	// phase A: stable rank of every key among the keys of its warp with the same digit
#pragma unroll
	for(int r = 0; r < RS_ITEMS; r++) {
		uint64_t i = wbase + (uint64_t)r * 32 + lane;
		bool ok = i < n;
		key[r] = ok? ki[i] : (K)0;
		uint32_t d = ok? ((uint32_t)(key[r] >> shift) & 0xFFu) : 256u;
		uint32_t peers = __match_any_sync(0xffffffffu, d);
		int leader = __ffs(peers) - 1;
		uint32_t old = 0;
		if(lane == leader) {
			old = cnt[warp][d];
			cnt[warp][d] = old + __popc(peers);
		}
		old = __shfl_sync(0xffffffffu, old, leader);
		lrank[r] = old + __popc(peers & lt);
		__syncwarp();
	}
	__syncthreads();
	// phase B: thread t owns digit t: destination base of every (warp, digit)
	{
		int d = threadIdx.x;
		uint64_t run = offsets[(arr * 256 + d) * gridDim.x + blockIdx.x] - arr * n;
		// destinations of one block span < 2^32 only relative to the block base, so keep the low bits relative to `run`
		uint32_t acc = 0;
#pragma unroll
		for(int w = 0; w < RS_WARPS; w++) {
			uint32_t c = cnt[w][d];
			cnt[w][d] = acc;
			acc += c;
		}
		// stash the 64-bit base of the digit in shared memory through two words
		__shared__ uint64_t digit_base[256];
		digit_base[d] = run;
		__syncthreads();
		// phase C
#pragma unroll
		for(int r = 0; r < RS_ITEMS; r++) {
			uint64_t i = wbase + (uint64_t)r * 32 + lane;
			if(i < n) {
				uint32_t dg = (uint32_t)(key[r] >> shift) & 0xFFu;
				uint64_t dst = digit_base[dg] + cnt[warp][dg] + lrank[r];
				ko[dst] = key[r];
				vo[dst] = vi[i];
			}
		}
	}
}

template <typename K>
int radix_sort_impl(abw_ctx* ctx, K* d_keys, K* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, uint32_t batch, uint64_t stride, int nbits)
{
	if(n == 0 || batch == 0 || nbits <= 0)
		return ABW_OK;
	unsigned int nblocks = abw_div_up(n, RS_TILE);
	DevBuf<unsigned long long> or_and;
	ABW_CUDA(ctx, or_and.alloc(2));
	unsigned long long init[2] = {0ull, ~0ull};
	ABW_CUDA(ctx, cudaMemcpyAsync(or_and.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
	{
		dim3 grid(min(nblocks, 4u * (unsigned)ctx->sm_count), batch);
		ABW_LAUNCH(ctx, k_rs_or_and<K>, grid, RS_THREADS, 0, d_keys, n, stride, or_and.p);
	}
	unsigned long long oa[2];
	ABW_CUDA(ctx, cudaMemcpyAsync(oa, or_and.p, sizeof(oa), cudaMemcpyDeviceToHost, ctx->stream));
	ABW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	unsigned long long varying = oa[0] ^ oa[1];

	DevBuf<uint32_t> hist;
	DevBuf<uint64_t> offs;
	uint64_t nh = (uint64_t)batch * 256 * nblocks;
	ABW_CUDA(ctx, hist.alloc(nh));
	ABW_CUDA(ctx, offs.alloc(nh));
	K* src_k = d_keys; K* dst_k = d_keys_tmp;
	uint32_t* src_v = d_vals; uint32_t* dst_v = d_vals_tmp;
	dim3 grid(nblocks, batch);
	for(int shift = 0; shift < nbits; shift += 8) {
		if(((varying >> shift) & 0xFFull) == 0)
			continue;                               // every key has the same digit here: the pass is the identity
		ABW_LAUNCH(ctx, k_rs_hist<K>, grid, RS_THREADS, 0, src_k, n, stride, shift, hist.p);
		ABW_CHECK(abw_exclusive_scan_u32_to_u64(ctx, hist.p, offs.p, nh, nullptr));
		ABW_LAUNCH(ctx, k_rs_scatter<K>, grid, RS_THREADS, 0, src_k, src_v, dst_k, dst_v, n, stride, shift, offs.p);
		K* tk = src_k; src_k = dst_k; dst_k = tk;
		uint32_t* tv = src_v; src_v = dst_v; dst_v = tv;
	}
	if(src_k != d_keys) {
		uint64_t total = (uint64_t)(batch - 1) * stride + n;
		ABW_CUDA(ctx, cudaMemcpyAsync(d_keys, src_k, total * sizeof(K), cudaMemcpyDeviceToDevice, ctx->stream));
		ABW_CUDA(ctx, cudaMemcpyAsync(d_vals, src_v, total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
	}
	return ABW_OK;
}

}  // namespace

int abw_exclusive_scan_u64(abw_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total)
{
	return scan_impl<uint64_t>(ctx, d_in, d_out, n, d_total);
}

int abw_exclusive_scan_u32_to_u64(abw_ctx* ctx, const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total)
{
	return scan_impl<uint32_t>(ctx, d_in, d_out, n, d_total);
}

int abw_radix_sort_pairs_u64(abw_ctx* ctx, uint64_t* d_keys, uint64_t* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, uint32_t batch,
                             uint64_t stride, int nbits)
{
	return radix_sort_impl<uint64_t>(ctx, d_keys, d_keys_tmp, d_vals, d_vals_tmp, n, batch, stride, nbits);
}

int abw_radix_sort_pairs_u32(abw_ctx* ctx, uint32_t* d_keys, uint32_t* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, int nbits)
{
	return radix_sort_impl<uint32_t>(ctx, d_keys, d_keys_tmp, d_vals, d_vals_tmp, n, 1, n, nbits);
}
