// Shared host/device helpers of libabawaca_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header only: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <cstdio>
#include "../../include/abawaca_b200.h"

struct abw_ctx {
	int          device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr;       // host->device staging that overlaps kernels of `stream` (abw_h2d_async)
	cudaStream_t side_stream = nullptr;       // kernels of a search level that nothing on `stream` waits for until a later level (search.cu: k_finalize_terminal)
	cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
	std::vector<cudaEvent_t> copy_events;     // events of the tickets copy_retired + 1 ... (abw_h2d_async)
	uint64_t     copy_retired = 0;            // tickets retired by abw_ctx_synchronize so far
	int          sm_count = 148;
	uint64_t     launches = 0;
	uint64_t     arena_misses = 0;             // device blocks obtained from the driver (cudaMallocAsync) rather than from the context's cache
	std::string  err;
	// optional per-kernel timing (abw_profile_enable): every launch is bracketed by events and waited for
	bool         profiling = false;
	cudaEvent_t  ev_a = nullptr, ev_b = nullptr;
	std::map<std::string, std::pair<uint64_t, double>> prof;   // kernel -> (launches, total ms)
	// device blocks freed by this context, kept for reuse (see abw_arena_alloc): capacity -> block
	std::multimap<size_t, void*> free_blocks;
	std::map<void*, size_t> block_cap;                         // capacity of every block obtained through the arena, cached or in use
	// pinned staging for the many small host->device uploads of a search level: two halves used alternately (see abw_stage_flip)
	unsigned char* h_stage = nullptr;
	size_t stage_half = 0, stage_used = 0;
	int stage_side = 0;
	// small device->host results fetched by a kernel into mapped pinned memory instead of through the copy engine (abw_fetch, abw_sync)
	struct Fetch { void* h_dst; const unsigned char* slot; size_t bytes; };
	unsigned char* h_bounce = nullptr;
	size_t bounce_cap = 0, bounce_used = 0;
	std::vector<Fetch> pending;
	// progress words of a running split search, written by the device (mapped pinned memory, search.cu)
	void* h_prog = nullptr;
	// scratch of the one-pass prefix sums (scan_sort.cu): ticket, tile sums, tile prefixes, status words
	void* scan_scratch = nullptr;
	size_t scan_cap = 0;
	uint32_t scan_epoch = 0;
	uint64_t scan_tickets = 0;
};

// Small uploads (descriptors, job lists, tile tables) go through pinned memory so that cudaMemcpyAsync really is asynchronous.
// abw_stage_flip switches to the other half; the caller guarantees that the copies enqueued two flips ago have completed
// (the split search synchronises at least once per level and flips once per level).
void abw_stage_flip(abw_ctx* ctx);
cudaError_t abw_stage_upload(abw_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

// Small transfers and the copy engines.  Copies of one direction execute in the order they were issued, whatever their stream: an 8-byte
// cudaMemcpyAsync on the compute stream waits for every bulk copy (read records in, .lrn matrix out) that was enqueued before it on the copy
// stream, and the kernels behind it wait too -- measured: end-to-end step = H2D + resident step + D2H, no overlap at all.
// Therefore the small transfers bypass the engines: abw_stage_upload lets a kernel read the pinned staging slot, and
// abw_fetch lets a kernel write the result into mapped pinned memory, from where abw_sync (cudaStreamSynchronize + copy-out) hands it to the
// caller's buffer.  ABW_SMALL_COPIES=memcpy selects plain cudaMemcpyAsync (the round-1 behaviour).  Every device->host result of an entry point
// is requested with abw_fetch and every wait on the context stream is abw_sync; ABW_ENTER drops requests an earlier, failed call may have left behind.
bool abw_small_copies_by_kernel();
cudaError_t abw_fetch(abw_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
cudaError_t abw_sync(abw_ctx* ctx);
// host->device upload of a small array on the context stream: cudaMemcpyAsync by default, through the pinned staging area and a kernel (cudaMemcpyAsync with
// ABW_SMALL_COPIES=memcpy) (the caller waits for the stream before it changes h_src or returns)
cudaError_t abw_upload_small(abw_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

inline int abw_fail(abw_ctx* ctx, int code, const std::string& msg)
{
	if(ctx)
		ctx->err = msg;
	return code;
}

#define ABW_CUDA(ctx, call)                                                                                          \
	do {                                                                                                             \
		cudaError_t e__ = (call);                                                                                    \
		if(e__ != cudaSuccess) {                                                                                     \
			char b__[512];                                                                                           \
			snprintf(b__, sizeof(b__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
			return abw_fail((ctx), ABW_ERR_CUDA, b__);                                                               \
		}                                                                                                            \
	} while(0)

#define ABW_CHECK(expr)                                                                                              \
	do {                                                                                                             \
		int r__ = (expr);                                                                                            \
		if(r__ != ABW_OK)                                                                                            \
			return r__;                                                                                              \
	} while(0)

// every launch goes through this so that abw_kernel_launches() is an honest count
#define ABW_LAUNCH(ctx, kernel, grid, block, smem, ...)                                                              \
	do {                                                                                                             \
		if((ctx)->profiling)                                                                                         \
			cudaEventRecord((ctx)->ev_a, (ctx)->stream);                                                             \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                             \
		(ctx)->launches++;                                                                                           \
		ABW_CUDA((ctx), cudaGetLastError());                                                                         \
		if((ctx)->profiling) {                                                                                       \
			float ms__ = 0;                                                                                          \
			cudaEventRecord((ctx)->ev_b, (ctx)->stream);                                                             \
			ABW_CUDA((ctx), cudaEventSynchronize((ctx)->ev_b));                                                      \
			cudaEventElapsedTime(&ms__, (ctx)->ev_a, (ctx)->ev_b);                                                   \
			auto& e__ = (ctx)->prof[#kernel];                                                                        \
			e__.first++;                                                                                             \
			e__.second += ms__;                                                                                      \
		}                                                                                                            \
	} while(0)

// Context whose entry point is executing on this host thread (set by ABW_ENTER).  Device buffers come from a per-context arena: blocks are
// obtained with cudaMallocAsync on the context stream and, when released, cached by the context instead of going back to the driver, so
// that after the first pass of a given shape no entry point pays for cuMemCreate / pool growth (which showed up as sporadic 10-20 ms stalls).
// Every block is only ever used in the order of the context stream, which makes immediate reuse safe.
extern thread_local abw_ctx* abw_tls_ctx;
cudaError_t abw_arena_alloc(abw_ctx* ctx, size_t bytes, void** out);
void abw_arena_free(abw_ctx* ctx, void* p);

// one NVTX range per ABI entry point (named after the function), closed when the entry point returns
struct AbwNvtxRange {
	explicit AbwNvtxRange(const char* name) { nvtxRangePushA(name); }
	~AbwNvtxRange() { nvtxRangePop(); }
};

#define ABW_ENTER(ctx)                                                                                               \
	AbwNvtxRange nvtx_range__(__func__);                                                                             \
	do {                                                                                                             \
		ABW_CUDA((ctx), cudaSetDevice((ctx)->device));                                                               \
		abw_tls_ctx = (ctx);                                                                                         \
		(ctx)->pending.clear();                                                                                      \
		(ctx)->bounce_used = 0;                                                                                      \
	} while(0)

template <typename T>
struct DevBuf {
	T*       p = nullptr;
	size_t   n = 0;
	abw_ctx* owner = nullptr;
	DevBuf() {}
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	~DevBuf() { release(); }
	void release()
	{
		if(p)
			abw_arena_free(owner, p);
		p = nullptr;
		n = 0;
	}
	cudaError_t alloc(size_t count)
	{
		release();
		n = count;
		owner = abw_tls_ctx;
		return abw_arena_alloc(owner, (count ? count : 1) * sizeof(T), (void**)&p);
	}
};

static inline unsigned int abw_div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

// ---- device-wide primitives implemented in scan_sort.cu -------------------------------------------
// exclusive prefix sum of n uint64 values (in place allowed); d_total (may be null) receives the sum
int abw_exclusive_scan_u64(abw_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total);
// exclusive prefix sum of uint32 inputs into uint64 outputs
int abw_exclusive_scan_u32_to_u64(abw_ctx* ctx, const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total);

// Stable LSD radix sort of `batch` independent arrays of n (key, value) pairs each (array b at offset b*stride).
// Only key bits [0, nbits) are examined.  Result is left in (keys, vals); (keys_tmp, vals_tmp) are scratch of the same size.
// Digits are 8, 9 or 10 bits wide, placed only over key bits that actually vary (bits equal in all keys of all arrays are skipped).
int abw_radix_sort_pairs_u64(abw_ctx* ctx, uint64_t* d_keys, uint64_t* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, uint32_t batch,
                             uint64_t stride, int nbits);
int abw_radix_sort_pairs_u32(abw_ctx* ctx, uint32_t* d_keys, uint32_t* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, uint32_t batch,
                             uint64_t stride, int nbits);
// the same with the set of key bits that differ between any two keys (OR of all keys ^ AND of all keys) supplied by the caller: no inspection pass
int abw_radix_sort_pairs_u32_varying(abw_ctx* ctx, uint32_t* d_keys, uint32_t* d_keys_tmp, uint32_t* d_vals, uint32_t* d_vals_tmp, uint64_t n, uint32_t batch,
                                     uint64_t stride, unsigned long long varying);
