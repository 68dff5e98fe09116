// Feature stage of abawaca-build on the device: 2-bit packing, windowing, k-mer signatures, coverage.
//
// Data layout in HBM (abw_seqset): every scaffold starts at a base offset that is a multiple of 128, so that
//   packed : 2 bits/base, 16 bases per u32, base i of the scaffold at bits [2*(i%16), +2) of word base/16+i/16
//            (A=0 C=1 G=2 T=3; non-ACGT positions hold 0),
//   valid  : 1 bit/base, set for A/C/G/T (after upper-casing),
//   nmask  : 1 bit/base, set for the literal 'N' only (quirk Q2: IUPAC codes are bases for windowing, but break k-mers)
// are all 16-byte aligned per scaffold and can be fetched with 128-bit loads.
#include "common.cuh"
#include <algorithm>

struct abw_seqset {
	uint32_t nscaf = 0;
	uint64_t total_padded = 0;            // bases
	DevBuf<uint64_t> len, base;           // [nscaf], [nscaf+1]
	DevBuf<uint32_t> packed, valid, nmask;
	DevBuf<unsigned long long> countN, countGC;
	std::vector<uint64_t> h_len, h_base;
};

struct abw_segments {
	uint32_t nscaf = 0;
	uint64_t nseg = 0;
	DevBuf<uint64_t> seg_first;           // [nscaf+1]
	DevBuf<uint32_t> seg_scaf;            // [nseg]
	DevBuf<uint64_t> seg_start, seg_end, seg_nonN;   // 1-based inclusive, abawaca-build.cpp:216
	DevBuf<uint64_t> seg_gbase;           // absolute (padded) base index of the first base of the segment
	DevBuf<uint4> scaf_info;              // per scaffold {first window (low 32 bits), windows, non-N bases per window (u64)}: one load for the coverage kernels
};

namespace {

// ---------------------------------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather_bits_0_8_16_24(uint32_t v)   // bit0 of each byte -> 4 bits
{
	uint32_t t = (v | (v >> 7)) & 0x00030003u;
	return (t | (t >> 14)) & 0xFu;
}

// 4 ASCII characters -> 8 bits of 2-bit codes + 4 valid bits + 4 N bits (+ GC count, + lower-case-n flag)
__device__ __forceinline__ void classify4(uint32_t w, uint32_t& code8, uint32_t& valid4, uint32_t& n4, uint32_t& gc, uint32_t& bad)
{
	bad |= __vcmpeq4(w, 0x6E6E6E6Eu);                       // 'n' throws in the reference (String.cpp:47-49)
	// toupper() for letters only (String.cpp:44); bytes outside a-z are left alone
	uint32_t is_lower = __vcmpgeu4(w, 0x61616161u) & __vcmpleu4(w, 0x7A7A7A7Au);
	uint32_t u = w & ~(is_lower & 0x20202020u);
	uint32_t eA = __vcmpeq4(u, 0x41414141u), eC = __vcmpeq4(u, 0x43434343u), eG = __vcmpeq4(u, 0x47474747u), eT = __vcmpeq4(u, 0x54545454u);
	uint32_t eN = __vcmpeq4(u, 0x4E4E4E4Eu);
	uint32_t lo = (eC | eT) & 0x01010101u, hi = (eG | eT) & 0x01010101u;
	uint32_t x = lo | (hi << 1);                            // one 2-bit code per byte
	uint32_t y = (x | (x >> 6)) & 0x000F000Fu;
	code8 = (y | (y >> 12)) & 0xFFu;
	valid4 = gather_bits_0_8_16_24((eA | eC | eG | eT) & 0x01010101u);
	n4 = gather_bits_0_8_16_24(eN & 0x01010101u);
	gc += __popc((eC | eG) & 0x01010101u);
}

// one warp per scaffold, one lane per 32-base unit
__global__ void __launch_bounds__(256) k_pack(const unsigned char* __restrict__ ascii, uint64_t ascii_bytes, const uint64_t* __restrict__ offsets,
                                              const uint64_t* __restrict__ base, uint32_t nscaf, uint32_t* __restrict__ packed, uint32_t* __restrict__ valid,
                                              uint32_t* __restrict__ nmask, unsigned long long* __restrict__ countN, unsigned long long* __restrict__ countGC,
                                              int* __restrict__ err)
{
	const int lane = threadIdx.x & 31;
	const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for(uint64_t s = warp0; s < nscaf; s += nwarps) {
		const uint64_t off = offsets[s], len = offsets[s + 1] - off, b0 = base[s], units = (base[s + 1] - b0) >> 5;
		uint32_t nN = 0, nGC = 0, bad = 0;
		for(uint64_t u = lane; u < units; u += 32) {
			uint64_t p = u << 5;                            // first base of the unit within the scaffold
			uint32_t code_lo = 0, code_hi = 0, v = 0, nm = 0;
			if(p < len) {
				uint64_t a = off + p;
				uint32_t cnt = (uint32_t)min((uint64_t)32, len - p);
				uint32_t w[8];
				if(cnt == 32 && ((a & ~3ull) + 36 <= ascii_bytes)) {
					const uint32_t* src = (const uint32_t*)(ascii + (a & ~3ull));
					uint32_t sh = (uint32_t)(a & 3) * 8;
					uint32_t t[9];
#pragma unroll
					for(int i = 0; i < 9; i++)
						t[i] = __ldg(src + i);
#pragma unroll
					for(int i = 0; i < 8; i++)
						w[i] = __funnelshift_r(t[i], t[i + 1], sh);
				}
				else {
#pragma unroll
					for(int i = 0; i < 8; i++) {
						uint32_t x = 0;
#pragma unroll
						for(int j = 0; j < 4; j++) {
							uint32_t k = i * 4 + j;
							uint32_t c = (k < cnt)? (uint32_t)ascii[a + k] : 0u;
							x |= c << (8 * j);
						}
						w[i] = x;
					}
				}
#pragma unroll
				for(int i = 0; i < 8; i++) {
					uint32_t c8, v4, n4;
					classify4(w[i], c8, v4, n4, nGC, bad);
					if(i < 4) code_lo |= c8 << (8 * i); else code_hi |= c8 << (8 * (i - 4));
					v |= v4 << (4 * i);
					nm |= n4 << (4 * i);
				}
				nN += __popc(nm);
			}
			uint64_t unit = (b0 >> 5) + u;
			packed[2 * unit] = code_lo;
			packed[2 * unit + 1] = code_hi;
			valid[unit] = v;
			nmask[unit] = nm;
		}
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) {
			nN += __shfl_xor_sync(0xffffffffu, nN, o);
			nGC += __shfl_xor_sync(0xffffffffu, nGC, o);
			bad |= __shfl_xor_sync(0xffffffffu, bad, o);
		}
		if(lane == 0) {
			countN[s] = nN;
			countGC[s] = nGC;
			if(bad)
				atomicExch(err, 1);
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// windowing, abawaca-build.cpp:198-228
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void window_plan(uint64_t len, uint64_t nN, uint64_t window, uint64_t& nbps, uint64_t& count)
{
	uint64_t nonN = len - nN;
	uint64_t nsegments = nonN / window;                     // :202
	if(nsegments == 0)
		nsegments = 1;                                      // :204-205
	nbps = nonN / nsegments;                                // :206
	// a window closes every time nbps non-N characters were seen (:210-224); with no non-N base at all the
	// test `nbps_segment == nbps` is 0 == 0 at every character, so every character becomes a window
	count = (nbps > 0)? nonN / nbps : len;
}

constexpr uint32_t COV_REGULAR = 0x80000000u;              // scaf_info.w of a scaffold whose windows are all nbps characters long (no N)

__global__ void k_seg_count(const uint64_t* __restrict__ len, const unsigned long long* __restrict__ countN, uint32_t nscaf, uint64_t window, uint64_t* __restrict__ counts)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s < nscaf) {
		uint64_t nbps, count;
		window_plan(len[s], countN[s], window, nbps, count);
		counts[s] = count;
	}
}

// one warp per scaffold walks the N mask; window j ends at the (j*nbps)-th non-N character
__global__ void __launch_bounds__(256) k_seg_fill(const uint64_t* __restrict__ len, const unsigned long long* __restrict__ countN, const uint64_t* __restrict__ base,
                                                  const uint32_t* __restrict__ nmask, uint32_t nscaf, uint64_t window, const uint64_t* __restrict__ seg_first,
                                                  uint32_t* __restrict__ seg_scaf, uint64_t* __restrict__ seg_start, uint64_t* __restrict__ seg_end,
                                                  uint64_t* __restrict__ seg_nonN, uint64_t* __restrict__ seg_gbase, uint4* __restrict__ scaf_info)
{
	const int lane = threadIdx.x & 31;
	const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for(uint64_t s = warp0; s < nscaf; s += nwarps) {
		const uint64_t L = len[s], b0 = base[s], first = seg_first[s];
		uint64_t nbps, count;
		window_plan(L, countN[s], window, nbps, count);
		if(lane == 0)
			{
			// .w: bits 32.. of nbps (coverage then searches instead of guessing), or COV_REGULAR for a scaffold without N whose windows
			// are exactly [j nbps + 1, (j + 1) nbps]: coverage then needs no window table at all
			uint32_t hi = (uint32_t)min(nbps >> 32, (uint64_t)0x7FFFFFFFu);
			if(countN[s] == 0 && nbps > 0 && hi == 0)
				hi = COV_REGULAR;
			scaf_info[s] = make_uint4((uint32_t)first, (uint32_t)min(count, (uint64_t)0xFFFFFFFFu), (uint32_t)nbps, hi);
		}
		if(count == 0)
			continue;
		if(nbps == 0) {
			for(uint64_t i = lane; i < count; i += 32) {
				seg_scaf[first + i] = (uint32_t)s;
				seg_start[first + i] = i + 1;
				seg_end[first + i] = i + 1;
				seg_nonN[first + i] = 0;
				seg_gbase[first + i] = b0 + i;
			}
			continue;
		}
		if(lane == 0) {
			seg_start[first] = 1;
			seg_gbase[first] = b0;
		}
		for(uint64_t i = lane; i < count; i += 32) {
			seg_scaf[first + i] = (uint32_t)s;
			seg_nonN[first + i] = nbps;
		}
		const uint64_t units = (L + 31) >> 5;
		uint64_t carry = 0;
		for(uint64_t u0 = 0; u0 < units; u0 += 32) {
			uint64_t u = u0 + lane;
			uint32_t bits = 0;
			if(u < units) {
				uint32_t in_range = (L - (u << 5) >= 32)? 0xFFFFFFFFu : ((1u << (uint32_t)(L - (u << 5))) - 1u);
				bits = ~nmask[(b0 >> 5) + u] & in_range;
			}
			uint32_t c = __popc(bits), incl = c;
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
				if(lane >= o)
					incl += t;
			}
			uint64_t before = carry + incl - c, after = carry + incl;
			for(uint64_t j = before / nbps + 1; j * nbps <= after && j <= count; j++) {
				uint32_t k = (uint32_t)(j * nbps - before);          // the k-th set bit of this word closes window j
				uint32_t bit = __fns(bits, 0, k);
				uint64_t pos1 = (u << 5) + bit + 1;                   // 1-based, :216
				seg_end[first + j - 1] = pos1;
				if(j < count) {
					seg_start[first + j] = pos1 + 1;                  // :221-222
					seg_gbase[first + j] = b0 + pos1;
				}
			}
			carry += __shfl_sync(0xffffffffu, incl, 31);
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// k-mer signature, abawaca-build.cpp:103-174
// ---------------------------------------------------------------------------------------------------
// dimension tables: for each of the 180 canonical dims the k, and the two little-endian codes (mer, reverse complement)
// Kept in GLOBAL memory and copied to shared memory by every CTA: lanes index them with different addresses, which the
// indexed-constant path serialises (round 1 ncu: ADU pipe 72 % busy with these lookups when they were __constant__).
__device__ uint32_t g_dim_tab[ABW_NKMER_DIMS];              // k | code_a << 8 | code_b << 16
__device__ double   g_milli[1001];                          // m / 1000.0

constexpr int KM_WARPS = 8;
constexpr int KM_GROUPS = 4;                                // sub-histograms per warp: lanes 8g..8g+7 share one

// Per warp.  hist is bin-major and interleaved by group: word = bin * 4 + group, so two lanes can only hit the same bank when they are in the
// same group of eight (micro-benchmark scripts/ubench_hist.cu on B200, updates per clock and SM: 13.3 in this layout with fire-and-forget
// shared atomics, 14.0 with eight groups at twice the footprint, 8.6 for a private byte histogram per lane updated with load/add/store,
// 10.4 for 16-bit counters or for 5-mer bins counted at every second position).  5.4 KB per warp: five CTAs of eight warps per SM.
struct __align__(16) KmWarpSmem {
	uint32_t hist[256 * KM_GROUPS];
	uint32_t cnt4[256];
	uint32_t cnt3[64];
	uint32_t cnt2[16];
	uint32_t cnt1[4];
	uint32_t tot[4];
	float    inv_tot[4];            // 1 / total_k, only used to form an integer quotient that is then verified exactly
	uint32_t pad[4];
};

// A warp per window; lane l takes the 16-base words l, l+32, ... of the window straight from global memory (the packed stream and the
// validity bits of neighbouring words share cache lines).  A word whose 16 four-mers are all valid costs 16 shared atomics; the first and
// last words of the window and words next to a non-ACGT character go through the per-position path.
template <int KIND>
__global__ void __launch_bounds__(KM_WARPS * 32) k_kmer(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, const uint64_t* __restrict__ seg_gbase,
                                                        const uint64_t* __restrict__ seg_start, const uint64_t* __restrict__ seg_end, uint64_t nseg, int skip_A,
                                                        double* __restrict__ rows, uint64_t ld, uint32_t col0)
{
	__shared__ uint32_t sm_dim_tab[ABW_NKMER_DIMS];
	__shared__ KmWarpSmem sm_warp[KM_WARPS];
	KmWarpSmem* sm = sm_warp + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	for(int i = threadIdx.x; i < ABW_NKMER_DIMS; i += blockDim.x)
		sm_dim_tab[i] = g_dim_tab[i];
	for(int i = lane; i < 256 * KM_GROUPS; i += 32)
		sm->hist[i] = 0;
	__syncthreads();
	uint32_t* const myhist = sm->hist + (lane >> 3);        // + bin * KM_GROUPS
	const uint64_t warp0 = (uint64_t)blockIdx.x * KM_WARPS + (threadIdx.x >> 5), nwarps = (uint64_t)gridDim.x * KM_WARPS;
	for(uint64_t g = warp0; g < nseg; g += nwarps) {
		const uint64_t gb = seg_gbase[g];
		const uint64_t pos_end = gb + (seg_end[g] - seg_start[g] + 1);
		for(int i = lane; i < 64 + 16 + 4 + 4; i += 32)
			sm->cnt3[i] = 0;                                // cnt3, cnt2, cnt1, tot are contiguous: these receive the run-end windows
		__syncwarp();
		const uint64_t wend = (pos_end + 15) >> 4;
		for(uint64_t wa = (gb >> 4) + lane; wa < wend; wa += 32) {
			const uint32_t w_lo = __ldg(packed + wa), w_hi = __ldg(packed + wa + 1);
			const uint32_t va = __ldg(valid + (wa >> 1)), vb = __ldg(valid + (wa >> 1) + 1);
			uint32_t v = __funnelshift_r(va, vb, (uint32_t)(wa & 1) * 16) & 0x7FFFFu;       // 16 positions + 3 lookahead
			const uint64_t p0 = wa << 4;
			// positions outside [gb, pos_end) never hold a base of this window
			if(p0 < gb)
				v &= 0xFFFFFFFFu << (uint32_t)(gb - p0);
			if(pos_end - p0 < 19)
				v &= (1u << (uint32_t)(pos_end - p0)) - 1u;
			const uint32_t v1 = v & 0xFFFFu, v2 = v1 & (v >> 1), v3 = v2 & (v >> 2), v4 = v3 & (v >> 3);
			// s[i] = the window shifted by i bases: 4-mer t = 4 j + i is byte j of s[i]
			uint32_t sh[4];
			sh[0] = w_lo;
			sh[1] = __funnelshift_r(w_lo, w_hi, 2);
			sh[2] = __funnelshift_r(w_lo, w_hi, 4);
			sh[3] = __funnelshift_r(w_lo, w_hi, 6);
			if(v4 == 0xFFFFu) {
#pragma unroll
				for(int i = 0; i < 4; i++) {
#pragma unroll
					for(int j = 0; j < 4; j++)
						atomicAdd(myhist + ((sh[i] >> (8 * j)) & 0xFFu) * KM_GROUPS, 1u);
				}
			}
			else if(v1) {
				// windows broken by a non-ACGT character or by the end of the segment (:131,147 key = 0)
				const uint32_t e3 = v3 & ~v4, e2 = v2 & ~v3, e1 = v1 & ~v2;
#pragma unroll
				for(int i = 0; i < 4; i++) {
#pragma unroll
					for(int j = 0; j < 4; j++) {
						const int t = 4 * j + i;
						const uint32_t x = (sh[i] >> (8 * j)) & 0xFFu;
						if((v4 >> t) & 1u)
							atomicAdd(myhist + x * KM_GROUPS, 1u);
						else if((e3 >> t) & 1u)
							atomicAdd(&sm->cnt3[x & 63u], 1u);
						else if((e2 >> t) & 1u)
							atomicAdd(&sm->cnt2[x & 15u], 1u);
						else if((e1 >> t) & 1u)
							atomicAdd(&sm->cnt1[x & 3u], 1u);
					}
				}
			}
		}
		__syncwarp();
		// fold the sub-histograms: lane l owns bins l, l+32, ..., l+224 (four consecutive words each) and clears them
		{
			uint32_t t4 = 0;
#pragma unroll
			for(int j = 0; j < 8; j++) {
				const int bin = lane + 32 * j;
				uint4* hp = reinterpret_cast<uint4*>(sm->hist + bin * KM_GROUPS);
				const uint4 a = hp[0];
				hp[0] = make_uint4(0, 0, 0, 0);
				const uint32_t c = a.x + a.y + a.z + a.w;
				sm->cnt4[bin] = c;
				t4 += c;
			}
			t4 = __reduce_add_sync(0xffffffffu, t4);
			__syncwarp();
			// lower orders: every valid 4-mer start is a valid 3-mer start, etc.; cnt3/cnt2/cnt1 hold the run-end extras
			uint32_t a = sm->cnt4[lane] + sm->cnt4[lane + 64] + sm->cnt4[lane + 128] + sm->cnt4[lane + 192];
			uint32_t b = sm->cnt4[lane + 32] + sm->cnt4[lane + 96] + sm->cnt4[lane + 160] + sm->cnt4[lane + 224];
			a += sm->cnt3[lane];
			b += sm->cnt3[lane + 32];
			sm->cnt3[lane] = a;
			sm->cnt3[lane + 32] = b;
			const uint32_t t3 = __reduce_add_sync(0xffffffffu, a + b);
			__syncwarp();
			uint32_t c2 = 0;
			if(lane < 16) {
				c2 = sm->cnt2[lane] + sm->cnt3[lane] + sm->cnt3[lane + 16] + sm->cnt3[lane + 32] + sm->cnt3[lane + 48];
				sm->cnt2[lane] = c2;
			}
			const uint32_t t2 = __reduce_add_sync(0xffffffffu, c2);
			__syncwarp();
			uint32_t c1 = 0;
			if(lane < 4) {
				c1 = sm->cnt1[lane] + sm->cnt2[lane] + sm->cnt2[lane + 4] + sm->cnt2[lane + 8] + sm->cnt2[lane + 12];
				sm->cnt1[lane] = c1;
			}
			const uint32_t t1 = __reduce_add_sync(0xffffffffu, c1);
			if(lane < 4) {
				const uint32_t tk = (lane == 0)? t1 : (lane == 1)? t2 : (lane == 2)? t3 : t4;
				sm->tot[lane] = tk;
				sm->inv_tot[lane] = tk? 1.0f / (float)tk : 0.0f;
			}
			__syncwarp();
		}
		// 180 canonical dimensions: dims[canon] += count/total for the mer and for its reverse complement (:171)
		for(int d = lane; d < ABW_NKMER_DIMS; d += 32) {
			const uint32_t ent = sm_dim_tab[d];
			const int k = ent & 0xFFu;
			const uint32_t* cnt = (k == 4)? sm->cnt4 : (k == 3)? sm->cnt3 : (k == 2)? sm->cnt2 : sm->cnt1;
			const uint32_t ca = (ent >> 8) & 0xFFu, cb = (ent >> 16) & 0xFFu;
			const uint32_t c1 = cnt[ca], c2 = (cb != ca)? cnt[cb] : 0u;
			const uint32_t t = sm->tot[k - 1];
			double out = 0.0;
			if(t != 0 && (c1 + c2) != 0) {
				bool exact_path = (KIND != ABW_FEAT_TRUNC3) || (c1 + c2) > 4000000u || t > (1u << 28);
				if(!exact_path) {
					// int(1000*(c1/t + c2/t)) equals floor(1000*(c1+c2)/t) whenever the quotient is not an integer: the rounding
					// error of the three fp64 operations (< 1e-12) is far below the distance 1/t to the next integer.  The quotient
					// (at most 1000) is formed in single precision, off by at most one, and verified exactly in integers.
					const uint32_t num = 1000u * (c1 + c2);
					uint32_t m = (uint32_t)__float2uint_rz((float)num * sm->inv_tot[k - 1]);
					int32_t r = (int32_t)(num - m * t);
					if(r < 0) { m--; r += (int32_t)t; }
					else if((uint32_t)r >= t) { m++; r -= (int32_t)t; }
					if(r != 0)
						out = g_milli[m];
					else
						exact_path = true;
				}
				if(exact_path) {
					double x = (c1? __ddiv_rn((double)c1, (double)(int)t) : 0.0);
					if(c2)
						x = __dadd_rn(x, __ddiv_rn((double)c2, (double)(int)t));
					out = (KIND == ABW_FEAT_TRUNC3)? __ddiv_rn((double)__double2int_rz(__dmul_rn(1000.0, x)), 1000.0) : x;
				}
			}
			if(!(skip_A && d == 0))
				rows[g * ld + col0 + d - (skip_A? 1 : 0)] = out;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------
// coverage, abawaca-build.cpp:177-185, 231-244, 546-551
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool read_accepted(const abw_read& r, uint32_t max_snps, uint32_t nscaf)
{
	uint32_t flag = r.flag_nsnps & 0xFFFFu, nsnps = r.flag_nsnps >> 16;
	return !((flag & 0x4u) || (nsnps > max_snps) || (flag & 0x100u)) && r.scaf < nscaf;   // :546-550
}

// Windows hit by the reads of a thread, staged so that the loads of its COV_ITEMS reads are in flight together.
//   first window g0 = the first one of the scaffold whose end is >= s (windows before it are skipped by `continue`, :235-236); windows of a scaffold hold the
//   same number of non-N bases, so it is window (s-1)/nbps unless a run of N shifted the boundaries (then: binary search).
//   Windows of a scaffold are contiguous (start[g] = end[g-1] + 1, start of the first = 1), so only the ends are read.
// Returns per read: g0, number of windows cnt, and the end of window g0 - 1 (0 for the first window of the scaffold) and of g0.
constexpr int COV_ITEMS = 4;                                // consecutive reads per thread
struct ReadHit { uint32_t g0, cnt; uint64_t end_prev, end_cur; };
__device__ __forceinline__ void read_windows(const abw_read (&rd)[COV_ITEMS], uint32_t max_snps, uint32_t nscaf, const uint4* __restrict__ scaf_info,
                                             const uint64_t* __restrict__ seg_end, ReadHit (&hit)[COV_ITEMS])
{
	bool acc[COV_ITEMS];
	uint4 si[COV_ITEMS];
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		acc[j] = read_accepted(rd[j], max_snps, nscaf);
		si[j] = acc[j]? __ldg(scaf_info + rd[j].scaf) : make_uint4(0, 0, 1, 0);
		acc[j] = acc[j] && si[j].y > 0;
	}
	uint32_t g[COV_ITEMS];
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		const uint32_t s = rd[j].pos0;
		uint32_t q = 0;
		if(s > 0) {
			// an all-N scaffold (nbps 0) has one window per character
			if(si[j].w == 0 || si[j].w == COV_REGULAR)
				q = (s - 1) / max(si[j].z, 1u);
			else
				q = 0;                                      // more than 2^32 bases per window: every read position lies in the first window or is found by the search
		}
		g[j] = si[j].x + min(si[j].y - (acc[j]? 1u : 0u), q);
	}
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		hit[j].end_cur = acc[j]? __ldg(seg_end + g[j]) : 0ull;
		hit[j].end_prev = (acc[j] && g[j] > si[j].x)? __ldg(seg_end + g[j] - 1) : 0ull;
	}
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		hit[j].g0 = 0; hit[j].cnt = 0;
		if(!acc[j])
			continue;
		const uint64_t s = rd[j].pos0, e = (uint64_t)rd[j].pos0 + rd[j].len - 1;
		const uint32_t f0 = si[j].x, f1 = si[j].x + si[j].y;
		uint32_t gg = g[j];
		if(!((gg == f0 || hit[j].end_prev < s) && hit[j].end_cur >= s)) {
			uint32_t lo = f0, hi = f1;                      // runs of N moved the boundaries: binary search
			while(lo < hi) {
				const uint32_t mid = lo + ((hi - lo) >> 1);
				if(__ldg(seg_end + mid) < s) lo = mid + 1; else hi = mid;
			}
			gg = lo;
			hit[j].end_cur = (gg < f1)? __ldg(seg_end + gg) : 0ull;
			hit[j].end_prev = (gg > f0)? __ldg(seg_end + gg - 1) : 0ull;
		}
		hit[j].g0 = gg;
		// windows gg, gg+1, ... while their start is <= e (:233-237)
		uint32_t c = 0;
		uint64_t prev_end = hit[j].end_prev, cur_end = hit[j].end_cur;
		uint32_t w = gg;
		while(w < f1 && !(e < prev_end + 1)) {                 // first window of a scaffold: prev_end is 0 and the start 1
			c++;
			w++;
			prev_end = cur_end;
			if(w < f1 && !(e < prev_end + 1))
				cur_end = __ldg(seg_end + w);
		}
		hit[j].cnt = c;
	}
}

constexpr int COV_THREADS = 256;
constexpr int COV_TILE = COV_THREADS * COV_ITEMS;

__device__ __forceinline__ void load_reads(const abw_read* __restrict__ reads, uint64_t nreads, uint64_t r0, abw_read (&rd)[COV_ITEMS])
{
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		if(r0 + j < nreads) {
			const uint4 x = __ldg(reinterpret_cast<const uint4*>(reads + r0 + j));
			rd[j].scaf = x.x; rd[j].pos0 = x.y; rd[j].len = x.z; rd[j].flag_nsnps = x.w;
		}
		else {
			rd[j].scaf = 0xFFFFFFFFu; rd[j].pos0 = 0; rd[j].len = 0; rd[j].flag_nsnps = 0;
		}
	}
}

// pass 1: (window, read) pairs per tile of COV_TILE reads; per-scaffold read bases of the -c sample
__global__ void __launch_bounds__(COV_THREADS) k_cov_count(const abw_read* __restrict__ reads, uint64_t nreads, uint32_t max_snps, uint32_t nscaf,
                                                          const uint4* __restrict__ scaf_info, const uint64_t* __restrict__ seg_end, uint32_t* __restrict__ tile_counts,
                                                          unsigned long long* __restrict__ scaf_nbps)
{
	__shared__ uint32_t sm[COV_THREADS / 32];
	const uint64_t r0 = (uint64_t)blockIdx.x * COV_TILE + (uint64_t)threadIdx.x * COV_ITEMS;
	abw_read rd[COV_ITEMS];
	load_reads(reads, nreads, r0, rd);
	ReadHit hit[COV_ITEMS];
	read_windows(rd, max_snps, nscaf, scaf_info, seg_end, hit);
	uint32_t c = 0;
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		c += hit[j].cnt;
		if(scaf_nbps != nullptr && read_accepted(rd[j], max_snps, nscaf))
			atomicAdd(&scaf_nbps[rd[j].scaf], (unsigned long long)rd[j].len);     // integer: order free (:242-243)
	}
	c = __reduce_add_sync(0xffffffffu, c);
	if((threadIdx.x & 31) == 0)
		sm[threadIdx.x >> 5] = c;
	__syncthreads();
	if(threadIdx.x == 0) {
		uint32_t t = 0;
#pragma unroll
		for(int w = 0; w < COV_THREADS / 32; w++)
			t += sm[w];
		tile_counts[blockIdx.x] = t;
	}
}

// pass 2: the pairs, in read order (then window order): key = window, value = overlap
__global__ void __launch_bounds__(COV_THREADS) k_cov_emit(const abw_read* __restrict__ reads, uint64_t nreads, uint32_t max_snps, uint32_t nscaf,
                                                         const uint4* __restrict__ scaf_info, const uint64_t* __restrict__ seg_end, const uint64_t* __restrict__ tile_offs,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
	__shared__ uint32_t sm[COV_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t r0 = (uint64_t)blockIdx.x * COV_TILE + (uint64_t)threadIdx.x * COV_ITEMS;
	abw_read rd[COV_ITEMS];
	load_reads(reads, nreads, r0, rd);
	ReadHit hit[COV_ITEMS];
	read_windows(rd, max_snps, nscaf, scaf_info, seg_end, hit);
	uint32_t c = 0;
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++)
		c += hit[j].cnt;
	uint32_t incl = c;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if(lane >= o)
			incl += t;
	}
	if(lane == 31)
		sm[warp] = incl;
	__syncthreads();
	uint32_t wex = 0;
#pragma unroll
	for(int w = 0; w < COV_THREADS / 32; w++)
		if(w < warp)
			wex += sm[w];
	uint64_t o = tile_offs[blockIdx.x] + wex + incl - c;
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		const uint64_t s = rd[j].pos0, e = (uint64_t)rd[j].pos0 + rd[j].len - 1;
		uint64_t prev_end = hit[j].end_prev, cur_end = hit[j].end_cur;
		for(uint32_t k = 0; k < hit[j].cnt; k++) {
			const uint32_t g = hit[j].g0 + k;
			const uint64_t st = prev_end + 1, en = cur_end;             // first window of a scaffold: prev_end is 0 and the start 1
			const uint64_t s2 = (s < st)? st : s, e2 = (e > en)? en : e;   // 0-based read against 1-based window: quirk Q3, kept (:182-183)
			keys[o] = g;
			vals[o] = (uint32_t)(e2 - s2 + 1);                             // the overlap travels with the pair: no gather after the sort
			o++;
			prev_end = cur_end;
			if(k + 1 < hit[j].cnt)
				cur_end = __ldg(seg_end + g + 1);
		}
	}
}

// run of every window in the sorted pairs: [run[g].x, run[g].y), both zero (pre-cleared) for a window no read touches
__global__ void k_cov_runs(const uint32_t* __restrict__ keys, uint64_t npairs, uint2* __restrict__ run)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= npairs)
		return;
	const uint32_t k = keys[i];
	if(i == 0 || keys[i - 1] != k)
		run[k].x = (uint32_t)i;
	if(i + 1 == npairs || keys[i + 1] != k)
		run[k].y = (uint32_t)(i + 1);
}

// one thread per window: the reads that hit it, in SAM order (the pairs were sorted stably by window)
// `only` (may be null): windows whose flag is zero already hold their value (k_cov_quotient) and are left alone
template <int KIND, int BATCH>
__global__ void k_cov_accumulate(const uint32_t* __restrict__ vals, const uint2* __restrict__ run, uint64_t nseg,
                                 const uint64_t* __restrict__ seg_start, const uint64_t* __restrict__ seg_end, double* __restrict__ rows, uint64_t ld, uint32_t col,
                                 const uint32_t* __restrict__ only)
{
	uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= nseg)
		return;
	if(only != nullptr && ((only[g >> 5] >> (g & 31u)) & 1u) == 0)
		return;
	const double seglen = (double)(seg_end[g] - seg_start[g] + 1);
	double acc = 0.0, q = 0.0;
	uint32_t last = 0xFFFFFFFFu;
	const uint2 rg = run[g];
	for(uint32_t i = rg.x; i < rg.y; i += BATCH) {
		// BATCH loads in flight (8 when every window is summed, 32 when only the few flagged ones are and the longest run sets the time);
		// the additions stay strictly in SAM order (quirk Q5)
		const uint32_t m = min((uint32_t)BATCH, rg.y - i);
		uint32_t ov[BATCH];
#pragma unroll
		for(int j = 0; j < BATCH; j++)
			ov[j] = ((uint32_t)j < m)? __ldg(vals + i + j) : 0u;
#pragma unroll
		for(int j = 0; j < BATCH; j++) {
			if((uint32_t)j < m) {
				if(ov[j] != last) {                            // most reads lie entirely inside the window: same quotient, computed once
					q = __ddiv_rn((double)ov[j], seglen);      // :184
					last = ov[j];
				}
				acc = __dadd_rn(acc, q);
			}
		}
	}
	if(KIND == ABW_FEAT_TRUNC3)
		acc = __ddiv_rn((double)__double2int_rz(__dmul_rn(1000.0, acc)), 1000.0);
	rows[g * ld + col] = acc;
}

// ---------------------------------------------------------------------------------------------------
// Coverage written with three decimals (ABW_FEAT_TRUNC3, what abawaca-build prints, abawaca-build.cpp:578-607) without sorting the reads.
//
// The reference adds fl(ov_i / len) in SAM order and prints int(1000 * sum) / 1000.0.  With A = sum of the overlaps (an integer, order free)
// the exact value of 1000 * sum is 1000 A / len = m + r / len (m, r integers, 0 <= r < len).  The floating-point sum of n <= A terms differs
// from A / len by at most (n + 1) u A / len (u = 2^-53: one rounding per quotient, one per addition, one for the product with 1000), so the
// product the reference truncates lies within E = 1000 (A + 1) u A / len (1 + 1e-10) of m + r / len.  Whenever r != 0 and
// E < min(r, len - r) / len the truncation is m whatever the order of the reads was: those windows are finished by k_cov_quotient from the
// integer sum alone (the test is made with twice that bound).  The others (r == 0: the exact value IS a multiple of 0.001 and the rounding
// direction decides, a fraction of about gcd(1000, len) / len of the windows; or a huge sum) are flagged, and only the reads that touch a
// flagged window go through the stable sort and the in-order accumulation below.  ABW_FEAT_RAW always takes the sorted path.
// ---------------------------------------------------------------------------------------------------
// pass 1: integer sum of the overlaps per window (order free), first window and window count of every read for the second pass,
// per-scaffold read bases of the -c sample.  status[0] |= 1 when a read touches more than 255 windows (the caller then sorts everything).
// A scaffold without N (COV_REGULAR) has the windows [l nbps + 1, (l + 1) nbps]: its reads need the 16-byte scaffold record and nothing
// else.  The other scaffolds go through the window table one read at a time (guess, verify, binary search: as read_windows).
__device__ __noinline__ void cov_sum_irregular(uint32_t pos0, uint32_t rlen, const uint4 si, const uint64_t* __restrict__ seg_end,
                                               unsigned long long* __restrict__ sum_ov, uint32_t& g0_out, uint32_t& cnt_out)
{
	const uint64_t s = pos0, e = (uint64_t)pos0 + rlen - 1;
	const uint32_t f0 = si.x, f1 = si.x + si.y;
	uint32_t q = 0;
	if(pos0 > 0 && si.w == 0)
		q = (pos0 - 1) / max(si.z, 1u);                     // an all-N scaffold (nbps 0) has one window per character
	uint32_t gg = f0 + min(si.y - 1u, q);
	uint64_t cur_end = __ldg(seg_end + gg), prev_end = (gg > f0)? __ldg(seg_end + gg - 1) : 0ull;
	if(!((gg == f0 || prev_end < s) && cur_end >= s)) {
		uint32_t lo = f0, hi = f1;                          // runs of N moved the boundaries: binary search for the first window whose end is >= s
		while(lo < hi) {
			const uint32_t mid = lo + ((hi - lo) >> 1);
			if(__ldg(seg_end + mid) < s) lo = mid + 1; else hi = mid;
		}
		gg = lo;
		cur_end = (gg < f1)? __ldg(seg_end + gg) : 0ull;
		prev_end = (gg > f0)? __ldg(seg_end + gg - 1) : 0ull;
	}
	uint32_t c = 0, w = gg;
	while(w < f1 && !(e < prev_end + 1)) {                     // windows gg, gg+1, ... while their start is <= e (:233-237)
		const uint64_t st = prev_end + 1, en = cur_end;
		const uint64_t s2 = (s < st)? st : s, e2 = (e > en)? en : e;   // 0-based read against 1-based window: quirk Q3, kept (:182-183)
		atomicAdd(&sum_ov[w], (unsigned long long)(e2 - s2 + 1));
		c++;
		w++;
		prev_end = cur_end;
		if(w < f1)
			cur_end = __ldg(seg_end + w);
	}
	g0_out = gg;
	cnt_out = c;
}

__global__ void __launch_bounds__(COV_THREADS, 5) k_cov_sum(const abw_read* __restrict__ reads, uint64_t nreads, uint32_t max_snps, uint32_t nscaf,
                                                           const uint4* __restrict__ scaf_info, const uint64_t* __restrict__ seg_end,
                                                           unsigned long long* __restrict__ sum_ov, uint4* __restrict__ hit_g0, uchar4* __restrict__ hit_cnt,
                                                           unsigned long long* __restrict__ scaf_nbps, uint32_t* __restrict__ status)
{
	static_assert(COV_ITEMS == 4, "hit records are stored four at a time");
	const uint64_t t = (uint64_t)blockIdx.x * COV_THREADS + threadIdx.x;
	const uint64_t r0 = t * COV_ITEMS;
	uint4 x[COV_ITEMS], si[COV_ITEMS];                        // read: scaf, pos0, len, flag_nsnps; scaffold: first window, windows, nbps, kind
	bool acc[COV_ITEMS];
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++)
		x[j] = (r0 + j < nreads)? __ldg(reinterpret_cast<const uint4*>(reads + r0 + j)) : make_uint4(0xFFFFFFFFu, 0, 0, 0);
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		const uint32_t flag = x[j].w & 0xFFFFu, nsnps = x[j].w >> 16;
		acc[j] = !((flag & 0x4u) || (nsnps > max_snps) || (flag & 0x100u)) && x[j].x < nscaf;   // :546-550
		si[j] = acc[j]? __ldg(scaf_info + x[j].x) : make_uint4(0, 0, 1, 0);
	}
	uint32_t g0[COV_ITEMS], cn[COV_ITEMS];
#pragma unroll
	for(int j = 0; j < COV_ITEMS; j++) {
		g0[j] = 0;
		cn[j] = 0;
		if(!acc[j])
			continue;
		if(scaf_nbps != nullptr)
			atomicAdd(&scaf_nbps[x[j].x], (unsigned long long)x[j].z);        // integer: order free (:242-243)
		if(si[j].y == 0)
			continue;
		if(si[j].w == COV_REGULAR) {
			const uint32_t nb = si[j].z, pos0 = x[j].y;
			const uint32_t q = pos0? (pos0 - 1) / nb : 0u;     // position s >= 1 lies in window (s - 1) / nbps; position 0 before the first one
			if(q >= si[j].y)
				continue;                                   // behind the last window (the trailing bases no window holds)
			const uint64_t s = pos0, e = (uint64_t)pos0 + x[j].z - 1;
			uint32_t l = q, c = 0;
			uint64_t st = (uint64_t)q * nb + 1;
			while(l < si[j].y && !(e < st)) {
				const uint64_t en = st + nb - 1;
				const uint64_t s2 = (s < st)? st : s, e2 = (e > en)? en : e;   // quirk Q3, as above
				atomicAdd(&sum_ov[si[j].x + l], (unsigned long long)(e2 - s2 + 1));
				c++;
				l++;
				st = en + 1;
			}
			g0[j] = si[j].x + q;
			cn[j] = c;
		}
		else
			cov_sum_irregular(x[j].y, x[j].z, si[j], seg_end, sum_ov, g0[j], cn[j]);
		if(cn[j] > 255u) {
			atomicOr(status, 1u);
			cn[j] = 255u;
		}
	}
	hit_g0[t] = make_uint4(g0[0], g0[1], g0[2], g0[3]);
	hit_cnt[t] = make_uchar4((unsigned char)cn[0], (unsigned char)cn[1], (unsigned char)cn[2], (unsigned char)cn[3]);
}

// one thread per window: the value from the integer sum where the truncation cannot depend on the order of the reads, a flag elsewhere
__global__ void __launch_bounds__(256) k_cov_quotient(const unsigned long long* __restrict__ sum_ov, uint64_t nseg, const uint64_t* __restrict__ seg_start,
                                                      const uint64_t* __restrict__ seg_end, double* __restrict__ rows, uint64_t ld, uint32_t col,
                                                      uint32_t* __restrict__ flag_bits)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool f = false;
	if(g < nseg) {
		const uint64_t A = sum_ov[g], len = seg_end[g] - seg_start[g] + 1;
		if(A == 0)
			rows[g * ld + col] = 0.0;                          // no read: the sum is 0.0 and so is int(0.0) / 1000.0
		else {
			f = true;
			if(A < (1ull << 40)) {
				const uint64_t num = 1000ull * A, m = num / len, r = num - m * len;
				if(r != 0 && m < (1ull << 31)) {
					const uint64_t dist = (r < len - r)? r : len - r;
					const double bound = 2000.0 * (double)(A + 2) * (double)A * 1.1102230246251565e-16;   // 2 x 1000 (A + 1) u A, in units of 1 / len
					if(bound < (double)dist) {
						f = false;
						rows[g * ld + col] = __ddiv_rn((double)(int)m, 1000.0);
					}
				}
			}
		}
	}
	// one bit per window (28 KB for 227 k windows: the second pass tests it for every read out of L1)
	const uint32_t bits = __ballot_sync(0xffffffffu, f);
	if((threadIdx.x & 31) == 0 && g < nseg)
		flag_bits[g >> 5] = bits;
}

__device__ __forceinline__ uint32_t window_flagged(const uint32_t* __restrict__ flag_bits, uint32_t g)
{
	return (__ldg(flag_bits + (g >> 5)) >> (g & 31u)) & 1u;
}

// Second pass over the hit records: a thread takes FL_GROUPS groups of four consecutive reads, a CTA 256 * FL_GROUPS * 4 consecutive reads;
// pairs come out in read order.  The flag of the first window of every read is tested unconditionally (independent loads, all in flight);
// the few reads over several windows loop.  (Measured: these kernels are bound by the chain load record -> load flag of each thread, and
// four groups per thread were slower than one.)
constexpr int FL_GROUPS = 1;
constexpr int FL_TILE_GROUPS = COV_THREADS * FL_GROUPS;

__device__ __forceinline__ uint32_t flagged_hits(const uint4* __restrict__ hit_g0, const uchar4* __restrict__ hit_cnt, uint64_t ngroups, uint64_t t0,
                                                const uint32_t* __restrict__ flag, uint4 (&g4)[FL_GROUPS], uchar4 (&c4)[FL_GROUPS])
{
#pragma unroll
	for(int q = 0; q < FL_GROUPS; q++) {
		const bool in = t0 + q < ngroups;
		g4[q] = in? __ldg(hit_g0 + t0 + q) : make_uint4(0, 0, 0, 0);
		c4[q] = in? __ldg(hit_cnt + t0 + q) : make_uchar4(0, 0, 0, 0);
	}
	uint32_t c = 0;
#pragma unroll
	for(int q = 0; q < FL_GROUPS; q++) {
		const uint32_t fx = window_flagged(flag, g4[q].x), fy = window_flagged(flag, g4[q].y);      // window 0 for a read without window: any valid word
		const uint32_t fz = window_flagged(flag, g4[q].z), fw = window_flagged(flag, g4[q].w);
		c += (c4[q].x? fx : 0u) + (c4[q].y? fy : 0u) + (c4[q].z? fz : 0u) + (c4[q].w? fw : 0u);
		for(uint32_t k = 1; k < c4[q].x; k++) c += window_flagged(flag, g4[q].x + k);
		for(uint32_t k = 1; k < c4[q].y; k++) c += window_flagged(flag, g4[q].y + k);
		for(uint32_t k = 1; k < c4[q].z; k++) c += window_flagged(flag, g4[q].z + k);
		for(uint32_t k = 1; k < c4[q].w; k++) c += window_flagged(flag, g4[q].w + k);
	}
	return c;
}

// pass 2a: (flagged window, read) pairs per tile
__global__ void __launch_bounds__(COV_THREADS) k_cov_count_flagged(const uint4* __restrict__ hit_g0, const uchar4* __restrict__ hit_cnt, uint64_t ngroups,
                                                                  const uint32_t* __restrict__ flag, uint32_t* __restrict__ tile_counts)
{
	__shared__ uint32_t sm[COV_THREADS / 32];
	const uint64_t t0 = ((uint64_t)blockIdx.x * COV_THREADS + threadIdx.x) * FL_GROUPS;
	uint4 g4[FL_GROUPS];
	uchar4 c4[FL_GROUPS];
	uint32_t c = flagged_hits(hit_g0, hit_cnt, ngroups, t0, flag, g4, c4);
	c = __reduce_add_sync(0xffffffffu, c);
	if((threadIdx.x & 31) == 0)
		sm[threadIdx.x >> 5] = c;
	__syncthreads();
	if(threadIdx.x == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < COV_THREADS / 32; w++)
			tot += sm[w];
		tile_counts[blockIdx.x] = tot;
	}
}

// pass 2b: those pairs in read order (then window order): key = window, value = overlap
__global__ void __launch_bounds__(COV_THREADS) k_cov_emit_flagged(const abw_read* __restrict__ reads, const uint4* __restrict__ hit_g0, const uchar4* __restrict__ hit_cnt,
                                                                 uint64_t ngroups, const uint32_t* __restrict__ flag, const uint64_t* __restrict__ seg_start,
                                                                 const uint64_t* __restrict__ seg_end, const uint64_t* __restrict__ tile_offs,
                                                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
	__shared__ uint32_t sm[COV_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t t0 = ((uint64_t)blockIdx.x * COV_THREADS + threadIdx.x) * FL_GROUPS;
	uint4 g4[FL_GROUPS];
	uchar4 c4[FL_GROUPS];
	const uint32_t c = flagged_hits(hit_g0, hit_cnt, ngroups, t0, flag, g4, c4);
	uint32_t incl = c;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
		if(lane >= o)
			incl += x;
	}
	if(lane == 31)
		sm[warp] = incl;
	__syncthreads();
	if(c == 0)
		return;
	uint32_t wex = 0;
#pragma unroll
	for(int w = 0; w < COV_THREADS / 32; w++)
		if(w < warp)
			wex += sm[w];
	uint64_t o = tile_offs[blockIdx.x] + wex + incl - c;
#pragma unroll
	for(int q = 0; q < FL_GROUPS; q++) {
		const uint32_t g0[4] = {g4[q].x, g4[q].y, g4[q].z, g4[q].w}, cn[4] = {c4[q].x, c4[q].y, c4[q].z, c4[q].w};
#pragma unroll
		for(int j = 0; j < 4; j++) {
			for(uint32_t k = 0; k < cn[j]; k++) {
				const uint32_t g = g0[j] + k;
				if(window_flagged(flag, g) == 0)
					continue;
				const uint4 x = __ldg(reinterpret_cast<const uint4*>(reads + (t0 + q) * COV_ITEMS + j));   // scaf, pos0, len, flag_nsnps
				const uint64_t s = x.y, e = (uint64_t)x.y + x.z - 1;
				const uint64_t st = __ldg(seg_start + g), en = __ldg(seg_end + g);
				const uint64_t s2 = (s < st)? st : s, e2 = (e > en)? en : e;   // as in k_cov_emit
				keys[o] = g;
				vals[o] = (uint32_t)(e2 - s2 + 1);
				o++;
			}
		}
	}
}

bool g_tables_ready[64] = {};

int upload_tables(abw_ctx* ctx)
{
	if(ctx->device >= 0 && ctx->device < 64 && g_tables_ready[ctx->device])
		return ABW_OK;
	uint8_t dk[ABW_NKMER_DIMS];
	uint16_t da[ABW_NKMER_DIMS], db[ABW_NKMER_DIMS];
	int next = 0;
	int dim_of[5][256];
	for(int k = 1; k <= 4; k++) {
		int n = 1 << (2 * k);
		for(int code = 0; code < n; code++) {              // big-endian code = lexicographic order, abawaca-build.cpp:90-99
			int rc = 0, c = code;
			for(int i = 0; i < k; i++) { rc = (rc << 2) | (3 - (c & 3)); c >>= 2; }
			auto to_le = [k](int be) { int le = 0; for(int i = 0; i < k; i++) { le |= ((be >> (2 * (k - 1 - i))) & 3) << (2 * i); } return le; };
			if(rc < code)
				dim_of[k][code] = dim_of[k][rc];
			else {
				dim_of[k][code] = next;
				dk[next] = (uint8_t)k;
				da[next] = (uint16_t)to_le(code);
				db[next] = (uint16_t)to_le(rc);
				next++;
			}
		}
	}
	if(next != ABW_NKMER_DIMS)
		return abw_fail(ctx, ABW_ERR_ARG, "internal: canonical k-mer table size");
	double milli[1001];
	for(int m = 0; m <= 1000; m++)
		milli[m] = (double)m / 1000.0;
	uint32_t tab[ABW_NKMER_DIMS];
	for(int i = 0; i < ABW_NKMER_DIMS; i++)
		tab[i] = (uint32_t)dk[i] | ((uint32_t)da[i] << 8) | ((uint32_t)db[i] << 16);
	ABW_CUDA(ctx, cudaMemcpyToSymbol(g_dim_tab, tab, sizeof(tab)));
	ABW_CUDA(ctx, cudaMemcpyToSymbol(g_milli, milli, sizeof(milli)));
	if(ctx->device >= 0 && ctx->device < 64)
		g_tables_ready[ctx->device] = true;
	return ABW_OK;
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" {

static int pack_impl(abw_ctx* ctx, abw_seqset* s, const char* ascii, int ascii_on_device, const uint64_t* h_offsets, uint32_t nscaf)
{
	s->nscaf = nscaf;
	s->h_len.resize(nscaf);
	s->h_base.resize((size_t)nscaf + 1);
	uint64_t b = 0;
	for(uint32_t i = 0; i < nscaf; i++) {
		if(h_offsets[i + 1] < h_offsets[i])
			return abw_fail(ctx, ABW_ERR_ARG, "abw_pack_sequences: offsets must be non-decreasing");
		s->h_len[i] = h_offsets[i + 1] - h_offsets[i];
		s->h_base[i] = b;
		b += (s->h_len[i] + 127) & ~127ull;
	}
	s->h_base[nscaf] = b;
	s->total_padded = b;
	const uint64_t bytes = h_offsets[nscaf];
	DevBuf<unsigned char> d_ascii;
	DevBuf<uint64_t> d_off;
	DevBuf<int> d_err;
	const unsigned char* src = (const unsigned char*)ascii;
	uint64_t readable = bytes;
	if(!ascii_on_device) {
		ABW_CUDA(ctx, d_ascii.alloc(bytes + 64));
		ABW_CUDA(ctx, cudaMemcpyAsync(d_ascii.p, ascii, bytes, cudaMemcpyHostToDevice, ctx->stream));
		src = d_ascii.p;
		readable = bytes + 64;
	}
	if(abw_small_copies_by_kernel())
		abw_stage_flip(ctx);                                   // this call waits for the stream before it returns: the half used two flips ago is free
	ABW_CUDA(ctx, d_off.alloc((size_t)nscaf + 1));
	ABW_CUDA(ctx, abw_upload_small(ctx, d_off.p, h_offsets, sizeof(uint64_t) * ((size_t)nscaf + 1)));
	ABW_CUDA(ctx, s->len.alloc(nscaf));
	ABW_CUDA(ctx, s->base.alloc((size_t)nscaf + 1));
	ABW_CUDA(ctx, abw_upload_small(ctx, s->len.p, s->h_len.data(), sizeof(uint64_t) * nscaf));
	ABW_CUDA(ctx, abw_upload_small(ctx, s->base.p, s->h_base.data(), sizeof(uint64_t) * ((size_t)nscaf + 1)));
	// + slack so that the k-mer kernel may read one word past a segment
	ABW_CUDA(ctx, s->packed.alloc(b / 16 + 16));
	ABW_CUDA(ctx, s->valid.alloc(b / 32 + 16));
	ABW_CUDA(ctx, s->nmask.alloc(b / 32 + 16));
	ABW_CUDA(ctx, cudaMemsetAsync(s->packed.p + b / 16, 0, 16 * sizeof(uint32_t), ctx->stream));
	ABW_CUDA(ctx, cudaMemsetAsync(s->valid.p + b / 32, 0, 16 * sizeof(uint32_t), ctx->stream));
	ABW_CUDA(ctx, cudaMemsetAsync(s->nmask.p + b / 32, 0, 16 * sizeof(uint32_t), ctx->stream));
	ABW_CUDA(ctx, s->countN.alloc(nscaf));
	ABW_CUDA(ctx, s->countGC.alloc(nscaf));
	ABW_CUDA(ctx, d_err.alloc(1));
	ABW_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
	if(nscaf > 0) {
		unsigned int blocks = (unsigned int)std::min<uint64_t>(abw_div_up((uint64_t)nscaf * 32, 256), (uint64_t)ctx->sm_count * 16);
		ABW_LAUNCH(ctx, k_pack, blocks, 256, 0, src, readable, d_off.p, s->base.p, nscaf, s->packed.p, s->valid.p, s->nmask.p, s->countN.p, s->countGC.p, d_err.p);
	}
	int h_err = 0;
	ABW_CUDA(ctx, abw_fetch(ctx, &h_err, d_err.p, sizeof(int)));
	ABW_CUDA(ctx, abw_sync(ctx));
	if(h_err)
		return abw_fail(ctx, ABW_ERR_ILLEGAL_DNA, "Illegal_DNAString: lower-case 'n' in a sequence (String.cpp:47-49)");
	return ABW_OK;
}

int abw_pack_sequences(abw_ctx* ctx, const char* ascii, int ascii_on_device, const uint64_t* h_offsets, uint32_t nscaf, abw_seqset** out)
{
	if(!ctx || !out || !h_offsets || (!ascii && nscaf && h_offsets[nscaf] > 0))
		return abw_fail(ctx, ABW_ERR_ARG, "abw_pack_sequences: null argument");
	ABW_ENTER(ctx);
	abw_seqset* s = new abw_seqset();
	int rc = pack_impl(ctx, s, ascii, ascii_on_device, h_offsets, nscaf);
	if(rc != ABW_OK) {
		delete s;
		return rc;
	}
	*out = s;
	return ABW_OK;
}

void abw_seqset_destroy(abw_seqset* s) { delete s; }

int abw_seqset_stats(abw_ctx* ctx, const abw_seqset* s, uint64_t* h_count_N, uint64_t* h_count_GC)
{
	if(!ctx || !s)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_seqset_stats: null argument");
	if(h_count_N)
		ABW_CUDA(ctx, abw_fetch(ctx, h_count_N, s->countN.p, sizeof(uint64_t) * s->nscaf));
	if(h_count_GC)
		ABW_CUDA(ctx, abw_fetch(ctx, h_count_GC, s->countGC.p, sizeof(uint64_t) * s->nscaf));
	ABW_CUDA(ctx, abw_sync(ctx));
	return ABW_OK;
}

int abw_segment(abw_ctx* ctx, const abw_seqset* s, uint32_t window_size, abw_segments** out)
{
	if(!ctx || !s || !out || window_size == 0)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_segment: bad argument");
	ABW_ENTER(ctx);
	abw_segments* g = new abw_segments();
	g->nscaf = s->nscaf;
	int rc = [&]() -> int {
		ABW_CUDA(ctx, g->seg_first.alloc((size_t)s->nscaf + 1));
		DevBuf<uint64_t> counts, total;
		ABW_CUDA(ctx, counts.alloc(s->nscaf));
		ABW_CUDA(ctx, total.alloc(1));
		if(s->nscaf > 0)
			ABW_LAUNCH(ctx, k_seg_count, abw_div_up(s->nscaf, 256), 256, 0, s->len.p, s->countN.p, s->nscaf, (uint64_t)window_size, counts.p);
		ABW_CHECK(abw_exclusive_scan_u64(ctx, counts.p, g->seg_first.p, s->nscaf, total.p));
		ABW_CUDA(ctx, abw_fetch(ctx, &g->nseg, total.p, sizeof(uint64_t)));
		ABW_CUDA(ctx, abw_sync(ctx));
		ABW_CUDA(ctx, abw_upload_small(ctx, g->seg_first.p + s->nscaf, &g->nseg, sizeof(uint64_t)));
		ABW_CUDA(ctx, g->seg_scaf.alloc(g->nseg));
		ABW_CUDA(ctx, g->seg_start.alloc(g->nseg));
		ABW_CUDA(ctx, g->seg_end.alloc(g->nseg));
		ABW_CUDA(ctx, g->seg_nonN.alloc(g->nseg));
		ABW_CUDA(ctx, g->seg_gbase.alloc(g->nseg));
		ABW_CUDA(ctx, g->scaf_info.alloc(s->nscaf));
		ABW_CUDA(ctx, cudaMemsetAsync(g->scaf_info.p, 0, sizeof(uint4) * s->nscaf, ctx->stream));
		if(s->nscaf > 0 && g->nseg > 0) {
			unsigned int blocks = (unsigned int)std::min<uint64_t>(abw_div_up((uint64_t)s->nscaf * 32, 256), (uint64_t)ctx->sm_count * 16);
			ABW_LAUNCH(ctx, k_seg_fill, blocks, 256, 0, s->len.p, s->countN.p, s->base.p, s->nmask.p, s->nscaf, (uint64_t)window_size, g->seg_first.p,
			           g->seg_scaf.p, g->seg_start.p, g->seg_end.p, g->seg_nonN.p, g->seg_gbase.p, g->scaf_info.p);
		}
		ABW_CUDA(ctx, abw_sync(ctx));
		return ABW_OK;
	}();
	if(rc != ABW_OK) {
		delete g;
		return rc;
	}
	*out = g;
	return ABW_OK;
}

void abw_segments_destroy(abw_segments* g) { delete g; }

uint64_t abw_segments_count(const abw_segments* g) { return g? g->nseg : 0; }

int abw_segments_get(abw_ctx* ctx, const abw_segments* g, uint64_t* h_seg_first, uint32_t* h_seg_scaf, uint64_t* h_seg_start, uint64_t* h_seg_end, uint64_t* h_seg_nonN)
{
	if(!ctx || !g)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_segments_get: null argument");
	if(h_seg_first)
		ABW_CUDA(ctx, abw_fetch(ctx, h_seg_first, g->seg_first.p, sizeof(uint64_t) * ((size_t)g->nscaf + 1)));
	if(h_seg_scaf)
		ABW_CUDA(ctx, abw_fetch(ctx, h_seg_scaf, g->seg_scaf.p, sizeof(uint32_t) * g->nseg));
	if(h_seg_start)
		ABW_CUDA(ctx, abw_fetch(ctx, h_seg_start, g->seg_start.p, sizeof(uint64_t) * g->nseg));
	if(h_seg_end)
		ABW_CUDA(ctx, abw_fetch(ctx, h_seg_end, g->seg_end.p, sizeof(uint64_t) * g->nseg));
	if(h_seg_nonN)
		ABW_CUDA(ctx, abw_fetch(ctx, h_seg_nonN, g->seg_nonN.p, sizeof(uint64_t) * g->nseg));
	ABW_CUDA(ctx, abw_sync(ctx));
	return ABW_OK;
}

int abw_kmer_features(abw_ctx* ctx, const abw_seqset* s, const abw_segments* g, int kind, int skip_A, double* d_rows, uint64_t ld, uint32_t col0)
{
	if(!ctx || !s || !g || !d_rows)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_kmer_features: null argument");
	if(kind != ABW_FEAT_TRUNC3 && kind != ABW_FEAT_RAW)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_kmer_features: unknown kind");
	if(ld < (uint64_t)col0 + ABW_NKMER_DIMS - (skip_A? 1 : 0))
		return abw_fail(ctx, ABW_ERR_ARG, "abw_kmer_features: row stride too small");
	ABW_ENTER(ctx);
	ABW_CHECK(upload_tables(ctx));
	if(g->nseg == 0)
		return ABW_OK;
	unsigned int blocks = (unsigned int)std::min<uint64_t>(abw_div_up(g->nseg, KM_WARPS), (uint64_t)ctx->sm_count * 5 * 4);
	if(kind == ABW_FEAT_TRUNC3)
		ABW_LAUNCH(ctx, k_kmer<ABW_FEAT_TRUNC3>, blocks, KM_WARPS * 32, 0, s->packed.p, s->valid.p, g->seg_gbase.p, g->seg_start.p, g->seg_end.p, g->nseg, skip_A, d_rows, ld, col0);
	else
		ABW_LAUNCH(ctx, k_kmer<ABW_FEAT_RAW>, blocks, KM_WARPS * 32, 0, s->packed.p, s->valid.p, g->seg_gbase.p, g->seg_start.p, g->seg_end.p, g->nseg, skip_A, d_rows, ld, col0);
	return ABW_OK;
}

// every (window, read) pair through the stable sort and the in-order accumulation; with `only` just the flagged windows, whose pairs
// are found from the hit records of k_cov_sum
static int coverage_sorted(abw_ctx* ctx, const abw_segments* g, const abw_read* rd, uint64_t nreads, uint32_t max_snps, int kind, double* d_rows, uint64_t ld,
                           uint32_t col, uint64_t* d_scaf_nbps, const uint32_t* only, const uint4* hit_g0, const uchar4* hit_cnt, const uint32_t* d_status,
                           bool* too_many_windows)
{
	DevBuf<uint32_t> tile_counts, keys, keys_tmp, vals, vals_tmp;
	DevBuf<uint64_t> tile_offs, total;
	DevBuf<uint2> run;
	const uint64_t ngroups = (uint64_t)abw_div_up(nreads, COV_TILE) * COV_THREADS;   // hit records (groups of four reads) written by k_cov_sum
	const unsigned int ntiles = only? abw_div_up(ngroups, FL_TILE_GROUPS) : abw_div_up(nreads, COV_TILE);
	ABW_CUDA(ctx, tile_counts.alloc(ntiles));
	ABW_CUDA(ctx, tile_offs.alloc(ntiles));
	ABW_CUDA(ctx, total.alloc(1));
	ABW_CUDA(ctx, run.alloc(g->nseg));
	ABW_CUDA(ctx, cudaMemsetAsync(run.p, 0, sizeof(uint2) * g->nseg, ctx->stream));
	uint64_t npairs = 0;
	if(nreads) {
		if(only)
			ABW_LAUNCH(ctx, k_cov_count_flagged, ntiles, COV_THREADS, 0, hit_g0, hit_cnt, ngroups, only, tile_counts.p);
		else
			ABW_LAUNCH(ctx, k_cov_count, ntiles, COV_THREADS, 0, rd, nreads, max_snps, g->nscaf, g->scaf_info.p, g->seg_end.p, tile_counts.p,
			           (unsigned long long*)d_scaf_nbps);
		ABW_CHECK(abw_exclusive_scan_u32_to_u64(ctx, tile_counts.p, tile_offs.p, ntiles, total.p));
		uint32_t status = 0;
		ABW_CUDA(ctx, abw_fetch(ctx, &npairs, total.p, sizeof(uint64_t)));
		if(d_status)
			ABW_CUDA(ctx, abw_fetch(ctx, &status, d_status, sizeof(uint32_t)));
		ABW_CUDA(ctx, abw_sync(ctx));
		if(status & 1u) {
			*too_many_windows = true;
			return ABW_OK;
		}
	}
	if(npairs >= (1ull << 32))
		return abw_fail(ctx, ABW_ERR_UNSUPPORTED, "abw_coverage: more than 2^32-1 (window, read) pairs per call; split the sample");
	ABW_CUDA(ctx, keys.alloc(npairs));
	ABW_CUDA(ctx, keys_tmp.alloc(npairs));
	ABW_CUDA(ctx, vals.alloc(npairs));
	ABW_CUDA(ctx, vals_tmp.alloc(npairs));
	if(npairs) {
		if(only)
			ABW_LAUNCH(ctx, k_cov_emit_flagged, ntiles, COV_THREADS, 0, rd, hit_g0, hit_cnt, ngroups, only, g->seg_start.p, g->seg_end.p, tile_offs.p, keys.p, vals.p);
		else
			ABW_LAUNCH(ctx, k_cov_emit, ntiles, COV_THREADS, 0, rd, nreads, max_snps, g->nscaf, g->scaf_info.p, g->seg_end.p, tile_offs.p, keys.p, vals.p);
		int nbits = 1;
		while(nbits < 32 && (1ull << nbits) < g->nseg)
			nbits++;
		// window ids are dense: every bit below nbits varies, no need to look
		ABW_CHECK(abw_radix_sort_pairs_u32(ctx, keys.p, keys_tmp.p, vals.p, vals_tmp.p, npairs, 1, npairs, -nbits));
		ABW_LAUNCH(ctx, k_cov_runs, abw_div_up(npairs, 256), 256, 0, keys.p, npairs, run.p);
	}
	if(g->nseg && (npairs || !only)) {
		if(kind == ABW_FEAT_TRUNC3 && only)
			ABW_LAUNCH(ctx, (k_cov_accumulate<ABW_FEAT_TRUNC3, 32>), abw_div_up(g->nseg, 64), 64, 0, vals.p, run.p, g->nseg, g->seg_start.p, g->seg_end.p, d_rows, ld, col, only);
		else if(kind == ABW_FEAT_TRUNC3)
			ABW_LAUNCH(ctx, (k_cov_accumulate<ABW_FEAT_TRUNC3, 8>), abw_div_up(g->nseg, 64), 64, 0, vals.p, run.p, g->nseg, g->seg_start.p, g->seg_end.p, d_rows, ld, col, only);
		else
			ABW_LAUNCH(ctx, (k_cov_accumulate<ABW_FEAT_RAW, 8>), abw_div_up(g->nseg, 64), 64, 0, vals.p, run.p, g->nseg, g->seg_start.p, g->seg_end.p, d_rows, ld, col, only);
	}
	return ABW_OK;
}

int abw_coverage(abw_ctx* ctx, const abw_segments* g, const abw_read* reads, uint64_t nreads, int reads_on_device, uint32_t max_snps,
                 int kind, double* d_rows, uint64_t ld, uint32_t col, uint64_t* d_scaf_nbps)
{
	if(!ctx || !g || !d_rows || (!reads && nreads))
		return abw_fail(ctx, ABW_ERR_ARG, "abw_coverage: null argument");
	if(nreads >= (1ull << 32))
		return abw_fail(ctx, ABW_ERR_UNSUPPORTED, "abw_coverage: more than 2^32-1 reads per call; split the sample");
	if(g->nseg >= (1ull << 32))
		return abw_fail(ctx, ABW_ERR_UNSUPPORTED, "abw_coverage: more than 2^32-1 windows");
	ABW_ENTER(ctx);
	DevBuf<abw_read> d_reads;
	const abw_read* rd = reads;
	if(!reads_on_device && nreads) {
		ABW_CUDA(ctx, d_reads.alloc(nreads));
		ABW_CUDA(ctx, cudaMemcpyAsync(d_reads.p, reads, sizeof(abw_read) * nreads, cudaMemcpyHostToDevice, ctx->stream));
		rd = d_reads.p;
	}
	// ABW_COVERAGE_SORT_ALL=1: every read through the sort also for three-decimal output (the first formulation; used by the tests to compare the two)
	static const bool sort_all = [] { const char* e = getenv("ABW_COVERAGE_SORT_ALL"); return e && *e && *e != '0'; }();
	bool too_many_windows = false;
	if(kind == ABW_FEAT_TRUNC3 && !sort_all && nreads && g->nseg) {
		DevBuf<unsigned long long> sum_ov;
		DevBuf<uint32_t> flag;                             // one bit per window
		DevBuf<uint4> hit_g0;
		DevBuf<uchar4> hit_cnt;
		DevBuf<uint32_t> status;
		const unsigned int ntiles = abw_div_up(nreads, COV_TILE);
		ABW_CUDA(ctx, sum_ov.alloc(g->nseg));
		ABW_CUDA(ctx, flag.alloc((g->nseg + 31) / 32));
		ABW_CUDA(ctx, hit_g0.alloc((size_t)ntiles * COV_THREADS));
		ABW_CUDA(ctx, hit_cnt.alloc((size_t)ntiles * COV_THREADS));
		ABW_CUDA(ctx, status.alloc(1));
		ABW_CUDA(ctx, cudaMemsetAsync(sum_ov.p, 0, sizeof(unsigned long long) * g->nseg, ctx->stream));
		ABW_CUDA(ctx, cudaMemsetAsync(status.p, 0, sizeof(uint32_t), ctx->stream));
		ABW_LAUNCH(ctx, k_cov_sum, ntiles, COV_THREADS, 0, rd, nreads, max_snps, g->nscaf, g->scaf_info.p, g->seg_end.p, sum_ov.p, hit_g0.p, hit_cnt.p,
		           (unsigned long long*)d_scaf_nbps, status.p);
		ABW_LAUNCH(ctx, k_cov_quotient, abw_div_up(g->nseg, 256), 256, 0, sum_ov.p, g->nseg, g->seg_start.p, g->seg_end.p, d_rows, ld, col, flag.p);
		ABW_CHECK(coverage_sorted(ctx, g, rd, nreads, max_snps, kind, d_rows, ld, col, nullptr, flag.p, hit_g0.p, hit_cnt.p, status.p, &too_many_windows));
		if(too_many_windows)                                   // a read over more than 255 windows: everything again through the sort (the read bases are counted already)
			ABW_CHECK(coverage_sorted(ctx, g, rd, nreads, max_snps, kind, d_rows, ld, col, nullptr, nullptr, nullptr, nullptr, nullptr, &too_many_windows));
	}
	else
		ABW_CHECK(coverage_sorted(ctx, g, rd, nreads, max_snps, kind, d_rows, ld, col, d_scaf_nbps, nullptr, nullptr, nullptr, nullptr, &too_many_windows));
	// results stay on the device and every later entry point works in the order of the context stream: only a host buffer of reads has to be
	// released by the time the call returns
	if(!reads_on_device)
		ABW_CUDA(ctx, abw_sync(ctx));
	return ABW_OK;
}

}  // extern "C"
